"""Two ICP workloads back to back (for an ncu comparison): predicted transform, then identity transform."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic, _lib
from delora_b200.pipeline import ScanPairPipeline

B, W, H = 8, 2048, 64
cfg = synthetic.fov_config(h=H, w=W)
hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
pairs = [synthetic.make_pair(i, w_raw=2048) for i in range(B)]
nmax = max(max(p[0].shape[1], p[1].shape[1]) for p in pairs)
pipe = ScanPairPipeline(B, nmax, H, W, hf, vf)
pipe.load([p[0] for p in pairs], [p[1] for p in pairs], torch.stack([p[3] for p in pairs]))
pipe.step()
torch.cuda.synchronize()
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
hw = H * W
def icp():
    _lib.check(L.delora_icp_dense_fwd_bwd(pipe.pts_grid.data_ptr() + B * hw * 16, pipe.nrm_grid.data_ptr() + B * hw * 16,
                                          pipe.transform.data_ptr(), pipe.pts_grid.data_ptr(), pipe.nrm_grid.data_ptr(), B, H, W,
                                          hf[0], hf[1], vf[0], vf[1], 1.0, 6, pipe.losses.data_ptr(), pipe.grad_T.data_ptr(),
                                          pipe.icp_scratch.data_ptr(), st), "icp")
for _ in range(3):
    icp()
torch.cuda.synchronize()
pipe.transform.copy_(torch.eye(4, device="cuda")[:3, :].reshape(1, 12).repeat(B, 1))
for _ in range(3):
    icp()
torch.cuda.synchronize()
print("done")

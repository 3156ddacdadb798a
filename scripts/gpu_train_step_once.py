"""A few full synthetic training steps (tcgen05 trunk) for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic
from delora_b200.train_step import SyntheticTrainStep
B, W = 16, 2048
cfg = synthetic.fov_config(h=64, w=W, device="cuda")
pairs = [synthetic.make_pair(i, w_raw=2048) for i in range(4)]
n_max = max(max(p[0].shape[1], p[1].shape[1]) for p in pairs)
pts = torch.zeros((2 * B, 3, n_max)); cnt = torch.zeros((2 * B,), dtype=torch.int32)
for i in range(B):
    s1, s2, _, _ = pairs[i % 4]
    pts[i, :, :s1.shape[1]] = s1; pts[B + i, :, :s2.shape[1]] = s2
    cnt[i], cnt[B + i] = s1.shape[1], s2.shape[1]
torch.manual_seed(0)
ts = SyntheticTrainStep(cfg, B, n_max, use_tensor_cores=True)
ts.load(pts.cuda(), cnt.cuda())
for _ in range(3):
    ts.step()
torch.cuda.synchronize()
print("done")

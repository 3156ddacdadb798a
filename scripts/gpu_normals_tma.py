"""Time the two staging variants of normals_7x11_kernel (coalesced loads vs one TMA box + repack from shared memory) on
the bench shape (16 images of 64 x 2048), L2 flushed between launches, CUDA events; results must be bit-identical."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import ops, synthetic

L = ops._lib.lib()
H, W, B = 64, 2048, 16
cfg = synthetic.fov_config(h=H, w=W, device="cuda")
pairs = [synthetic.make_pair(i, w_raw=2048) for i in range(4)]
n_max = max(max(p[0].shape[1], p[1].shape[1]) for p in pairs)
pts = torch.zeros((B, 3, n_max)); cnt = torch.zeros((B,), dtype=torch.int32)
for i in range(B):
    s = pairs[i % 4][i // 4 % 2]
    pts[i, :, :s.shape[1]] = s; cnt[i] = s.shape[1]
image, _ = ops.project(pts.cuda(), cnt.cuda(), H, W, cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for rep in range(2):
    for mode in (0, 1):
        L.delora_normals_select_staging(mode)
        ts = []
        for i in range(23):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = ops.normals(image, grids=True)
            e1.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        out[mode] = res
        print(f"rep {rep} staging {'TMA box + smem repack' if mode else 'coalesced loads      '}: median {ts[len(ts)//2]:.1f} us, min {ts[0]:.1f} us", flush=True)
L.delora_normals_select_staging(0)
print("bit-identical:", all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(out[0], out[1])))

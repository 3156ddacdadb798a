"""Time the full synthetic training step (config #3 shape) with and without the tcgen05 trunk."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic
from delora_b200.train_step import SyntheticTrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W = 2048
cfg = synthetic.fov_config(h=64, w=W, device="cuda")
pairs = [synthetic.make_pair(i, w_raw=2048) for i in range(4)]
n_max = max(max(p[0].shape[1], p[1].shape[1]) for p in pairs)
pts = torch.zeros((2 * B, 3, n_max)); cnt = torch.zeros((2 * B,), dtype=torch.int32)
for i in range(B):
    s1, s2, _, _ = pairs[i % 4]
    pts[i, :, :s1.shape[1]] = s1; pts[B + i, :, :s2.shape[1]] = s2
    cnt[i], cnt[B + i] = s1.shape[1], s2.shape[1]
for use_tc in (True, False):
    torch.manual_seed(0)
    ts = SyntheticTrainStep(cfg, B, n_max, use_tensor_cores=use_tc)
    ts.load(pts.cuda(), cnt.cuda())
    for _ in range(3): loss, parts = ts.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K): loss, parts = ts.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"use_tensor_cores={use_tc}: {ms:.2f} ms/step -> {B/(ms*1e-3):.1f} pairs/s ; loss {loss.item():.6f} pairs/sample {parts[0,3].item():.0f}", flush=True)

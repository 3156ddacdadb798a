"""Per-kernel device time of the training step in its real (warm-cache, back-to-back) setting, from torch's profiler
(CUPTI activity records, no serialisation, no cache flushes -- unlike the ncu launch lists under profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
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
ts = SyntheticTrainStep(cfg, B, n_max)
ts.load(pts.cuda(), cnt.cuda())
for _ in range(3): ts.step()
torch.cuda.synchronize()
K = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(K): ts.step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"sum of kernel time per step {tot / K:.1f} us ({len(rows)} distinct kernels)")
print("| kernel | launches/step | us/step | share |\n|---|---|---|---|")
for k, c, t in rows[:45]:
    print(f"| `{k[:90]}` | {c / K:.1f} | {t / K:.1f} | {100 * t / tot:.1f} % |")

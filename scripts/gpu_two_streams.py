"""Config #2 step (8 pairs, 64 x 2048) as ONE batched pipeline vs two half-batches on two streams (the latency-bound
projection / block search of one half next to the FP32-bound normals of the other)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic
from delora_b200.pipeline import ScanPairPipeline

H, W = 64, 2048
cfg = synthetic.fov_config(h=H, w=W)
hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
pairs = [synthetic.make_pair(i, w_raw=2048) for i in range(8)]
n_max = max(max(p[0].shape[1], p[1].shape[1]) for p in pairs)


def build(idx):
    p = ScanPairPipeline(len(idx), n_max, H, W, hf, vf, device="cuda")
    p.load([pairs[i][0] for i in idx], [pairs[i][1] for i in idx], torch.stack([pairs[i][3] for i in idx]))
    return p


one = build(list(range(8)))
halves = [build([0, 1, 2, 3]), build([4, 5, 6, 7])]
quarters = [build([0, 1]), build([2, 3]), build([4, 5]), build([6, 7])]
streams = [torch.cuda.Stream() for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(parts):
    if len(parts) == 1:
        parts[0].step()
        return
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    for p, s in zip(parts, streams):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            p.step()
        e = torch.cuda.Event(); e.record(s)
        cur.wait_event(e)


for name, parts in (("1 x 8 pairs", [one]), ("2 x 4 pairs", halves), ("4 x 2 pairs", quarters), ("1 x 8 pairs", [one]), ("2 x 4 pairs", halves)):
    for _ in range(5): run(parts)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(parts); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name}: median {ts[15]:.1f} us, min {ts[0]:.1f} us per 8 pairs", flush=True)
l1 = torch.cat([p.losses for p in halves]); l0 = one.losses
print("losses identical:", torch.equal(l0, l1))

"""torchrun --nproc-per-node 1 scripts/gpu_ar_interference.py
What the co-resident all-reduce kernel costs the training step on ONE GPU (a one-rank group: no coupling between
ranks, all loads local): step time without it, with everything but the kernel, and with several CTA shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from delora_b200 import synthetic
from delora_b200.train_step import SyntheticTrainStep

B, W = 16, 2048
cfg = synthetic.fov_config(h=64, w=W, device=dev)
pairs = [synthetic.make_pair(i, w_raw=2048) for i in range(4)]
n_max = 131072
pts = torch.zeros((2 * B, 3, n_max)); cnt = torch.zeros((2 * B,), dtype=torch.int32)
for i in range(B):
    s1, s2, _, _ = pairs[i % 4]
    n1, n2 = min(s1.shape[1], n_max), min(s2.shape[1], n_max)
    pts[i, :, :n1] = s1[:, :n1]; pts[B + i, :, :n2] = s2[:, :n2]
    cnt[i], cnt[B + i] = n1, n2
pts, cnt = pts.to(dev), cnt.to(dev)
CONFIGS = [("none", None), ("noop", (16, 128, True)), ("16x128", (16, 128, False)), ("148x32", (148, 32, False)),
           ("148x64", (148, 64, False)), ("4x128", (4, 128, False)), ("74x64", (74, 64, False))]
for rep in range(2):
    for name, c in CONFIGS:
        os.environ["DELORA_AR_SELF"] = "1" if c else "0"
        if c:
            os.environ["DELORA_AR_CTAS"], os.environ["DELORA_AR_THREADS"] = str(c[0]), str(c[1])
            os.environ["DELORA_AR_NOOP"] = "1" if c[2] else "0"
        torch.manual_seed(0)
        ts = SyntheticTrainStep(cfg, B, n_max, grad_sync="bucketed")
        ts.load(pts, cnt)
        for _ in range(3): ts.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        e0.record()
        for _ in range(K): ts.step()
        e1.record(); torch.cuda.synchronize()
        print(f"rep {rep} {name:8s} peer={getattr(ts.sync, 'peer', None) is not None}: {e0.elapsed_time(e1)/K:.3f} ms/step", flush=True)
        del ts
        torch.cuda.empty_cache()
dist.destroy_process_group()

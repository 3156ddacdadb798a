"""Under torchrun: time the full training step (B = 16 per GPU, 64x2048) with the bucketed all-reduce, without any
collective and with the blocking flat all-reduce.  NCCL knobs come from the environment (e.g. NCCL_MAX_CTAS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.distributed.init_process_group("nccl", device_id=dev)
from delora_b200 import synthetic  # noqa: E402
from delora_b200.train_step import SyntheticTrainStep  # noqa: E402

B = 16
cfg = synthetic.fov_config(h=64, w=2048, device=dev)
pairs = [synthetic.make_pair(rank * 4 + i, w_raw=2048) for i in range(4)]
n_max = max(max(p[0].shape[1], p[1].shape[1]) for p in pairs)
pts = torch.zeros((2 * B, 3, n_max)); cnt = torch.zeros((2 * B,), dtype=torch.int32)
for i in range(B):
    s1, s2, _, _ = pairs[i % 4]
    pts[i, :, :s1.shape[1]] = s1; pts[B + i, :, :s2.shape[1]] = s2
    cnt[i], cnt[B + i] = s1.shape[1], s2.shape[1]
out = {}
for mode in ("bucketed", "none", "flat", "bucketed"):
    torch.manual_seed(0)
    ts = SyntheticTrainStep(cfg, B, n_max, grad_sync=mode)
    ts.load(pts.to(dev), cnt.to(dev))
    for _ in range(4):
        ts.step()
    torch.cuda.synchronize(); torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 15
    for _ in range(K):
        ts.step()
    e1.record(); torch.cuda.synchronize(); torch.distributed.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    out.setdefault(mode, []).append(round(float(t[0]), 3))
    del ts
if rank == 0:
    print("world", world, "NCCL_MAX_CTAS", os.environ.get("NCCL_MAX_CTAS"), "NCCL_ALGO", os.environ.get("NCCL_ALGO"), out, flush=True)
torch.distributed.destroy_process_group()

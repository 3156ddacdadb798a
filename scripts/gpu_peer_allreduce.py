"""torchrun --nproc-per-node N scripts/gpu_peer_allreduce.py [--train]
Peer-memory (NVLink / NVSwitch multicast) gradient all-reduce on a multi-GPU node: rendezvous, exactness against NCCL,
time of the 47.5 MB reduction alone, and (--train) the DDP training step with each transport."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
from delora_b200.parallel_grad import _PeerTransport

def say(*a):
    if rank == 0:
        print(*a, flush=True)

total = 11_880_000 // 128 * 128
t0 = time.time()
pt = _PeerTransport(total, dev, None, n_ctas=int(os.environ.get("AR_CTAS", "16")))
say(f"rendezvous ok in {time.time()-t0:.2f}s: world {world}, multicast_ptr {'yes' if pt.multicast else 'NO (peer loads)'}")
for use_mc in ([True, False] if pt.multicast else [False]):
    mc_saved = pt.multicast
    if not use_mc:
        pt.multicast = 0
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(total, generator=g, device=dev)
    ref = src.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    pt.flat.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    for (s, e) in [(0, 1280), (1280, 3_000_064), (3_000_064, total)]:
        pt.launch(s, e)
    pt.wait()
    torch.cuda.synchronize()
    pt.check()
    err = (pt.flat - ref).abs().max().item()
    gathered = [torch.empty_like(pt.flat) for _ in range(world)] if rank == 0 else None
    dist.gather(pt.flat, gathered, dst=0)
    same = all(torch.equal(gathered[0], t) for t in gathered) if rank == 0 else True
    say(f"[{'multicast' if use_mc else 'peer'}] max |peer - nccl| = {err:.3e} (|x| ~ 1); identical on all ranks: {same}")
    # time of one whole-buffer reduction, nothing else running
    for ctas in (4, 16, 32):
        pt.n_ctas = ctas
        for _ in range(3):
            pt.launch(0, total)
        pt.wait(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pt.stream)
        for _ in range(10):
            pt.launch(0, total)
        e1.record(pt.stream)
        pt.wait(); torch.cuda.synchronize()
        say(f"[{'multicast' if use_mc else 'peer'}] {ctas:2d} CTAs: {e0.elapsed_time(e1)/10*1e3:.0f} us per 47.5 MB all-reduce")
    # small bucket latency (the exposed one)
    pt.n_ctas = 16
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pt.stream)
    for _ in range(20):
        pt.launch(0, 157_696)
    e1.record(pt.stream)
    pt.wait(); torch.cuda.synchronize()
    say(f"[{'multicast' if use_mc else 'peer'}] 0.63 MB bucket: {e0.elapsed_time(e1)/20*1e3:.1f} us")
    pt.multicast = mc_saved
pt.check()

if "--train" in sys.argv:
    from delora_b200 import synthetic
    from delora_b200.train_step import SyntheticTrainStep
    B, W = 16, 2048
    cfg = synthetic.fov_config(h=64, w=W, device=dev)
    pairs = [synthetic.make_pair(i + 4 * rank, w_raw=2048) for i in range(4)]
    n_max = 131072
    pts = torch.zeros((2 * B, 3, n_max)); cnt = torch.zeros((2 * B,), dtype=torch.int32)
    for i in range(B):
        s1, s2, _, _ = pairs[i % 4]
        n1, n2 = min(s1.shape[1], n_max), min(s2.shape[1], n_max)
        pts[i, :, :n1] = s1[:, :n1]; pts[B + i, :, :n2] = s2[:, :n2]
        cnt[i], cnt[B + i] = n1, n2
    res = {}
    for mode in ("bucketed", "bucketed-nccl", "none", "bucketed", "bucketed-nccl", "none"):
        torch.manual_seed(0)
        ts = SyntheticTrainStep(cfg, B, n_max, grad_sync=mode)
        ts.load(pts.to(dev), cnt.to(dev))
        for _ in range(3): loss, _ = ts.step()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        e0.record()
        for _ in range(K): loss, _ = ts.step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w0 = next(ts.model.parameters()).detach().flatten()[:8].clone()
        wl = [torch.empty_like(w0) for _ in range(world)]
        dist.all_gather(wl, w0)
        if getattr(ts.sync, "peer", None) is not None:
            ts.sync.peer.check()
        say(f"train step [{mode:13s}] transport {getattr(ts.sync, 'transport', '-'):15s}: {t.item():.3f} ms/step, loss {loss.item():.6f}, "
            f"weights identical across ranks: {all(torch.equal(wl[0], x) for x in wl)} {getattr(ts.sync, 'transport_note', '')}")
        if getattr(ts.sync, "peer", None) is not None and mode == "bucketed":
            # where inside the step each bucket's reduction ran: offsets relative to the step's first kernel
            ts.sync.peer.trace = []
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            s0.record(); ts.step(); s1.record(); torch.cuda.synchronize()
            for (a, b, ready, t0, t1) in ts.sync.peer.trace:
                say(f"    bucket [{a:9d}, {b:9d}) {4*(b-a)/1e6:6.2f} MB: ready at {s0.elapsed_time(ready):.3f} ms, kernel "
                    f"{s0.elapsed_time(t0):.3f} -> {s0.elapsed_time(t1):.3f} ms (step ends {s0.elapsed_time(s1):.3f})")
            ts.sync.peer.trace = None
        del ts
        torch.cuda.empty_cache()
dist.destroy_process_group()

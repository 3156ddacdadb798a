"""Data-parallel gradient synchronisation (the reference itself is single-process,
`src/deploy/deployer.py:329-342` divides the summed loss by the batch size -- with equal per-rank batches the mean
over ranks of the per-rank mean reproduces it).  The scan-pair kernels are rank-local; the gradient all-reduce is the
only collective of the training step.  One process per GPU, launched with torchrun.

`BucketedGradAllReduce` (the training path): every gradient is a view of ONE persistent flat fp32 buffer (11.88 M
elements = 47.5 MB for the default model) laid out in BACKWARD order as a few buckets (heads + fc | layer4 | layer3 |
layer2 | layer1 + stem).  The tensor-core encoder writes its weight gradients straight into the flat slices (no copies)
and announces each group from inside its backward, so layer4 (33.6 MB) is reduced while the backward of layers 3 -> 1
still runs; only the last small bucket is exposed.

Transport of a bucket (`transport=`):
  "peer"  the flat buffer is symmetric memory (mapped at every peer of the node, NVSwitch multicast where available)
          and `delora_grad_allreduce_f32` (csrc/grad_allreduce.cu) reduces it in place over NVLink: a 128-thread,
          shared-memory-free kernel whose CTAs are resident NEXT TO the persistent tcgen05 convolution CTAs.  NCCL's
          CTAs cannot share an SM with those, which left 0.25 ms of the collective exposed per step at 8 GPUs.
  "nccl"  `dist.all_reduce(AVG, async_op=True)` per bucket (any backend; gloo on CPU uses SUM + divide).
  "auto"  "peer" on CUDA when the symmetric-memory rendezvous succeeds on EVERY rank, else "nccl".
`FlatGradAllReduce` (one blocking all-reduce after backward) is kept for the CPU / gloo tests and as the measured
baseline of the overlap."""
import ctypes
import os

import torch
import torch.distributed as dist


def _world(process_group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(process_group), dist.get_rank(process_group)
    return 1, 0


class FlatGradAllReduce:
    def __init__(self, module, process_group=None):
        self.world, self.rank = _world(process_group)
        self.group = process_group
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        if self.world > 1:      # identical start on every rank
            for p in self.params:
                dist.broadcast(p.data, src=0, group=self.group)

    def all_reduce(self):
        """Average the gradients over the ranks in place (sum / world, equal per-rank batch)."""
        if self.world == 1:
            return
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(self.world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + n].view_as(p))
            off += n

    finish = all_reduce


class _PeerTransport:
    """Symmetric flat buffer + flag words + the launch of csrc/grad_allreduce.cu.  torch supplies the plumbing only
    (CUDA VMM allocation, handle exchange over the process group's store, the multicast binding)."""

    BUCKET_ALIGN = 128          # floats: buckets start on 512-byte boundaries (16-byte vectors, whole sectors)

    def __init__(self, total, device, group, n_ctas=16, n_threads=128):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        L = _lib.lib()
        pg = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
        self.flat = symm_mem.empty(total, dtype=torch.float32, device=device)
        self.flat.zero_()
        self.flags = symm_mem.empty(L.delora_grad_allreduce_flag_words(), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        hb = symm_mem.rendezvous(self.flat, pg.group_name)
        hf = symm_mem.rendezvous(self.flags, pg.group_name)
        self._handles = (hb, hf)
        u64 = ctypes.c_uint64 * self.world
        self.peer_bufs = u64(*[int(x) for x in hb.buffer_ptrs])
        self.peer_flags = u64(*[int(x) for x in hf.buffer_ptrs])
        self.multicast = int(hb.multicast_ptr or 0)
        self.n_ctas = int(os.environ.get("DELORA_AR_CTAS", n_ctas))
        self.n_threads = int(os.environ.get("DELORA_AR_THREADS", n_threads))
        self.noop = os.environ.get("DELORA_AR_NOOP") == "1"        # experiments: everything but the kernel
        self.seq = 0
        # the collective's own stream: it waits for the bucket's producers, then runs beside the rest of the backward
        self.stream = torch.cuda.Stream(device=device, priority=-1)
        self._done = None           # (the rendezvous above is the barrier "every rank has zeroed its flags")
        self.trace = None           # set to [] to collect (start, end, t_start_event, t_end_event) per launch

    def launch(self, start, end):
        cur = torch.cuda.current_stream(self.flat.device)
        ready = torch.cuda.Event(enable_timing=self.trace is not None)
        ready.record(cur)
        self.stream.wait_event(ready)
        self.seq += 1
        L = self._lib.lib()
        if self.trace is not None:
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record(self.stream)
        if not self.noop:
            self._lib.check(L.delora_grad_allreduce_f32(self.peer_bufs, self.peer_flags, self.multicast, self.rank,
                                                        self.world, start, end - start, 1.0 / self.world,
                                                        self.seq & 0x7fffffff, self.n_ctas, self.n_threads,
                                                        self.status.data_ptr(), self.stream.cuda_stream),
                            "delora_grad_allreduce_f32")
        self._done = torch.cuda.Event(enable_timing=self.trace is not None)
        self._done.record(self.stream)
        if self.trace is not None:
            self.trace.append((start, end, ready, t0, self._done))

    def wait(self):
        if self._done is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(self._done)
            self._done = None

    def check(self):
        """Host-synchronising: raises if a peer failed to arrive in some collective since the last check."""
        code = int(self.status.item())
        if code:
            self.status.zero_()
            raise RuntimeError(f"delora_grad_allreduce_f32: rank {code - 1} did not arrive within the time limit")


class BucketedGradAllReduce:
    """Flat gradient buffer + bucketed, overlapped all-reduce.  Usage per step (ONE backward per step: the flat slices
    are overwritten, not accumulated into):
        optimizer.zero_grad(set_to_none=True); loss.backward(); sync.finish(); optimizer.step()
    `finish()` waits for the outstanding collectives (stream-ordered, no host sync) and makes every `p.grad` the
    view of the flat buffer that holds the averaged gradient.  With WORLD_SIZE == 1 the buffer and the direct
    gradient writes are still used (no copies), only the collectives are skipped."""

    def __init__(self, module, encoder=None, process_group=None, enabled=True, transport="auto"):
        self.world, self.rank = _world(process_group)
        self.group = process_group
        self.enabled = bool(enabled)
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError(f"transport must be auto, peer or nccl, got {transport!r}")
        if transport == "peer" and not torch.cuda.is_available():
            raise ValueError("transport='peer' needs CUDA devices on one NVLink node")
        params = [p for p in module.parameters() if p.requires_grad]
        trunk = list(encoder.trunk_parameters()) if encoder is not None else []
        trunk_ids = {id(p) for p in trunk}
        others = [p for p in params if id(p) not in trunk_ids]
        # buckets in the order their gradients become ready during backward
        buckets = [others]
        if trunk:
            groups = encoder.trunk_parameter_groups()          # list of lists of trunk indices, backward order
            for g in groups:
                buckets.append([trunk[i] for i in g])
            self._trunk_bucket = {}
            for bi, g in enumerate(groups):
                for i in g:
                    self._trunk_bucket[i] = bi + 1
        self.buckets = [b for b in buckets]
        device = params[0].device
        align = _PeerTransport.BUCKET_ALIGN
        total = sum((sum(p.numel() for p in b) + align - 1) // align * align for b in self.buckets)
        self.peer, self.transport, self.transport_note = None, "nccl", ""
        # DELORA_AR_SELF=1 (experiments): run the peer kernel even in a one-rank group, to separate its own cost
        # (launches, co-residency with the convolutions) from the coupling between ranks
        solo = self.world == 1 and os.environ.get("DELORA_AR_SELF") == "1" and dist.is_initialized()
        if (self.world > 1 or solo) and self.enabled and transport != "nccl" and device.type == "cuda":
            ok = 1
            try:
                self.peer = _PeerTransport(total, device, process_group)
            except Exception as e:                       # no VMM / fabric handle support, rendezvous refused, ...
                ok, self.peer, self.transport_note = 0, None, f"{type(e).__name__}: {e}"[:300]
            agree = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
            if int(agree.item()) == 0:
                if transport == "peer":
                    raise RuntimeError("peer-memory gradient transport unavailable on some rank: " + self.transport_note)
                self.peer = None
            else:
                self.transport = "peer-multicast" if self.peer.multicast else "peer"
        self.flat = self.peer.flat if self.peer is not None else torch.zeros(total, dtype=torch.float32, device=device)
        self.views, self.ranges, self._bucket_of = {}, [], {}
        off = 0
        for bi, b in enumerate(self.buckets):
            start = off
            for p in b:
                n = p.numel()
                self.views[id(p)] = self.flat[off:off + n].view_as(p)
                self._bucket_of[id(p)] = bi
                off += n
            off = (off + align - 1) // align * align       # padding stays zero: reduced along, never read
            self.ranges.append((start, off))
        self.params = params
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._launched = [False] * len(self.buckets)
        if self.world > 1:
            for p in params:
                dist.broadcast(p.data, src=0, group=self.group)
        # gradients produced by autograd itself (heads, fc, or everything when the tensor-core trunk is not used):
        # copied into their flat slice as soon as they are accumulated
        for p in params:
            p.register_post_accumulate_grad_hook(self._on_autograd_grad)
        self._trunk_params = trunk
        self.encoder = encoder
        if encoder is not None:
            encoder.grad_views = [self.views[id(p)] for p in trunk]
            encoder.grad_ready = self._on_trunk_grads
        self._reset()

    # ------------------------------------------------------------------ per-step state
    def _reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._ready = set()

    def _mark(self, p):
        if id(p) in self._ready:
            return
        self._ready.add(id(p))
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        if self._launched[bi]:
            return
        self._launched[bi] = True
        if (self.world == 1 and self.peer is None) or not self.enabled:
            return
        s, e = self.ranges[bi]
        if e == s:
            return
        if self.peer is not None:
            self.peer.launch(s, e)
            return
        chunk = self.flat[s:e]
        if dist.get_backend(self.group) == "nccl":
            self._works.append((dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
        else:
            self._works.append((dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True), chunk))

    def _on_autograd_grad(self, p):
        if id(p) in self._ready:       # written in place by the encoder and possibly already being reduced: whatever
            return                     # autograd stored in p.grad (the view itself or a copy of it) is replaced in finish()
        view = self.views[id(p)]
        if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)
        self._mark(p)

    def _on_trunk_grads(self, trunk_indices):
        """Called by the encoder's backward right after the weight gradients of `trunk_indices` were enqueued
        (written into `encoder.grad_views` on the current stream)."""
        for i in trunk_indices:
            self._mark(self._trunk_params[i])

    def finish(self):
        """Wait (stream-ordered) for the collectives and point every p.grad at its averaged flat view."""
        for bi in range(len(self.buckets)):        # parameters that received no gradient this step count as zeros
            if not self._launched[bi]:
                for p in self.buckets[bi]:
                    if id(p) not in self._ready:
                        self.views[id(p)].zero_()
                self._launch(bi)
        for work, chunk in self._works:
            work.wait()
            if chunk is not None:
                chunk.div_(self.world)
        if self.peer is not None:
            self.peer.wait()
        for p in self.params:
            view = self.views[id(p)]
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
        self._reset()

    all_reduce = finish


def make_grad_sync(model, mode="bucketed", process_group=None):
    """The gradient synchroniser of a training loop: "bucketed" (flat buffer, overlapped per-bucket all-reduce over peer
    memory when available, else NCCL; also used on one GPU with the tensor-core encoder, where it only removes gradient
    copies), "bucketed-nccl" (the same with NCCL forced: the measured baseline of the peer-memory kernel), "flat" (one
    blocking all-reduce after backward), "none" (no collective: the measured reference for the exposed all-reduce time)."""
    encoder = model._tensor_core_path() if hasattr(model, "_tensor_core_path") else None
    world, _ = _world(process_group)
    if mode == "flat" or (encoder is None and world == 1):
        return FlatGradAllReduce(model, process_group)
    return BucketedGradAllReduce(model, encoder=encoder, process_group=process_group, enabled=(mode != "none"),
                                 transport="nccl" if mode == "bucketed-nccl" else "auto")


def shard_pairs(num_pairs, rank, world):
    """Contiguous shard of pair indices for `rank` (the bench / loaders shard scan pairs, not tensors)."""
    per = (num_pairs + world - 1) // world
    return list(range(min(rank * per, num_pairs), min((rank + 1) * per, num_pairs)))

"""Data-parallel gradient synchronisation: ONE flat all-reduce of every parameter gradient per
step (11.88 M fp32 = 47.5 MB for the default model), NCCL over NVLink on GPUs, gloo in the CPU
tests.  The scan-pair kernels are rank-local; this is the only collective of the training step
(SURVEY.md §8(e)).  One process per GPU, launched with torchrun."""

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    def __init__(self, module, process_group=None):
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
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


def shard_pairs(num_pairs, rank, world):
    """Contiguous shard of pair indices for `rank` (the bench / loaders shard scan pairs, not tensors)."""
    per = (num_pairs + world - 1) // world
    return list(range(min(rank * per, num_pairs), min((rank + 1) * per, num_pairs)))

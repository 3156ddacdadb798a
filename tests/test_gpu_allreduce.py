"""-m gpu: the peer-memory gradient all-reduce kernel (csrc/grad_allreduce.cu) on ONE device: the "ranks" are kernels
on separate streams of the same GPU that see each other's buffers and flag words through plain device pointers (the
unicast path; the NVSwitch multicast path needs a multi-GPU node, scripts/gpu_peer_allreduce.py)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(world, total, ranges, n_ctas, rounds=1, seed=0):
    from delora_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device=DEV).manual_seed(seed)
    bufs = [torch.randn(total, generator=g, device=DEV) for _ in range(world)]
    flags = [torch.zeros(L.delora_grad_allreduce_flag_words(), dtype=torch.int32, device=DEV) for _ in range(world)]
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    u64 = ctypes.c_uint64 * world
    pb, pf = u64(*[b.data_ptr() for b in bufs]), u64(*[f.data_ptr() for f in flags])
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    seq = 0
    want = [b.clone() for b in bufs]
    for _ in range(rounds):
        for (s, e) in ranges:
            seq += 1
            acc = torch.zeros(e - s, device=DEV)
            for q in range(world):                       # the kernel adds the ranks in this order
                acc = acc + want[q][s:e]
            acc = acc * (1.0 / world)
            for q in range(world):
                want[q][s:e] = acc
            for r in range(world):
                _lib.check(L.delora_grad_allreduce_f32(pb, pf, 0, r, world, s, e - s, 1.0 / world, seq, n_ctas,
                                                       128 if n_ctas <= 8 else 32, status.data_ptr(), streams[r].cuda_stream), "grad_allreduce")
    torch.cuda.synchronize()
    assert int(status.item()) == 0, "a rank timed out"
    return bufs, want


@pytest.mark.parametrize("world,n_ctas", [(1, 4), (2, 8), (4, 16), (8, 4), (2, 148)])
def test_peer_allreduce_average_is_exact_and_identical_on_all_ranks(world, n_ctas, cuda_lib):
    total = 1 << 20
    ranges = [(0, 4096), (4096, 4096 + 300 * 128), (524288, 1 << 20), (262144, 262144 + 4)]
    bufs, want = _run(world, total, ranges, n_ctas, rounds=2, seed=world)
    for q in range(world):
        assert torch.equal(bufs[q], want[q])
        assert torch.equal(bufs[q][:4096 + 300 * 128], bufs[0][:4096 + 300 * 128])


def test_peer_allreduce_rejects_bad_arguments(cuda_lib):
    from delora_b200 import _lib
    L = _lib.lib()
    buf = torch.zeros(1024, device=DEV)
    flags = torch.zeros(L.delora_grad_allreduce_flag_words(), dtype=torch.int32, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    u64 = ctypes.c_uint64 * 1
    pb, pf = u64(buf.data_ptr()), u64(flags.data_ptr())
    assert L.delora_grad_allreduce_f32(pb, pf, 0, 0, 1, 2, 64, 1.0, 1, 4, 128, status.data_ptr(), None) != 0     # offset % 4
    assert L.delora_grad_allreduce_f32(pb, pf, 0, 0, 1, 0, 64, 1.0, 1, 0, 128, status.data_ptr(), None) != 0     # n_ctas
    assert L.delora_grad_allreduce_f32(pb, pf, 0, 1, 1, 0, 64, 1.0, 1, 4, 128, status.data_ptr(), None) != 0     # rank >= world
    assert L.delora_grad_allreduce_f32(pb, pf, 0, 0, 1, 0, 64, 1.0, 1, 4, 48, status.data_ptr(), None) != 0      # n_threads
    assert b"grad_allreduce" in L.delora_last_error()

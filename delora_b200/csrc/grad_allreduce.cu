// Gradient all-reduce over NVLink / NVSwitch peer memory for the data-parallel training step (sm_100a).
//
// the reference is single-process (src/deploy/trainer.py:23-24 steps one optimizer on one
// device); the data-parallel step averages the weight gradients of equal per-rank batches, which reproduces
// the reference's `loss / batch_size` (src/deploy/deployer.py:329-342) over the global batch.
//
// Why not NCCL here: the tcgen05 convolution kernels are persistent, one 224/352-thread CTA with ~226 KB of shared
// memory per SM.  An NCCL CTA (640 threads, tens of KB of shared memory) cannot share an SM with them, so a
// collective issued during the backward only runs in the gaps between convolution kernels and then delays the next
// one (measured: 0.25 ms of exposed all-reduce per step at 8 GPUs, profiles/r02_train_step.md).  This kernel is built
// to be CO-RESIDENT with them instead: 128 threads, no shared memory, < 64 registers, so its CTAs are placed next
// to the running convolution CTAs and the reduction proceeds while the backward continues.
//
// Data path (every gradient lives in ONE flat fp32 buffer that is symmetric memory, i.e. mapped at every peer and,
// where the fabric offers it, bound to an NVSwitch multicast object):
//   rank r owns elements [r * per, (r + 1) * per) of the bucket.
//   multicast:  v = multimem.ld_reduce.add.v4.f32 [mc + i]   -- the switch adds the 8 copies in flight
//               multimem.st.v4.f32 [mc + i] = v * scale       -- and writes the average back to all 8
//   peer (no multicast):  v = sum over ranks q (fixed order) of ld [peer_q + i];  st [peer_q + i] = v * scale
// so every element crosses the links once in each direction and all ranks end up with bit-identical averages.
// Ordering between ranks: one flag word per (phase, CTA, source rank) in symmetric memory, written with
// st.release.sys and polled with ld.acquire.sys; CTA c of every rank synchronises only with CTA c of the peers
// (start: all gradients of the bucket are written; end: all slices are stored), monotonically increasing
// sequence numbers, no reset.  A wait that exceeds kArTimeoutNs sets *status and carries on: a lost peer shows
// up as an error on the host, never as a hung GPU.
#include "common.cuh"

namespace delora {

constexpr int kArMaxThreads = 128;
constexpr int kArMaxWorld = 16;
constexpr int kArMaxCtas = 160;
constexpr unsigned long long kArTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct ArParams {
    unsigned long long peer_buf[kArMaxWorld];     // unicast address of the flat buffer at every rank
    unsigned long long peer_flag[kArMaxWorld];    // unicast address of the flag words at every rank
    float* mc;                                    // multicast address of the flat buffer (nullptr: peer path)
    int rank, world;
    long long off, count;                         // bucket = elements [off, off + count) of the flat buffer
    float scale;
    unsigned seq;
    int* status;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Flag word of (phase, cta, source rank) inside one rank's flag array.
__device__ __forceinline__ int ar_flag_index(int phase, int cta, int src) {
    return (phase * kArMaxCtas + cta) * kArMaxWorld + src;
}

// All ranks' CTA `cta` meet: thread q tells rank q "rank `rank` reached `seq`", then waits for rank q's word here.
__device__ __forceinline__ void ar_meet(const ArParams& p, int phase, int cta) {
    const int q = threadIdx.x;
    if (q < p.world) {
        unsigned* theirs = reinterpret_cast<unsigned*>(p.peer_flag[q]) + ar_flag_index(phase, cta, p.rank);
        st_release_sys(theirs, p.seq);
        const unsigned* mine = reinterpret_cast<const unsigned*>(p.peer_flag[p.rank]) + ar_flag_index(phase, cta, q);
        const unsigned long long t0 = global_timer_ns();
        while ((int)(ld_acquire_sys(mine) - p.seq) < 0) {
            if (global_timer_ns() - t0 > kArTimeoutNs) { atomicExch(p.status, 1 + q); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

template <bool MC>
__global__ void __launch_bounds__(kArMaxThreads)
grad_allreduce_kernel(ArParams p) {
    const int cta = blockIdx.x;
    ar_meet(p, 0, cta);                                             // every rank's bucket is complete

    const long long n4 = p.count >> 2;
    const long long per = (n4 + p.world - 1) / p.world;
    const long long lo = (p.off >> 2) + (long long)p.rank * per;
    const long long hi_raw = (p.off >> 2) + ((long long)(p.rank + 1) * per < n4 ? (long long)(p.rank + 1) * per : n4);
    const long long hi = hi_raw > lo ? hi_raw : lo;
    constexpr int U = 4;                                            // independent 16-byte requests in flight per thread
    const int nt = blockDim.x;
    const long long step = (long long)gridDim.x * nt * U;
    for (long long i0 = lo + (long long)cta * nt * U + threadIdx.x; i0 < hi; i0 += step) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + (long long)u * nt;
            if (i < hi) {
                if (MC) {
                    v[u] = multimem_ld_reduce_add(reinterpret_cast<const float4*>(p.mc) + i);
                } else {
                    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int q = 0; q < p.world; ++q) {
                        const float4 t = ld_relaxed_sys(reinterpret_cast<const float4*>(p.peer_buf[q]) + i);
                        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
                    }
                    v[u] = s;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + (long long)u * nt;
            if (i < hi) {
                const float4 o = make_float4(v[u].x * p.scale, v[u].y * p.scale, v[u].z * p.scale, v[u].w * p.scale);
                if (MC) {
                    multimem_st(reinterpret_cast<float4*>(p.mc) + i, o);
                } else {
                    for (int q = 0; q < p.world; ++q) st_relaxed_sys(reinterpret_cast<float4*>(p.peer_buf[q]) + i, o);
                }
            }
        }
    }
    __threadfence_system();                                         // this thread's stores before the CTA's flag
    __syncthreads();
    ar_meet(p, 1, cta);                                             // every rank's slice has landed everywhere
}

}  // namespace delora

using namespace delora;

extern "C" int delora_grad_allreduce_flag_words(void) { return 2 * kArMaxCtas * kArMaxWorld; }

extern "C" int delora_grad_allreduce_f32(const uint64_t* peer_bufs, const uint64_t* peer_flags, uint64_t multicast_ptr,
                                         int rank, int world, long long offset, long long count, float scale,
                                         uint32_t seq, int n_ctas, int n_threads, int32_t* status, void* stream) {
    DELORA_CHECK_ARG(peer_bufs && peer_flags && status, "delora_grad_allreduce_f32: null pointer");
    DELORA_CHECK_ARG(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world,
                     "delora_grad_allreduce_f32: bad rank %d / world %d (max %d)", rank, world, kArMaxWorld);
    DELORA_CHECK_ARG(offset >= 0 && count >= 0 && (offset & 3) == 0 && (count & 3) == 0,
                     "delora_grad_allreduce_f32: offset %lld / count %lld must be multiples of 4 floats", offset, count);
    DELORA_CHECK_ARG(n_ctas >= 1 && n_ctas <= kArMaxCtas, "delora_grad_allreduce_f32: n_ctas %d outside [1, %d]", n_ctas,
                     kArMaxCtas);
    DELORA_CHECK_ARG(n_threads >= 32 && n_threads <= kArMaxThreads && n_threads % 32 == 0,
                     "delora_grad_allreduce_f32: n_threads %d must be 32, 64, 96 or 128", n_threads);
    if (count == 0) return 0;
    ArParams p;
    for (int q = 0; q < kArMaxWorld; ++q) {
        p.peer_buf[q] = q < world ? peer_bufs[q] : 0ull;
        p.peer_flag[q] = q < world ? peer_flags[q] : 0ull;
        DELORA_CHECK_ARG(q >= world || (p.peer_buf[q] && p.peer_flag[q] && (p.peer_buf[q] & 15) == 0),
                         "delora_grad_allreduce_f32: peer %d has a null or unaligned mapping", q);
    }
    DELORA_CHECK_ARG((multicast_ptr & 15) == 0, "delora_grad_allreduce_f32: unaligned multicast address");
    p.mc = reinterpret_cast<float*>(multicast_ptr);
    p.rank = rank; p.world = world; p.off = offset; p.count = count; p.scale = scale; p.seq = seq; p.status = status;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.mc)
        grad_allreduce_kernel<true><<<n_ctas, n_threads, 0, st>>>(p);
    else
        grad_allreduce_kernel<false><<<n_ctas, n_threads, 0, st>>>(p);
    DELORA_CHECK_LAUNCH("grad_allreduce_kernel");
    return 0;
}

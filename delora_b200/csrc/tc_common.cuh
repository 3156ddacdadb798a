// PTX wrappers shared by the tcgen05 convolution kernels (conv_tc.cu, conv_rows.cu, conv_wgrad.cu): mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05.mma / commit / ld, shared-memory matrix descriptors.
//
// Descriptor semantics used here were measured on B200 with scripts/umma_probe.cu (profiles/r02_umma_probe.log):
// with SWIZZLE_128B the tensor core derives the swizzle phase from the ABSOLUTE shared-memory address, so an operand
// may start at any multiple of 128 B inside a TMA-written tile with the descriptor's base_offset field left at 0
// (setting it to (addr >> 7) & 7 gives wrong results).  This is what lets one halo tile serve all filter taps.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include "common.cuh"

namespace delora {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// one lane of a converged warp (elect.sync): the tcgen05 issue pattern -- the WHOLE warp runs the loop so that
// descriptors and barrier addresses stay in uniform registers, only the tcgen05 instructions are predicated
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// dynamic shared memory base rounded up to 1024 B WITHOUT leaving the shared address space (pointer arithmetic on the
// extern array; a round trip through uintptr_t makes every later access a generic LD/ST)
#define DELORA_ALIGNED_SMEM(raw) ((raw) + ((1024u - (smem_u32(raw) & 1023u)) & 1023u))
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- cta_group::2 (CTA pair) variants; mechanics verified by scripts/umma_probe2.cu (profiles/r02_umma_probe2.log):
// both CTAs execute the allocation, each CTA's TMA writes its OWN shared memory but completes its bytes on the LEADER's
// (even rank) mbarrier -- the local barrier address with bit 24 cleared is the leader's copy in shared::cluster space --
// one thread of the leader issues the MMA for both SMs and its commit is multicast to the same barrier in both CTAs.
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {      // arrive on the leader CTA's copy of `bar`
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kLeaderMask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & kLeaderMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & kLeaderMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_out, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor: start address >> 4 | LBO (unused for swizzled
// K-major) = 1 | SBO = 1024 B (8 rows x 128 B) | version 1 | base_offset 0 | layout_type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// MN-major, 128-byte-swizzled descriptor: rows of 64 M/N elements (128 B), K runs across rows; LBO = byte offset
// between 64-element blocks along M/N, SBO = 1024 B between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor, kind::f16: D = F32 (bit 4), A = B = BF16 (bits 7, 10), a_major / b_major (bits 15, 16:
// 0 = K-major, 1 = MN-major), N >> 3 (bits 17..22), M >> 4 (bits 24..28)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.0f);
    if (act == 2) {           // MUFU.TANH: 2^-11 relative error, far below the bf16 output rounding
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v));
        return t;
    }
    return v;
}

// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
inline PFN_cuTensorMapEncodeTiled_v12000 get_tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

// Epilogue of one output pixel's 32 consecutive channels (fp32 accumulators in v[32]): residual add, activation or
// activation derivative of the saved forward output (dgrad modes 3 = tanh', 4 = relu'), bf16 conversion, NHWC store
// and the circular halo copies of the padded layout.  `off` = element offset of the pixel's first channel of this
// group in y / residual / saved; halo_right / halo_left: also store at pixel + Wout / pixel - Wout.
struct EpiloguePrefetch { uint4 r[4]; uint4 s[4]; };
// issue the residual / saved loads of one pixel's 32 channels early (before the TMEM read + transpose)
__device__ __forceinline__ void epilogue_prefetch32(EpiloguePrefetch& pf, const __nv_bfloat16* __restrict__ residual,
                                                    const __nv_bfloat16* __restrict__ saved, size_t off, int act) {
    if (residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(residual + off);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) pf.r[j4] = __ldg(rp + j4);
    }
    if (act >= 3) {
        const uint4* sp = reinterpret_cast<const uint4*>(saved + off);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) pf.s[j4] = __ldg(sp + j4);
    }
}
// same as epilogue_store32 with the residual / saved values already in registers
__device__ __forceinline__ void epilogue_finish32(float (&v)[32], const EpiloguePrefetch& pf, bool has_residual,
                                                  __nv_bfloat16* __restrict__ y, size_t off, int act, bool halo_right,
                                                  bool halo_left, size_t halo_elems) {
    if (has_residual) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pf.r[j4]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                v[j4 * 8 + e * 2] += f.x; v[j4 * 8 + e * 2 + 1] += f.y;
            }
        }
    }
    if (act >= 3) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pf.s[j4]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __bfloat1622float2(h[e]);
                const float d0 = (act == 3) ? fmaf(-a.x, a.x, 1.0f) : (a.x > 0.0f ? 1.0f : 0.0f);
                const float d1 = (act == 3) ? fmaf(-a.y, a.y, 1.0f) : (a.y > 0.0f ? 1.0f : 0.0f);
                v[j4 * 8 + e * 2] *= d0; v[j4 * 8 + e * 2 + 1] *= d1;
            }
        }
    }
    uint4 out[4];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            h[e] = __floats2bfloat162_rn(apply_act(v[j4 * 8 + e * 2], act), apply_act(v[j4 * 8 + e * 2 + 1], act));
        out[j4] = *reinterpret_cast<uint4*>(h);
    }
    uint4* yp = reinterpret_cast<uint4*>(y + off);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) yp[j4] = out[j4];
    if (halo_right) {
        uint4* hp = reinterpret_cast<uint4*>(y + off + halo_elems);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) hp[j4] = out[j4];
    }
    if (halo_left) {
        uint4* hp = reinterpret_cast<uint4*>(y + off - halo_elems);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) hp[j4] = out[j4];
    }
}

__device__ __forceinline__ void epilogue_store32(float (&v)[32], const __nv_bfloat16* __restrict__ residual,
                                                 const __nv_bfloat16* __restrict__ saved, __nv_bfloat16* __restrict__ y,
                                                 size_t off, int act, bool halo_right, bool halo_left, size_t halo_elems) {
    if (residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(residual + off);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 rv = __ldg(rp + j4);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                v[j4 * 8 + e * 2] += f.x; v[j4 * 8 + e * 2 + 1] += f.y;
            }
        }
    }
    if (act >= 3) {
        const uint4* sp = reinterpret_cast<const uint4*>(saved + off);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 sv = __ldg(sp + j4);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&sv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __bfloat1622float2(h[e]);
                const float d0 = (act == 3) ? fmaf(-a.x, a.x, 1.0f) : (a.x > 0.0f ? 1.0f : 0.0f);
                const float d1 = (act == 3) ? fmaf(-a.y, a.y, 1.0f) : (a.y > 0.0f ? 1.0f : 0.0f);
                v[j4 * 8 + e * 2] *= d0; v[j4 * 8 + e * 2 + 1] *= d1;
            }
        }
    }
    uint4 out[4];
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            h[e] = __floats2bfloat162_rn(apply_act(v[j4 * 8 + e * 2], act), apply_act(v[j4 * 8 + e * 2 + 1], act));
        out[j4] = *reinterpret_cast<uint4*>(h);
    }
    uint4* yp = reinterpret_cast<uint4*>(y + off);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) yp[j4] = out[j4];
    if (halo_right) {
        uint4* hp = reinterpret_cast<uint4*>(y + off + halo_elems);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) hp[j4] = out[j4];
    }
    if (halo_left) {
        uint4* hp = reinterpret_cast<uint4*>(y + off - halo_elems);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) hp[j4] = out[j4];
    }
}

}  // namespace delora

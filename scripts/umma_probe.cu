// Probe of tcgen05 shared-memory descriptor semantics on sm_100a (development tool, not product code).
//
// Questions answered (each prints PASS/FAIL per variant):
//  T1  K-major SWIZZLE_128B operand whose start address is shifted by s x 128 B (s = 0..9) inside a TMA-written
//      tile: does the MMA read rows s..s+127 correctly (a) with base_offset = 0, (b) with base_offset = (addr>>7)&7 ?
//      -> decides whether the 9 filter taps of a convolution can be issued from ONE halo tile in shared memory.
//  T2  MN-major SWIZZLE_128B operand (wgrad: K = pixels) with the K start shifted by s pixel rows, same two variants.
//  T3  K-major SWIZZLE_64B operands with 64-byte rows (K = 32 per tile) -- the stem layout.
//  T4  TMA load through a tensor map with OVERLAPPING strides (dim-1 stride 32 B < dim-0 extent 64 B): the
//      im2col-by-strides view of the 8-channel stem input.
//  T5  TMA store (shared -> global) of a SWIZZLE_128B tile.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/umma_probe scripts/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* src, const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
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

// descriptor: start | LBO | SBO | version 1 | base_offset | layout
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off,
                                              uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)layout << 61;
    return d;
}

struct ProbeArgs {
    int test;        // 1 K-major SW128 shifted rows, 2 MN-major shifted K rows, 3 SW64 K-major
    int shift;       // rows
    int use_base;    // 0/1
    int a_rows;      // rows of the A source box
    int N;           // 64
};

// smem: A region 24 KB (up to 192 rows x 128 B), B region 16 KB, C staging 16 KB
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ out,
             ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + 32768;
    uint64_t* bar = (uint64_t*)(smem + 65536);
    uint64_t* bar2 = bar + 1;
    uint32_t* tptr = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        if (p.test == 1) {
            // A: [a_rows][64] bf16 K-major rows of 128 B (box 64 x a_rows), B: [N][64]
            mbar_expect_tx(bar, (uint32_t)(p.a_rows * 128 + p.N * 128));
            tma_load_2d(sa, &map_a, bar, 0, 0);
            tma_load_2d(sb, &map_b, bar, 0, 0);
            mbar_wait(bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a0 = smem_u32(sa) + p.shift * 128;
            const uint32_t bo = p.use_base ? ((a0 >> 7) & 7) : 0;
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_desc(a0 + k * 32, 16, 1024, bo, 2);
                const uint64_t db = make_desc(smem_u32(sb) + k * 32, 16, 1024, 0, 2);
                mma_bf16(tmem, da, db, idesc, k > 0);
            }
            tcgen05_commit(bar2);
        } else if (p.test == 2) {
            // A = dZ [64 px][128 co] as two [64 px][64 co] blocks (8 KB each), B = X [a_rows px][64 ci]; K = pixels,
            // both MN-major; B's K start shifted by `shift` pixel rows
            mbar_expect_tx(bar, (uint32_t)(2 * 8192 + p.a_rows * 128));
            tma_load_2d(sa, &map_a, bar, 0, 0);
            tma_load_2d(sa + 8192, &map_a, bar, 64, 0);
            tma_load_2d(sb, &map_b, bar, 0, 0);
            mbar_wait(bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int k = 0; k < 4; ++k) {
                const uint32_t a0 = smem_u32(sa) + k * 2048;
                const uint32_t b0 = smem_u32(sb) + k * 2048 + p.shift * 128;
                const uint32_t bo = p.use_base ? ((b0 >> 7) & 7) : 0;
                const uint64_t da = make_desc(a0, 8192, 1024, 0, 2);
                const uint64_t db = make_desc(b0, 8192, 1024, bo, 2);
                mma_bf16(tmem, da, db, idesc, k > 0);
            }
            tcgen05_commit(bar2);
        } else if (p.test == 3) {
            // SW64: A [128][32] rows of 64 B (+ shift rows), B [N][32]; K = 32 -> 2 MMAs; SBO = 512 B (8 rows x 64 B)
            mbar_expect_tx(bar, (uint32_t)(p.a_rows * 64 + p.N * 64));
            tma_load_2d(sa, &map_a, bar, 0, 0);
            tma_load_2d(sb, &map_b, bar, 0, 0);
            mbar_wait(bar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t a0 = smem_u32(sa) + p.shift * 64;
            const uint32_t bo = p.use_base ? ((a0 >> 7) & 7) : 0;
            for (int k = 0; k < 2; ++k) {
                const uint64_t da = make_desc(a0 + k * 32, 16, 512, bo, 4);
                const uint64_t db = make_desc(smem_u32(sb) + k * 32, 16, 512, 0, 4);
                mma_bf16(tmem, da, db, idesc, k > 0);
            }
            tcgen05_commit(bar2);
        }
    }
    __syncwarp();
    mbar_wait(bar2, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
        for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(acc[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

// T4/T5: TMA load with overlapping strides into SW64 smem, de-swizzle by hand to check; TMA store of a SW128 tile
__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                 __nv_bfloat16* __restrict__ dump, int mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 32768);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (mode == 4) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, 128 * 64);
            tma_load_2d(smem, &map_in, bar, 0, 3);      // 32 elems x 128 rows starting at row 3
        }
        mbar_wait(bar, 0);
        // SW64: 16-byte chunk index c (0..3) of row r is stored at chunk c ^ ((r >> 1) & 3)
        for (int i = threadIdx.x; i < 128 * 32; i += 128) {
            const int r = i / 32, e = i % 32;
            const int chunk = (e / 8) ^ ((r >> 1) & 3);
            dump[i] = reinterpret_cast<const __nv_bfloat16*>(smem)[r * 32 + chunk * 8 + (e % 8)];
        }
    } else {
        // fill a SW128 [128][64] tile by hand: element (r, e) -> chunk (e/8) ^ (r & 7); value = r * 64 + e
        for (int i = threadIdx.x; i < 128 * 64; i += 128) {
            const int r = i / 64, e = i % 64;
            const int chunk = (e / 8) ^ (r & 7);
            reinterpret_cast<__nv_bfloat16*>(smem)[r * 64 + chunk * 8 + (e % 8)] = __float2bfloat16((float)((r * 64 + e) % 251));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_store_2d(smem, &map_out, 0, 5);         // rows 5..132 of the output
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("no cuTensorMapEncodeTiled\n"); exit(2);
    }
    return (PFN_cuTensorMapEncodeTiled_v12000)p;
}

static CUtensorMap make_map_2d(void* ptr, uint64_t d0, uint64_t d1, uint64_t stride1_bytes, uint32_t b0, uint32_t b1,
                               CUtensorMapSwizzle sw, int* rc_out = nullptr) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
    CUtensorMap m;
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc_out) *rc_out = (int)rc;
    else if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); exit(2); }
    return m;
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    CK(cudaSetDevice(0));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    const int R = 192, N = 64;
    std::vector<float> hx(R * 128), hw(N * 64);
    srand(1);
    for (auto& v : hx) v = bf((float)(rand() % 17 - 8) / 8.0f);
    for (auto& v : hw) v = bf((float)(rand() % 13 - 6) / 4.0f);
    std::vector<__nv_bfloat16> bx(hx.size()), bw(hw.size());
    for (size_t i = 0; i < hx.size(); ++i) bx[i] = __float2bfloat16(hx[i]);
    for (size_t i = 0; i < hw.size(); ++i) bw[i] = __float2bfloat16(hw[i]);
    __nv_bfloat16 *dx, *dw; float* dout;
    CK(cudaMalloc(&dx, bx.size() * 2)); CK(cudaMalloc(&dw, bw.size() * 2)); CK(cudaMalloc(&dout, 128 * 128 * 4));
    CK(cudaMemcpy(dx, bx.data(), bx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, bw.data(), bw.size() * 2, cudaMemcpyHostToDevice));
    std::vector<float> hout(128 * 128);

    // ---------------- T1
    printf("T1 K-major SW128, A start shifted by s rows (x128 B)\n");
    {
        // X viewed as [R][64] (first 64 of 128 columns, row stride 256 B)
        CUtensorMap ma = make_map_2d(dx, 64, R, 256, 64, 144, CU_TENSOR_MAP_SWIZZLE_128B);
        CUtensorMap mb = make_map_2d(dw, 64, N, 128, 64, N, CU_TENSOR_MAP_SWIZZLE_128B);
        for (int use_base = 0; use_base < 2; ++use_base)
            for (int s = 0; s <= 9; ++s) {
                ProbeArgs p{1, s, use_base, 144, N};
                CK(cudaMemset(dout, 0, 128 * 128 * 4));
                probe_kernel<<<1, 128, 80 * 1024>>>(ma, mb, dout, p);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hout.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                for (int i = 0; i < 128; ++i)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < 64; ++k) ref += (double)hx[(i + s) * 128 + k] * hw[n * 64 + k];
                        maxerr = fmax(maxerr, fabs(ref - hout[i * N + n]));
                    }
                printf("  base_offset=%s shift=%d maxerr=%.4f %s\n", use_base ? "(a>>7)&7" : "0", s, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
            }
    }
    // ---------------- T2
    printf("T2 MN-major SW128 (K = pixel rows), B K-start shifted by s rows\n");
    {
        // dZ = X viewed as [64 px][128 co] (full 128 columns); Xop = second matrix [R px][64 ci]: reuse hx columns 64..127 of rows
        CUtensorMap ma = make_map_2d(dx, 128, R, 256, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
        CUtensorMap mb = make_map_2d(dx + 64, 64, R, 256, 64, 80, CU_TENSOR_MAP_SWIZZLE_128B);
        for (int use_base = 0; use_base < 2; ++use_base)
            for (int s = 0; s <= 9; ++s) {
                ProbeArgs p{2, s, use_base, 80, 64};
                CK(cudaMemset(dout, 0, 128 * 128 * 4));
                probe_kernel<<<1, 128, 80 * 1024>>>(ma, mb, dout, p);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hout.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                for (int co = 0; co < 128; ++co)
                    for (int ci = 0; ci < 64; ++ci) {
                        double ref = 0;
                        for (int px = 0; px < 64; ++px) ref += (double)hx[px * 128 + co] * hx[(px + s) * 128 + 64 + ci];
                        maxerr = fmax(maxerr, fabs(ref - hout[co * 64 + ci]));
                    }
                printf("  base_offset=%s shift=%d maxerr=%.4f %s\n", use_base ? "(a>>7)&7" : "0", s, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
            }
    }
    // ---------------- T3
    printf("T3 K-major SW64 (64-byte rows, K = 32), A start shifted by s rows (x64 B)\n");
    {
        CUtensorMap ma = make_map_2d(dx, 32, R, 256, 32, 144, CU_TENSOR_MAP_SWIZZLE_64B);
        CUtensorMap mb = make_map_2d(dw, 32, N, 128, 32, N, CU_TENSOR_MAP_SWIZZLE_64B);
        for (int use_base = 0; use_base < 2; ++use_base)
            for (int s = 0; s <= 8; s += (s < 2 ? 1 : 2)) {
                ProbeArgs p{3, s, use_base, 144, N};
                CK(cudaMemset(dout, 0, 128 * 128 * 4));
                probe_kernel<<<1, 128, 80 * 1024>>>(ma, mb, dout, p);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hout.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                for (int i = 0; i < 128; ++i)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < 32; ++k) ref += (double)hx[(i + s) * 128 + k] * hw[n * 64 + k];
                        maxerr = fmax(maxerr, fabs(ref - hout[i * N + n]));
                    }
                printf("  base_offset=%s shift=%d maxerr=%.4f %s\n", use_base ? "(a>>7)&7" : "0", s, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
            }
    }
    // ---------------- T4: overlapping strides
    printf("T4 TMA load with overlapping strides (row stride 32 B, row extent 64 B)\n");
    {
        int rc = 0;
        // flat buffer of bf16 = hx; view [rows][32] with row stride 16 elements (32 B)
        CUtensorMap mi = make_map_2d(dx, 32, 1000, 32, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B, &rc);
        if (rc != 0) printf("  encode rejected overlapping strides: rc=%d FAIL\n", rc);
        else {
            __nv_bfloat16* ddump; CK(cudaMalloc(&ddump, 128 * 32 * 2));
            CUtensorMap mo = make_map_2d(dx, 64, R, 256, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
            tma_probe_kernel<<<1, 128, 48 * 1024>>>(mi, mo, ddump, 4);
            CK(cudaDeviceSynchronize());
            std::vector<__nv_bfloat16> hd(128 * 32);
            CK(cudaMemcpy(hd.data(), ddump, hd.size() * 2, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int e = 0; e < 32; ++e)
                    if (__bfloat162float(hd[r * 32 + e]) != hx[(r + 3) * 16 + e]) ++bad;
            printf("  mismatches=%d %s\n", bad, bad == 0 ? "PASS" : "FAIL");
        }
    }
    // ---------------- T5: TMA store
    printf("T5 TMA store of a SW128 [128][64] tile\n");
    {
        __nv_bfloat16* dy; CK(cudaMalloc(&dy, 200 * 64 * 2)); CK(cudaMemset(dy, 0, 200 * 64 * 2));
        CUtensorMap mo = make_map_2d(dy, 64, 200, 128, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        CUtensorMap mi = make_map_2d(dx, 64, R, 256, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        tma_probe_kernel<<<1, 128, 48 * 1024>>>(mi, mo, nullptr, 5);
        CK(cudaDeviceSynchronize());
        std::vector<__nv_bfloat16> hy(200 * 64);
        CK(cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r = 0; r < 200; ++r)
            for (int e = 0; e < 64; ++e) {
                const float want = (r >= 5 && r < 133) ? (float)(((r - 5) * 64 + e) % 251) : 0.0f;
                if (__bfloat162float(hy[r * 64 + e]) != want) ++bad;
            }
        printf("  mismatches=%d %s\n", bad, bad == 0 ? "PASS" : "FAIL");
    }
    printf("probe done\n");
    return 0;
}

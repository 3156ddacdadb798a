// Probe 2 (development tool): cta_group::2 mechanics and tcgen05.mma issue rates on sm_100a.
//  T6  2-CTA GEMM D[256][256] = X[256][64] * W[256][64]^T: tcgen05.alloc.cta_group::2 in both CTAs, peer TMA loads
//      signalling the leader's mbarrier, tcgen05.mma.cta_group::2 (M = 256), multicast commit, per-CTA TMEM read-back.
//  T7  MMA rate: cycles per tcgen05.mma (SS mode, K = 16) for cta_group::1 M=128 N=64/128/256 and cta_group::2 M=256
//      N=64/128/256, all SMs busy, operands resident in shared memory (no TMA traffic): shows the shared-memory-port limit.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/umma_probe2 scripts/umma_probe2.cu -lcuda
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
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: data into THIS CTA's shared memory, completion bytes on the barrier at `bar_addr` (a shared::cluster
// address: the local address with bit 24 cleared = the even (leader) CTA's copy of the barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void commit_1sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_1sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- T6
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // 128 rows x 128 B
    uint8_t* sb = smem + 16384;         // 128 rows x 128 B (this CTA's half of N)
    uint64_t* full = (uint64_t*)(smem + 32768);
    uint64_t* done = full + 1;
    uint32_t* tptr = (uint32_t*)(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        const uint32_t leader_full = smem_u32(full) & 0xFEFFFFFFu;
        if (rank == 0) mbar_expect_tx(full, 4 * 16384);
        tma_load_2d_2sm(sa, &map_x, leader_full, 0, (int)rank * 128);
        tma_load_2d_2sm(sb, &map_w, leader_full, 0, (int)rank * 128);
        if (rank == 0) {
            mbar_wait(full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = make_idesc(256, 256);
            for (int k = 0; k < 4; ++k)
                mma_2sm(tmem, make_desc(smem_u32(sa) + k * 32), make_desc(smem_u32(sb) + k * 32), idesc, k > 0);
            commit_2sm(done, 3);
        }
    }
    __syncwarp();
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
        for (int j = 0; j < 32; ++j) out[(size_t)(rank * 128 + warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(acc[j]);
    }
    if (threadIdx.x == 0) out[256 * 256 + rank] = __uint_as_float(tmem);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

// ---------------------------------------------------------------- T7
template <int CG>
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int M, int N, int iters, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // 16 KB
    uint8_t* sb = smem + 16384;         // 32 KB
    uint64_t* done = (uint64_t*)(smem + 49152);
    uint32_t* tptr = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = (CG == 2) ? cluster_rank() : 0;
    for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // small bf16 values
    if (threadIdx.x == 0) {
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(256u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(256u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = make_idesc(M, N);
        const uint64_t da = make_desc(smem_u32(sa)), db = make_desc(smem_u32(sb));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (CG == 2) mma_2sm(tmem, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, 1);
            else mma_1sm(tmem, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, 1);
        }
        if (CG == 2) commit_2sm(done, 1); else commit_1sm(done);
        mbar_wait(done, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
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
static CUtensorMap make_map_2d(void* ptr, uint64_t d0, uint64_t d1, uint64_t stride1_bytes, uint32_t b0, uint32_t b1) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
    CUtensorMap m;
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); exit(2); }
    return m;
}
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int CG>
static void run_rate(int M, int N, long long* dcyc) {
    const int iters = 4096;
    const int grid = 148;
    CK(cudaFuncSetAttribute(mma_rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaMemset(dcyc, 0, 148 * 8));
    CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG>, M, N, iters, dcyc));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(148);
    CK(cudaMemcpy(h.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; int n = 0; double sum = 0;
    for (int i = 0; i < 148; ++i) if (h[i] > 0) { mx = fmax(mx, (double)h[i]); sum += h[i]; ++n; }
    const double cyc = sum / n / iters;
    const double ideal = (double)M * N / (256.0 * CG);      // per pair for CG = 2
    printf("  cta_group::%d M=%d N=%d: %.1f cycles/MMA (ideal %.0f) -> %.0f %% of the tensor peak; smem read %.0f B/clk/SM\n", CG, M,
           N, cyc, ideal, 100.0 * ideal / cyc, (M / CG * 32.0 + N / CG * 32.0) / cyc);
}

int main() {
    CK(cudaSetDevice(0));
    // ---------------- T6
    printf("T6 cta_group::2 GEMM 256x256x64\n");
    {
        std::vector<float> hx(256 * 64), hw(256 * 64);
        srand(2);
        for (auto& v : hx) v = bf((float)(rand() % 17 - 8) / 8.0f);
        for (auto& v : hw) v = bf((float)(rand() % 13 - 6) / 4.0f);
        std::vector<__nv_bfloat16> bx(hx.size()), bw(hw.size());
        for (size_t i = 0; i < hx.size(); ++i) bx[i] = __float2bfloat16(hx[i]);
        for (size_t i = 0; i < hw.size(); ++i) bw[i] = __float2bfloat16(hw[i]);
        __nv_bfloat16 *dx, *dw; float* dout;
        CK(cudaMalloc(&dx, bx.size() * 2)); CK(cudaMalloc(&dw, bw.size() * 2)); CK(cudaMalloc(&dout, (256 * 256 + 2) * 4));
        CK(cudaMemcpy(dx, bx.data(), bx.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dw, bw.data(), bw.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dout, 0, (256 * 256 + 2) * 4));
        CUtensorMap mx = make_map_2d(dx, 64, 256, 128, 64, 128);
        CUtensorMap mw = make_map_2d(dw, 64, 256, 128, 64, 128);
        CK(cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
        gemm_2cta_kernel<<<2, 128, 48 * 1024>>>(mx, mw, dout);
        CK(cudaDeviceSynchronize());
        std::vector<float> ho(256 * 256 + 2);
        CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int i = 0; i < 256; ++i)
            for (int n = 0; n < 256; ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)hx[i * 64 + k] * hw[n * 64 + k];
                maxerr = fmax(maxerr, fabs(ref - ho[i * 256 + n]));
            }
        unsigned t0, t1;
        memcpy(&t0, &ho[256 * 256], 4); memcpy(&t1, &ho[256 * 256 + 1], 4);
        printf("  maxerr=%.4f %s (tmem base cta0=0x%x cta1=0x%x)\n", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL", t0, t1);
    }
    // ---------------- T7
    printf("T7 tcgen05.mma issue rate, SS mode, all SMs\n");
    {
        long long* dcyc; CK(cudaMalloc(&dcyc, 148 * 8));
        run_rate<1>(128, 64, dcyc);
        run_rate<1>(128, 128, dcyc);
        run_rate<1>(128, 256, dcyc);
        run_rate<1>(64, 256, dcyc);
        run_rate<2>(256, 64, dcyc);
        run_rate<2>(256, 128, dcyc);
        run_rate<2>(256, 256, dcyc);
        run_rate<2>(128, 256, dcyc);
    }
    printf("probe2 done\n");
    return 0;
}

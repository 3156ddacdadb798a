// Encoder weight gradient, second-generation kernel (round 2).
//
//   dW[co][tap][ci] = sum over output pixels of dZ[pix][co] * X[pix * stride + tap][ci]
// (autograd's backward of the reference's nn.Conv2d layers w.r.t. their weights, src/models/resnet_modified.py:40,
// :126-134).  GEMM with K = PIXELS, both operands pixel-major NHWC tiles (MN-major for the tensor core).
//
// The first-generation kernel (conv_tc.cu) streams a dZ tile AND one X tile per filter tap through shared memory: 104 B
// of fill per tensor-core clock at 64 channels, 162-290 TFLOP/s on the 64/128-channel layers (profiles/r01_*).  Here a
// K-row (one image row segment of NS pixels) is loaded ONCE and the taps are formed in the MMA descriptors:
//   * tap q of a row = the same X tile starting q pixels (q * 128 B) later -- MN-major operands may start at any
//     128-byte offset (scripts/umma_probe.cu T2);
//   * the "leading byte offset" between the 64-channel blocks of an operand is free, so ONE operand can be made of the
//     blocks (tap q0, tap q1, tap q2) of one tile (LBO = 128 B, overlapping views) or of two different rows;
//   mode 0 (Cout % 128 == 0): M = 128 output channels (dZ), N = 192 = 3 taps x 64 input channels per MMA, one filter
//          row r and up to two 64-channel input blocks per CTA (384 TMEM columns);
//   mode 1 (Cout == 64):      M = 128 = 2 taps x 64 input channels (X), N = 64 output channels (dZ), all 9 taps in one
//          CTA as 5 tap pairs (320 TMEM columns) -- M = 64 would halve the tensor rate;
//   strided layers read X through an element-strided tensor map (even / odd columns as separate tiles).
// The per-K-step MMA list ("plan") is built on the host.  Split-K over pixel rows: persistent CTAs, grid = tiles x
// splits <= #SMs, fp32 partials, a coalesced fixed-order reduction (deterministic) into the torch layout.
// Warp roles: 0 TMA producer, 1 MMA issuer (warp-uniform loop, elected lane), 2..5 epilogue.
#include <string.h>
#include "tc_common.cuh"

namespace delora {

constexpr int kWgThreads = 192;
constexpr int kWgMaxLoads = 6, kWgMaxMma = 6, kWgMaxStages = 8;

struct WgLoad { int is_x, ch_blk, dw, dh, smem_off, bytes; };
struct WgMma { int a_off, b_off, a_lbo, b_lbo, N, acc_col, ci_blk; int tap[3]; };

struct Wg2Params {
    int B, Hout, Wout, Cin, Cout, taps;
    int sh, sw;
    int NS, segs, k_rows;
    int mode;
    int co_tiles, ci_tiles, n_tiles, splits;
    int ci_per_tile, co_per_tile;
    int stage_bytes, stages, stage_tx;
    int n_loads, n_mma, acc_alloc;
    WgLoad loads[kWgMaxLoads];
    WgMma mma[kWgMaxMma];
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad2_tc_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                      float* __restrict__ partial, const __grid_constant__ Wg2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = DELORA_ALIGNED_SMEM(smem_raw);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* empty_bar = full_bar + kWgMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kWgMaxStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x % p.n_tiles, split = blockIdx.x / p.n_tiles;
    const int ci_t = tile % p.ci_tiles, co_t = (tile / p.ci_tiles) % p.co_tiles, r = tile / (p.ci_tiles * p.co_tiles);
    const int co0 = co_t * p.co_per_tile, ci0 = ci_t * p.ci_per_tile;
    const int per = (p.k_rows + p.splits - 1) / p.splits;
    const int k_begin = split * per, k_end = min(p.k_rows, k_begin + per);
    const int n_iter = max(0, k_end - k_begin);
    const int S = p.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, (uint32_t)p.acc_alloc);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tma_prefetch_desc(&map_dz);
            tma_prefetch_desc(&map_x);
            uint32_t s = 0, ph = 0;
            for (int it = 0; it < n_iter; ++it) {
                int kt = k_begin + it;
                const int seg = kt % p.segs; kt /= p.segs;
                const int h = kt % p.Hout;
                const int b = kt / p.Hout;
                const int w0 = seg * p.NS;
                mbar_wait(empty_bar + s, ph ^ 1);
                mbar_expect_tx(full_bar + s, (uint32_t)p.stage_tx);
                uint8_t* st = smem + s * p.stage_bytes;
                for (int l = 0; l < p.n_loads; ++l) {
                    const WgLoad& ld = p.loads[l];
                    if (ld.is_x)
                        tma_load_4d(st + ld.smem_off, &map_x, full_bar + s, ci0 + 64 * ld.ch_blk, w0 * p.sw + ld.dw,
                                    h * p.sh + ld.dh + (p.mode == 0 ? r : 0), b);
                    else
                        tma_load_4d(st + ld.smem_off, &map_dz, full_bar + s, co0 + 64 * ld.ch_blk, w0 + 1, h + 1, b);
                }
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        const bool issuer = elect_one();
        const int ksteps = p.NS >> 4, n_mma = p.n_mma;
        const uint64_t desc_hi = make_smem_desc_mn(0, 0);
        uint64_t da_hi[kWgMaxMma], db_hi[kWgMaxMma];
        uint32_t a_off[kWgMaxMma], b_off[kWgMaxMma], idesc[kWgMaxMma], acc_col[kWgMaxMma];
#pragma unroll
        for (int e = 0; e < kWgMaxMma; ++e) {
            const WgMma& m = p.mma[e];
            da_hi[e] = desc_hi | ((uint64_t)((m.a_lbo >> 4) & 0x3FFF) << 16);
            db_hi[e] = desc_hi | ((uint64_t)((m.b_lbo >> 4) & 0x3FFF) << 16);
            a_off[e] = (uint32_t)m.a_off >> 4; b_off[e] = (uint32_t)m.b_off >> 4;
            idesc[e] = make_idesc(128, m.N > 0 ? m.N : 64, 1, 1);
            acc_col[e] = (uint32_t)m.acc_col;
        }
        const uint32_t st_lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4, st_step = (uint32_t)p.stage_bytes >> 4;
        uint32_t s = 0, ph = 0;
        for (int it = 0; it < n_iter; ++it) {
            mbar_wait(full_bar + s, ph);
            tc_fence_after();
            if (issuer) {
                const uint32_t st_lo = st_lo0 + s * st_step;
                for (int k = 0; k < ksteps; ++k) {            // 16 pixels = 16 rows of 128 B = 2048 B = 128 units
                    const uint32_t kofs = (uint32_t)k * 128u;
                    const uint32_t accum = (it > 0 || k > 0) ? 1u : 0u;
#pragma unroll
                    for (int e = 0; e < kWgMaxMma; ++e)
                        if (e < n_mma)
                            tcgen05_mma_bf16(tmem_base + acc_col[e], da_hi[e] | (uint64_t)(st_lo + a_off[e] + kofs),
                                             db_hi[e] | (uint64_t)(st_lo + b_off[e] + kofs), idesc[e], accum);
                }
                tcgen05_commit(empty_bar + s);
                if (it == n_iter - 1) tcgen05_commit(tmem_full_bar);
            }
            __syncwarp();
            if (++s == (uint32_t)S) { s = 0; ph ^= 1; }
        }
    } else {
        // ===================== epilogue: fp32 partial of this (tile, split) =====================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;                           // accumulator row (TMEM lane)
        float* __restrict__ out_split = partial + (size_t)split * p.taps * p.Cout * p.Cin;
        if (n_iter > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
        }
        for (int e = 0; e < p.n_mma; ++e) {
            const WgMma& mm = p.mma[e];
            for (int c0 = 0; c0 < mm.N; c0 += 32) {
                uint32_t acc[32];
                if (n_iter > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mm.acc_col + c0), acc);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = 0u;
                }
                if (p.mode == 0) {
                    // rows = output channels, columns = [tap block][64 input channels]
                    const int tap = 3 * r + mm.tap[c0 >> 6];
                    float4* o4 = reinterpret_cast<float4*>(out_split + ((size_t)tap * p.Cout + co0 + m) * p.Cin + ci0 +
                                                           64 * mm.ci_blk + (c0 & 63));
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o4[j] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                            __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
                } else if (p.mode == 2) {
                    // stem: rows = [filter row block][k = q * 16 + c], columns = output channels
                    const int fr = mm.tap[m >> 6], k = m & 63, q = k >> 4, c = k & 15;
                    if (fr >= 0 && q < 3) {
                        float* o = out_split + ((size_t)(fr * 3 + q) * p.Cout + co0 + c0) * p.Cin + c;
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[(size_t)j * p.Cin] = __uint_as_float(acc[j]);
                    }
                } else {
                    // rows = [tap block][64 input channels], columns = output channels: a warp's 32 lanes are 32
                    // consecutive input channels of one (tap, co) -> coalesced 128-byte stores
                    const int tap = mm.tap[m >> 6];
                    if (tap >= 0) {
                        float* o = out_split + ((size_t)tap * p.Cout + co0 + c0) * p.Cin + ci0 + 64 * mm.ci_blk + (m & 63);
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[(size_t)j * p.Cin] = __uint_as_float(acc[j]);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.acc_alloc);
}

// sum the split-K slices in a fixed order (deterministic) and write the torch layout dW[co][ci][r][s].
// blockDim = (32, 8): thread (x, y) owns 4 consecutive input channels of one (tap, co) and the slices y, y + 8, ...;
// the 8 partial sums are combined through shared memory in y order.  Reads are 512 contiguous bytes per warp.
__global__ void __launch_bounds__(256)
wgrad2_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int Cout, int Cin, int Cin_true,
                     float* __restrict__ dw) {
    __shared__ float4 sm[8][32];
    const size_t total4 = (size_t)taps * Cout * Cin / 4;
    const size_t q = (size_t)blockIdx.x * 32 + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < total4) {
        const float4* src = reinterpret_cast<const float4*>(partial) + q;
        for (int s = threadIdx.y; s < splits; s += 8) {
            const float4 v = __ldg(src + (size_t)s * total4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && q < total4) {
#pragma unroll
        for (int y = 1; y < 8; ++y) {
            const float4 v = sm[y][threadIdx.x];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const size_t i = q * 4;
        const int ci = (int)(i % Cin);
        const int co = (int)((i / Cin) % Cout);
        const int tap = (int)(i / ((size_t)Cin * Cout));
        const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (ci + j < Cin_true) dw[((size_t)co * Cin_true + ci + j) * taps + tap] = vals[j];
    }
}

// stem: dW[co][c][tap] = sum over slices of (partial[tap][co][c] + partial[tap][co][c + 8]) -- channels c and c + 8
// of the 16-channel input are bf16(x) and its rounding residual, which share one weight (conv_stem.cu)
__global__ void __launch_bounds__(256)
stem_reduce_kernel(const float* __restrict__ partial, int splits, int Cin_true, float* __restrict__ dw) {
    const int i = blockIdx.x * 256 + threadIdx.x;            // (tap, co, c < 8)
    if (i >= 9 * 64 * 8) return;
    const int c = i & 7, co = (i >> 3) & 63, tap = i >> 9;
    if (c >= Cin_true) return;
    const size_t total = (size_t)9 * 64 * 16, o = ((size_t)tap * 64 + co) * 16 + c;
    float acc = 0.0f;
    for (int s = 0; s < splits; ++s) acc += __ldg(partial + (size_t)s * total + o) + __ldg(partial + (size_t)s * total + o + 8);
    dw[((size_t)co * Cin_true + c) * 9 + tap] = acc;
}

bool wgrad2_eligible(int Cin, int Cout, int ksize, int stride_h, int stride_w) {
    if (ksize != 3 || Cin % 64 != 0) return false;
    if (Cout % 128 == 0) return true;
    return Cout == 64 && stride_h == 1 && stride_w == 1;
}

static void wgrad2_shape(int B, int Hout, int Wout, int Cin, int Cout, int sw, Wg2Params* p) {
    p->mode = (Cout % 128 == 0) ? 0 : 1;
    const int ns_cap = (sw == 2) ? 64 : 128;
    p->NS = Wout >= ns_cap ? ns_cap : (Wout + 15) / 16 * 16;
    p->segs = (Wout + p->NS - 1) / p->NS;
    p->k_rows = B * Hout * p->segs;
    if (p->mode == 0) {
        const int nb = Cin >= 128 ? 2 : 1;
        p->ci_per_tile = 64 * nb; p->co_per_tile = 128;
        p->ci_tiles = Cin / p->ci_per_tile; p->co_tiles = Cout / 128;
        p->n_tiles = 3 * p->ci_tiles * p->co_tiles;
    } else {
        p->ci_per_tile = 64; p->co_per_tile = 64;
        p->ci_tiles = Cin / 64; p->co_tiles = Cout / 64;
        p->n_tiles = p->ci_tiles * p->co_tiles;
    }
    int splits = kNumSMs / p->n_tiles;
    if (splits < 1) splits = 1;
    if (splits > p->k_rows) splits = p->k_rows > 0 ? p->k_rows : 1;
    p->splits = splits;
}

int64_t wgrad2_scratch_floats(int B, int Hout, int Wout, int Cin, int Cout, int sw) {
    Wg2Params p;
    wgrad2_shape(B, Hout, Wout, Cin, Cout, sw, &p);
    return (int64_t)p.splits * 9 * Cout * Cin;
}

int wgrad2_launch(const void* x, const void* dz, float* dw, float* scratch, int B, int Hin, int Win, int Cin, int Cin_true,
                  int Cout, int stride_h, int stride_w, cudaStream_t st) {
    Wg2Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.taps = 9; p.sh = stride_h; p.sw = stride_w;
    p.Hout = (Hin - 1) / stride_h + 1; p.Wout = (Win - 1) / stride_w + 1;
    wgrad2_shape(B, p.Hout, p.Wout, Cin, Cout, stride_w, &p);
    const int NS = p.NS;
    const int DB = NS * 128;                                        // one 64-channel block of dZ
    const int xpix = (stride_w == 2) ? NS + 1 : NS + 2;             // pixels per X tile
    const int TS = (xpix * 128 + 1023) / 1024 * 1024;
    int nl = 0, nm = 0;
    if (p.mode == 0) {
        const int nb = p.ci_per_tile / 64;
        p.loads[nl++] = WgLoad{0, 0, 0, 0, 0, DB};
        p.loads[nl++] = WgLoad{0, 1, 0, 0, DB, DB};
        const int xo = 2 * DB;
        for (int j = 0; j < nb; ++j) {
            if (stride_w == 1) {
                p.loads[nl++] = WgLoad{1, j, 0, 0, xo + j * TS, xpix * 128};
                WgMma m = {0, xo + j * TS, DB, 128, 192, j * 192, j, {0, 1, 2}};
                p.mma[nm++] = m;
            } else {
                // even columns (taps q = 0 and, one pixel later, q = 2), odd columns (tap q = 1)
                p.loads[nl++] = WgLoad{1, j, 0, 0, xo + (2 * j) * TS, xpix * 128};
                p.loads[nl++] = WgLoad{1, j, 1, 0, xo + (2 * j + 1) * TS, xpix * 128};
                WgMma me = {0, xo + (2 * j) * TS, DB, 128, 128, j * 192, j, {0, 2, -1}};
                WgMma mo = {0, xo + (2 * j + 1) * TS, DB, 128, 64, j * 192 + 128, j, {1, -1, -1}};
                p.mma[nm++] = me; p.mma[nm++] = mo;
            }
        }
        p.stage_bytes = xo + nb * (stride_w == 2 ? 2 : 1) * TS;
        p.acc_alloc = nb * 192 <= 256 ? 256 : 512;
    } else {
        p.loads[nl++] = WgLoad{0, 0, 0, 0, 0, DB};
        const int xo = DB;
        for (int rr = 0; rr < 3; ++rr) p.loads[nl++] = WgLoad{1, 0, 0, rr, xo + rr * TS, xpix * 128};
        // A = X views: pairs of taps as the two 64-channel blocks of M; the block distance (LBO) is 128 B for
        // horizontally adjacent taps and TS - 256 B from (row r, q = 2) to (row r + 1, q = 0)
        const WgMma plan[5] = {
            {xo + 0 * TS + 0, 0, 128, 0, 64, 0, 0, {0, 1, -1}},
            {xo + 0 * TS + 256, 0, TS - 256, 0, 64, 64, 0, {2, 3, -1}},
            {xo + 1 * TS + 128, 0, 128, 0, 64, 128, 0, {4, 5, -1}},
            {xo + 2 * TS + 0, 0, 128, 0, 64, 192, 0, {6, 7, -1}},
            {xo + 2 * TS + 256, 0, 128, 0, 64, 256, 0, {8, -1, -1}},
        };
        for (int e = 0; e < 5; ++e) p.mma[nm++] = plan[e];
        p.stage_bytes = xo + 3 * TS;
        p.acc_alloc = 512;
    }
    p.n_loads = nl; p.n_mma = nm;
    p.stage_tx = 0;
    for (int l = 0; l < nl; ++l) p.stage_tx += p.loads[l].bytes;
    int stages = (225 * 1024 - 512) / p.stage_bytes;
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    DELORA_CHECK_ARG(stages >= 2, "conv_wgrad2: stage of %d bytes does not fit twice", p.stage_bytes);
    p.stages = stages;

    PFN_cuTensorMapEncodeTiled_v12000 encode = get_tensor_map_encoder();
    DELORA_CHECK_ARG(encode != nullptr, "conv_wgrad2: cuTensorMapEncodeTiled not available");
    CUtensorMap map_dz, map_x;
    {
        const int Hp = p.Hout + 2, Wp = p.Wout + 2;
        // extents stop at the last REAL pixel: a ragged K row reads zeros from the out-of-bounds fill, not the halo
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)(Wp - 1), (cuuint64_t)(Hp - 1), (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wp * Cout * 2, (cuuint64_t)Hp * Wp * Cout * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)NS, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult rc = encode(&map_dz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dz), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DELORA_CHECK_ARG(rc == CUDA_SUCCESS, "conv_wgrad2: tensor map (dz) failed: %d", (int)rc);
    }
    {
        const int Hp = Hin + 2, Wp = Win + 2;
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)Wp * Cin * 2, (cuuint64_t)Hp * Wp * Cin * 2};
        // with a traversal stride the box spans xpix * stride_w elements and loads every stride_w-th one
        cuuint32_t box[4] = {64, (cuuint32_t)(xpix * stride_w), 1, 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, 1, 1};
        CUresult rc = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DELORA_CHECK_ARG(rc == CUDA_SUCCESS, "conv_wgrad2: tensor map (x) failed: %d", (int)rc);
    }
    const size_t smem = (size_t)p.stages * p.stage_bytes + (2 * kWgMaxStages + 1) * 8 + 16 + 1024;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        DELORA_CHECK_ARG(e == cudaSuccess, "conv_wgrad2: shared-memory opt-in failed: %s", cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    conv_wgrad2_tc_kernel<<<p.n_tiles * p.splits, kWgThreads, smem, st>>>(map_dz, map_x, scratch, p);
    DELORA_CHECK_LAUNCH("conv_wgrad2_tc_kernel");
    const size_t total4 = (size_t)9 * Cout * Cin / 4;
    wgrad2_reduce_kernel<<<(unsigned)((total4 + 31) / 32), dim3(32, 8), 0, st>>>(scratch, p.splits, 9, Cout, Cin, Cin_true, dw);
    DELORA_CHECK_LAUNCH("wgrad2_reduce_kernel");
    return 0;
}

// stem (conv_stem.cu): X tiles come from the overlapping-stride map (one 128-byte row = the 4 x 16-channel window of
// an output pixel), M = 128 = two filter rows x 64 k, N = 64 output channels, two MMAs per K-step
static void wgrad2_stem_shape(int B, int H, int Wout, Wg2Params* p) {
    p->mode = 2;
    p->NS = 128; p->segs = (Wout + 127) / 128; p->k_rows = B * H * p->segs;
    p->ci_per_tile = 16; p->co_per_tile = 64; p->ci_tiles = 1; p->co_tiles = 1; p->n_tiles = 1;
    p->splits = p->k_rows < kNumSMs ? (p->k_rows > 0 ? p->k_rows : 1) : kNumSMs;
}

int64_t wgrad2_stem_scratch_floats(int B, int H, int Wout) {
    Wg2Params p;
    wgrad2_stem_shape(B, H, Wout, &p);
    return (int64_t)p.splits * 9 * 64 * 16;
}

int wgrad2_launch_stem(const CUtensorMap& map_x, const void* dz, float* dw, float* scratch, int B, int H, int Wout,
                       int Cin_true, cudaStream_t st) {
    Wg2Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.Cin = 16; p.Cout = 64; p.taps = 9; p.sh = 1; p.sw = 1; p.Hout = H; p.Wout = Wout;
    wgrad2_stem_shape(B, H, Wout, &p);
    const int T = 16384;
    p.loads[0] = WgLoad{0, 0, 0, 0, 0, T};
    for (int rr = 0; rr < 3; ++rr) p.loads[1 + rr] = WgLoad{1, 0, 0, rr, T + rr * T, T};
    const WgMma m0 = {T, 0, T, 0, 64, 0, 0, {0, 1, -1}};           // filter rows 0, 1
    const WgMma m1 = {3 * T, 0, T, 0, 64, 64, 0, {2, -1, -1}};     // filter row 2 (+ an ignored block)
    p.mma[0] = m0; p.mma[1] = m1;
    p.n_loads = 4; p.n_mma = 2; p.stage_bytes = 4 * T; p.stage_tx = 4 * T; p.stages = 3; p.acc_alloc = 128;
    PFN_cuTensorMapEncodeTiled_v12000 encode = get_tensor_map_encoder();
    DELORA_CHECK_ARG(encode != nullptr, "stem wgrad: cuTensorMapEncodeTiled not available");
    CUtensorMap map_dz;
    {
        const int Hp = H + 2, Wp = Wout + 2;
        cuuint64_t dims[4] = {64, (cuuint64_t)(Wp - 1), (cuuint64_t)(Hp - 1), (cuuint64_t)B};
        cuuint64_t strides[3] = {128, (cuuint64_t)Wp * 128, (cuuint64_t)Hp * Wp * 128};
        cuuint32_t box[4] = {64, 128, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult rc = encode(&map_dz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dz), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DELORA_CHECK_ARG(rc == CUDA_SUCCESS, "stem wgrad: tensor map (dz) failed: %d", (int)rc);
    }
    // + one tile of slack after the last stage: the ignored second block of the last MMA reads it
    const size_t smem = (size_t)p.stages * p.stage_bytes + T + (2 * kWgMaxStages + 1) * 8 + 16 + 1024;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        DELORA_CHECK_ARG(e == cudaSuccess, "stem wgrad: shared-memory opt-in failed: %s", cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    conv_wgrad2_tc_kernel<<<p.n_tiles * p.splits, kWgThreads, smem, st>>>(map_dz, map_x, scratch, p);
    DELORA_CHECK_LAUNCH("conv_wgrad2_tc_kernel (stem)");
    stem_reduce_kernel<<<(9 * 64 * 8 + 255) / 256, 256, 0, st>>>(scratch, p.splits, Cin_true, dw);
    DELORA_CHECK_LAUNCH("stem_reduce_kernel");
    return 0;
}

}  // namespace delora

// Per-pixel surface normals of a range image (sm_100a).
//
// Replaces preprocessing.normal_computation.NormalsComputer.compute_normal_vectors
// (reference: src/preprocessing/normal_computation.py:89-122 gather of the (2a+1)x(2b+1)
// edge-clamped patch; :53-87 range gate, >= min neighbours, eigen-decomposition, orientation)
// and utility.linalg.cov (src/utility/linalg.py:33-56 zero-aware mean / covariance).
//
// One CTA stages a (TH+2a) x (TW+2b) tile of (x, y, z, |p|) as float4 in shared memory, already
// edge-clamped (one LDS.128 per tap); every thread owns TWO vertically adjacent pixels as one packed
// fp32x2 pair (FFMA2 / FADD2 / FMUL2).  ONE pass over the taps in coordinates relative to the pixel
// itself: gated count, sum and second moments, then C = (sum w d d^T - n m m^T) / (n - 1) -- the
// reference (linalg.py:33-56) subtracts the mean in a second sweep; shifting by the centre first makes
// the one-pass form as accurate -- then a cyclic Jacobi eigen-solve of the 3x3
// covariance in registers, smallest-eigenvalue eigenvector, flipped toward the sensor.
// The kernel is FP32-pipe / issue bound (~2.4 kFLOP per pixel against 28 B of HBM traffic), not HBM
// bound: bench.py reports it against the fp32 roofline; see DESIGN.md.
#include "common.cuh"
#include "tc_common.cuh"

namespace delora {

constexpr int kNormTW = 32;
constexpr int kNormTH = 8;

struct Sym3 { float a00, a01, a02, a11, a12, a22; };

// Cyclic Jacobi for a symmetric 3x3; returns the eigenvector of the smallest eigenvalue.
// The rotation parameters use the fast reciprocal / rsqrt units (MUFU): a Jacobi rotation only has
// to be orthogonal (c^2 + s^2 = 1 to fp32 rounding, guaranteed by s = t*c, c = rsqrt(1 + t^2)) and
// to shrink the off-diagonal entry; a 2-ulp error in the angle costs at most an extra sweep.  The
// IEEE-exact divisions/square roots of the first version were ~45 % of the kernel's instructions.
__device__ __forceinline__ void smallest_eigenvector(Sym3 m, float& nx, float& ny, float& nz) {
    float a[3][3] = {{m.a00, m.a01, m.a02}, {m.a01, m.a11, m.a12}, {m.a02, m.a12, m.a22}};
    float v[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
        const float off = fabsf(a[0][1]) + fabsf(a[0][2]) + fabsf(a[1][2]);
        const float diag = fabsf(a[0][0]) + fabsf(a[1][1]) + fabsf(a[2][2]);
        // fp32 convergence: off-diagonal mass below one ulp of the diagonal mass (Jacobi converges
        // quadratically: typically 3-4 sweeps; a stricter test never fires in fp32 and just burns sweeps)
        if (off <= 3e-8f * diag) break;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = (k == 2) ? 1 : 0;
            const int q = (k == 0) ? 1 : 2;
            const float apq = a[p][q];
            // theta = (aqq - app) / (2 apq);  t = sgn(theta) / (|theta| + sqrt(theta^2 + 1))
            // written without a division by apq:  t = sgn * 2|apq| / (|d| + sqrt(d^2 + 4 apq^2)), d = aqq - app
            const float d = a[q][q] - a[p][p];
            const float two_apq = 2.0f * apq;
            const float h2 = fmaf(d, d, two_apq * two_apq);
            const float hyp = h2 * rsqrtf(fmaxf(h2, 1e-38f));
            const float sgn = ((d >= 0.0f) == (apq >= 0.0f)) ? 1.0f : -1.0f;
            const float t = (apq == 0.0f) ? 0.0f : sgn * __fdividef(fabsf(two_apq), fabsf(d) + hyp);
            const float c = rsqrtf(fmaf(t, t, 1.0f));
            const float sn = t * c;
            const float tau = __fdividef(sn, 1.0f + c);
            a[p][p] -= t * apq;
            a[q][q] += t * apq;
            a[p][q] = 0.0f; a[q][p] = 0.0f;
            const int r = 3 - p - q;
            const float arp = a[r][p], arq = a[r][q];
            a[r][p] = arp - sn * (arq + tau * arp);
            a[r][q] = arq + sn * (arp - tau * arq);
            a[p][r] = a[r][p]; a[q][r] = a[r][q];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float vip = v[i][p], viq = v[i][q];
                v[i][p] = vip - sn * (viq + tau * vip);
                v[i][q] = viq + sn * (vip - tau * viq);
            }
        }
    }
    int j = 0;
    float best = a[0][0];
    if (a[1][1] < best) { best = a[1][1]; j = 1; }
    if (a[2][2] < best) { j = 2; }
    nx = (j == 0) ? v[0][0] : (j == 1 ? v[0][1] : v[0][2]);
    ny = (j == 0) ? v[1][0] : (j == 1 ? v[1][1] : v[1][2]);
    nz = (j == 0) ? v[2][0] : (j == 1 ? v[2][1] : v[2][2]);
    const float inv = rsqrtf(nx * nx + ny * ny + nz * nz);
    nx *= inv; ny *= inv; nz *= inv;
}

// ---------------------------------------------------------------------------------------------
// Fast path for the reference's 7 x 11 patch (config/config_datasets.yaml:33,60).
// Tile = 32 columns x 16 rows per CTA of 256 threads; every thread owns kFastPix = 2 vertically adjacent
// pixels (one packed fp32x2 pair), so one shared-memory read of a neighbour feeds both patches (8 x 11
// LDS.128 for 2 x 77 taps).  Four pixels per thread (two pairs, 92 registers, 20 warps/SM) cost the same
// instructions per pixel and measured 3 % slower than two (62 registers, 32 warps/SM).  The halo is staged ALREADY CLAMPED (positions
// outside the image replicate the edge pixel, exactly the reference's index clamp,
// normal_computation.py:104-111), so the tap loops use constant offsets and no index math.
// tile.w = |p|, or +inf for an all-zero pixel: "neighbour present" (range gate passed and
// not (0,0,0), linalg.py:34-37) is then the single test !(|q.w - c.w| > eps).
constexpr int kFastTW = 32, kFastTH = 16, kFastPix = 2, kFastA = 3, kFastB = 5;
constexpr int kFastThreads = kFastTW * (kFastTH / kFastPix);            // 256
constexpr int kFastTileW = kFastTW + 2 * kFastB;                        // 42
constexpr int kFastTileH = kFastTH + 2 * kFastA;                        // 22

// Everything after the staging: the tap loop, the eigen-solve and the stores (shared by both staging variants).
__device__ __forceinline__ void normals_7x11_body(const float4* __restrict__ tile, int bi, int u0, int v0, int H, int W,
                                                  float eps_range, int min_nb, float* __restrict__ normals,
                                                  float4* __restrict__ pts_grid, float4* __restrict__ nrm_grid) {
    const size_t HW = (size_t)H * W;
    const float inf = __int_as_float(0x7f800000);
    const int lu = threadIdx.x % kFastTW, lr = threadIdx.x / kFastTW;       // lr: which group of 4 rows
    const int u = u0 + lu, vb = v0 + lr * kFastPix;
    if (u >= W || vb >= H) return;
    // thread's window: tile rows [lr*4, lr*4 + 9], tile cols [lu, lu + 10]
    const float4* __restrict__ base = tile + (lr * kFastPix) * kFastTileW + lu;

    float4 c[kFastPix];
    bool valid[kFastPix];
#pragma unroll
    for (int k = 0; k < kFastPix; ++k) {
        c[k] = base[(k + kFastA) * kFastTileW + kFastB];
        valid[k] = (vb + k < H) && c[k].x != 0.0f && c[k].y != 0.0f && c[k].z != 0.0f;     // :35
    }
    // The pixels are processed as PAIRS in packed fp32x2 arithmetic (Blackwell FFMA2 /
    // FADD2 / FMUL2: one instruction per two fp32 lanes, IEEE round-to-nearest per lane, so the
    // numbers are those of the scalar code).  A neighbour outside one pixel's 7 rows gets weight 0.
    constexpr int kPairs = kFastPix / 2;
    // ONE pass over the taps, in coordinates relative to the pixel itself (d = q - c, |d| <~ 1 m where the
    // coordinates are up to ~100 m): gated count, sum and second moments; the covariance follows as
    //     C = (sum w d d^T - n m m^T) / (n - 1),   m = sum w d / n.
    // The reference (utility/linalg.py:33-56) subtracts the mean in a second pass; shifting by the centre
    // first makes the one-pass form as accurate (the cancellation n m m^T vs sum d d^T is between numbers of
    // the size of the neighbourhood, not of the scene) and saves the second sweep over the 110 shared-memory
    // vectors and its gates: -26 % instructions.  Measured against LAPACK on the goldens: same error
    // distribution as the two-pass kernel (tests/test_gpu_parity.py prints it).
    float2 sx[kPairs], sy[kPairs], sz[kPairs], cnt[kPairs];
    float2 a00[kPairs], a01[kPairs], a02[kPairs], a11[kPairs], a12[kPairs], a22[kPairs];
    float2 ncx[kPairs], ncy[kPairs], ncz[kPairs], cw[kPairs];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) {
        sx[p] = sy[p] = sz[p] = cnt[p] = make_float2(0.f, 0.f);
        a00[p] = a01[p] = a02[p] = a11[p] = a12[p] = a22[p] = make_float2(0.f, 0.f);
        ncx[p] = make_float2(-c[2 * p].x, -c[2 * p + 1].x);
        ncy[p] = make_float2(-c[2 * p].y, -c[2 * p + 1].y);
        ncz[p] = make_float2(-c[2 * p].z, -c[2 * p + 1].z);
        cw[p] = make_float2(c[2 * p].w, c[2 * p + 1].w);
    }
#pragma unroll 1
    for (int du = 0; du < 2 * kFastB + 1; ++du) {
#pragma unroll
        for (int r = 0; r < kFastPix + 2 * kFastA; ++r) {
            const float4 q = base[r * kFastTileW + du];
#pragma unroll
            for (int p = 0; p < kPairs; ++p) {
                const bool in0 = (r - 2 * p >= 0) && (r - 2 * p <= 2 * kFastA);             // compile-time
                const bool in1 = (r - 2 * p - 1 >= 0) && (r - 2 * p - 1 <= 2 * kFastA);
                if (in0 || in1) {
                    const bool g0 = in0 && !(fabsf(q.w - cw[p].x) > eps_range);             // :56-59 + linalg.py:34-37
                    const bool g1 = in1 && !(fabsf(q.w - cw[p].y) > eps_range);
                    const float2 w = make_float2(g0 ? 1.0f : 0.0f, g1 ? 1.0f : 0.0f);
                    const float2 dx = __fadd2_rn(make_float2(q.x, q.x), ncx[p]);
                    const float2 dy = __fadd2_rn(make_float2(q.y, q.y), ncy[p]);
                    const float2 dz = __fadd2_rn(make_float2(q.z, q.z), ncz[p]);
                    const float2 ex = __fmul2_rn(dx, w), ey = __fmul2_rn(dy, w), ez = __fmul2_rn(dz, w);
                    sx[p] = __fadd2_rn(sx[p], ex); sy[p] = __fadd2_rn(sy[p], ey); sz[p] = __fadd2_rn(sz[p], ez);
                    cnt[p] = __fadd2_rn(cnt[p], w);
                    a00[p] = __ffma2_rn(ex, dx, a00[p]); a01[p] = __ffma2_rn(ex, dy, a01[p]);
                    a02[p] = __ffma2_rn(ex, dz, a02[p]); a11[p] = __ffma2_rn(ey, dy, a11[p]);
                    a12[p] = __ffma2_rn(ey, dz, a12[p]); a22[p] = __ffma2_rn(ez, dz, a22[p]);
                }
            }
        }
    }
    int n[kFastPix];
    bool go[kFastPix];
    Sym3 s[kFastPix];
#pragma unroll
    for (int k = 0; k < kFastPix; ++k) {
        const int p = k / 2;
        const float fn = (k & 1) ? cnt[p].y : cnt[p].x;
        const float ssx = (k & 1) ? sx[p].y : sx[p].x, ssy = (k & 1) ? sy[p].y : sy[p].x, ssz = (k & 1) ? sz[p].y : sz[p].x;
        n[k] = (int)fn;
        go[k] = valid[k] && n[k] >= min_nb;                                       // :67-69
        const float inv_n = __fdiv_rn(1.0f, fmaxf(fn, 1.0f));
        const float mx = ssx * inv_n, my = ssy * inv_n, mz = ssz * inv_n;         // mean of d
        s[k] = (k & 1) ? Sym3{a00[p].y, a01[p].y, a02[p].y, a11[p].y, a12[p].y, a22[p].y}
                       : Sym3{a00[p].x, a01[p].x, a02[p].x, a11[p].x, a12[p].x, a22[p].x};
        s[k].a00 = fmaf(-ssx, mx, s[k].a00); s[k].a01 = fmaf(-ssx, my, s[k].a01); s[k].a02 = fmaf(-ssx, mz, s[k].a02);
        s[k].a11 = fmaf(-ssy, my, s[k].a11); s[k].a12 = fmaf(-ssy, mz, s[k].a12); s[k].a22 = fmaf(-ssz, mz, s[k].a22);
    }
#pragma unroll
    for (int k = 0; k < kFastPix; ++k) {
        const int v = vb + k;
        if (v >= H) continue;
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (go[k]) {
            const float f = __fdiv_rn(1.0f, (float)(n[k] - 1));                   // linalg.py:39, :56
            Sym3 t = s[k];
            t.a00 *= f; t.a01 *= f; t.a02 *= f; t.a11 *= f; t.a12 *= f; t.a22 *= f;
            smallest_eigenvector(t, nx, ny, nz);                                  // torch.symeig, evec[:, :, 0]
            if (nx * c[k].x + ny * c[k].y + nz * c[k].z > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }   // :79-81
        }
        const size_t pix = (size_t)v * W + u;
        if (normals) {
            float* __restrict__ out = normals + (size_t)bi * 3 * HW + pix;
            out[0] = nx; out[HW] = ny; out[2 * HW] = nz;
        }
        if (pts_grid) {
            pts_grid[(size_t)bi * HW + pix] = valid[k] ? make_float4(c[k].x, c[k].y, c[k].z, __int_as_float((int)pix))
                                                       : make_float4(inf, inf, inf, __int_as_float(-1));
            const bool has = (nx != 0.0f) | (ny != 0.0f) | (nz != 0.0f);          // icp_losses.py:48-52
            nrm_grid[(size_t)bi * HW + pix] = make_float4(nx, ny, nz, has ? 1.0f : 0.0f);
        }
    }
}

// Staging variant 1 (default): coalesced loads, index clamp, repack -- all in one pass.
__global__ void __launch_bounds__(kFastThreads)
normals_7x11_kernel(const float* __restrict__ image, int C_img, int H, int W, float eps_range, int min_nb,
                    float* __restrict__ normals, float4* __restrict__ pts_grid, float4* __restrict__ nrm_grid) {
    __shared__ float4 tile[kFastTileH * kFastTileW];
    const int bi = blockIdx.z;
    const int u0 = blockIdx.x * kFastTW, v0 = blockIdx.y * kFastTH;
    const size_t HW = (size_t)H * W;
    const float* __restrict__ img = image + (size_t)bi * C_img * HW;
    const float inf = __int_as_float(0x7f800000);

    for (int i = threadIdx.x; i < kFastTileH * kFastTileW; i += kFastThreads) {
        const int ly = i / kFastTileW, lx = i - ly * kFastTileW;
        const int gv = min(max(v0 - kFastA + ly, 0), H - 1), gu = min(max(u0 - kFastB + lx, 0), W - 1);
        const size_t o = (size_t)gv * W + gu;
        const float x = __ldg(img + o), y = __ldg(img + HW + o), z = __ldg(img + 2 * HW + o);
        const bool zero = (x == 0.0f) & (y == 0.0f) & (z == 0.0f);
        tile[i] = make_float4(x, y, z, zero ? inf : range3(x, y, z));
    }
    __syncthreads();
    normals_7x11_body(tile, bi, u0, v0, H, W, eps_range, min_nb, normals, pts_grid, nrm_grid);
}

// Staging variant 2 (delora_normals_select_staging(1); built to MEASURE what TMA buys here): ONE
// cp.async.bulk.tensor.4d box brings the three channel planes of the halo tile into shared memory (out-of-image
// positions zero-filled), then the same repack as above runs from shared memory -- the edge clamp becomes a clamped
// read of the planes, the zero test and |p| stay.  The tensor-map box cannot do that repack, so the pass over the
// 22 x 42 positions remains; what TMA removes is the address arithmetic and the three global loads per position.
// The box starts 8 (not 5) columns left of the tile: TMA wants the first element of a box row on a 16-byte boundary
// (u0 - 5 faulted with "illegal instruction" on the UTMALDG), so the planes are 48 floats wide.
constexpr int kFastPlaneOff = 8, kFastPlaneW = 48;
__global__ void __launch_bounds__(kFastThreads)
normals_7x11_tma_kernel(const __grid_constant__ CUtensorMap map_img, int H, int W, float eps_range, int min_nb,
                        float* __restrict__ normals, float4* __restrict__ pts_grid, float4* __restrict__ nrm_grid) {
    __shared__ __align__(128) float planes[3][kFastTileH][kFastPlaneW];
    __shared__ float4 tile[kFastTileH * kFastTileW];
    __shared__ __align__(8) uint64_t bar;
    const int bi = blockIdx.z;
    const int u0 = blockIdx.x * kFastTW, v0 = blockIdx.y * kFastTH;
    const float inf = __int_as_float(0x7f800000);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, (uint32_t)sizeof(planes));
        tma_load_4d(&planes[0][0][0], &map_img, &bar, u0 - kFastPlaneOff, v0 - kFastA, 0, bi);
    }
    __syncthreads();                                                     // barrier initialised before anybody polls it
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < kFastTileH * kFastTileW; i += kFastThreads) {
        const int ly = i / kFastTileW, lx = i - ly * kFastTileW;
        const int sv = min(max(v0 - kFastA + ly, 0), H - 1) - (v0 - kFastA);     // the reference's index clamp, in tile
        const int su = min(max(u0 - kFastB + lx, 0), W - 1) - (u0 - kFastPlaneOff);   // coordinates (always inside the box)
        const float x = planes[0][sv][su], y = planes[1][sv][su], z = planes[2][sv][su];
        const bool zero = (x == 0.0f) & (y == 0.0f) & (z == 0.0f);
        tile[i] = make_float4(x, y, z, zero ? inf : range3(x, y, z));
    }
    __syncthreads();
    normals_7x11_body(tile, bi, u0, v0, H, W, eps_range, min_nb, normals, pts_grid, nrm_grid);
}

// Generic patch sizes (and the first version of the kernel): one pixel per thread, runtime loops.
// A, B: half sizes of the patch when known at compile time (loops unroll); -1 = runtime.
template <int A_, int B_>
__global__ void __launch_bounds__(kNormTW * kNormTH)
normals_kernel(const float* __restrict__ image, int C_img, int H, int W, int a_rt, int b_rt,
               float eps_range, int min_nb, float* __restrict__ normals, float4* __restrict__ pts_grid,
               float4* __restrict__ nrm_grid) {
    extern __shared__ float4 tile[];
    const int a = (A_ >= 0) ? A_ : a_rt;
    const int b = (B_ >= 0) ? B_ : b_rt;
    const int tw = kNormTW + 2 * b, th = kNormTH + 2 * a;
    const int bi = blockIdx.z;
    const int u0 = blockIdx.x * kNormTW, v0 = blockIdx.y * kNormTH;
    const size_t HW = (size_t)H * W;
    const float* __restrict__ img = image + (size_t)bi * C_img * HW;

    // stage the halo tile; out-of-image positions are never addressed (neighbour coordinates are
    // clamped to the image first, normal_computation.py:104-111), so they are left unset.
    for (int i = threadIdx.x; i < tw * th; i += kNormTW * kNormTH) {
        const int ly = i / tw, lx = i - ly * tw;
        const int gv = v0 - a + ly, gu = u0 - b + lx;
        if (gv >= 0 && gv < H && gu >= 0 && gu < W) {
            const size_t o = (size_t)gv * W + gu;
            const float x = __ldg(img + o), y = __ldg(img + HW + o), z = __ldg(img + 2 * HW + o);
            tile[i] = make_float4(x, y, z, range3(x, y, z));
        }
    }
    __syncthreads();

    const int lu = threadIdx.x % kNormTW, lv = threadIdx.x / kNormTW;
    const int u = u0 + lu, v = v0 + lv;
    if (u >= W || v >= H) return;
    const float4 c = tile[(lv + a) * tw + (lu + b)];
    float nx = 0.f, ny = 0.f, nz = 0.f;
    const bool valid = c.x != 0.0f && c.y != 0.0f && c.z != 0.0f;     // normal_computation.py:35
    if (valid) {
        const int ntaps = (2 * a + 1) * (2 * b + 1);
        // clamped window in tile coordinates
        float sx = 0.f, sy = 0.f, sz = 0.f;
        int n = 0;
#pragma unroll 1
        for (int dv = -a; dv <= a; ++dv) {
            const int vv = min(max(v + dv, 0), H - 1) - (v0 - a);
            const float4* __restrict__ row = tile + vv * tw;
#pragma unroll
            for (int du = -b; du <= b; ++du) {
                const int uu = min(max(u + du, 0), W - 1) - (u0 - b);
                const float4 q = row[uu];
                const bool gated = fabsf(q.w - c.w) > eps_range;          // :56-59
                const bool present = !gated && ((q.x != 0.0f) | (q.y != 0.0f) | (q.z != 0.0f));   // linalg.py:34-37
                if (present) { sx += q.x; sy += q.y; sz += q.z; ++n; }
            }
        }
        if (n >= min_nb) {                                            // :67-69
            const float k = (float)ntaps, fn = (float)n;
            // torch.mean(dim) * K / n   (linalg.py:41-42)
            const float mx = __fdiv_rn(__fmul_rn(__fdiv_rn(sx, k), k), fn);
            const float my = __fdiv_rn(__fmul_rn(__fdiv_rn(sy, k), k), fn);
            const float mz = __fdiv_rn(__fmul_rn(__fdiv_rn(sz, k), k), fn);
            Sym3 s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int dv = -a; dv <= a; ++dv) {
                const int vv = min(max(v + dv, 0), H - 1) - (v0 - a);
                const float4* __restrict__ row = tile + vv * tw;
#pragma unroll
                for (int du = -b; du <= b; ++du) {
                    const int uu = min(max(u + du, 0), W - 1) - (u0 - b);
                    const float4 q = row[uu];
                    const bool gated = fabsf(q.w - c.w) > eps_range;
                    const bool present = !gated && ((q.x != 0.0f) | (q.y != 0.0f) | (q.z != 0.0f));
                    if (present) {
                        const float dx = q.x - mx, dy = q.y - my, dz = q.z - mz;    // linalg.py:43-46
                        s.a00 = fmaf(dx, dx, s.a00); s.a01 = fmaf(dx, dy, s.a01); s.a02 = fmaf(dx, dz, s.a02);
                        s.a11 = fmaf(dy, dy, s.a11); s.a12 = fmaf(dy, dz, s.a12); s.a22 = fmaf(dz, dz, s.a22);
                    }
                }
            }
            const float f = __fdiv_rn(1.0f, (float)(n - 1));           // linalg.py:39, :56
            s.a00 *= f; s.a01 *= f; s.a02 *= f; s.a11 *= f; s.a12 *= f; s.a22 *= f;
            smallest_eigenvector(s, nx, ny, nz);                      // torch.symeig, evec[:, :, 0]  (:70-76)
            if (nx * c.x + ny * c.y + nz * c.z > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }   // :79-81
        }
    }
    const size_t pix = (size_t)v * W + u;
    if (normals) {
        float* __restrict__ out = normals + (size_t)bi * 3 * HW + pix;
        out[0] = nx; out[HW] = ny; out[2 * HW] = nz;
    }
    if (pts_grid) {
        // dense float4 grids for the ICP kernel: one valid point per cell, empty cells at +inf
        const float inf = __int_as_float(0x7f800000);
        pts_grid[(size_t)bi * HW + pix] = valid ? make_float4(c.x, c.y, c.z, __int_as_float((int)pix))
                                                : make_float4(inf, inf, inf, __int_as_float(-1));
        const bool has = (nx != 0.0f) | (ny != 0.0f) | (nz != 0.0f);                     // icp_losses.py:48-52
        nrm_grid[(size_t)bi * HW + pix] = make_float4(nx, ny, nz, has ? 1.0f : 0.0f);
    }
}

}  // namespace delora

using namespace delora;

static int g_normals_staging = 0;      // 0: coalesced loads (default), 1: TMA box + repack from shared memory
extern "C" int delora_normals_select_staging(int mode) {
    const int old = g_normals_staging;
    if (mode == 0 || mode == 1) g_normals_staging = mode;
    return old;
}

extern "C" int delora_normals_fwd(const float* image, int B, int C_img, int H, int W, int nb_h, int nb_w,
                                  float epsilon_range, int min_neighbors, float* normals, delora_f4* pts_grid,
                                  delora_f4* nrm_grid, void* stream) {
    DELORA_CHECK_ARG(image && (normals || pts_grid), "delora_normals_fwd: null pointer");
    DELORA_CHECK_ARG((pts_grid == nullptr) == (nrm_grid == nullptr), "delora_normals_fwd: pts_grid and nrm_grid go together");
    DELORA_CHECK_ARG(B > 0 && B <= 65535 && C_img >= 3 && H > 0 && W > 0, "delora_normals_fwd: bad shape");
    const int a = nb_h / 2, b = nb_w / 2;                              // int(side/2): normal_computation.py:97-98
    DELORA_CHECK_ARG(a >= 0 && b >= 0 && a <= 16 && b <= 32, "delora_normals_fwd: neighbourhood %dx%d unsupported",
                     nb_h, nb_w);
    const size_t smem = (size_t)(kNormTW + 2 * b) * (kNormTH + 2 * a) * sizeof(float4);
    dim3 grid((W + kNormTW - 1) / kNormTW, (H + kNormTH - 1) / kNormTH, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (a == kFastA && b == kFastB) {
        dim3 gridf((W + kFastTW - 1) / kFastTW, (H + kFastTH - 1) / kFastTH, B);
        bool tma_done = false;
        if (g_normals_staging == 1 && (W % 4) == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0) {
            PFN_cuTensorMapEncodeTiled_v12000 encode = get_tensor_map_encoder();
            DELORA_CHECK_ARG(encode != nullptr, "delora_normals_fwd: cuTensorMapEncodeTiled unavailable");
            CUtensorMap m;
            cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C_img, (cuuint64_t)B};
            cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C_img * H * W * 4};
            cuuint32_t box[4] = {(cuuint32_t)kFastPlaneW, (cuuint32_t)kFastTileH, 3, 1};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(image), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            DELORA_CHECK_ARG(rc == CUDA_SUCCESS, "delora_normals_fwd: tensor map failed: %d", (int)rc);
            normals_7x11_tma_kernel<<<gridf, kFastThreads, 0, st>>>(m, H, W, epsilon_range, min_neighbors, normals,
                                                                    (float4*)pts_grid, (float4*)nrm_grid);
            tma_done = true;
        }
        if (!tma_done)
            normals_7x11_kernel<<<gridf, kFastThreads, 0, st>>>(image, C_img, H, W, epsilon_range, min_neighbors, normals,
                                                                (float4*)pts_grid, (float4*)nrm_grid);
    } else {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(normals_kernel<-1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem);
            DELORA_CHECK_ARG(e == cudaSuccess, "delora_normals_fwd: smem opt-in failed: %s", cudaGetErrorString(e));
        }
        normals_kernel<-1, -1><<<grid, kNormTW * kNormTH, smem, st>>>(image, C_img, H, W, a, b, epsilon_range,
                                                                      min_neighbors, normals, (float4*)pts_grid,
                                                                      (float4*)nrm_grid);
    }
    DELORA_CHECK_LAUNCH("normals_kernel");
    return 0;
}

// Fused SE(3) transform + exact NN + ICP losses fwd/bwd on DENSE range-image grids (sm_100a).
//
// Same contract as icp.cu (reference: src/deploy/deployer.py:181-189, src/losses/icp_losses.py:28-240),
// specialised for the case the training step actually has: source and target are the valid pixels
// of two projected H x W range images, i.e. at most one point per spherical cell.  Points and
// normals live as float4 per pixel ((x, y, z, pixel id) / (nx, ny, nz, has_normal); empty pixels
// are (+inf, +inf, +inf, -1)), written by the normals kernel, so no list compaction and no
// CSR index are needed and a window of cells is a fixed-stride walk over one array.
//
// One thread per SOURCE PIXEL; a warp = 32 adjacent pixels.  The search window is described by
// warp-uniform extents (rows down/up, columns left/right of every lane's own centre cell) and
// grows one strip at a time while ANY lane's exactness guard still fails, on the side the first
// failing lane needs.  Control flow is warp-uniform (no divergence), adjacent lanes read adjacent
// cells (coalesced, L1-resident), and the guard is the same proof as in icp.cu: all unsearched
// targets lie beyond a border half-plane (azimuth) or cone (elevation) of the lane's window.
#include <stdlib.h>
#include "icp_common.cuh"

namespace delora {

constexpr int kDenseThreads = 128;

__device__ __forceinline__ int wrap_col(int c, int W) {
    c += (c < 0) ? W : 0;
    c -= (c >= W) ? W : 0;
    return c;
}

// lower bound of the distance from the source point to everything beyond a window border that is
// `dpx` pixels (of `rad_per_px` radians) away; `radius` is |p| (elevation cones) or |p_xy| (azimuth planes)
__device__ __forceinline__ float border_bound(float dpx, float rad_per_px, float radius) {
    const float d = fminf(fmaxf(dpx * rad_per_px, 0.0f), kHalfPiF);
    return radius * __sinf(d);
}

// Branch-free running (smallest, second smallest) fp32 squared distance + index of the smallest.
// The whole search runs on this; float64 is only consulted afterwards if the two smallest are
// within kNNBand of each other (nn_exact_rescan).
struct NN2 {
    float m1, m2;
    int j1;
};

__device__ __forceinline__ void nn2_eval(const float4 t, int j, float sx, float sy, float sz, NN2& s,
                                         bool ok = true) {
    const float dx = sx - t.x, dy = sy - t.y, dz = sz - t.z;
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float d2f = ok ? d2 : __int_as_float(0x7f800000);      // out-of-grid rows / idle lanes never win
    const bool lt = d2f < s.m1;
    s.m2 = lt ? s.m1 : fminf(s.m2, d2f);
    s.j1 = lt ? j : s.j1;
    s.m1 = lt ? d2f : s.m1;
}

__device__ __forceinline__ float4 inf4() {
    const float inf = __int_as_float(0x7f800000);
    return make_float4(inf, inf, inf, __int_as_float(-1));
}

// Rare path: two candidates within the fp32 ambiguity band.  Re-rank every candidate of the
// final window whose fp32 distance is inside the band in float64 (lowest pixel id on exact ties).
__device__ __noinline__ int nn_exact_rescan(const float4* __restrict__ tg, int H, int W, int rc, int cc, int e_dn,
                                            int e_up, int e_lf, int e_rt, float sx, float sy, float sz,
                                            float thresh) {
    double best = 1.0e300;
    int bj = -1;
    for (int dr = -e_dn; dr <= e_up; ++dr) {
        const int row = rc + dr;
        if (row < 0 || row >= H) continue;
        for (int dc = -e_lf; dc <= e_rt; ++dc) {
            const int j = row * W + wrap_col(cc + dc, W);
            const float4 t = __ldg(tg + j);
            const float dx = sx - t.x, dy = sy - t.y, dz = sz - t.z;
            const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d2f <= thresh) {
                const double d2 = nn_d2_exact(sx, sy, sz, t.x, t.y, t.z);
                if (d2 < best || (d2 == best && j < bj)) { best = d2; bj = j; }
            }
        }
    }
    return bj;
}

// ---------------------------------------------------------------------------------------------
// Range pyramid for the far / misaligned case.  The strip search above certifies exactness with a
// purely ANGULAR bound, so a source point that is d away from the target surface has to scan every
// cell within the angle asin(d/r) -- thousands of cells when the predicted transform is still poor
// (d ~ 1 m), although almost all of them are provably farther than d once their RANGE is taken into
// account.  Blocks of 4 x 16 cells carry [min, max] of |t|; a block whose lower bound
//     min over rho in [rmin, rmax] of (r_s - rho)^2 + 4 r_s rho sin^2(theta/2),
//     sin^2(theta/2) >= max( sin^2(gap_el/2), cos(e_s) * min cos(e_blk) * sin^2(gap_az/2) )
// exceeds the current best distance cannot hold the nearest neighbour and is skipped.
constexpr int kBlkH = 4, kBlkW = 16;
__device__ unsigned int g_dbg[8];     // phase-2 statistics (owners, blocks tested, blocks scanned, max scanned/owner, warps)
constexpr int kDefaultMaxStrips = 64; // strip expansions before a lane switches to the block search (measured:
                                      // 24 -> +25 % on well-aligned pairs from serialised owners, 64 -> +2 %)

__global__ void __launch_bounds__(256)
block_range_kernel(const float4* __restrict__ grid, int H, int W, int nbh, int nbw, float2* __restrict__ blk) {
    const int b = blockIdx.y;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nbh * nbw) return;
    const int br = warp / nbw, bc = warp % nbw;
    float lo = 3.0e38f, hi = -1.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        const int row = br * kBlkH + (c >> 4), col = bc * kBlkW + (c & 15);
        if (row < H && col < W) {
            const float4 t = __ldg(grid + (size_t)b * H * W + (size_t)row * W + col);
            if (__float_as_int(t.w) >= 0) {
                const float rr = sqrtf(fmaf(t.z, t.z, fmaf(t.y, t.y, t.x * t.x)));
                lo = fminf(lo, rr); hi = fmaxf(hi, rr);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // widen by a relative 1e-6 so that the bound stays conservative against the fp32 norm rounding
    if (lane == 0) blk[(size_t)b * nbh * nbw + warp] = make_float2(lo * (1.0f - 1e-6f), hi * (1.0f + 1e-6f));
}

// first / last pixel column of the (unwrapped) block column bc; bc in [-nbw, 2 nbw)
__device__ __forceinline__ int blk_u_lo(int bc, int nbw, int W) {
    const int k = (bc < 0) ? -1 : (bc >= nbw) ? 1 : 0;
    return (bc - k * nbw) * kBlkW + k * W;
}
__device__ __forceinline__ int blk_u_hi(int bc, int nbw, int W) {
    const int k = (bc < 0) ? -1 : (bc >= nbw) ? 1 : 0;
    return min((bc - k * nbw) * kBlkW + kBlkW - 1, W - 1) + k * W;
}

__device__ __forceinline__ void nn2_merge(NN2& a, float bm1, float bm2, int bj) {
    if (bj >= 0 && bj == a.j1) { a.m2 = fminf(a.m2, bm2); return; }     // the same cell seen twice is not a tie
    const bool take = (bm1 < a.m1) || (bm1 == a.m1 && bj >= 0 && (a.j1 < 0 || bj < a.j1));
    const float lose = take ? a.m1 : bm1;                   // the larger of the two minima
    a.m2 = fminf(fminf(a.m2, bm2), (bj >= 0 && a.j1 >= 0) ? lose : 3.0e38f);
    if (take) { a.m1 = bm1; a.j1 = bj; }
}

template <bool PO2PO>
__global__ void __launch_bounds__(kDenseThreads, 8)
icp_dense_kernel(const float4* __restrict__ src_grid, const float4* __restrict__ src_ngrid,
                 const float* __restrict__ T, const float4* __restrict__ tgt_grid,
                 const float4* __restrict__ tgt_ngrid, const float2* __restrict__ blk_range, int nbh, int nbw,
                 int max_strips, GridParams g, uint32_t flags, float* __restrict__ partial_rows, int rows_per_pair) {
    const int b = blockIdx.y;
    const int H = g.H, W = g.W, HW = H * W;
    const float4* __restrict__ tg = tgt_grid + (size_t)b * HW;
    const float4* __restrict__ tn = tgt_ngrid + (size_t)b * HW;
    const int i = blockIdx.x * kDenseThreads + threadIdx.x;
    const int warp_in_pair = i >> 5;
    constexpr float kInf = 3.0e38f;
    constexpr float kSlack = 2e-3f;     // px; covers the fp32 error of the cell binning

    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), m = p;
    bool active = false;
    if (i < HW) {
        p = __ldg(src_grid + (size_t)b * HW + i);
        active = __float_as_int(p.w) >= 0;
        if (active) m = __ldg(src_ngrid + (size_t)b * HW + i);
    }
    float sx = 0.f, sy = 0.f, sz = 0.f, nsx = 0.f, nsy = 0.f, nsz = 0.f;
    if (active) {
        const Rigid rt = load_rigid(T + (size_t)b * 12);
        apply_rigid(rt, p, m, sx, sy, sz, nsx, nsy, nsz);
    }
    int best_j = -1;
    if (__ballot_sync(0xffffffffu, active) != 0u) {
        float us = 0.f, vs = 0.f;
        pixel_coords(g, sx, sy, sz, us, vs);
        if (!(us == us)) us = 0.0f;
        if (!(vs == vs)) vs = 0.0f;
        const float rxy = sqrtf(fmaf(sx, sx, sy * sy));
        const float r = sqrtf(fmaf(sz, sz, fmaf(sx, sx, sy * sy)));
        const int cc = (int)fminf(fmaxf(rintf(us), 0.0f), g.wm1);
        const int rc = (int)fminf(fmaxf(rintf(vs), 0.0f), g.hm1);
        // distance (px) from the source direction to the four borders of its own centre cell
        const float f_dn = (vs - (float)rc) + 0.5f - kSlack, f_up = ((float)rc - vs) + 0.5f - kSlack;
        const float f_lf = (us - (float)cc) + 0.5f - kSlack, f_rt = ((float)cc - us) + 0.5f - kSlack;

        NN2 nn;
        nn.m1 = kInf; nn.m2 = kInf; nn.j1 = -1;
        // warp-uniform window extents around every lane's own (rc, cc); the 3 x 5 start window is
        // fully unrolled: 15 independent loads in flight
        int e_dn = 1, e_up = 1, e_lf = 2, e_rt = 2;
        bool ext_private = false;      // set for lanes whose extents were widened by the block search
        (void)ext_private;
        if (H >= 3 && W >= 5) {
            int col[5];
#pragma unroll
            for (int dc = -2; dc <= 2; ++dc) col[dc + 2] = wrap_col(cc + dc, W);
#pragma unroll
            for (int dr = -1; dr <= 1; ++dr) {
                const int row = rc + dr;
                const bool rok = active && row >= 0 && row < H;
                const int rbase = min(max(row, 0), H - 1) * W;          // always a valid address
#pragma unroll
                for (int dc = 0; dc < 5; ++dc) nn2_eval(__ldg(tg + rbase + col[dc]), rbase + col[dc], sx, sy, sz, nn, rok);
            }
        } else {
            e_dn = e_up = e_lf = e_rt = 0;
            nn2_eval(__ldg(tg + rc * W + cc), rc * W + cc, sx, sy, sz, nn, active);
        }
        float b_dn = (rc - e_dn > 0) ? border_bound(f_dn + (float)e_dn, g.dv_rad, r) : kInf;
        float b_up = (rc + e_up < H - 1) ? border_bound(f_up + (float)e_up, g.dv_rad, r) : kInf;
        // a border at or beyond the +-180 deg seam is seam_px closer than its unwrapped pixel distance
        float b_lf = (e_lf + e_rt + 1 >= W) ? kInf
                   : border_bound(f_lf + (float)e_lf - (cc - e_lf <= 0 ? g.seam_px : 0.0f), g.du_rad, rxy);
        float b_rt = (e_lf + e_rt + 1 >= W) ? kInf
                   : border_bound(f_rt + (float)e_rt - (cc + e_rt >= W - 1 ? g.seam_px : 0.0f), g.du_rad, rxy);
        unsigned pending = 0u;                                   // lanes that still fail after kMaxStrips strips
        for (int strip = 0;; ++strip) {
            const float bmin = fminf(fminf(b_dn, b_up), fminf(b_lf, b_rt));
            const bool done = !active || bmin >= kInf ||
                              (nn.j1 >= 0 && sqrtf(nn.m1) * 1.00001f <= bmin * 0.9995f);
            const unsigned failing = __ballot_sync(0xffffffffu, !done);
            if (failing == 0u) break;
            if (strip >= max_strips) { pending = failing; break; }
            const int my_side = (bmin == b_dn) ? 0 : (bmin == b_up) ? 1 : (bmin == b_lf) ? 2 : 3;
            const int side = __shfl_sync(0xffffffffu, my_side, __ffs(failing) - 1);
            if (side < 2) {
                const int row = (side == 0) ? rc - (++e_dn) : rc + (++e_up);
                const bool rok = active && row >= 0 && row < H;
                const int rbase = min(max(row, 0), H - 1) * W;
                const int n = e_lf + e_rt + 1;
                int col = wrap_col(cc - e_lf, W);
#pragma unroll 4
                for (int k = 0; k < n; ++k) {
                    nn2_eval(__ldg(tg + rbase + col), rbase + col, sx, sy, sz, nn, rok);
                    ++col;
                    col = (col == W) ? 0 : col;
                }
                if (side == 0) b_dn = (rc - e_dn > 0) ? border_bound(f_dn + (float)e_dn, g.dv_rad, r) : kInf;
                else           b_up = (rc + e_up < H - 1) ? border_bound(f_up + (float)e_up, g.dv_rad, r) : kInf;
            } else {
                const int step = min(2, W - (e_lf + e_rt + 1));
                const int c0 = wrap_col((side == 2) ? cc - e_lf - 1 : cc + e_rt + 1, W);
                const int c1 = wrap_col((side == 2) ? cc - e_lf - 2 : cc + e_rt + 2, W);
                const int n = e_dn + e_up + 1;
                const bool two = (step == 2);
#pragma unroll 2
                for (int k = 0; k < n; ++k) {
                    const int row = rc - e_dn + k;
                    const bool rok = active && row >= 0 && row < H;
                    const int rbase = min(max(row, 0), H - 1) * W;
                    nn2_eval(__ldg(tg + rbase + c0), rbase + c0, sx, sy, sz, nn, rok);
                    nn2_eval(__ldg(tg + rbase + c1), rbase + c1, sx, sy, sz, nn, rok && two);
                }
                if (side == 2) e_lf += step; else e_rt += step;
                const bool full_w = (e_lf + e_rt + 1 >= W);
                b_lf = full_w ? kInf : border_bound(f_lf + (float)e_lf - (cc - e_lf <= 0 ? g.seam_px : 0.0f), g.du_rad, rxy);
                b_rt = full_w ? kInf : border_bound(f_rt + (float)e_rt - (cc + e_rt >= W - 1 ? g.seam_px : 0.0f), g.du_rad, rxy);
            }
        }
        // ---------------- phase 2: warp-cooperative, range-pruned block search for the remaining lanes
        const float2* __restrict__ blk = blk_range + (size_t)b * nbh * nbw;
        const int lane = threadIdx.x & 31;
        if (pending && (threadIdx.x & 31) == 0) { atomicAdd(&g_dbg[5], 1u); atomicMax(&g_dbg[6], (unsigned)__popc(pending)); }
        while (pending) {
            const int owner = __ffs(pending) - 1;
            pending &= pending - 1;
            const float osx = __shfl_sync(0xffffffffu, sx, owner), osy = __shfl_sync(0xffffffffu, sy, owner),
                        osz = __shfl_sync(0xffffffffu, sz, owner);
            const float ous = __shfl_sync(0xffffffffu, us, owner), ovs = __shfl_sync(0xffffffffu, vs, owner);
            const float orr = __shfl_sync(0xffffffffu, r, owner), orxy = __shfl_sync(0xffffffffu, rxy, owner);
            const int orc = __shfl_sync(0xffffffffu, rc, owner), occ = __shfl_sync(0xffffffffu, cc, owner);
            float obest = __shfl_sync(0xffffffffu, nn.m1, owner);         // running best d^2 (fp32), warp-uniform
            const int obr = orc / kBlkH, obc = occ / kBlkW;
            const float cos_es = orr > 0.0f ? orxy / orr : 1.0f;
            NN2 loc;
            loc.m1 = kInf; loc.m2 = kInf; loc.j1 = -1;
            // Block rectangle that certifies the CURRENT best: grow it (warp-uniform scalar loops) until all
            // four borders are at least d0 away; the best can only shrink while the rectangle is scanned, so
            // one pass over its blocks is enough (no ring-by-ring dependency chain).
            const float d0 = (obest < kInf) ? sqrtf(obest) * (1.00001f / 0.9995f) : kInf;
            int br_lo = obr, br_hi = obr, bc_lo = obc, bc_hi = obc;            // bc_* unwrapped
            while (br_lo > 0 && border_bound(ovs - ((float)(br_lo * kBlkH) - 0.5f) - kSlack, g.dv_rad, orr) < d0) --br_lo;
            while (br_hi < nbh - 1 &&
                   border_bound(((float)min(br_hi * kBlkH + kBlkH - 1, H - 1) + 0.5f) - ovs - kSlack, g.dv_rad, orr) < d0)
                ++br_hi;
            // (unwrapped block columns: W need not be a multiple of kBlkW, so the pixel span of block column
            //  bc outside [0, nbw) is that of its wrapped twin shifted by +-W, not bc * kBlkW)
            while (bc_hi - bc_lo + 1 < nbw &&
                   border_bound(ous - ((float)blk_u_lo(bc_lo, nbw, W) - 0.5f) - kSlack - (bc_lo <= 0 ? g.seam_px : 0.0f),
                                g.du_rad, orxy) < d0) --bc_lo;
            while (bc_hi - bc_lo + 1 < nbw &&
                   border_bound(((float)blk_u_hi(bc_hi, nbw, W) + 0.5f) - ous - kSlack - (bc_hi >= nbw - 1 ? g.seam_px : 0.0f),
                                g.du_rad, orxy) < d0) ++bc_hi;
            const int nbc = bc_hi - bc_lo + 1, nblk = (br_hi - br_lo + 1) * nbc;
            int dbg_scanned = 0;
            for (int base = 0; base < nblk; base += 32) {
                const int idx = base + lane;
                const bool in = idx < nblk;
                const int br = br_lo + (in ? idx / nbc : 0);
                int bc = bc_lo + (in ? idx % nbc : 0);
                bc += (bc < 0) ? nbw : 0;
                bc -= (bc >= nbw) ? nbw : 0;
                bool ok = in;
                float lb2 = kInf;
                if (ok) {
                    const float2 rg = __ldg(blk + br * nbw + bc);
                    if (rg.y >= 0.0f) {
                        const float v_lo = (float)(br * kBlkH) - 0.5f, v_hi = (float)min(br * kBlkH + kBlkH - 1, H - 1) + 0.5f;
                        const float u_lo = (float)(bc * kBlkW) - 0.5f, u_hi = (float)min(bc * kBlkW + kBlkW - 1, W - 1) + 0.5f;
                        const float gv = fmaxf(fmaxf(v_lo - ovs, ovs - v_hi) - kSlack, 0.0f) * g.dv_rad;
                        // azimuth gap ON THE CIRCLE: the direct way, or the other way round through the seam
                        const float d1 = u_lo - ous, d2 = ous - u_hi;
                        const float gpx = fminf(fmaxf(fmaxf(d1, d2), 0.0f), g.circ_px + fminf(d1, d2));
                        const float gu = fmaxf(gpx - kSlack, 0.0f) * g.du_rad;
                        const float sv = __sinf(fminf(0.5f * gv, kHalfPiF)), su = __sinf(fminf(0.5f * gu, kHalfPiF));
                        const float e_lo = g.vf0 + v_lo * g.dv_rad, e_hi = g.vf0 + v_hi * g.dv_rad;
                        const float c_blk = fmaxf(fminf(__cosf(e_lo), __cosf(e_hi)), 0.0f);
                        const float S = fmaxf(sv * sv, cos_es * c_blk * su * su);
                        const float rho = fminf(fmaxf(orr * (1.0f - 2.0f * S), rg.x), rg.y);
                        const float dr_ = orr - rho;
                        lb2 = fmaf(dr_, dr_, 4.0f * orr * rho * S) * 0.999f;
                    } else {
                        ok = false;                                       // empty block
                    }
                }
                unsigned need = __ballot_sync(0xffffffffu, ok && lb2 <= obest * 1.001f);
                dbg_scanned += __popc(need);
                while (need) {
                    const int sel = __ffs(need) - 1;
                    need &= need - 1;
                    const int sbr = __shfl_sync(0xffffffffu, br, sel), sbc = __shfl_sync(0xffffffffu, bc, sel);
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int c = lane + 32 * h2;
                        const int row = sbr * kBlkH + (c >> 4), col = sbc * kBlkW + (c & 15);
                        const bool cok = row < H && col < W;
                        const int j = min(row, H - 1) * W + min(col, W - 1);
                        nn2_eval(__ldg(tg + j), j, osx, osy, osz, loc, cok);
                    }
                }
                // tighten the running best with what this chunk found (prunes the following chunks harder)
                float wm = loc.m1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wm = fminf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
                obest = fminf(obest, wm);
            }
            if (lane == 0) {
                atomicAdd(&g_dbg[0], 1u); atomicAdd(&g_dbg[1], (unsigned)nblk); atomicAdd(&g_dbg[2], (unsigned)dbg_scanned);
                atomicMax(&g_dbg[3], (unsigned)dbg_scanned); atomicMax(&g_dbg[4], (unsigned)nblk);
            }
            // merge the 32 partial results, then into the owner's own phase-1 result
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float bm1 = __shfl_xor_sync(0xffffffffu, loc.m1, o), bm2 = __shfl_xor_sync(0xffffffffu, loc.m2, o);
                const int bj = __shfl_xor_sync(0xffffffffu, loc.j1, o);
                nn2_merge(loc, bm1, bm2, bj);
            }
            if (lane == owner) {
                nn2_merge(nn, loc.m1, loc.m2, loc.j1);
                // extents of everything that has been examined (for the float64 tie re-ranking)
                e_dn = max(e_dn, rc - br_lo * kBlkH);
                e_up = max(e_up, min(br_hi * kBlkH + kBlkH - 1, H - 1) - rc);
                e_lf = max(e_lf, min(cc - blk_u_lo(bc_lo, nbw, W), W - 1));
                e_rt = max(e_rt, min(blk_u_hi(bc_hi, nbw, W) - cc, W - 1));
                ext_private = true;
            }
        }
        best_j = nn.j1;
        if (active && best_j >= 0 && nn.m2 <= nn.m1 * (1.0f + kNNBand))
            best_j = nn_exact_rescan(tg, H, W, rc, cc, e_dn, e_up, e_lf, e_rt, sx, sy, sz, nn.m1 * (1.0f + kNNBand));
    }
    float acc[kIcpAcc];
#pragma unroll
    for (int k = 0; k < kIcpAcc; ++k) acc[k] = 0.0f;
    if (active && best_j >= 0) {
        float4 pd, nd;
        accumulate_pair(PO2PO ? flags : (flags & ~DELORA_LOSS_PO2PO), p, m, sx, sy, sz, nsx, nsy, nsz,
                        __ldg(tg + best_j), __ldg(tn + best_j), acc, pd, nd);
    }
    if (warp_in_pair < rows_per_pair)
        write_warp_partials<(PO2PO ? kIcpAcc : 24)>(
            acc, partial_rows + ((size_t)b * rows_per_pair + warp_in_pair) * DELORA_ICP_PARTIAL);
}

}  // namespace delora

using namespace delora;

extern "C" int delora_debug_counters(unsigned int* out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, delora::g_dbg, 8 * sizeof(unsigned int));
    if (reset) { unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(delora::g_dbg, z, sizeof(z)); }
    return 0;
}

extern "C" int delora_icp_dense_fwd_bwd(const delora_f4* src_grid, const delora_f4* src_ngrid, const float* T,
                                        const delora_f4* tgt_grid, const delora_f4* tgt_ngrid, int B, int H, int W,
                                        double hfov0, double hfov1, double vfov0, double vfov1, float lambda_po2pl,
                                        uint32_t flags, float* losses, float* grad_T, float* scratch, void* stream) {
    DELORA_CHECK_ARG(src_grid && src_ngrid && T && tgt_grid && tgt_ngrid && losses && grad_T && scratch,
                     "delora_icp_dense_fwd_bwd: null pointer");
    DELORA_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0, "delora_icp_dense_fwd_bwd: bad shape");
    const GridParams g = make_grid(H, W, hfov0, hfov1, vfov0, vfov1, 1);
    const int HW = H * W;
    const int rows = (HW + 31) / 32;                  // one partial row per warp of 32 source pixels
    const IcpScratch sc = icp_scratch(scratch, B, rows);
    dim3 grid((HW + kDenseThreads - 1) / kDenseThreads, B);
    cudaStream_t st = (cudaStream_t)stream;
    int max_strips = kDefaultMaxStrips;
    if (const char* e = getenv("DELORA_ICP_MAX_STRIPS")) max_strips = atoi(e);      // tuning knob (tests / profiling)
    // range pyramid of the TARGET grids, kept behind the partial rows / column sums / counters of the scratch
    const int nbh = (H + kBlkH - 1) / kBlkH, nbw = (W + kBlkW - 1) / kBlkW;
    float2* blk = reinterpret_cast<float2*>(scratch + (size_t)B * delora_icp_partial_rows(HW) * DELORA_ICP_PARTIAL +
                                            (size_t)B * DELORA_ICP_PARTIAL + (((size_t)B + 1) & ~(size_t)1));
    {
        dim3 gb((nbh * nbw * 32 + 255) / 256, B);
        block_range_kernel<<<gb, 256, 0, st>>>((const float4*)tgt_grid, H, W, nbh, nbw, blk);
        DELORA_CHECK_LAUNCH("block_range_kernel");
    }
    if (flags & DELORA_LOSS_PO2PO) {
        icp_dense_kernel<true><<<grid, kDenseThreads, 0, st>>>((const float4*)src_grid, (const float4*)src_ngrid, T,
                                                               (const float4*)tgt_grid, (const float4*)tgt_ngrid, blk,
                                                               nbh, nbw, max_strips, g, flags, sc.rows, rows);
    } else {
        icp_dense_kernel<false><<<grid, kDenseThreads, 0, st>>>((const float4*)src_grid, (const float4*)src_ngrid, T,
                                                                (const float4*)tgt_grid, (const float4*)tgt_ngrid, blk,
                                                                nbh, nbw, max_strips, g, flags, sc.rows, rows);
    }
    DELORA_CHECK_LAUNCH("icp_dense_kernel");
    return launch_icp_finalize(scratch, B, rows, lambda_po2pl, flags, losses, grad_T, st);
}

// Shared device/host helpers for the delora_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/delora_b200.h"

namespace delora {

void set_error(const char* fmt, ...);

#define DELORA_CHECK_ARG(cond, ...)                        \
    do {                                                   \
        if (!(cond)) {                                     \
            ::delora::set_error(__VA_ARGS__);              \
            return 1;                                      \
        }                                                  \
    } while (0)

#define DELORA_CHECK_LAUNCH(name)                                                   \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ::delora::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return 2;                                                               \
        }                                                                           \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// Spherical grid parameters shared by the projection, the cell binning and the NN search.
// The float constants are produced exactly as torch produces them from the reference's Python
// doubles: the scalar operand of `tensor - float` / `tensor / float` is rounded to fp32 once.
struct GridParams {
    int H, W;
    float hf0, hspan, hinv, wm1;   // u = (atan2(y,x) - hf0) / hspan * wm1
    float vf0, vspan, vinv, hm1;   // v = (atan2(z,|xy|) - vf0) / vspan * hm1
    float du_rad, dv_rad;          // radians per pixel step (hspan/wm1, vspan/hm1)
    int div_mode;
};

inline GridParams make_grid(int H, int W, double hf0, double hf1, double vf0, double vf1, int div_mode) {
    GridParams g;
    g.H = H; g.W = W;
    g.hf0 = (float)hf0; g.hspan = (float)(hf1 - hf0); g.hinv = 1.0f / g.hspan; g.wm1 = (float)(W - 1);
    g.vf0 = (float)vf0; g.vspan = (float)(vf1 - vf0); g.vinv = 1.0f / g.vspan; g.hm1 = (float)(H - 1);
    g.du_rad = (W > 1) ? (float)((hf1 - hf0) / (double)(W - 1)) : 1.0f;
    g.dv_rad = (H > 1) ? (float)((vf1 - vf0) / (double)(H - 1)) : 1.0f;
    g.div_mode = div_mode;
    return g;
}

// range = torch.norm(xyz, dim=1) on CPU == sqrt((x*x + y*y) + z*z), no FMA contraction
// (probed bit-exact against torch 2.11 CPU on 2e5 random points).
__device__ __forceinline__ float range3(float x, float y, float z) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}
__device__ __forceinline__ float range2(float x, float y) {
    return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}

// (u, v) with the reference's op order: src/utility/projection.py:21-31.
__device__ __forceinline__ void pixel_coords(const GridParams& g, float x, float y, float z,
                                             float& u, float& v) {
    float au = __fsub_rn(atan2f(y, x), g.hf0);
    float av = __fsub_rn(atan2f(z, range2(x, y)), g.vf0);
    if (g.div_mode == 0) {
        u = __fmul_rn(__fdiv_rn(au, g.hspan), g.wm1);
        v = __fmul_rn(__fdiv_rn(av, g.vspan), g.hm1);
    } else {
        u = __fmul_rn(__fmul_rn(au, g.hinv), g.wm1);
        v = __fmul_rn(__fmul_rn(av, g.vinv), g.hm1);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace delora

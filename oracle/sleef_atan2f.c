/* TEST INFRASTRUCTURE (oracle): scalar C restatement of SLEEF's `Sleef_atan2f*_u10` (FMA builds:
 * avx2 / avx512f), the routine behind `torch.atan2` on CPU float32 tensors
 * (ATen Vectorized<float>::atan2 -> Sleef_atan2f8_u10 / Sleef_atan2f16_u10), which the reference
 * calls in src/utility/projection.py:23,27.  SLEEF is a third-party dependency of torch and is not
 * vendored in the reference; the algorithm is restated from its published source (sleefsimdsp.c:
 * xatan2f_u1 / atan2kf_u1, df.h double-float helpers) and was checked instruction by instruction
 * against the disassembly of this image's libtorch_cpu.so (constants read from .rodata).
 * Pinned: bit-identical to libtorch's Sleef_atan2f8_u10avx2 on 8e6 random pairs and to
 * torch.atan2 (tests/test_oracle_golden.py::test_sleef_restatement_matches_torch_atan2).
 * The CUDA twin is `sleef_atan2f_u10` in delora_b200/csrc/common.cuh (same operations, IEEE
 * round-to-nearest intrinsics), which makes the GPU projection bit-identical to the CPU reference.
 *
 * Build (tests do this): gcc -O2 -shared -fPIC -mfma -ffp-contract=off oracle/sleef_atan2f.c -o oracle/_ref/libsleef_atan2f.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } f2;

static inline f2 dfdiv(f2 n, f2 d) {
    float t = 1.0f / d.x;
    float s = n.x * t;
    float u = fmaf(t, n.x, -s);
    float v = fmaf(-d.y, t, fmaf(-d.x, t, 1.0f));
    f2 r = {s, fmaf(s, v, fmaf(n.y, t, u))};
    return r;
}
static inline f2 dfsqu(f2 x) { float s = x.x * x.x; f2 r = {s, fmaf(x.x + x.x, x.y, fmaf(x.x, x.x, -s))}; return r; }
static inline f2 dfnorm(f2 t) { float s = t.x + t.y; f2 r = {s, (t.x - s) + t.y}; return r; }
static inline f2 dfmul(f2 x, f2 y) { float s = x.x * y.x; f2 r = {s, fmaf(x.x, y.y, fmaf(x.y, y.x, fmaf(x.x, y.x, -s)))}; return r; }
static inline f2 dfmulf(f2 x, float y) { float s = x.x * y; f2 r = {s, fmaf(x.y, y, fmaf(x.x, y, -s))}; return r; }
static inline f2 dfadd_ff(float x, float y) { float s = x + y; f2 r = {s, (x - s) + y}; return r; }
static inline f2 dfadd_ff2(float x, f2 y) { float s = x + y.x; f2 r = {s, ((x - s) + y.x) + y.y}; return r; }
static inline f2 dfadd_22(f2 x, f2 y) { float s = x.x + y.x; f2 r = {s, (((x.x - s) + y.x) + x.y) + y.y}; return r; }

static inline float mulsign(float x, float y) {   /* x with its sign flipped if y is negative (bit op) */
    uint32_t a, b; memcpy(&a, &x, 4); memcpy(&b, &y, 4); a ^= (b & 0x80000000u); memcpy(&x, &a, 4); return x;
}

float delora_sleef_atan2f_u10(float y0, float x0) {
    float x = x0, y = y0;
    if (fabsf(x) < 2.9387372783541830947e-39f) { x *= 16777216.0f; y *= 16777216.0f; }
    /* atan2kf_u1(|y|, x) */
    f2 yy = {fabsf(y), 0.0f}, xx = {x, 0.0f};
    int q = (xx.x < 0) ? -2 : 0;
    if (xx.x < 0) { xx.x = -xx.x; xx.y = -xx.y; }
    int p = xx.x < yy.x;
    if (p) q += 1;
    f2 s = p ? (f2){-xx.x, -xx.y} : yy;
    f2 t = p ? yy : xx;
    s = dfdiv(s, t);
    t = dfnorm(dfsqu(s));
    float u = -0.00176397908944636583328247f;
    u = fmaf(u, t.x, 0.0107900900766253471374512f);
    u = fmaf(u, t.x, -0.0309564601629972457885742f);
    u = fmaf(u, t.x, 0.0577365085482597351074219f);
    u = fmaf(u, t.x, -0.0838950723409652709960938f);
    u = fmaf(u, t.x, 0.109463557600975036621094f);
    u = fmaf(u, t.x, -0.142626821994781494140625f);
    u = fmaf(u, t.x, 0.199983194470405578613281f);
    t = dfmul(t, dfadd_ff(-0.333332866430282592773438f, u * t.x));
    t = dfmul(s, dfadd_ff2(1.0f, t));
    t = dfadd_22(dfmulf((f2){1.5707963705062866211f, -4.3711388286737928865e-08f}, (float)q), t);
    float r = t.x + t.y;
    /* xatan2f_u1 post-processing */
    r = mulsign(r, x);
    const float pio2 = 1.5707963267948966f, pio4 = 0.78539816339744831f, pi = 3.14159265358979323846f;
    if (isinf(x) || x == 0.0f) r = pio2 - (isinf(x) ? mulsign(pio2, x) : 0.0f);
    if (isinf(y)) r = pio2 - (isinf(x) ? mulsign(pio4, x) : 0.0f);
    if (y == 0.0f) r = signbit(x) ? pi : 0.0f;
    if (isnan(x) || isnan(y)) return NAN;
    return mulsign(r, y);
}

void delora_sleef_atan2f_array(const float* y, const float* x, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = delora_sleef_atan2f_u10(y[i], x[i]);
}

"""Development check of the row-block convolution kernel (csrc/conv_rows.cu) on a B200:
new kernel vs the first-generation kernel (csrc/conv_tc.cu) on identical inputs, plus CUDA-event timings.
Usage: python scripts/gpu_conv_rows_check.py [quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from delora_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
L = _lib.lib()


def rand_padded(b, h, w, c, gen, scale=0.5):
    x = ops.padded_nhwc_zeros(b, h, w, c, DEV)
    x[:, 1:h + 1, 1:w + 1] = (torch.randn(b, h, w, c, device=DEV, generator=gen) * scale).to(torch.bfloat16)
    x[:, 1:h + 1, 0] = x[:, 1:h + 1, w]
    x[:, 1:h + 1, w + 1] = x[:, 1:h + 1, 1]
    return x


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check_fprop(b, cin, cout, h, w, k, act, use_res, time_it=True):
    gen = torch.Generator(device=DEV).manual_seed(cin + cout + h + w)
    x = rand_padded(b, h, w, cin, gen)
    wt = (torch.randn(cout, k * k, cin, device=DEV, generator=gen) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    res = rand_padded(b, h, w, cout, gen) if use_res else None
    saved = rand_padded(b, h, w, cout, gen, 0.3) if act >= 3 else None
    L.delora_conv_select_kernel(0)
    y_old = ops.conv2d_fprop(x, wt, h, w, k, (1, 1), act, res, saved=saved)
    L.delora_conv_select_kernel(1)
    y_new = ops.conv2d_fprop(x, wt, h, w, k, (1, 1), act, res, saved=saved)
    torch.cuda.synchronize()
    d = (y_new.float() - y_old.float()).abs()
    ref = y_old.float().abs().max().item()
    bad = (d > 2e-2 * max(1.0, ref)).sum().item()
    msg = f"fprop B{b} {cin}->{cout} {h}x{w} k{k} act{act} res{int(use_res)}: maxdiff {d.max().item():.4g} (ref max {ref:.3g}) bad {bad}"
    if time_it:
        out = ops.padded_nhwc_zeros(b, h, w, cout, DEV)
        L.delora_conv_select_kernel(0)
        t_old = timed(lambda: ops.conv2d_fprop(x, wt, h, w, k, (1, 1), act, res, out, saved))
        L.delora_conv_select_kernel(2)
        t_one = timed(lambda: ops.conv2d_fprop(x, wt, h, w, k, (1, 1), act, res, out, saved))
        L.delora_conv_select_kernel(1)
        t_new = timed(lambda: ops.conv2d_fprop(x, wt, h, w, k, (1, 1), act, res, out, saved))
        fl = 2.0 * b * h * w * cout * cin * k * k
        msg += (f" | old {t_old * 1e3:.1f} us ({fl / t_old / 1e9:.0f} TF/s)  rows-1cta {t_one * 1e3:.1f} us "
                f"({fl / t_one / 1e9:.0f})  rows-pairs {t_new * 1e3:.1f} us ({fl / t_new / 1e9:.0f} TF/s)")
    print(("OK   " if bad == 0 else "FAIL ") + msg, flush=True)
    return bad == 0


def check_dgrad(b, cin, cout, h, w, stride, act, use_res, time_it=True):
    """cin/cout/h/w of the FORWARD conv (input h x w, cin channels)."""
    gen = torch.Generator(device=DEV).manual_seed(cin + cout + h + w + 1)
    ho, wo = ops.conv_out_size(h, stride[0]), ops.conv_out_size(w, stride[1])
    dz = rand_padded(b, ho, wo, cout, gen)
    wf = (torch.randn(cin, 9, cout, device=DEV, generator=gen) / (cout * 9) ** 0.5).to(torch.bfloat16)
    res = rand_padded(b, h, w, cin, gen) if use_res else None
    saved = rand_padded(b, h, w, cin, gen, 0.3) if act >= 3 else None
    L.delora_conv_select_kernel(0)

    def old(out=None):
        src = dz if stride == (1, 1) else ops.zero_upsample(dz, ho, wo, stride, None, (h, w))
        return ops.conv2d_fprop(src, wf, h, w, 3, (1, 1), act, res, out, saved)
    y_old = old()
    L.delora_conv_select_kernel(1)
    y_new = ops.conv2d_dgrad(dz, wf, h, w, stride, act, res, None, saved)
    L.delora_conv_select_kernel(0)
    torch.cuda.synchronize()
    d = (y_new.float() - y_old.float()).abs()
    ref = y_old.float().abs().max().item()
    bad = (d > 2e-2 * max(1.0, ref)).sum().item()
    msg = f"dgrad B{b} fwd {cin}->{cout} in {h}x{w} s{stride} act{act} res{int(use_res)}: maxdiff {d.max().item():.4g} (ref max {ref:.3g}) bad {bad}"
    if time_it:
        out = ops.padded_nhwc_zeros(b, h, w, cin, DEV)
        t_old = timed(lambda: old(out))
        L.delora_conv_select_kernel(2)
        t_one = timed(lambda: ops.conv2d_dgrad(dz, wf, h, w, stride, act, res, out, saved)) if cin % 128 == 0 else float("nan")
        L.delora_conv_select_kernel(1)
        t_new = timed(lambda: ops.conv2d_dgrad(dz, wf, h, w, stride, act, res, out, saved))
        fl = 2.0 * b * ho * wo * cout * cin * 9
        msg += (f" | old {t_old * 1e3:.1f} us  rows-1cta {t_one * 1e3:.1f} us  rows-pairs {t_new * 1e3:.1f} us "
                f"({fl / t_new / 1e9:.0f} TF/s useful)")
    L.delora_conv_select_kernel(1)
    print(("OK   " if bad == 0 else "FAIL ") + msg, flush=True)
    return bad == 0


def check_wgrad(b, cin, cout, h, w, stride, time_it=True):
    gen = torch.Generator(device=DEV).manual_seed(cin + cout + h + w + 2)
    ho, wo = ops.conv_out_size(h, stride[0]), ops.conv_out_size(w, stride[1])
    x = rand_padded(b, h, w, cin, gen)
    dz = rand_padded(b, ho, wo, cout, gen)
    L.delora_conv_select_kernel(0)
    g_old = ops.conv2d_wgrad(x, dz, h, w, 3, stride).clone()
    L.delora_conv_select_kernel(1)
    g_new = ops.conv2d_wgrad(x, dz, h, w, 3, stride).clone()
    torch.cuda.synchronize()
    d = (g_new - g_old).abs().max().item()
    ref = g_old.abs().max().item()
    ok = d <= 1e-3 * ref
    msg = f"wgrad B{b} {cin}->{cout} in {h}x{w} s{stride}: maxdiff {d:.4g} (ref max {ref:.4g})"
    if time_it:
        L.delora_conv_select_kernel(0)
        t_old = timed(lambda: ops.conv2d_wgrad(x, dz, h, w, 3, stride))
        L.delora_conv_select_kernel(1)
        t_new = timed(lambda: ops.conv2d_wgrad(x, dz, h, w, 3, stride))
        fl = 2.0 * b * ho * wo * cout * cin * 9
        msg += f" | old {t_old * 1e3:.1f} us ({fl / t_old / 1e9:.0f} TF/s)  new {t_new * 1e3:.1f} us ({fl / t_new / 1e9:.0f} TF/s)"
    L.delora_conv_select_kernel(1)
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return ok


def check_stem(b, h, w, time_it=True):
    import torch.nn.functional as F
    gen = torch.Generator(device=DEV).manual_seed(h + w)
    img1 = torch.randn(b, 4, h, w, device=DEV, generator=gen) * 5.0
    img2 = torch.randn(b, 4, h, w, device=DEV, generator=gen) * 5.0
    wt = torch.randn(64, 8, 3, 3, device=DEV, generator=gen) / 72 ** 0.5
    x16 = ops.images_to_nhwc16(img1, img2)
    wst = torch.empty((3, 64, 64), dtype=torch.bfloat16, device=DEV)
    ops.stem_weight_prep(wt, wst)
    y = ops.stem_fprop(x16, wst, h, w, ops.ACT_TANH)
    x = torch.cat([img1, img2], 1)                 # the stem carries bf16(x) + bf16(x - bf16(x)): ~fp32 input precision
    ref = torch.tanh(F.conv2d(F.pad(x, (1, 1, 0, 0), mode="circular"), wt.to(torch.bfloat16).float(), stride=(1, 2), padding=(1, 0)))
    got = ops.nhwc_to_nchw(y, h, w // 2)
    e1 = (got - ref).abs().max().item()
    yf = y.float()
    halo_ok = torch.equal(yf[:, 1:-1, 0], yf[:, 1:-1, w // 2]) and torch.equal(yf[:, 1:-1, w // 2 + 1], yf[:, 1:-1, 1])
    # wgrad vs autograd
    xr = x.clone().requires_grad_(False)
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    dz = (torch.randn(b, 64, h, w // 2, device=DEV, generator=gen) * 0.5).to(torch.bfloat16).float()
    F.conv2d(F.pad(xr, (1, 1, 0, 0), mode="circular"), wr, stride=(1, 2), padding=(1, 0)).backward(dz)
    dzp = ops.padded_nhwc_zeros(b, h, w // 2, 64, DEV)
    dzp[:, 1:h + 1, 1:w // 2 + 1] = dz.permute(0, 2, 3, 1).to(torch.bfloat16)
    dzp[:, 1:h + 1, 0] = dzp[:, 1:h + 1, w // 2]
    dzp[:, 1:h + 1, w // 2 + 1] = dzp[:, 1:h + 1, 1]
    dw = ops.stem_wgrad(x16, dzp, h, w, 8)
    e2 = (dw - wr.grad).abs().max().item() / wr.grad.abs().max().item()
    ok = e1 < 8e-3 and halo_ok and e2 < 1e-3
    msg = f"stem B{b} {h}x{w}: fprop maxerr {e1:.4g} halo {halo_ok} wgrad rel {e2:.3g}"
    if time_it:
        out = ops.padded_nhwc_zeros(b, h, w // 2, 64, DEV)
        t_f = timed(lambda: ops.stem_fprop(x16, wst, h, w, ops.ACT_TANH, out))
        t_w = timed(lambda: ops.stem_wgrad(x16, dzp, h, w, 8))
        t_i = timed(lambda: ops.images_to_nhwc16(img1, img2))
        msg += f" | to_nhwc16 {t_i * 1e3:.1f} us  fprop {t_f * 1e3:.1f} us  wgrad {t_w * 1e3:.1f} us"
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return ok


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ok = True
    t0 = time.time()
    # small correctness cases first (ragged widths / heights, 1x1, both activations, backward epilogues)
    ok &= check_fprop(1, 64, 128, 8, 128, 3, 0, False, False)
    ok &= check_fprop(2, 128, 128, 16, 128, 3, 1, True, False)
    ok &= check_fprop(1, 128, 256, 7, 90, 3, 2, True, False)
    ok &= check_fprop(1, 256, 256, 5, 45, 3, 3, True, False)
    ok &= check_fprop(2, 512, 512, 32, 23, 3, 4, False, False)
    ok &= check_fprop(1, 512, 512, 32, 64, 3, 2, True, False)
    ok &= check_fprop(1, 256, 128, 9, 200, 1, 0, False, False)
    ok &= check_dgrad(1, 128, 256, 8, 128, (1, 2), 3, True, False)
    ok &= check_dgrad(1, 256, 512, 16, 64, (2, 2), 3, True, False)
    ok &= check_dgrad(1, 256, 512, 15, 46, (2, 2), 0, False, False)
    ok &= check_dgrad(1, 128, 128, 8, 96, (1, 1), 4, True, False)
    ok &= check_fprop(1, 64, 256, 3, 40, 3, 2, True, False)
    ok &= check_fprop(3, 128, 512, 9, 23, 3, 1, True, False)
    ok &= check_dgrad(2, 256, 128, 10, 180, (1, 2), 3, True, False)
    ok &= check_wgrad(1, 64, 64, 8, 128, (1, 1), False)
    ok &= check_wgrad(2, 64, 64, 6, 180, (1, 1), False)
    ok &= check_wgrad(1, 128, 128, 8, 128, (1, 1), False)
    ok &= check_wgrad(2, 64, 128, 16, 256, (1, 2), False)
    ok &= check_wgrad(1, 256, 512, 16, 128, (2, 2), False)
    ok &= check_wgrad(1, 256, 512, 64, 45, (2, 2), False)
    ok &= check_wgrad(2, 512, 512, 32, 23, (1, 1), False)
    ok &= check_wgrad(1, 128, 256, 64, 90, (1, 2), False)
    ok &= check_fprop(1, 64, 64, 8, 128, 3, 2, True, False)
    ok &= check_fprop(2, 64, 64, 5, 200, 3, 1, True, False)
    ok &= check_fprop(1, 128, 128, 3, 130, 3, 3, True, False)
    ok &= check_fprop(1, 128, 64, 7, 256, 3, 4, False, False)
    ok &= check_dgrad(1, 64, 128, 6, 512, (1, 2), 3, True, False)
    ok &= check_dgrad(2, 128, 64, 9, 256, (1, 1), 0, False, False)
    ok &= check_stem(1, 8, 256, False)
    ok &= check_stem(2, 16, 180, False)
    ok &= check_stem(1, 64, 720, False)
    print(f"-- small cases done in {time.time() - t0:.1f} s, ok={ok}", flush=True)
    if not quick:
        # bench shapes (B = 16, 64x2048 image): L2 128ch @64x256, L3 256ch @64x128, L4 512ch @32x64
        check_fprop(16, 64, 64, 64, 512, 3, 2, True)
        check_fprop(16, 64, 64, 64, 512, 3, 3, True)
        check_dgrad(16, 64, 128, 64, 512, (1, 2), 3, True)
        check_fprop(16, 128, 128, 64, 256, 3, 2, True)
        check_fprop(16, 256, 256, 64, 128, 3, 2, True)
        check_fprop(16, 512, 512, 32, 64, 3, 2, True)
        check_fprop(16, 128, 128, 64, 256, 3, 3, True)
        check_dgrad(16, 128, 256, 64, 256, (1, 2), 3, True)
        check_dgrad(16, 256, 512, 64, 128, (2, 2), 3, True)
    if not quick:
        check_stem(16, 64, 2048)
        check_wgrad(16, 64, 64, 64, 512, (1, 1))
        check_wgrad(16, 128, 128, 64, 256, (1, 1))
        check_wgrad(16, 64, 128, 64, 512, (1, 2))
        check_wgrad(16, 256, 256, 64, 128, (1, 1))
        check_wgrad(16, 128, 256, 64, 256, (1, 2))
        check_wgrad(16, 512, 512, 32, 64, (1, 1))
        check_wgrad(16, 256, 512, 64, 128, (2, 2))
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

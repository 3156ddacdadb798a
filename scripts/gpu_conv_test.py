"""Parity of the tcgen05 implicit-GEMM convolution against torch (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from delora_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def to_padded_nhwc(x):            # x [B,C,H,W] fp32 -> [B,H+2,W+2,C] bf16 (circular W, zero H)
    xp = F.pad(x, (1, 1, 0, 0), mode="circular")
    xp = F.pad(xp, (0, 0, 1, 1))
    return xp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def ref_conv(x, w, ksize, stride, act, residual):
    xb = x.to(torch.bfloat16).float()
    wb = w.to(torch.bfloat16).float()
    if ksize == 3:
        y = F.conv2d(F.pad(xb, (1, 1, 0, 0), mode="circular"), wb, stride=stride, padding=(1, 0))
    else:
        y = F.conv2d(xb, wb, stride=stride)
    if residual is not None:
        y = y + residual.to(torch.bfloat16).float()
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.tanh(y)
    return y


def run(b, cin, cout, h, w, ksize, stride, act, use_res):
    x = torch.randn(b, cin, h, w, device=dev) * 0.5
    wt = torch.randn(cout, cin, ksize, ksize, device=dev) * (1.0 / (cin * ksize * ksize) ** 0.5)
    ho, wo = h // stride[0], w // stride[1]
    res = torch.randn(b, cout, ho, wo, device=dev) * 0.5 if use_res else None
    xn = to_padded_nhwc(x)
    wn = wt.permute(0, 2, 3, 1).reshape(cout, ksize * ksize, cin).contiguous().to(torch.bfloat16)
    rn = to_padded_nhwc(res) if use_res else None
    y = ops.conv2d_fprop(xn, wn, h, w, ksize, stride, act, rn)
    torch.cuda.synchronize()
    got = ops.nhwc_to_nchw(y, ho, wo)
    ref = ref_conv(x, wt, ksize, stride, act, res)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    # halo columns must be the circular copies, halo rows zero
    yf = y.float()
    halo = (yf[:, 1:-1, 0] - yf[:, 1:-1, wo]).abs().max().item() + (yf[:, 1:-1, wo + 1] - yf[:, 1:-1, 1]).abs().max().item()
    zr = yf[:, 0].abs().max().item() + yf[:, -1].abs().max().item()
    ok = err <= 2e-2 * max(scale, 1.0) and halo == 0.0 and zr == 0.0
    print(f"B={b} Cin={cin} Cout={cout} {h}x{w} k={ksize} s={stride} act={act} res={use_res}: max|err|={err:.3e} "
          f"(ref max {scale:.2f}) halo={halo} zero_rows={zr} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


cases = [
    (1, 64, 64, 8, 128, 3, (1, 1), 0, False),
    (2, 64, 64, 16, 256, 3, (1, 1), 2, True),
    (2, 64, 128, 16, 256, 3, (1, 2), 2, False),
    (2, 64, 128, 16, 256, 1, (1, 2), 0, False),
    (2, 128, 128, 16, 128, 3, (1, 1), 1, True),
    (2, 256, 512, 16, 128, 3, (2, 2), 2, False),
    (2, 256, 512, 16, 128, 1, (2, 2), 0, False),
    (1, 512, 512, 32, 64, 3, (1, 1), 2, True),
    (2, 64, 64, 64, 1024, 3, (1, 2), 2, False),
]
allok = True
for c in cases:
    try:
        allok &= run(*c)
    except Exception as e:
        print("EXC", c, e, flush=True)
        allok = False
print("ALL OK" if allok else "SOME FAILED")

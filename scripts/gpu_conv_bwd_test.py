"""dgrad / wgrad on tcgen05 vs torch autograd (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from delora_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def to_padded_nhwc(x):
    xp = F.pad(F.pad(x, (1, 1, 0, 0), mode="circular"), (0, 0, 1, 1))
    return xp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def conv_ref(x, w, k, stride):
    if k == 3:
        return F.conv2d(F.pad(x, (1, 1, 0, 0), mode="circular"), w, stride=stride, padding=(1, 0))
    return F.conv2d(x, w, stride=stride)


def run(b, cin, cout, h, w, k, stride):
    x = (torch.randn(b, cin, h, w, device=dev) * 0.5).to(torch.bfloat16).float().requires_grad_(True)
    wt = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16).float().requires_grad_(True)
    ho, wo = h // stride[0], w // stride[1]
    dz = (torch.randn(b, cout, ho, wo, device=dev) * 0.5).to(torch.bfloat16).float()
    y = conv_ref(x, wt, k, stride)
    y.backward(dz)
    # ---- wgrad
    dw = ops.conv2d_wgrad(to_padded_nhwc(x.detach()), to_padded_nhwc(dz), h, w, k, stride)
    torch.cuda.synchronize()
    e_w = (dw - wt.grad).abs().max().item() / wt.grad.abs().max().item()
    # ---- dgrad: conv of the (zero-upsampled) output gradient with the flipped, transposed filter
    wf = wt.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(cin, k * k, cout).contiguous().to(torch.bfloat16)
    dzn = to_padded_nhwc(dz)
    if stride != (1, 1):
        dzn = ops.zero_upsample(dzn, ho, wo, stride, out_hw=(h, w))
    dx = ops.conv2d_fprop(dzn, wf, h, w, k, (1, 1), ops.ACT_NONE)
    torch.cuda.synchronize()
    got = ops.nhwc_to_nchw(dx, h, w)
    e_x = (got - x.grad).abs().max().item() / x.grad.abs().max().item()
    ok = e_w < 1e-2 and e_x < 1.5e-2
    print(f"B={b} {cin}->{cout} {h}x{w} k{k} s{stride}: wgrad rel err {e_w:.2e}, dgrad rel err {e_x:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


cases = [
    (2, 64, 64, 8, 128, 3, (1, 1)),
    (2, 64, 128, 16, 256, 3, (1, 2)),
    (2, 64, 128, 16, 256, 1, (1, 2)),
    (2, 128, 128, 16, 128, 3, (1, 1)),
    (2, 256, 512, 16, 128, 3, (2, 2)),
    (2, 256, 512, 16, 128, 1, (2, 2)),
    (1, 512, 512, 32, 64, 3, (1, 1)),
]
allok = True
for c in cases:
    try:
        allok &= run(*c)
    except Exception as e:
        print("EXC", c, repr(e)[:300], flush=True)
        allok = False
print("ALL OK" if allok else "SOME FAILED")

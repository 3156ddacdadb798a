"""Run ONE convolution configuration a few times (for ncu): python scripts/gpu_conv_one.py MODE B CIN COUT H W [iters]
MODE: 0 first-generation kernel, 2 row-block single CTA, 1 row-block CTA pairs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from delora_b200 import _lib, ops  # noqa: E402

mode, b, cin, cout, h, w = [int(a) for a in sys.argv[1:7]]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 3
L = _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(1)
x = ops.padded_nhwc_zeros(b, h, w, cin, "cuda")
x[:, 1:h + 1] = (torch.randn(b, h, w + 2, cin, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
res = ops.padded_nhwc_zeros(b, h, w, cout, "cuda")
res[:, 1:h + 1] = (torch.randn(b, h, w + 2, cout, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)
wt = (torch.randn(cout, 9, cin, device="cuda", generator=gen) / (cin * 9) ** 0.5).to(torch.bfloat16)
out = ops.padded_nhwc_zeros(b, h, w, cout, "cuda")
L.delora_conv_select_kernel(mode)
for _ in range(iters):
    ops.conv2d_fprop(x, wt, h, w, 3, (1, 1), ops.ACT_TANH, res, out)
torch.cuda.synchronize()
print("done", mode, b, cin, cout, h, w)

"""Time the stem's max-pool backward (B = 16, 64 x 1024 x 64, the training-step shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import ops

L = ops._lib.lib()
b, h, w, c = 16, 64, 1024, 64
wo = w // 2
y0 = (torch.randn((b, h + 2, w + 2, c), device="cuda") * 1.5).half()
pooled = ops.padded_nhwc_zeros(b, h, wo, c, "cuda")
idx = torch.empty((b, h, wo, c), dtype=torch.uint8, device="cuda")
L.delora_maxpool_w_idx_nhwc_bf16(y0.data_ptr(), b, h, w, c, pooled.data_ptr(), idx.data_ptr(), 2, 1, ops._stream())
dy = torch.randn((b, h + 2, wo + 2, c), device="cuda").bfloat16()
dz = torch.empty((b, h + 2, w + 2, c), dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.delora_maxpool_w_bwd_nhwc_bf16(dy.data_ptr(), idx.data_ptr(), y0.data_ptr(), b, h, w, c, 6, dz.data_ptr(), 1, ops._stream())
    e1.record(); torch.cuda.synchronize()
    if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
byt = dy.numel() * 2 + idx.numel() + y0.numel() * 2 + dz.numel() * 2
print(f"maxpool_bwd: median {ts[len(ts)//2]:.1f} us, min {ts[0]:.1f} us; algorithmic {byt/1e6:.1f} MB -> {byt/ts[len(ts)//2]/1e3:.0f} GB/s")

"""tcgen05 encoder vs the torch (cuDNN, fp32) model + timing (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic
from delora_b200.models.model import OdometryModel
from delora_b200.models.tc_encoder import TensorCoreEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
H = 64
cfg = synthetic.fov_config(h=H, w=W, device="cuda")
cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
            "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False})
torch.manual_seed(0)
model = OdometryModel(cfg).cuda().eval()
enc = TensorCoreEncoder(model)
g = torch.Generator(device="cuda").manual_seed(1)
# range-image-like inputs: xyz in metres + range, many zeros
img1 = torch.randn(B, 4, H, W, device="cuda", generator=g) * 5.0
img2 = torch.randn(B, 4, H, W, device="cuda", generator=g) * 5.0
with torch.no_grad():
    ref_feats = model.forward_features(image_1=img1[:2], image_2=img2[:2])
    t_ref, q_ref = model(image_1=img1[:2], image_2=img2[:2])
feats = enc.features(img1[:2].contiguous(), img2[:2].contiguous())
from delora_b200 import ops
for i, ((x, h, w), r) in enumerate(zip(feats, ref_feats[:4])):
    got = ops.nhwc_to_nchw(x, h, w)
    err = (got - r).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), r.flatten(), dim=0).item()
    print(f"x{i+1}: shape {tuple(r.shape)} max|err|={err:.3e} (ref max {r.abs().max().item():.2f}) cosine={cos:.6f}")
t, q = enc.forward(img1[:2].contiguous(), img2[:2].contiguous())
print("translation err", (t - t_ref).abs().max().item(), "quaternion err", (q - q_ref).abs().max().item(), "ref", t_ref[0].tolist())

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
gflop = 96.17 * (W / 2048.0) * B
ms = timeit(lambda: enc.features(img1, img2))
print(f"tcgen05 encoder fwd B={B} 64x{W}: {ms:.3f} ms -> {gflop/ms:.1f} TFLOP/s (dense bf16)")
with torch.no_grad():
    ms_t = timeit(lambda: model.forward_features(image_1=img1, image_2=img2))
    print(f"torch fp32 (cuDNN) fwd: {ms_t:.3f} ms -> {gflop/ms_t:.1f} TFLOP/s")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ms_b = timeit(lambda: model.forward_features(image_1=img1, image_2=img2))
    print(f"torch bf16 autocast (cuDNN) fwd: {ms_b:.3f} ms -> {gflop/ms_b:.1f} TFLOP/s")

# ---- per-layer timing
print("per-layer (B=%d):" % B)
import itertools
def time_conv(cin, cout, h, w, k, stride, act, res, iters=20):
    x = torch.randn(B, h + 2, w + 2, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, k * k, cin, device="cuda") * 0.05).to(torch.bfloat16)
    ho, wo = h // stride[0], w // stride[1]
    out = ops.padded_nhwc_zeros(B, ho, wo, cout, "cuda")
    r = torch.randn(B, ho + 2, wo + 2, cout, device="cuda").to(torch.bfloat16) if res else None
    ms = timeit(lambda: ops.conv2d_fprop(x, wt, h, w, k, stride, act, r, out), iters)
    fl = 2.0 * B * ho * wo * cout * cin * k * k
    print(f"  conv {cin:3d}->{cout:3d} in {h}x{w} k{k} s{stride}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  (M tiles {B*ho*wo//128}, BN {128 if cout%128==0 else 64})")
    return ms
tot = 0
tot += time_conv(64, 64, H, W, 3, (1, 2), 2, False)
for _ in range(1): pass
tot += 4 * time_conv(64, 64, H, W // 4, 3, (1, 1), 2, True)
tot += time_conv(64, 128, H, W // 4, 3, (1, 2), 2, False) + time_conv(64, 128, H, W // 4, 1, (1, 2), 0, False)
tot += 3 * time_conv(128, 128, H, W // 8, 3, (1, 1), 2, True)
tot += time_conv(128, 256, H, W // 8, 3, (1, 2), 2, False) + time_conv(128, 256, H, W // 8, 1, (1, 2), 0, False)
tot += 3 * time_conv(256, 256, H, W // 16, 3, (1, 1), 2, True)
tot += time_conv(256, 512, H, W // 16, 3, (2, 2), 2, False) + time_conv(256, 512, H, W // 16, 1, (2, 2), 0, False)
tot += 3 * time_conv(512, 512, H // 2, W // 32, 3, (1, 1), 2, True)
print(f"sum of conv layers: {tot:.3f} ms")
ms = timeit(lambda: ops.images_to_nhwc(img1, img2, 64)); print(f"images_to_nhwc: {ms*1e3:.1f} us")
xs = ops.padded_nhwc_zeros(B, H, W // 2, 64, "cuda")
ms = timeit(lambda: ops.maxpool_w(xs, H, W // 2)); print(f"maxpool (incl. zeros alloc): {ms*1e3:.1f} us")

"""tcgen05 encoder training path (fwd + bwd) vs torch fp32 autograd (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from delora_b200 import synthetic
from delora_b200.models.model import OdometryModel
from delora_b200.models.tc_encoder import TensorCoreEncoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W = int(sys.argv[2]) if len(sys.argv) > 2 else 512
H = 64
cfg = synthetic.fov_config(h=H, w=W, device="cuda")
cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
            "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False})
torch.manual_seed(0)
model = OdometryModel(cfg).cuda()
enc = TensorCoreEncoder(model)
g = torch.Generator(device="cuda").manual_seed(1)
img1 = torch.randn(B, 4, H, W, device="cuda", generator=g) * 5.0
img2 = torch.randn(B, 4, H, W, device="cuda", generator=g) * 5.0
wsel = torch.randn(B, 512, device="cuda", generator=g)

# reference: torch fp32
model.zero_grad()
feats = model.forward_features(image_1=img1, image_2=img2)
pooled_ref = feats[3].mean(dim=(2, 3))
(pooled_ref * wsel).sum().backward()
ref_grads = [p.grad.clone() for p in enc.trunk_parameters()]
model.zero_grad()
pooled = enc.pooled_features(img1, img2)
(pooled * wsel).sum().backward()
torch.cuda.synchronize()
print("pooled: max err", (pooled - pooled_ref).abs().max().item(), "cos", F.cosine_similarity(pooled.flatten(), pooled_ref.flatten(), dim=0).item())
names = []
r = model.resnet
names.append("conv1")
for li in range(1, 5):
    for bi, blk in enumerate(getattr(r, f"layer{li}")):
        names += [f"layer{li}.{bi}.conv1", f"layer{li}.{bi}.conv2"] + ([f"layer{li}.{bi}.downsample"] if blk.downsample is not None else [])
worst = 1.0
for n, p, rg in zip(names, enc.trunk_parameters(), ref_grads):
    cos = F.cosine_similarity(p.grad.flatten(), rg.flatten(), dim=0).item()
    rel = ((p.grad - rg).norm() / rg.norm()).item()
    worst = min(worst, cos)
    print(f"  {n:24s} grad cosine {cos:.5f} rel L2 err {rel:.3e} |g| {rg.norm().item():.3e}")
print("worst cosine", worst, "OK" if worst > 0.99 else "FAIL")

def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def step_tc():
    model.zero_grad(set_to_none=True)
    (enc.pooled_features(img1, img2) * wsel).sum().backward()
def step_torch():
    model.zero_grad(set_to_none=True)
    (model.forward_features(image_1=img1, image_2=img2)[3].mean(dim=(2, 3)) * wsel).sum().backward()
def step_torch_bf16():
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        f = model.forward_features(image_1=img1, image_2=img2)[3]
    (f.float().mean(dim=(2, 3)) * wsel).sum().backward()
gflop = 3 * 96.17 * (W / 2048.0) * B
for name, fn in (("tcgen05 fwd+bwd", step_tc), ("torch fp32 fwd+bwd", step_torch), ("torch bf16 autocast fwd+bwd", step_torch_bf16)):
    ms = timeit(fn)
    print(f"{name}: {ms:.2f} ms  ({gflop/ms:.1f} TFLOP/s at 3x fwd flops)")

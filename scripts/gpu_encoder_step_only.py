"""One tcgen05 encoder training step (fwd + bwd), B=16 64x2048, for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic
from delora_b200.models.model import OdometryModel
from delora_b200.models.tc_encoder import TensorCoreEncoder
B, W, H = 16, 2048, 64
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
for _ in range(2):
    model.zero_grad(set_to_none=True)
    (enc.pooled_features(img1, img2) * wsel).sum().backward()
torch.cuda.synchronize()
print("done")

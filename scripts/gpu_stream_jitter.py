"""Where do the rare slow frames of the streaming leg come from?  Per-phase host timestamps of 600 frames."""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from delora_b200 import synthetic
from delora_b200.deploy.stream import OdometryStream
from delora_b200.models.model import OdometryModel
H, W = 64, 2048
cfg = synthetic.fov_config(h=H, w=W, device="cuda")
cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
            "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False,
            "use_tensor_core_encoder": True})
torch.manual_seed(0)
model = OdometryModel(cfg).cuda().eval()
frames = []
for i in range(4):
    s1, s2, _, _ = synthetic.make_pair(i, w_raw=2048)
    frames += [s1, s2]
n_max = max(f.shape[1] for f in frames)
st = OdometryStream(model, cfg, "kitti", n_max, use_cuda_graph=True)
for i in range(8):
    st.push(frames[i % 8])
gc.collect(); gc.freeze()
lat = []
for i in range(600):
    t0 = time.perf_counter()
    st.push(frames[i % 8])
    lat.append((time.perf_counter() - t0) * 1e3)
slow = [(i, round(x, 2)) for i, x in enumerate(lat) if x > 2.0]
srt = sorted(lat)
print("p50", srt[300], "p99", srt[594], "max", srt[-1], "slow frames", slow)
# same loop, device-side timing only (events around the graph replay)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
for i in range(300):
    ev[i][0].record(); st.graph.replay(); ev[i][1].record()
torch.cuda.synchronize()
d = sorted(a.elapsed_time(b) for a, b in ev)
print("device-only graph replay ms: p50", d[150], "p99", d[297], "max", d[-1])

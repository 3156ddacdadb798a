"""Error distribution of the CUDA normals vs the oracle (LAPACK) as a function of the eigen-gap: calibrates the
bars of tests/test_gpu_parity.py::test_normals_and_lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from delora_b200 import ops, synthetic  # noqa: E402
from oracle import delora_oracle as orc  # noqa: E402

torch.set_num_threads(8)
for (h, w, wraw, rings) in ((16, 180, 192, 16), (64, 720, 1875, 64), (64, 2048, 2048, 64), (64, 2250, 2048, 64), (64, 512, 600, 64), (16, 720, 800, 16)):
    vf = (-15.0, 15.0) if h == 16 else (-24.5, 2.0)
    cfg = synthetic.fov_config(h=h, w=w, vfov_deg=vf)
    s1, _, _, _ = synthetic.make_pair(3, w_raw=wraw, rings=rings, vfov_deg=vf)
    hf, vfr = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
    image = orc.project_to_img(s1[None], h, w, hf, vfr)[0]
    n_o, enough, locs, aux = orc.compute_normal_vectors(image, return_aux=True)
    nrm = ops.normals(image.cuda())
    valid = (image[0, 0] != 0) & (image[0, 1] != 0) & (image[0, 2] != 0)
    n_g = nrm[0].permute(1, 2, 0)[valid.cuda()].cpu()
    has = enough
    err = (n_g - n_o).norm(dim=1)[has]
    ev = aux["eigenvalues"]
    gap = ((ev[:, 1] - ev[:, 0]) / ev[:, 2].clamp_min(1e-30))
    q = torch.quantile(err, torch.tensor([0.5, 0.99, 0.999])).tolist()
    print(f"{h}x{w}: n={int(has.sum())} median={q[0]:.2e} p99={q[1]:.2e} p99.9={q[2]:.2e} max={err.max():.2e}; "
          f"mask equal={bool(torch.equal((n_g != 0).any(1), has))}; max(err*gap)={float((err * gap).max()):.2e}; "
          f"max err [gap>=1e-2]={float(err[gap >= 1e-2].max()):.2e} [gap>=1e-3]={float(err[gap >= 1e-3].max()):.2e} "
          f"min gap={float(gap.min()):.2e}", flush=True)

"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import hashlib
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def digest(t):
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


def unsort_uv(cloud, u_sorted, v_sorted):
    """Oracle (u, v) come in stable range order; scatter them back to the original point order."""
    rng = torch.norm(cloud[None][:, :3, :], dim=1)
    order = torch.argsort(rng, dim=1, stable=True)[0]
    u = torch.empty_like(u_sorted)
    v = torch.empty_like(v_sorted)
    u[order] = u_sorted
    v[order] = v_sorted
    return u, v, rng[0]


def case_inputs(meta):
    from delora_b200 import synthetic
    cfg = synthetic.fov_config(h=meta["H"], w=meta["W"], vfov_deg=tuple(meta["vfov_deg"]))
    pair = synthetic.make_pair(meta["pair_index"], w_raw=meta["w_raw"], rings=meta["rings"],
                               vfov_deg=tuple(meta["vfov_deg"]))
    return cfg, pair

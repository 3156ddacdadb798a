"""-m "not gpu": the oracle against the golden vectors the unmodified reference produced
(oracle/gen_golden.py).  Bit-exact for projection / normals / point lists / pair counts; losses to
1e-6 rel and the transform gradient to 1e-5 rel (autograd summation order is not pinned)."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, case_inputs, digest, unsort_uv
from oracle import delora_oracle as orc

CASES = ["small_16x180", "kitti_64x720", "kitti_64x2048"]


@pytest.fixture(scope="module", params=CASES)
def case(request, golden):
    meta = golden[request.param]
    cfg, (scan_1, scan_2, t_gt, t_pred) = case_inputs(meta)
    assert [digest(scan_1), digest(scan_2)] == meta["inputs_sha256"], "synthetic generator drifted"
    out = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg)
    return request.param, meta, cfg, (scan_1, scan_2, t_pred), out


def test_projection_images_bit_exact(case):
    name, meta, cfg, _, out = case
    assert digest(out["image_1"][0]) == meta["sha256"]["image_1"]
    assert digest(out["image_2"][0]) == meta["sha256"]["image_2"]


def test_projection_uv_and_indices(case):
    name, meta, cfg, (scan_1, _, _), _ = case
    h, w = meta["H"], meta["W"]
    image, u, v, idx, i2p = orc.project_to_img(scan_1[None], h, w, cfg["horizontal_field_of_view"],
                                               cfg["kitti"]["vertical_field_of_view"])
    uo, vo, _ = unsort_uv(scan_1, u[0], v[0])
    assert digest(uo) == meta["sha256"]["u_1"]
    assert digest(vo) == meta["sha256"]["v_1"]
    assert digest(idx) == meta["sha256"]["idx_1"]          # (range, index) order == stable sort
    assert idx.shape[0] == meta["K"][0]
    assert torch.equal(image[0][:, i2p[0, :, 0], i2p[0, :, 1]][:3], scan_1[:, idx])


def test_normals_and_lists_bit_exact(case):
    name, meta, _, _, out = case
    for k in ("1", "2"):
        assert digest(out["points_" + k]) == meta["sha256"]["points_" + k]
        assert digest(out["normals_" + k]) == meta["sha256"]["normals_" + k]
    assert out["points_1"].shape[0] == meta["P"][0]


def test_losses_and_gradient(case):
    name, meta, _, _, out = case
    assert out["num_pairs"] == meta["num_pairs"]
    assert out["loss_po2pl"] == pytest.approx(meta["loss_po2pl"], rel=1e-6)
    assert out["loss_pl2pl"] == pytest.approx(meta["loss_pl2pl"], rel=1e-6)
    g = np.asarray(meta["grad_T"])
    assert np.abs(out["grad_T"].numpy() - g).max() <= 1e-5 * np.abs(g).max()


def test_small_case_full_tensors():
    z = np.load(os.path.join(GOLDEN, "small_16x180.npz"))
    import json
    meta = json.load(open(os.path.join(GOLDEN, "golden.json")))["small_16x180"]
    cfg, (scan_1, scan_2, _, t_pred) = case_inputs(meta)
    out = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg)
    assert np.array_equal(out["image_1"][0].numpy(), z["image_1"])
    assert np.array_equal(out["normals_2"].numpy(), z["normals_2"])
    assert np.array_equal(out["points_2"].numpy(), z["points_2"])
    # kept source points (plotting["scan_2_transformed"], icp_losses.py:153-156): same set, same order
    src = orc.transform_point_cloud(t_pred.view(1, 4, 4), out["points_2"].t()[None])[0]
    assert np.allclose(src[:, out["kept_mask"]].numpy(), z["kept_source_points"], rtol=0, atol=1e-6)


def test_nn_kdtree_equals_bruteforce():
    import json
    meta = json.load(open(os.path.join(GOLDEN, "golden.json")))["small_16x180"]
    cfg, (scan_1, scan_2, _, t_pred) = case_inputs(meta)
    out = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg, nn_method="brute")
    out2 = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg, nn_method="kdtree")
    assert torch.equal(out["nn_index"], out2["nn_index"])


@pytest.mark.parametrize("name", ["edge_16x180", "tie_16x512"])
def test_stress_clouds(name, golden):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = golden[name]
    from delora_b200 import synthetic
    cfg = synthetic.fov_config(h=meta["H"], w=meta["W"], vfov_deg=tuple(meta["vfov_deg"]))
    cloud = torch.from_numpy(z["cloud"])
    image, u, v, idx, _ = orc.project_to_img(cloud[None], meta["H"], meta["W"], cfg["horizontal_field_of_view"],
                                             cfg["kitti"]["vertical_field_of_view"])
    assert np.array_equal(image[0].numpy(), z["image"])           # pixel VALUES are tie-independent
    uo, vo, rng = unsort_uv(cloud, u[0], v[0])
    assert np.array_equal(uo.numpy(), z["u"]) and np.array_equal(vo.numpy(), z["v"])
    assert idx.shape[0] == meta["K"]
    if name.startswith("edge"):
        assert np.array_equal(idx.numpy(), z["idx"])
    else:
        # equal-range ties: the reference's unstable argsort picks arbitrarily; the chosen points
        # must be indistinguishable (same coordinates) and ours is the lowest index
        ref_idx = torch.from_numpy(z["idx"])
        assert torch.equal(cloud[:, idx], cloud[:, ref_idx])
        assert bool((idx <= ref_idx).all())


def test_quaternion_to_T(golden):
    z = np.load(os.path.join(GOLDEN, "quaternion.npz"))
    t = orc.transformation_matrix_quaternion(torch.from_numpy(z["translation"]), torch.from_numpy(z["quaternion"]))
    assert np.array_equal(t.numpy(), z["T"])


def test_sleef_restatement_matches_torch_atan2():
    """oracle/sleef_atan2f.c (scalar twin of the CUDA `sleef_atan2f_u10`) against torch.atan2 on CPU:
    bit-identical wherever torch runs its vectorised (SLEEF) path, i.e. everywhere except the < 32
    trailing elements torch hands to the scalar libm."""
    import ctypes
    import subprocess
    from helpers import ROOT
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)      # oracle/Makefile -> oracle/_ref/
    lib = os.path.join(ROOT, "oracle", "_ref", "libsleef_atan2f.so")
    L = ctypes.CDLL(lib)
    L.delora_sleef_atan2f_array.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_long]

    def mine(y, x):
        o = np.empty_like(y)
        L.delora_sleef_atan2f_array(y.ctypes.data, x.ctypes.data, o.ctypes.data, len(y))
        return o
    rng = np.random.default_rng(0)
    n = 32 * 62500                                            # multiple of 32: no scalar tail with one thread
    for scale in (80.0, 1e-3, 1e6):
        x = ((rng.random(n) - 0.5) * scale).astype(np.float32)
        y = ((rng.random(n) - 0.5) * scale).astype(np.float32)
        ref = torch.atan2(torch.from_numpy(y), torch.from_numpy(x)).numpy()
        assert np.array_equal(mine(y, x).view(np.int32), ref.view(np.int32))
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-40, -1e-40, 1e-30, 3e38, 5.0, -7.5], np.float32)
    yy, xx = [a.ravel().copy() for a in np.meshgrid(sp, sp)]
    pad = (-len(yy)) % 32
    yy, xx = np.concatenate((yy, np.ones(pad, np.float32))), np.concatenate((xx, np.ones(pad, np.float32)))
    ref = torch.atan2(torch.from_numpy(yy), torch.from_numpy(xx)).numpy()
    assert np.array_equal(mine(yy, xx).view(np.int32), ref.view(np.int32))
    nan = np.array([np.nan, 1.0] * 16, np.float32)
    assert np.isnan(mine(nan, nan[::-1].copy())).all()


def test_offline_preprocessing_golden():
    """The oracle's composition (4-channel projection -> normals -> row-major lists) against the files the
    reference's own `Preprocesser.preprocess_data()` wrote for two synthetic KITTI .bin scans
    (oracle/gen_golden.py:preprocess_golden)."""
    from delora_b200 import synthetic
    z = np.load(os.path.join(GOLDEN, "preprocess_16x200.npz"))
    cfg = synthetic.preprocessing_config("unused", "unused")
    for k in range(2):
        scan = torch.from_numpy(synthetic.kitti_bin_scan(60 + k)).t().contiguous()
        assert scan.shape[0] == 4
        img = orc.project_to_img(scan[None], 16, 200, cfg["horizontal_field_of_view"],
                                 cfg["kitti"]["vertical_field_of_view"])[0]
        assert img.shape[1] == 5                                   # x, y, z, reflectance, range
        normals, _, points = orc.compute_normal_vectors(img)
        assert np.array_equal(points.numpy(), z[f"points_{k}"])
        assert np.array_equal(normals.numpy(), z[f"normals_{k}"])


def test_projection_properties_random_clouds():
    """Oracle projection on random small clouds (no reference needed): every image pixel holds the in-FOV point of
    minimum range that rounds to it (lowest index on ties), the survivor list is sorted by (range, index), and
    (u, v) follow compute_2D_coordinates' op sequence."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(st.integers(0, 10_000), st.integers(1, 300))
    def check(seed, n):
        g = torch.Generator().manual_seed(seed)
        cloud = (torch.rand(3, n, generator=g) - 0.5) * torch.tensor([[40.0], [40.0], [6.0]])
        cloud[:, : n // 4] = cloud[:, n // 4: 2 * (n // 4)]                       # exact duplicates: ties
        h, w = 8, 32
        hf, vf = [-math.pi * 179.9 / 180, math.pi * 179.9 / 180], [-0.4, 0.3]
        image, u, v, idx, i2p = orc.project_to_img(cloud[None], h, w, hf, vf)
        rng = torch.norm(cloud, dim=0)
        uu = (torch.atan2(cloud[1], cloud[0]) - hf[0]) / (hf[1] - hf[0]) * (w - 1)
        vv = (torch.atan2(cloud[2], torch.norm(cloud[:2], dim=0)) - vf[0]) / (vf[1] - vf[0]) * (h - 1)
        ru, rv = torch.round(uu), torch.round(vv)
        inside = (ru >= 0) & (ru <= w - 1) & (rv >= 0) & (rv <= h - 1)
        best = {}
        for i in range(n):
            if inside[i]:
                key = (int(rv[i]), int(ru[i]))
                if key not in best or (float(rng[i]), i) < (float(rng[best[key]]), best[key]):
                    best[key] = i
        assert idx.shape[0] == len(best)
        for (r, c), i in best.items():
            assert torch.equal(image[0, :3, r, c], cloud[:, i]) and float(image[0, 3, r, c]) == float(rng[i])
        assert int((image[0, 3] != 0).sum()) == len([i for i in best.values() if float(rng[i]) != 0.0])
        keys = [(float(rng[i]), int(i)) for i in idx]
        assert keys == sorted(keys) and sorted(int(i) for i in idx) == sorted(best.values())
    import math
    check()

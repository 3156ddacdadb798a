#!/usr/bin/env python3
"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference on committed inputs.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py

Inputs come from `delora_b200/synthetic.py` (seeded, deterministic); outputs are what the
reference's own classes return:
  utility.projection.ImageProjectionLayer.forward          (src/utility/projection.py:108)
  preprocessing.normal_computation.NormalsComputer.compute_normal_vectors   (:89)
  deploy.deployer.Deployer.{transform,rotate}_point_cloud_transformation_matrix  (:181-189)
  losses.icp_losses.ICPLosses.forward                      (src/losses/icp_losses.py:28)
  models.model_parts.GeometryHandler.get_transformation_matrix_quaternion (kornia stub, see ref_harness)
Small cases store full tensors; KITTI-sized cases store SHA-256 digests + scalars so the
fixtures stay small.  The script also checks the oracle against what it just generated and
prints the verdict.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from delora_b200 import synthetic  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (pair index, w_raw, rings, vfov_deg, H, W)
    "small_16x180": (0, 192, 16, (-15.0, 15.0), 16, 180),
    "kitti_64x720": (1, 1875, 64, synthetic.KITTI_VFOV_DEG, 64, 720),
    "kitti_64x2048": (2, 2048, 64, synthetic.KITTI_VFOV_DEG, 64, 2048),
}


def canonical_projection_outputs(cloud, u, v, idx, stable=False):
    """The reference sorts by range with an UNSTABLE argsort (src/utility/projection.py:63), so
    the order among equal-range points is arbitrary.  Canonical forms used for pinning:
    u, v scattered back to the original point order; idx ordered by (range, index).
    `torch.argsort` is deterministic for a given input and build, so re-running it here
    recovers the permutation the reference used."""
    rng = torch.norm(cloud[None][:, :3, :], dim=1)
    sort_ref = torch.argsort(rng, dim=1, stable=stable)[0]   # stable=True only for the oracle's outputs
    u_orig = torch.empty_like(u)
    v_orig = torch.empty_like(v)
    u_orig[sort_ref] = u
    v_orig[sort_ref] = v
    order = np.lexsort((idx.numpy(), rng[0][idx].numpy()))
    return u_orig, v_orig, idx[torch.from_numpy(order)]


def digest(t):
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


def ref_pair(ref, cfg, scan_1, scan_2, t_pred, dataset="kitti"):
    """The reference's own path for one pair: 2x projection, 2x normals, transform, ICP losses,
    backward to the 4x4 transform."""
    import deploy.deployer
    proj = ref.projection.ImageProjectionLayer(config=cfg)
    nc = ref.normal_computation.NormalsComputer(config=cfg, dataset_name=dataset)
    out = {}
    imgs, lists = [], []
    for k, scan in ((1, scan_1), (2, scan_2)):
        image, u, v, idx, i2p = proj(input=scan[None].clone(), dataset=dataset)
        normals, has_n, points = nc.compute_normal_vectors(image=image.clone())
        uc, vc, idxc = canonical_projection_outputs(scan, u[0], v[0], idx)
        out[f"image_{k}"], out[f"u_{k}"], out[f"v_{k}"] = image[0], uc, vc
        out[f"idx_{k}"], out[f"i2p_{k}"] = idxc, i2p[0]
        out[f"normals_{k}"], out[f"has_normal_{k}"], out[f"points_{k}"] = normals, has_n, points
        imgs.append(image)
        lists.append((points, normals))
    dep = object.__new__(deploy.deployer.Deployer)
    tm = t_pred.clone().view(1, 4, 4).requires_grad_(True)
    (p1, n1), (p2, n2) = lists
    src = dep.transform_point_cloud_transformation_matrix(transformation_matrix=tm,
                                                          point_cloud=p2.t()[None].contiguous())
    src_n = dep.rotate_point_cloud_transformation_matrix(transformation_matrix=tm,
                                                         point_cloud=n2.t()[None].contiguous())
    icp = ref.icp_losses.ICPLosses(config=cfg)
    losses, plotting = icp(source_point_cloud_transformed=src, source_normal_list_transformed=src_n,
                           target_point_cloud=p1.t()[None].contiguous(),
                           target_normal_list=n1.t()[None].contiguous(),
                           compute_pointwise_loss_bool=False)
    loss = cfg["lambda_po2pl"] * losses["loss_po2pl"] + losses["loss_pl2pl"]
    loss.sum().backward()
    out["loss_po2pl"] = losses["loss_po2pl"].detach().reshape(1)
    out["loss_pl2pl"] = losses["loss_pl2pl"].detach().reshape(1)
    out["grad_T"] = tm.grad[0, :3, :].clone()
    out["num_pairs"] = torch.tensor([plotting["scan_2_transformed"].shape[2]])
    out["kept_source_points"] = plotting["scan_2_transformed"][0].detach()
    return out


def main():
    # torch's CPU elementwise kernels run SLEEF on full vectors and the scalar libm on the < 32-element
    # tail of every per-thread chunk, so the float (u, v) of a few elements depend on the thread count.
    # One thread makes the goldens reproducible on any machine (the tail is then only N mod 32 elements).
    torch.set_num_threads(1)
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_harness.reference_modules()
    from oracle import delora_oracle as orc
    summary = {}
    for name, (index, w_raw, rings, vfov, h, w) in CASES.items():
        cfg = synthetic.fov_config(h=h, w=w, vfov_deg=vfov)
        scan_1, scan_2, t_gt, t_pred = synthetic.make_pair(index, w_raw=w_raw, rings=rings, vfov_deg=vfov)
        out = ref_pair(ref, cfg, scan_1, scan_2, t_pred)
        meta = {"pair_index": index, "w_raw": w_raw, "rings": rings, "vfov_deg": list(vfov), "H": h, "W": w,
                "n_points": [scan_1.shape[1], scan_2.shape[1]],
                "K": [int(out["idx_1"].shape[0]), int(out["idx_2"].shape[0])],
                "P": [int(out["points_1"].shape[0]), int(out["points_2"].shape[0])],
                "num_pairs": int(out["num_pairs"][0]),
                "loss_po2pl": float(out["loss_po2pl"][0]), "loss_pl2pl": float(out["loss_pl2pl"][0]),
                "grad_T": out["grad_T"].numpy().astype(np.float64).tolist(),
                "inputs_sha256": [digest(scan_1), digest(scan_2)],
                "sha256": {k: digest(out[k]) for k in
                           ("image_1", "image_2", "u_1", "v_1", "idx_1", "idx_2",
                            "normals_1", "normals_2", "has_normal_1", "has_normal_2", "points_1", "points_2")}}
        if name.startswith("small"):
            np.savez_compressed(os.path.join(GOLDEN, name + ".npz"),
                                **{k: v.detach().cpu().numpy() for k, v in out.items()})
        else:
            # keep a strided sample of the normals so the GPU parity test has float goldens at size
            stride = 97
            np.savez_compressed(os.path.join(GOLDEN, name + "_sample.npz"),
                                normals_1=out["normals_1"][::stride].numpy(),
                                points_1=out["points_1"][::stride].numpy(),
                                has_normal_1=out["has_normal_1"][::stride].numpy(), stride=np.array([stride]))
        summary[name] = meta
        # ---- pin the oracle against what the reference just produced
        o = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg)
        img_o = orc.project_to_img(scan_1[None], h, w, cfg["horizontal_field_of_view"],
                                   cfg["kitti"]["vertical_field_of_view"])
        checks = {
            "image_1": torch.equal(o["image_1"][0], out["image_1"]),
            "image_2": torch.equal(o["image_2"][0], out["image_2"]),
            "u_1": torch.equal(canonical_projection_outputs(scan_1, img_o[1][0], img_o[2][0], img_o[3], stable=True)[0], out["u_1"]),
            "v_1": torch.equal(canonical_projection_outputs(scan_1, img_o[1][0], img_o[2][0], img_o[3], stable=True)[1], out["v_1"]),
            "idx_1": torch.equal(img_o[3], out["idx_1"]),
            "points_1": torch.equal(o["points_1"], out["points_1"]),
            "normals_1": torch.equal(o["normals_1"], out["normals_1"]),
            "normals_2": torch.equal(o["normals_2"], out["normals_2"]),
            "num_pairs": o["num_pairs"] == meta["num_pairs"],
            "loss_po2pl_rel": abs(o["loss_po2pl"] - meta["loss_po2pl"]) / meta["loss_po2pl"],
            "loss_pl2pl_rel": abs(o["loss_pl2pl"] - meta["loss_pl2pl"]) / meta["loss_pl2pl"],
            "grad_T_rel": float((o["grad_T"] - out["grad_T"]).abs().max() / out["grad_T"].abs().max()),
        }
        print(name, json.dumps(checks))
    # ---- projection-only stress clouds (full tensors, tiny)
    for name, cloud, (h, w, vfov) in (
            ("edge_16x180", synthetic.edge_stress_cloud(), (16, 180, (-15.0, 15.0))),
            ("tie_16x512", synthetic.tie_stress_cloud(), (16, 512, (-15.0, 15.0)))):
        cfg = synthetic.fov_config(h=h, w=w, vfov_deg=vfov)
        proj = ref.projection.ImageProjectionLayer(config=cfg)
        image, u, v, idx, i2p = proj(input=cloud[None].clone(), dataset="kitti")
        uc, vc, idxc = canonical_projection_outputs(cloud, u[0], v[0], idx)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), cloud=cloud.numpy(), image=image[0].numpy(),
                            u=uc.numpy(), v=vc.numpy(), idx=idxc.numpy())
        io = orc.project_to_img(cloud[None], h, w, cfg["horizontal_field_of_view"],
                                cfg["kitti"]["vertical_field_of_view"])
        print(name, "image", torch.equal(io[0], image), "idx", torch.equal(io[3], idxc),
              "idx_mismatch", int((io[3] != idxc).sum()) if io[3].shape == idxc.shape else "shape")
        summary[name] = {"H": h, "W": w, "vfov_deg": list(vfov), "K": int(idx.shape[0])}
    # ---- quaternion -> T (kornia 0.3.0 restatement) vs the in-repo quat2mat and scipy
    g = torch.Generator().manual_seed(7)
    quat = torch.randn(16, 4, generator=g)
    trans = torch.randn(16, 3, generator=g)
    t_ref = ref.model_parts.GeometryHandler.get_transformation_matrix_quaternion(
        translation=trans, quaternion=quat, device="cpu")
    from scipy.spatial.transform import Rotation
    r_scipy = torch.from_numpy(Rotation.from_quat(quat.numpy().astype(np.float64)).as_matrix()).float()
    print("quat: kornia-stub vs scipy max abs", float((t_ref[:, :3, :3] - r_scipy).abs().max()))
    np.savez_compressed(os.path.join(GOLDEN, "quaternion.npz"), quaternion=quat.numpy(), translation=trans.numpy(),
                        T=t_ref.numpy())
    # ---- encoder + heads: reference model (small width) with fixed seed -> state_dict + outputs
    mcfg = synthetic.fov_config(h=16, w=64, vfov_deg=(-15.0, 15.0))
    mcfg.update({"pre_feature_extraction": False, "resnet_outputs": 64, "use_dropout": False, "layers": [2, 2, 2, 2],
                 "factor_fewer_resnet_channels": 8, "activation_fct": "tanh", "use_single_mlp_at_output": False})
    torch.manual_seed(1234)
    ref_model = ref.model.OdometryModel(config=mcfg)
    g2 = torch.Generator().manual_seed(99)
    img_1 = torch.randn(2, 4, 16, 64, generator=g2)
    img_2 = torch.randn(2, 4, 16, 64, generator=g2)
    with torch.no_grad():
        tr, rot = ref_model(image_1=img_1, image_2=img_2)
        feats = ref_model.forward_features(image_1=img_1, image_2=img_2)
    sd = {"sd::" + k: v.numpy() for k, v in ref_model.state_dict().items()}
    np.savez_compressed(os.path.join(GOLDEN, "model_small.npz"), image_1=img_1.numpy(), image_2=img_2.numpy(),
                        translation=tr.numpy(), rotation=rot.numpy(), x1=feats[0].numpy(), x4=feats[3].numpy(), **sd)
    summary["model_small"] = {"params": int(sum(p.numel() for p in ref_model.parameters())),
                              "keys": sorted(ref_model.state_dict().keys())}
    summary["_versions"] = {"torch": torch.__version__, "numpy": np.__version__,
                            "scipy": __import__("scipy").__version__, "numba": __import__("numba").__version__,
                            "reference": "leggedrobotics/delora @ 15a25ee (SURVEY.md header)"}
    with open(os.path.join(GOLDEN, "golden.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print("written to", GOLDEN)


def preprocess_golden():
    """The reference's offline stage end to end (src/preprocessing/preprocesser.py:50-103 through
    src/data/kitti_scans.py): two synthetic KITTI .bin scans -> its own scans/ and normals/ .npy files."""
    import tempfile
    torch.set_num_threads(1)
    pp = ref_harness.reference_preprocesser()
    with tempfile.TemporaryDirectory() as tmp:
        velo = os.path.join(tmp, "raw", "00", "velodyne")
        os.makedirs(velo)
        for k in range(2):
            synthetic.kitti_bin_scan(60 + k).tofile(os.path.join(velo, format(k, "06d") + ".bin"))
        cfg = synthetic.preprocessing_config(os.path.join(tmp, "raw"), os.path.join(tmp, "pre"))
        pp.Preprocesser(config=cfg).preprocess_data()
        out = {}
        for k in range(2):
            out[f"points_{k}"] = np.load(os.path.join(tmp, "pre", "00", "scans", format(k, "06d") + ".npy"))
            out[f"normals_{k}"] = np.load(os.path.join(tmp, "pre", "00", "normals", format(k, "06d") + ".npy"))
    np.savez_compressed(os.path.join(GOLDEN, "preprocess_16x200.npz"), **out)
    # pin the oracle's composition (projection of the 4-channel cloud, normals, row-major lists)
    from oracle import delora_oracle as orc
    for k in range(2):
        scan = torch.from_numpy(synthetic.kitti_bin_scan(60 + k)).t().contiguous()
        img = orc.project_to_img(scan[None], 16, 200, cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"])[0]
        normals, _, points = orc.compute_normal_vectors(img)
        print("preprocess scan", k, "points", np.array_equal(points.numpy(), out[f"points_{k}"]),
              "normals", np.array_equal(normals.numpy(), out[f"normals_{k}"]), out[f"points_{k}"].shape)


def log_images_golden():
    """The logging / visualisation projections of the training step (src/deploy/deployer.py:73-89 create_images,
    :316-320 log_img_2_transformed): the reference's own Deployer.create_images on the small case, fed by its
    ICPLosses with compute_pointwise_loss_bool=True."""
    torch.set_num_threads(1)
    ref = ref_harness.reference_modules()
    import deploy.deployer
    idx, w_raw, rings, vfov, h, w = CASES["small_16x180"]
    cfg = synthetic.fov_config(h=h, w=w, vfov_deg=vfov)
    scan_1, scan_2, _, t_pred = synthetic.make_pair(idx, w_raw=w_raw, rings=rings, vfov_deg=vfov)
    proj = ref.projection.ImageProjectionLayer(config=cfg)
    nc = ref.normal_computation.NormalsComputer(config=cfg, dataset_name="kitti")
    lists = []
    for scan in (scan_1, scan_2):
        image = proj(input=scan[None].clone(), dataset="kitti")[0]
        normals, _, points = nc.compute_normal_vectors(image=image.clone())
        lists.append((points.t()[None].contiguous(), normals.t()[None].contiguous()))
    (p1, n1), (p2, n2) = lists
    dep = object.__new__(deploy.deployer.Deployer)
    dep.config, dep.img_projection = cfg, proj
    tm = t_pred.clone().view(1, 4, 4)
    src = dep.transform_point_cloud_transformation_matrix(transformation_matrix=tm, point_cloud=p2)
    src_n = dep.rotate_point_cloud_transformation_matrix(transformation_matrix=tm, point_cloud=n2)
    icp = ref.icp_losses.ICPLosses(config=cfg)
    losses, plotting = icp(source_point_cloud_transformed=src, source_normal_list_transformed=src_n,
                           target_point_cloud=p1, target_normal_list=n1, compute_pointwise_loss_bool=True)
    img_2_t, _, v_pixel, _, _ = proj(input=src, dataset="kitti")
    dep.create_images(preprocessed_data={"scan_1": p1, "normal_list_1": n1, "dataset": "kitti"}, losses=losses,
                      plotting=plotting)
    out = {"points_1": p1[0].numpy(), "normals_1": n1[0].numpy(), "points_2": p2[0].numpy(), "normals_2": n2[0].numpy(),
           "t_pred": t_pred.numpy(), "log_img_2_transformed": img_2_t[0].numpy(),
           "log_pointwise_loss": dep.log_pointwise_loss[0].detach().numpy(),
           "log_normals_target": dep.log_normals_target[0].numpy(),
           "log_normals_transformed_source": dep.log_normals_transformed_source[0].detach().numpy(),
           "visible_pixels": np.array([int(((torch.round(v_pixel) < h) & (v_pixel > 0)).sum())])}
    np.savez_compressed(os.path.join(GOLDEN, "log_images_16x180.npz"), **out)
    print("log images", {k: v.shape for k, v in out.items()})


def poses_golden():
    """Pose chaining (src/utility/poses.py:11-58) of 40 slightly non-orthonormal relative transforms."""
    ref_harness.install()
    import utility.poses as ref_poses
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(11)
    rel = []
    for _ in range(40):
        t = np.eye(4)[None].copy()
        t[0, :3, :3] = Rotation.from_euler("zyx", rng.normal(0, [0.03, 0.005, 0.005])).as_matrix()
        t[0, :3, :3] += rng.normal(0, 1e-7, (3, 3))          # slightly off SO(3), as a network output is
        t[0, :3, 3] = rng.normal([1.0, 0, 0], [0.2, 0.05, 0.02])
        rel.append(t.astype(np.float32))
    out = ref_poses.compute_poses([t.copy() for t in rel])
    np.savez_compressed(os.path.join(GOLDEN, "poses.npz"), relative=np.stack(rel), poses=out)
    print("poses", out.shape)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--poses-only":
        poses_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--log-images-only":
        log_images_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--preprocess-only":
        preprocess_golden()
        sys.exit(0)
    main()
    preprocess_golden()
    log_images_golden()
    poses_golden()

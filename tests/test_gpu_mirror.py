"""-m gpu: the reference-API mirror (same class names, signatures, return shapes as the
reference's modules) against the oracle / golden vectors."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, case_inputs
from oracle import delora_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cfg_for(meta, **extra):
    from delora_b200 import synthetic
    cfg = synthetic.fov_config(h=meta["H"], w=meta["W"], vfov_deg=tuple(meta["vfov_deg"]), device=DEV)
    cfg.update(extra)
    return cfg


@pytest.mark.parametrize("name", ["small_16x180", "kitti_64x720"])
def test_image_projection_layer_five_outputs(name, golden, cuda_lib):
    from delora_b200.utility.projection import ImageProjectionLayer
    meta = golden[name]
    _, (scan_1, _, _, _) = case_inputs(meta)
    cfg = cfg_for(meta)
    layer = ImageProjectionLayer(config=cfg)
    image, u, v, idx, i2p = layer(input=scan_1[None].to(DEV), dataset="kitti")
    hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
    # integer stage against the oracle on the kernel's own (u, v); float (u, v) within 1e-3 px
    rng = torch.norm(scan_1, dim=0)
    order = torch.argsort(rng, stable=True)
    u_orig = torch.empty_like(u[0].cpu()); u_orig[order] = u[0].cpu()
    v_orig = torch.empty_like(v[0].cpu()); v_orig[order] = v[0].cpu()
    io, uo, vo, idxo, i2po = orc.project_to_img(scan_1[None], meta["H"], meta["W"], hf, vf, uv_override=(u_orig, v_orig))
    assert image.shape == io.shape and image.dtype == torch.float32 and str(image.device).startswith("cuda")
    assert torch.equal(image.cpu(), io)
    assert idx.dtype == torch.int64 and torch.equal(idx.cpu(), idxo), "survivors in ascending (range, index) order"
    assert i2p.shape == i2po.shape and torch.equal(i2p.cpu(), i2po)
    assert u.shape == (1, scan_1.shape[1]) and torch.equal(u.cpu(), uo) and torch.equal(v.cpu(), vo)
    uc, vc, _, = orc.project_to_img(scan_1[None], meta["H"], meta["W"], hf, vf)[1:4]
    assert (u.cpu() - uc).abs().max() < 1e-3 and (v.cpu() - vc).abs().max() < 1e-3


def test_normals_computer_triple(golden, cuda_lib):
    from delora_b200.preprocessing.normal_computation import NormalsComputer
    meta = golden["kitti_64x720"]
    z = np.load(os.path.join(GOLDEN, "kitti_64x720_sample.npz"))
    _, (scan_1, _, _, _) = case_inputs(meta)
    cfg = cfg_for(meta)
    hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
    image = orc.project_to_img(scan_1[None], meta["H"], meta["W"], hf, vf)[0]
    nc = NormalsComputer(config=cfg, dataset_name="kitti")
    normals, has_normal, points = nc.compute_normal_vectors(image=image.to(DEV))
    assert normals.shape == (meta["P"][0], 3) and has_normal.dtype == torch.bool and points.shape == normals.shape
    stride = int(z["stride"][0])                                   # strided sample of the REFERENCE's output
    assert np.array_equal(points[::stride].cpu().numpy(), z["points_1"])
    assert np.array_equal(has_normal[::stride].cpu().numpy(), z["has_normal_1"])
    err = np.linalg.norm(normals[::stride].cpu().numpy() - z["normals_1"], axis=1)
    assert np.quantile(err, 0.99) < 2e-4 and (err > 1e-2).mean() < 5e-3


@pytest.mark.parametrize("po2po,normal_loss", [(False, "squared"), (True, "linear")])
def test_icp_losses_module_forward_backward(po2po, normal_loss, golden, cuda_lib):
    from delora_b200.losses.icp_losses import ICPLosses
    meta = golden["small_16x180"]
    cfg, (scan_1, scan_2, _, t_pred) = case_inputs(meta)
    out = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg, backward=False)
    gcfg = cfg_for(meta, point_to_point_loss=po2po, normal_loss=normal_loss)
    tm = t_pred.view(1, 4, 4)
    src = orc.transform_point_cloud(tm, out["points_2"].t()[None]).contiguous()
    src_n = orc.rotate_point_cloud(tm, out["normals_2"].t()[None]).contiguous()
    tgt, tgt_n = out["points_1"].t()[None].contiguous(), out["normals_1"].t()[None].contiguous()
    # oracle with autograd
    so, sno = src.clone().requires_grad_(True), src_n.clone().requires_grad_(True)
    lo, aux = orc.icp_losses(so, sno, tgt, tgt_n, point_to_point_loss=po2po, normal_loss=normal_loss, return_aux=True)
    (2.0 * lo["loss_po2pl"] + 0.5 * lo["loss_pl2pl"] + lo["loss_po2po"]).sum().backward()
    # mirror
    sg, sng = src.clone().to(DEV).requires_grad_(True), src_n.clone().to(DEV).requires_grad_(True)
    mod = ICPLosses(config=gcfg)
    losses, plotting = mod(source_point_cloud_transformed=sg, source_normal_list_transformed=sng,
                           target_point_cloud=tgt.to(DEV), target_normal_list=tgt_n.to(DEV),
                           compute_pointwise_loss_bool=True)
    assert set(losses) == {"loss_po2po", "loss_po2pl", "loss_po2pl_pointwise", "loss_pl2pl"}
    (2.0 * losses["loss_po2pl"] + 0.5 * losses["loss_pl2pl"] + losses["loss_po2po"]).sum().backward()
    for k in ("loss_po2pl", "loss_pl2pl", "loss_po2po"):
        assert float(losses[k].detach()) == pytest.approx(float(lo[k].detach()), rel=1e-5, abs=1e-12), k
    assert plotting["scan_2_transformed"].shape == (1, 3, aux["num_pairs"])
    assert torch.equal(plotting["scan_2_transformed"].detach().cpu(), aux["source_points_where_normals"].detach())
    assert losses["loss_po2pl_pointwise"].shape == (1, 3, aux["num_pairs"])
    assert torch.allclose(sg.grad.cpu(), so.grad, rtol=1e-4, atol=1e-9)
    assert torch.allclose(sng.grad.cpu(), sno.grad, rtol=1e-4, atol=1e-9)


def test_geometry_handler_autograd(cuda_lib):
    from delora_b200.models.model_parts import GeometryHandler
    z = np.load(os.path.join(GOLDEN, "quaternion.npz"))
    q = torch.from_numpy(z["quaternion"]).to(DEV).requires_grad_(True)
    t = torch.from_numpy(z["translation"]).to(DEV).requires_grad_(True)
    T = GeometryHandler.get_transformation_matrix_quaternion(translation=t, quaternion=q, device=DEV)
    assert np.abs(T.detach().cpu().numpy() - z["T"]).max() < 5e-7
    w = torch.randn(16, 4, 4, generator=torch.Generator().manual_seed(3))
    (T * w.to(DEV)).sum().backward()
    qo = torch.from_numpy(z["quaternion"]).requires_grad_(True)
    to = torch.from_numpy(z["translation"]).requires_grad_(True)
    (orc.transformation_matrix_quaternion(to, qo) * w).sum().backward()
    assert torch.allclose(q.grad.cpu(), qo.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(t.grad.cpu(), to.grad, rtol=1e-6, atol=1e-7)


def tiny_training_config(tmp_path, batch_size):
    from delora_b200 import synthetic
    cfg = synthetic.fov_config(h=16, w=180, vfov_deg=(-15.0, 15.0), device=DEV)
    cfg.update({"pre_feature_extraction": False, "resnet_outputs": 64, "use_dropout": False, "layers": [1, 1, 1, 1],
                "factor_fewer_resnet_channels": 8, "activation_fct": "tanh", "use_single_mlp_at_output": False,
                "batch_size": batch_size, "learning_rate": 1e-4, "store_dataset_in_RAM": True,
                "num_dataloader_workers": 0, "normalization_scaling": False, "inference_only": False,
                "unsupervised_at_start": True, "mode": "training", "experiment": "test",
                "training_run_name": "pytest_run", "run_name": "pytest_run", "checkpoint": None,
                "checkpoint_dir": str(tmp_path)})
    cfg["kitti"]["preprocessed_path"] = str(tmp_path / "pre")
    cfg["kitti"]["data_identifiers"] = [0]
    return cfg


def write_preprocessed(tmp_path, n_scans=3):
    """Preprocessed lists made by the oracle (CPU): [P,3] points + normals per scan."""
    from delora_b200 import synthetic
    seq = tmp_path / "pre" / "00"
    (seq / "scans").mkdir(parents=True)
    (seq / "normals").mkdir(parents=True)
    cfg = synthetic.fov_config(h=16, w=180, vfov_deg=(-15.0, 15.0))
    hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
    for k in range(n_scans):
        s1, s2, _, _ = synthetic.make_pair(40 + k, w_raw=192, rings=16, vfov_deg=(-15.0, 15.0))
        img = orc.project_to_img((s1 if k % 2 == 0 else s2)[None], 16, 180, hf, vf)[0]
        normals, _, points = orc.compute_normal_vectors(img)
        np.save(seq / "scans" / (format(k, "06d") + ".npy"), points.numpy())
        np.save(seq / "normals" / (format(k, "06d") + ".npy"), normals.numpy())


def test_deployer_step_matches_per_sample_reference_flow(tmp_path, cuda_lib):
    """Batched step (B=2) vs the reference's per-sample flow restated with the oracle: projection of
    the lists, sub-selection by the returned indices, ICP losses with the SAME predicted transforms,
    and the cumulative `loss_pc` quirk of src/deploy/deployer.py:312."""
    from delora_b200.deploy.trainer import Trainer
    write_preprocessed(tmp_path, n_scans=3)
    cfg = tiny_training_config(tmp_path, batch_size=2)
    torch.manual_seed(0)
    trainer = Trainer(config=cfg)
    trainer.training_bool = False                                   # forward only
    dicts = [trainer.dataset[0], trainer.dataset[1]]
    for d in dicts:
        for k in d:
            if hasattr(d[k], "to"):
                d[k] = d[k].to(DEV)
    el = Trainer.new_epoch_losses()
    el, T = trainer.step(preprocessed_dicts=dicts, epoch_losses=el)
    T = T.detach().cpu()
    hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]
    run = {"po2pl": 0.0, "pl2pl": 0.0, "pc": 0.0}
    for j, d in enumerate(dicts):
        lists = {}
        for key_s, key_n in (("scan_1", "normal_list_1"), ("scan_2", "normal_list_2")):
            s, n = d[key_s].cpu(), d[key_n].cpu()
            idx = orc.project_to_img(s, 16, 180, hf, vf)[3]
            lists[key_s], lists[key_n] = s[:, :, idx], n[:, :, idx]           # deployer.py:258-261
        src = orc.transform_point_cloud(T[j:j + 1], lists["scan_2"])
        src_n = orc.rotate_point_cloud(T[j:j + 1], lists["normal_list_2"])
        lo = orc.icp_losses(src, src_n, lists["scan_1"].contiguous(), lists["normal_list_1"].contiguous())
        run["po2pl"] += float(lo["loss_po2pl"])
        run["pl2pl"] += float(lo["loss_pl2pl"])
        run["pc"] += run["po2pl"] + run["pl2pl"]                            # :309-312 (running sums!)
    assert float(np.asarray(el["loss_po2pl_epoch"]).reshape(-1)[0]) == pytest.approx(run["po2pl"] / 2, rel=2e-5)
    assert float(np.asarray(el["loss_pl2pl_epoch"]).reshape(-1)[0]) == pytest.approx(run["pl2pl"] / 2, rel=2e-5)
    assert float(np.asarray(el["loss_point_cloud_epoch"]).reshape(-1)[0]) == pytest.approx(run["pc"] / 2, rel=2e-5)


def test_trainer_runs_and_checkpoints(tmp_path, cuda_lib):
    from delora_b200.deploy.trainer import Trainer
    write_preprocessed(tmp_path, n_scans=3)
    cfg = tiny_training_config(tmp_path, batch_size=1)
    torch.manual_seed(0)
    trainer = Trainer(config=cfg)
    before = [p.detach().clone() for p in trainer.model.parameters()]
    hist = trainer.train(max_epochs=2)
    assert len(hist) == 2 and all(np.isfinite(hist))
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, trainer.model.parameters())), "weights moved"
    ckpt = torch.load(tmp_path / "pytest_run_latest_checkpoint.pth", weights_only=False)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss", "parameters"}
    assert "resnet.conv1.weight" in ckpt["model_state_dict"]
    assert "fully_connected_rotation.1.weight" in ckpt["model_state_dict"]


def test_offline_preprocesser_matches_reference_files(tmp_path, cuda_lib):
    """`Preprocesser.preprocess_data()` on two synthetic KITTI .bin scans against the .npy files the REFERENCE's
    own Preprocesser wrote for the same input (tests/golden/preprocess_16x200.npz): point lists bit-exact,
    has-normal masks exact, normals to float tolerance; the output tree is what the training dataset reads."""
    from delora_b200 import synthetic
    from delora_b200.preprocessing.preprocesser import Preprocesser
    z = np.load(os.path.join(GOLDEN, "preprocess_16x200.npz"))
    velo = tmp_path / "raw" / "00" / "velodyne"
    velo.mkdir(parents=True)
    for k in range(2):
        synthetic.kitti_bin_scan(60 + k).tofile(velo / (format(k, "06d") + ".bin"))
    cfg = synthetic.preprocessing_config(tmp_path / "raw", tmp_path / "pre", device=DEV)
    cfg["preprocessing_batch_size"] = 2
    Preprocesser(config=cfg).preprocess_data()
    for k in range(2):
        pts = np.load(tmp_path / "pre" / "00" / "scans" / (format(k, "06d") + ".npy"))
        nrm = np.load(tmp_path / "pre" / "00" / "normals" / (format(k, "06d") + ".npy"))
        assert pts.dtype == np.float32 and nrm.dtype == np.float32 and pts.shape == nrm.shape
        assert np.array_equal(pts, z[f"points_{k}"])
        has_ref = (z[f"normals_{k}"] != 0).any(axis=1)
        assert np.array_equal((nrm != 0).any(axis=1), has_ref)
        err = np.linalg.norm(nrm - z[f"normals_{k}"], axis=1)[has_ref]
        assert np.quantile(err, 0.99) <= 1e-6 and err.max() <= 5e-5
    # per-scan entry point of the reference (same signature), single scan
    pre = Preprocesser(config=cfg)
    pre.config["dataset"] = "kitti"
    from delora_b200.preprocessing.normal_computation import NormalsComputer
    pre.normals_computer = NormalsComputer(config=cfg, dataset_name="kitti")
    (tmp_path / "one" / "scans").mkdir(parents=True)
    (tmp_path / "one" / "normals").mkdir(parents=True)
    pre.scans_name, pre.normals_name = str(tmp_path / "one" / "scans"), str(tmp_path / "one" / "normals")
    scan = torch.from_numpy(synthetic.kitti_bin_scan(60)).t().contiguous()[None]
    pre.apply_preprocessing_step(scan=scan, index=7)
    assert np.array_equal(np.load(tmp_path / "one" / "scans" / "000007.npy"), z["points_0"])


@pytest.mark.parametrize("use_graph", [False, True])
def test_odometry_stream_one_projection_per_frame(use_graph, cuda_lib):
    """`deploy.stream.OdometryStream`: frame k costs one projection (frame k-1's image is cached) + encoder +
    quaternion->T, optionally replayed as one CUDA graph; the relative transforms must equal the pairwise
    inference path (both projections recomputed, eager launches) bit for bit."""
    from delora_b200 import ops, synthetic
    from delora_b200.deploy.stream import OdometryStream
    from delora_b200.models.model import OdometryModel
    from delora_b200.models.model_parts import GeometryHandler
    h, w = 64, 512
    cfg = synthetic.fov_config(h=h, w=w, device=DEV)
    cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False, "layers": [2, 2, 2, 2],
                "factor_fewer_resnet_channels": 1, "activation_fct": "tanh", "use_single_mlp_at_output": False,
                "use_tensor_core_encoder": True})
    torch.manual_seed(3)
    model = OdometryModel(cfg).to(DEV).eval()
    frames = []
    for i in range(3):
        s1, s2, _, _ = synthetic.make_pair(80 + i, w_raw=512)
        frames += [s1, s2]
    n_max = max(f.shape[1] for f in frames)
    stream = OdometryStream(model, cfg, "kitti", n_max, use_cuda_graph=use_graph)
    got = [stream.push(f) for f in frames]
    assert got[0] is None and all(g.shape == (1, 4, 4) for g in got[1:])
    hf, vf = cfg["horizontal_field_of_view"], cfg["kitti"]["vertical_field_of_view"]

    def image_of(f):
        pts = torch.zeros((1, 3, n_max), device=DEV)
        pts[0, :, :f.shape[1]] = f.to(DEV)
        return ops.project(pts, torch.tensor([f.shape[1]], dtype=torch.int32, device=DEV), h, w, hf, vf)[0]

    for k in range(1, len(frames)):
        with torch.no_grad():
            t, q = model(image_1=image_of(frames[k - 1]), image_2=image_of(frames[k]))
            ref = GeometryHandler.get_transformation_matrix_quaternion(translation=t, quaternion=q, device=DEV)
        assert np.array_equal(got[k], ref.cpu().numpy()), k
    poses = stream.poses()
    assert poses.shape == (len(frames), 4, 4) and np.allclose(poses[0], np.eye(4))


def test_tester_inference_collects_transforms(tmp_path, cuda_lib):
    from delora_b200.deploy.tester import Tester
    from delora_b200.models.model import OdometryModel
    write_preprocessed(tmp_path, n_scans=4)
    cfg = tiny_training_config(tmp_path, batch_size=1)
    torch.manual_seed(0)
    ckpt = tmp_path / "model.pth"
    torch.save({"model_state_dict": OdometryModel(cfg).state_dict()}, ckpt)
    cfg.update({"checkpoint": str(ckpt), "inference_only": True, "mode": "testing"})
    tester = Tester(config=cfg)
    tester.test()
    rel = tester.computed_transformations_datasets[0][0]
    assert len(rel) == len(tester.dataset) == 3 and all(t.shape == (1, 4, 4) for t in rel)
    # the same pairs through Deployer.step directly
    d = tester.dataset[1]
    for k in d:
        if hasattr(d[k], "to"):
            d[k] = d[k].to(DEV)
    with torch.no_grad():
        direct = tester.step(preprocessed_dicts=[d])
    assert np.array_equal(direct.cpu().numpy(), rel[1])
    assert tester.poses().shape == (4, 4, 4)
    cfg["checkpoint"] = None
    with pytest.raises(Exception, match="No checkpoint"):
        Tester(config=cfg)


@pytest.mark.parametrize("scaling", [False, True])
def test_step_with_padded_batch_equals_list_of_dicts(scaling, tmp_path, cuda_lib):
    """Device-side batching (data/batching.py): `Deployer.step(PaddedBatch)` must give exactly what the reference-shaped
    list-of-dicts path gives (same kernels, same inputs), with and without `normalization_scaling`."""
    from delora_b200.data import batching
    from delora_b200.deploy.trainer import Trainer
    write_preprocessed(tmp_path, n_scans=4)
    cfg = tiny_training_config(tmp_path, batch_size=3)
    cfg["normalization_scaling"] = scaling
    torch.manual_seed(0)
    trainer = Trainer(config=cfg)
    trainer.training_bool = False
    items = [trainer.dataset[i] for i in range(3)]
    padded = batching.padded_collate([trainer.dataset[i] for i in range(3)]).to(DEV)
    for d in items:
        for k in d:
            if hasattr(d[k], "to"):
                d[k] = d[k].to(DEV)
    el_a, t_a = trainer.step(preprocessed_dicts=items, epoch_losses=Trainer.new_epoch_losses())
    el_b, t_b = trainer.step(preprocessed_dicts=padded, epoch_losses=Trainer.new_epoch_losses())
    assert torch.equal(t_a, t_b)
    for k in ("loss_epoch", "loss_point_cloud_epoch", "loss_po2pl_epoch", "loss_pl2pl_epoch", "visible_pixels_epoch"):
        assert np.array_equal(np.asarray(el_a[k]), np.asarray(el_b[k])), k
    # the prefetching loader feeds the trainer end to end
    cfg2 = tiny_training_config(tmp_path, batch_size=2)
    trainer2 = Trainer(config=cfg2)
    hist = trainer2.train(max_epochs=1)
    assert len(hist) == 1 and np.isfinite(hist[0])


def test_log_images_match_reference_create_images(cuda_lib):
    """`Deployer.create_images` + `log_img_2_transformed` on the device against the images the REFERENCE's
    own Deployer.create_images / ImageProjectionLayer produced for the same lists (tests/golden/log_images_16x180.npz,
    written by oracle/gen_golden.py --log-images-only).  The transformed source points are recomputed on the GPU
    (fp32 matmul), so a point within 1e-6 px of a pixel boundary may land in the neighbouring pixel: at most 0.3 % of
    the pixels may differ, all the others must agree to 1e-5."""
    from delora_b200 import synthetic
    from delora_b200.deploy.deployer import Deployer
    from delora_b200.losses.icp_losses import ICPLosses
    from delora_b200.utility.projection import ImageProjectionLayer
    z = np.load(os.path.join(GOLDEN, "log_images_16x180.npz"))
    cfg = synthetic.fov_config(h=16, w=180, vfov_deg=(-15.0, 15.0), device=DEV)
    dep = object.__new__(Deployer)
    dep.config, dep.device = cfg, DEV
    dep.img_projection = ImageProjectionLayer(config=cfg)
    dep.lossPointCloud = ICPLosses(config=cfg)
    t = lambda k: torch.from_numpy(z[k])[None].to(DEV).contiguous()
    p1, n1, p2, n2 = t("points_1"), t("normals_1"), t("points_2"), t("normals_2")
    tm = torch.from_numpy(z["t_pred"]).view(1, 4, 4).to(DEV)
    src = dep.transform_point_cloud_transformation_matrix(tm, p2).contiguous()
    src_n = dep.rotate_point_cloud_transformation_matrix(tm, n2).contiguous()
    losses, plotting = dep.lossPointCloud(source_point_cloud_transformed=src, source_normal_list_transformed=src_n,
                                          target_point_cloud=p1, target_normal_list=n1, compute_pointwise_loss_bool=True)
    img2t, _, v_pix, _, _ = dep.img_projection(input=src, dataset="kitti")
    dep.create_images(preprocessed_data={"scan_1": p1, "normal_list_1": n1, "dataset": "kitti"}, losses=losses,
                      plotting=plotting)
    got = {"log_img_2_transformed": img2t[0], "log_pointwise_loss": dep.log_pointwise_loss[0],
           "log_normals_target": dep.log_normals_target[0],
           "log_normals_transformed_source": dep.log_normals_transformed_source[0]}
    for k, g in got.items():
        ref = torch.from_numpy(z[k])
        assert tuple(g.shape) == tuple(ref.shape), k
        differ = ((g.cpu() - ref).abs() > 1e-5 * (1.0 + ref.abs())).any(dim=0)
        print(f"[{k}] pixels differing from the reference: {int(differ.sum())} of {differ.numel()}")
        assert differ.float().mean().item() <= 3e-3, k
    assert torch.equal(got["log_normals_target"].cpu(), torch.from_numpy(z["log_normals_target"])), "no recomputation here"
    visible = int(((torch.round(v_pix) < 16) & (v_pix > 0)).sum())
    assert abs(visible - int(z["visible_pixels"][0])) <= 2


def test_icp_losses_po2po_alone_branch(golden, cuda_lib):
    """`po2po_alone: True` (src/losses/icp_losses.py:36-46): every source point against its nearest target point,
    no normals involved; value and gradient against the oracle / torch autograd."""
    from delora_b200.losses.icp_losses import ICPLosses
    meta = golden["small_16x180"]
    cfg, (scan_1, scan_2, _, t_pred) = case_inputs(meta)
    gcfg = cfg_for(meta, po2po_alone=True, point_to_point_loss=True, point_to_plane_loss=False, plane_to_plane_loss=False)
    out = orc.pair_forward_backward(scan_1, scan_2, t_pred, cfg)
    tm = t_pred.view(1, 4, 4)
    src = orc.transform_point_cloud(tm, out["points_2"].t()[None]).contiguous()
    tgt = out["points_1"].t()[None].contiguous()
    so = src.clone().requires_grad_(True)
    nn = torch.from_numpy(orc.nearest_neighbors(tgt[0].t().numpy(), src[0].t().numpy()))
    lo = torch.nn.functional.mse_loss(so, tgt[:, :, nn])
    lo.backward()
    sg = src.to(DEV).requires_grad_(True)
    zeros = torch.zeros_like(sg)
    losses, plotting = ICPLosses(gcfg)(source_point_cloud_transformed=sg, source_normal_list_transformed=zeros,
                                       target_point_cloud=tgt.to(DEV), target_normal_list=torch.zeros_like(tgt).to(DEV),
                                       compute_pointwise_loss_bool=False)
    assert plotting is None and float(losses["loss_po2pl"]) == 0.0 and float(losses["loss_pl2pl"]) == 0.0
    assert float(losses["loss_po2po"]) == pytest.approx(float(lo), rel=1e-5)
    losses["loss_po2po"].sum().backward()
    assert torch.allclose(sg.grad.cpu(), so.grad, rtol=1e-4, atol=1e-10)


def test_trainer_default_path_launches_tcgen05_kernels(tmp_path, cuda_lib):
    """With the reference's configuration keys only (no opt-in flag) the full-width model trains on the tensor-core
    encoder: the profiler's kernel list of one training step holds the tcgen05 convolution kernels and NO cuDNN /
    cuBLAS convolution kernel."""
    from delora_b200.deploy.trainer import Trainer
    write_preprocessed(tmp_path, n_scans=3)
    cfg = tiny_training_config(tmp_path, batch_size=2)
    cfg.update({"factor_fewer_resnet_channels": 1, "layers": [2, 2, 2, 2], "resnet_outputs": 1000})
    assert "use_tensor_core_encoder" not in cfg
    torch.manual_seed(0)
    trainer = Trainer(config=cfg)
    dicts = [trainer.dataset[0], trainer.dataset[1]]
    for d in dicts:
        for k in d:
            if hasattr(d[k], "to"):
                d[k] = d[k].to(DEV)
    before = trainer.model.resnet.layer3[0].conv1.weight.detach().clone()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        trainer.optimizer.zero_grad()
        trainer.step(preprocessed_dicts=dicts, epoch_losses=Trainer.new_epoch_losses())
        torch.cuda.synchronize()
    names = {e.key for e in prof.key_averages()}
    ours = [n for n in names if "delora::conv_" in n or "delora::stem_" in n]
    assert any("conv_rows_tc_kernel" in n or "conv_fprop_tc_kernel" in n for n in ours), names
    assert any("conv_wgrad2_tc_kernel" in n or "conv_wgrad_tc_kernel" in n for n in ours), names
    assert any("stem_fprop_tc_kernel" in n for n in ours), names
    lib_conv = [n for n in names if ("cudnn" in n.lower() or "xmma" in n.lower() or "implicit_gemm" in n.lower()
                                     or "conv2d" in n.lower() or "wgrad" in n.lower() or "dgrad" in n.lower())
                and "delora::" not in n]
    assert not lib_conv, f"library convolution kernels on the default training path: {lib_conv}"
    assert not torch.equal(before, trainer.model.resnet.layer3[0].conv1.weight.detach()), "weights moved"

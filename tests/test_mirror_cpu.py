"""-m "not gpu": host-side logic of the reference-API mirror (model parameter names / forward vs
the reference's own model, dataset indexing, CLI config merge, gradient all-reduce over gloo)."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT


def small_model_config():
    from delora_b200 import synthetic
    cfg = synthetic.fov_config(h=16, w=64, vfov_deg=(-15.0, 15.0))
    cfg.update({"pre_feature_extraction": False, "resnet_outputs": 64, "use_dropout": False, "layers": [2, 2, 2, 2],
                "factor_fewer_resnet_channels": 8, "activation_fct": "tanh", "use_single_mlp_at_output": False})
    return cfg


def test_model_state_dict_interchanges_with_reference(golden):
    """Load the REFERENCE model's state_dict (golden) into the mirror and reproduce its outputs."""
    from delora_b200.models.model import OdometryModel
    z = np.load(os.path.join(GOLDEN, "model_small.npz"))
    model = OdometryModel(config=small_model_config())
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    assert sorted(model.state_dict().keys()) == golden["model_small"]["keys"]
    model.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in model.parameters()) == golden["model_small"]["params"]
    with torch.no_grad():
        tr, rot = model(image_1=torch.from_numpy(z["image_1"]), image_2=torch.from_numpy(z["image_2"]))
        feats = model.forward_features(image_1=torch.from_numpy(z["image_1"]), image_2=torch.from_numpy(z["image_2"]))
    assert np.allclose(feats[0].numpy(), z["x1"], rtol=1e-5, atol=1e-6)
    assert np.allclose(feats[3].numpy(), z["x4"], rtol=1e-5, atol=1e-6)
    assert np.allclose(tr.numpy(), z["translation"], rtol=1e-5, atol=1e-6)
    assert np.allclose(rot.numpy(), z["rotation"], rtol=1e-5, atol=1e-6)


def test_full_size_model_parameter_count():
    """11,675,112 encoder + 100,504 rotation + 100,403 translation parameters."""
    from delora_b200.models.model import OdometryModel
    cfg = small_model_config()
    cfg.update({"resnet_outputs": 1000, "factor_fewer_resnet_channels": 1})
    model = OdometryModel(config=cfg)
    assert sum(p.numel() for p in model.resnet.parameters()) == 11675112
    assert sum(p.numel() for p in model.fully_connected_rotation.parameters()) == 100504
    assert sum(p.numel() for p in model.fully_connected_translation.parameters()) == 100403


def write_tiny_dataset(root, n_scans=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    seq = os.path.join(root, "00")
    os.makedirs(os.path.join(seq, "scans"))
    os.makedirs(os.path.join(seq, "normals"))
    for k in range(n_scans):
        p = 50 + 7 * k
        np.save(os.path.join(seq, "scans", format(k, "06d") + ".npy"), torch.randn(p, 3, generator=g).numpy())
        np.save(os.path.join(seq, "normals", format(k, "06d") + ".npy"), torch.randn(p, 3, generator=g).numpy())


def test_dataset_layout_and_keys(tmp_path):
    from delora_b200.data.dataset import PreprocessedPointCloudDataset
    write_tiny_dataset(str(tmp_path))
    cfg = {"datasets": ["kitti"], "store_dataset_in_RAM": False,
           "kitti": {"preprocessed_path": str(tmp_path), "data_identifiers": [0]}}
    ds = PreprocessedPointCloudDataset(config=cfg)
    assert len(ds) == 3                                           # consecutive pairs (t, t+1)
    item = ds[1]
    assert set(item) == {"index", "index_dataset", "index_sequence", "index_scan", "dataset", "normal_list_1",
                         "normal_list_2", "scan_1", "scan_2"}
    assert item["scan_1"].shape == (1, 3, 57) and item["scan_2"].shape == (1, 3, 64)
    assert item["dataset"] == "kitti" and item["index_scan"] == 1
    cfg["store_dataset_in_RAM"] = True
    ds2 = PreprocessedPointCloudDataset(config=cfg)
    assert torch.equal(ds2[1]["scan_2"], item["scan_2"])
    cfg["kitti"]["data_identifiers"] = [7]
    with pytest.raises(Exception, match="does not exist"):
        PreprocessedPointCloudDataset(config=cfg)


def test_cli_config_merge_matches_reference_yaml(tmp_path):
    """bin/run_training.py::build_config on a copy of the reference-format YAML files."""
    sys.path.insert(0, os.path.join(ROOT, "bin"))
    import run_training
    cdir = tmp_path / "config"
    cdir.mkdir()
    (cdir / "config_datasets.yaml").write_text(
        "horizontal_field_of_view: [ -179.9, 179.9 ]\nepsilon_range: 0.5\n"
        "kitti:\n  training_identifiers: [ 0 ]\n  vertical_field_of_view: [ -24.5, 2.0 ]\n"
        "  vertical_cells: 64\n  horizontal_cells: 720\n")
    (cdir / "deployment_options.yaml").write_text('datasets: ["kitti"]\ndevice: "cpu"\nexperiment: "e"\n')
    (cdir / "hyperparameters.yaml").write_text("batch_size: 1\nlearning_rate: 0.00001\n")
    cfg = run_training.build_config("run", "exp", "", config_dir=str(cdir))
    assert cfg["kitti"]["vertical_field_of_view"][0] == pytest.approx(-24.5 * np.pi / 180.0)
    assert cfg["horizontal_field_of_view"][1] == pytest.approx(179.9 * np.pi / 180.0)
    assert cfg["kitti"]["data_identifiers"] == [0] and cfg["mode"] == "training" and cfg["experiment"] == "exp"
    assert cfg["device"] == torch.device("cpu") and cfg["run_name"] == "run" and cfg["checkpoint"] is None


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from delora_b200.parallel_grad import FlatGradAllReduce, shard_pairs
    torch.manual_seed(100 + rank)                                  # different init per rank -> broadcast must fix it
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    sync = FlatGradAllReduce(model)
    x = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 10.0
    shard = shard_pairs(8, rank, world)
    model(x[shard]).pow(2).sum().backward()
    sync.all_reduce()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    w0 = torch.cat([p.data.reshape(-1) for p in model.parameters()])
    torch.save({"grad": flat, "w": w0, "shard": shard}, os.path.join(out, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    """Two gloo ranks, each on its own shard of the pairs: after the flat all-reduce both hold the
    average of the two shard gradients == the gradient of the half-summed full batch."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert r0["shard"] == [0, 1, 2, 3] and r1["shard"] == [4, 5, 6, 7]
    assert torch.equal(r0["w"], r1["w"]), "parameters are broadcast from rank 0"
    assert torch.allclose(r0["grad"], r1["grad"], rtol=0, atol=0)
    # single-process check
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 3))
    x = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 10.0
    (model(x).pow(2).sum() / 2).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(r0["grad"], ref, rtol=1e-5, atol=1e-6)


class _FakeTrunk:
    """Stands in for the tensor-core encoder on the CPU: two 'trunk' weights whose gradients are written straight
    into the flat views and announced group by group from inside backward (parallel_grad's encoder protocol)."""

    def __init__(self):
        self.w = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(3, 3))]
        self.grad_views, self.grad_ready = None, None

    def trunk_parameters(self):
        return self.w

    def trunk_parameter_groups(self):
        return [[1], [0]]                                   # backward finishes w[1] first

    def apply(self, x):
        enc = self

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, w0, w1):
                h = torch.tanh(x @ w0.t())
                ctx.save_for_backward(x, w0, w1, h)
                return h @ w1.t()

            @staticmethod
            def backward(ctx, g):
                x, w0, w1, h = ctx.saved_tensors
                g1 = g.t() @ h
                out1 = enc.grad_views[1] if enc.grad_views is not None else torch.empty_like(w1)
                out1.copy_(g1)
                if enc.grad_ready is not None:
                    enc.grad_ready([1])
                gh = (g @ w1) * (1 - h * h)
                out0 = enc.grad_views[0] if enc.grad_views is not None else torch.empty_like(w0)
                out0.copy_(gh.t() @ x)
                if enc.grad_ready is not None:
                    enc.grad_ready([0])
                return None, out0, out1
        return Fn.apply(x, self.w[0], self.w[1])


class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.trunk = _FakeTrunk()
        self.w0, self.w1 = self.trunk.w
        self.head = torch.nn.Linear(3, 2)

    def forward(self, x):
        return self.head(self.trunk.apply(x))


def _bucket_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from delora_b200.parallel_grad import BucketedGradAllReduce, shard_pairs
    torch.manual_seed(200 + rank)
    model = _FakeModel()
    sync = BucketedGradAllReduce(model, encoder=model.trunk)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    x = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 10.0
    shard = shard_pairs(8, rank, world)
    grads = []
    for _ in range(2):                                     # two steps: the per-step bucket state must re-arm
        opt.zero_grad(set_to_none=True)
        model(x[shard]).pow(2).sum().backward()
        sync.finish()
        assert all(p.grad.data_ptr() == sync.views[id(p)].data_ptr() for p in model.parameters())
        grads.append(torch.cat([p.grad.reshape(-1).clone() for p in model.parameters()]))
        opt.step()
    torch.save({"grads": grads, "w": torch.cat([p.data.reshape(-1) for p in model.parameters()])},
               os.path.join(out, f"b{rank}.pt"))
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_world_size_2_gloo(tmp_path):
    """BucketedGradAllReduce (flat gradient buffer, per-bucket all-reduce launched from inside backward): two gloo
    ranks end up with the gradient of the half-summed full batch, for two consecutive steps, and stay in lock step."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_bucket_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    assert torch.equal(r0["w"], r1["w"])
    for g0, g1 in zip(r0["grads"], r1["grads"]):
        assert torch.equal(g0, g1)
    torch.manual_seed(200)
    model = _FakeModel()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    x = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 10.0
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        (model(x).pow(2).sum() / 2).backward()
        ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        assert torch.allclose(r0["grads"][step], ref, rtol=1e-5, atol=1e-6), step
        opt.step()


def test_bucket_layout_and_transport_arguments():
    """Flat gradient buffer: buckets start on 128-float boundaries (the peer-memory kernel moves 16-byte vectors and
    splits a bucket between the ranks), the views tile the buckets without overlap; transport names are checked and
    the peer transport is refused without CUDA."""
    from delora_b200.parallel_grad import BucketedGradAllReduce
    model = _FakeModel()
    sync = BucketedGradAllReduce(model, encoder=model.trunk)
    assert sync.transport == "nccl" and sync.peer is None          # one process, CPU: nothing to rendezvous with
    assert all(s % 128 == 0 and e % 128 == 0 and e > s for s, e in sync.ranges)
    assert sync.ranges[-1][1] == sync.flat.numel()
    base, seen = sync.flat.data_ptr(), []
    for bi, bucket in enumerate(sync.buckets):
        s, e = sync.ranges[bi]
        for p in bucket:
            off = (sync.views[id(p)].data_ptr() - base) // 4
            assert s <= off and off + p.numel() <= e
            seen.append((off, off + p.numel()))
    seen.sort()
    assert all(a[1] <= b[0] for a, b in zip(seen, seen[1:]))
    with pytest.raises(ValueError):
        BucketedGradAllReduce(_FakeModel(), transport="smoke-signals")
    if not torch.cuda.is_available():
        with pytest.raises(ValueError):
            BucketedGradAllReduce(_FakeModel(), transport="peer")


def test_nvtx_ranges_are_opt_in(monkeypatch):
    """`ops.nvtx_range` (DELORA_NVTX=1): pushes / pops a named range only when enabled; usable without a GPU."""
    from delora_b200 import ops
    calls = []
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda name: calls.append(("push", name)))
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda: calls.append(("pop",)))
    monkeypatch.setattr(ops, "NVTX", False)
    with ops.nvtx_range("normals"):
        pass
    assert calls == []
    monkeypatch.setattr(ops, "NVTX", True)
    with pytest.raises(KeyError):
        with ops.nvtx_range("normals"):
            raise KeyError("propagates, range still closed")
    assert calls == [("push", "normals"), ("pop",)]


def test_kitti_bin_reader(tmp_path):
    """`data.kitti_scans.KITTIPointCloudDataset`: sorted *.bin files -> [4, N] float32 (src/data/kitti_scans.py:35-50)."""
    from delora_b200 import synthetic
    from delora_b200.data.kitti_scans import KITTIPointCloudDataset
    velo = tmp_path / "03" / "velodyne"
    velo.mkdir(parents=True)
    scans = [synthetic.kitti_bin_scan(70 + k) for k in range(3)]
    for k in (2, 0, 1):
        scans[k].tofile(velo / (format(k, "06d") + ".bin"))
    ds = KITTIPointCloudDataset(base_dir=str(tmp_path), identifier=3, device="cpu")
    assert len(ds) == 3 and ds.num_elements == 3
    for k in range(3):
        t = ds[k]
        assert t.shape == (4, scans[k].shape[0]) and t.dtype == torch.float32
        assert np.array_equal(t.t().numpy(), scans[k])


def test_compute_poses_golden(tmp_path):
    """`utility.poses.compute_poses` against the reference's output (tests/golden/poses.npz) + the KITTI pose file."""
    from delora_b200.utility import poses
    z = np.load(os.path.join(GOLDEN, "poses.npz"))
    out = poses.compute_poses([t.copy() for t in z["relative"]])
    assert out.shape == z["poses"].shape == (41, 4, 4)
    assert np.abs(out - z["poses"]).max() < 1e-12
    poses.write_poses_to_text_file(str(tmp_path / "poses.txt"), out)
    rows = np.loadtxt(tmp_path / "poses.txt")
    assert rows.shape == (41, 12) and np.allclose(rows, out.reshape(41, 16)[:, :12])


def test_padded_collate_and_prefetch_host_logic():
    """data.batching: list of dataset items -> one staging buffer with [2B,3,Nmax] points / normals + counts
    (scan_1 of every sample first, then scan_2), zero padding, metadata kept; PrefetchLoader on CPU is a pass-through."""
    from delora_b200.data import batching
    g = torch.Generator().manual_seed(5)
    items = []
    for i, (n1, n2) in enumerate(((7, 5), (3, 9), (6, 6))):
        items.append({"index": i, "index_dataset": 0, "index_sequence": 1, "index_scan": i, "dataset": "kitti",
                      "scan_1": torch.randn(1, 3, n1, generator=g), "scan_2": torch.randn(1, 3, n2, generator=g),
                      "normal_list_1": torch.randn(1, 3, n1, generator=g), "normal_list_2": torch.randn(1, 3, n2, generator=g)})
    batch = batching.padded_collate(items)
    assert len(batch) == 3 and batch.n_max == 9 and batch.dataset == "kitti"
    assert batch.points.shape == (6, 3, 9) and batch.counts.tolist() == [7, 3, 6, 5, 9, 6]
    for i, d in enumerate(items):
        assert torch.equal(batch.points[i, :, :d["scan_1"].shape[2]], d["scan_1"][0])
        assert torch.equal(batch.points[3 + i, :, :d["scan_2"].shape[2]], d["scan_2"][0])
        assert torch.equal(batch.normals[3 + i, :, :d["scan_2"].shape[2]], d["normal_list_2"][0])
        assert float(batch.points[i, :, d["scan_1"].shape[2]:].abs().sum()) == 0.0
        assert batch.meta[i]["index_scan"] == i and "scan_1" not in batch.meta[i]
    # views alias one flat buffer (a single copy ships everything)
    assert batch.points.data_ptr() == batch.flat.data_ptr()
    loader = torch.utils.data.DataLoader(items, batch_size=2, collate_fn=batching.padded_collate)
    out = list(batching.PrefetchLoader(loader, "cpu"))
    assert [len(b) for b in out] == [2, 1] and out[1].counts.tolist() == [6, 6]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from helpers import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_pipeline_staging_layout_views():
    """`ScanPairPipeline.staging_layout` / `input_views`: one flat buffer, 256-byte aligned sections, typed views
    that alias it (what lets a host caller ship a step with a single copy)."""
    from delora_b200.pipeline import ScanPairPipeline
    lay = ScanPairPipeline.staging_layout(3, 3, 1001)
    assert lay["n_points"] % 256 == 0 and lay["transform"] % 256 == 0 and lay["bytes"] % 256 == 0
    assert lay["n_points"] >= 2 * 3 * 3 * 1001 * 4 and lay["transform"] >= lay["n_points"] + 2 * 3 * 4
    flat = torch.zeros((lay["bytes"],), dtype=torch.uint8)
    points, n_points, transform = ScanPairPipeline.input_views(flat, lay)
    assert points.shape == (6, 3, 1001) and n_points.shape == (6,) and transform.shape == (3, 12)
    points[5, 2, 1000] = 1.5
    n_points[5] = 77
    transform[2, 11] = -2.0
    raw = flat.numpy()
    assert raw[:2 * 3 * 3 * 1001 * 4].view("float32")[-1] == 1.5
    assert raw[lay["n_points"]:lay["n_points"] + 24].view("int32")[5] == 77
    assert raw[lay["transform"]:lay["transform"] + 3 * 48].view("float32")[-1] == -2.0
    with pytest.raises(RuntimeError, match="CUDA"):
        ScanPairPipeline(1, 16, 4, 8, (-3.0, 3.0), (-0.4, 0.1), device="cpu")


def test_testing_cli_config_prefers_checkpoint_parameters(tmp_path):
    """bin/run_testing.py::build_config (reference bin/run_testing.py:19-91): YAML merge, testing identifiers,
    checkpoint-carried network parameters, dropout off, mode / unsupervised flags."""
    sys.path.insert(0, os.path.join(ROOT, "bin"))
    import run_testing
    cdir = tmp_path / "config"
    cdir.mkdir()
    (cdir / "config_datasets.yaml").write_text(
        "horizontal_field_of_view: [ -179.9, 179.9 ]\n"
        "kitti:\n  training_identifiers: [ 0 ]\n  testing_identifiers: [ 9, 10 ]\n"
        "  vertical_field_of_view: [ -24.5, 2.0 ]\n  vertical_cells: 64\n  horizontal_cells: 720\n")
    (cdir / "deployment_options.yaml").write_text(
        'datasets: ["kitti"]\ndevice: "cpu"\nexperiment: "e"\ninference_only: True\nstore_dataset_in_RAM: False\n')
    (cdir / "hyperparameters.yaml").write_text("batch_size: 4\nuse_dropout: True\nlayers: [2, 2, 2, 2]\n")
    plain = tmp_path / "plain.pth"
    torch.save({"model_state_dict": {}}, plain)
    cfg = run_testing.build_config("run", "exp", str(plain), config_dir=str(cdir))
    assert cfg["kitti"]["data_identifiers"] == [9, 10] and cfg["mode"] == "testing" and cfg["unsupervised_at_start"]
    assert cfg["use_dropout"] is False and cfg["checkpoint"] == str(plain) and cfg["experiment"] == "exp"
    assert cfg["kitti"]["vertical_field_of_view"][1] == pytest.approx(2.0 * np.pi / 180.0)
    # a checkpoint that carries the training run's parameters: those win, angles are NOT converted twice
    params = {"kitti": {"vertical_field_of_view": [-0.4, 0.03], "vertical_cells": 64, "horizontal_cells": 720,
                        "training_identifiers": [0]},
              "horizontal_field_of_view": [-3.14, 3.14], "layers": [1, 1, 1, 1], "use_dropout": False, "batch_size": 8}
    rich = tmp_path / "rich.pth"
    torch.save({"model_state_dict": {}, "parameters": params}, rich)
    cfg = run_testing.build_config("run", "", str(rich), config_dir=str(cdir))
    assert cfg["layers"] == [1, 1, 1, 1] and cfg["kitti"]["vertical_field_of_view"] == [-0.4, 0.03]
    assert cfg["kitti"]["data_identifiers"] == [9, 10] and cfg["inference_only"] is True
    assert cfg["device"] == torch.device("cpu") and cfg["store_dataset_in_RAM"] is False

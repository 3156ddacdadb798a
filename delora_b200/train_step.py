"""Full training step on synthetic scan pairs (BASELINE config #3 / #4 shape): projection of the raw
scans -> normals -> tcgen05 encoder forward -> heads -> quaternion->T -> fused ICP losses with dL/dT ->
backward through heads and the tcgen05 encoder (dgrad / wgrad), whose weight gradients land in one flat buffer and
are all-reduced bucket by bucket over NCCL WHILE the backward still runs (WORLD_SIZE > 1, parallel_grad.py) -> Adam.  Used by bench.py (`train_step` key) and tests; mirrors Deployer.step
(src/deploy/deployer.py:237-342) + Trainer's optimizer (src/deploy/trainer.py:23-24) with the normals
computed in-line from the projected images (no preprocessed dataset on the bench box)."""
import torch

from . import ops
from .deploy.deployer import _FusedIcp
from .models.model import OdometryModel
from .models.model_parts import GeometryHandler
from .parallel_grad import make_grad_sync


class SyntheticTrainStep:
    def __init__(self, cfg, batch, n_max, dataset="kitti", use_tensor_cores=True, lr=1e-5, identity_init=True,
                 grad_sync="bucketed", autocast_bf16=False):
        self.cfg = dict(cfg)
        self.cfg.update({"pre_feature_extraction": False, "resnet_outputs": 1000, "use_dropout": False,
                         "layers": [2, 2, 2, 2], "factor_fewer_resnet_channels": 1, "activation_fct": "tanh",
                         "use_single_mlp_at_output": False, "use_tensor_core_encoder": bool(use_tensor_cores)})
        ds = self.cfg[dataset]
        self.B, self.N = int(batch), int(n_max)
        self.H, self.W = ds["vertical_cells"], ds["horizontal_cells"]
        self.hf, self.vf = tuple(self.cfg["horizontal_field_of_view"]), tuple(ds["vertical_field_of_view"])
        self.device = torch.device(self.cfg["device"])
        self.model = OdometryModel(self.cfg).to(self.device)
        if identity_init:
            # The reference first fits the identity transform (loss < 1e-2, src/deploy/trainer.py:184-186) and
            # only then switches to the geometric losses, so the unsupervised phase starts from T ~ I.
            # Reproduce that state directly: zero the last head layers, quaternion bias = (0, 0, 0, 1).
            with torch.no_grad():
                for head, bias in ((self.model.fully_connected_rotation, [0.0, 0.0, 0.0, 1.0]),
                                   (self.model.fully_connected_translation, [0.0, 0.0, 0.0])):
                    head[3].weight.zero_()
                    head[3].bias.copy_(torch.tensor(bias, device=self.device))
        self.autocast_bf16 = bool(autocast_bf16)       # torch / cuDNN path only: the reference stack under bf16 autocast
        # same optimizer as the reference (src/deploy/trainer.py:23-24); on CUDA torch's fused implementation (one
        # kernel for all parameters instead of ~7 multi-tensor launches, identical state_dict)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr, fused=self.device.type == "cuda")
        self.sync = make_grad_sync(self.model, grad_sync)
        self.scratch = ops.icp_scratch(self.B, self.H * self.W, self.device)
        self.points = torch.zeros((2 * self.B, 3, self.N), dtype=torch.float32, device=self.device)
        self.n_points = torch.zeros((2 * self.B,), dtype=torch.int32, device=self.device)

    def load(self, points, n_points):
        self.points.copy_(points, non_blocking=True)
        self.n_points.copy_(n_points, non_blocking=True)

    def step(self):
        b, h, w = self.B, self.H, self.W
        with ops.nvtx_range("projection"):
            image, _ = ops.project(self.points, self.n_points, h, w, self.hf, self.vf)
        with ops.nvtx_range("normals"):
            _, pts_grid, nrm_grid = ops.normals(image, self.cfg["kitti"]["neighborhood_side_length"],
                                                self.cfg["epsilon_range"],
                                                self.cfg["min_num_points_in_neighborhood_to_determine_point_class"],
                                                grids=True)
        self.optimizer.zero_grad(set_to_none=True)
        with ops.nvtx_range("model forward"):
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast_bf16):
                translations, quaternions = self.model(image_1=image[:b].contiguous(), image_2=image[b:].contiguous())
            translations, quaternions = translations.float(), quaternions.float()
            T = GeometryHandler.get_transformation_matrix_quaternion(translations, quaternions, self.device)
        with ops.nvtx_range("icp losses"):
            total, parts = _FusedIcp.apply(T, pts_grid[b:].contiguous(), nrm_grid[b:].contiguous(),
                                           pts_grid[:b].contiguous(), nrm_grid[:b].contiguous(),
                                           (h, w, self.hf, self.vf), float(self.cfg["lambda_po2pl"]),
                                           ops.LOSS_PO2PL | ops.LOSS_PL2PL, self.scratch)
            loss = total.mean()
        with ops.nvtx_range("backward (+ gradient buckets)"):
            loss.backward()
            self.sync.finish()
        with ops.nvtx_range("optimizer"):
            self.optimizer.step()
        return loss.detach(), parts

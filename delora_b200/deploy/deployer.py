"""`deploy.deployer.Deployer` re-designed for the GPU (reference: src/deploy/deployer.py).

Same constructor and `step(preprocessed_dicts, epoch_losses, log_images_bool)` contract, but the
per-sample Python loops with host round trips (:245-268 projection, :290-312 transform + losses)
become one batched device-side pass:

    pad/stack the B scan pairs -> ONE projection launch for the 2B scans -> gather the per-point
    normals through the pixel->point index map -> encoder -> quaternion -> T (CUDA fwd/bwd) ->
    ONE fused ICP kernel (SE(3) transform, exact NN, po2pl/pl2pl(/po2po), gradient w.r.t. T).

Reference quirks kept on purpose, each behind a config key:
  * `loss_pc` accumulates a running cumulative sum inside the batch loop (:312)
    -> sample j (0-based) gets weight (B - j)/B.  `plain_batch_mean: True` switches to 1/B.
  * identity pre-training compares only the LAST sample's transform with I (:324-327, :334-336).
"""
import torch

from .. import ops
from ..data import batching
from ..data import dataset as dataset_module
from ..losses import icp_losses
from ..models import model as model_module
from ..models import model_parts
from ..utility import projection


class _FusedIcp(torch.autograd.Function):
    """losses(T) for a batch of pairs on dense grids; backward = the kernel's own dL/dT."""

    @staticmethod
    def forward(ctx, transforms, src_grid, src_ngrid, tgt_grid, tgt_ngrid, grid, lam, flags, scratch):
        h, w, hf, vf = grid
        t12 = transforms[:, :3, :].reshape(-1, 12).detach().float().contiguous()
        losses, grad_t = ops.icp_dense_fwd_bwd(src_grid, src_ngrid, t12, tgt_grid, tgt_ngrid, h, w, hf, vf,
                                               lambda_po2pl=lam, flags=flags, scratch=scratch)
        ctx.save_for_backward(grad_t)
        total = losses[:, 0] + lam * losses[:, 1] + losses[:, 2]            # deployer.py:309-312
        ctx.mark_non_differentiable(losses)
        return total, losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        (grad_t,) = ctx.saved_tensors
        b = grad_t.shape[0]
        g = torch.zeros((b, 4, 4), dtype=torch.float32, device=grad_t.device)
        g[:, :3, :] = grad_t.view(b, 3, 4) * g_total.view(b, 1, 1)
        return g, None, None, None, None, None, None, None, None


class Deployer(object):

    def __init__(self, config):
        self.config = config
        self.device = config["device"]
        self.batch_size = config["batch_size"]
        self.dataset = dataset_module.PreprocessedPointCloudDataset(config=config)
        self.steps_per_epoch = int(len(self.dataset) / self.batch_size)
        self.img_projection = projection.ImageProjectionLayer(config=config)
        self.model = model_module.OdometryModel(config=self.config).to(self.device)
        self.geometry_handler = model_parts.GeometryHandler(config=config)
        self.lossTransformation = torch.nn.MSELoss()
        self.lossPointCloud = icp_losses.ICPLosses(config=self.config)      # drop-in operator (used by callers)
        self.training_bool = False
        self._scratch = {}
        self.log_img_1, self.log_img_2 = [], []

    @staticmethod
    def list_collate(batch_dicts):
        return [batch_dict for batch_dict in batch_dicts]

    # ---- reference helpers kept for API compatibility (src/deploy/deployer.py:181-189) ----------
    def rotate_point_cloud_transformation_matrix(self, transformation_matrix, point_cloud):
        return transformation_matrix[:, :3, :3].matmul(point_cloud[:, :3, :])

    def transform_point_cloud_transformation_matrix(self, transformation_matrix, point_cloud):
        out = self.rotate_point_cloud_transformation_matrix(transformation_matrix, point_cloud)
        return out + transformation_matrix[:, :3, 3].view(-1, 3, 1)

    def normalize_input(self, preprocessed_data):                           # :222-235
        means = [torch.mean(torch.norm(preprocessed_data[k], dim=1), dim=1, keepdim=True) for k in ("scan_1", "scan_2")]
        normalization_mean = torch.mean(torch.cat(means, dim=1), dim=1)
        preprocessed_data["scan_1"] /= normalization_mean
        preprocessed_data["scan_2"] /= normalization_mean
        preprocessed_data["scaling_factor"] = normalization_mean
        return preprocessed_data, normalization_mean

    # ---- batched device-side collate ---------------------------------------------------------
    def _stack(self, dicts):
        dev = torch.device(self.device)
        b = len(dicts)
        n_max = max(max(d["scan_1"].shape[2], d["scan_2"].shape[2]) for d in dicts)
        pts = torch.zeros((2 * b, 3, n_max), dtype=torch.float32, device=dev)
        nrm = torch.zeros((2 * b, 3, n_max), dtype=torch.float32, device=dev)
        cnt = torch.zeros((2 * b,), dtype=torch.int32)
        for i, d in enumerate(dicts):
            for half, key_s, key_n in ((0, "scan_1", "normal_list_1"), (1, "scan_2", "normal_list_2")):
                n = d[key_s].shape[2]
                pts[half * b + i, :, :n] = d[key_s][0].to(dev)
                nrm[half * b + i, :, :n] = d[key_n][0].to(dev)
                cnt[half * b + i] = n
        return pts, nrm, cnt.to(dev, non_blocking=True)

    def _flags(self):
        f = 0
        if self.config["point_to_point_loss"]:
            f |= ops.LOSS_PO2PO
        if self.config["point_to_plane_loss"]:
            f |= ops.LOSS_PO2PL
        if self.config["plane_to_plane_loss"]:
            f |= ops.LOSS_PL2PL | (ops.NORMAL_LINEAR if self.config["normal_loss"] == "linear" else 0)
        return f

    def _normalize_batch(self, pts, counts_host, b):
        """`normalize_input` (:222-235) for a padded batch: per sample, the mean of the two scans' mean ranges --
        each mean reduced over exactly that scan's points, like the reference (the projection downstream is
        discrete, so the factor must not differ in the last bit)."""
        means = torch.stack([torch.mean(torch.norm(pts[i:i + 1, :, :n], dim=1), dim=1, keepdim=True)
                             for i, n in enumerate(counts_host)])                      # [2B,1,1]
        factor = torch.stack([torch.mean(torch.cat((means[i], means[b + i]), dim=1), dim=1) for i in range(b)]).view(b)
        pts /= torch.cat((factor, factor)).view(2 * b, 1, 1)
        return factor

    def step(self, preprocessed_dicts, epoch_losses=None, log_images_bool=False):
        """`preprocessed_dicts`: the reference's list of per-sample dicts, or a `data.batching.PaddedBatch`
        (already on the device: no per-sample work at all)."""
        padded = isinstance(preprocessed_dicts, batching.PaddedBatch)
        b = len(preprocessed_dicts)
        dataset = preprocessed_dicts.dataset if padded else preprocessed_dicts[0]["dataset"]
        ds = self.config[dataset]
        h, w = ds["vertical_cells"], ds["horizontal_cells"]
        hf, vf = self.config["horizontal_field_of_view"], ds["vertical_field_of_view"]
        scaling = None
        if padded:
            dev = torch.device(self.device)
            batch = preprocessed_dicts if preprocessed_dicts.flat.device == dev else preprocessed_dicts.to(dev)
            pts, nrm, cnt = batch.points, batch.normals, batch.counts
            if self.config["normalization_scaling"]:
                scaling = self._normalize_batch(pts, batch.counts_host, b)
            n_last = batch.counts_host[2 * b - 1]
        else:
            if self.config["normalization_scaling"]:
                for d in preprocessed_dicts:
                    self.normalize_input(preprocessed_data=d)
                scaling = torch.cat([d["scaling_factor"].reshape(1) for d in preprocessed_dicts]).to(self.device)
            pts, nrm, cnt = self._stack(preprocessed_dicts)
            n_last = int(preprocessed_dicts[-1]["scan_2"].shape[2])
        image, index_map = self.img_projection.project_batch(pts, cnt, dataset)    # [2B,4,H,W]
        images_model_1, images_model_2 = image[:b], image[b:]
        self.log_img_1, self.log_img_2 = images_model_1[-1:, :3], images_model_2[-1:, :3]

        translations, rotation_representation = self.model(image_1=images_model_1, image_2=images_model_2)
        computed_transformations = self.geometry_handler.get_transformation_matrix_quaternion(
            translation=translations, quaternion=rotation_representation, device=self.device)

        if self.config["inference_only"]:
            if scaling is not None:
                computed_transformations[:, :3, 3] *= scaling.view(b, 1)
            return computed_transformations

        pts_grid, nrm_grid = ops.grids_from_projection(pts, nrm, index_map)
        key = (b, h * w)
        if key not in self._scratch:
            self._scratch[key] = ops.icp_scratch(b, h * w, pts.device)
        total, parts = _FusedIcp.apply(computed_transformations, pts_grid[b:].contiguous(), nrm_grid[b:].contiguous(),
                                       pts_grid[:b].contiguous(), nrm_grid[:b].contiguous(), (h, w, hf, vf),
                                       float(self.config["lambda_po2pl"]), self._flags(), self._scratch[key])
        if self.config.get("plain_batch_mean", False):
            weights = torch.full((b,), 1.0 / b, device=total.device)
        else:   # running cumulative sum of the reference (:312): sample j is counted (B - j) times, then / B
            weights = torch.arange(b, 0, -1, device=total.device, dtype=torch.float32) / b
        losses = {"loss_pc": (weights * total).sum().reshape(1),
                  "loss_po2po": parts[:, 0].sum().reshape(1) / b,
                  "loss_po2pl": float(self.config["lambda_po2pl"]) * parts[:, 1].sum().reshape(1) / b,
                  "loss_pl2pl": parts[:, 2].sum().reshape(1) / b}
        if not self.config["unsupervised_at_start"]:                        # identity fitting (:324-327, :334-336)
            eye = torch.eye(4, device=total.device).view(1, 4, 4)
            loss = self.lossTransformation(input=computed_transformations[-1:], target=eye) / b
        else:
            loss = losses["loss_pc"]
        if self.training_bool:
            loss.sum().backward()
            self.optimizer.step()
        if scaling is not None:
            computed_transformations[:, :3, 3] *= scaling.view(b, 1)
        if epoch_losses is not None:
            # `visible_pixels` (:365-367): points of the last transformed source scan with 0 < v < H
            with torch.no_grad():
                src = pts[2 * b - 1:2 * b, :, :n_last]
                moved = self.transform_point_cloud_transformation_matrix(computed_transformations[-1:].detach(), src)
                _, v_pix, _ = ops.project_uv(moved.contiguous(), cnt[2 * b - 1:2 * b].contiguous(), h, w, hf, vf)
                visible = ((torch.round(v_pix[0, :n_last]) < h) & (v_pix[0, :n_last] > 0)).sum()
            epoch_losses["loss_epoch"] += loss.detach().cpu().numpy()
            epoch_losses["loss_point_cloud_epoch"] += losses["loss_pc"].detach().cpu().numpy()
            epoch_losses["loss_po2po_epoch"] += losses["loss_po2po"].detach().cpu().numpy()
            epoch_losses["loss_po2pl_epoch"] += losses["loss_po2pl"].detach().cpu().numpy()
            epoch_losses["loss_pl2pl_epoch"] += losses["loss_pl2pl"].detach().cpu().numpy()
            epoch_losses["visible_pixels_epoch"] += float(visible)
            return epoch_losses, computed_transformations
        return losses, computed_transformations

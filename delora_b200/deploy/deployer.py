"""`deploy.deployer.Deployer` re-designed for the GPU (reference: src/deploy/deployer.py).

Same constructor and `step(preprocessed_dicts, epoch_losses, log_images_bool)` contract, but the
per-sample Python loops with host round trips (:245-268 projection, :290-312 transform + losses)
become one batched device-side pass:

    pad/stack the B scan pairs -> ONE projection launch for the 2B scans -> gather the per-point
    normals through the pixel->point index map -> encoder -> quaternion -> T (CUDA fwd/bwd) ->
    ONE fused ICP kernel (SE(3) transform, exact NN, po2pl/pl2pl(/po2po), gradient w.r.t. T).

Logging / visualisation (reference :73-89 `create_images`, :316-320 `log_img_2_transformed`,
:349-356): only when `log_images_bool` is set -- once per epoch in the reference's Trainer -- the first sample is
additionally run through the list-based ICPLosses operator (kept pairs, pointwise residuals) and its 6- and
9-channel clouds are projected on the device; a normal step does no extra projection beyond the `visible_pixels`
coordinates (no image, no sort).

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
        self.log_img_2_transformed = self.log_pointwise_loss = None
        self.log_normals_target = self.log_normals_transformed_source = None

    @staticmethod
    def list_collate(batch_dicts):
        return [batch_dict for batch_dict in batch_dicts]

    def create_images(self, preprocessed_data, losses, plotting):
        """src/deploy/deployer.py:73-89: images of the target normals and of the transformed source points /
        normals / pointwise point-to-plane residuals at the kept pairs -- two device-side projections of a
        6-channel and a 9-channel cloud (the reference: numba + host round trips)."""
        image_1_at_normals, _, _, _, _ = self.img_projection(
            input=torch.cat((preprocessed_data["scan_1"], preprocessed_data["normal_list_1"]), dim=1),
            dataset=preprocessed_data["dataset"])
        image_2, _, _, _, _ = self.img_projection(
            input=torch.cat((plotting["scan_2_transformed"], plotting["normals_2_transformed"],
                             losses["loss_po2pl_pointwise"]), dim=1).detach(),
            dataset=preprocessed_data["dataset"])
        self.log_pointwise_loss = image_2[:, 6:9]
        self.log_normals_target = image_1_at_normals[:, 3:6]
        self.log_normals_transformed_source = image_2[:, 3:6]

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
        if self.config.get("po2po_alone", False):       # :36-46: every source point against its NN, no normals
            return ops.LOSS_PO2PO
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
        if self.config.get("po2po_alone", False):
            # the fused kernel takes point-to-point pairs where neither side has a normal: without normals that is
            # every source point, the reference's po2po_alone branch (src/losses/icp_losses.py:36-46)
            nrm_grid = torch.zeros_like(nrm_grid)
        key = (b, h * w)
        if key not in self._scratch:
            self._scratch[key] = ops.icp_scratch(b, h * w, pts.device)
        total, parts = _FusedIcp.apply(computed_transformations, pts_grid[b:].contiguous(), nrm_grid[b:].contiguous(),
                                       pts_grid[:b].contiguous(), nrm_grid[:b].contiguous(), (h, w, hf, vf),
                                       float(self.config["lambda_po2pl"]), self._flags(), self._scratch[key])
        # the reference divides by the CONFIGURED batch size (:329-338), also for a short last batch
        bs = float(self.batch_size)
        if self.config.get("plain_batch_mean", False):
            weights = torch.full((b,), 1.0 / bs, device=total.device)
        else:   # running cumulative sum of the reference (:312): sample j is counted (B - j) times, then / batch_size
            weights = torch.arange(b, 0, -1, device=total.device, dtype=torch.float32) / bs
        losses = {"loss_pc": (weights * total).sum().reshape(1),
                  "loss_po2po": parts[:, 0].sum().reshape(1) / bs,
                  "loss_po2pl": float(self.config["lambda_po2pl"]) * parts[:, 1].sum().reshape(1) / bs,
                  "loss_pl2pl": parts[:, 2].sum().reshape(1) / bs}
        if not self.config["unsupervised_at_start"]:                        # identity fitting (:324-327, :334-336)
            eye = torch.eye(4, device=total.device).view(1, 4, 4)
            loss = self.lossTransformation(input=computed_transformations[-1:], target=eye) / bs
        else:
            loss = losses["loss_pc"]
        if self.training_bool:
            loss.sum().backward()
            self.optimizer.step()
        if epoch_losses is not None:
            # `visible_pixels` (:365-367): points of a transformed source scan with 0 < v < H, projected with the
            # transform BEFORE the translation is rescaled (:344-346 comes after the transform of :294 in the reference).
            # Normal steps: the LAST sample (the reference's loop variable, :349-352); log steps: the FIRST (:317-319).
            with torch.no_grad():
                k = 0 if log_images_bool else b - 1
                n_k = int((batch.counts_host if padded else [d["scan_1"].shape[2] for d in preprocessed_dicts]
                           + [d["scan_2"].shape[2] for d in preprocessed_dicts])[b + k])
                src = pts[b + k:b + k + 1, :, :n_k]
                t_k = computed_transformations[k:k + 1].detach()
                moved = self.transform_point_cloud_transformation_matrix(t_k, src).contiguous()
                if log_images_bool and not self.config.get("po2po_alone", False):
                    # logging path (once per epoch): image of the transformed source scan (:316-320), then the list-
                    # based loss operator on sample 0 for the kept pairs and the pointwise residuals (:302-307), then
                    # the two multi-channel projections of create_images (:353-356)
                    self.log_img_2_transformed, _, v_all, _, _ = self.img_projection(input=moved, dataset=dataset)
                    v_pix = v_all
                    n_1 = int((batch.counts_host if padded else [d["scan_1"].shape[2] for d in preprocessed_dicts])[0])
                    tgt, tgt_n = pts[0:1, :, :n_1].contiguous(), nrm[0:1, :, :n_1].contiguous()
                    src_n = self.rotate_point_cloud_transformation_matrix(t_k, nrm[b:b + 1, :, :n_k]).contiguous()
                    losses_0, plotting_0 = self.lossPointCloud(
                        source_point_cloud_transformed=moved, source_normal_list_transformed=src_n,
                        target_point_cloud=tgt, target_normal_list=tgt_n, compute_pointwise_loss_bool=True)
                    losses["loss_po2pl_pointwise"] = losses_0["loss_po2pl_pointwise"]
                    self.create_images(preprocessed_data={"scan_1": tgt, "normal_list_1": tgt_n, "dataset": dataset},
                                       losses=losses, plotting=plotting_0)
                else:
                    _, v_pix, _ = ops.project_uv(moved, cnt[b + k:b + k + 1].contiguous(), h, w, hf, vf)
                visible = ((torch.round(v_pix[0, :n_k]) < h) & (v_pix[0, :n_k] > 0)).sum()
        if scaling is not None:
            computed_transformations[:, :3, 3] *= scaling.view(b, 1)
        if epoch_losses is not None:
            epoch_losses["loss_epoch"] += loss.detach().cpu().numpy()
            epoch_losses["loss_point_cloud_epoch"] += losses["loss_pc"].detach().cpu().numpy()
            epoch_losses["loss_po2po_epoch"] += losses["loss_po2po"].detach().cpu().numpy()
            epoch_losses["loss_po2pl_epoch"] += losses["loss_po2pl"].detach().cpu().numpy()
            epoch_losses["loss_pl2pl_epoch"] += losses["loss_pl2pl"].detach().cpu().numpy()
            epoch_losses["visible_pixels_epoch"] += float(visible)
            return epoch_losses, computed_transformations
        return losses, computed_transformations

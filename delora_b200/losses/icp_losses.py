"""Drop-in `losses.icp_losses.ICPLosses` (reference: src/losses/icp_losses.py).

`forward(source_point_cloud_transformed, source_normal_list_transformed, target_point_cloud,
target_normal_list, compute_pointwise_loss_bool) -> (losses, plotting)` with the reference's
dict keys and gradient flow (to the two source tensors only).  The cKDTree build + two queries
(:34, :70-80), the boolean-mask compactions (:55-60, :114-121) and the three sub-losses
(:168-179, :196-206, :224-240) are one counting-sort (spherical cell grid) + one fused CUDA
kernel; the nearest neighbours are the exact float64 Euclidean NN, as cKDTree returns them.
"""
import torch

from .. import ops


class _IcpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, src_n, tgt, tgt_n, grid, flags):
        h, w, hf, vf = grid
        dev = src.device
        ns = torch.tensor([src.shape[2]], dtype=torch.int32, device=dev)
        nt = torch.tensor([tgt.shape[2]], dtype=torch.int32, device=dev)
        s4, sn4 = ops.pack_lists(src.detach().float().contiguous(), src_n.detach().float().contiguous(), ns)
        t4, tn4, cs = ops.grid_build(tgt.detach().float().contiguous(), tgt_n.detach().float().contiguous(), nt, h, w,
                                     hf, vf)
        losses, _, nn_index, pdir, ndir = ops.icp_fwd_bwd(s4, sn4, ns, None, t4, tn4, cs, h, w, hf, vf,
                                                          flags=flags, pointwise=True)
        ctx.save_for_backward(pdir, ndir, ns, losses)
        ctx.mark_non_differentiable(nn_index)
        return losses[0, 0:1].clone(), losses[0, 1:2].clone(), losses[0, 2:3].clone(), nn_index, pdir

    @staticmethod
    def backward(ctx, g_po2po, g_po2pl, g_pl2pl, _g_idx, _g_dir):
        pdir, ndir, ns, losses = ctx.saved_tensors
        up = torch.cat((g_po2po.reshape(1), g_po2pl.reshape(1), g_pl2pl.reshape(1))).float().reshape(1, 3).contiguous()
        gp, gn = ops.icp_point_grads(pdir, ndir, ns, losses, up)
        return gp, gn, None, None, None, None


class ICPLosses(torch.nn.Module):

    def __init__(self, config):
        super().__init__()
        self.config = config
        if self.config["plane_to_plane_loss"] and self.config["normal_loss"] not in ("linear", "squared"):
            raise Exception("The normal loss which is defined here is not admissible.")
        ds = self.config[self.config["datasets"][0]]
        # any spherical grid gives the exact NN; the sensor's own image grid keeps <= ~1 point per cell
        self.grid = (int(ds["vertical_cells"]), int(ds["horizontal_cells"]),
                     tuple(self.config["horizontal_field_of_view"]), tuple(ds["vertical_field_of_view"]))

    def _flags(self):
        f = 0
        if self.config["point_to_point_loss"]:
            f |= ops.LOSS_PO2PO
        if self.config["point_to_plane_loss"]:
            f |= ops.LOSS_PO2PL
        if self.config["plane_to_plane_loss"]:
            f |= ops.LOSS_PL2PL
            if self.config["normal_loss"] == "linear":
                f |= ops.NORMAL_LINEAR
        return f

    def forward(self, source_point_cloud_transformed, source_normal_list_transformed, target_point_cloud,
                target_normal_list, compute_pointwise_loss_bool):
        if self.config["po2po_alone"]:
            # src/losses/icp_losses.py:36-46: every source point against its NN, no normals involved
            zeros_s = torch.zeros_like(source_point_cloud_transformed)
            zeros_t = torch.zeros_like(target_point_cloud)
            loss_po2po, _, _, _, _ = _IcpFunction.apply(source_point_cloud_transformed, zeros_s, target_point_cloud,
                                                        zeros_t, self.grid, ops.LOSS_PO2PO)
            dev = loss_po2po.device
            losses = {"loss_po2po": loss_po2po, "loss_po2pl": torch.zeros(1, device=dev),
                      "loss_po2pl_pointwise": torch.zeros(1, device=dev), "loss_pl2pl": torch.zeros(1, device=dev)}
            return losses, None
        loss_po2po, loss_po2pl, loss_pl2pl, nn_index, pdir = _IcpFunction.apply(
            source_point_cloud_transformed, source_normal_list_transformed, target_point_cloud, target_normal_list,
            self.grid, self._flags())
        dev = loss_po2pl.device
        kept = pdir[0, :, 3] == 1.0                                     # (source normal) & (target normal): :110-121
        source_points_where_normals = source_point_cloud_transformed[:, :, kept]
        source_normals_where_normals = source_normal_list_transformed[:, :, kept]
        pointwise = torch.zeros(1, device=dev)
        if compute_pointwise_loss_bool and self.config["point_to_plane_loss"]:
            tgt_kept = target_point_cloud[:, :, nn_index[0, kept].long()]
            pointwise = (source_points_where_normals - tgt_kept)        # :197, returned un-detached like the reference
        losses = {
            "loss_po2po": loss_po2po if self.config["point_to_point_loss"] else torch.zeros(1, device=dev),
            "loss_po2pl": loss_po2pl if self.config["point_to_plane_loss"] else torch.zeros(1, device=dev),
            "loss_po2pl_pointwise": pointwise,
            "loss_pl2pl": loss_pl2pl if self.config["plane_to_plane_loss"] else torch.zeros(1, device=dev),
        }
        plotting = {"scan_2_transformed": source_points_where_normals,
                    "normals_2_transformed": source_normals_where_normals}
        return losses, plotting

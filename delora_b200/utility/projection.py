"""Drop-in `utility.projection.ImageProjectionLayer` (reference: src/utility/projection.py).

Same constructor, same `forward(input, dataset)` signature, same five return values, backed by
the CUDA projection / (u,v) / radix-sort kernels of libdelora_b200.so.  Differences, all
deliberate:
  * equal-range ties inside a pixel are resolved deterministically (lowest point index) where the
    reference's unstable `torch.argsort` is arbitrary (:63);
  * nothing leaves the GPU except the survivor count (the reference round-trips the whole index
    list through numba on the host, :84-91).
"""
import torch

from .. import ops


class ImageProjectionLayer(torch.nn.Module):

    def __init__(self, config):
        super().__init__()
        self.device = config["device"]
        self.config = config
        self.horizontal_field_of_view = config["horizontal_field_of_view"]

    def _dims(self, dataset):
        ds = self.config[dataset]
        return ds["vertical_cells"], ds["horizontal_cells"], ds["vertical_field_of_view"]

    def project_batch(self, points, n_points, dataset):
        """Batched core: points [B,C,N] fp32 (padded), n_points [B] int32 -> image [B,C+1,H,W],
        index_map [B,H,W] int32.  This is what the training step uses (one launch per batch)."""
        h, w, vfov = self._dims(dataset)
        return ops.project(points, n_points, h, w, self.horizontal_field_of_view, vfov)

    def project_to_img(self, point_cloud, dataset):
        """src/utility/projection.py:48-106.  point_cloud [1,C,N] -> (image [1,C+1,H,W],
        u [1,N], v [1,N] (range-sorted, un-rounded), point_indices [K] int64 ascending range,
        image_to_pointcloud_indices [1,K,2] int64 (v,u))."""
        h, w, vfov = self._dims(dataset)
        dev = torch.device(self.device)
        pts = point_cloud.detach().to(device=dev, dtype=torch.float32).contiguous()
        if pts.dim() != 3 or pts.shape[0] != 1:
            raise Exception("ImageProjectionLayer expects a [1, C, N] point cloud (the reference only uses batch "
                            "element 0: src/utility/projection.py:67,78)")
        n = pts.shape[2]
        n_dev = torch.tensor([n], dtype=torch.int32, device=dev)
        hf = self.horizontal_field_of_view
        image, index_map = ops.project(pts, n_dev, h, w, hf, vfov)
        u, v, rng = ops.project_uv(pts, n_dev, h, w, hf, vfov)
        order = ops.sort_by_range(rng, n_dev)[0].long()            # ascending (range, index): :63-67
        # survivors = points that own a pixel, listed in range order (:94-96, :105)
        imap = index_map[0].reshape(-1)
        pix = torch.nonzero(imap >= 0)[:, 0]
        pixel_of_point = torch.full((n,), -1, dtype=torch.long, device=dev)
        pixel_of_point[imap[pix].long()] = pix
        sorted_pixels = pixel_of_point[order]
        keep = sorted_pixels >= 0
        point_indices = order[keep]
        kept_pix = sorted_pixels[keep]
        image_to_pointcloud_indices = torch.stack((kept_pix // w, kept_pix % w), dim=1)[None]
        return image, u[:, order], v[:, order], point_indices, image_to_pointcloud_indices

    def forward(self, input, dataset):
        return self.project_to_img(point_cloud=input, dataset=dataset)

"""Drop-in `preprocessing.normal_computation.NormalsComputer`
(reference: src/preprocessing/normal_computation.py; covariance: src/utility/linalg.py:33-56).

`compute_normal_vectors(image)` returns the same triple (normals [P,3], has_normal [P] bool,
points [P,3]) for the P valid pixels in row-major order; the 7x11 (or configured) edge-clamped
patch, range gate, >= min-neighbour rule, smallest-eigenvector and sensor-facing flip all run in
one CUDA kernel, followed by a scan-based compaction.  No host LAPACK, no 77-step Python loop.
"""
import torch

from .. import ops


class NormalsComputer:
    def __init__(self, config, dataset_name):
        self.config = config
        self.dataset_name = dataset_name

    def _params(self):
        return (tuple(self.config[self.dataset_name]["neighborhood_side_length"]),
                float(self.config["epsilon_range"]),
                int(self.config["min_num_points_in_neighborhood_to_determine_point_class"]))

    def compute_normal_images(self, image):
        """Batched core: image [B,C,H,W] -> normals [B,3,H,W] (zeros where no normal)."""
        nb, eps, min_nb = self._params()
        return ops.normals(image, nb, eps, min_nb)

    def compute_normal_vectors(self, image):
        dev = torch.device(self.config["device"])
        img = image.detach().to(device=dev, dtype=torch.float32).contiguous()
        h = self.config[self.dataset_name]["vertical_cells"]
        w = self.config[self.dataset_name]["horizontal_cells"]
        if img.shape[2] != h or img.shape[3] != w:
            raise Exception("image size does not match the configured vertical/horizontal cells")
        img = img[:1]                                              # the reference reads image[0] only (:33, :94)
        nrm = self.compute_normal_images(img)
        pts4, nrm4, _, counts = ops.lists_from_images(img, nrm)
        p = int(counts[0])                                         # data-dependent output size: one host sync
        normals = nrm4[0, :p, :3].contiguous()
        has_normal = nrm4[0, :p, 3] != 0
        points = pts4[0, :p, :3].contiguous()
        return normals, has_normal, points

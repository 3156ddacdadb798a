"""Drop-in `preprocessing.preprocesser.Preprocesser` (reference: src/preprocessing/preprocesser.py).

Offline stage of the reference: raw scan [1,4,N] -> projection at `horizontal_cells_preprocessing`
(KITTI: 64 x 2250) -> per-pixel normals -> `<preprocessed_path>/<seq>/{scans,normals}/%06d.npy` with
the [P,3] point list and the [P,3] normals of the valid pixels in row-major order (:50-68).  The
reference spends ~1-2 s per scan here (argsort + numba de-duplication on the host, a 77-step Python
neighbour gather, LAPACK `symeig` of P 3x3 matrices); this one runs the projection, normals and
compaction kernels of libdelora_b200.so for a whole batch of scans per launch and only the .npy
writes stay on the host.

`apply_preprocessing_step(scan, index)` keeps the reference's per-scan signature;
`apply_preprocessing_batch(scans, indices)` is the batched form the KITTI walker uses.
Rosbag input and the matplotlib preview (`visualize_single_img_preprocessing`) are out of scope
(SURVEY.md §2): asking for them raises.
"""
import os

import numpy as np
import torch

from .. import ops
from ..data import kitti_scans
from ..utility import projection
from . import normal_computation


class Preprocesser:

    def __init__(self, config):
        self.config = config
        self.device = config["device"]
        self.img_projection = projection.ImageProjectionLayer(config=config)
        self.batch_size = int(config.get("preprocessing_batch_size", 16))

    def ensure_dir(self, file_path):
        directory = os.path.dirname(file_path)
        if not os.path.exists(directory):
            os.makedirs(directory)

    def _save(self, index, normals, points):
        np.save(os.path.join(self.normals_name, format(int(index), '06d') + ".npy"), normals)
        np.save(os.path.join(self.scans_name, format(int(index), '06d') + ".npy"), points)

    def preprocess_scans(self, scans):
        """Core: list of [C,N_i] (or [1,C,N_i]) clouds -> list of (normals [P_i,3], points [P_i,3]) numpy arrays."""
        if self.config.get("visualize_single_img_preprocessing", False):
            raise Exception("visualize_single_img_preprocessing is not supported by the B200 preprocesser")
        dev = torch.device(self.device)
        dataset = self.config["dataset"]
        clouds = [s[0] if s.dim() == 3 else s for s in scans]
        c = clouds[0].shape[0]
        n_max = max(int(s.shape[1]) for s in clouds)
        b = len(clouds)
        host = torch.zeros((b, c, n_max), dtype=torch.float32).pin_memory()
        counts = torch.empty((b,), dtype=torch.int32)
        for i, s in enumerate(clouds):
            host[i, :, :s.shape[1]] = s
            counts[i] = s.shape[1]
        pts = host.to(dev, non_blocking=True)
        n_dev = counts.to(dev)
        image, _ = self.img_projection.project_batch(pts, n_dev, dataset)            # closest point per pixel
        nrm_img = self.normals_computer.compute_normal_images(image)
        pts4, nrm4, _, valid = ops.lists_from_images(image, nrm_img)                  # row-major valid pixels
        valid = valid.cpu()
        pts4, nrm4 = pts4.cpu(), nrm4.cpu()
        out = []
        for i in range(b):
            p = int(valid[i])
            out.append((nrm4[i, :p, :3].contiguous().numpy(), pts4[i, :p, :3].contiguous().numpy()))
        return out

    def apply_preprocessing_batch(self, scans, indices):
        for (normals, points), index in zip(self.preprocess_scans(scans), indices):
            self._save(index, normals, points)

    def apply_preprocessing_step(self, scan, index):
        self.apply_preprocessing_batch([scan], [index])

    def preprocess_data(self):
        for index_of_dataset, dataset_name in enumerate(self.config["datasets"]):
            self.config["dataset"] = dataset_name
            self.config[dataset_name]["horizontal_cells"] = self.config[dataset_name]["horizontal_cells_preprocessing"]
            self.normals_computer = normal_computation.NormalsComputer(config=self.config, dataset_name=dataset_name)
            for index_of_sequence, data_identifier in enumerate(self.config[dataset_name]["data_identifiers"]):
                self.config[dataset_name]["data_identifier"] = data_identifier
                ident = format(data_identifier, '02d') if isinstance(data_identifier, int) else str(data_identifier)
                name = os.path.join(self.config[dataset_name]["preprocessed_path"], ident + "/")
                self.normals_name = os.path.join(name, "normals/")
                self.ensure_dir(file_path=self.normals_name)
                self.scans_name = os.path.join(name, "scans/")
                self.ensure_dir(file_path=self.scans_name)
                if self.config[dataset_name]["dataset_type"] == "kitti":
                    kitti_scans.KITTIDatasetPreprocessor(config=self.config, dataset_name=dataset_name,
                                                         preprocessing_fct=self.apply_preprocessing_step,
                                                         preprocessing_fct_batch=self.apply_preprocessing_batch,
                                                         batch_size=self.batch_size).preprocess()
                elif self.config[dataset_name]["dataset_type"] == "rosbag":
                    raise Exception('Dataset type "rosbag" is outside the B200 hot path (needs ROS); convert the bag '
                                    'to KITTI-style .bin scans first.')
                else:
                    raise Exception('Dataset type not yet supported. Currently only "kitti" and "rosbag" available.')

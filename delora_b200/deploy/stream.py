"""Streaming odometry inference: the per-frame path of `Tester.test_dataset` / the ROS node
(reference: src/deploy/tester.py:38-107 -> src/deploy/deployer.py:237-281, :370-375;
src/ros_utils/odometry_publisher.py:137-147), one frame at a time, batch 1.

Per frame the reference projects BOTH scans of the pair again (the previous frame's projection is recomputed),
runs the encoder through cuDNN and synchronises with the host three times.  Here a frame costs ONE projection
(the previous frame's range image is kept on the device), the encoder forward on the tcgen05 kernels and the
quaternion -> T kernel, captured once in a CUDA graph (25 launches replayed as one) between a pinned-memory
H2D of the raw scan and a 64-byte D2H of the relative transform.  BASELINE config #5.
"""
import numpy as np
import torch

from .. import ops
from ..models import model_parts
from ..utility import poses as poses_module


class OdometryStream:
    def __init__(self, model, config, dataset, n_max, use_cuda_graph=True):
        self.model = model.eval()
        self.config = config
        self.dataset = dataset
        ds = config[dataset]
        self.h, self.w = int(ds["vertical_cells"]), int(ds["horizontal_cells"])
        self.hf, self.vf = config["horizontal_field_of_view"], ds["vertical_field_of_view"]
        self.device = torch.device(config["device"])
        if self.device.type != "cuda":
            raise RuntimeError("OdometryStream runs on CUDA devices only (no CPU fallback)")
        self.n_max = int(n_max)
        self.host_scan = torch.zeros((1, 3, self.n_max), dtype=torch.float32).pin_memory()
        self.host_count = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.host_T = torch.zeros((1, 4, 4), dtype=torch.float32).pin_memory()
        self._host_scan_np, self._host_count_np = self.host_scan.numpy(), self.host_count.numpy()
        self.points = torch.zeros((1, 3, self.n_max), dtype=torch.float32, device=self.device)
        self.count = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.prev_image = torch.zeros((1, 4, self.h, self.w), dtype=torch.float32, device=self.device)
        self.T = torch.zeros((1, 4, 4), dtype=torch.float32, device=self.device)
        self.frames = 0
        self.relative = []                       # [1,4,4] numpy per frame pair, like Tester's lists
        self.use_cuda_graph = bool(use_cuda_graph)
        self.graph = None
        # the model's own configuration decides the encoder path (tensor cores by default when eligible,
        # models/model.py); the stream never edits the shared config dict

    # everything of a frame that runs on the device; static shapes and buffers (graph-capturable)
    def _frame(self):
        image, _ = ops.project(self.points, self.count, self.h, self.w, self.hf, self.vf)     # closest point per pixel
        with torch.no_grad():
            translation, quaternion = self.model(image_1=self.prev_image, image_2=image)
            t = model_parts.GeometryHandler.get_transformation_matrix_quaternion(
                translation=translation, quaternion=quaternion, device=self.device)
        self.T.copy_(t)
        self.prev_image.copy_(image)

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        saved = self.prev_image.clone()          # taken BEFORE the side stream is released: the warm-up frames write prev_image
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                   # warm-up: allocator, tensor-map cache, cuBLAS handles
                self._frame()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._frame()
        self.prev_image.copy_(saved)

    def push(self, scan):
        """scan: [3,N] or [1,3,N] float32 host tensor (N <= n_max).  Returns the relative transform
        T_{k-1,k} as a [1,4,4] numpy array, or None for the first frame of a stream."""
        scan = scan[0] if scan.dim() == 3 else scan
        n = int(scan.shape[1])
        if n > self.n_max:
            raise Exception("scan has more points than the stream was sized for")
        # plain memcpy into the pinned staging buffer: a torch copy of this size wakes the intra-op thread pool, whose
        # spinning workers can exhaust a container's CPU quota and stall the stream for tens of ms every 100 ms
        np.copyto(self._host_scan_np[0, :, :n], scan[:3].numpy())
        self._host_count_np[0] = n
        self.points.copy_(self.host_scan, non_blocking=True)
        self.count.copy_(self.host_count, non_blocking=True)
        first = self.frames == 0
        if first:
            # no previous frame yet: project only
            image, _ = ops.project(self.points, self.count, self.h, self.w, self.hf, self.vf)
            self.prev_image.copy_(image)
            torch.cuda.current_stream().synchronize()
            self.frames += 1
            return None
        tc = self.model._tensor_core_path()
        if tc is not None:
            tc._refresh_weights()                # bf16 filter copies follow the fp32 parameters (no-op when unchanged)
        if self.use_cuda_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._frame()
        self.host_T.copy_(self.T, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.frames += 1
        t = self.host_T.numpy().copy()
        self.relative.append(t)
        return t

    def poses(self):
        """World-frame KITTI poses of the stream so far (src/utility/poses.py:11-58)."""
        return poses_module.compute_poses(self.relative)

    def reset(self):
        self.frames = 0
        self.relative = []

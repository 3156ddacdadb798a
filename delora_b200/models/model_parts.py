"""Drop-in `models.model_parts` (reference: src/models/model_parts.py): `CircularPad` and
`GeometryHandler`.  The quaternion -> 4x4 step (kornia 0.3.0 `quaternion_to_rotation_matrix`,
(x, y, z, w) order, L2-normalised; src/models/model_parts.py:29-44) is one CUDA kernel forward and
one backward, wired into autograd."""
import torch

from .. import ops


class CircularPad(torch.nn.Module):
    def __init__(self, padding=(1, 1, 0, 0)):
        super().__init__()
        self.padding = padding

    def forward(self, input):
        return torch.nn.functional.pad(input=input, pad=self.padding, mode="circular")


class _QuatToT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, translation, quaternion):
        q = quaternion.detach().float().contiguous()
        t = translation.detach().float().contiguous()
        ctx.save_for_backward(q)
        return ops.quat_to_T(q, t)

    @staticmethod
    def backward(ctx, grad_t):
        (q,) = ctx.saved_tensors
        gq, gt = ops.quat_to_T_bwd(q, grad_t.float().contiguous())
        return gt, gq


class GeometryHandler:
    def __init__(self, config):
        self.device = config["device"]

    @staticmethod
    def quaternion_to_rot_matrix(quaternion):
        b = quaternion.shape[0]
        zero = torch.zeros((b, 3), dtype=torch.float32, device=quaternion.device)
        return _QuatToT.apply(zero, quaternion)[:, :3, :3]

    @staticmethod
    def get_transformation_matrix_quaternion(translation, quaternion, device):
        """src/models/model_parts.py:37-44: T = [[R(q), t], [0, 1]]."""
        return _QuatToT.apply(translation, quaternion)

"""Import the UNMODIFIED reference (`/root/reference/src`) as a checker.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/gen_golden.py` (run in the build container, where
`/root/reference` is mounted) to produce the golden vectors under `tests/golden/`, and by the
`-m "not gpu"` tests that re-pin the oracle when the reference is present.  Nothing under
`delora_b200/` imports this file and nothing here runs on the GPU box.

Shims injected before import; no reference file is edited:
  * ``torch.symeig`` was removed from torch>=2.0  ->  ``torch.linalg.eigh`` (same LAPACK
    ``syevd`` family; ascending eigenvalues) — call site src/preprocessing/normal_computation.py:70.
  * ``kornia`` 0.3.0 (conda/DeLORA-py3.9.yml:53) is not installed -> stub with the published
    0.3.0 algorithm of ``quaternion_to_rotation_matrix`` ((x, y, z, w) order, L2-normalised,
    eps 1e-12) and ``angle_axis_to_rotation_matrix`` — call sites src/models/model_parts.py:31,35.
  * ``mlflow``, ``qqdm``, ``matplotlib``, ``pykitti`` stubs (import-time only).
"""
import os
import sys
import types

import torch

REFERENCE_SRC = "/root/reference/src"


def reference_available():
    return os.path.isdir(REFERENCE_SRC)


def kornia_quaternion_to_rotation_matrix(quaternion):
    """kornia==0.3.0 `kornia.geometry.conversions.quaternion_to_rotation_matrix` restated from
    its published source: normalise (F.normalize, eps=1e-12), unpack (x, y, z, w), build R from
    the doubled products.  The in-repo statement of the same convention is
    src/ros_utils/odometry_publisher.py:113-126 (`quat2mat`)."""
    q = torch.nn.functional.normalize(quaternion, p=2.0, dim=-1, eps=1e-12)
    x, y, z, w = torch.chunk(q, chunks=4, dim=-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.tensor(1.0, device=quaternion.device)
    matrix = torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                          txy + twz, one - (txx + tzz), tyz - twx,
                          txz - twy, tyz + twx, one - (txx + tyy)], dim=-1).view(-1, 3, 3)
    if quaternion.dim() == 1:
        matrix = matrix.squeeze(0)
    return matrix


def kornia_angle_axis_to_rotation_matrix(angle_axis):
    """kornia==0.3.0 `angle_axis_to_rotation_matrix` (Rodrigues with a first-order Taylor
    branch for theta^2 <= 1e-6)."""
    def _rodrigues(aa, theta2, eps=1e-6):
        theta = torch.sqrt(theta2)
        wxyz = aa / (theta + eps)
        wx, wy, wz = torch.chunk(wxyz, 3, dim=1)
        c, s = torch.cos(theta), torch.sin(theta)
        r00 = c + wx * wx * (1.0 - c)
        r10 = wz * s + wx * wy * (1.0 - c)
        r20 = -wy * s + wx * wz * (1.0 - c)
        r01 = wx * wy * (1.0 - c) - wz * s
        r11 = c + wy * wy * (1.0 - c)
        r21 = wx * s + wy * wz * (1.0 - c)
        r02 = wy * s + wx * wz * (1.0 - c)
        r12 = -wx * s + wy * wz * (1.0 - c)
        r22 = c + wz * wz * (1.0 - c)
        return torch.cat([r00, r01, r02, r10, r11, r12, r20, r21, r22], dim=1).view(-1, 3, 3)

    def _taylor(aa):
        rx, ry, rz = torch.chunk(aa, 3, dim=1)
        k1 = torch.ones_like(rx)
        return torch.cat([k1, -rz, ry, rz, k1, -rx, -ry, rx, k1], dim=1).view(-1, 3, 3)

    _aa = angle_axis.unsqueeze(1)
    theta2 = torch.matmul(_aa, _aa.transpose(1, 2)).squeeze(1)
    rot_normal = _rodrigues(angle_axis, theta2)
    rot_taylor = _taylor(angle_axis)
    mask = (theta2 > 1e-6).view(-1, 1, 1).to(theta2.device).type_as(theta2)
    return mask * rot_normal + (1.0 - mask) * rot_taylor


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_installed = False


def install():
    """Make `import utility.projection` etc. resolve to the reference's own modules."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference sources not mounted at " + REFERENCE_SRC)
    # torch>=2.0 keeps a `torch.symeig` that only raises; replace it unconditionally.
    def symeig(a, eigenvectors=False, upper=True):
        return torch.linalg.eigh(a, UPLO="U" if upper else "L")
    torch.symeig = symeig
    if "kornia" not in sys.modules:
        _stub("kornia", quaternion_to_rotation_matrix=kornia_quaternion_to_rotation_matrix,
              angle_axis_to_rotation_matrix=kornia_angle_axis_to_rotation_matrix)
    for name in ("mlflow", "mlflow.pytorch", "mlflow.tracking", "qqdm", "pykitti"):
        if name not in sys.modules:
            _stub(name)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            _stub("matplotlib")
            _stub("matplotlib.pyplot")
            _stub("mpl_toolkits")
            _stub("mpl_toolkits.mplot3d", Axes3D=object)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _installed = True


def reference_preprocesser():
    """The reference's offline stage (`preprocessing.preprocesser.Preprocesser`, src/preprocessing/preprocesser.py),
    imported unmodified.  Extra import-time stubs: `pykitti.utils.load_velo_scan` restated from pykitti 0.3.1
    (`np.fromfile(file, dtype=np.float32).reshape((-1, 4))`; call site src/data/kitti_scans.py:44) and empty ROS
    modules for `data.rosbag_scans`."""
    install()
    import numpy as np

    def load_velo_scan(file):
        return np.fromfile(file, dtype=np.float32).reshape((-1, 4))

    utils = types.ModuleType("pykitti.utils")
    utils.load_velo_scan = load_velo_scan
    utils.yield_velo_scans = lambda files: (load_velo_scan(f) for f in files)
    sys.modules["pykitti"].utils = utils
    sys.modules["pykitti.utils"] = utils
    for name in ("rosbag", "rospy", "ros_numpy", "sensor_msgs", "sensor_msgs.msg", "sensor_msgs.point_cloud2"):
        if name not in sys.modules:
            _stub(name)
    import preprocessing.preprocesser
    return preprocessing.preprocesser


def reference_modules():
    """Returns the reference's hot-path modules, imported unmodified."""
    install()
    import utility.projection
    import utility.linalg
    import preprocessing.normal_computation
    import losses.icp_losses
    import models.model
    import models.model_parts
    import models.resnet_modified
    return types.SimpleNamespace(projection=utility.projection, linalg=utility.linalg,
                                 normal_computation=preprocessing.normal_computation,
                                 icp_losses=losses.icp_losses, model=models.model,
                                 model_parts=models.model_parts,
                                 resnet_modified=models.resnet_modified)

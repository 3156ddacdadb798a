"""Synthetic KITTI-shaped LiDAR scan pairs (SURVEY.md §8(d) "Synthetic inputs").

Scene: the sensor sits 1.73 m above a ground plane inside a 40 m x 40 m axis-aligned
box room with infinitely tall walls.  A scan is a ring-structured set of rays
(`rings` elevations uniform in the vertical FOV x `w_raw` azimuths uniform in
(-179.9 deg, 179.9 deg)), every ray jittered by U(-0.3, 0.3) of a ray-grid cell,
ray-cast to the scene, perturbed by N(0, 1 cm) noise per coordinate; 2 % of the rays
are dropped.  Scan t+1 is rendered from the ego-motion (0.5 m forward, 1 deg yaw).

Everything is generated with a seeded CPU ``torch.Generator`` (seed = 1000 + pair
index), so the same arrays come out here, in the tests and on the GPU box.

This file is host-side input generation only; nothing here is a compute path.
"""
import math

import numpy as np
import torch

SENSOR_HEIGHT = 1.73
ROOM_HALF = 20.0
KITTI_VFOV_DEG = (-24.5, 2.0)
HFOV_DEG = (-179.9, 179.9)


def fov_config(h=64, w=2048, vfov_deg=KITTI_VFOV_DEG, hfov_deg=HFOV_DEG, dataset="kitti",
               device="cpu", neighborhood=(7, 11)):
    """A reference-style flat config dict (angles already in radians, as the reference's
    bin scripts hand them to the operators: bin/run_training.py:62-67)."""
    return {
        "device": device,
        "horizontal_field_of_view": [hfov_deg[0] * (np.pi / 180.0), hfov_deg[1] * (np.pi / 180.0)],
        "epsilon_range": 0.5,
        "min_num_points_in_neighborhood_to_determine_point_class": 10,
        "epsilon_plane": 0.01,
        "epsilon_line": 0.01,
        "datasets": [dataset],
        dataset: {
            "vertical_field_of_view": [vfov_deg[0] * (np.pi / 180.0), vfov_deg[1] * (np.pi / 180.0)],
            "vertical_cells": h,
            "horizontal_cells": w,
            "neighborhood_side_length": list(neighborhood),
        },
        # loss switches: config/hyperparameters.yaml:14-19
        "point_to_point_loss": False,
        "point_to_plane_loss": True,
        "plane_to_plane_loss": True,
        "po2po_alone": False,
        "normal_loss": "squared",
        "lambda_po2pl": 1.0,
    }


def _rot_z(yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def render_scan(gen, position_xy=(0.0, 0.0), yaw=0.0, rings=64, w_raw=2048,
                vfov_deg=KITTI_VFOV_DEG, drop_fraction=0.02, noise_sigma=0.01):
    """Ray-cast one scan; returns float32 [3, N] in the sensor frame (ring-major order)."""
    n = rings * w_raw
    el0, el1 = math.radians(vfov_deg[0]), math.radians(vfov_deg[1])
    az0, az1 = math.radians(HFOV_DEG[0]), math.radians(HFOV_DEG[1])
    ring = torch.arange(rings, dtype=torch.float64).repeat_interleave(w_raw)
    col = torch.arange(w_raw, dtype=torch.float64).repeat(rings)
    jit = (torch.rand(2, n, generator=gen, dtype=torch.float64) - 0.5) * 0.6
    el = el0 + (ring + jit[0]) / (rings - 1) * (el1 - el0)
    az = az0 + (col + jit[1]) / (w_raw - 1) * (az1 - az0)
    d_s = torch.stack((torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)))
    d_w = torch.from_numpy(_rot_z(yaw)) @ d_s
    px, py = position_xy
    inf = torch.full((n,), float("inf"), dtype=torch.float64)
    t_ground = torch.where(d_w[2] < 0, -SENSOR_HEIGHT / d_w[2], inf)
    t_x = torch.where(d_w[0] > 0, (ROOM_HALF - px) / d_w[0],
                      torch.where(d_w[0] < 0, (-ROOM_HALF - px) / d_w[0], inf))
    t_y = torch.where(d_w[1] > 0, (ROOM_HALF - py) / d_w[1],
                      torch.where(d_w[1] < 0, (-ROOM_HALF - py) / d_w[1], inf))
    t = torch.minimum(t_ground, torch.minimum(t_x, t_y))
    pts = d_s * t + noise_sigma * torch.randn(3, n, generator=gen, dtype=torch.float64)
    keep = torch.rand(n, generator=gen, dtype=torch.float64) >= drop_fraction
    return pts[:, keep].to(torch.float32).contiguous()


def transform_matrix(dx, dy, dz, yaw, pitch=0.0, roll=0.0):
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    t = np.eye(4, dtype=np.float64)
    t[:3, :3] = rz @ ry @ rx
    t[:3, 3] = (dx, dy, dz)
    return t


def make_pair(index, w_raw=2048, rings=64, vfov_deg=KITTI_VFOV_DEG):
    """Scan pair `index`: (scan_1 [3,N1], scan_2 [3,N2], T_gt [4,4], T_pred [4,4]) float32.

    T maps scan_2 coordinates into the frame of scan_1 (the reference transforms the
    source = scan_2 and matches it against the target = scan_1: deployer.py:294-307).
    """
    gen = torch.Generator(device="cpu")
    gen.manual_seed(1000 + int(index))
    yaw = math.radians(1.0)
    scan_1 = render_scan(gen, (0.0, 0.0), 0.0, rings, w_raw, vfov_deg)
    scan_2 = render_scan(gen, (0.5, 0.0), yaw, rings, w_raw, vfov_deg)
    t_gt = transform_matrix(0.5, 0.0, 0.0, yaw)
    t_pred = transform_matrix(0.5 + 0.05, 0.0, 0.0, yaw + math.radians(0.2))
    return scan_1, scan_2, torch.from_numpy(t_gt).float(), torch.from_numpy(t_pred).float()


def tie_stress_cloud(index=0, w_raw=512, rings=16, vfov_deg=(-15.0, 15.0), n_dup=200):
    """A cloud with exact duplicate points (equal fp32 range in one pixel) appended."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(5000 + int(index))
    scan = render_scan(gen, (0.0, 0.0), 0.0, rings, w_raw, vfov_deg)
    pick = torch.randperm(scan.shape[1], generator=gen)[:n_dup]
    return torch.cat((scan, scan[:, pick]), dim=1).contiguous()


def edge_stress_cloud(h=16, w=180, vfov_deg=(-15.0, 15.0), r=7.5):
    """Points placed on/near pixel-boundary azimuths and elevations (k + 0.5 pixels), the
    image seam (|azimuth| -> 180 deg), outside the vertical FOV, and points with zero
    coordinates (they survive the projection but are not 'valid' pixels for the normals)."""
    hf0, hf1 = math.radians(HFOV_DEG[0]), math.radians(HFOV_DEG[1])
    vf0, vf1 = math.radians(vfov_deg[0]), math.radians(vfov_deg[1])
    pts = []
    r0 = r

    def next_r():
        # every point gets its own range, so no equal-range ties blur the rounding-boundary cases
        return r0 + 0.013 * len(pts)
    for k in range(0, w - 1, 7):
        for dv in (0.0, 0.25, 0.5):
            for eps in (-1e-3, 0.0, 1e-3):
                r = next_r()
                az = hf0 + (k + 0.5 + eps) / (w - 1) * (hf1 - hf0)
                el = vf0 + ((k % (h - 1)) + dv) / (h - 1) * (vf1 - vf0)
                pts.append((r * math.cos(el) * math.cos(az), r * math.cos(el) * math.sin(az), r * math.sin(el)))
    for az_deg in (-180.0, -179.99, -179.9, 179.9, 179.95, 180.0):
        r = next_r()
        az = math.radians(az_deg)
        pts.append((r * math.cos(az), r * math.sin(az), 0.1))
    for el_deg in (vfov_deg[0] - 1.5, vfov_deg[0] - 0.9, vfov_deg[1] + 0.9, vfov_deg[1] + 1.5, 89.0, -89.0):
        r = next_r()
        el = math.radians(el_deg)
        pts.append((r * math.cos(el), 0.3 * math.cos(el), r * math.sin(el)))
    pts += [(5.0, 0.0, 0.0), (0.0, 5.0, 0.0), (3.0, 3.0, 0.0), (0.0, 0.0, 4.0), (-4.0, 0.0, 0.5), (0.0, 0.0, 0.0)]
    return torch.tensor(pts, dtype=torch.float32).t().contiguous()


def kitti_bin_scan(index, w_raw=192, rings=16, vfov_deg=(-15.0, 15.0)):
    """A synthetic scan in KITTI's velodyne .bin layout: [N, 4] float32 rows (x, y, z, reflectance)."""
    scan_1, _, _, _ = make_pair(index, w_raw=w_raw, rings=rings, vfov_deg=vfov_deg)
    g = torch.Generator().manual_seed(5000 + index)
    refl = torch.rand(scan_1.shape[1], generator=g)
    return torch.cat((scan_1, refl[None]), dim=0).t().contiguous().numpy().astype(np.float32)


def preprocessing_config(data_path, preprocessed_path, h=16, w_pre=200, vfov_deg=(-15.0, 15.0), identifiers=(0,),
                         device="cpu"):
    """Config of the offline stage as bin/preprocess_data.py builds it (angles in radians)."""
    cfg = fov_config(h=h, w=w_pre, vfov_deg=vfov_deg, device=device)
    cfg["kitti"].update({"horizontal_cells_preprocessing": w_pre, "data_identifiers": list(identifiers),
                         "data_path": str(data_path), "preprocessed_path": str(preprocessed_path),
                         "dataset_type": "kitti"})
    cfg["visualize_single_img_preprocessing"] = False
    return cfg

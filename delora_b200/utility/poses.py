"""Drop-in `utility.poses` (reference: src/utility/poses.py): chaining of the per-frame relative transforms
into KITTI world-frame poses, and the 12-column KITTI pose file.

`compute_poses(computed_transformations)`: T_0k(lidar) = T_0,k-1 . T_k-1,k, rotation re-projected onto SO(3) through a
unit quaternion after every step (:44-50), expressed in the KITTI camera/world axes with the fixed
lidar->world permutation (:20-29, :51-53).  Host-side float64 numpy + scipy, as in the reference: this is a
4x4 product per frame, not GPU work.
"""
import csv

import numpy as np
import scipy.spatial.transform

LIDAR_TO_WORLD = np.asarray([[[0, -1, 0, 0],
                              [0, 0, -1, 0],
                              [1, 0, 0, 0],
                              [0, 0, 0, 1]]], dtype=np.float64)


def check_validity_so3(r):
    det_valid = np.isclose(np.linalg.det(r), [1.0], atol=1e-6)
    inv_valid = np.allclose(r.transpose().dot(r), np.eye(3), atol=1e-6)
    return det_valid and inv_valid


def compute_poses(computed_transformations):
    world_to_lidar = np.transpose(LIDAR_TO_WORLD, (0, 2, 1))
    t_lidar = np.eye(4, dtype=np.float64)[None].copy()
    poses = [np.eye(4, dtype=np.float64)[None]]
    for t_rel in computed_transformations:
        t_lidar = np.matmul(t_lidar, np.asarray(t_rel))
        quat = scipy.spatial.transform.Rotation.from_matrix(t_lidar[0, :3, :3]).as_quat()
        quat /= np.linalg.norm(quat)
        t_lidar[0, :3, :3] = scipy.spatial.transform.Rotation.from_quat(quat).as_matrix()
        t_world = np.matmul(np.matmul(LIDAR_TO_WORLD, t_lidar), world_to_lidar)
        if not check_validity_so3(r=t_world[0, :3, :3]):
            raise Exception("Pose is not valid!")
        poses.append(t_world)
    return np.concatenate(poses, axis=0)


def write_poses_to_text_file(file_name, poses):
    with open(file_name, "w", newline="") as txt_file:
        writer = csv.writer(txt_file, delimiter=" ")
        for pose in poses:
            writer.writerow(pose.reshape(16)[:12])

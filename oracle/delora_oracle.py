"""CPU restatement (oracle) of DeLORA's hot path.  TEST INFRASTRUCTURE — NOT A PRODUCT PATH.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this file.  `delora_b200/` never does.

The reference is Python/PyTorch, so this restatement is torch-CPU + numpy + scipy: the same
floating-point op sequence as the reference for everything that fixes a result bit (range,
(u, v), rounding, covariance, eigen-solve, loss), and plain vectorised numpy for the integer
work (first-come occupancy, masks, gathers).  Each function cites the reference lines it follows
(paths relative to /root/reference).

Pinning: `oracle/gen_golden.py` runs the UNMODIFIED reference (via `oracle/ref_harness.py`) on
committed synthetic inputs and stores its outputs / digests in `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against them (bit-exact for the projection,
normals, kept-pair sets and NN indices; losses and gradients to 1e-6 rel).  The reference itself
has no tests or golden vectors, and three third-party libraries sit on the path
(scipy cKDTree, LAPACK syevd through torch, kornia 0.3.0): those are restated/called as the
reference calls them and are "parity unpinned" by the reference's own repository.

Deliberate definitions where the reference is non-deterministic:
  * equal-range ties inside one pixel: the reference's `torch.argsort` is unstable
    (src/utility/projection.py:63); the oracle uses a STABLE sort = lowest point index wins.
  * exact-distance NN ties: cKDTree order is arbitrary; the oracle's brute-force path picks the
    lowest target index.
"""
import numpy as np
import torch

try:  # scipy is the reference's own dependency for the NN step (src/losses/icp_losses.py:6)
    import scipy.spatial
except ImportError:  # pragma: no cover
    scipy = None


# --------------------------------------------------------------------------------------
# a1-a3  projection            src/utility/projection.py:21-106
# --------------------------------------------------------------------------------------
def compute_2d_coordinates(point_cloud, width_pixel, height_pixel, hfov, vfov):
    """src/utility/projection.py:21-31 — same torch op order (sub, div, mul; atan2; norm)."""
    u = ((torch.atan2(point_cloud[:, 1, :], point_cloud[:, 0, :]) - hfov[0]) / (hfov[1] - hfov[0])
         * (width_pixel - 1))
    v = ((torch.atan2(point_cloud[:, 2, :], torch.norm(point_cloud[:, :2, :], dim=1)) - vfov[0])
         / (vfov[1] - vfov[0]) * (height_pixel - 1))
    return u, v


def project_to_img(point_cloud, height_pixel, width_pixel, hfov, vfov, device="cpu", uv_override=None):
    """src/utility/projection.py:48-106.

    point_cloud [1,C,N] -> (image [1,C+1,H,W], u [1,N], v [1,N], point_indices [K] int64
    ascending range, image_to_pointcloud_indices [1,K,2] int64 (v,u)).
    `device` only selects where the torch float ops run (the tests also run this on "cuda" to
    compare against what the reference computes with its default ``device: "cuda"``).
    `uv_override=(u, v)` ([N] each, ORIGINAL point order) replaces the computed float coordinates;
    the tests use it to check the integer stage (min-range selection, scatter, indices) of the
    CUDA kernel bit-exactly on the kernel's own (u, v).
    """
    point_cloud = point_cloud.to(device)
    b, c, n = point_cloud.shape
    pcr = torch.zeros((b, c + 1, n), device=device)                       # :56-59
    pcr[:, :c, :] = point_cloud
    pcr[:, -1, :] = torch.norm(pcr[:, :3, :], dim=1)                      # :59-60
    sort_indices = torch.argsort(pcr[:, c, :], dim=1, stable=True)        # :63 (stable: see header)
    pcr = pcr[:, :, sort_indices[0]]                                      # :67
    u, v = compute_2d_coordinates(pcr, width_pixel, height_pixel, hfov, vfov)   # :69
    if uv_override is not None:
        u = uv_override[0].to(device)[sort_indices[0]][None]
        v = uv_override[1].to(device)[sort_indices[0]][None]
    ru, rv = torch.round(u), torch.round(v)
    inside = (ru <= width_pixel - 1) & (ru >= 0) & (rv <= height_pixel - 1) & (rv >= 0)   # :74-75
    u_f = ru[inside].long().cpu().numpy()                                 # :76-77, :87-88
    v_f = rv[inside].long().cpu().numpy()
    pcr_in = pcr[:, :, inside[0]]                                         # :78
    # remove_duplicate_indices (:34-43): serial first-come occupancy over the range-sorted list
    # == first occurrence of each flat pixel id.
    flat = v_f * width_pixel + u_f
    _, first = np.unique(flat, return_index=True)
    unique_bool = np.zeros(len(flat), dtype=bool)
    unique_bool[first] = True
    unique_t = torch.from_numpy(unique_bool).to(device)
    i2p = np.stack((v_f[unique_bool], u_f[unique_bool]), axis=1)[None].astype(np.int64)
    pcr_k = pcr_in[:, :, unique_t]                                        # :95
    image = torch.zeros((b, c + 1, height_pixel, width_pixel), device=device)   # :98-100
    vk = torch.from_numpy(v_f[unique_bool]).to(device)
    uk = torch.from_numpy(u_f[unique_bool]).to(device)
    image[:, :, vk, uk] = pcr_k                                           # :102-103
    point_indices = sort_indices[inside][unique_t]                        # :105
    return image, u, v, point_indices, torch.from_numpy(i2p).to(device)


# --------------------------------------------------------------------------------------
# a4-a5  normals               src/preprocessing/normal_computation.py:30-122, src/utility/linalg.py:33-56
# --------------------------------------------------------------------------------------
def cov_zero_aware(point_neighbors):
    """src/utility/linalg.py:33-56, 3-D branch.  point_neighbors [P,3,K]."""
    not_zero = ((point_neighbors[:, 0, :] != 0) | (point_neighbors[:, 1, :] != 0)
                | (point_neighbors[:, 2, :] != 0))                        # :34-37
    number_neighbours = torch.sum(not_zero, dim=1)                        # :38
    factor = torch.ones(1) / (number_neighbours - 1)                      # :39
    mean = (torch.mean(point_neighbors, dim=2, keepdim=True) * point_neighbors.shape[2]
            / number_neighbours.view(-1, 1, 1))                           # :41-42
    difference = point_neighbors - mean                                   # :43
    difference.permute(0, 2, 1)[~not_zero] = 0.0                          # :44-45
    squared = difference.matmul(difference.permute(0, 2, 1))              # :46, :54
    return factor.view(-1, 1, 1) * squared, number_neighbours             # :56


def compute_normal_vectors(image, neighborhood=(7, 11), epsilon_range=0.5, min_neighbors=10,
                           return_aux=False):
    """src/preprocessing/normal_computation.py:89-122 + :30-41 + :53-87.

    image [1,>=3,H,W] -> (normals [P,3], has_normal [P] bool, points [P,3]); P = pixels with
    x!=0 & y!=0 & z!=0 in row-major order.
    """
    img = image[0, :3].cpu()
    _, h, w = img.shape
    flat = img.reshape(3, h * w).transpose(0, 1)
    valid = (flat[:, 0] != 0) & (flat[:, 1] != 0) & (flat[:, 2] != 0)    # :35
    pix = torch.nonzero(valid)[:, 0]
    v_c = pix // w
    u_c = pix % w
    a = int(neighborhood[0] / 2)                                          # :97
    b = int(neighborhood[1] / 2)                                          # :98
    dv = torch.arange(-a, a + 1).repeat_interleave(2 * b + 1)             # loop order :99-100
    du = torch.arange(-b, b + 1).repeat(2 * a + 1)
    vv = (v_c[None, :] + dv[:, None]).clamp_(0, h - 1)                    # edge clamp :104-111
    uu = (u_c[None, :] + du[:, None]).clamp_(0, w - 1)
    point_neighbors = img[:, vv, uu].permute(1, 0, 2).contiguous()        # [K,3,P]  :112-117
    point_locations = img[:, v_c, u_c].view(1, 3, -1)                     # :119
    # covariance_eigen_decomposition :53-87
    deviates = torch.abs(torch.norm(point_neighbors, dim=1)
                         - torch.norm(point_locations, dim=1)) > epsilon_range      # :56-57
    point_neighbors.permute(0, 2, 1)[deviates] = 0.0                      # :59
    cov, number_neighbours = cov_zero_aware(point_neighbors.permute(2, 1, 0))       # :61-63
    enough = number_neighbours >= min_neighbors                           # :67-68
    cov_kept = cov[enough]
    eigenvalues, eigenvectors = torch.linalg.eigh(cov_kept, UPLO="U")     # torch.symeig :70
    normals_kept = eigenvectors[:, :, 0].clone()                          # :76
    locs = point_locations[0].permute(1, 0)                               # :78
    dots = normals_kept.view(-1, 1, 3).matmul(locs[enough].view(-1, 3, 1)).reshape(-1)   # :79-80
    normals_kept[dots > 0] *= -1                                          # :81
    normals = torch.zeros_like(locs)                                      # :84
    normals[enough] = normals_kept                                        # :85
    if return_aux:
        aux = {"cov": cov, "number_neighbours": number_neighbours, "eigenvalues": eigenvalues,
               "enough": enough, "pixel_index": pix}
        return normals, enough, locs.contiguous(), aux
    return normals, enough, locs.contiguous()


# --------------------------------------------------------------------------------------
# a7-a8  quaternion -> T, SE(3) transforms     src/models/model_parts.py:29-44, src/deploy/deployer.py:181-189
# --------------------------------------------------------------------------------------
def quaternion_to_rotation_matrix(quaternion):
    """kornia==0.3.0 (conda/DeLORA-py3.9.yml:53) `quaternion_to_rotation_matrix`, (x,y,z,w),
    restated from its published source; call site src/models/model_parts.py:31.  The in-repo
    statement of the convention is src/ros_utils/odometry_publisher.py:113-126."""
    q = torch.nn.functional.normalize(quaternion, p=2.0, dim=-1, eps=1e-12)
    x, y, z, w = torch.chunk(q, chunks=4, dim=-1)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.tensor(1.0)
    return torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                        txy + twz, one - (txx + tzz), tyz - twx,
                        txz - twy, tyz + twx, one - (txx + tyy)], dim=-1).view(-1, 3, 3)


def transformation_matrix_quaternion(translation, quaternion):
    """src/models/model_parts.py:37-44."""
    rot = quaternion_to_rotation_matrix(quaternion)
    t = torch.zeros((rot.shape[0], 4, 4))
    t[:, :3, :3] = rot
    t[:, 3, 3] = 1
    t[:, :3, 3] = translation
    return t


def rotate_point_cloud(transformation_matrix, point_cloud):
    """src/deploy/deployer.py:181-182."""
    return transformation_matrix[:, :3, :3].matmul(point_cloud[:, :3, :])


def transform_point_cloud(transformation_matrix, point_cloud):
    """src/deploy/deployer.py:184-189."""
    out = rotate_point_cloud(transformation_matrix, point_cloud)
    return out + transformation_matrix[:, :3, 3].view(-1, 3, 1)


# --------------------------------------------------------------------------------------
# a9-a12 ICP losses            src/losses/icp_losses.py:24-240
# --------------------------------------------------------------------------------------
def nearest_neighbors(target_xyz, source_xyz, method="kdtree"):
    """Exact 3-D Euclidean NN in float64 (src/losses/icp_losses.py:24-26, :34).

    "kdtree": scipy.spatial.cKDTree(target).query(source) — the reference's own call.
    "brute":  float64 brute force, lowest target index on exact ties (small cases only).
    target_xyz [Nt,3], source_xyz [Ns,3] float32 numpy -> int64 [Ns]."""
    if len(source_xyz) == 0:
        return np.zeros((0,), dtype=np.int64)
    if method == "kdtree":
        tree = scipy.spatial.cKDTree(target_xyz)
        return tree.query(source_xyz)[1].astype(np.int64)
    t = target_xyz.astype(np.float64)
    out = np.empty(len(source_xyz), dtype=np.int64)
    chunk = max(1, int(2e7 // max(1, len(t))))
    for i0 in range(0, len(source_xyz), chunk):
        s = source_xyz[i0:i0 + chunk].astype(np.float64)
        d2 = ((s[:, None, :] - t[None, :, :]) ** 2).sum(-1)
        out[i0:i0 + chunk] = np.argmin(d2, axis=1)
    return out


def icp_losses(source_points_t, source_normals_t, target_points, target_normals,
               point_to_point_loss=False, point_to_plane_loss=True, plane_to_plane_loss=True,
               normal_loss="squared", nn_method="kdtree", return_aux=False):
    """src/losses/icp_losses.py:28-158 (the `po2po_alone: False` branch) with the sub-losses
    :168-179 (po2po), :196-206 (po2pl), :224-240 (pl2pl).  All inputs [1,3,N] float32 torch
    tensors; the two source tensors may require grad.  Returns the reference's `losses` dict
    (+ aux: NN indices for every source point, kept-pair mask)."""
    src_has_n = ((source_normals_t[:, 0, :] != 0) | (source_normals_t[:, 1, :] != 0)
                 | (source_normals_t[:, 2, :] != 0))[0]                   # :48-50
    tgt_has_n = ((target_normals[:, 0, :] != 0) | (target_normals[:, 1, :] != 0)
                 | (target_normals[:, 2, :] != 0))[0]                     # :51-52
    tgt_np = target_points[0].permute(1, 0).detach().cpu().numpy()
    src_np = source_points_t[0].permute(1, 0).detach().cpu().numpy()
    nn_all = torch.from_numpy(nearest_neighbors(tgt_np, src_np, nn_method))   # :70-80 (one query; same result per point)
    # 2) source has a normal, and so does its target  (:102-121)
    idx_n = nn_all[src_has_n]
    keep = tgt_has_n[idx_n]
    s_pts = source_points_t[:, :, src_has_n][:, :, keep]
    s_nrm = source_normals_t[:, :, src_has_n][:, :, keep]
    t_pts = target_points[:, :, idx_n][:, :, keep]
    t_nrm = target_normals[:, :, idx_n][:, :, keep]
    zero = torch.zeros(1)
    loss_po2po, loss_po2pl, loss_pl2pl = zero, zero, zero
    mse = torch.nn.MSELoss()
    if point_to_point_loss:                                               # :83-99, :168-179
        idx_nn = nn_all[~src_has_n]
        keep_nn = ~tgt_has_n[idx_nn]
        loss_po2po = mse(source_points_t[:, :, ~src_has_n][:, :, keep_nn],
                         target_points[:, :, idx_nn][:, :, keep_nn])
    if point_to_plane_loss:                                               # :196-206
        dist = (s_pts - t_pts).permute(2, 0, 1).matmul(t_nrm.permute(2, 1, 0))
        loss_po2pl = mse(dist, torch.zeros(dist.shape))
    if plane_to_plane_loss:                                               # :224-240
        sn = s_nrm.permute(2, 0, 1)
        if normal_loss == "linear":
            dots = torch.matmul(sn, t_nrm.permute(2, 1, 0))
            loss_pl2pl = mse(1 - dots, torch.zeros(dots.shape))
        elif normal_loss == "squared":
            d = torch.norm(sn - t_nrm.permute(2, 0, 1), dim=2, keepdim=True)
            loss_pl2pl = mse(d, torch.zeros(d.shape))
        else:
            raise Exception("The normal loss which is defined here is not admissible.")
    losses = {"loss_po2po": loss_po2po, "loss_po2pl": loss_po2pl, "loss_pl2pl": loss_pl2pl}
    if return_aux:
        kept_mask = torch.zeros(src_has_n.shape[0], dtype=torch.bool)
        kept_mask[torch.nonzero(src_has_n)[:, 0][keep]] = True
        aux = {"nn_index": nn_all, "kept_mask": kept_mask, "num_pairs": int(keep.sum()),
               "source_points_where_normals": s_pts, "source_normals_where_normals": s_nrm}
        return losses, aux
    return losses


# --------------------------------------------------------------------------------------
# the per-pair hot path (BASELINE.json configs[0] / configs[1]):
#   2x projection -> 2x normals -> SE(3) transform of the source -> ICP losses (+ backward to T)
# --------------------------------------------------------------------------------------
def pair_forward_backward(scan_1, scan_2, transform, cfg, dataset="kitti", lambda_po2pl=1.0,
                          backward=True, nn_method="kdtree", timings=None):
    """scan_k [3,N] float32, transform [4,4].  Follows Deployer.step for one sample
    (src/deploy/deployer.py:252-261, :294-312) with the normals computed in-line from the
    projected image as Preprocesser.apply_preprocessing_step does
    (src/preprocessing/preprocesser.py:52,60-61).  Returns dict(loss_po2pl, loss_pl2pl, loss,
    grad_T [3,4], num_pairs)."""
    import time
    ds = cfg[dataset]
    h, w = ds["vertical_cells"], ds["horizontal_cells"]
    hf, vf = cfg["horizontal_field_of_view"], ds["vertical_field_of_view"]
    t0 = time.perf_counter()
    img1 = project_to_img(scan_1[None], h, w, hf, vf)[0]
    img2 = project_to_img(scan_2[None], h, w, hf, vf)[0]
    t1 = time.perf_counter()
    nb = ds["neighborhood_side_length"]
    n1, _, p1 = compute_normal_vectors(img1, nb, cfg["epsilon_range"],
                                       cfg["min_num_points_in_neighborhood_to_determine_point_class"])
    n2, _, p2 = compute_normal_vectors(img2, nb, cfg["epsilon_range"],
                                       cfg["min_num_points_in_neighborhood_to_determine_point_class"])
    t2 = time.perf_counter()
    tm = transform.clone().view(1, 4, 4).requires_grad_(backward)
    src = transform_point_cloud(tm, p2.t()[None])
    src_n = rotate_point_cloud(tm, n2.t()[None])
    losses, aux = icp_losses(src, src_n, p1.t()[None].contiguous(), n1.t()[None].contiguous(),
                             nn_method=nn_method, return_aux=True)
    loss = lambda_po2pl * losses["loss_po2pl"] + losses["loss_pl2pl"]    # deployer.py:309-312 (B=1)
    grad = None
    if backward:
        loss.sum().backward()
        grad = tm.grad[0, :3, :].clone()
    t3 = time.perf_counter()
    if timings is not None:
        timings.append({"projection_s": t1 - t0, "normals_s": t2 - t1, "loss_s": t3 - t2})
    return {"loss_po2pl": float(losses["loss_po2pl"].detach()), "loss_pl2pl": float(losses["loss_pl2pl"].detach()),
            "loss": float(loss.sum().detach()), "grad_T": grad, "num_pairs": aux["num_pairs"],
            "points_1": p1, "normals_1": n1, "points_2": p2, "normals_2": n2,
            "image_1": img1, "image_2": img2, "nn_index": aux["nn_index"], "kept_mask": aux["kept_mask"]}

"""CPU restatement (numpy / torch-CPU) of the four reference stages.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites
the reference lines it follows (paths relative to the MUVO checkout).  Two
flavours exist where the reference is slow:

* ``*_loop``  : follows the reference's control flow (incl. its per-voxel Python
                loop); this is what ``bench.py`` times as ``cpu_baseline`` /
                ``--impl reference`` (kind = "port").
* ``*_fast``  : same results, vectorised; used by tests at larger sizes.

Tie contract: the reference uses ``np.argsort`` with its default
(unstable) kind, so results on *exact* key ties are implementation-defined.
The oracle uses ``kind='stable'`` -- "lowest original index wins" -- which is
what the CUDA path implements.
"""
from __future__ import annotations

import math

import numpy as np
import torch

ROADLINE_ID = 6  # data/data_preprocessing.py:42-68,198  (LABEL_CLASS == 'roadlines')

__all__ = [
    "ROADLINE_ID", "voxel_grid_params", "voxel_filter_loop", "voxel_filter_fast", "densify_voxels",
    "label_remap_table", "range_projection", "range_projection_indices", "pack_range_view",
    "bev_intrinsics", "gen_dx_bx", "frustum_grid", "frustum_geometry", "bev_cell_ids",
    "cumsum_trick", "quick_cumsum_backward", "voxel_pooling_cumsum", "voxel_pooling_exact",
    "frustum_pooling_forward", "decode_depth_image", "depth2pcd", "merge_pcd_arrays", "lidar_prep", "label_pyramids", "ssc_counts", "ssc_counts_loop", "ssc_add_batch_counts",
    "ssc_stats_from_counts", "sem_scal_loss", "geo_scal_loss", "scal_sums", "scatter_mean", "scatter_max",
]


# --------------------------------------------------------------------------- (a)
def voxel_grid_params(voxel_resolution, voxel_size, offset):
    """Final additive offset and exclusive upper bound, both float64.

    data/data_preprocessing.py:173-178: ``offset += res * size / 2`` then
    ``0 <= pcd + offset < size * res``.  The caller's offset is never mutated.
    """
    size = np.asarray(voxel_size)
    res = np.asarray(voxel_resolution)
    off = np.array(offset, dtype=np.float64, copy=True)
    off = off + res * size / 2
    upper = (size * res).astype(np.float64)
    return off, upper


def _voxel_prepare(pcd, sem, voxel_resolution, voxel_size, offset):
    """Shared front half: bounds filter, divmod, linear id, stable sort.

    data/data_preprocessing.py:176-189.
    """
    pcd = np.asarray(pcd)
    sem = np.asarray(sem)
    res = np.asarray(voxel_resolution)
    off, upper = voxel_grid_params(voxel_resolution, voxel_size, offset)
    shifted = pcd + off                                             # :177 (float64)
    inside = ((0 <= shifted) & (shifted < upper)).all(axis=1)      # :178
    shifted, sem_in = shifted[inside], sem[inside]
    dx, dy, _ = np.asarray(voxel_size)
    cell, rem = np.divmod(shifted, res)                             # :183
    lin = cell[:, 0] + cell[:, 1] * dx + cell[:, 2] * dx * dy       # :184
    order = np.argsort(lin, kind="stable")                          # :187 (stable: see module doc)
    return lin[order], cell[order], rem[order], sem_in[order]


def voxel_filter_loop(pcd, sem, voxel_resolution, voxel_size, offset, roadline_id=ROADLINE_ID):
    """Per-voxel loop port of ``voxel_filter`` (data/data_preprocessing.py:172-228).

    Returns ``(voxels uint16 (n,3), semantics uint8 (n,))`` ordered by linear id
    ``x + y*Dx + z*Dx*Dy``.
    """
    lin, cell, rem, sem_s = _voxel_prepare(pcd, sem, voxel_resolution, voxel_size, offset)
    ids, first = np.unique(lin, return_index=True)                  # :192
    n_occ, n_all = ids.shape[0], lin.shape[0]
    voxels = np.zeros((n_occ, 3), dtype=np.uint16)
    labels = np.zeros((n_occ,), dtype=np.uint8)
    sem_flat = sem_s.reshape(n_all, -1)[:, 0] if n_all else sem_s.reshape(0)
    for i in range(n_occ):                                          # :202
        lo = first[i]
        hi = first[i + 1] if i < n_occ - 1 else n_all               # :210
        dist = np.sum(rem[lo:hi] ** 2, axis=1)                      # :212  ((x^2+y^2)+z^2, float64)
        seg = sem_flat[lo:hi]
        if np.isin(seg, roadline_id).any():                         # :217
            labels[i] = roadline_id
        else:
            labels[i] = seg[np.argmin(dist)]
        voxels[i] = cell[lo]                                        # :219
    return voxels, labels


def voxel_filter_fast(pcd, sem, voxel_resolution, voxel_size, offset, roadline_id=ROADLINE_ID):
    """Vectorised equivalent of :func:`voxel_filter_loop` (same outputs)."""
    lin, cell, rem, sem_s = _voxel_prepare(pcd, sem, voxel_resolution, voxel_size, offset)
    n_all = lin.shape[0]
    if n_all == 0:
        return np.zeros((0, 3), np.uint16), np.zeros((0,), np.uint8)
    sem_flat = sem_s.reshape(n_all, -1)[:, 0]
    sq = rem ** 2
    dist = (sq[:, 0] + sq[:, 1]) + sq[:, 2]                         # same order as np.sum(axis=1)
    # within each voxel: smallest dist, first (lowest original index) among equals
    order = np.lexsort((np.arange(n_all), dist, lin))
    lin_o = lin[order]
    starts = np.flatnonzero(np.r_[True, lin_o[1:] != lin_o[:-1]])
    winner = order[starts]
    labels = sem_flat[winner].astype(np.uint8)
    has_road = np.maximum.reduceat((sem_flat[order] == roadline_id).astype(np.uint8), starts)
    labels[has_road > 0] = roadline_id
    voxels = cell[winner].astype(np.uint16)
    return voxels, labels


def label_remap_table(label_map: dict) -> np.ndarray:
    """muvo/data/dataset.py:281-283: 256-entry-compatible remap built from LABEL_MAP."""
    tab = np.full((max(label_map.keys()) + 1), max(label_map.values()), dtype=np.uint8)
    tab[list(label_map.keys())] = list(label_map.values())
    return tab


def densify_voxels(voxel_data, voxel_size, remap=None):
    """``(n,4) uint16 [x,y,z,sem]`` -> dense uint8 grid.  muvo/data/dataset.py:317-327."""
    pts = voxel_data[:, :-1]
    lab = voxel_data[:, -1].copy()
    lab[lab == 255] = 0                                             # :323
    if remap is not None:
        lab = remap[lab]                                            # :324
    grid = np.zeros(tuple(voxel_size), dtype=np.uint8)
    grid[pts[:, 0], pts[:, 1], pts[:, 2]] = lab                     # :326
    return grid


# --------------------------------------------------------------------------- N1: camera + LiDAR cloud in front of (a)
EGO_VEHICLE_DIMENSION = [4.902, 2.128, 1.511]          # data/data_preprocessing.py:5


def lidar_prep(points_xyz, obj_tag, lidar_position, remap=None, ego_dimension=EGO_VEHICLE_DIMENSION):
    """The LiDAR-side prep in front of stage (b), muvo/data/dataset.py:278-290, on copies of the inputs:
    ``convert_coor_lidar`` (data/data_preprocessing.py:119-122: in-place ``+=`` on the float32 array, ``y *= -1``),
    ``remap[ObjTag]`` (:281-283) and the ego-box drop (:286-290).  Returns ``(points float32 (m,3), semantics uint8 (m,))``."""
    pcd = np.array(points_xyz, dtype=np.float32, copy=True)
    pcd += np.asarray(lidar_position)                               # data_preprocessing.py:120
    pcd[:, 1] *= -1                                                 # :121
    sem = np.asarray(obj_tag).reshape(-1)
    if remap is not None:
        sem = np.asarray(remap)[sem]                                # dataset.py:283
    if ego_dimension is not None:
        x, y, z = ego_dimension
        ego_box = np.array([[-x / 2, -y / 2, 0], [x / 2, y / 2, z]])    # :287
        ego_idx = ((ego_box[0] < pcd) & (pcd < ego_box[1])).all(axis=1)  # :288
        sem, pcd = sem[~ego_idx], pcd[~ego_idx]                     # :289-290
    return pcd, sem.astype(np.uint8)


def decode_depth_image(img):
    """``read_img`` without the file access (data/data_preprocessing.py:72-77): ``img`` = ``cv2.imread(file, -1)``,
    uint8 ``(H, W, 4)`` [B, G, R, semantic] -> (depth float64 (H,W) in metres, semantic uint8 (H,W))."""
    depth_color = img[..., :-1].astype(float)
    semantic = img[..., -1]
    depth = 1000 * ((256 ** 2 * depth_color[..., 2] + 256 * depth_color[..., 1] + depth_color[..., 0]) / (256 ** 3 - 1))
    return depth, semantic


def depth2pcd(depth, semantic, fov, range=100):
    """data/data_preprocessing.py:87-106, float64 throughout; pixels in row-major order."""
    h, w = depth.shape
    f = w / (2.0 * np.tan(fov * np.pi / 360.0))
    cx, cy = w / 2.0, h / 2.0
    d = depth.reshape(-1)
    valid = d < 1000                                                             # :94
    yy, xx = np.divmod(np.arange(h * w), w)
    d, xx, yy = d[valid], xx[valid], yy[valid]
    x, y = (xx - cx) * d / f, (yy - cy) * d / f                                  # :101
    pts = np.stack([x, y, d], axis=1)
    sem = semantic.reshape(-1, 1)[valid]
    keep = np.sqrt((x * x + y * y) + d * d) < range                              # :104  np.linalg.norm(axis=1)
    return pts[keep], sem[keep]


def merge_pcd_arrays(img, lidar_xyz, lidar_sem, camera_pos, lidar_pos, fov=110, mask_ego=True):
    """``merge_pcd`` (data/data_preprocessing.py:125-139) on arrays instead of files: camera cloud (float64) followed
    by the LiDAR cloud (float32 values), ego-box points removed, order preserved.  Returns ``(pcd float64 (n,3),
    semantic uint8 (n,1))``.  The 4x4 float32 ``mat @ pcd.T`` of convert_coor_img (:109-119) has one non-zero
    coefficient of +-1 per row besides the translation, so each coordinate is a single rounded float64 sum."""
    depth, semantic = decode_depth_image(img)
    ip, isem = depth2pcd(depth, semantic, fov)
    forward, right, up = (np.float64(np.float32(v)) for v in camera_pos)           # mat is np.float32 (:111)
    img_pcd = np.stack([ip[:, 2] + forward, -ip[:, 0] + (-right), -ip[:, 1] + up], axis=1)
    lp = np.array(lidar_xyz, dtype=np.float32, copy=True)
    lp += np.asarray(lidar_pos)                                                  # :122 float32 in-place add
    lp[:, 1] *= -1                                                               # :123
    pcd = np.concatenate([img_pcd, lp], axis=0)
    sem = np.concatenate([isem, np.asarray(lidar_sem).reshape(-1, 1)], axis=0)
    if mask_ego:
        x, y, z = EGO_VEHICLE_DIMENSION
        box = np.array([[-x / 2, -y / 2, 0], [x / 2, y / 2, z]])
        ego = ((box[0] < pcd) & (pcd < box[1])).all(axis=1)                      # :136
        pcd, sem = pcd[~ego], sem[~ego]
    return pcd, sem


def _nearest_idx(n_in, n_out):
    """PyTorch `nearest` source indices (torchvision resize NEAREST / F.interpolate): min(floor(dst * float32(in/out)), in-1)."""
    scale = np.float32(n_in) / np.float32(n_out)
    return np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64), n_in - 1)


def label_pyramids(range_xyzd=None, range_sem=None, voxel=None, scale=50.0):
    """LIDAR_RE / LIDAR_SEG / VOXEL_SEG blocks of ``PreProcess.forward`` (muvo/models/preprocess.py:151-186) on numpy arrays
    with a leading frame axis: level 1 (= input, range view divided by ``scale`` in float32), levels 2 and 4 by repeated
    nearest-neighbour halving."""
    out = {}

    def halve(a, axes):
        for ax in axes:
            a = np.take(a, _nearest_idx(a.shape[ax], a.shape[ax] // 2), axis=ax)
        return a
    if range_xyzd is not None:
        l1 = (np.asarray(range_xyzd, dtype=np.float32) / np.float32(scale)).astype(np.float32)
        l2 = halve(l1, (2, 3)); l4 = halve(l2, (2, 3))
        out.update(range_view_label_1=l1, range_view_label_2=l2, range_view_label_4=l4)
    if range_sem is not None:
        s2 = halve(np.asarray(range_sem), (1, 2)); s4 = halve(s2, (1, 2))
        out.update(range_view_seg_label_1=np.asarray(range_sem), range_view_seg_label_2=s2, range_view_seg_label_4=s4)
    if voxel is not None:
        v2 = halve(np.asarray(voxel), (1, 2, 3)); v4 = halve(v2, (1, 2, 3))
        out.update(voxel_label_1=np.asarray(voxel), voxel_label_2=v2, voxel_label_4=v4)
    return out


# --------------------------------------------------------------------------- (b)
def range_projection_indices(points, H=64, W=1024, fov_down=-30, fov_up=10, lidar_position=(1, 0, 2)):
    """Per-point ``(proj_h, proj_w, depth64)``.  muvo/utils/geometry_utils.py:167-200."""
    up = fov_up / 180.0 * np.pi                                     # :168
    down = fov_down / 180.0 * np.pi                                 # :169
    fov = up - down                                                 # :170
    lidar = np.asarray(lidar_position)
    carla = points * np.array([1, -1, 1])                           # :177 (promotes to float64)
    carla -= lidar                                                  # :178
    depth = np.linalg.norm(carla, 2, axis=1)                        # :180  sqrt((x^2+y^2)+z^2)
    x = carla[:, 0]
    y = -carla[:, 1]                                                # :183
    z = carla[:, 2]
    yaw = np.arctan2(y, x)                                          # :186
    pitch = np.arcsin(z / depth)                                    # :187
    pw = 0.5 * (1.0 - yaw / np.pi)                                  # :189
    ph = 1.0 - (pitch + abs(down)) / fov                            # :190
    pw *= W
    ph *= H
    pw = np.maximum(0, np.minimum(W - 1, np.floor(pw))).astype(np.int32)   # :194-196
    ph = np.maximum(0, np.minimum(H - 1, np.floor(ph))).astype(np.int32)   # :198-200
    return ph, pw, depth


def range_projection(points, semantics, H=64, W=1024, fov_down=-30, fov_up=10, lidar_position=(1, 0, 2)):
    """``PointCloud.do_range_projection`` (muvo/utils/geometry_utils.py:175-220).

    Returns ``range_depth (H,W) f32`` (-1 empty), ``range_xyz (H,W,3) f32`` (input
    ego-frame xyz of the nearest point), ``range_sem (H,W) u8``.
    """
    points = np.asarray(points)
    semantics = np.asarray(semantics)
    ph, pw, depth = range_projection_indices(points, H, W, fov_down, fov_up, lidar_position)
    order = np.argsort(depth, kind="stable")[::-1]                  # :203 far -> near; near written last
    range_depth = np.full((H, W), -1, dtype=np.float32)
    range_xyz = np.full((H, W, 3), 0, dtype=np.float32)
    range_sem = np.full((H, W), 0, dtype=np.uint8)
    range_depth[ph[order], pw[order]] = depth[order]                # :217
    range_xyz[ph[order], pw[order]] = points[order]                 # :218
    range_sem[ph[order], pw[order]] = semantics[order]              # :219
    return range_depth, range_xyz, range_sem


def pack_range_view(range_depth, range_xyz):
    """muvo/data/dataset.py:301-303: ``(4,H,W)`` float32 = x, y, z, depth."""
    return np.ascontiguousarray(
        np.concatenate([range_xyz, range_depth[..., None]], axis=-1).transpose((2, 0, 1)))


# --------------------------------------------------------------------------- (c)
def bev_intrinsics(size, scale, offsetx):
    """muvo/utils/geometry_utils.py:8-19."""
    return np.array([[1 / scale, 0, size[0] / 2 + offsetx],
                     [0, -1 / scale, size[1] / 2],
                     [0, 0, 1]], dtype=np.float32)


def gen_dx_bx(size, scale, offsetx):
    """muvo/models/frustum_pooling.py:10-20."""
    xb = [-size[0] * scale / 2 - offsetx * scale, size[0] * scale / 2 - offsetx * scale, scale]
    yb = [-size[1] * scale / 2, size[1] * scale / 2, scale]
    zb = [-10.0, 10.0, 20.0]
    rows = [xb, yb, zb]
    dx = torch.Tensor([r[2] for r in rows])
    bx = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    nx = torch.LongTensor([np.round((r[1] - r[0]) / r[2]) for r in rows])
    return dx, bx, nx


def frustum_grid(dbound, fH, fW, downsample):
    """muvo/models/frustum_pooling.py:89-109: ``(D,fH,fW,3)`` of (u, v, depth)."""
    ds = torch.arange(dbound[0], dbound[1], dbound[2], dtype=torch.float32)
    D = len(ds)
    og_h, og_w = fH * downsample, fW * downsample
    d = ds.view(-1, 1, 1).expand(-1, fH, fW)
    u = torch.linspace(0, og_w - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    v = torch.linspace(0, og_h - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((u, v, d), -1)


def _intrinsics_inverse(k):
    """muvo/utils/geometry_utils.py:22-34 (closed form, not torch.inverse)."""
    fx, fy, cx, cy = k[..., 0, 0], k[..., 1, 1], k[..., 0, 2], k[..., 1, 2]
    one, zero = torch.ones_like(fx), torch.zeros_like(fx)
    return torch.stack((torch.stack((1 / fx, zero, -cx / fx), -1),
                        torch.stack((zero, 1 / fy, -cy / fy), -1),
                        torch.stack((zero, zero, one), -1)), -2)


def frustum_geometry(frustum, intrinsics, pose):
    """muvo/models/frustum_pooling.py:111-129,204-206 -> ``(B,N,D,H,W,3)`` ego xyz."""
    rots = pose[..., :3, :3]
    trans = pose[..., :3, 3:]
    B, N = trans.shape[:2]
    pts = frustum.unsqueeze(0).unsqueeze(0).unsqueeze(-1)
    pts = torch.cat((pts[:, :, :, :, :, :2] * pts[:, :, :, :, :, 2:3], pts[:, :, :, :, :, 2:3]), 5)
    combine = rots.matmul(_intrinsics_inverse(intrinsics))
    pts = combine.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
    pts = pts + trans.view(B, N, 1, 1, 1, 3)
    return pts


def bev_cell_ids(geom, bev_intr, bx, dx):
    """muvo/models/frustum_pooling.py:139-146: affine + ``.long()`` (truncation)."""
    g = geom.reshape(-1, 3).clone()
    g[:, 0] = g[:, 0] * bev_intr[0, 0] + bev_intr[0, 2]
    g[:, 1] = g[:, 1] * bev_intr[1, 1] + bev_intr[1, 2]
    g[:, 2] = (g[:, 2] - bx[2] + dx[2] / 2.) / dx[2]
    return g.long()


def cumsum_trick(x, geom, ranks):
    """muvo/models/frustum_pooling.py:23-31 (== QuickCumsum.forward :36-42, VoxelsSumming layers.py:329-341)."""
    x = x.cumsum(0)
    kept = torch.ones(x.shape[0], dtype=torch.bool)
    kept[:-1] = ranks[1:] != ranks[:-1]
    x, geom = x[kept], geom[kept]
    x = torch.cat((x[:1], x[1:] - x[:-1]))
    return x, geom


def quick_cumsum_backward(grad_seg, ranks):
    """muvo/models/frustum_pooling.py:52-60: ``grad_x[i] = grad_seg[segment(i)]``."""
    kept = torch.ones(ranks.shape[0], dtype=torch.bool)
    kept[:-1] = ranks[1:] != ranks[:-1]
    back = torch.cumsum(kept, 0)
    back[kept] -= 1
    return grad_seg[back]


def _pool_prepare(geom, x, mask, bev_intr, bx, dx, nx):
    B, N, D, H, W, C = x.shape
    n_prime = B * N * D * H * W
    xf = x.reshape(n_prime, C)                                      # :136
    cells = bev_cell_ids(geom, bev_intr, bx, dx)                    # :139-146
    batch_ix = torch.arange(B).repeat_interleave(n_prime // B).view(-1, 1)   # :148-149
    cells = torch.cat((cells, batch_ix), 1)                         # :150
    if len(mask) > 0:                                               # :153-156
        m = mask.reshape(n_prime).bool()
        xf, cells = xf[m], cells[m]
    kept = ((cells[:, 0] >= 0) & (cells[:, 0] < nx[0]) & (cells[:, 1] >= 0) & (cells[:, 1] < nx[1])
            & (cells[:, 2] >= 0) & (cells[:, 2] < nx[2]))           # :159-161
    xf, cells = xf[kept], cells[kept]
    ranks = (cells[:, 0] * (nx[1] * nx[2] * B) + cells[:, 1] * (nx[2] * B)
             + cells[:, 2] * B + cells[:, 3])                       # :166-169
    order = ranks.argsort(stable=True)                              # :170
    return xf[order], cells[order], ranks[order], (B, C)


def voxel_pooling_cumsum(geom, x, mask, bev_intr, bx, dx, nx):
    """fp32 cumsum-trick pooling, the reference's own arithmetic (frustum_pooling.py:131-187)."""
    xf, cells, ranks, (B, C) = _pool_prepare(geom, x, mask, bev_intr, bx, dx, nx)
    nxl = [int(v) for v in nx]
    final = torch.zeros((B, C, nxl[2], nxl[1], nxl[0]), dtype=xf.dtype)
    if xf.shape[0] > 0:
        xs, cs = cumsum_trick(xf, cells, ranks)                     # :174-177
        final[cs[:, 3], :, cs[:, 2], cs[:, 1], cs[:, 0]] = xs       # :180-182
    return torch.cat(final.unbind(dim=2), 1)                        # :185


def voxel_pooling_exact(geom, x, mask, bev_intr, bx, dx, nx):
    """Same pooling with a float64 direct segment sum (accuracy yardstick)."""
    xf, cells, ranks, (B, C) = _pool_prepare(geom, x, mask, bev_intr, bx, dx, nx)
    nxl = [int(v) for v in nx]
    final = torch.zeros((B * nxl[2] * nxl[1] * nxl[0], C), dtype=torch.float64)
    flat = ((cells[:, 3] * nxl[2] + cells[:, 2]) * nxl[1] + cells[:, 1]) * nxl[0] + cells[:, 0]
    final.index_add_(0, flat, xf.double())
    final = final.view(B, nxl[2], nxl[1], nxl[0], C).permute(0, 4, 1, 2, 3)
    return torch.cat(final.unbind(dim=2), 1).contiguous()


def frustum_pooling_forward(x, intrinsics, pose, mask, size, scale, offsetx, dbound, downsample, exact=False):
    """``FrustumPooling.forward`` (muvo/models/frustum_pooling.py:189-209) as a pure function."""
    bev_k = torch.tensor(bev_intrinsics(size, scale, offsetx))
    dx, bx, nx = gen_dx_bx(size, scale, offsetx)
    fH, fW = x.shape[-3:-1]
    frustum = frustum_grid(dbound, fH, fW, downsample)
    geom = frustum_geometry(frustum, intrinsics, pose)
    fn = voxel_pooling_exact if exact else voxel_pooling_cumsum
    return fn(geom, x, mask, bev_k, bx, dx, nx)


# --------------------------------------------------------------------------- (d)
def ssc_counts(pred, target, n_classes, nonempty=None, nonsurface=None, ignore255=False):
    """Counts behind ``SSCMetrics`` as one confusion histogram (muvo/metrics.py:143-216).

    Returns int64 ``[3 + 3C]`` = completion (tp, fp, fn), then tp[C], fp[C], fn[C].
    ``nonempty`` restricts both families, ``nonsurface`` only completion (:79-95);
    ``ignore255=True`` reproduces ``add_batch``'s ``mask = y_true != 255`` (:79,:90).
    With no mask at all, target==255 voxels are rewritten to (0,0) (:150-151,:184-185)
    and therefore count as a class-0 true positive, exactly as the reference does.
    """
    p = np.asarray(pred).reshape(-1).astype(np.int64)
    t = np.asarray(target).reshape(-1).astype(np.int64)
    is255 = t == 255
    p = np.where(is255, 0, p)
    t = np.where(is255, 0, t)
    sem_valid = np.ones(t.shape, bool)
    if ignore255:
        sem_valid &= ~is255
    if nonempty is not None:
        sem_valid &= np.asarray(nonempty).reshape(-1).astype(bool)
    comp_valid = sem_valid.copy()
    if nonsurface is not None:
        comp_valid &= np.asarray(nonsurface).reshape(-1).astype(bool)
    C = int(n_classes)
    out = np.zeros(3 + 3 * C, dtype=np.int64)
    bt, bp = (t > 0) & comp_valid, (p > 0) & comp_valid
    out[0] = np.count_nonzero(bt & bp)
    out[1] = np.count_nonzero(~(t > 0) & bp)
    out[2] = np.count_nonzero(bt & ~(p > 0))
    ts, ps = t[sem_valid], p[sem_valid]
    for j in range(C):
        out[3 + j] = np.count_nonzero((ts == j) & (ps == j))
        out[3 + C + j] = np.count_nonzero((ts != j) & (ps == j))
        out[3 + 2 * C + j] = np.count_nonzero((ts == j) & (ps != j))
    return out


def ssc_counts_loop(pred, target, n_classes, nonempty=None, nonsurface=None, ignore255=False):
    """Frame-by-frame / class-by-class torch port of muvo/metrics.py:143-216 (the timed CPU baseline)."""
    pred = torch.as_tensor(pred)
    target = torch.as_tensor(target)
    C = int(n_classes)
    bs = pred.shape[0]
    sem_mask = None
    if ignore255:
        sem_mask = target != 255
    if nonempty is not None:
        ne = torch.as_tensor(nonempty).bool()
        sem_mask = ne if sem_mask is None else (sem_mask & ne)
    comp_mask = sem_mask
    if nonsurface is not None:
        ns = torch.as_tensor(nonsurface).bool()
        comp_mask = ns if comp_mask is None else (comp_mask & ns)
    out = np.zeros(3 + 3 * C, dtype=np.int64)
    p = pred.clone()
    t = target.clone()
    p[t == 255] = 0
    t[t == 255] = 0
    t = t.reshape(bs, -1)
    p = p.reshape(bs, -1)
    b_pred = p.new_zeros(p.shape)
    b_true = t.new_zeros(t.shape)
    b_pred[p > 0] = 1
    b_true[t > 0] = 1
    for i in range(bs):
        yt, yp = b_true[i], b_pred[i]
        if comp_mask is not None:
            sel = comp_mask[i].reshape(-1)
            yt, yp = yt[sel], yp[sel]
        out[0] += torch.stack(torch.where(torch.logical_and(yt == 1, yp == 1))).numel()
        out[1] += torch.stack(torch.where(torch.logical_and(yt != 1, yp == 1))).numel()
        out[2] += torch.stack(torch.where(torch.logical_and(yt == 1, yp != 1))).numel()
    for i in range(bs):
        yt, yp = t[i], p[i]
        if sem_mask is not None:
            sel = sem_mask[i].reshape(-1)
            yt, yp = yt[sel], yp[sel]
        for j in range(C):
            out[3 + j] += torch.stack(torch.where(torch.logical_and(yt == j, yp == j))).numel()
            out[3 + C + j] += torch.stack(torch.where(torch.logical_and(yt != j, yp == j))).numel()
            out[3 + 2 * C + j] += torch.stack(torch.where(torch.logical_and(yt == j, yp != j))).numel()
    return out


def ssc_add_batch_counts(pred, target, n_classes, nonempty=None, nonsurface=None):
    """Counts accumulated by one ``SSCMetrics.add_batch`` call (muvo/metrics.py:77-100)."""
    return ssc_counts(pred, target, n_classes, nonempty, nonsurface, ignore255=True)


def ssc_stats_from_counts(counts, n_classes):
    """muvo/metrics.py:102-121 applied to an int64 count vector (float32 running sums as the reference)."""
    C = int(n_classes)
    tp, fp, fn = (int(v) for v in counts[:3])
    if tp != 0:
        precision, recall, iou = tp / (tp + fp), tp / (tp + fn), tp / (tp + fp + fn)
    else:
        precision, recall, iou = 0, 0, 0
    tps = torch.zeros(C) + torch.as_tensor(counts[3:3 + C].astype(np.int32))
    fps = torch.zeros(C) + torch.as_tensor(counts[3 + C:3 + 2 * C].astype(np.int32))
    fns = torch.zeros(C) + torch.as_tensor(counts[3 + 2 * C:3 + 3 * C].astype(np.int32))
    iou_ssc = tps / (tps + fps + fns + 1e-5)
    return {"precision": precision, "recall": recall, "iou": iou, "iou_ssc": iou_ssc,
            "iou_ssc_mean": torch.mean(iou_ssc[1:])}


# ---------------------------------------------------------------------------------------------
# N4: SemScalLoss / GeoScalLoss (muvo/losses.py:191-287), float64 restatement
# ---------------------------------------------------------------------------------------------
def _softmax64(prediction):
    """F.softmax(prediction, dim=1) (losses.py:205, :268) of float32 logits, evaluated in float64."""
    x = np.asarray(prediction, dtype=np.float64)
    x = x - x.max(axis=1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=1, keepdims=True)


def _bce_to_one(x):
    """F.binary_cross_entropy(x, ones) (losses.py:230-249, :284-286): -log(x) with the log clamped at -100."""
    with np.errstate(divide="ignore"):
        return -max(np.log(x), -100.0)


def sem_scal_loss(prediction, target, ignore_index=255):
    """SemScalLoss.forward, losses.py:199-251.  prediction (b,s,c,x,y,z) logits, target (b,s,x,y,z) labels."""
    b, s, c = prediction.shape[:3]
    p_all = _softmax64(np.asarray(prediction).reshape((b * s, c) + tuple(prediction.shape[3:])))
    tgt = np.asarray(target).reshape((b * s,) + tuple(prediction.shape[3:]))
    mask = tgt != ignore_index                                            # :208
    loss, count = 0.0, 0.0
    for i in range(c):                                                    # :210
        p = p_all[:, i][mask]                                             # :213-217
        t = (tgt[mask] == i).astype(np.float64)                           # :220-221
        if t.sum() > 0:                                                   # :224
            count += 1.0
            nominator = (p * t).sum()
            loss_class = 0.0
            if p.sum() > 0:                                               # :229
                precision = nominator / p.sum()
                if 0 <= precision <= 1:
                    loss_class += _bce_to_one(precision)
            recall = nominator / t.sum()                                  # :236-237
            if 0 <= recall <= 1:
                loss_class += _bce_to_one(recall)
            if (1 - t).sum() > 0:                                         # :243
                specificity = ((1 - p) * (1 - t)).sum() / (1 - t).sum()
                if 0 <= specificity <= 1:
                    loss_class += _bce_to_one(specificity)
            loss += loss_class
    return loss / count                                                   # :251 (ZeroDivisionError when no class is present)


def geo_scal_loss(prediction, target, ignore_index=255):
    """GeoScalLoss.forward, losses.py:259-287."""
    b, s, c = prediction.shape[:3]
    p_all = _softmax64(np.asarray(prediction).reshape((b * s, c) + tuple(prediction.shape[3:])))
    tgt = np.asarray(target).reshape((b * s,) + tuple(prediction.shape[3:]))
    mask = tgt != ignore_index                                            # :275
    empty = p_all[:, 0][mask]                                             # :271, :279
    nonempty = 1.0 - empty
    nt = (tgt != 0)[mask].astype(np.float64)                              # :276-277
    inter = (nt * nonempty).sum()                                         # :281
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = inter / nonempty.sum()
        recall = inter / nt.sum()
        spec = ((1 - nt) * empty).sum() / (1 - nt).sum()
    return _bce_to_one(precision) + _bce_to_one(recall) + _bce_to_one(spec)


def scal_sums(prediction, target, ignore_index=255):
    """The 3C+1 scalars both losses are functions of: sum_p[C], nom[C], cnt[C], n_valid (float64)."""
    b, s, c = prediction.shape[:3]
    p_all = _softmax64(np.asarray(prediction).reshape((b * s, c) + tuple(prediction.shape[3:])))
    tgt = np.asarray(target).reshape((b * s,) + tuple(prediction.shape[3:]))
    mask = tgt != ignore_index
    out = np.zeros(3 * c + 1)
    for i in range(c):
        p = p_all[:, i][mask]
        hit = tgt[mask] == i
        out[i], out[c + i], out[2 * c + i] = p.sum(), p[hit].sum(), hit.sum()
    out[3 * c] = mask.sum()
    return out


# ---------------------------------------------------------------------------------------------
# N4: torch_scatter reductions used by the PointPillar encoder (muvo/models/common.py:703, :731)
# torch_scatter is a third-party dependency of the reference (not vendored, no pinned version in requirements.txt);
# these restate its documented semantics for dim=0.  PARITY UNPINNED: the package is not installable here.
# ---------------------------------------------------------------------------------------------
def scatter_mean(src, index, dim_size=None):
    """out[m] = mean of src rows with index == m (float64 accumulation); rows that receive nothing are 0."""
    src = np.asarray(src, dtype=np.float64)
    index = np.asarray(index).reshape(-1)
    M = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.size else 0)
    out = np.zeros((M, src.shape[1]))
    np.add.at(out, index, src)
    cnt = np.bincount(index, minlength=M).astype(np.float64)
    return out / np.maximum(cnt, 1.0)[:, None]


def scatter_max(src, index, dim_size=None):
    """(max, arg): per output row and feature the maximum over the source rows with index == m and the LOWEST source row
    attaining it; rows that receive nothing are 0 with arg = N."""
    src = np.asarray(src)
    index = np.asarray(index).reshape(-1)
    N, F = src.shape
    M = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.size else 0)
    out = np.zeros((M, F), dtype=src.dtype)
    arg = np.full((M, F), N, dtype=np.int64)
    seen = np.zeros(M, dtype=bool)
    for n in range(N):                                   # small inputs only
        m = index[n]
        if not seen[m]:
            out[m], arg[m], seen[m] = src[n], n, True
        else:
            better = src[n] > out[m]
            out[m] = np.where(better, src[n], out[m])
            arg[m] = np.where(better, n, arg[m])
    return out, arg


def _self_check():  # pragma: no cover - quick sanity when run directly
    P = np.array([(-48, -48, -6), (47.99, 47.99, 25.99), (48, 0, 0), (0, 0, 26), (0, 0, -6.01), (0.1, 0.1, 0.1),
                  (0.4, 0.4, 0.4), (0.26, 0.26, 0.26), (1.05, 0, 0), (1.45, 0.45, 0.45)])
    S = np.array([1, 2, 3, 4, 5, 7, 8, 9, 10, 6], dtype=np.uint8)
    print(voxel_filter_loop(P, S, 0.5, [192, 192, 64], [0.0, 0, -10.0]))
    print(voxel_filter_fast(P, S, 0.5, [192, 192, 64], [0.0, 0, -10.0]))
    print(math.pi)


if __name__ == "__main__":  # pragma: no cover
    _self_check()

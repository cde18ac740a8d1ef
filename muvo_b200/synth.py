"""Seeded synthetic inputs shaped like MUVO's CARLA data.

No dataset or checkpoint is available offline, so tests and ``bench.py`` use these
generators.  Everything is ``numpy.random.default_rng(seed)`` / ``torch.Generator``
driven; seed convention = ``1000 * config + frame index``.
"""
from __future__ import annotations

import numpy as np

LIDAR_POSITION = (1.0, 0.0, 2.0)          # muvo/config.py:85, data/data_preprocess.yaml:10
RANGE_H, RANGE_W = 64, 1024               # muvo/config.py:88,90
RANGE_FOV = (-30, 10)                     # muvo/config.py:87
VOXEL_SIZE = (192, 192, 64)               # data/data_preprocess.yaml:13, muvo/config.py:103
VOXEL_RES = 0.5                           # data/data_preprocess.yaml:12
VOXEL_OFFSET = (0.0, 0, -10.0)            # data/generate_voxels.py:66-68 with offset_z=-20 px
# muvo/constants.py:180-204 (LABEL_MAP): every raw CARLA tag except 0 (unlabeled) and 13 (sky) -> 1
LABEL_MAP = {k: (0 if k in (0, 13) else 1) for k in range(23)}


def label_remap256() -> np.ndarray:
    """256-entry version of the dataset remap (muvo/data/dataset.py:281-283,323): 255 -> 0."""
    tab = np.full(256, max(LABEL_MAP.values()), dtype=np.uint8)
    for k, v in LABEL_MAP.items():
        tab[k] = v
    tab[255] = 0
    return tab


def carla_lidar_frame(n_points: int, seed: int, beams: int = 64):
    """One CARLA-style semantic-LiDAR sweep in the EGO frame (after ``convert_coor_lidar``).

    64 beams with pitch linspace(-30 deg, +10 deg), uniform azimuth; each ray hits the
    ground plane (LiDAR 2 m above it) or a per-azimuth-sector wall, whichever is
    nearer; rays hitting nothing are dropped and resampled.  Points come out in SCAN ORDER (channel-major,
    azimuth ascending), as CARLA's ray-cast LiDAR delivers them (carla_gym/.../lidar/ray_cast_semantic.py:197-219).  Returns
    ``(points float32 (n,3), semantics uint8 (n,))`` with raw CARLA tags: road 7,
    road lines 6 (stripes), walls 1, 2 % vehicles 10.
    """
    rng = np.random.default_rng(seed)
    n_sectors = 64
    wall_dist = rng.uniform(5.0, 80.0, n_sectors)
    wall_height = rng.uniform(3.0, 15.0, n_sectors)
    pitches = np.deg2rad(np.linspace(RANGE_FOV[0], RANGE_FOV[1], beams))
    pts, sems, keys = [], [], []
    have = 0
    while have < n_points:
        m = int((n_points - have) * 1.6) + 64
        beam = rng.integers(0, beams, m)
        pitch = pitches[beam]
        az = rng.uniform(-np.pi, np.pi, m)
        sector = np.minimum(((az + np.pi) / (2 * np.pi) * n_sectors).astype(np.int64), n_sectors - 1)
        with np.errstate(divide="ignore"):
            r_ground = np.where(pitch < 0, 2.0 / -np.sin(pitch), np.inf)
        r_wall = wall_dist[sector] / np.cos(pitch)
        wall_ok = (2.0 + r_wall * np.sin(pitch)) < wall_height[sector]
        r_wall = np.where(wall_ok, r_wall, np.inf)
        r = np.minimum(r_ground, r_wall)
        hit_ground = r_ground <= r_wall
        ok = np.isfinite(r)
        r = np.clip(r + rng.normal(0.0, 0.02, m), 0.5, 100.0)
        # CARLA LiDAR frame: x forward, y right, z up
        x = r * np.cos(pitch) * np.cos(az)
        y = r * np.cos(pitch) * np.sin(az)
        z = r * np.sin(pitch)
        p = np.stack([x, y, z], 1).astype(np.float32)[ok]
        sem = np.where(hit_ground, 7, 1).astype(np.uint8)[ok]
        pts.append(p)
        sems.append(sem)
        keys.append(np.stack([beam[ok].astype(np.float64), az[ok]], 1))
        have += p.shape[0]
    p = np.concatenate(pts)[:n_points]
    sem = np.concatenate(sems)[:n_points]
    key = np.concatenate(keys)[:n_points]
    order = np.lexsort((key[:, 1], key[:, 0]))          # channel-major, azimuth ascending
    p, sem = p[order], sem[order]
    # data/data_preprocessing.py:119-122 (convert_coor_lidar): += lidar_pos ; y *= -1 (float32)
    p = p + np.asarray(LIDAR_POSITION, dtype=np.float32)
    p[:, 1] *= -1
    ground = sem == 7
    sem[ground & (np.mod(np.abs(p[:, 1]), 3.5) < 0.15)] = 6
    sem[rng.random(n_points) < 0.02] = 10
    return np.ascontiguousarray(p), sem


def lidar_batch(n_frames: int, n_min: int, n_max: int, seed0: int):
    """Ragged batch: concatenated points/semantics + ``frame_offsets int64 [F+1]``."""
    rng = np.random.default_rng(seed0 + 999_983)
    sizes = rng.integers(n_min, n_max + 1, n_frames) if n_max > n_min else np.full(n_frames, n_min)
    pts, sems = [], []
    for f in range(n_frames):
        p, s = carla_lidar_frame(int(sizes[f]), seed0 + f)
        pts.append(p)
        sems.append(s)
    offsets = np.zeros(n_frames + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    return np.concatenate(pts), np.concatenate(sems), offsets


def carla_depth_image(seed: int, h: int = 600, w: int = 960):
    """CARLA-style encoded depth + semantic image as ``cv2.imread(file, -1)`` returns it: uint8 ``(h, w, 4)`` with
    depth = 1000 * (R + 256 G + 256^2 B) / (256^3 - 1) metres in channels [B, G, R] = img[..., :3] read as
    ``depth_color[..., 2], [..., 1], [..., 0]`` (data/data_preprocessing.py:72-77) and the semantic tag in channel 3.
    Ground plane below the camera, a far wall, sky (depth 1000 -> dropped), a few exact-range and near-ego pixels."""
    rng = np.random.default_rng(seed)
    v, u = np.mgrid[0:h, 0:w]
    f = w / (2.0 * np.tan(110 * np.pi / 360.0))
    ray_y = (v - h / 2.0) / f                               # image y down
    with np.errstate(divide="ignore", invalid="ignore"):
        d_ground = np.where(ray_y > 1e-3, 2.0 / ray_y, np.inf)      # camera 2 m above the ground
    d_wall = rng.uniform(15.0, 90.0) + 5.0 * np.sin(u / 37.0)
    depth = np.minimum(d_ground, d_wall) + rng.normal(0, 0.01, (h, w))
    depth = np.clip(depth, 0.3, 1000.0)
    depth[: h // 6] = 1000.0                                 # sky
    depth[rng.random((h, w)) < 0.01] = rng.uniform(0.5, 3.0)  # close clutter (some of it inside the ego box)
    code = np.round(depth / 1000.0 * (256 ** 3 - 1)).astype(np.int64)
    code = np.clip(code, 0, 256 ** 3 - 1)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    img[..., 0] = code & 255                                  # depth_color[..., 0]
    img[..., 1] = (code >> 8) & 255
    img[..., 2] = (code >> 16) & 255
    img[..., 3] = np.where(depth >= 1000.0, 13, np.where(d_ground < d_wall, 7, 1)).astype(np.uint8)
    return img


def muvo_camera():
    """Cropped intrinsics / extrinsics used by the BEV lift at muvo.yml geometry.

    muvo/utils/geometry_utils.py:64-91 with IMAGE.FOV=100, SIZE=(600,960), camera at
    (1,0,2) (muvo/config.py:111-114), then the crop shift of preprocess.py:244-248
    (left=64, top=138).
    """
    f = 960 / (2 * np.tan(100 * np.pi / 360.0))
    K = np.float32([[f, 0, 960 / 2], [0, f, 600 / 2], [0, 0, 1]])
    K[0, 2] -= 64
    K[1, 2] -= 138
    E = np.float32([[0, 0, 1, 1.0], [-1, 0, 0, -0.0], [0, -1, 0, 2.0], [0, 0, 0, 1]])
    return K, E


BEV_POOL_ARGS = dict(size=(48, 48), scale=0.8, offsetx=-16.0, dbound=[1.0, 38.0, 1.0], downsample=8)  # mile.py:37-43


def bev_inputs(batch_frames: int, channels: int, seed: int, fH: int = 40, fW: int = 104, D: int = 37,
               topk: int = 10, device="cpu", dtype=None):
    """cfg3 inputs: ``feat ~ N(0,1)``, ``depth = softmax(N(0,1))`` and the top-k depth mask.

    Returns ``(feat (B,C,fH,fW), depth (B,D,fH,fW), mask bool (B,D,fH,fW), K (B,3,3), E (B,4,4))``;
    the lifted tensor is built by the caller exactly as muvo/models/mile.py:517-521.
    """
    import torch
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(batch_frames, channels, fH, fW, generator=g)
    depth = torch.randn(batch_frames, D, fH, fW, generator=g).softmax(dim=1)
    if topk and topk > 0:
        bins = depth.topk(topk, dim=1)[1]                       # mile.py:512
        mask = torch.zeros(depth.shape, dtype=torch.bool)
        mask.scatter_(1, bins, 1)                               # mile.py:513-514
    else:
        mask = torch.zeros(0)
    K, E = muvo_camera()
    K = torch.from_numpy(K).expand(batch_frames, 3, 3).contiguous()
    E = torch.from_numpy(E).expand(batch_frames, 4, 4).contiguous()
    out = [feat, depth, mask, K, E]
    if dtype is not None:
        out[0], out[1] = out[0].to(dtype), out[1].to(dtype)
    return tuple(t.to(device) for t in out)


def lift(feat, depth):
    """muvo/models/mile.py:517-521: outer product -> strided ``(B,1,D,H,W,C)`` view."""
    x = (depth.unsqueeze(1) * feat.unsqueeze(2)).type_as(feat)
    return x.unsqueeze(1).permute(0, 1, 3, 4, 5, 2)


def occupancy_pair(n_frames: int, n_classes: int, seed: int, size=VOXEL_SIZE):
    """cfg4 inputs: ``y_true`` uint8 (5 % occupied, 0.1 % ignore=255), ``y_pred`` int64 in [0,C)."""
    rng = np.random.default_rng(seed)
    shape = (n_frames,) + tuple(size)
    u = rng.random(shape, dtype=np.float32)
    y_true = np.zeros(shape, dtype=np.uint8)
    occ = u < 0.05
    if n_classes > 2:
        y_true[occ] = rng.integers(1, n_classes, int(occ.sum()), dtype=np.uint8)
    else:
        y_true[occ] = 1
    y_true[u > 0.999] = 255
    # argmax of N(0,1) logits over C classes == uniform class draw
    y_pred = rng.integers(0, n_classes, shape, dtype=np.int64)
    return y_pred, y_true

"""Stages (a) voxelisation and (b) range-view projection: host side.

Drop-in replacements (same names / arguments / returns as the reference):

* :func:`voxel_filter`                  <- ``data/data_preprocessing.py:172-228``
* :class:`PointCloud` ``.do_range_projection`` <- ``muvo/utils/geometry_utils.py:167-220``

plus the batched device API the B200 pipeline actually uses
(:func:`sensor_to_grid`: ragged frames in, dense grids + range images out, one
read of the point stream).  All arithmetic happens in ``libmuvo_b200.so``;
this module only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, _ws

ROADLINE_ID = 6  # data/data_preprocessing.py:198: np.where(LABEL_CLASS == 'roadlines')[0][0]


@dataclass(frozen=True)
class GridSpec:
    """(voxel_resolution, voxel_size, offset) of ``voxel_filter`` (data_preprocessing.py:172)."""
    voxel_resolution: float = 0.5
    voxel_size: Sequence[int] = (192, 192, 64)
    offset: Sequence[float] = (0.0, 0.0, -10.0)
    roadline_id: int = ROADLINE_ID

    def to_c(self) -> _lib.MuvoGrid:
        size = np.asarray(self.voxel_size)
        res = np.asarray(self.voxel_resolution)
        # data_preprocessing.py:176-178, evaluated by numpy in float64 exactly as the reference does;
        # the caller's offset object is never mutated.
        off = np.array(self.offset, dtype=np.float64, copy=True)
        off = off + res * size / 2
        upper = (size * res).astype(np.float64)
        g = _lib.MuvoGrid()
        g.res = float(res)
        for k in range(3):
            g.offset[k] = float(off[k])
            g.upper[k] = float(upper[k])
            g.size[k] = int(size[k])
        g.roadline_id = int(self.roadline_id)
        return g

    @property
    def n_voxels(self) -> int:
        return int(np.prod(np.asarray(self.voxel_size, dtype=np.int64)))


@dataclass(frozen=True)
class RangeSpec:
    """Constructor arguments of ``PointCloud`` (geometry_utils.py:167-173)."""
    H: int = 64
    W: int = 1024
    fov_down: float = -30
    fov_up: float = 10
    lidar_position: Sequence[float] = (1, 0, 2)

    def to_c(self) -> _lib.MuvoRangeCfg:
        up = self.fov_up / 180.0 * np.pi         # :168
        down = self.fov_down / 180.0 * np.pi     # :169
        c = _lib.MuvoRangeCfg()
        c.H, c.W = int(self.H), int(self.W)
        c.fov_down_abs = float(abs(down))        # :190
        c.fov = float(up - down)                 # :170
        lp = np.asarray(self.lidar_position, dtype=np.float64)
        for k in range(3):
            c.lidar_pos[k] = float(lp[k])
        return c


def _as_offsets(frame_offsets, n_points: int, device) -> torch.Tensor:
    if frame_offsets is None:
        return torch.tensor([0, n_points], dtype=torch.int64, device=device)
    if isinstance(frame_offsets, torch.Tensor):
        off = frame_offsets.to(device=device, dtype=torch.int64)
    else:
        off = torch.as_tensor(np.asarray(frame_offsets, dtype=np.int64), device=device)
    return off.contiguous()


@dataclass(frozen=True)
class LidarPrep:
    """LiDAR-side prep of ``muvo/data/dataset.py:275-290`` fused into the range projection (``sensor_to_grid(...,
    lidar_prep=)``): ``convert_coor_lidar(points_xyz, lidar_position)``, ``remap[ObjTag]`` and the ego-box drop."""
    lidar_position: Sequence[float] = (1.0, 0.0, 2.0)            # cfg.POINTS.LIDAR_POSITION (muvo/config.py:85)
    ego_dimension: Optional[Sequence[float]] = (4.902, 2.128, 1.511)   # constants.py:8; None = keep every point
    remap: Optional[np.ndarray] = None                            # 256-entry uint8 table (dataset.py:281-283) or None

    def to_c(self, dev):
        c = _lib.MuvoLidarPrep()
        keep = None
        for k in range(3):
            c.add[k] = float(self.lidar_position[k])
        if self.ego_dimension is not None:
            x, y, z = (float(v) for v in self.ego_dimension)
            lo, hi = np.array([[-x / 2, -y / 2, 0], [x / 2, y / 2, z]])          # dataset.py:287, evaluated by numpy
            for k in range(3):
                c.box_lo[k], c.box_hi[k] = float(lo[k]), float(hi[k])
            c.use_ego_box = 1
        if self.remap is not None:
            tab = np.asarray(self.remap, dtype=np.uint8).reshape(-1)
            if tab.size < 256:                                                   # remap tables cover the raw tags only
                tab = np.concatenate([tab, np.full(256 - tab.size, tab.max() if tab.size else 0, np.uint8)])
            keep = torch.from_numpy(np.ascontiguousarray(tab[:256])).to(dev)
            c.remap256 = keep.data_ptr()
        return c, keep


def sensor_to_grid(points: torch.Tensor, semantics: torch.Tensor, frame_offsets=None, *,
                   grid: Optional[GridSpec] = None, range_spec: Optional[RangeSpec] = None,
                   dense: bool = True, sparse: bool = False, remap: Optional[torch.Tensor] = None,
                   layout: str = "xyzd", want_diag: bool = False, out: Optional[dict] = None,
                   packed_sparse: bool = False, lidar_prep: Optional[LidarPrep] = None) -> dict:
    """Batched (a)+(b) on the current CUDA stream.  Nothing synchronises.

    points ``(P,3)`` float32 (float64 allowed when only ``grid`` is given), semantics ``(P,)`` uint8,
    ``frame_offsets`` int64 ``[F+1]`` (device tensor, or host sequence).  Returns a dict with
    ``voxel (F,Dx,Dy,Dz) u8`` | ``voxel_sparse (P,4) u16`` + ``n_occ (F,) i64`` and
    ``range_xyzd (F,4,H,W) f32`` + ``range_sem (F,H,W) u8`` (layout "xyzd") or
    ``range_depth (F,H,W)``, ``range_xyz (F,H,W,3)``, ``range_sem`` (layout "hwc").
    ``out`` may carry preallocated tensors under the same keys.  Frame f's sparse rows start at row
    ``frame_offsets[f]``; with ``packed_sparse`` the frames' lists are stored back to back instead and
    ``sparse_start (F+1,) i64`` gives the first row of every frame (a read-back then moves only the rows used).
    ``lidar_prep`` (range-only calls): ``points`` / ``semantics`` are the RAW sweep (LiDAR frame, CARLA tags) and the prep of
    dataset.py:275-290 happens inside the kernels; ``range_xyz*`` then holds the converted points, ``range_sem`` the remapped tags.
    """
    _lib.require_cuda(points, semantics)
    if grid is None and range_spec is None:
        raise ValueError("nothing to do: pass grid and/or range_spec")
    if lidar_prep is not None and (grid is not None or range_spec is None or points.dtype != torch.float32):
        raise ValueError("lidar_prep belongs to range-only calls on float32 points (muvo/data/dataset.py:275-300)")
    lib = _lib.load()
    dev = points.device
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("points must be (P, 3)")
    if points.dtype not in (torch.float32, torch.float64):
        raise TypeError("points must be float32 or float64")
    if points.dtype == torch.float64 and range_spec is not None:
        raise TypeError("range projection takes float32 points (muvo/data/dataset.py:275-300 feeds float32)")
    points = points.contiguous()
    semantics = semantics.reshape(-1).contiguous()
    if semantics.dtype != torch.uint8:
        raise TypeError("semantics must be uint8")
    P = points.shape[0]
    if semantics.shape[0] != P:
        raise ValueError("points / semantics length mismatch")
    off = _as_offsets(frame_offsets, P, dev)
    F = off.numel() - 1
    res = dict(out) if out else {}
    with torch.cuda.device(dev):
        stream = _lib.current_stream(dev)
        g_c = grid.to_c() if grid is not None else None
        r_c = range_spec.to_c() if range_spec is not None else None
        nbytes = C.c_size_t(0)
        _lib.check(lib.muvo_points_workspace_bytes(P, F, C.byref(g_c) if g_c else None, C.byref(r_c) if r_c else None,
                                                   C.byref(nbytes)), "muvo_points_workspace_bytes")
        sig = (F, tuple(int(v) for v in grid.voxel_size) if grid is not None else None,
               (int(range_spec.H), int(range_spec.W)) if range_spec is not None else None)
        ws = _ws.get(nbytes.value, dev, stream, sig)
        diag = torch.zeros(_lib.DIAG_COUNT, dtype=torch.int64, device=dev) if want_diag else None
        dense_t = sparse_t = nocc_t = depth_t = xyz_t = sem_t = start_t = None
        if grid is not None:
            dx, dy, dz = (int(v) for v in grid.voxel_size)
            if dense:
                dense_t = res.get("voxel")
                if dense_t is None:
                    dense_t = torch.empty((F, dx, dy, dz), dtype=torch.uint8, device=dev)
            if sparse:
                sparse_t = res.get("voxel_sparse")
                if sparse_t is None:
                    sparse_t = torch.empty((max(P, 1), 4), dtype=torch.int16, device=dev)   # viewed as uint16 by the caller
            nocc_t = res.get("n_occ")
            if nocc_t is None:
                nocc_t = torch.empty((F,), dtype=torch.int64, device=dev)
            if sparse and packed_sparse:
                start_t = res.get("sparse_start")
                if start_t is None:
                    start_t = torch.empty((F + 1,), dtype=torch.int64, device=dev)
            if remap is not None:
                remap = remap.to(device=dev, dtype=torch.uint8).contiguous()
                if remap.numel() != 256:
                    raise ValueError("remap must have 256 entries")
        if range_spec is not None:
            H, W = int(range_spec.H), int(range_spec.W)
            sem_t = res.get("range_sem")
            if sem_t is None:
                sem_t = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
            if layout == "xyzd":
                xyz_t = res.get("range_xyzd")
                if xyz_t is None:
                    xyz_t = torch.empty((F, 4, H, W), dtype=torch.float32, device=dev)
                lay = _lib.RANGE_LAYOUT_XYZD
            elif layout == "hwc":
                xyz_t = res.get("range_xyz")
                if xyz_t is None:
                    xyz_t = torch.empty((F, H, W, 3), dtype=torch.float32, device=dev)
                depth_t = res.get("range_depth")
                if depth_t is None:
                    depth_t = torch.empty((F, H, W), dtype=torch.float32, device=dev)
                lay = _lib.RANGE_LAYOUT_HWC
            else:
                raise ValueError("layout must be 'xyzd' or 'hwc'")
        p = _lib.ptr
        if grid is not None and range_spec is not None:
            rc = lib.muvo_points_fused(p(points), p(semantics), p(off), F, P, C.byref(g_c), p(remap), C.byref(r_c), lay,
                                       p(dense_t), p(sparse_t), p(nocc_t), p(start_t), p(depth_t), p(xyz_t), p(sem_t), p(diag),
                                       ws.data_ptr(), ws.numel(), stream)
        elif grid is not None:
            dt = _lib.F32 if points.dtype == torch.float32 else _lib.F64
            rc = lib.muvo_voxelize(p(points), dt, p(semantics), p(off), F, P, C.byref(g_c), p(remap), p(dense_t),
                                   p(sparse_t), p(nocc_t), p(start_t), p(diag), ws.data_ptr(), ws.numel(), stream)
        elif lidar_prep is not None:
            prep_c, _keep = lidar_prep.to_c(dev)
            rc = lib.muvo_range_project_lidar(p(points), p(semantics), p(off), F, P, C.byref(r_c), C.byref(prep_c), lay, p(depth_t),
                                              p(xyz_t), p(sem_t), p(diag), ws.data_ptr(), ws.numel(), stream)
        else:
            rc = lib.muvo_range_project(p(points), p(semantics), p(off), F, P, C.byref(r_c), lay, p(depth_t), p(xyz_t),
                                        p(sem_t), p(diag), ws.data_ptr(), ws.numel(), stream)
        if rc != 0:
            _ws.invalidate(dev, stream)
        _lib.check(rc, "muvo points kernels")
    if dense_t is not None:
        res["voxel"] = dense_t
    if sparse_t is not None:
        res["voxel_sparse"] = sparse_t
    if nocc_t is not None:
        res["n_occ"] = nocc_t
    if start_t is not None:
        res["sparse_start"] = start_t
    if range_spec is not None:
        res["range_sem"] = sem_t
        if layout == "xyzd":
            res["range_xyzd"] = xyz_t
        else:
            res["range_xyz"], res["range_depth"] = xyz_t, depth_t
    if diag is not None:
        res["diag"] = diag
    res["frame_offsets"] = off
    return res


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MuvoError("muvo_b200 kernels need a CUDA device (sm_100a); no CPU fallback exists")
    return torch.device("cuda", torch.cuda.current_device())


def voxel_filter(pcd, sem, voxel_resolution, voxel_size, offset):
    """Drop-in for ``voxel_filter`` (data/data_preprocessing.py:172-228): NumPy in, NumPy out.

    Returns ``(voxels uint16 (n,3), semantics uint8 (n,))`` ordered by ``x + y*Dx + z*Dx*Dy``.
    Unlike the reference, a caller-supplied ``offset`` ndarray is not modified in place.
    """
    pcd = np.asarray(pcd)
    if pcd.ndim != 2 or pcd.shape[1] != 3:
        raise ValueError("pcd must be (N, 3)")
    if pcd.dtype != np.float32:
        pcd = pcd.astype(np.float64, copy=False)     # `pcd + offset` promotes to float64 in the reference (:177)
    sem = np.asarray(sem).reshape(pcd.shape[0], -1)[:, 0] if pcd.shape[0] else np.zeros((0,), np.uint8)
    sem = sem.astype(np.uint8, copy=False)
    dev = _default_device()
    spec = GridSpec(voxel_resolution, tuple(int(v) for v in np.asarray(voxel_size)), tuple(np.asarray(offset, dtype=np.float64).tolist()))
    pts_t = torch.from_numpy(np.ascontiguousarray(pcd)).to(dev)
    sem_t = torch.from_numpy(np.ascontiguousarray(sem)).to(dev)
    r = sensor_to_grid(pts_t, sem_t, None, grid=spec, dense=False, sparse=True)
    # ONE synchronisation for the read-back: the count and an optimistic number of rows travel together into pinned memory
    # (a CARLA frame occupies 10-30 k voxels); only a frame with more occupied voxels than that needs a second copy
    P = int(pts_t.shape[0])
    cap = min(P, 1 << 16)
    h = _pinned_rows(cap)
    h_n, h_rows = h[:8].view(torch.int64), h[8:8 + cap * 8].view(torch.int16).view(cap, 4)
    h_n.copy_(r["n_occ"][:1], non_blocking=True)
    if cap:
        h_rows.copy_(r["voxel_sparse"][:cap], non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    n = int(h_n[0])
    rows = h_rows[:min(n, cap)].numpy().view(np.uint16)
    if n > cap:
        rows = np.concatenate([rows, r["voxel_sparse"][cap:n].cpu().numpy().view(np.uint16)], 0)
    return np.ascontiguousarray(rows[:, :3]), rows[:, 3].astype(np.uint8)


_pinned_cache: dict = {}


def _pinned_rows(cap: int) -> torch.Tensor:
    """Pinned staging for voxel_filter's read-back (8 bytes of count + cap rows of 8 bytes), one buffer per thread and size class."""
    import threading
    size = 1 << max(12, int(8 + cap * 8 - 1).bit_length())
    key = (threading.get_ident(), size)
    buf = _pinned_cache.get(key)
    if buf is None:
        buf = torch.empty(size, dtype=torch.uint8).pin_memory()
        _pinned_cache[key] = buf
    return buf


EGO_VEHICLE_DIMENSION = [4.902, 2.128, 1.511]          # data/data_preprocessing.py:5


def label_pyramids(range_xyzd=None, range_sem=None, voxel=None, scale: float = 50.0) -> dict:
    """The range-view / voxel label pyramids of ``PreProcess.forward`` (muvo/models/preprocess.py:151-186) for a batch
    of frames: ``range_xyzd (F,4,H,W) f32`` -> ``range_view_label_1`` (= ``/ scale``), ``_2``, ``_4``; ``range_sem (F,H,W)
    u8`` -> ``range_view_seg_label_2/_4``; ``voxel (F,X,Y,Z) u8`` -> ``voxel_label_2/_4``.  Device tensors in / out."""
    lib = _lib.load()
    ref = range_xyzd if range_xyzd is not None else voxel
    if ref is None:
        raise ValueError("nothing to do")
    _lib.require_cuda(ref)
    dev = ref.device
    out = {}
    F = int(ref.shape[0])
    H = W = X = Y = Z = 0
    rv1 = rv2 = rv4 = s2 = s4 = v2 = v4 = None
    if range_xyzd is not None:
        range_xyzd = range_xyzd.contiguous().float()
        _, four, H, W = range_xyzd.shape
        if four != 4:
            raise ValueError("range_xyzd must be (F, 4, H, W)")
        rv1 = torch.empty_like(range_xyzd)
        rv2 = torch.empty((F, 4, H // 2, W // 2), dtype=torch.float32, device=dev)
        rv4 = torch.empty((F, 4, H // 4, W // 4), dtype=torch.float32, device=dev)
        out.update(range_view_label_1=rv1, range_view_label_2=rv2, range_view_label_4=rv4)
        if range_sem is not None:
            range_sem = range_sem.contiguous()
            s2 = torch.empty((F, H // 2, W // 2), dtype=torch.uint8, device=dev)
            s4 = torch.empty((F, H // 4, W // 4), dtype=torch.uint8, device=dev)
            out.update(range_view_seg_label_1=range_sem, range_view_seg_label_2=s2, range_view_seg_label_4=s4)
    if voxel is not None:
        voxel = voxel.contiguous()
        _, X, Y, Z = voxel.shape
        v2 = torch.empty((F, X // 2, Y // 2, Z // 2), dtype=torch.uint8, device=dev)
        v4 = torch.empty((F, X // 4, Y // 4, Z // 4), dtype=torch.uint8, device=dev)
        out.update(voxel_label_1=voxel, voxel_label_2=v2, voxel_label_4=v4)
    p = _lib.ptr
    with torch.cuda.device(dev):
        rc = lib.muvo_label_pyramids(p(range_xyzd), p(range_sem) if range_xyzd is not None else None, p(voxel), F, H, W, X, Y, Z,
                                     float(scale), p(rv1), p(rv2), p(rv4), p(s2), p(s4), p(v2), p(v4), _lib.current_stream(dev))
    _lib.check(rc, "muvo_label_pyramids")
    return out


def merge_pcd_device(img, lidar_xyz, lidar_sem, camera_pos, lidar_pos, fov=110, mask_ego=True, device=None):
    """Camera + LiDAR cloud of ``merge_pcd`` (data/data_preprocessing.py:125-139) built on the GPU.

    ``img`` = ``cv2.imread(depth_file, -1)`` (uint8 ``(H,W,4)``: encoded depth in B,G,R + semantic tag), ``lidar_xyz``
    float32 ``(N,3)`` in the LiDAR frame with ``lidar_sem`` uint8.  Returns DEVICE tensors ``(xyz float64 (n,3),
    sem uint8 (n,))`` in the reference's order (one host sync for ``n``, the data-dependent length)."""
    if not torch.cuda.is_available():
        raise _lib.MuvoError("muvo_b200 kernels need a CUDA device (sm_100a); no CPU fallback exists")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    img_t = torch.as_tensor(np.ascontiguousarray(img, dtype=np.uint8)) if not isinstance(img, torch.Tensor) else img
    if img_t.dim() != 3 or img_t.shape[2] != 4 or img_t.dtype != torch.uint8:
        raise ValueError("img must be uint8 (H, W, 4) as cv2.imread(file, -1) returns it")
    H, W = int(img_t.shape[0]), int(img_t.shape[1])
    img_t = img_t.to(dev).contiguous()
    lx = torch.as_tensor(np.ascontiguousarray(lidar_xyz, dtype=np.float32)) if not isinstance(lidar_xyz, torch.Tensor) else lidar_xyz
    ls = torch.as_tensor(np.ascontiguousarray(np.asarray(lidar_sem).reshape(-1), dtype=np.uint8)) if not isinstance(lidar_sem, torch.Tensor) else lidar_sem
    lx, ls = lx.to(dev, torch.float32).contiguous(), ls.to(dev, torch.uint8).reshape(-1).contiguous()
    N = int(lx.shape[0])
    focal = float(W / (2.0 * np.tan(fov * np.pi / 360.0)))                       # :89, evaluated by numpy as the reference does
    cam = (C.c_double * 3)(*[float(np.float32(v)) for v in camera_pos])          # the 4x4 matrix is np.float32 (:111)
    lid = (C.c_double * 3)(*[float(v) for v in lidar_pos])
    box = None
    if mask_ego:
        x, y, z = EGO_VEHICLE_DIMENSION
        box = (C.c_double * 6)(-x / 2, -y / 2, 0.0, x / 2, y / 2, z)
    n_max = H * W + N
    xyz = torch.empty((max(n_max, 1), 3), dtype=torch.float64, device=dev)
    sem = torch.empty((max(n_max, 1),), dtype=torch.uint8, device=dev)
    n_out = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        nb = C.c_size_t(0)
        _lib.check(lib.muvo_merge_pcd_workspace_bytes(H, W, N, C.byref(nb)), "muvo_merge_pcd_workspace_bytes")
        ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        rc = lib.muvo_merge_pcd(_lib.ptr(img_t), H, W, focal, 100.0, cam, _lib.ptr(lx), _lib.ptr(ls), N, lid, box, xyz.data_ptr(),
                                sem.data_ptr(), n_out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.current_stream(dev))
    _lib.check(rc, "muvo_merge_pcd")
    n = int(n_out.item())
    return xyz[:n], sem[:n]


def merge_pcd_batch(frames, camera_pos, lidar_pos, fov=110, mask_ego=True, device=None):
    """``merge_pcd`` for a list of decoded frames ``[(img, lidar_xyz, lidar_sem), ...]`` packed back to back on the device
    with NO host synchronisation in between (``muvo_merge_pcd_at`` chains the row offsets on the device).  Returns
    ``(xyz float64 (cap,3), sem uint8 (cap,), row_offsets int64 (N+1,))`` device tensors; ``row_offsets`` is the
    ``frame_offsets`` of :func:`sensor_to_grid` (read it once to learn the total)."""
    if not torch.cuda.is_available():
        raise _lib.MuvoError("muvo_b200 kernels need a CUDA device (sm_100a); no CPU fallback exists")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    N = len(frames)
    cam = (C.c_double * 3)(*[float(np.float32(v)) for v in camera_pos])
    lid = (C.c_double * 3)(*[float(v) for v in lidar_pos])
    box = None
    if mask_ego:
        x, y, z = EGO_VEHICLE_DIMENSION
        box = (C.c_double * 6)(-x / 2, -y / 2, 0.0, x / 2, y / 2, z)
    staged, cap, ws_need = [], 0, 0
    for img, lxyz, lsem in frames:
        img_t = torch.as_tensor(np.ascontiguousarray(img, dtype=np.uint8))
        if img_t.dim() != 3 or img_t.shape[2] != 4:
            raise ValueError("img must be uint8 (H, W, 4) as cv2.imread(file, -1) returns it")
        lx = torch.as_tensor(np.ascontiguousarray(lxyz, dtype=np.float32))
        ls = torch.as_tensor(np.ascontiguousarray(np.asarray(lsem).reshape(-1), dtype=np.uint8))
        H, W, n = int(img_t.shape[0]), int(img_t.shape[1]), int(lx.shape[0])
        nb = C.c_size_t(0)
        _lib.check(lib.muvo_merge_pcd_workspace_bytes(H, W, n, C.byref(nb)), "muvo_merge_pcd_workspace_bytes")
        ws_need = max(ws_need, nb.value)
        cap += H * W + n
        staged.append((img_t.to(dev, non_blocking=True), lx.to(dev, non_blocking=True), ls.to(dev, non_blocking=True), H, W, n))
    xyz = torch.empty((max(cap, 1), 3), dtype=torch.float64, device=dev)
    sem = torch.empty((max(cap, 1),), dtype=torch.uint8, device=dev)
    offs = torch.zeros((N + 1,), dtype=torch.int64, device=dev)
    ws = torch.empty(max(ws_need, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = _lib.current_stream(dev)
        for f, (img_t, lx, ls, H, W, n) in enumerate(staged):
            focal = float(W / (2.0 * np.tan(fov * np.pi / 360.0)))
            rc = lib.muvo_merge_pcd_at(_lib.ptr(img_t), H, W, focal, 100.0, cam, _lib.ptr(lx), _lib.ptr(ls), n, lid, box, xyz.data_ptr(),
                                       sem.data_ptr(), cap, offs.data_ptr(), f, ws.data_ptr(), ws.numel(), stream)
            _lib.check(rc, "muvo_merge_pcd_at")
    return xyz, sem, offs


def merge_pcd_arrays(img, lidar_xyz, lidar_sem, camera_pos, lidar_pos, fov=110, mask_ego=True):
    """``merge_pcd`` on arrays, NumPy in / NumPy out: ``(pcd float64 (n,3), semantic uint8 (n,1))``."""
    xyz, sem = merge_pcd_device(img, lidar_xyz, lidar_sem, camera_pos, lidar_pos, fov, mask_ego)
    return xyz.cpu().numpy(), sem.cpu().numpy()[:, None]


def merge_pcd(depth_file, lidar_file, camera_pos, lidar_pos, fov=110, mask_ego=True):
    """Drop-in for ``merge_pcd`` (data/data_preprocessing.py:125-139): same arguments (file names) and returns."""
    import cv2
    img = cv2.imread(depth_file, -1)
    data = np.load(lidar_file, allow_pickle=True).item()                         # load_lidar, :80-84
    return merge_pcd_arrays(img, data['points_xyz'], data['ObjTag'], camera_pos, lidar_pos, fov, mask_ego)


def voxelize_one(depth_file, lidar_file, cfg, save_name, pipe=None):
    """Drop-in for ``voxelize_one`` (data/generate_voxels.py:64-77): merge on the GPU, voxelise the float64 cloud on the
    GPU without a round trip through the host, save the ``(n,4) uint16`` array."""
    import cv2
    img = cv2.imread(depth_file, -1)
    data = np.load(lidar_file, allow_pickle=True).item()
    xyz, sem = merge_pcd_device(img, data['points_xyz'], data['ObjTag'], cfg.camera_position, cfg.lidar_position, cfg.fov)
    offset_x = cfg.bev_offset_forward * cfg.bev_resolution
    offset_z = cfg.offset_z * cfg.voxel_resolution
    spec = GridSpec(cfg.voxel_resolution, tuple(cfg.voxel_size), (offset_x, 0, offset_z))
    r = sensor_to_grid(xyz, sem, None, grid=spec, dense=False, sparse=True)
    n = int(r["n_occ"][0].item())
    out = r["voxel_sparse"][:n].cpu().numpy().view(np.uint16)
    np.save(f'{save_name}', out)
    if pipe is not None:
        pipe.send(['x'])
    return out


def densify_voxels(voxel_data, voxel_size, remap=None, frame_offsets=None, device=None):
    """Saved sparse voxels -> dense grid, ``muvo/data/dataset.py:317-327``: ``voxel_data (n,4)`` uint16 / int16 rows
    ``[x, y, z, label]`` as ``voxelize_one`` saves them (NumPy array or device tensor; several files concatenated with
    ``frame_offsets [F+1]``), label 255 -> 0 then ``remap[label]`` (dataset.py:322-323), written with the reference's
    "last row wins" rule.  Returns a DEVICE tensor ``(F, Dx, Dy, Dz) uint8`` (``[0][None]`` is dataset.py:327's array)."""
    lib = _lib.load()
    dev = torch.device(device) if device is not None else (voxel_data.device if isinstance(voxel_data, torch.Tensor) and voxel_data.is_cuda
                                                           else _default_device())
    if isinstance(voxel_data, torch.Tensor):
        vd = voxel_data.to(dev)
        vd = vd.view(torch.int16) if vd.dtype in (torch.int16, torch.uint16) else vd.to(torch.int16)
    else:
        vd = torch.from_numpy(np.ascontiguousarray(np.asarray(voxel_data).astype(np.uint16, copy=False)).view(np.int16)).to(dev)
    if vd.dim() != 2 or vd.shape[1] != 4:
        raise ValueError("voxel_data must be (n, 4) rows [x, y, z, label]")
    vd = vd.contiguous()
    n = int(vd.shape[0])
    off = _as_offsets(frame_offsets, n, dev)
    F = off.numel() - 1
    dx, dy, dz = (int(v) for v in voxel_size)
    tab = None
    if remap is not None:
        t = np.asarray(remap, dtype=np.uint8).reshape(-1)
        if t.size < 256:
            t = np.concatenate([t, np.full(256 - t.size, t.max() if t.size else 0, np.uint8)])
        tab = torch.from_numpy(np.ascontiguousarray(t[:256])).to(dev)
    out = torch.empty((F, dx, dy, dz), dtype=torch.uint8, device=dev)
    scratch = torch.empty((max(n, 1),), dtype=torch.uint8, device=dev)
    n_bad = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.muvo_densify_sparse(_lib.ptr(vd), _lib.ptr(off), F, n, dx, dy, dz, _lib.ptr(tab), _lib.ptr(out), _lib.ptr(scratch),
                                     _lib.ptr(n_bad), _lib.current_stream(dev))
    _lib.check(rc, "muvo_densify_sparse")
    out.n_bad = n_bad                    # rows outside the grid (numpy: IndexError); a device counter, read it if you care
    return out


def lidar_range_view(points_xyz, obj_tag, pc=None, lidar_position=(1.0, 0.0, 2.0), remap=None, frame_offsets=None,
                     layout: str = "xyzd", device=None):
    """``muvo/data/dataset.py:275-305`` for one or more raw semantic-LiDAR sweeps in ONE pass of the point kernels:
    ``convert_coor_lidar`` + ``remap[ObjTag]`` + ego-box drop + ``do_range_projection`` (+ the ``(4,H,W)`` packing of
    :301-303 with ``layout="xyzd"``).  ``points_xyz (P,3)`` float32 in the LiDAR frame and ``obj_tag (P,)`` uint8, NumPy or
    device tensors (the inputs are not modified, unlike ``convert_coor_lidar``).  Returns :func:`sensor_to_grid`'s dict
    of device tensors."""
    dev = torch.device(device) if device is not None else (points_xyz.device if isinstance(points_xyz, torch.Tensor) and points_xyz.is_cuda
                                                           else _default_device())
    pts = points_xyz if isinstance(points_xyz, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points_xyz, dtype=np.float32))
    tag = obj_tag if isinstance(obj_tag, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(obj_tag).reshape(-1), dtype=np.uint8))
    pts, tag = pts.to(dev, torch.float32), tag.to(dev, torch.uint8).reshape(-1)
    if pc is None:
        pc = PointCloud(lidar_position=lidar_position)
    spec = pc.spec if isinstance(pc, PointCloud) else _RadianRangeSpec(pc.H, pc.W, float(pc.fov_down), float(pc.fov),
                                                                      tuple(float(v) for v in np.asarray(pc.lidar_position, dtype=np.float64).reshape(-1)))
    return sensor_to_grid(pts, tag, frame_offsets, range_spec=spec, layout=layout, want_diag=True,
                          lidar_prep=LidarPrep(lidar_position=tuple(float(v) for v in lidar_position), remap=remap))


def voxelize_one_array(pcd, sem, voxel_resolution, voxel_size, offset):
    """``(n,4) uint16 [x,y,z,label]`` as saved by ``voxelize_one`` (data/generate_voxels.py:64-73)."""
    vox, lab = voxel_filter(pcd, sem, voxel_resolution, voxel_size, offset)
    return np.concatenate([vox, lab[:, None]], axis=1)


@dataclass(frozen=True)
class _RadianRangeSpec:
    """Range description taken verbatim (radians) from a PointCloud-like object, so that a patched reference
    instance and :class:`PointCloud` feed bit-identical constants to the kernel."""
    H: int
    W: int
    fov_down_rad: float
    fov_rad: float
    lidar_position: tuple

    def to_c(self) -> _lib.MuvoRangeCfg:
        c = _lib.MuvoRangeCfg()
        c.H, c.W = int(self.H), int(self.W)
        c.fov_down_abs = float(abs(self.fov_down_rad))
        c.fov = float(self.fov_rad)
        for k in range(3):
            c.lidar_pos[k] = float(self.lidar_position[k])
        return c


def do_range_projection(pc, points, semantics):
    """``PointCloud.do_range_projection`` for any object exposing H, W, fov_down, fov, lidar_position
    (the reference class or ours): geometry_utils.py:175-220, NumPy in / NumPy out."""
    points = np.asarray(points)
    if points.ndim != 2 or points.shape[1] != 3:
        raise ValueError("points must be (N, 3)")
    if points.dtype != np.float32:
        raise TypeError("do_range_projection expects float32 points (as produced by muvo/data/dataset.py:275-290)")
    semantics = np.asarray(semantics).reshape(-1).astype(np.uint8, copy=False)
    spec = _RadianRangeSpec(pc.H, pc.W, float(pc.fov_down), float(pc.fov),
                            tuple(float(v) for v in np.asarray(pc.lidar_position, dtype=np.float64).reshape(-1)))
    dev = _default_device()
    pts_t = torch.from_numpy(np.ascontiguousarray(points)).to(dev)
    sem_t = torch.from_numpy(np.ascontiguousarray(semantics)).to(dev)
    r = sensor_to_grid(pts_t, sem_t, None, range_spec=spec, layout="hwc", want_diag=True)
    diag = r["diag"].cpu().numpy()
    if diag[_lib.DIAG_DROPPED_NONFINITE] > 0:
        # the reference hits `IndexError` here: NaN -> int32 min index (geometry_utils.py:187-217)
        raise IndexError("point(s) at the sensor origin or with non-finite coordinates cannot be projected")
    try:
        pc.last_diag = diag
    except Exception:
        pass
    return (r["range_depth"][0].cpu().numpy(), r["range_xyz"][0].cpu().numpy(), r["range_sem"][0].cpu().numpy())


class PointCloud(object):
    """Drop-in for ``muvo.utils.geometry_utils.PointCloud`` (geometry_utils.py:166-244)."""

    def __init__(self, H=64, W=1024, fov_down=-30, fov_up=10, lidar_position=(1, 0, 2)):
        self.fov_up = fov_up / 180.0 * np.pi  # in rad
        self.fov_down = fov_down / 180.0 * np.pi
        self.fov = self.fov_up - self.fov_down
        self.H = H
        self.W = W
        self.lidar_position = np.asarray(lidar_position)
        self._spec = _RadianRangeSpec(H, W, float(self.fov_down), float(self.fov),
                                      tuple(float(v) for v in np.asarray(lidar_position, dtype=np.float64).reshape(-1)))

    @property
    def spec(self):
        return self._spec

    def do_range_projection(self, points, semantics):
        """NumPy in / NumPy out: ``(range_depth (H,W) f32, range_xyz (H,W,3) f32, range_sem (H,W) u8)``."""
        return do_range_projection(self, points, semantics)

    def project_batch(self, points: torch.Tensor, semantics: torch.Tensor, frame_offsets=None, layout: str = "xyzd"):
        """Device-side batched projection (no host round trip); see :func:`sensor_to_grid`."""
        return sensor_to_grid(points, semantics, frame_offsets, range_spec=self._spec, layout=layout)

    def restore_pcd_coor(self, range_depth):
        """Inverse projection (geometry_utils.py:223-244); plain NumPy, not on the hot path."""
        rows = np.arange(self.H, dtype=float)[:, None] / self.H
        cols = np.arange(self.W, dtype=float)[None, :] / self.W
        pitch = (1.0 - rows) * self.fov - abs(self.fov_down)
        yaw = (1.0 - cols / 0.5) * np.pi
        pitch = np.broadcast_to(pitch, (self.H, self.W))[None, None]
        yaw = np.broadcast_to(yaw, (self.H, self.W))[None, None]
        depth = range_depth
        z = depth * np.sin(pitch)
        planar = depth * np.cos(pitch)
        x = planar * np.cos(yaw)
        y = planar * np.sin(yaw)
        pts = np.stack([x, -y, z], axis=-1)
        pts = pts + self.lidar_position.reshape((1, 1, 1, 1, -1))
        pts = pts * np.array([1, -1, 1]).reshape((1, 1, 1, 1, -1))
        return np.concatenate([pts, depth[..., None]], axis=-1)

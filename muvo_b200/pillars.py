""""Next" row N4, last part: the scatter reductions of the PointPillar encoder on the GPU.

Drop-ins for ``torch_scatter.scatter_mean(src, index, dim=0)`` and ``torch_scatter.scatter_max(src, index, dim=0)``
as ``muvo/models/common.py`` uses them (``DynamicPointNet.forward`` :703, ``PointPillarNet.decorate`` :731), plus
``PointPillarNet``'s geometric helpers (``grid_locations`` :735-745, ``decorate`` :721-733, ``scatter_points`` :756-761)
as functions.  ``torch_scatter`` is a third-party dependency of the reference that is neither vendored nor installed in
this image, so parity is against its documented semantics (restated in ``oracle/``) and ``torch.scatter_reduce``.
Only ``dim=0`` on 2-D floating ``src`` is supported -- what the reference calls; float16 / bfloat16 inputs (the
reference's default ``PRECISION='16-mixed'`` feeds ``scatter_max`` half tensors, common.py:702-703) are reduced in
float32 and returned, with their gradients, in the input dtype.  Both reductions are deterministic.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

_IDX = {torch.int64: _lib.I64, torch.int32: _lib.I32}
_flags: dict = {}


def _flag(dev: torch.device) -> torch.Tensor:
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _flags:
        _flags[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return _flags[key]


def _prep(src: torch.Tensor, index: torch.Tensor, dim: int, dim_size: Optional[int]):
    _lib.require_cuda(src, index)
    if dim not in (0, -src.dim()):
        raise NotImplementedError("muvo_b200 scatter ops support dim=0 only (the PointPillar call sites)")
    if src.dim() != 2 or src.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError("src must be a 2-D float32 / float16 / bfloat16 tensor")
    if index.dtype not in _IDX:
        raise TypeError("index must be int64 or int32")
    if index.dim() == 2:                       # torch_scatter broadcasts an index of src's shape; the reference passes 1-D
        index = index[:, 0]
    if index.dim() != 1 or index.shape[0] != src.shape[0]:
        raise ValueError("index must have one entry per source row")
    M = int(dim_size) if dim_size is not None else (int(index.max().item()) + 1 if index.numel() else 0)
    return src.contiguous(), index.contiguous(), M


def _workspace(lib, M: int, F: int, dev) -> torch.Tensor:
    need = C.c_size_t(0)
    _lib.check(lib.muvo_pillar_workspace_bytes(M, F, C.byref(need)), "muvo_pillar_workspace_bytes")
    return torch.empty(int(need.value), dtype=torch.uint8, device=dev)


def check_indices(device=None) -> None:
    """Raises ``IndexError`` if any scatter call since the last check saw an out-of-range index (synchronises)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    f = _flag(dev)
    if int(f.item()):
        f.zero_()
        raise IndexError("scatter index out of range")


class _ScatterMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, M):
        N, F = int(src.shape[0]), int(src.shape[1])
        dev = src.device
        in_dtype = src.dtype
        src = src.float()                                          # the reduction runs in float32 (as under autocast)
        out = torch.empty((M, F), dtype=torch.float32, device=dev)
        count = torch.empty((M,), dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            ws = _workspace(lib, M, F, dev)
            rc = lib.muvo_pillar_scatter_mean(_lib.ptr(src), _lib.ptr(index), _IDX[index.dtype], N, F, M, _lib.ptr(out),
                                              _lib.ptr(count), ws.data_ptr(), ws.numel(), _flag(dev).data_ptr(),
                                              _lib.current_stream(dev))
        _lib.check(rc, "muvo_pillar_scatter_mean")
        ctx.save_for_backward(index, count)
        ctx.shape = (N, F, M, in_dtype)
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, g):
        index, count = ctx.saved_tensors
        N, F, M, in_dtype = ctx.shape
        dev = g.device
        gs = torch.empty((N, F), dtype=torch.float32, device=dev)
        g = g.contiguous().float()
        with torch.cuda.device(dev):
            rc = _lib.load().muvo_pillar_scatter_mean_bwd(_lib.ptr(g), _lib.ptr(index), _IDX[index.dtype], _lib.ptr(count), N, F, M,
                                                          _lib.ptr(gs), _lib.current_stream(dev))
        _lib.check(rc, "muvo_pillar_scatter_mean_bwd")
        return gs.to(in_dtype), None, None


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, M):
        N, F = int(src.shape[0]), int(src.shape[1])
        dev = src.device
        in_dtype = src.dtype
        src = src.float()                                          # exact for float16 / bfloat16 values: the max is unchanged
        out = torch.empty((M, F), dtype=torch.float32, device=dev)
        arg = torch.empty((M, F), dtype=torch.int64, device=dev)
        lib = _lib.load()
        ws = _workspace(lib, M, F, dev)
        with torch.cuda.device(dev):
            rc = lib.muvo_pillar_scatter_max(_lib.ptr(src), _lib.ptr(index), _IDX[index.dtype], N, F, M, _lib.ptr(out), _lib.ptr(arg),
                                             ws.data_ptr(), ws.numel(), _flag(dev).data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "muvo_pillar_scatter_max")
        ctx.save_for_backward(index, arg)
        ctx.shape = (N, F, M, in_dtype)
        ctx.mark_non_differentiable(arg)
        return out.to(in_dtype), arg

    @staticmethod
    def backward(ctx, g, _g_arg):
        index, arg = ctx.saved_tensors
        N, F, M, in_dtype = ctx.shape
        dev = g.device
        gs = torch.empty((N, F), dtype=torch.float32, device=dev)
        g = g.contiguous().float()
        with torch.cuda.device(dev):
            rc = _lib.load().muvo_pillar_scatter_max_bwd(_lib.ptr(g), _lib.ptr(index), _IDX[index.dtype], _lib.ptr(arg), N, F, M,
                                                         _lib.ptr(gs), _lib.current_stream(dev))
        _lib.check(rc, "muvo_pillar_scatter_max_bwd")
        return gs.to(in_dtype), None, None


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim: int = 0, out=None, dim_size: Optional[int] = None) -> torch.Tensor:
    """``torch_scatter.scatter_mean(src, index, dim=0)`` (common.py:731): ``[M, F]``, empty rows 0."""
    if out is not None:
        raise NotImplementedError("out= is not supported")
    s, i, M = _prep(src, index, dim, dim_size)
    return _ScatterMean.apply(s, i, M)


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = 0, out=None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``torch_scatter.scatter_max(src, index, dim=0)`` (common.py:703): ``(max [M, F], argmax [M, F] int64)``; empty rows
    are 0 with argmax ``N``; ties resolve to the lowest source row."""
    if out is not None:
        raise NotImplementedError("out= is not supported")
    s, i, M = _prep(src, index, dim, dim_size)
    return _ScatterMax.apply(s, i, M)


# ---- PointPillarNet's geometric helpers (muvo/models/common.py:708-761) as functions of its constructor arguments
def pillar_grid_locations(points: torch.Tensor, min_x=-10, max_x=70, min_y=-40, max_y=40, pixels_per_meter=4):
    """``PointPillarNet.grid_locations`` (:735-745): in-range points and their ``(x, y)`` pillar coordinates (int64)."""
    keep = (points[:, 0] >= min_x) & (points[:, 0] < max_x) & (points[:, 1] >= min_y) & (points[:, 1] < max_y)
    points = points[keep, :]
    coords = (points[:, [0, 1]] - torch.tensor([min_x, min_y], device=points.device)) * pixels_per_meter
    return points, coords.long()


def pillar_decorate(points: torch.Tensor, unique_coords: torch.Tensor, inverse_indices: torch.Tensor, min_x=-10, min_y=-40,
                    pixels_per_meter=4) -> torch.Tensor:
    """``PointPillarNet.decorate`` (:721-733) with the cluster mean on the GPU kernel."""
    dtype = points.dtype
    x_centers = unique_coords[inverse_indices][:, 2:3].to(dtype) / pixels_per_meter + min_x
    y_centers = unique_coords[inverse_indices][:, 1:2].to(dtype) / pixels_per_meter + min_y
    xyz = points[:, :3]
    points_cluster = xyz - scatter_mean(xyz, inverse_indices, dim=0, dim_size=int(unique_coords.shape[0]))[inverse_indices]
    return torch.cat([points, points_cluster, xyz[:, :1] - x_centers, xyz[:, 1:2] - y_centers], dim=-1)


def pillar_scatter_points(features: torch.Tensor, coords: torch.Tensor, batch_size: int, ny: int, nx: int) -> torch.Tensor:
    """``PointPillarNet.scatter_points`` (:756-761): pillar features onto the ``(B, F, ny, nx)`` canvas (y flipped)."""
    canvas = torch.zeros(batch_size, features.shape[1], ny, nx, dtype=features.dtype, device=features.device)
    canvas[coords[:, 0], :, torch.clamp(ny - 1 - coords[:, 1], 0, ny - 1), torch.clamp(coords[:, 2], 0, nx - 1)] = features
    return canvas

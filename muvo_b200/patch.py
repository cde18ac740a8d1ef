"""Monkey-patch a MUVO checkout in place: ``muvo_b200.patch()`` swaps the four hot-path entry
points for the B200 kernels; ``unpatch()`` restores them.  The reference
tree is never edited on disk.
"""
from __future__ import annotations

import sys

_saved: list = []


def _swap(obj, name, new):
    if hasattr(obj, name):
        _saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, new)


def patch(verbose: bool = False) -> int:
    """Patch every already-imported reference module; returns the number of symbols replaced.

    ``data/generate_voxels.py`` does ``from data_preprocessing import *``; import it AFTER calling
    ``patch()`` (or call ``patch()`` again) so that its namespace picks up the replacement too.
    """
    from . import frustum_pooling as fp, metrics, points
    n0 = len(_saved)
    m = sys.modules.get("data_preprocessing")
    if m is not None:
        _swap(m, "voxel_filter", points.voxel_filter)
        _swap(m, "merge_pcd", points.merge_pcd)                   # N1: camera + LiDAR cloud on the device
    m = sys.modules.get("generate_voxels")
    if m is not None:
        _swap(m, "voxel_filter", points.voxel_filter)
        _swap(m, "merge_pcd", points.merge_pcd)
        _swap(m, "voxelize_one", points.voxelize_one)             # merge + voxelise without a host round trip
    m = sys.modules.get("muvo.utils.geometry_utils")
    if m is not None and hasattr(m, "PointCloud"):
        _swap(m.PointCloud, "do_range_projection", points.do_range_projection)
    m = sys.modules.get("muvo.models.frustum_pooling")
    if m is not None:
        for name in ("FrustumPooling", "QuickCumsum", "cumsum_trick", "quick_cumsum"):
            _swap(m, name, getattr(fp, name))
    m = sys.modules.get("muvo.models.mile")
    if m is not None:
        _swap(m, "FrustumPooling", fp.FrustumPooling)
    m = sys.modules.get("muvo.layers.layers")
    if m is not None:
        _swap(m, "VoxelsSumming", fp.VoxelsSumming)
    m = sys.modules.get("muvo.metrics")
    if m is not None:
        _swap(m, "SSCMetrics", metrics.SSCMetrics)
    m = sys.modules.get("muvo.trainer")
    if m is not None:
        _swap(m, "SSCMetrics", metrics.SSCMetrics)
    from . import losses                                          # N4: SemScalLoss / GeoScalLoss reductions
    for mod in ("muvo.losses", "muvo.trainer"):
        m = sys.modules.get(mod)
        if m is not None:
            for name in ("SemScalLoss", "GeoScalLoss"):
                if hasattr(m, name):
                    _swap(m, name, getattr(losses, name))
    from . import pillars                                         # N4: torch_scatter calls of the PointPillar encoder
    m = sys.modules.get("muvo.models.common")
    if m is not None:
        _swap(m, "scatter_mean", pillars.scatter_mean)
        _swap(m, "scatter_max", pillars.scatter_max)
    if verbose:
        for obj, name, _ in _saved[n0:]:
            print(f"muvo_b200.patch: {getattr(obj, '__name__', obj)}.{name}")
    return len(_saved) - n0


def unpatch() -> None:
    while _saved:
        obj, name, old = _saved.pop()
        setattr(obj, name, old)

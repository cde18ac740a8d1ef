"""Monkey-patch a MUVO checkout in place: ``muvo_b200.patch()`` swaps the four hot-path entry
points for the B200 kernels; ``unpatch()`` restores them.  The reference
tree is never edited on disk.
"""
from __future__ import annotations

import sys

_saved: list = []


def _swap(obj, name, new):
    if hasattr(obj, name) and getattr(obj, name) is not new:
        _saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, new)


def _in_worker_or_bad_fork() -> bool:
    """True inside a DataLoader worker, or in a forked child of a process that had already initialised CUDA
    (where any CUDA call raises "Cannot re-initialize CUDA in forked subprocess")."""
    import torch
    try:
        if torch.utils.data.get_worker_info() is not None:
            return True
    except Exception:
        pass
    bad_fork = getattr(torch.cuda, "_is_in_bad_fork", None)
    return bool(bad_fork()) if callable(bad_fork) else False


def _range_projection_dispatch(original):
    """``PointCloud.do_range_projection`` replacement that is safe at the reference's own call site.

    The reference calls it per sample from ``Dataset.__getitem__`` (muvo/data/dataset.py:300), i.e. inside forked
    DataLoader workers (``N_WORKERS`` > 0), where CUDA cannot be used: there the reference's own NumPy method runs
    unchanged.  In the main process (``num_workers=0``, offline tools, tests) the B200 kernel runs.  The batched
    GPU path for training is ``muvo_b200.sensor_to_grid`` / ``HostPipeline`` in the main process."""
    from . import points

    def do_range_projection(self, pts, semantics):
        if _in_worker_or_bad_fork():
            return original(self, pts, semantics)
        return points.do_range_projection(self, pts, semantics)

    do_range_projection.__wrapped__ = original
    return do_range_projection


def patch(verbose: bool = False, range_projection: str = "auto") -> int:
    """Patch every already-imported reference module; returns the number of symbols replaced.

    ``data/generate_voxels.py`` does ``from data_preprocessing import *`` (``generate_voxels.py:16``), so the star-imported
    names live in its own namespace: every namespace that holds a copy (``data_preprocessing``, ``data.data_preprocessing``,
    ``generate_voxels``, ``data.generate_voxels`` and ``__main__`` when the script itself is running) is patched.
    Call ``patch()`` after the reference modules are imported (or again later: it is idempotent per symbol).

    ``range_projection``: "auto" (default) = GPU kernel in the main process, the reference's NumPy method inside
    DataLoader workers / forked children; "always" = GPU kernel everywhere (needs ``num_workers=0`` or the spawn start
    method); "never" = leave ``PointCloud.do_range_projection`` alone.
    """
    from . import frustum_pooling as fp, metrics, points
    if range_projection not in ("auto", "always", "never"):
        raise ValueError("range_projection must be 'auto', 'always' or 'never'")
    n0 = len(_saved)
    mains = [sys.modules.get("__main__")]
    for name in ("data_preprocessing", "data.data_preprocessing"):
        m = sys.modules.get(name)
        if m is not None:
            for name, new in (("voxel_filter", points.voxel_filter), ("merge_pcd", points.merge_pcd)):   # N1: merge on the device
                if hasattr(m, name) and not getattr(getattr(m, name), "__module__", "").startswith("muvo_b200"):
                    _swap(m, name, new)
    for m in [sys.modules.get("generate_voxels"), sys.modules.get("data.generate_voxels")] + mains:
        if m is None or getattr(m, "__name__", "").startswith("muvo_b200"):
            continue
        if m in mains and not all(hasattr(m, k) for k in ("voxelize_one", "voxel_filter", "merge_pcd")):
            continue                                              # __main__ only when it is the reference script itself
        for name, new in (("voxel_filter", points.voxel_filter), ("merge_pcd", points.merge_pcd),
                          ("voxelize_one", points.voxelize_one)):   # voxelize_one: merge + voxelise without a host round trip
            if hasattr(m, name) and not getattr(getattr(m, name), "__module__", "").startswith("muvo_b200"):
                _swap(m, name, new)
    m = sys.modules.get("muvo.utils.geometry_utils")
    if m is not None and hasattr(m, "PointCloud") and range_projection != "never":
        cur = m.PointCloud.do_range_projection
        if not hasattr(cur, "__wrapped__") and getattr(cur, "__module__", "") != points.__name__:
            _swap(m.PointCloud, "do_range_projection",
                  points.do_range_projection if range_projection == "always" else _range_projection_dispatch(cur))
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

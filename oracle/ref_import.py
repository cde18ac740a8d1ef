"""Import the UNMODIFIED reference functions.

TEST / BASELINE INFRASTRUCTURE ONLY -- the product (``muvo_b200/``) never imports this.
Resolution order: ``$MUVO_REFERENCE_ROOT``, ``/root/reference`` (the build container), ``baseline/_ref`` (the unmodified
copy that ``tools/install_reference.py`` / ``__graft_entry__.build()`` leave next to the repo; git-ignored, but it travels
to the GPU box, where ``/root/reference`` does not exist).  Used by ``tests/golden/make_golden.py`` (golden vectors),
``tests/test_oracle_vs_reference.py`` / ``tests/test_reference_integration.py`` (skipped when nothing resolves) and by
``bench.py``'s reference arm and ``cpu_baseline`` / stage baselines (falling back to the oracle port otherwise).

Optional third-party imports of the reference that are not installed here
(open3d, carla, chamferdist, timm, torch_scatter) are replaced by ``MagicMock``
stubs *before* import; none of them is touched by the four hot-path functions.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _resolve() -> str:
    cands = [os.environ.get("MUVO_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "muvo", "metrics.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _resolve()
_STUBS = ["open3d", "carla", "chamferdist", "timm", "timm.models", "timm.models.resnet", "torch_scatter"]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "muvo", "metrics.py"))


def load() -> types.SimpleNamespace:
    """Returns a namespace with the reference entry points of SURVEY.md section 8(a)."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "data")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import data_preprocessing  # type: ignore  # data/data_preprocessing.py
    from muvo.utils import geometry_utils  # type: ignore
    from muvo.models import frustum_pooling  # type: ignore
    from muvo.layers import layers  # type: ignore
    from muvo import metrics  # type: ignore
    from muvo import losses  # type: ignore
    return types.SimpleNamespace(
        data_preprocessing=data_preprocessing,
        geometry_utils=geometry_utils,
        frustum_pooling=frustum_pooling,
        layers=layers,
        metrics=metrics,
        voxel_filter=data_preprocessing.voxel_filter,
        convert_coor_lidar=data_preprocessing.convert_coor_lidar,
        merge_pcd=data_preprocessing.merge_pcd,
        read_img=data_preprocessing.read_img,
        PointCloud=geometry_utils.PointCloud,
        calculate_geometry=geometry_utils.calculate_geometry,
        FrustumPooling=frustum_pooling.FrustumPooling,
        QuickCumsum=frustum_pooling.QuickCumsum,
        cumsum_trick=frustum_pooling.cumsum_trick,
        VoxelsSumming=layers.VoxelsSumming,
        SSCMetrics=metrics.SSCMetrics,
        losses=losses,
        SemScalLoss=losses.SemScalLoss,
        GeoScalLoss=losses.GeoScalLoss,
    )

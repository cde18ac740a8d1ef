"""Import the UNMODIFIED reference functions (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call this.
It is used by ``tests/golden/make_golden.py`` (to freeze golden vectors) and by
``tests/test_oracle_vs_reference.py`` (skipped when the checkout is absent).

Optional third-party imports of the reference that are not installed here
(open3d, carla, chamferdist, timm, torch_scatter) are replaced by ``MagicMock``
stubs *before* import; none of them is touched by the four hot-path functions.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("MUVO_REFERENCE_ROOT", "/root/reference")
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

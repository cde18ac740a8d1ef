"""muvo_b200 -- B200 (sm_100a) kernels for MUVO's geometric sensor-to-grid hot path.

Four stages behind the reference's own Python call signatures:

(a) ``voxel_filter``                       LiDAR / merged point cloud -> occupancy voxels
(b) ``PointCloud.do_range_projection``     points -> 64x1024 range view, nearest point wins
(c) ``FrustumPooling`` / ``QuickCumsum`` / ``VoxelsSumming``   lift-splat BEV pooling, fwd + bwd
(d) ``SSCMetrics``                         occupancy IoU tp/fp/fn counts (+ NCCL all-reduce)

All arithmetic runs in ``libmuvo_b200.so`` (hand-written CUDA, C ABI in ``include/muvo_b200.h``);
there is no CPU or PyTorch fallback.  Build with ``python -m muvo_b200.build``.
"""
from ._lib import MuvoError, load as load_library  # noqa: F401
from .points import (GridSpec, RangeSpec, PointCloud, sensor_to_grid, voxel_filter, voxelize_one_array,  # noqa: F401
                     merge_pcd, merge_pcd_arrays, merge_pcd_device, voxelize_one, label_pyramids)
from .metrics import SSCMetrics, ssc_counts, ssc_counts_from_logits, all_reduce_counts  # noqa: F401
from .frustum_pooling import (FrustumPooling, QuickCumsum, VoxelsSumming, cumsum_trick, quick_cumsum, gen_dx_bx,  # noqa: F401
                              bev_pool, lift_splat, bev_params_to_intrinsics, intrinsics_inverse)
from .losses import SemScalLoss, GeoScalLoss, scal_losses, scal_sums  # noqa: F401
from .pillars import scatter_mean, scatter_max, pillar_grid_locations, pillar_decorate, pillar_scatter_points  # noqa: F401
from .patch import patch, unpatch  # noqa: F401
from .distributed import shard_frames, init_distributed  # noqa: F401

__version__ = "0.1.0"

"""Per-kernel CUDA-event times of the sparse-output path (what HostPipeline / voxel_filter run) at cfg2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muvo_b200 import _lib, synth  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

dev = torch.device("cuda", 0)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 96
if len(sys.argv) > 2:
    _lib.load().muvo_debug_set_tuning(5, int(sys.argv[2]))
pts, sem, off = synth.lidar_batch(F, 60000, 100000, 2000)
tp, ts, to = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)
stream = _lib.current_stream(dev)
out = {}
for it in range(2):
    with _lib.profile(stream) as p:
        r = sensor_to_grid(tp, ts, to, grid=GridSpec(), range_spec=RangeSpec(lidar_position=(1.0, 0.0, 2.0)), dense=False, sparse=True,
                           layout="hwc", out=out, packed_sparse=True)
    out = {k: v for k, v in r.items() if k != "diag"}
print(F, "frames:", "  ".join(f"{k}={1e3 * ms:.1f}" for k, ms in p.kernels), " total", round(1e3 * sum(ms for _, ms in p.kernels), 1), "us")

"""cfg2: per-kernel CUDA-event times of the point path with both stages, voxel only, range only, and the neighbour filter off.
python tools/points_split.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muvo_b200 import _lib, synth  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
pts, sem, off = synth.lidar_batch(96, 60000, 100000, 2000)
tp, ts, to = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)
remap = torch.from_numpy(synth.label_remap256()).to(dev)
stream = _lib.current_stream(dev)
G, R = GridSpec(), RangeSpec(lidar_position=(1.0, 0.0, 2.0))


def run(name, reps=10, **kw):
    out = {}

    def step():
        nonlocal out
        r = sensor_to_grid(tp, ts, to, remap=remap, layout="xyzd", out=out, **kw)
        out = {k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem") if k in r}
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    acc = {}
    for _ in range(reps):
        with _lib.profile(stream) as p:
            step()
        for i, (k, ms) in enumerate(p.kernels):
            acc.setdefault((i, k), []).append(ms)
    parts = "  ".join(f"{k}={1e3 * sum(x) / len(x):.1f}" for (i, k), x in sorted(acc.items()))
    print(f"{name:28s} step {1e3 * e0.elapsed_time(e1) / reps:7.1f} us | {parts}", flush=True)


run("both", grid=G, range_spec=R)
run("voxel only", grid=G)
run("range only", range_spec=R)
lib.muvo_debug_set_tuning(1, 1)
run("both, no neighbour filter", grid=G, range_spec=R)
run("voxel only, no filter", grid=G)
lib.muvo_debug_set_tuning(1, 0)
# shuffled points inside every frame: what the path costs when the cloud is NOT in scan order
o = to.cpu().tolist()
g = torch.Generator().manual_seed(1)
perm = torch.cat([torch.randperm(o[i + 1] - o[i], generator=g) + o[i] for i in range(len(o) - 1)]).to(dev)
tp, ts = tp[perm].contiguous(), ts[perm].contiguous()
run("both, shuffled frames", grid=G, range_spec=R)
run("voxel only, shuffled", grid=G)
run("range only, shuffled", range_spec=R)

"""Timing experiments on the point pass (library built with MUVO_NVCC_EXTRA=-DMUVO_TIMING_KNOBS; results are wrong by design):
tuning key 1 bit 1 = no table atomics, bit 2 = red.max instead of atom.max.  python tools/points_knobs.py"""
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
    acc = {}
    for _ in range(reps):
        with _lib.profile(stream) as p:
            step()
        for i, (k, ms) in enumerate(p.kernels):
            acc.setdefault((i, k), []).append(ms)
    parts = "  ".join(f"{k}={1e3 * sum(x) / len(x):.1f}" for (i, k), x in sorted(acc.items()))
    print(f"{name:34s} | {parts}", flush=True)


for knob, what in ((0, "atom.max (shipped)"), (2, "no table atomics"), (4, "red.max, no return")):
    lib.muvo_debug_set_tuning(1, knob)
    run(f"both, {what}", grid=G, range_spec=R)
    run(f"voxel only, {what}", grid=G)
    run(f"range only, {what}", range_spec=R)
lib.muvo_debug_set_tuning(1, 0)

"""cfg2 step, per-kernel CUDA-event times for several values of a tuning knob:
python tools/points_sweep.py [--key 0] [--values 0,2,3,4,5,6] [--frames 96]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muvo_b200 import _lib, synth  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--key", type=int, default=0)
ap.add_argument("--values", default="0,2,3,4,5,6")
ap.add_argument("--frames", type=int, default=96)
ap.add_argument("--nmin", type=int, default=60000)
ap.add_argument("--nmax", type=int, default=100000)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--set", default="", help="other knobs, e.g. 0=3,1=0")
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
ap_alt = int(os.environ.get("MUVO_SWEEP_ALT", "1"))       # > 1: rotate over that many DIFFERENT batches (no table sector is reused step to step)
batches = []
for b in range(ap_alt):
    pts, sem, off = synth.lidar_batch(a.frames, a.nmin, a.nmax, 2000 + 7919 * b)
    batches.append((torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)))
tp, ts, to = batches[0]
step_no = 0
remap = torch.from_numpy(synth.label_remap256()).to(dev)
stream = _lib.current_stream(dev)
out = {}


def step():
    global out, step_no, tp, ts, to
    tp, ts, to = batches[step_no % len(batches)]
    step_no += 1
    r = sensor_to_grid(tp, ts, to, grid=GridSpec(), range_spec=RangeSpec(lidar_position=(1.0, 0.0, 2.0)), remap=remap,
                       layout="xyzd", out=out)
    out = {k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem")}


for kv in [x for x in a.set.split(",") if x]:
    lib.muvo_debug_set_tuning(int(kv.split("=")[0]), int(kv.split("=")[1]))
for v in [int(x) for x in a.values.split(",")]:
    lib.muvo_debug_set_tuning(a.key, v)
    if a.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            step()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"knob{a.key}={v}: graph replay step {1e3 * e0.elapsed_time(e1) / a.reps:.1f} us", flush=True)
        continue
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    acc = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    for _ in range(a.reps):
        with _lib.profile(stream) as p:
            step()
        for i, (k, ms) in enumerate(p.kernels):
            acc.setdefault((i, k), []).append(ms)
    parts = "  ".join(f"{k}={1e3 * sum(x) / len(x):.1f}" for (i, k), x in sorted(acc.items()))
    print(f"knob{a.key}={v}: step {1e3 * e0.elapsed_time(e1) / a.reps:.1f} us | {parts}", flush=True)
lib.muvo_debug_set_tuning(a.key, 0)

"""cfg2: one fused call vs the two stages as separate calls on two streams (voxel-only | range-only), forked / joined with events.
python tools/points_two_streams.py"""
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
G, R = GridSpec(), RangeSpec(lidar_position=(1.0, 0.0, 2.0))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
outs = [{}, {}, {}]


def call(slot, **kw):
    r = sensor_to_grid(tp, ts, to, remap=remap, layout="xyzd", out=outs[slot], **kw)
    outs[slot] = {k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem") if k in r}


def fused():
    call(0, grid=G, range_spec=R)


def split():
    main = torch.cuda.current_stream()
    s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s1):
        call(1, grid=G)
    with torch.cuda.stream(s2):
        call(2, range_spec=R)
    main.wait_stream(s1); main.wait_stream(s2)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def graphed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    return g.replay


print(f"fused call            : {timeit(fused):7.1f} us", flush=True)
for knob in (0, 1, 2, 3):
    lib.muvo_debug_set_tuning(0, knob)
    print(f"two streams, knob0={knob}  : {timeit(split):7.1f} us", flush=True)
lib.muvo_debug_set_tuning(0, 0)
gf = graphed(fused)
print(f"fused, graph replay   : {timeit(gf):7.1f} us", flush=True)
for knob in (0, 2):
    lib.muvo_debug_set_tuning(0, knob)
    gs = graphed(split)
    print(f"two streams graph, knob0={knob}: {timeit(gs):7.1f} us", flush=True)
lib.muvo_debug_set_tuning(0, 0)
a = sensor_to_grid(tp, ts, to, remap=remap, layout="xyzd", grid=G, range_spec=R)
split(); torch.cuda.synchronize()
print("equal:", torch.equal(a["voxel"], outs[1]["voxel"]), torch.equal(a["range_xyzd"], outs[2]["range_xyzd"]),
      torch.equal(a["range_sem"], outs[2]["range_sem"]))

"""cfg2: voxel-only call on one stream, range-only call on a second stream that starts `delay` us later (torch.cuda._sleep), so that
the range point pass (issue bound) runs next to the dense emit (DRAM bound).  python tools/points_staggered.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muvo_b200 import _lib, synth  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
batches = []
for b in range(3):
    pts, sem, off = synth.lidar_batch(96, 60000, 100000, 2000 + 7919 * b)
    batches.append((torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)))
remap = torch.from_numpy(synth.label_remap256()).to(dev)
G, R = GridSpec(), RangeSpec(lidar_position=(1.0, 0.0, 2.0))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
outs = [{}, {}, {}]
step_no = 0


def call(slot, b, **kw):
    tp, ts, to = batches[b]
    r = sensor_to_grid(tp, ts, to, remap=remap, layout="xyzd", out=outs[slot], **kw)
    outs[slot] = {k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem") if k in r}


def fused():
    global step_no
    call(0, step_no % 3, grid=G, range_spec=R)
    step_no += 1


def make_split(delay_cycles, range_first):
    def split():
        global step_no
        b = step_no % 3
        step_no += 1
        main = torch.cuda.current_stream()
        s1.wait_stream(main); s2.wait_stream(main)
        first, second = (s2, s1) if range_first else (s1, s2)
        with torch.cuda.stream(second):
            if delay_cycles:
                torch.cuda._sleep(delay_cycles)
        with torch.cuda.stream(s1):
            call(1, b, grid=G)
        with torch.cuda.stream(s2):
            call(2, b, range_spec=R)
        main.wait_stream(s1); main.wait_stream(s2)
    return split


def timeit(fn, reps=21):
    for _ in range(6):
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
    global step_no
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gs = []
    for b in range(3):                       # one graph per batch (the batch index is baked into a capture)
        step_no = b
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        gs.append(g)
    k = [0]

    def replay():
        gs[k[0] % 3].replay(); k[0] += 1
    return replay


print(f"fused (graph replay)                         : {timeit(graphed(fused)):7.1f} us", flush=True)
for knob in (0, 2, 3):
    lib.muvo_debug_set_tuning(0, knob)
    for delay_us in (0, 40, 70, 100):
        for rf in (False, True):
            t = timeit(graphed(make_split(int(delay_us * 1965), rf)))
            print(f"point CTAs/SM knob {knob}, {'range' if rf else 'voxel'} call first, other delayed {delay_us:3d} us : {t:7.1f} us", flush=True)
lib.muvo_debug_set_tuning(0, 0)

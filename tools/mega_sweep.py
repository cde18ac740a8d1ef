"""Time the dataflow point kernel over tuning settings: python tools/mega_sweep.py "4=8 5=4" "4=16 5=8" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from muvo_b200 import _lib, synth
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid

lib = _lib.load()
dev = torch.device("cuda", 0)
frames = int(os.environ.get("FRAMES", "96"))
pts, sem, off = bench.make_batch(0, frames)
d_pts, d_sem, d_off = (torch.from_numpy(x).to(dev) for x in (pts, sem, off))
grid, rs = GridSpec(), RangeSpec(lidar_position=bench.LIDAR)
remap = torch.from_numpy(synth.label_remap256()).to(dev)
out = {}
def step():
    global out
    out = sensor_to_grid(d_pts, d_sem, d_off, grid=grid, range_spec=rs, dense=True, remap=remap, layout="xyzd", out=out)
for cfg in sys.argv[1:] or [""]:
    for k in range(8):
        lib.muvo_debug_set_tuning(k, 0)
    for kv in cfg.split():
        k, v = kv.split("=")
        lib.muvo_debug_set_tuning(int(k), int(v))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"[{cfg:>24s}] {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us/step", flush=True)

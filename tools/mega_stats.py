"""Per-CTA phase breakdown of the dataflow point kernel (points_mega.cu) on the cfg2 batch.

    python tools/mega_stats.py [--frames 96] [--steps 5] [--tuning k=v ...]
"""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from muvo_b200 import _lib, synth
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=96)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--nmin", type=int, default=60000)
ap.add_argument("--nmax", type=int, default=100000)
ap.add_argument("--tuning", nargs="*", default=[])
ap.add_argument("--mode", default="both", choices=["both", "vox", "range"])
args = ap.parse_args()
lib = _lib.load()
lib.muvo_debug_set_tuning(3, 2)
for kv in args.tuning:
    k, v = kv.split("=")
    lib.muvo_debug_set_tuning(int(k), int(v))
dev = torch.device("cuda", 0)
pts, sem, off = bench.make_batch(0, args.frames, args.nmin, args.nmax)
d_pts, d_sem, d_off = (torch.from_numpy(x).to(dev) for x in (pts, sem, off))
grid = GridSpec() if args.mode in ("both", "vox") else None
rs = RangeSpec(lidar_position=bench.LIDAR) if args.mode in ("both", "range") else None
remap = torch.from_numpy(synth.label_remap256()).to(dev)
out = {}
def step():
    global out
    out = sensor_to_grid(d_pts, d_sem, d_off, grid=grid, range_spec=rs, dense=True, remap=remap, layout="xyzd", out=out)
for _ in range(3):
    step()
torch.cuda.synchronize()
lib.muvo_debug_mega_stats(None, 0, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
n = 1024
buf = (C.c_ulonglong * (n * 20))()
lib.muvo_debug_mega_stats(buf, n, 1)
st = np.frombuffer(buf, dtype=np.uint64).reshape(n, 20).astype(np.float64) / args.steps
st = st[st[:, 11] > 0]
names = ["cons_wait", "P", "ER", "ED", "drain", "prod_wait_stage", "prod_wait_ready", "prod_wait_slot", "nP", "nER", "nED", "total", "q_entries", "q_slow_vox", "q_slow_pix", "q_band_vox", "q_band_pix", "drainA", "drainB", "-"]
print(f"{ms*1e3:.1f} us/step, {len(st)} CTAs, {pts.shape[0]} points, {args.frames} frames")
for i, nm in enumerate(names):
    c = st[:, i]
    print(f"  {nm:16s} mean {c.mean():10.0f}  min {c.min():10.0f}  max {c.max():10.0f}   (sum {c.sum():.3g})")
tot_units = st[:, 8:11].sum(0)
print("  cycles/unit: P %.0f  ER %.0f  ED %.0f" % tuple(st[:, 1 + k].sum() / max(tot_units[k], 1) for k in range(3)))

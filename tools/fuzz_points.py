"""Randomised configurations of stages (a) + (b) against the oracle (bit-exact): grid sizes / resolutions / offsets, range-image
shapes incl. H*W not a multiple of 4, fields of view, sensor positions, ragged batches with empty and one-point frames, scan-ordered
and shuffled clouds, float64 points for the voxeliser.  python tools/fuzz_points.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (the checker)
from muvo_b200 import synth  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
dev = torch.device("cuda", 0)
bad = 0
for case in range(n_cases):
    res = float(rng.choice([0.5, 0.25, 0.2, 0.4, 1.0, 0.3]))
    size = [int(rng.choice([16, 48, 100, 192, 256])), int(rng.choice([16, 50, 192, 160])), int(rng.choice([8, 20, 64, 32]))]
    offset = [float(rng.choice([0.0, 1.3, -2.0])), float(rng.choice([0.0, 0.7])), float(rng.choice([-10.0, 0.0, -3.5]))]
    H, W = int(rng.choice([64, 32, 30, 16, 128])), int(rng.choice([1024, 1000, 512, 2048, 250]))
    fov_down, fov_up = float(rng.choice([-30, -25, -45, -10])), float(rng.choice([10, 2, 15, 30]))
    lidar = [float(rng.choice([1.0, 0.0, -0.5])), float(rng.choice([0.0, 0.3])), float(rng.choice([2.0, 1.5, 0.0]))]
    F = int(rng.integers(1, 6))
    sizes = [int(rng.choice([0, 1, 7, 500, 3000, 20000])) for _ in range(F)]
    pts_l, sem_l = [], []
    for f, n in enumerate(sizes):
        if n == 0:
            pts_l.append(np.zeros((0, 3), np.float32)); sem_l.append(np.zeros((0,), np.uint8)); continue
        p, s = synth.carla_lidar_frame(n, int(rng.integers(1 << 30)))
        if rng.random() < 0.4:
            perm = rng.permutation(n); p, s = p[perm], s[perm]
        if rng.random() < 0.3:                      # coarse coordinates: many exact ties and points on voxel faces
            p = (np.round(p / res * 2) * res / 2).astype(np.float32)
        pts_l.append(p); sem_l.append(s)
    pts, sem = np.concatenate(pts_l), np.concatenate(sem_l)
    off = np.r_[0, np.cumsum(sizes)].astype(np.int64)
    if len(pts) == 0:
        continue
    f64 = rng.random() < 0.25
    special = rng.random() < 0.2                   # NaN / +-inf / huge / denormal coordinates: the voxeliser drops what is outside
    if special:                                    # (the reference's range projection raises on them, so voxel-only calls)
        k = rng.integers(0, len(pts), min(len(pts), 12))
        vals = np.array([np.nan, np.inf, -np.inf, 1e30, -1e30, 1e-42, -0.0, 3.4e38], np.float32)
        pts = pts.copy()
        pts[k, rng.integers(0, 3, len(k))] = vals[rng.integers(0, len(vals), len(k))]
    remap = synth.label_remap256() if rng.random() < 0.5 else None
    grid = GridSpec(voxel_resolution=res, voxel_size=tuple(size), offset=tuple(offset))
    rs = RangeSpec(H=H, W=W, fov_down=fov_down, fov_up=fov_up, lidar_position=tuple(lidar))
    tp = torch.from_numpy(pts.astype(np.float64) if f64 else pts).to(dev)
    ts = torch.from_numpy(sem).to(dev)
    layout = str(rng.choice(["xyzd", "hwc"]))
    sparse = bool(rng.random() < 0.5)
    kw = dict(grid=grid, remap=torch.from_numpy(remap) if remap is not None else None, layout=layout, dense=True, sparse=sparse)
    do_range = not f64 and not special
    if do_range:
        kw["range_spec"] = rs
    r = sensor_to_grid(tp, ts, off, **kw)
    torch.cuda.synchronize()
    ok = True
    for f in range(F):
        p, s = pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]]
        pin = p.astype(np.float64) if f64 else p
        v0, l0 = O.voxel_filter_fast(pin, s, res, list(size), list(offset))
        want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), tuple(size), remap) if remap is not None else None
        if want is None:
            want = np.zeros(tuple(size), np.uint8)
            lab = l0.copy(); lab[lab == 255] = 0
            want[v0[:, 0], v0[:, 1], v0[:, 2]] = lab
        ok &= np.array_equal(r["voxel"][f].cpu().numpy(), want) and int(r["n_occ"][f]) == len(v0)
        if sparse:
            rows = r["voxel_sparse"].cpu().numpy().view(np.uint16)[off[f]:off[f] + len(v0)]
            ok &= np.array_equal(rows[:, :3], v0) and np.array_equal(rows[:, 3].astype(np.uint8), l0)
        if do_range:
            if len(p) == 0:
                d0, x0, s0 = -np.ones((H, W), np.float32), np.zeros((H, W, 3), np.float32), np.zeros((H, W), np.uint8)
            else:
                d0, x0, s0 = O.range_projection(p, s, H=H, W=W, fov_down=fov_down, fov_up=fov_up, lidar_position=lidar)
            if layout == "xyzd":
                ok &= np.array_equal(r["range_xyzd"][f].cpu().numpy(), O.pack_range_view(d0, x0))
            else:
                ok &= np.array_equal(r["range_depth"][f].cpu().numpy(), d0) and np.array_equal(r["range_xyz"][f].cpu().numpy(), x0)
            ok &= np.array_equal(r["range_sem"][f].cpu().numpy(), s0)
    if not ok:
        bad += 1
        print("MISMATCH case", case, dict(res=res, size=size, offset=offset, H=H, W=W, fov=(fov_down, fov_up), lidar=lidar, sizes=sizes,
                                          f64=f64, layout=layout, sparse=sparse, remap=remap is not None), flush=True)
print(f"{n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)

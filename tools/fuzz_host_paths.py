"""Randomised runs of the host-facing paths against the oracle: HostPipeline (ragged batches of very different sizes back to back
through the same slots, sparse / dense, both layouts, device_out), merge_pcd on random images / sweeps, lidar_range_view (LiDAR
prep fused into the point kernels) on ragged batches.  python tools/fuzz_host_paths.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (the checker)
from muvo_b200 import synth  # noqa: E402
from muvo_b200.pipeline import HostPipeline  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, lidar_range_view, merge_pcd_arrays  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rng = np.random.default_rng(seed)
dev = torch.device("cuda", 0)
GRID = (0.5, [192, 192, 64], [0.0, 0.0, -10.0])
LIDAR = [1.0, 0.0, 2.0]
bad = 0


def batch(F, lo, hi):
    sizes = [int(rng.integers(lo, hi + 1)) if rng.random() > 0.15 else 0 for _ in range(F)]
    if sum(sizes) == 0:
        sizes[0] = 5
    pl, sl = [], []
    for n in sizes:
        if n == 0:
            pl.append(np.zeros((0, 3), np.float32)); sl.append(np.zeros((0,), np.uint8)); continue
        p, s = synth.carla_lidar_frame(n, int(rng.integers(1 << 30)))
        pl.append(p); sl.append(s)
    return np.concatenate(pl), np.concatenate(sl), np.r_[0, np.cumsum(sizes)].astype(np.int64)


def check_frame(r, f, p, s, layout, sparse, dense, remap, host):
    get = (lambda t: t.numpy()) if host else (lambda t: t.cpu().numpy())
    ok = True
    v0, l0 = O.voxel_filter_fast(p, s, *GRID)
    ok &= int(get(r["n_occ"])[f]) == len(v0)
    if sparse:
        st = get(r["sparse_start"])
        rows = get(r["voxel_sparse"]).view(np.uint16)[st[f]:st[f] + len(v0)]
        ok &= np.array_equal(rows[:, :3], v0) and np.array_equal(rows[:, 3].astype(np.uint8), l0)
    if dense:
        want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), (192, 192, 64), remap)
        ok &= np.array_equal(get(r["voxel"])[f], want)
    if len(p):
        d0, x0, s0 = O.range_projection(p, s, lidar_position=LIDAR)
    else:
        d0, x0, s0 = -np.ones((64, 1024), np.float32), np.zeros((64, 1024, 3), np.float32), np.zeros((64, 1024), np.uint8)
    if layout == "xyzd":
        ok &= np.array_equal(get(r["range_xyzd"])[f], O.pack_range_view(d0, x0))
    else:
        ok &= np.array_equal(get(r["range_depth"])[f], d0) and np.array_equal(get(r["range_xyz"])[f], x0)
    ok &= np.array_equal(get(r["range_sem"])[f], s0)
    return bool(ok)


for case in range(n_cases):
    try:
        # ---- HostPipeline: three batches of different sizes through the same slots
        layout = str(rng.choice(["hwc", "xyzd"]))
        sparse = bool(rng.random() < 0.6)
        dense = (not sparse) or bool(rng.random() < 0.3)
        device_out = bool(rng.random() < 0.3)
        remap = synth.label_remap256()
        pipe = HostPipeline(dev, grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), dense=dense, sparse=sparse, layout=layout,
                            remap=remap if dense else None, depth=int(rng.integers(1, 4)), host_threads=int(rng.integers(0, 4)), device_out=device_out)
        batches = [batch(int(rng.integers(1, 5)), 100, int(rng.choice([800, 5000, 30000]))) for _ in range(3)]
        pending = []
        for b in batches:
            if len(pending) >= len(pipe.slots):
                bb = pending.pop(0)
                r = pipe.result()
                for f in range(len(bb[2]) - 1):
                    if not check_frame(r, f, bb[0][bb[2][f]:bb[2][f + 1]], bb[1][bb[2][f]:bb[2][f + 1]], layout, sparse, dense, remap, not device_out):
                        bad += 1; print("PIPELINE MISMATCH", case, dict(layout=layout, sparse=sparse, dense=dense, device_out=device_out), flush=True); break
            pipe.submit(*b); pending.append(b)
        while pending:
            bb = pending.pop(0)
            r = pipe.result()
            for f in range(len(bb[2]) - 1):
                if not check_frame(r, f, bb[0][bb[2][f]:bb[2][f + 1]], bb[1][bb[2][f]:bb[2][f + 1]], layout, sparse, dense, remap, not device_out):
                    bad += 1; print("PIPELINE MISMATCH", case, dict(layout=layout, sparse=sparse, dense=dense, device_out=device_out), flush=True); break
        pipe.close()
        # ---- merge_pcd on a random image + sweep
        h, w = int(rng.choice([60, 120, 600])), int(rng.choice([96, 200, 960]))
        img = synth.carla_depth_image(int(rng.integers(1 << 30)), h, w)
        n = int(rng.choice([0, 1, 3000, 40000]))
        if n:
            p1, s1 = synth.carla_lidar_frame(n, int(rng.integers(1 << 30)))
            lid = p1.copy(); lid[:, 1] *= -1; lid -= np.asarray(LIDAR, np.float32)
        else:
            lid, s1 = np.zeros((0, 3), np.float32), np.zeros((0,), np.uint8)
        fov = float(rng.choice([110, 90, 60]))
        me = bool(rng.random() < 0.7)
        gp, gs = merge_pcd_arrays(img, lid, s1, [1.0, 0.0, 2.0], LIDAR, fov=fov, mask_ego=me)
        wp, ws = O.merge_pcd_arrays(img, lid, s1, [1.0, 0.0, 2.0], LIDAR, fov=fov, mask_ego=me)
        if not (np.array_equal(gp, wp) and np.array_equal(gs.reshape(-1), np.asarray(ws).reshape(-1))):
            bad += 1; print("MERGE MISMATCH", case, dict(h=h, w=w, n=n, fov=fov, mask_ego=me), flush=True)
        # ---- lidar_range_view on a ragged batch of raw sweeps
        bp, bs, bo = batch(int(rng.integers(1, 4)), 50, 20000)
        raw = bp.copy(); raw[:, 1] *= -1; raw -= np.asarray(LIDAR, np.float32)              # back to the LiDAR frame
        rm = synth.label_remap256() if rng.random() < 0.6 else None
        lay = str(rng.choice(["hwc", "xyzd"]))
        r = lidar_range_view(raw, bs, lidar_position=tuple(LIDAR), remap=rm, frame_offsets=bo, layout=lay)
        for f in range(len(bo) - 1):
            pp, ss = O.lidar_prep(raw[bo[f]:bo[f + 1]], bs[bo[f]:bo[f + 1]], LIDAR, rm)
            if len(pp):
                d0, x0, s0 = O.range_projection(pp, ss, lidar_position=LIDAR)
            else:
                d0, x0, s0 = -np.ones((64, 1024), np.float32), np.zeros((64, 1024, 3), np.float32), np.zeros((64, 1024), np.uint8)
            if lay == "xyzd":
                ok = np.array_equal(r["range_xyzd"][f].cpu().numpy(), O.pack_range_view(d0, x0))
            else:
                ok = np.array_equal(r["range_depth"][f].cpu().numpy(), d0) and np.array_equal(r["range_xyz"][f].cpu().numpy(), x0)
            ok &= np.array_equal(r["range_sem"][f].cpu().numpy(), s0)
            if not ok:
                bad += 1; print("LIDAR PREP MISMATCH", case, dict(layout=lay, remap=rm is not None, sizes=np.diff(bo).tolist()), flush=True); break
    except Exception as e:  # noqa: BLE001
        bad += 1
        import traceback
        print("EXCEPTION in case", case, repr(e)[:300], flush=True)
        traceback.print_exc(limit=4)
print(f"{n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)

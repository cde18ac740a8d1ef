"""BASELINE.json configs[4] (scaling sweep): 1 M-point frames, batch 64 x seq 12 = 768 frames, voxelise + project +
BEV-pool, frames sharded `frame % world == rank`, streamed through each GPU in chunks.

    python tools/cfg5_sweep.py [--frames 768] [--chunk 24] [--bev-chunk 24] [--distinct 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/cfg5_sweep.py ...

Inputs are device resident (the chunk buffers are reused: `--distinct` different synthetic 1 M-point frames are tiled to
fill a chunk, generating 768 distinct frames on the host would take ten minutes); every chunk is a full pass of the
four point kernels / the pool kernels over `chunk` frames.  Timed with CUDA events, max over ranks, one JSON line.
This is a measurement tool for the fifth config, not the driver's bench line (that is cfg2, bench.py).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muvo_b200  # noqa: E402
from muvo_b200 import synth  # noqa: E402
from muvo_b200.distributed import init_distributed, shard_frames  # noqa: E402
from muvo_b200.frustum_pooling import bev_pool  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=768)
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--chunk", type=int, default=24, help="frames per point-kernel pass")
ap.add_argument("--bev-chunk", type=int, default=24, help="frames per BEV-pool pass (236 MB of lifted features each)")
ap.add_argument("--distinct", type=int, default=4)
a = ap.parse_args()

rank, world, dev = init_distributed()
mine = len(shard_frames(a.frames, rank, world))            # frames this rank owns
# ---- points: (a)+(b) on chunks of 1 M-point frames
frames = [synth.carla_lidar_frame(a.points, 5000 + 17 * rank + i) for i in range(a.distinct)]
pts = np.concatenate([frames[i % a.distinct][0] for i in range(a.chunk)])
sem = np.concatenate([frames[i % a.distinct][1] for i in range(a.chunk)])
off = (np.arange(a.chunk + 1, dtype=np.int64) * a.points)
tp, ts, to = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)
remap = torch.from_numpy(synth.label_remap256()).to(dev)
out = {}


def points_pass():
    global out
    r = sensor_to_grid(tp, ts, to, grid=GridSpec(), range_spec=RangeSpec(lidar_position=(1.0, 0.0, 2.0)), remap=remap,
                       layout="xyzd", out=out)
    out = {k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem")}


def timed(fn, n):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


n_pass = -(-mine // a.chunk)
ms_points = timed(points_pass, n_pass)
# ---- (c) BEV pool fwd + bwd at C = 384 on chunks of lifted features
B, C = a.bev_chunk, 384
feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000 + rank, device=dev)
fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
x = synth.lift(feat, depth).detach().requires_grad_(True)
fp.initialize_frustum(x)
cell = fp.cell_ids(fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None]), mask)
gout = torch.randn((B, C, 48, 48), device=dev)


def bev_pass():
    o = bev_pool(x, cell, 2304)
    torch.autograd.grad(o, x, gout.view(o.shape))


n_bev = -(-mine // a.bev_chunk)
ms_bev = timed(bev_pass, n_bev)
if rank == 0:
    n_pts_frame = 37 * 40 * 104
    line = {"config": "cfg5: 1M-point frames, 64 x 12 frames, voxelise + project + BEV pool (C=384)", "n_gpus": world,
            "frames_total": a.frames, "frames_per_rank": mine, "points_per_frame": a.points,
            "points": {"chunk_frames": a.chunk, "passes": n_pass, "ms_total": ms_points,
                       "points_per_s": world * n_pass * a.chunk * a.points / (ms_points * 1e-3),
                       "frames_per_s": world * n_pass * a.chunk / (ms_points * 1e-3),
                       "algorithmic_GBps_per_gpu": n_pass * a.chunk * (13 * a.points + 192 * 192 * 64 + 65536 * 17) / ms_points / 1e6},
            "bev_pool_fwd_bwd": {"chunk_frames": B, "passes": n_bev, "ms_total": ms_bev,
                                 "frames_per_s": world * n_bev * B / (ms_bev * 1e-3),
                                 "dense_GBps_per_gpu": n_bev * B * (2 * n_pts_frame * C * 4 + 2 * C * 2304 * 4) / ms_bev / 1e6},
            "data": "synthetic, device resident, chunk buffers reused"}
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()

"""Run a few cfg2 steps (for ncu): python tools/prof_step.py [--frames 96] [--steps 4] [--stage points|bev|ssc]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muvo_b200  # noqa: E402
from muvo_b200 import synth  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=96)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--stage", default="points")
ap.add_argument("--nmin", type=int, default=60000)
ap.add_argument("--nmax", type=int, default=100000)
ap.add_argument("--tune", action="append", default=[], help="k=v for muvo_debug_set_tuning")
a = ap.parse_args()
for kv in a.tune:
    from muvo_b200 import _lib
    k, v = kv.split("=")
    _lib.check(_lib.load().muvo_debug_set_tuning(int(k), int(v)), "set_tuning")
dev = torch.device("cuda", 0)
if a.stage == "points":
    pts, sem, off = synth.lidar_batch(a.frames, a.nmin, a.nmax, 2000)
    tp, ts, to = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)
    remap = torch.from_numpy(synth.label_remap256()).to(dev)
    out = {}
    for _ in range(a.steps):
        r = sensor_to_grid(tp, ts, to, grid=GridSpec(), range_spec=RangeSpec(lidar_position=(1.0, 0.0, 2.0)), remap=remap,
                           layout="xyzd", out=out)
        out = {k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem")}
elif a.stage == "bev":
    from muvo_b200.frustum_pooling import bev_pool
    B, C = 6, 384
    feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000, device=dev)
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
    x = synth.lift(feat, depth).detach().requires_grad_(True)
    fp.initialize_frustum(x)
    cell = fp.cell_ids(fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None]), mask)
    for _ in range(a.steps):
        out = bev_pool(x, cell, 2304)
        torch.autograd.grad(out, x, torch.ones_like(out))
elif a.stage == "misc":         # every kernel that ships without a dedicated profile: sparse emit, merge, pyramids, densify, pillar scatter,
    # LiDAR-prep range path, streamed BEV pool (module call) -- one call each at bench-like sizes
    from muvo_b200 import pillars
    from muvo_b200.points import densify_voxels, label_pyramids, lidar_range_view, merge_pcd_batch
    pts, sem, off = synth.lidar_batch(24, a.nmin, a.nmax, 2000)
    tp, ts, to = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev), torch.from_numpy(off).to(dev)
    frames = []
    for k in range(4):
        img = synth.carla_depth_image(5100 + k)
        p1, s1 = synth.carla_lidar_frame(60000, 5200 + k)
        lid = p1.copy(); lid[:, 1] *= -1; lid -= torch.tensor([1.0, 0.0, 2.0]).numpy()
        frames.append((img, lid, s1))
    feat, depth, mask, K, E = synth.bev_inputs(6, 384, 3000, device=dev)
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
    x = synth.lift(feat, depth).detach()
    src = torch.randn(200_000, 64, device=dev)
    idx = torch.randint(0, 12000, (200_000,), device=dev)
    for _ in range(a.steps):
        r = sensor_to_grid(tp, ts, to, grid=GridSpec(), dense=False, sparse=True, packed_sparse=True)                 # k_emit_sparse, k_frame_prefix, scan role
        rows = r["voxel_sparse"].view(torch.int16)
        dense = densify_voxels(r["voxel_sparse"][:int(r["sparse_start"][-1])], (192, 192, 64), synth.label_remap256(),
                               frame_offsets=r["sparse_start"])                                                         # k_densify_*
        rv = lidar_range_view(tp, ts, frame_offsets=to, remap=synth.label_remap256())                                    # prep variant of the point pass
        label_pyramids(rv["range_xyzd"], rv["range_sem"], dense)                                                        # k_range_pyramid, k_voxel_pyramid
        merge_pcd_batch(frames, [1.0, 0.0, 2.0], [1.0, 0.0, 2.0])                                                       # k_merge_*
        pillars.scatter_mean(src, idx, dim_size=12000); pillars.scatter_max(src, idx, dim_size=12000)                   # k_pillar_*
        fp(x, K[:, None], E[:, None], mask)                                                                             # k_chunk_compact + k_pool_stream
elif a.stage == "other":        # (d) counts, N4 scal losses fwd + bwd, N2 fused lift-splat fwd + bwd at the bench shapes
    from muvo_b200.frustum_pooling import lift_splat
    from muvo_b200.losses import scal_losses
    from muvo_b200.metrics import ssc_counts
    yp, yt = synth.occupancy_pair(16, 2, 4000)
    tp, tt = torch.from_numpy(yp).to(dev), torch.from_numpy(yt).to(dev)
    logits = torch.randn((1, 16, 2, 192, 192, 64), device=dev).requires_grad_(True)
    feat, depth, mask, K, E = synth.bev_inputs(6, 384, 3000, device=dev)
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
    fp.initialize_frustum(synth.lift(feat[:1], depth[:1]))
    cell = fp.cell_ids(fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None]), mask)
    fl, dl = feat.detach().requires_grad_(True), depth.detach().requires_grad_(True)
    for _ in range(a.steps):
        ssc_counts(tp, tt, 2, ignore255=True)
        sem, geo = scal_losses(logits, tt.view(1, 16, 192, 192, 64))
        torch.autograd.grad(sem + geo, logits)
        ol = lift_splat(fl, dl, cell, 2304)
        torch.autograd.grad(ol, (fl, dl), torch.ones_like(ol))
else:
    from muvo_b200.metrics import ssc_counts
    yp, yt = synth.occupancy_pair(16, 2, 4000)
    tp, tt = torch.from_numpy(yp).to(dev), torch.from_numpy(yt).to(dev)
    for _ in range(a.steps):
        ssc_counts(tp, tt, 2, ignore255=True)
torch.cuda.synchronize()
print("done")

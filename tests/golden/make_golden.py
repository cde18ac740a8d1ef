"""Freeze golden vectors by running the UNMODIFIED reference (build container only).

Usage (from the repo root, with /root/reference mounted):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz`` (inputs + the reference's own outputs).  The files
are committed; the GPU box has no reference checkout and only reads them.
Versions used when the committed fixtures were produced are stored inside each
file (``meta``).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from muvo_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore", category=DeprecationWarning)
META = np.array(f"numpy {np.__version__}; torch {torch.__version__}; reference fzi-forschungszentrum-informatik/muvo")


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, meta=META, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def golden_voxel(R):
    grid = dict(voxel_resolution=0.5, voxel_size=[192, 192, 64], offset=[0.0, 0, -10.0])
    # known-answer vector, SURVEY.md A.1 item 9
    P = np.array([(-48, -48, -6), (47.99, 47.99, 25.99), (48, 0, 0), (0, 0, 26), (0, 0, -6.01), (0.1, 0.1, 0.1),
                  (0.4, 0.4, 0.4), (0.26, 0.26, 0.26), (1.05, 0, 0), (1.45, 0.45, 0.45)], dtype=np.float64)
    S = np.array([1, 2, 3, 4, 5, 7, 8, 9, 10, 6], dtype=np.uint8)
    kv, kl = R.voxel_filter(P.copy(), S, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    pts, sem = synth.carla_lidar_frame(12000, 1000)
    v32, l32 = R.voxel_filter(pts.copy(), sem, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    # float64 merged-cloud style input with (N,1) semantics (data_preprocessing.py:130-131)
    rng = np.random.default_rng(77)
    p64 = np.concatenate([pts[:6000].astype(np.float64) + rng.normal(0, 1e-3, (6000, 3)),
                          rng.uniform([-50, -50, -8], [50, 50, 28], (6000, 3))])
    s64 = rng.integers(0, 23, (12000, 1)).astype(np.uint8)
    v64, l64 = R.voxel_filter(p64.copy(), s64, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    # a different (still power-of-two) resolution and a smaller grid
    va, la = R.voxel_filter(pts.copy(), sem, 0.25, [96, 128, 32], [2.0, 0, -1.0])
    # non power-of-two resolution exercises the general np.divmod path
    vb, lb = R.voxel_filter(p64.copy(), s64, 0.2, [200, 200, 40], [0.0, 0, -1.0])
    save("voxel.npz", known_pts=P, known_sem=S, known_vox=kv, known_lab=kl,
         pts32=pts, sem32=sem, vox32=v32, lab32=l32, pts64=p64, sem64=s64, vox64=v64, lab64=l64,
         vox_alt=va, lab_alt=la, vox_np2=vb, lab_np2=lb)


def golden_range(R):
    pc = R.PointCloud(64, 1024, -30, 10, [1.0, 0.0, 2.0])
    K = np.array([(11, 0, 2), (1, 5, 2), (1, -5, 2), (11, 0, 30), (11, 0, -30), (-9, 0.0, 2), (-9, -0.0, 2),
                  (2, 1, 2), (0, 1, 2), (2, -1, 2), (0, -1, 2)], dtype=np.float32)
    KS = np.arange(1, len(K) + 1, dtype=np.uint8)
    kd, kx, ks = pc.do_range_projection(K.copy(), KS)
    pts, sem = synth.carla_lidar_frame(12000, 2000)
    d, x, s = pc.do_range_projection(pts.copy(), sem)
    # dense frame: many points per pixel -> exercises the nearest-wins resolve
    rng = np.random.default_rng(5)
    dense = (rng.normal(0, 1, (30000, 3)) * np.array([20, 20, 3]) + np.array([1, 0, 2])).astype(np.float32)
    dsem = rng.integers(0, 23, 30000).astype(np.uint8)
    dd, dx, dsm = pc.do_range_projection(dense.copy(), dsem)
    pc2 = R.PointCloud(32, 256, -25, 3, [0.5, 0.25, 1.75])
    d2, x2, s2 = pc2.do_range_projection(pts.copy(), sem)
    save("range.npz", known_pts=K, known_sem=KS, known_depth=kd, known_xyz=kx, known_semimg=ks,
         pts=pts, sem=sem, depth=d, xyz=x, semimg=s, dense_pts=dense, dense_sem=dsem, dense_depth=dd, dense_xyz=dx,
         dense_semimg=dsm, alt_depth=d2, alt_xyz=x2, alt_semimg=s2)


def golden_bev(R):
    out = {}
    # QuickCumsum / VoxelsSumming / cumsum_trick known answers
    x = torch.tensor([[1., 10.], [2., 20.], [3., 30.], [4., 40.]], requires_grad=True)
    ranks = torch.tensor([0, 0, 1, 3])
    geom = torch.arange(16).view(4, 4)
    xs, gs = R.QuickCumsum.apply(x, geom, ranks)
    (xs * torch.tensor([[1.], [2.], [3.]])).sum().backward()
    out.update(qc_x=x.detach().numpy(), qc_ranks=ranks.numpy(), qc_geom=geom.numpy(), qc_xseg=xs.detach().numpy(),
               qc_gseg=gs.numpy(), qc_gradx=x.grad.numpy())
    # random sorted-rank segment sums
    g = torch.Generator().manual_seed(11)
    xr = torch.randn(5000, 24, generator=g)
    rk = torch.sort(torch.randint(0, 700, (5000,), generator=g))[0]
    gm = torch.randint(0, 48, (5000, 4), generator=g)
    xs2, gs2 = R.cumsum_trick(xr, gm, rk)
    out.update(seg_x=xr.numpy(), seg_ranks=rk.numpy(), seg_geom=gm.numpy(), seg_xseg=xs2.numpy(), seg_gseg=gs2.numpy())
    # whole module at muvo.yml geometry (small C), train (QuickCumsum) and eval (cumsum_trick), fwd + bwd
    B, C = 1, 6
    feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000)
    for tag, m in (("mask", mask), ("nomask", torch.zeros(0))):
        fp = R.FrustumPooling(**synth.BEV_POOL_ARGS)
        fp.train()
        f = feat.clone().requires_grad_(True)
        d = depth.clone().requires_grad_(True)
        o = fp(synth.lift(f, d), K[:, None], E[:, None], m)
        gout = torch.randn(o.shape, generator=torch.Generator().manual_seed(12))
        (o * gout).sum().backward()
        out.update({f"pool_{tag}_out": o.detach().numpy(), f"pool_{tag}_gout": gout.numpy(),
                    f"pool_{tag}_gfeat": f.grad.numpy(),
                    f"pool_{tag}_gdepth_sub": d.grad.numpy()[:, :, ::4, ::4].copy()})  # subsampled to keep the file small
    out.update(pool_feat=feat.numpy(), pool_depth=depth.numpy(), pool_mask=mask.numpy(), pool_K=K.numpy(),
               pool_E=E.numpy())
    fp = R.FrustumPooling(**synth.BEV_POOL_ARGS)
    out.update(bev_intrinsics=fp.bev_intrinsics.numpy(), dx=fp.dx.numpy(), bx=fp.bx.numpy(), nx=fp.nx.numpy(),
               ds=fp.ds.numpy())
    # cell ids the reference derives for one frame (for geometry parity)
    fp.initialize_frustum(synth.lift(feat, depth))
    geom = fp.get_geometry(E[:1, None, :3, :3], E[:1, None, :3, 3:], K[:1, None])
    g2 = geom.view(-1, 3).clone()
    g2[:, 0] = g2[:, 0] * fp.bev_intrinsics[0, 0] + fp.bev_intrinsics[0, 2]
    g2[:, 1] = g2[:, 1] * fp.bev_intrinsics[1, 1] + fp.bev_intrinsics[1, 2]
    g2[:, 2] = (g2[:, 2] - fp.bx[2] + fp.dx[2] / 2.) / fp.dx[2]
    out.update(cells=g2.long().numpy().astype(np.int16))
    save("bev.npz", **out)


def golden_ssc(R):
    out = {}
    for C in (2, 9):
        yp, yt = synth.occupancy_pair(3, C, 4000 + C, size=(24, 20, 8))
        m = R.SSCMetrics(C)
        tp, tt = torch.from_numpy(yp), torch.from_numpy(yt)
        a = m.get_score_completion(tp, tt)
        b = m.get_score_semantic_and_completion(tp, tt)
        raw = np.r_[list(a), b[0].numpy(), b[1].numpy(), b[2].numpy()].astype(np.int64)
        rng = np.random.default_rng(C)
        ne = rng.random(yt.shape) < 0.8
        ns = rng.random(yt.shape) < 0.7
        a = m.get_score_completion(tp, tt, torch.from_numpy(ne))
        b = m.get_score_semantic_and_completion(tp, tt, torch.from_numpy(ne))
        masked = np.r_[list(a), b[0].numpy(), b[1].numpy(), b[2].numpy()].astype(np.int64)
        m.reset()
        m.add_batch(tp, tt)
        m.add_batch(tp, tt, torch.from_numpy(ne), torch.from_numpy(ns))
        st = m.get_stats()
        acc = np.r_[m.completion_tp, m.completion_fp, m.completion_fn, m.tps.numpy(), m.fps.numpy(), m.fns.numpy()]
        out.update({f"c{C}_pred": yp.astype(np.int8), f"c{C}_true": yt, f"c{C}_raw": raw, f"c{C}_nonempty": ne,
                    f"c{C}_nonsurface": ns, f"c{C}_masked": masked, f"c{C}_acc2": acc.astype(np.float64),
                    f"c{C}_iou": np.float64(st["iou"]), f"c{C}_precision": np.float64(st["precision"]),
                    f"c{C}_recall": np.float64(st["recall"]), f"c{C}_iou_ssc": st["iou_ssc"].numpy(),
                    f"c{C}_iou_ssc_mean": st["iou_ssc_mean"].numpy()})
    save("ssc.npz", **out)


def golden_merge(R):
    """N1: the reference's merge_pcd (file based) on a synthetic encoded depth image + LiDAR sweep, then its voxel_filter."""
    import tempfile
    import cv2
    img = synth.carla_depth_image(5000, h=150, w=240)                    # small fixture; full size runs in the tests via the oracle
    pts, sem = synth.carla_lidar_frame(8000, 5001)
    lid = pts.copy(); lid[:, 1] *= -1; lid -= np.float32([1, 0, 2])      # back to the LiDAR frame (the file format)
    with tempfile.TemporaryDirectory() as td:
        cv2.imwrite(os.path.join(td, "d.png"), img)
        np.save(os.path.join(td, "l.npy"), {"points_xyz": lid.copy(), "ObjTag": sem}, allow_pickle=True)
        pcd, s = R.merge_pcd(os.path.join(td, "d.png"), os.path.join(td, "l.npy"), [1.0, 0.0, 2.0], [1.0, 0.0, 2.0], fov=110)
        pcd2, s2 = R.merge_pcd(os.path.join(td, "d.png"), os.path.join(td, "l.npy"), [1.5, 0.25, 1.75], [1.0, 0.0, 2.0], fov=90,
                               mask_ego=False)
    vox, lab = R.voxel_filter(pcd.copy(), s, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    save("merge.npz", img=img, lidar_xyz=lid, lidar_sem=sem, pcd=pcd, sem=s, pcd_nomask=pcd2, sem_nomask=s2, vox=vox, lab=lab)


def scal_inputs(C, seed, shape=(1, 3, 12, 10, 8), ignore_frac=0.02):
    """Seeded logits (b,s,C,x,y,z) f32 and uint8 targets with a few 255s -- stored in the fixture, not regenerated."""
    g = torch.Generator().manual_seed(seed)
    b, s = shape[:2]
    pred = torch.randn((b, s, C) + tuple(shape[2:]), generator=g) * 2.0
    tgt = torch.randint(0, C, (b, s) + tuple(shape[2:]), generator=g).to(torch.uint8)
    tgt[torch.rand(tgt.shape, generator=g) < 0.6] = 0                      # mostly empty, like occupancy grids
    tgt[torch.rand(tgt.shape, generator=g) < ignore_frac] = 255
    return pred, tgt


def golden_scal(R):
    """N4: SemScalLoss / GeoScalLoss of the reference (fp32, CPU) with their gradients w.r.t. the logits."""
    out = {}
    warnings.filterwarnings("ignore", category=FutureWarning)
    for C in (2, 5, 9):
        pred, tgt = scal_inputs(C, 7000 + C)
        for name, cls in (("sem", R.SemScalLoss), ("geo", R.GeoScalLoss)):
            p = pred.clone().requires_grad_(True)
            loss = cls()(p, tgt)
            loss.backward()
            out[f"c{C}_{name}_loss"] = np.float64(loss.item())
            out[f"c{C}_{name}_grad"] = p.grad.numpy()
        out[f"c{C}_pred"] = pred.numpy()
        out[f"c{C}_target"] = tgt.numpy()
    # class 1 absent from the targets: SemScalLoss skips it (count), and no ignore voxels at all
    pred, tgt = scal_inputs(3, 7100, ignore_frac=0.0)
    tgt[tgt == 1] = 0
    p = pred.clone().requires_grad_(True)
    loss = R.SemScalLoss()(p, tgt)
    loss.backward()
    out.update(absent_pred=pred.numpy(), absent_target=tgt.numpy(), absent_sem_loss=np.float64(loss.item()), absent_sem_grad=p.grad.numpy())
    save("scal.npz", **out)


def golden_lidar(R):
    """N1, LiDAR side: the reference's own ``convert_coor_lidar`` + the literal dataset.py:281-290 lines +
    ``PointCloud.do_range_projection`` on a raw sweep (LiDAR frame, raw CARLA tags, some points inside the ego box); and the
    sparse -> dense voxel glue of dataset.py:317-327 (``muvo.data.dataset`` itself needs lightning / the CARLA dataframes)."""
    pts, sem = synth.carla_lidar_frame(12000, 6000)
    raw = pts.copy()
    raw[:, 1] *= -1
    raw -= np.float32([1.0, 0.0, 2.0])                              # back to the LiDAR frame (float32, as CARLA stores it)
    rng = np.random.default_rng(61)
    raw[:400] = (rng.uniform([-3.4, -1.0, -2.0], [1.4, 1.0, -0.4], (400, 3))).astype(np.float32)   # in / around the ego box
    tag = sem.copy()
    tag[:400] = rng.integers(0, 23, 400)
    lidar_position = [1.0, 0.0, 2.0]
    LABEL_MAP = synth.LABEL_MAP
    points = R.convert_coor_lidar(raw.copy(), lidar_position)                       # dataset.py:278
    remap = np.full((max(LABEL_MAP.keys()) + 1), max(LABEL_MAP.values()), dtype=np.uint8)   # :281
    remap[list(LABEL_MAP.keys())] = list(LABEL_MAP.values())                        # :282
    semantics = remap[tag]                                                          # :283
    x, y, z = [4.902, 2.128, 1.511]                                                 # constants.py:8
    ego_box = np.array([[-x / 2, -y / 2, 0], [x / 2, y / 2, z]])                    # :287
    ego_idx = ((ego_box[0] < points) & (points < ego_box[1])).all(axis=1)           # :288
    semantics = semantics[~ego_idx]
    points = points[~ego_idx]
    pc = R.PointCloud(64, 1024, -30, 10, lidar_position)
    d, xyz, sm = pc.do_range_projection(points, semantics)                          # :300
    xyzd = np.concatenate([xyz, d[..., None]], axis=-1).transpose((2, 0, 1))        # :302-303
    # sparse -> dense (:317-327) on a voxel_filter output with one 255 label and a duplicated row
    v, l = R.voxel_filter(pts.copy(), sem, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    voxel_data = np.concatenate([v, l[:, None].astype(np.uint16)], axis=1)
    voxel_data[5, 3] = 255
    voxel_data = np.concatenate([voxel_data, voxel_data[10:12] * np.uint16([1, 1, 1, 0]) + np.uint16([0, 0, 0, 13]),
                                 voxel_data[20:21]])       # (label 13 -> 0: the later row must win; an exact repeat)
    voxel_points = voxel_data[:, :-1]
    voxel_semantics = voxel_data[:, -1].copy()
    voxel_semantics[voxel_semantics == 255] = 0                                     # :322
    voxel_semantics = remap[voxel_semantics]                                        # :323
    voxels = np.zeros([192, 192, 64], dtype=np.uint8)                               # :324
    voxels[voxel_points[:, 0], voxel_points[:, 1], voxel_points[:, 2]] = voxel_semantics   # :325
    save("lidar.npz", raw=raw, tag=tag, remap=remap, n_ego=np.int64(ego_idx.sum()), points=points, semantics=semantics,
         depth=d, xyz=xyz, semimg=sm, xyzd=xyzd, voxel_data=voxel_data, voxels_packed=np.packbits(voxels > 0),
         voxels_nz_idx=np.flatnonzero(voxels).astype(np.int32), voxels_nz_val=voxels.reshape(-1)[np.flatnonzero(voxels)])


def main():
    R = ref_import.load()
    golden_lidar(R)
    golden_merge(R)
    golden_voxel(R)
    golden_range(R)
    golden_bev(R)
    golden_ssc(R)
    golden_scal(R)


if __name__ == "__main__":
    main()

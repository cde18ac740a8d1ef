"""The CPU oracle against the committed golden vectors (produced by the unmodified reference,
tests/golden/make_golden.py) and the known-answer vectors of SURVEY.md Appendix A."""
import numpy as np
import torch

import oracle as O
from muvo_b200 import synth

GRID = (0.5, [192, 192, 64], [0.0, 0, -10.0])


def test_voxel_known_answers(golden):
    g = golden("voxel.npz")
    for fn in (O.voxel_filter_loop, O.voxel_filter_fast):
        v, l = fn(g["known_pts"], g["known_sem"], *GRID)
        assert np.array_equal(v, g["known_vox"]) and np.array_equal(l, g["known_lab"])
    assert g["known_vox"].tolist() == [[0, 0, 0], [96, 96, 12], [98, 96, 12], [191, 191, 63]]
    assert g["known_lab"].tolist() == [1, 7, 6, 2]


def test_voxel_golden_frames(golden):
    g = golden("voxel.npz")
    for fn in (O.voxel_filter_loop, O.voxel_filter_fast):
        v, l = fn(g["pts32"], g["sem32"], *GRID)
        assert v.dtype == np.uint16 and l.dtype == np.uint8
        assert np.array_equal(v, g["vox32"]) and np.array_equal(l, g["lab32"])
        v, l = fn(g["pts64"], g["sem64"], *GRID)
        assert np.array_equal(v, g["vox64"]) and np.array_equal(l, g["lab64"])
        v, l = fn(g["pts32"], g["sem32"], 0.25, [96, 128, 32], [2.0, 0, -1.0])
        assert np.array_equal(v, g["vox_alt"]) and np.array_equal(l, g["lab_alt"])
        v, l = fn(g["pts64"], g["sem64"], 0.2, [200, 200, 40], [0.0, 0, -1.0])
        assert np.array_equal(v, g["vox_np2"]) and np.array_equal(l, g["lab_np2"])


def test_voxel_edge_cases():
    v, l = O.voxel_filter_fast(np.zeros((0, 3), np.float32), np.zeros((0,), np.uint8), *GRID)
    assert v.shape == (0, 3) and l.shape == (0,)
    # all points outside the grid
    v, l = O.voxel_filter_loop(np.full((5, 3), 1000.0, np.float32), np.ones(5, np.uint8), *GRID)
    assert v.shape == (0, 3)
    # exact ties: lowest original index wins (stable contract)
    p = np.array([[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.1, 0.1, 0.1]], np.float32)
    for fn in (O.voxel_filter_loop, O.voxel_filter_fast):
        _, l = fn(p, np.array([3, 4, 5], np.uint8), *GRID)
        assert l.tolist() == [3]
    # offset list is not mutated
    off = [0.0, 0, -10.0]
    O.voxel_filter_fast(p, np.array([3, 4, 5], np.uint8), 0.5, [192, 192, 64], off)
    assert off == [0.0, 0, -10.0]


def test_densify(golden):
    g = golden("voxel.npz")
    data = np.concatenate([g["vox32"], g["lab32"][:, None].astype(np.uint16)], 1)
    grid = O.densify_voxels(data, (192, 192, 64), synth.label_remap256())
    assert grid.dtype == np.uint8 and grid.shape == (192, 192, 64)
    assert grid.sum() == (synth.label_remap256()[g["lab32"]] > 0).sum()


def test_range_golden(golden):
    g = golden("range.npz")
    d, x, s = O.range_projection(g["known_pts"], g["known_sem"], lidar_position=[1.0, 0.0, 2.0])
    assert np.array_equal(d, g["known_depth"]) and np.array_equal(x, g["known_xyz"]) and np.array_equal(s, g["known_semimg"])
    hw = {tuple(r) for r in np.argwhere(d >= 0)}
    # SURVEY.md A.2 item 8 + axis/diagonal bins
    assert {(16, 512), (16, 256), (16, 768), (0, 512), (63, 512), (16, 0), (16, 1023)} <= hw
    d, x, s = O.range_projection(g["pts"], g["sem"], lidar_position=[1.0, 0.0, 2.0])
    assert np.array_equal(d, g["depth"]) and np.array_equal(x, g["xyz"]) and np.array_equal(s, g["semimg"])
    d, x, s = O.range_projection(g["dense_pts"], g["dense_sem"], lidar_position=[1.0, 0.0, 2.0])
    assert np.array_equal(d, g["dense_depth"]) and np.array_equal(x, g["dense_xyz"]) and np.array_equal(s, g["dense_semimg"])
    d, x, s = O.range_projection(g["pts"], g["sem"], 32, 256, -25, 3, [0.5, 0.25, 1.75])
    assert np.array_equal(d, g["alt_depth"]) and np.array_equal(x, g["alt_xyz"]) and np.array_equal(s, g["alt_semimg"])


def test_bev_known_answers(golden):
    g = golden("bev.npz")
    xs, gs = O.cumsum_trick(torch.from_numpy(g["qc_x"]), torch.from_numpy(g["qc_geom"]), torch.from_numpy(g["qc_ranks"]))
    assert np.array_equal(xs.numpy(), g["qc_xseg"]) and np.array_equal(gs.numpy(), g["qc_gseg"])
    assert g["qc_xseg"].tolist() == [[3, 30], [3, 30], [4, 40]]
    gx = O.quick_cumsum_backward(torch.tensor([[1., 1.], [2., 2.], [3., 3.]]), torch.from_numpy(g["qc_ranks"]))
    assert np.array_equal(gx.numpy(), g["qc_gradx"])
    xs, gs = O.cumsum_trick(torch.from_numpy(g["seg_x"]), torch.from_numpy(g["seg_geom"]), torch.from_numpy(g["seg_ranks"]))
    assert np.array_equal(xs.numpy(), g["seg_xseg"]) and np.array_equal(gs.numpy(), g["seg_gseg"])


def test_bev_module_golden(golden):
    g = golden("bev.npz")
    feat, depth, mask = (torch.from_numpy(g[k]) for k in ("pool_feat", "pool_depth", "pool_mask"))
    K, E = torch.from_numpy(g["pool_K"]), torch.from_numpy(g["pool_E"])
    # the seeded generator reproduces the stored inputs
    f2, d2, m2, K2, E2 = synth.bev_inputs(1, 6, 3000)
    assert torch.equal(f2, feat) and torch.equal(d2, depth) and torch.equal(m2, mask) and torch.equal(K2, K)
    for tag, m in (("mask", mask), ("nomask", torch.zeros(0))):
        x = synth.lift(feat, depth)
        o = O.frustum_pooling_forward(x, K[:, None], E[:, None], m, **synth.BEV_POOL_ARGS)
        assert np.array_equal(o.numpy(), g[f"pool_{tag}_out"])
        oe = O.frustum_pooling_forward(x, K[:, None], E[:, None], m, exact=True, **synth.BEV_POOL_ARGS)
        ref = torch.from_numpy(g[f"pool_{tag}_out"]).double()
        assert (oe - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert np.allclose(O.bev_intrinsics(**{k: synth.BEV_POOL_ARGS[k] for k in ("size", "scale", "offsetx")}), g["bev_intrinsics"])
    dx, bx, nx = O.gen_dx_bx(synth.BEV_POOL_ARGS["size"], synth.BEV_POOL_ARGS["scale"], synth.BEV_POOL_ARGS["offsetx"])
    assert np.array_equal(dx.numpy(), g["dx"]) and np.array_equal(bx.numpy(), g["bx"]) and np.array_equal(nx.numpy(), g["nx"])
    # geometry -> cell ids
    fr = O.frustum_grid(synth.BEV_POOL_ARGS["dbound"], 40, 104, 8)
    geom = O.frustum_geometry(fr, K[:1, None], E[:1, None])
    cells = O.bev_cell_ids(geom, torch.from_numpy(g["bev_intrinsics"]), bx, dx)
    assert np.array_equal(cells.numpy().astype(np.int16), g["cells"])


def test_ssc_golden(golden):
    g = golden("ssc.npz")
    for C in (2, 9):
        yp, yt = g[f"c{C}_pred"].astype(np.int64), g[f"c{C}_true"]
        assert np.array_equal(O.ssc_counts(yp, yt, C), g[f"c{C}_raw"])
        assert np.array_equal(O.ssc_counts_loop(yp, yt, C), g[f"c{C}_raw"])
        ne, ns = g[f"c{C}_nonempty"], g[f"c{C}_nonsurface"]
        assert np.array_equal(O.ssc_counts(yp, yt, C, nonempty=ne), g[f"c{C}_masked"])
        assert np.array_equal(O.ssc_counts_loop(yp, yt, C, nonempty=ne), g[f"c{C}_masked"])
        acc = O.ssc_add_batch_counts(yp, yt, C) + O.ssc_add_batch_counts(yp, yt, C, ne, ns)
        assert np.array_equal(acc.astype(np.float64), g[f"c{C}_acc2"])
        st = O.ssc_stats_from_counts(acc, C)
        assert st["iou"] == float(g[f"c{C}_iou"]) and st["precision"] == float(g[f"c{C}_precision"])
        assert np.allclose(st["iou_ssc"].numpy(), g[f"c{C}_iou_ssc"], rtol=1e-6)


def test_merge_pcd_oracle_matches_reference_golden(golden):
    """N1: the numpy restatement of merge_pcd reproduces the reference's own output (frozen by make_golden.py)."""
    g = golden("merge.npz")
    pcd, sem = O.merge_pcd_arrays(g["img"], g["lidar_xyz"], g["lidar_sem"], [1.0, 0.0, 2.0], [1.0, 0.0, 2.0], fov=110)
    assert pcd.dtype == np.float64 and np.array_equal(pcd, g["pcd"]) and np.array_equal(sem, g["sem"])
    pcd2, sem2 = O.merge_pcd_arrays(g["img"], g["lidar_xyz"], g["lidar_sem"], [1.5, 0.25, 1.75], [1.0, 0.0, 2.0], fov=90,
                                    mask_ego=False)
    assert np.array_equal(pcd2, g["pcd_nomask"]) and np.array_equal(sem2, g["sem_nomask"])
    v, l = O.voxel_filter_fast(pcd, sem, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    assert np.array_equal(v, g["vox"]) and np.array_equal(l, g["lab"])


def test_label_pyramids_oracle_matches_torch_nearest():
    """N1: the pyramid restatement equals what PreProcess computes with torchvision resize NEAREST / F.interpolate
    (muvo/models/preprocess.py:151-186), also for sizes that are not multiples of 4."""
    import torch.nn.functional as Fn
    rng = np.random.default_rng(3)
    for (H, W, X, Y, Z) in ((64, 1024, 192, 192, 64), (10, 22, 13, 9, 6)):
        xyzd = rng.normal(0, 30, (3, 4, H, W)).astype(np.float32)
        sem = rng.integers(0, 23, (3, H, W)).astype(np.uint8)
        vox = rng.integers(0, 3, (2, X, Y, Z)).astype(np.uint8)
        got = O.label_pyramids(xyzd, sem, vox, scale=50.0)
        l1 = torch.from_numpy(xyzd).float() / 50.0
        l2 = Fn.interpolate(l1, (H // 2, W // 2), mode="nearest"); l4 = Fn.interpolate(l2, (H // 4, W // 4), mode="nearest")
        assert np.array_equal(got["range_view_label_1"], l1.numpy()) and np.array_equal(got["range_view_label_2"], l2.numpy())
        assert np.array_equal(got["range_view_label_4"], l4.numpy())
        s1 = torch.from_numpy(sem)[:, None]
        s2 = Fn.interpolate(s1, (H // 2, W // 2), mode="nearest"); s4 = Fn.interpolate(s2, (H // 4, W // 4), mode="nearest")
        assert np.array_equal(got["range_view_seg_label_2"], s2[:, 0].numpy()) and np.array_equal(got["range_view_seg_label_4"], s4[:, 0].numpy())
        v1 = torch.from_numpy(vox)[:, None]
        v2 = Fn.interpolate(v1, (X // 2, Y // 2, Z // 2), mode="nearest"); v4 = Fn.interpolate(v2, (X // 4, Y // 4, Z // 4), mode="nearest")
        assert np.array_equal(got["voxel_label_2"], v2[:, 0].numpy()) and np.array_equal(got["voxel_label_4"], v4[:, 0].numpy())


def test_scal_losses_golden(golden):
    """N4: float64 restatement vs the reference's fp32 SemScalLoss / GeoScalLoss (losses.py:191-287)."""
    g = golden("scal.npz")
    for C in (2, 5, 9):
        pred, tgt = g[f"c{C}_pred"], g[f"c{C}_target"]
        assert abs(O.sem_scal_loss(pred, tgt) - float(g[f"c{C}_sem_loss"])) <= 1e-6 * float(g[f"c{C}_sem_loss"])
        assert abs(O.geo_scal_loss(pred, tgt) - float(g[f"c{C}_geo_loss"])) <= 1e-6 * float(g[f"c{C}_geo_loss"])
        s = O.scal_sums(pred, tgt)
        assert s[3 * C] == (tgt != 255).sum() and abs(s[:C].sum() - s[3 * C]) < 1e-6
    assert abs(O.sem_scal_loss(g["absent_pred"], g["absent_target"]) - float(g["absent_sem_loss"])) <= 1e-6 * float(g["absent_sem_loss"])


def test_scatter_oracle_known_answers():
    """N4 (PointPillar): restated torch_scatter semantics on a hand-checked case; torch.scatter_reduce as a second opinion."""
    src = np.array([[1., 2], [3, 1], [3, 5], [0, 0]], dtype=np.float32)
    idx = np.array([0, 2, 2, 0])
    assert O.scatter_mean(src, idx, 4).tolist() == [[0.5, 1.0], [0, 0], [3.0, 3.0], [0, 0]]
    mx, arg = O.scatter_max(src, idx, 4)
    assert mx.tolist() == [[1, 2], [0, 0], [3, 5], [0, 0]] and arg.tolist() == [[0, 0], [4, 4], [1, 2], [4, 4]]
    t = torch.zeros((4, 2)).scatter_reduce(0, torch.from_numpy(idx)[:, None].expand(-1, 2), torch.from_numpy(src), "amax", include_self=False)
    assert np.array_equal(t.numpy(), mx)


def test_lidar_prep_and_densify_match_reference_golden(golden):
    """N1 (LiDAR side): the restatement of dataset.py:278-290 / :317-327 against the reference's own functions run on a raw
    sweep (tests/golden/make_golden.py:golden_lidar)."""
    g = golden("lidar.npz")
    p, s = O.lidar_prep(g["raw"], g["tag"], [1.0, 0.0, 2.0], g["remap"])
    assert p.dtype == np.float32 and np.array_equal(p, g["points"]) and np.array_equal(s, g["semantics"])
    assert len(g["raw"]) - len(p) == int(g["n_ego"]) > 0
    d, x, sm = O.range_projection(p, s, lidar_position=[1.0, 0.0, 2.0])
    assert np.array_equal(d, g["depth"]) and np.array_equal(x, g["xyz"]) and np.array_equal(sm, g["semimg"])
    assert np.array_equal(O.pack_range_view(d, x), g["xyzd"])
    grid = O.densify_voxels(g["voxel_data"], (192, 192, 64), g["remap"])
    assert np.array_equal(np.flatnonzero(grid), g["voxels_nz_idx"]) and np.array_equal(grid.reshape(-1)[g["voxels_nz_idx"]], g["voxels_nz_val"])

"""Host-side logic that needs no GPU: spec constants, sharding, patching, synthetic generators, loud failure."""
import sys
import types

import numpy as np
import pytest
import torch

import muvo_b200
import oracle as O
from muvo_b200 import synth
from muvo_b200.points import GridSpec, RangeSpec, PointCloud


def test_grid_spec_matches_oracle_params():
    for res, size, off in ((0.5, [192, 192, 64], [0.0, 0, -10.0]), (0.2, [200, 200, 40], [0.0, 0, -1.0]),
                           (0.25, [96, 128, 32], [2.0, 0, -1.0])):
        g = GridSpec(res, tuple(size), tuple(off)).to_c()
        o, u = O.voxel_grid_params(res, size, off)
        assert list(g.offset) == o.tolist() and list(g.upper) == u.tolist() and list(g.size) == size
    user = np.array([0.0, 0.0, -10.0])
    GridSpec(0.5, (192, 192, 64), user).to_c()
    assert user.tolist() == [0.0, 0.0, -10.0]          # never mutated
    GridSpec(0.5, (192, 192, 64), [0, 0, -10]).to_c()  # all-int offsets are accepted


def test_range_spec_constants():
    c = RangeSpec(64, 1024, -30, 10, (1, 0, 2)).to_c()
    assert c.fov_down_abs == abs(-30 / 180.0 * np.pi) and c.fov == 10 / 180.0 * np.pi - (-30 / 180.0 * np.pi)
    pc = PointCloud(64, 1024, -30, 10, [1, 0, 2])
    c2 = pc.spec.to_c()
    assert (c2.fov, c2.fov_down_abs, list(c2.lidar_pos)) == (c.fov, c.fov_down_abs, [1.0, 0.0, 2.0])
    assert pc.H == 64 and pc.W == 1024 and pc.fov == c.fov


def test_restore_pcd_coor_shape():
    pc = PointCloud(8, 16, -30, 10, [1, 0, 2])
    out = pc.restore_pcd_coor(np.ones((1, 1, 8, 16)))
    assert out.shape == (1, 1, 8, 16, 4)


def test_no_cuda_is_loud():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(muvo_b200.MuvoError):
        muvo_b200.voxel_filter(np.zeros((4, 3), np.float32), np.zeros(4, np.uint8), 0.5, [192, 192, 64], [0.0, 0, -10.0])
    with pytest.raises(muvo_b200.MuvoError):
        PointCloud().do_range_projection(np.ones((4, 3), np.float32), np.zeros(4, np.uint8))
    with pytest.raises(muvo_b200.MuvoError):
        muvo_b200.SSCMetrics(2).add_batch(torch.zeros(1, 4, 4, 4, dtype=torch.int64), torch.zeros(1, 4, 4, 4, dtype=torch.uint8))
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS)
    feat, depth, mask, K, E = synth.bev_inputs(1, 2, 1, fH=4, fW=6)
    with pytest.raises(muvo_b200.MuvoError):
        fp(synth.lift(feat, depth), K[:, None], E[:, None], mask)


def test_product_never_imports_oracle():
    import os, re
    root = os.path.dirname(os.path.abspath(muvo_b200.__file__))
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "/root/reference" not in text, f


def test_frustum_pooling_buffers_match_reference_contract(golden):
    g = golden("bev.npz")
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS)
    assert list(fp.state_dict().keys()) == ["bev_intrinsics"]          # only persistent buffer (checkpoint compat)
    assert np.array_equal(fp.bev_intrinsics.numpy(), g["bev_intrinsics"])
    assert np.array_equal(fp.dx.numpy(), g["dx"]) and np.array_equal(fp.bx.numpy(), g["bx"])
    assert np.array_equal(fp.nx.numpy(), g["nx"]) and np.array_equal(fp.ds.numpy(), g["ds"])
    assert fp.nx_constant == [48, 48, 1] and fp.D == 37
    # geometry + cell ids are torch ops: check them on CPU against the reference's cells
    feat, depth, mask, K, E = synth.bev_inputs(1, 2, 3000)
    fp.initialize_frustum(synth.lift(feat, depth))
    geom = fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None])
    cell = fp.cell_ids(geom, torch.zeros(0))[0].numpy()
    ref = g["cells"].astype(np.int64)
    inb = (ref[:, 0] >= 0) & (ref[:, 0] < 48) & (ref[:, 1] >= 0) & (ref[:, 1] < 48) & (ref[:, 2] == 0)
    want = np.where(inb, ref[:, 1] * 48 + ref[:, 0], -1)
    assert np.array_equal(cell, want)
    assert (cell >= 0).sum() > 100000     # ~72 % of 153 920 in bounds


def test_shard_frames():
    for n, w in ((96, 8), (97, 8), (5, 8), (768, 4)):
        seen = []
        for r in range(w):
            seen += muvo_b200.shard_frames(n, r, w)
        assert sorted(seen) == list(range(n))
        rr = sorted(sum((muvo_b200.shard_frames(n, r, w, contiguous=False) for r in range(w)), []))
        assert rr == list(range(n))
    with pytest.raises(ValueError):
        muvo_b200.shard_frames(4, 4, 4)


def test_patch_and_unpatch_swap_reference_symbols():
    fake_dp = types.ModuleType("data_preprocessing")
    fake_dp.voxel_filter = lambda *a: "ref"
    fake_dp2 = types.ModuleType("data.data_preprocessing")      # the name muvo/data/dataset.py:16 imports it under
    fake_dp2.voxel_filter = fake_dp.voxel_filter
    fake_gv = types.ModuleType("data.generate_voxels")          # star-import copy of generate_voxels.py:16
    fake_gv.voxel_filter = fake_dp.voxel_filter
    fake_gv.voxelize_one = lambda *a: "ref"
    fake_m = types.ModuleType("muvo.metrics")
    fake_m.SSCMetrics = object
    fake_g = types.ModuleType("muvo.utils.geometry_utils")

    class RefPC:
        def do_range_projection(self, p, s):
            return "ref"
    fake_g.PointCloud = RefPC
    names = {"data_preprocessing": fake_dp, "data.data_preprocessing": fake_dp2, "data.generate_voxels": fake_gv,
             "muvo.metrics": fake_m, "muvo.utils.geometry_utils": fake_g}
    sys.modules.update(names)
    try:
        n = muvo_b200.patch()
        assert n == 6
        assert muvo_b200.patch() == 0                            # idempotent
        assert fake_dp.voxel_filter is muvo_b200.voxel_filter and fake_dp2.voxel_filter is muvo_b200.voxel_filter
        assert fake_gv.voxel_filter is muvo_b200.voxel_filter and fake_gv.voxelize_one is muvo_b200.points.voxelize_one
        assert fake_m.SSCMetrics is muvo_b200.SSCMetrics
        # "auto": a dispatcher that keeps the reference's NumPy method for DataLoader workers / forked children
        disp = RefPC.do_range_projection
        assert disp is not muvo_b200.points.do_range_projection and disp.__wrapped__(RefPC(), 0, 0) == "ref"
        P = sys.modules["muvo_b200.patch"]
        saved = P._in_worker_or_bad_fork
        P._in_worker_or_bad_fork = lambda: True
        try:
            assert RefPC().do_range_projection(0, 0) == "ref"
        finally:
            P._in_worker_or_bad_fork = saved
        muvo_b200.unpatch()
        assert fake_dp.voxel_filter() == "ref" and fake_m.SSCMetrics is object and RefPC().do_range_projection(0, 0) == "ref"
        assert muvo_b200.patch(range_projection="always") == 6
        assert RefPC.do_range_projection is muvo_b200.points.do_range_projection
        muvo_b200.unpatch()
        assert muvo_b200.patch(range_projection="never") == 5 and RefPC().do_range_projection(0, 0) == "ref"
    finally:
        muvo_b200.unpatch()
        for k in names:
            sys.modules.pop(k, None)


class _RangeDataset(torch.utils.data.Dataset):
    """Calls do_range_projection per sample, like muvo/data/dataset.py:300."""

    def __init__(self, pc):
        self.pc = pc

    def __len__(self):
        return 4

    def __getitem__(self, i):
        pts, sem = synth.carla_lidar_frame(2000, 7000 + i)
        d, x, s = self.pc.do_range_projection(pts, sem)
        return torch.from_numpy(d), torch.from_numpy(s)


def test_patched_range_projection_in_dataloader_workers():
    """the reference runs do_range_projection in forked DataLoader workers (N_WORKERS > 0), where CUDA cannot
    be used; the default patch must keep working there (reference NumPy method), not crash."""
    ref_import = pytest.importorskip("oracle.ref_import")
    if not ref_import.available():
        pytest.skip("reference checkout not mounted")
    gu = ref_import.load().geometry_utils
    pc = gu.PointCloud(64, 1024, -30, 10, [1.0, 0.0, 2.0])
    want = [_RangeDataset(pc)[i] for i in range(4)]
    sys.modules.setdefault("muvo.utils.geometry_utils", gu)
    try:
        muvo_b200.patch()
        assert hasattr(gu.PointCloud.do_range_projection, "__wrapped__")
        dl = torch.utils.data.DataLoader(_RangeDataset(pc), batch_size=1, num_workers=2, shuffle=False)
        got = [(d[0], s[0]) for d, s in dl]
        for (d0, s0), (d1, s1) in zip(want, got):
            assert torch.equal(d0, d1) and torch.equal(s0, s1)
    finally:
        muvo_b200.unpatch()


def test_synth_generators_are_deterministic_and_in_contract():
    p1, s1 = synth.carla_lidar_frame(5000, 1000)
    p2, s2 = synth.carla_lidar_frame(5000, 1000)
    assert np.array_equal(p1, p2) and np.array_equal(s1, s2)
    assert p1.dtype == np.float32 and s1.dtype == np.uint8 and p1.shape == (5000, 3)
    assert len(np.unique(p1, axis=0)) == 5000                                   # no duplicate points
    assert np.all(np.linalg.norm(p1 - np.float32([1, 0, 2]), axis=1) > 0.4)      # none at the sensor origin
    assert set(np.unique(s1)) <= {1, 6, 7, 10}
    pts, sem, off = synth.lidar_batch(5, 1000, 2000, 2000)
    assert off[0] == 0 and off[-1] == pts.shape[0] == sem.shape[0] and np.all(np.diff(off) >= 1000)
    assert synth.label_remap256()[255] == 0 and synth.label_remap256()[13] == 0 and synth.label_remap256()[7] == 1
    yp, yt = synth.occupancy_pair(1, 2, 1, size=(16, 16, 8))
    assert yp.dtype == np.int64 and yt.dtype == np.uint8 and set(np.unique(yt)) <= {0, 1, 255}


def test_sscmetrics_accumulate_arithmetic_matches_reference(golden):
    """The accumulate/compute/get_stats half is pure host code: feed it golden counts, compare golden stats."""
    g = golden("ssc.npz")
    for C in (2, 9):
        yp, yt = g[f"c{C}_pred"].astype(np.int64), g[f"c{C}_true"]
        m = muvo_b200.SSCMetrics(C)
        m._accumulate(torch.from_numpy(O.ssc_add_batch_counts(yp, yt, C)))
        m._accumulate(torch.from_numpy(O.ssc_add_batch_counts(yp, yt, C, g[f"c{C}_nonempty"], g[f"c{C}_nonsurface"])))
        st = m.get_stats()
        assert st["iou"] == float(g[f"c{C}_iou"]) and st["recall"] == float(g[f"c{C}_recall"])
        assert np.array_equal(st["iou_ssc"].numpy(), g[f"c{C}_iou_ssc"])
        assert np.array_equal(st["iou_ssc_mean"].numpy(), g[f"c{C}_iou_ssc_mean"])
        acc = np.r_[m.completion_tp, m.completion_fp, m.completion_fn, m.tps.numpy(), m.fps.numpy(), m.fns.numpy()]
        assert np.array_equal(acc.astype(np.float64), g[f"c{C}_acc2"])

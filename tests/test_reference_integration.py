"""``muvo_b200.patch()`` against the REAL reference modules, driven through the reference's own call sites.

Runs wherever an unmodified checkout resolves (``$MUVO_REFERENCE_ROOT``, ``/root/reference`` or the ``baseline/_ref`` copy that
``__graft_entry__.build()`` leaves next to the repo -- that one travels to the GPU box) AND a CUDA device is present; skipped
otherwise.  Before ``patch()`` the reference's functions produce the expected values on the CPU; after it the SAME module
attributes (``data_preprocessing.voxel_filter``, the star-imported names inside ``data/generate_voxels.py``:16,
``PointCloud.do_range_projection``, ``muvo.models.frustum_pooling.FrustumPooling`` as ``mile.py:37-43`` constructs it,
``muvo.metrics.SSCMetrics``) run on the B200 and must return the same results; ``unpatch()`` restores every symbol.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import pytest
import torch

import muvo_b200
from muvo_b200 import synth
from oracle import ref_import

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_import.available():
        pytest.skip("no reference checkout (MUVO_REFERENCE_ROOT / /root/reference / baseline/_ref)")
    ns = ref_import.load()
    for name in ("hydra", "omegaconf", "clearml"):                 # only data/generate_voxels.py's CLI wrapper needs them
        sys.modules.setdefault(name, MagicMock())
    import generate_voxels                                         # data/generate_voxels.py: `from data_preprocessing import *`
    ns.generate_voxels = generate_voxels
    yield ns
    muvo_b200.unpatch()


def _files(tmp_path):
    import cv2
    img = synth.carla_depth_image(7000, h=150, w=240)
    pts, sem = synth.carla_lidar_frame(8000, 7001)
    lid = pts.copy(); lid[:, 1] *= -1; lid -= np.float32([1, 0, 2])
    depth_file, lidar_file = str(tmp_path / "depth_semantic_000000001.png"), str(tmp_path / "points_semantic_000000001.npy")
    assert cv2.imwrite(depth_file, img)
    np.save(lidar_file, {"points_xyz": lid, "ObjTag": sem}, allow_pickle=True)
    return depth_file, lidar_file


CFG = types.SimpleNamespace(camera_position=[1.0, 0.0, 2.0], lidar_position=[1.0, 0.0, 2.0], fov=110, bev_offset_forward=0,
                            bev_resolution=0.2, offset_z=-20, voxel_resolution=0.5, voxel_size=[192, 192, 64])


def test_patch_swaps_the_real_modules_and_results_match(ref, lib, tmp_path):
    muvo_b200.unpatch()
    dp, gv, gu, fpm, mt = ref.data_preprocessing, ref.generate_voxels, ref.geometry_utils, ref.frustum_pooling, ref.metrics
    assert gv.voxel_filter is dp.voxel_filter and not dp.voxel_filter.__module__.startswith("muvo_b200")
    # ---- expected values: the unpatched reference on the CPU
    pts, sem = synth.carla_lidar_frame(20000, 7100)
    want_v, want_l = dp.voxel_filter(pts.copy(), sem, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    depth_file, lidar_file = _files(tmp_path)
    pipe = types.SimpleNamespace(sent=[], send=lambda m: pipe.sent.append(m))
    gv.voxelize_one(depth_file, lidar_file, CFG, str(tmp_path / "ref_voxel.npy"), pipe)
    want_file = np.load(str(tmp_path / "ref_voxel.npy"))
    want_rv = gu.PointCloud(64, 1024, -30, 10, [1.0, 0.0, 2.0]).do_range_projection(pts.copy(), sem)
    feat, depth, mask, K, E = synth.bev_inputs(2, 8, 7200)
    ref_fp = fpm.FrustumPooling(**synth.BEV_POOL_ARGS)                       # mile.py:37-43
    x_cpu = synth.lift(feat, depth)
    want_bev = ref_fp(x_cpu, K[:, None], E[:, None], mask)
    state = ref_fp.state_dict()
    yp, yt = synth.occupancy_pair(2, 9, 7300, size=(48, 48, 16))
    m_ref = mt.SSCMetrics(9)
    m_ref.add_batch(torch.from_numpy(yp), torch.from_numpy(yt))
    want_stats = m_ref.get_stats()
    # ---- patch the imported reference
    n = muvo_b200.patch()
    assert n >= 8
    assert muvo_b200.patch() == 0                                            # idempotent
    assert dp.voxel_filter is muvo_b200.points.voxel_filter and gv.voxel_filter is muvo_b200.points.voxel_filter   # star import too
    assert gv.merge_pcd is muvo_b200.points.merge_pcd and gv.voxelize_one is muvo_b200.points.voxelize_one
    assert fpm.FrustumPooling is muvo_b200.FrustumPooling and mt.SSCMetrics is muvo_b200.SSCMetrics
    # (a) through the reference's module attribute
    got_v, got_l = dp.voxel_filter(pts.copy(), sem, 0.5, [192, 192, 64], [0.0, 0, -10.0])
    assert np.array_equal(got_v, want_v) and np.array_equal(got_l, want_l)
    # (a) through data/generate_voxels.py's own entry point (files in, file out, progress pipe)
    pipe.sent.clear()
    gv.voxelize_one(depth_file, lidar_file, CFG, str(tmp_path / "our_voxel.npy"), pipe)
    assert np.array_equal(np.load(str(tmp_path / "our_voxel.npy")), want_file) and pipe.sent == [['x']]
    # (b) the reference's class, our method (main process -> GPU kernel)
    pc = gu.PointCloud(64, 1024, -30, 10, [1.0, 0.0, 2.0])
    got_rv = pc.do_range_projection(pts.copy(), sem)
    for a, b in zip(got_rv, want_rv):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    # (c) constructed as mile.py does, the reference's checkpoint loads strictly, same pooled features
    our_fp = fpm.FrustumPooling(**synth.BEV_POOL_ARGS)
    assert isinstance(our_fp, muvo_b200.FrustumPooling)
    our_fp.load_state_dict(state, strict=True)
    our_fp = our_fp.cuda()
    got_bev = our_fp(x_cpu.cuda(), K.cuda()[:, None], E.cuda()[:, None], mask.cuda())
    assert got_bev.shape == want_bev.shape
    assert (got_bev.cpu() - want_bev).abs().max() <= 1e-5 * want_bev.abs().max()
    assert torch.equal(got_bev.cpu() == 0, want_bev == 0)
    # (d) same statistics from the same calls
    m = mt.SSCMetrics(9)
    m.add_batch(torch.from_numpy(yp).cuda(), torch.from_numpy(yt).cuda())
    st = m.get_stats()
    assert st["iou"] == want_stats["iou"] and st["precision"] == want_stats["precision"] and st["recall"] == want_stats["recall"]
    assert torch.equal(st["iou_ssc"], want_stats["iou_ssc"]) and st["iou_ssc_mean"] == want_stats["iou_ssc_mean"]
    # ---- and back
    muvo_b200.unpatch()
    assert not dp.voxel_filter.__module__.startswith("muvo_b200") and gv.voxel_filter is dp.voxel_filter
    assert fpm.FrustumPooling is not muvo_b200.FrustumPooling and mt.SSCMetrics is not muvo_b200.SSCMetrics

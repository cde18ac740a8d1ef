"""Oracle restatement vs the LIVE reference (only where /root/reference is mounted)."""
import numpy as np
import pytest
import torch

import oracle as O
from oracle import ref_import
from muvo_b200 import synth

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def R():
    return ref_import.load()


def test_voxel_filter_matches_reference(R):
    for seed, n in ((1000, 30000), (1001, 5000)):
        p, s = synth.carla_lidar_frame(n, seed)
        v0, l0 = R.voxel_filter(p.copy(), s, 0.5, [192, 192, 64], [0.0, 0, -10.0])
        for fn in (O.voxel_filter_loop, O.voxel_filter_fast):
            v, l = fn(p, s, 0.5, [192, 192, 64], [0.0, 0, -10.0])
            assert np.array_equal(v, v0) and np.array_equal(l, l0)


def test_range_projection_matches_reference(R):
    pc = R.PointCloud(64, 1024, -30, 10, [1.0, 0.0, 2.0])
    p, s = synth.carla_lidar_frame(40000, 2001)
    ref = pc.do_range_projection(p.copy(), s)
    got = O.range_projection(p, s, lidar_position=[1.0, 0.0, 2.0])
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def test_frustum_pooling_matches_reference(R):
    feat, depth, mask, K, E = synth.bev_inputs(2, 8, 3001)
    fp = R.FrustumPooling(**synth.BEV_POOL_ARGS)
    fp.eval()
    ref = fp(synth.lift(feat, depth), K[:, None], E[:, None], mask)
    got = O.frustum_pooling_forward(synth.lift(feat, depth), K[:, None], E[:, None], mask, **synth.BEV_POOL_ARGS)
    assert torch.equal(ref, got)


def test_ssc_matches_reference(R):
    yp, yt = synth.occupancy_pair(2, 9, 4001, size=(48, 48, 16))
    m = R.SSCMetrics(9)
    tp, tt = torch.from_numpy(yp), torch.from_numpy(yt)
    a = m.get_score_completion(tp, tt)
    b = m.get_score_semantic_and_completion(tp, tt)
    ref = np.r_[list(a), b[0].numpy(), b[1].numpy(), b[2].numpy()]
    assert np.array_equal(O.ssc_counts(yp, yt, 9), ref)


def test_scal_losses_match_reference(R):
    """N4: float64 restatement vs the reference's fp32 SemScalLoss / GeoScalLoss (losses.py:191-287) on a fresh input."""
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)
    g = torch.Generator().manual_seed(99)
    pred = torch.randn((2, 2, 4, 10, 8, 6), generator=g) * 2
    tgt = torch.randint(0, 4, (2, 2, 10, 8, 6), generator=g).to(torch.uint8)
    tgt[torch.rand(tgt.shape, generator=g) < 0.03] = 255
    sem, geo = R.SemScalLoss()(pred, tgt).item(), R.GeoScalLoss()(pred, tgt).item()
    assert abs(O.sem_scal_loss(pred.numpy(), tgt.numpy()) - sem) <= 2e-6 * abs(sem)
    assert abs(O.geo_scal_loss(pred.numpy(), tgt.numpy()) - geo) <= 2e-6 * abs(geo)


def test_merge_pcd_matches_reference(R, tmp_path):
    """N1: merge_pcd restatement vs the reference's file-based function (data_preprocessing.py:125-139)."""
    import cv2
    img = synth.carla_depth_image(5100, h=60, w=96)
    pts, sem = synth.carla_lidar_frame(3000, 5101)
    lid = pts.copy(); lid[:, 1] *= -1; lid -= np.float32([1, 0, 2])
    cv2.imwrite(str(tmp_path / "d.png"), img)
    np.save(str(tmp_path / "l.npy"), {"points_xyz": lid.copy(), "ObjTag": sem}, allow_pickle=True)
    pcd, s = R.merge_pcd(str(tmp_path / "d.png"), str(tmp_path / "l.npy"), [1.0, 0.0, 2.0], [1.0, 0.0, 2.0], fov=110)
    p0, s0 = O.merge_pcd_arrays(img, lid, sem, [1.0, 0.0, 2.0], [1.0, 0.0, 2.0], fov=110)
    assert np.array_equal(pcd, p0) and np.array_equal(np.asarray(s).reshape(-1), np.asarray(s0).reshape(-1))

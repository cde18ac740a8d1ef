"""GPU parity for stages (a) voxelisation and (b) range-view projection, through the reference-facing
wrappers (which call the C ABI).  Bit-exact against the golden vectors and the CPU oracle."""
import os
import numpy as np
import pytest
import torch

import muvo_b200
import oracle as O
from muvo_b200 import synth
from muvo_b200.points import GridSpec, RangeSpec, PointCloud, sensor_to_grid

pytestmark = pytest.mark.gpu
GRID = (0.5, [192, 192, 64], [0.0, 0, -10.0])
LIDAR = [1.0, 0.0, 2.0]


def dev():
    return torch.device("cuda", 0)


# ------------------------------------------------------------------ (a)
def test_voxel_known_answers(golden, lib):
    g = golden("voxel.npz")
    v, l = muvo_b200.voxel_filter(g["known_pts"], g["known_sem"], *GRID)
    assert v.dtype == np.uint16 and l.dtype == np.uint8
    assert v.tolist() == [[0, 0, 0], [96, 96, 12], [98, 96, 12], [191, 191, 63]] and l.tolist() == [1, 7, 6, 2]


def test_voxel_golden(golden, lib):
    g = golden("voxel.npz")
    v, l = muvo_b200.voxel_filter(g["pts32"], g["sem32"], *GRID)
    assert np.array_equal(v, g["vox32"]) and np.array_equal(l, g["lab32"])
    v, l = muvo_b200.voxel_filter(g["pts64"], g["sem64"], *GRID)          # float64 cloud, (N,1) semantics
    assert np.array_equal(v, g["vox64"]) and np.array_equal(l, g["lab64"])
    v, l = muvo_b200.voxel_filter(g["pts32"], g["sem32"], 0.25, [96, 128, 32], [2.0, 0, -1.0])
    assert np.array_equal(v, g["vox_alt"]) and np.array_equal(l, g["lab_alt"])
    v, l = muvo_b200.voxel_filter(g["pts64"], g["sem64"], 0.2, [200, 200, 40], [0.0, 0, -1.0])   # general np.divmod path
    assert np.array_equal(v, g["vox_np2"]) and np.array_equal(l, g["lab_np2"])


@pytest.mark.parametrize("n,seed", [(100000, 1000), (1, 1001), (33, 1002), (4097, 1003)])
def test_voxel_vs_oracle(n, seed, lib):
    p, s = synth.carla_lidar_frame(n, seed)
    v0, l0 = O.voxel_filter_fast(p, s, *GRID)
    v, l = muvo_b200.voxel_filter(p, s, *GRID)
    assert np.array_equal(v, v0) and np.array_equal(l, l0)
    arr = muvo_b200.voxelize_one_array(p, s, *GRID)
    assert arr.dtype == np.uint16 and arr.shape == (len(v0), 4)


def test_voxel_edge_cases(lib):
    v, l = muvo_b200.voxel_filter(np.zeros((0, 3), np.float32), np.zeros((0,), np.uint8), *GRID)
    assert v.shape == (0, 3) and l.shape == (0,)
    v, l = muvo_b200.voxel_filter(np.full((7, 3), 1000.0, np.float32), np.ones(7, np.uint8), *GRID)
    assert v.shape == (0, 3)
    # exact ties -> lowest original index; a roadline point anywhere in the voxel overrides
    p = np.array([[0.1, 0.1, 0.1]] * 5, np.float32)
    _, l = muvo_b200.voxel_filter(p, np.array([3, 4, 5, 9, 9], np.uint8), *GRID)
    assert l.tolist() == [3]
    _, l = muvo_b200.voxel_filter(p, np.array([3, 4, 5, 6, 9], np.uint8), *GRID)
    assert l.tolist() == [6]
    # many points in one voxel (contended resolve) + unaligned tail
    rng = np.random.default_rng(0)
    p = (rng.random((5003, 3)) * 0.5).astype(np.float32)
    s = rng.integers(0, 23, 5003).astype(np.uint8)
    s[s == 6] = 7
    v0, l0 = O.voxel_filter_fast(p, s, *GRID)
    v, l = muvo_b200.voxel_filter(p, s, *GRID)
    assert len(v0) == 1 and np.array_equal(v, v0) and np.array_equal(l, l0)
    # boundaries: lower inclusive, upper exclusive
    p = np.array([[-48, -48, -6], [48, 0, 0], [47.999996, 47.999996, 25.999998], [0, 0, 26]], np.float32)
    v0, l0 = O.voxel_filter_fast(p, np.array([1, 2, 3, 4], np.uint8), *GRID)
    v, l = muvo_b200.voxel_filter(p, np.array([1, 2, 3, 4], np.uint8), *GRID)
    assert np.array_equal(v, v0) and np.array_equal(l, l0) and len(v) == 2


def _batch(F, nmin, nmax, seed):
    pts, sem, off = synth.lidar_batch(F, nmin, nmax, seed)
    return pts, sem, off


def test_batched_dense_sparse_and_range_vs_oracle(lib):
    pts, sem, off = _batch(5, 3000, 9000, 2100)
    # insert an empty frame in the middle
    off = np.r_[off[:3], off[2], off[3:]]
    F = len(off) - 1
    remap = synth.label_remap256()
    r = sensor_to_grid(torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev()), off, grid=GridSpec(),
                       range_spec=RangeSpec(lidar_position=tuple(LIDAR)), dense=True, sparse=True,
                       remap=torch.from_numpy(remap), layout="hwc", want_diag=True)
    torch.cuda.synchronize()
    dense, sparse, nocc = r["voxel"].cpu().numpy(), r["voxel_sparse"].cpu().numpy().view(np.uint16), r["n_occ"].cpu().numpy()
    n_in = 0
    for f in range(F):
        p, s = pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]]
        v0, l0 = O.voxel_filter_fast(p, s, *GRID)
        assert nocc[f] == len(v0)
        rows = sparse[off[f]:off[f] + nocc[f]]
        assert np.array_equal(rows[:, :3], v0) and np.array_equal(rows[:, 3].astype(np.uint8), l0)
        want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), (192, 192, 64), remap)
        assert np.array_equal(dense[f], want)
        d0, x0, s0 = O.range_projection(p, s, lidar_position=LIDAR)
        assert np.array_equal(r["range_depth"][f].cpu().numpy(), d0)
        assert np.array_equal(r["range_xyz"][f].cpu().numpy(), x0)
        assert np.array_equal(r["range_sem"][f].cpu().numpy(), s0)
        sh = p.astype(np.float64) + np.array([48.0, 48.0, 6.0])
        n_in += int(((sh >= 0) & (sh < np.array([96.0, 96.0, 32.0]))).all(1).sum())
    diag = r["diag"].cpu().numpy()
    assert diag[3] == n_in and diag[0] == 0


def test_dense_only_fast_path_and_workspace_self_cleaning(lib):
    """Dense-order bitmap path; repeated calls on the same (self-cleaning) workspace with different batches."""
    remap = torch.from_numpy(synth.label_remap256())
    outs = []
    for seed, F in ((2200, 4), (2300, 2), (2200, 4)):
        pts, sem, off = _batch(F, 2000, 6000, seed)
        r = sensor_to_grid(torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev()), off, grid=GridSpec(),
                           range_spec=RangeSpec(lidar_position=tuple(LIDAR)), remap=remap, layout="xyzd")
        torch.cuda.synchronize()
        for f in range(F):
            p, s = pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]]
            v0, l0 = O.voxel_filter_fast(p, s, *GRID)
            want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), (192, 192, 64), remap.numpy())
            assert np.array_equal(r["voxel"][f].cpu().numpy(), want)
            d0, x0, s0 = O.range_projection(p, s, lidar_position=LIDAR)
            assert np.array_equal(r["range_xyzd"][f].cpu().numpy(), O.pack_range_view(d0, x0))
            assert np.array_equal(r["range_sem"][f].cpu().numpy(), s0)
        outs.append((r["voxel"].clone(), r["range_xyzd"].clone()))
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])     # deterministic


def test_raw_labels_without_remap_and_small_odd_grid(lib):
    """No remap (raw CARLA tags in the dense grid) and a grid whose size is not a multiple of 16/1024."""
    pts, sem, off = _batch(2, 4000, 5000, 2400)
    spec = GridSpec(0.5, (50, 30, 7), (1.0, 0, -1.0))
    r = sensor_to_grid(torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev()), off, grid=spec)
    torch.cuda.synchronize()
    for f in range(2):
        p, s = pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]]
        v0, l0 = O.voxel_filter_fast(p, s, 0.5, [50, 30, 7], [1.0, 0, -1.0])
        want = np.zeros((50, 30, 7), np.uint8)
        want[v0[:, 0], v0[:, 1], v0[:, 2]] = l0
        assert np.array_equal(r["voxel"][f].cpu().numpy(), want)
        assert int(r["n_occ"][f]) == len(v0)


# ------------------------------------------------------------------ (b)
def test_range_known_and_golden(golden, lib):
    g = golden("range.npz")
    pc = PointCloud(64, 1024, -30, 10, LIDAR)
    d, x, s = pc.do_range_projection(g["known_pts"], g["known_sem"])
    assert d.dtype == np.float32 and x.shape == (64, 1024, 3) and s.dtype == np.uint8
    assert np.array_equal(d, g["known_depth"]) and np.array_equal(x, g["known_xyz"]) and np.array_equal(s, g["known_semimg"])
    d, x, s = pc.do_range_projection(g["pts"], g["sem"])
    assert np.array_equal(d, g["depth"]) and np.array_equal(x, g["xyz"]) and np.array_equal(s, g["semimg"])
    d, x, s = pc.do_range_projection(g["dense_pts"], g["dense_sem"])
    assert np.array_equal(d, g["dense_depth"]) and np.array_equal(x, g["dense_xyz"]) and np.array_equal(s, g["dense_semimg"])
    pc2 = PointCloud(32, 256, -25, 3, [0.5, 0.25, 1.75])
    d, x, s = pc2.do_range_projection(g["pts"], g["sem"])
    assert np.array_equal(d, g["alt_depth"]) and np.array_equal(x, g["alt_xyz"]) and np.array_equal(s, g["alt_semimg"])


def test_range_axes_diagonals_and_signed_zero(lib):
    """Exact bin edges (axes / diagonals) and the sign of zero behind the sensor."""
    L = np.float32(LIDAR)
    dirs = []
    for a in (1.0, 2.5, 7.0):
        for sx, sy in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, 1), (-1, -1)):
            for z in (0.0, 0.5, -1.0):
                dirs.append((a * sx, a * sy, z))
    pts = (np.array(dirs, np.float32) + L).astype(np.float32)
    pts = np.concatenate([pts, np.float32([[-9, 0.0, 2], [-9, -0.0, 2]])])
    sem = (np.arange(len(pts)) % 23).astype(np.uint8)
    pc = PointCloud(64, 1024, -30, 10, LIDAR)
    d, x, s = pc.do_range_projection(pts, sem)
    d0, x0, s0 = O.range_projection(pts, sem, lidar_position=LIDAR)
    assert np.array_equal(d, d0) and np.array_equal(x, x0) and np.array_equal(s, s0)
    ph, pw, _ = O.range_projection_indices(np.float32([[-9, 0.0, 2], [-9, -0.0, 2]]), lidar_position=LIDAR)
    assert pw.tolist() == [0, 1023]
    assert pc.last_diag[1] > 0          # points exactly on a column edge are reported


def test_range_random_directions_vs_oracle(lib):
    rng = np.random.default_rng(11)
    pts = (rng.normal(0, 1, (200000, 3)) * np.array([30, 30, 4]) + np.array(LIDAR)).astype(np.float32)
    sem = rng.integers(0, 23, len(pts)).astype(np.uint8)
    pc = PointCloud(64, 1024, -30, 10, LIDAR)
    d, x, s = pc.do_range_projection(pts, sem)
    d0, x0, s0 = O.range_projection(pts, sem, lidar_position=LIDAR)
    assert np.array_equal(d, d0) and np.array_equal(x, x0) and np.array_equal(s, s0)
    assert pc.last_diag[0] == 0


def test_range_ties_and_origin(lib):
    pc = PointCloud(64, 1024, -30, 10, LIDAR)
    # exact duplicates: lowest index wins
    pts = np.float32([[11, 0, 2]] * 4 + [[6, 0, 2]] * 3)
    d, x, s = pc.do_range_projection(pts, np.uint8([1, 2, 3, 4, 5, 6, 7]))
    assert s[16, 512] == 5 and d[16, 512] == 5.0
    # equal depth from mirrored points lands in different pixels
    d0, x0, s0 = O.range_projection(pts, np.uint8([1, 2, 3, 4, 5, 6, 7]), lidar_position=LIDAR)
    assert np.array_equal(s, s0) and np.array_equal(d, d0)
    with pytest.raises(IndexError):
        pc.do_range_projection(np.float32([[1, 0, 2], [5, 5, 5]]), np.uint8([1, 2]))     # a point AT the sensor
    # empty input -> all-empty images
    d, x, s = pc.do_range_projection(np.zeros((0, 3), np.float32), np.zeros((0,), np.uint8))
    assert np.all(d == -1) and np.all(x == 0) and np.all(s == 0)


def test_error_paths(lib):
    with pytest.raises(TypeError):
        PointCloud().do_range_projection(np.ones((3, 3), np.float64), np.zeros(3, np.uint8))
    with pytest.raises(ValueError):
        muvo_b200.voxel_filter(np.ones((3, 2), np.float32), np.zeros(3, np.uint8), *GRID)
    with pytest.raises(muvo_b200.MuvoError):     # > 65535 voxels per axis cannot be uint16 coordinates
        muvo_b200.voxel_filter(np.ones((3, 3), np.float32), np.zeros(3, np.uint8), 0.5, [70000, 2, 2], [0.0, 0, 0])


# ------------------------------------------------------------------ full BASELINE size: properties
def test_full_size_batch_properties(lib):
    """cfg2 shape (8 x 12 frames, 60-100 k points each): size-independent invariants + oracle on a few frames."""
    F = 96
    pts, sem, off = synth.lidar_batch(F, 60000, 100000, 2000)
    tp, ts = torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev())
    r = sensor_to_grid(tp, ts, off, grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), layout="xyzd",
                       want_diag=True)
    r2 = sensor_to_grid(tp, ts, off, grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), layout="xyzd")
    torch.cuda.synchronize()
    vox, xyzd, rsem = r["voxel"], r["range_xyzd"], r["range_sem"]
    assert torch.equal(vox, r2["voxel"]) and torch.equal(xyzd, r2["range_xyzd"]) and torch.equal(rsem, r2["range_sem"])
    # (1) raw labels are never 0 in this generator -> occupied voxel count == n_occ
    assert torch.equal((vox != 0).flatten(1).sum(1), r["n_occ"])
    # (2) independent float64 voxel ids in torch: same set of occupied voxels per frame
    sh = tp.double() + torch.tensor([48.0, 48.0, 6.0], device=dev(), dtype=torch.float64)
    inside = ((sh >= 0) & (sh < torch.tensor([96.0, 96.0, 32.0], device=dev(), dtype=torch.float64))).all(1)
    ijk = torch.floor(sh * 2.0).long()
    frame = torch.bucketize(torch.arange(len(pts), device=dev()), torch.from_numpy(off[1:]).to(dev()), right=True)
    lin = ((frame * 192 + ijk[:, 0]) * 192 + ijk[:, 1]) * 64 + ijk[:, 2]
    occ = torch.zeros(F * 192 * 192 * 64, dtype=torch.bool, device=dev())
    occ[lin[inside]] = True
    assert torch.equal(occ.view(F, 192, 192, 64), vox != 0)
    assert int(r["diag"][3]) == int(inside.sum())
    # (3) range image: stored depth == |xyz - lidar| of the stored point; empty pixels are (-1, 0, 0, 0)
    x, y, z, d = xyzd[:, 0], xyzd[:, 1], xyzd[:, 2], xyzd[:, 3]
    filled = d >= 0
    dd = torch.sqrt(((x.double() - 1.0) ** 2 + (-y.double() - 0.0) ** 2) + (z.double() - 2.0) ** 2).float()
    assert torch.equal(dd[filled], d[filled])
    assert torch.all(x[~filled] == 0) and torch.all(rsem[~filled] == 0) and torch.all(d[~filled] == -1)
    # (4) oracle on three frames
    for f in (0, 47, 95):
        p, s = pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]]
        v0, l0 = O.voxel_filter_fast(p, s, *GRID)
        want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), (192, 192, 64))
        assert np.array_equal(vox[f].cpu().numpy(), want)
        d0, x0, s0 = O.range_projection(p, s, lidar_position=LIDAR)
        assert np.array_equal(xyzd[f].cpu().numpy(), O.pack_range_view(d0, x0))
        assert np.array_equal(rsem[f].cpu().numpy(), s0)


def test_cfg5_million_point_frames_vs_oracle(lib):
    """cfg5 frame size (64 beams x 15 625 azimuths = 1 M points per frame): three ragged frames, full oracle parity on
    every output (dense + sorted sparse voxels, range view) -- many points per voxel / pixel, long atomic chains."""
    sizes = [1_000_000, 999_999, 1_000_000]
    frames = [synth.carla_lidar_frame(n, 5000 + i) for i, n in enumerate(sizes)]
    pts = np.concatenate([f[0] for f in frames]); sem = np.concatenate([f[1] for f in frames])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    r = sensor_to_grid(torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev()), off, grid=GridSpec(),
                       range_spec=RangeSpec(lidar_position=tuple(LIDAR)), dense=True, sparse=True, layout="xyzd")
    torch.cuda.synchronize()
    for f, (p, sm) in enumerate(frames):
        v0, l0 = O.voxel_filter_fast(p, sm, *GRID)
        n = int(r["n_occ"][f])
        assert n == len(v0)
        rows = r["voxel_sparse"][off[f]:off[f] + n].cpu().numpy().view(np.uint16)
        assert np.array_equal(rows[:, :3], v0) and np.array_equal(rows[:, 3], l0.astype(np.uint16))
        want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), (192, 192, 64))
        assert np.array_equal(r["voxel"][f].cpu().numpy(), want)
        d0, x0, s0 = O.range_projection(p, sm, lidar_position=LIDAR)
        assert np.array_equal(r["range_xyzd"][f].cpu().numpy(), O.pack_range_view(d0, x0))
        assert np.array_equal(r["range_sem"][f].cpu().numpy(), s0)


# ------------------------------------------------------------------ near-ties: the packed atomicMax + exact tie protocol
def test_voxel_near_ties_resolve_exactly(lib):
    """Thousands of points whose |p mod res|^2 agree in the top 32 key bits (and many exact ties) in a handful
    of voxels: the fast packed order is ambiguous there, the tie protocol must reproduce the exact arg-min."""
    rng = np.random.default_rng(21)
    n = 6000
    base = np.float32([0.25, 0.25, 0.25]) + rng.integers(0, 3, (n, 1)).astype(np.float32) * np.float32(0.5)
    ulp = np.spacing(np.float32(0.25))
    pts = (base + rng.integers(-4, 5, (n, 3)).astype(np.float32) * ulp).astype(np.float32)
    sem = rng.integers(0, 23, n).astype(np.uint8)
    for with_road in (False, True):
        s = sem.copy()
        if not with_road:
            s[s == 6] = 7
        v0, l0 = O.voxel_filter_fast(pts, s, *GRID)
        v, l = muvo_b200.voxel_filter(pts, s, *GRID)
        assert len(v0) == 3 and np.array_equal(v, v0) and np.array_equal(l, l0)
    # all points identical: pure index tie-break under heavy contention
    same = np.repeat(np.float32([[3.3, -7.1, 0.2]]), 4000, 0)
    s = (np.arange(4000) % 5 + 1).astype(np.uint8)
    v, l = muvo_b200.voxel_filter(same, s, *GRID)
    assert l.tolist() == [1]


def test_range_near_ties_resolve_exactly(lib):
    rng = np.random.default_rng(22)
    n = 5000
    ulp = np.spacing(np.float32(11.0))
    pts = np.zeros((n, 3), np.float32)
    pts[:, 0] = np.float32(11.0) + rng.integers(-6, 7, n).astype(np.float32) * ulp
    pts[:, 2] = 2.0
    # a second bundle in another pixel, plus exact duplicates
    pts[n // 2:, 1] = np.float32(5.0)
    pts[n // 2:, 0] = np.float32(1.0) + rng.integers(-6, 7, n - n // 2).astype(np.float32) * np.spacing(np.float32(1.0))
    sem = rng.integers(1, 23, n).astype(np.uint8)
    pc = PointCloud(64, 1024, -30, 10, LIDAR)
    d, x, s = pc.do_range_projection(pts, sem)
    d0, x0, s0 = O.range_projection(pts, sem, lidar_position=LIDAR)
    assert np.array_equal(d, d0) and np.array_equal(x, x0) and np.array_equal(s, s0)
    assert (d >= 0).sum() <= 4


def test_f32_pixel_path_never_disagrees_with_float64(lib):
    """The f32 polynomial pixel path may only decide points that the float64 formula puts in the same pixel;
    everything near a bin edge must be deferred.  40 M points: random directions at all ranges, points generated ON
    column/row edges (+- a few f32 ulps), and tiny / huge magnitudes."""
    import ctypes as C
    rng = torch.Generator(device="cuda").manual_seed(5)
    d = dev()
    n = 10_000_000
    cfgs = [RangeSpec(), RangeSpec(H=128, W=2048, fov_down=-25, fov_up=15, lidar_position=(0.5, -0.25, 1.75))]
    for spec in cfgs:
        L = torch.tensor(spec.lidar_position, dtype=torch.float64, device=d)
        clouds = []
        # (1) isotropic directions, log-uniform range 0.05 .. 300 m
        v = torch.randn(n, 3, generator=rng, device=d, dtype=torch.float64)
        v /= v.norm(dim=1, keepdim=True)
        rr = torch.exp(torch.empty(n, 1, device=d, dtype=torch.float64).uniform_(np.log(0.05), np.log(300.0), generator=rng))
        clouds.append(v * rr)
        # (2) LiDAR-like: pitch inside the FOV, uniform azimuth
        az = torch.empty(n, device=d, dtype=torch.float64).uniform_(-np.pi, np.pi, generator=rng)
        pit = torch.empty(n, device=d, dtype=torch.float64).uniform_(np.deg2rad(spec.fov_down - 2), np.deg2rad(spec.fov_up + 2), generator=rng)
        r2 = torch.empty(n, device=d, dtype=torch.float64).uniform_(0.5, 100.0, generator=rng)
        clouds.append(torch.stack([r2 * pit.cos() * az.cos(), r2 * pit.cos() * az.sin(), r2 * pit.sin()], 1))
        # (3) directions exactly on column edges and row edges, perturbed by ~1e-7 relative
        kcol = torch.randint(0, spec.W + 1, (n,), generator=rng, device=d).double()
        yaw = np.pi * (1.0 - 2.0 * kcol / spec.W)
        krow = torch.randint(0, spec.H + 1, (n,), generator=rng, device=d).double()
        fov = np.deg2rad(spec.fov_up - spec.fov_down)
        pitch_e = (1.0 - krow / spec.H) * fov - abs(np.deg2rad(spec.fov_down))
        on_col = torch.stack([r2 * pit.cos() * yaw.cos(), r2 * pit.cos() * yaw.sin(), r2 * pit.sin()], 1)
        on_row = torch.stack([r2 * pitch_e.cos() * az.cos(), r2 * pitch_e.cos() * az.sin(), r2 * pitch_e.sin()], 1)
        jit = 1.0 + 2e-7 * torch.randn(n, 3, generator=rng, device=d, dtype=torch.float64)
        clouds.append(on_col * jit)
        clouds.append(on_row * jit)
        counts = torch.zeros(4, dtype=torch.int64, device=d)
        cfg = spec.to_c()
        for c in clouds:
            # LiDAR frame -> ego frame as the reference inverts it (geometry_utils.py:177-178): x + L0, -(y + L1), z + L2
            ego = torch.stack([c[:, 0] + L[0], -(c[:, 1] + L[1]), c[:, 2] + L[2]], 1).float().contiguous()
            rc = lib.muvo_debug_pixel_check(ego.data_ptr(), ego.shape[0], C.byref(cfg), counts.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream)
            assert rc == 0
        # extreme magnitudes must all be deferred or dropped, never decided in f32
        ext = torch.tensor([[1e-20, 1e-20, 1e-20], [1e19, 1e19, 1e19], [3e38, 0, 0], [1e-30, 0, 1e-10]], device=d).float()
        ext = (ext * torch.tensor([1.0, -1.0, 1.0], device=d) + L.float() * torch.tensor([1.0, -1.0, 1.0], device=d)).contiguous()
        c_ext = torch.zeros(4, dtype=torch.int64, device=d)
        assert lib.muvo_debug_pixel_check(ext.data_ptr(), ext.shape[0], C.byref(cfg), c_ext.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream) == 0
        assert c_ext[2].item() == 0
        n_fast, n_slow, n_bad, n_drop = counts.tolist()
        assert n_bad == 0
        assert n_fast + n_slow + n_drop == 4 * n
        # clouds (1)+(2) are generic: only a small fraction may take the float64 path
        assert n_slow < 0.6 * 4 * n and n_fast > 0.4 * 4 * n
    # generic clouds alone: deferral rate is 2*eps per axis (~0.4 %)
    counts = torch.zeros(4, dtype=torch.int64, device=d)
    spec = RangeSpec()
    cfg = spec.to_c()
    pts, _ = synth.carla_lidar_frame(200000, 77)
    ego = torch.from_numpy(pts).to(d)
    assert lib.muvo_debug_pixel_check(ego.data_ptr(), ego.shape[0], C.byref(cfg), counts.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream) == 0
    n_fast, n_slow, n_bad, n_drop = counts.tolist()
    assert n_bad == 0 and n_slow < 0.02 * 200000


def test_host_pipeline_packed_sparse_and_range(lib):
    """numpy in -> pinned staging -> kernels -> pinned host out: three batches through two slots; the second and third
    read back only ~1.25 x the rows the first one needed (packed sparse lists), one of them with MORE occupied voxels
    than that (top-up path)."""
    from muvo_b200.pipeline import HostPipeline
    pipe = HostPipeline(dev(), grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), dense=False, sparse=True,
                        layout="hwc", host_threads=3)
    batches = [_batch(3, 2000, 3000, 2200), _batch(4, 2000, 3000, 2300), _batch(4, 20000, 30000, 2400)]
    outs = []
    pipe.submit(*batches[0])
    outs.append({k: v.clone() for k, v in pipe.result().items()})     # first batch alone: establishes the row cap
    pipe.submit(*batches[1])
    pipe.submit(*batches[2])
    outs.append({k: v.clone() for k, v in pipe.result().items()})
    outs.append({k: v.clone() for k, v in pipe.result().items()})
    for (pts, sem, off), r in zip(batches, outs):
        F = len(off) - 1
        start, nocc = r["sparse_start"].numpy(), r["n_occ"].numpy()
        rows_all = r["voxel_sparse"].numpy().view(np.uint16)
        assert start[0] == 0 and np.array_equal(np.diff(start), nocc)
        for f in range(F):
            p, s = pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]]
            v0, l0 = O.voxel_filter_fast(p, s, *GRID)
            rows = rows_all[start[f]:start[f] + nocc[f]]
            assert nocc[f] == len(v0)
            assert np.array_equal(rows[:, :3], v0) and np.array_equal(rows[:, 3].astype(np.uint8), l0)
            d0, x0, s0 = O.range_projection(p, s, lidar_position=LIDAR)
            assert np.array_equal(r["range_depth"][f].numpy(), d0)
            assert np.array_equal(r["range_xyz"][f].numpy(), x0)
            assert np.array_equal(r["range_sem"][f].numpy(), s0)


def test_host_pipeline_pinned_inputs_skip_staging_and_errors_surface(lib):
    """Inputs that already live in pinned memory are the DMA source themselves; same results as from numpy.  Inconsistent
    frame_offsets are rejected by submit(); a failure on the worker thread is re-raised by result()."""
    from muvo_b200.pipeline import HostPipeline
    pipe = HostPipeline(dev(), grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), dense=False, sparse=True, layout="hwc")
    pts, sem, off = _batch(3, 2000, 3000, 2500)
    ppts, psem, poff = HostPipeline.pinned_inputs(len(pts), len(off) - 1)
    ppts.numpy()[...] = pts; psem.numpy()[...] = sem; poff.numpy()[...] = off
    assert ppts.is_pinned()
    pipe.submit(pts, sem, off)
    a = {k: v.clone() for k, v in pipe.result().items()}
    pipe.submit(ppts, psem, poff)
    b = {k: v.clone() for k, v in pipe.result().items()}
    assert a.keys() == b.keys()
    n = int(a["sparse_start"][-1])
    assert torch.equal(a["voxel_sparse"][:n], b["voxel_sparse"][:n])
    for k in a:
        if k != "voxel_sparse":
            assert torch.equal(a[k], b[k]), k
    bad = off.copy(); bad[-1] += 5
    with pytest.raises(ValueError):
        pipe.submit(pts, sem, bad)                                       # checked on the caller's thread
    pipe.submit(pts, np.array(["x"] * len(sem)), off)                    # fails on the worker thread (not convertible to uint8)
    with pytest.raises(Exception):
        pipe.result()
    pipe.submit(pts, sem, off)                                           # the pipeline is usable afterwards
    c = pipe.result()
    assert torch.equal(c["n_occ"], a["n_occ"])


def test_cuda_graph_replay_of_the_point_kernels(lib):
    """The four kernels are chained by programmatic dependent launch; capturing the call in a CUDA graph and replaying it on
    new input values (same buffers) must still give the oracle's result, twice in a row (workspace self-cleaning)."""
    pts, sem, off = _batch(4, 3000, 5000, 2600)
    pts2, sem2, _ = _batch(4, 3000, 5000, 2700)
    n = min(len(pts), len(pts2))
    off = np.array([0, n // 4, n // 2, 3 * n // 4, n], dtype=np.int64)
    tp, ts, to = torch.from_numpy(pts[:n].copy()).to(dev()), torch.from_numpy(sem[:n].copy()).to(dev()), torch.from_numpy(off).to(dev())
    out = {}

    def step():
        r = sensor_to_grid(tp, ts, to, grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), layout="xyzd", out=out)
        out.update({k: r[k] for k in ("voxel", "n_occ", "range_xyzd", "range_sem")})
        return r

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
    for src_p, src_s in ((pts2[:n], sem2[:n]), (pts[:n], sem[:n])):
        tp.copy_(torch.from_numpy(src_p.copy())); ts.copy_(torch.from_numpy(src_s.copy()))
        g.replay()
        torch.cuda.synchronize()
        for f in range(4):
            p, s_ = src_p[off[f]:off[f + 1]], src_s[off[f]:off[f + 1]]
            v0, l0 = O.voxel_filter_fast(p, s_, *GRID)
            want = O.densify_voxels(np.concatenate([v0, l0[:, None].astype(np.uint16)], 1), (192, 192, 64))
            assert np.array_equal(out["voxel"][f].cpu().numpy(), want)
            d0, x0, s0 = O.range_projection(p, s_, lidar_position=LIDAR)
            assert np.array_equal(out["range_xyzd"][f].cpu().numpy(), O.pack_range_view(d0, x0))
            assert np.array_equal(out["range_sem"][f].cpu().numpy(), s0)


# ------------------------------------------------------------------ N1: camera + LiDAR cloud in front of (a)
def test_merge_pcd_golden_and_full_size_vs_oracle(golden, lib):
    g = golden("merge.npz")
    pcd, sem = muvo_b200.merge_pcd_arrays(g["img"], g["lidar_xyz"], g["lidar_sem"], [1.0, 0.0, 2.0], [1.0, 0.0, 2.0], fov=110)
    assert pcd.dtype == np.float64 and sem.shape == (len(pcd), 1)
    assert np.array_equal(pcd, g["pcd"]) and np.array_equal(sem, g["sem"])
    pcd2, sem2 = muvo_b200.merge_pcd_arrays(g["img"], g["lidar_xyz"], g["lidar_sem"], [1.5, 0.25, 1.75], [1.0, 0.0, 2.0], fov=90,
                                            mask_ego=False)
    assert np.array_equal(pcd2, g["pcd_nomask"]) and np.array_equal(sem2, g["sem_nomask"])
    # merged float64 cloud straight into the voxeliser (device tensors, no host round trip) == the reference's chain
    xyz_d, sem_d = muvo_b200.merge_pcd_device(g["img"], g["lidar_xyz"], g["lidar_sem"], [1.0, 0.0, 2.0], [1.0, 0.0, 2.0])
    r = sensor_to_grid(xyz_d, sem_d, None, grid=GridSpec(), dense=False, sparse=True)
    n = int(r["n_occ"][0].item())
    rows = r["voxel_sparse"][:n].cpu().numpy().view(np.uint16)
    assert np.array_equal(rows[:, :3], g["vox"]) and np.array_equal(rows[:, 3].astype(np.uint8), g["lab"])
    # full CARLA image size (600 x 960) + a 60 k LiDAR sweep against the oracle
    img = synth.carla_depth_image(5100)
    pts, s = synth.carla_lidar_frame(60000, 5101)
    lid = pts.copy(); lid[:, 1] *= -1; lid -= np.float32([1, 0, 2])
    got_p, got_s = muvo_b200.merge_pcd_arrays(img, lid, s, [1.0, 0.0, 2.0], [1.0, 0.0, 2.0])
    want_p, want_s = O.merge_pcd_arrays(img, lid, s, [1.0, 0.0, 2.0], [1.0, 0.0, 2.0])
    assert np.array_equal(got_p, want_p) and np.array_equal(got_s, want_s)
    # empty inputs
    e_p, e_s = muvo_b200.merge_pcd_arrays(np.zeros((0, 0, 4), np.uint8), np.zeros((0, 3), np.float32), np.zeros((0,), np.uint8),
                                          [1.0, 0.0, 2.0], [1.0, 0.0, 2.0])
    assert e_p.shape == (0, 3) and e_s.shape == (0, 1)


def test_generate_voxels_cli_on_a_synthetic_run_directory(lib, tmp_path):
    """N3: the batch voxeliser walks <root>/**/Town*/*/, reads pd_dataframe.pkl and writes voxel/voxel_#########.npy in the
    format of voxelize_one; every file equals merge_pcd + voxel_filter of the oracle."""
    import cv2
    import pandas as pd
    from muvo_b200 import generate_voxels as gv
    run = tmp_path / "trainval" / "train" / "Town01" / "0000"
    (run / "depth_semantic").mkdir(parents=True)
    (run / "points_semantic").mkdir()
    rows, inputs = [], []
    for k in range(3):
        img = synth.carla_depth_image(5200 + k, h=120, w=200)
        pts, s = synth.carla_lidar_frame(5000, 5300 + k)
        lid = pts.copy(); lid[:, 1] *= -1; lid -= np.float32([1, 0, 2])
        name = f"{k:09d}"
        cv2.imwrite(str(run / "depth_semantic" / f"depth_semantic_{name}.png"), img)
        np.save(str(run / "points_semantic" / f"points_semantic_{name}.npy"), {"points_xyz": lid, "ObjTag": s}, allow_pickle=True)
        rows.append({"depth_semantic_path": f"depth_semantic/depth_semantic_{name}.png",
                     "points_semantic_path": f"points_semantic/points_semantic_{name}.npy"})
        inputs.append((img, lid, s))
    pd.DataFrame(rows).to_pickle(run / "pd_dataframe.pkl")
    assert gv.main(["--root", str(tmp_path), "--io-threads", "2"]) == 0
    frame = pd.read_pickle(run / "pd_dataframe.pkl")
    assert list(frame["voxel_path"]) == [f"voxel/voxel_{k:09d}.npy" for k in range(3)]
    for k, (img, lid, s) in enumerate(inputs):
        got = np.load(run / frame["voxel_path"][k])
        pcd, sem = O.merge_pcd_arrays(img, lid, s, [1.0, 0.0, 2.0], [1.0, 0.0, 2.0], fov=110)
        v, l = O.voxel_filter_fast(pcd, sem, 0.5, [192, 192, 64], [0 * 0.2, 0, -20 * 0.5])
        assert got.dtype == np.uint16 and np.array_equal(got, np.concatenate([v, l[:, None].astype(np.uint16)], 1))
    assert gv.main(["--root", str(tmp_path / "nothing_here")]) == 1


def test_label_pyramids_match_oracle(lib):
    """N1: range-view / voxel label pyramids behind the point kernels (PreProcess.forward, preprocess.py:151-186), bit-exact."""
    pts, sem, off = _batch(3, 20000, 30000, 2600)
    r = sensor_to_grid(torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev()), off, grid=GridSpec(),
                       range_spec=RangeSpec(lidar_position=tuple(LIDAR)), remap=torch.from_numpy(synth.label_remap256()), layout="xyzd")
    got = muvo_b200.label_pyramids(r["range_xyzd"], r["range_sem"], r["voxel"], scale=50.0)
    want = O.label_pyramids(r["range_xyzd"].cpu().numpy(), r["range_sem"].cpu().numpy(), r["voxel"].cpu().numpy(), scale=50.0)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    rng = np.random.default_rng(4)                      # sizes that are not multiples of 4
    xyzd = torch.from_numpy(rng.normal(0, 30, (2, 4, 10, 22)).astype(np.float32)).to(dev())
    vox = torch.from_numpy(rng.integers(0, 3, (2, 13, 9, 6)).astype(np.uint8)).to(dev())
    got = muvo_b200.label_pyramids(xyzd, None, vox, scale=50.0)
    want = O.label_pyramids(xyzd.cpu().numpy(), None, vox.cpu().numpy(), scale=50.0)
    for k in want:
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


# ----------------------------------------------------------------------------- N1: LiDAR-side prep fused into (b), densify
def test_lidar_range_view_fused_prep_matches_reference_golden(golden, lib):
    """muvo/data/dataset.py:275-305 for a raw sweep in one pass of the point kernels (convert_coor_lidar + LABEL_MAP remap +
    ego-box drop inside the kernels): bit-equal to the reference's own functions (tests/golden/lidar.npz), single frame and
    a ragged batch of three, both layouts; the caller's arrays are not modified."""
    from muvo_b200.points import lidar_range_view
    g = golden("lidar.npz")
    raw, tag = g["raw"].copy(), g["tag"].copy()
    r = lidar_range_view(raw, tag, lidar_position=(1.0, 0.0, 2.0), remap=g["remap"], layout="hwc")
    assert np.array_equal(raw, g["raw"]) and np.array_equal(tag, g["tag"])
    assert np.array_equal(r["range_depth"][0].cpu().numpy(), g["depth"])
    assert np.array_equal(r["range_xyz"][0].cpu().numpy(), g["xyz"])
    assert np.array_equal(r["range_sem"][0].cpu().numpy(), g["semimg"])
    n = len(raw)
    cut = [0, 5000, 5000 + 3000, n]
    rb = lidar_range_view(np.concatenate([raw[:5000], raw[:3000], raw]), np.concatenate([tag[:5000], tag[:3000], tag]),
                          lidar_position=(1.0, 0.0, 2.0), remap=g["remap"], frame_offsets=np.array([0, 5000, 8000, 8000 + n]))
    assert np.array_equal(rb["range_xyzd"][2].cpu().numpy(), g["xyzd"])
    for f, m in ((0, 5000), (1, 3000)):
        p, s = O.lidar_prep(raw[:m], tag[:m], [1.0, 0.0, 2.0], g["remap"])
        d, x, sm = O.range_projection(p, s, lidar_position=[1.0, 0.0, 2.0])
        assert np.array_equal(rb["range_xyzd"][f].cpu().numpy(), O.pack_range_view(d, x))
        assert np.array_equal(rb["range_sem"][f].cpu().numpy(), sm)
    # without the ego box / remap the fused path equals convert_coor_lidar + plain projection
    from muvo_b200.points import LidarPrep, RangeSpec, sensor_to_grid
    r2 = sensor_to_grid(torch.from_numpy(raw).cuda(), torch.from_numpy(tag).cuda(), None, range_spec=RangeSpec(lidar_position=(1.0, 0.0, 2.0)),
                        layout="hwc", lidar_prep=LidarPrep(ego_dimension=None))
    p, s = O.lidar_prep(raw, tag, [1.0, 0.0, 2.0], None, ego_dimension=None)
    d, x, sm = O.range_projection(p, s, lidar_position=[1.0, 0.0, 2.0])
    assert np.array_equal(r2["range_depth"][0].cpu().numpy(), d) and np.array_equal(r2["range_sem"][0].cpu().numpy(), sm)


def test_densify_voxels_matches_reference_golden_and_last_row_wins(golden, lib):
    """dataset.py:317-327 on the device: 255 -> 0, remap, duplicates resolved like numpy's fancy assignment, files batched."""
    from muvo_b200.points import densify_voxels
    g = golden("lidar.npz")
    vd = g["voxel_data"]
    want = np.zeros(192 * 192 * 64, np.uint8)
    want[g["voxels_nz_idx"]] = g["voxels_nz_val"]
    got = densify_voxels(vd, (192, 192, 64), g["remap"])
    assert got.shape == (1, 192, 192, 64) and np.array_equal(got[0].cpu().numpy().reshape(-1), want)
    rng = np.random.default_rng(9)
    rows = np.stack([rng.integers(0, 12, 4000), rng.integers(0, 10, 4000), rng.integers(0, 6, 4000), rng.integers(0, 23, 4000)], 1).astype(np.uint16)
    rows[rng.random(4000) < 0.05, 3] = 255                                    # 720 voxels, 4000 rows: heavy duplication
    off = np.array([0, 1500, 1500, 4000])
    gb = densify_voxels(rows, (12, 10, 6), None, frame_offsets=off)
    for f in range(3):
        assert np.array_equal(gb[f].cpu().numpy(), O.densify_voxels(rows[off[f]:off[f + 1]], (12, 10, 6), None))
    assert int(gb.n_bad.item()) == 0
    bad = rows.copy(); bad[7, 0] = 12
    assert int(densify_voxels(bad, (12, 10, 6), None).n_bad.item()) == 1


def test_voxelize_one_drop_in_matches_reference_golden(golden, lib, tmp_path):
    """``muvo_b200.points.voxelize_one`` (the replacement of data/generate_voxels.py:64-77, files in / file out) against the
    reference's merge_pcd + voxel_filter run on the same files (tests/golden/merge.npz)."""
    import types
    cv2 = pytest.importorskip("cv2")
    from muvo_b200.points import voxelize_one
    g = golden("merge.npz")
    img, lx, ls = g["img"], g["lidar_xyz"], g["lidar_sem"]
    depth_file, lidar_file = str(tmp_path / "d.png"), str(tmp_path / "l.npy")
    assert cv2.imwrite(depth_file, img)
    np.save(lidar_file, {"points_xyz": lx, "ObjTag": ls}, allow_pickle=True)
    cfg = types.SimpleNamespace(camera_position=[1.0, 0.0, 2.0], lidar_position=[1.0, 0.0, 2.0], fov=110, bev_offset_forward=0,
                                bev_resolution=0.2, offset_z=-20, voxel_resolution=0.5, voxel_size=[192, 192, 64])
    out = voxelize_one(depth_file, lidar_file, cfg, str(tmp_path / "v.npy"))
    want = np.concatenate([g["vox"], g["lab"][:, None].astype(np.uint16)], 1)      # the reference's merge_pcd + voxel_filter (:64-73)
    assert out.dtype == np.uint16 and np.array_equal(out, want)
    assert np.array_equal(np.load(str(tmp_path / "v.npy")), want)


def test_batched_voxelisation_equals_frame_by_frame(lib):
    """generate_voxels.voxelize_frames: N merges packed on the device (muvo_merge_pcd_at), one launch of the point kernels
    over the ragged batch, two host syncs -- the same (n,4) arrays as one frame at a time, for frames of different sizes."""
    from muvo_b200.generate_voxels import load_config, voxelize_frame, voxelize_frames
    cfg = load_config()
    frames = []
    for k, (h, w, n) in enumerate([(150, 240, 8000), (100, 160, 3000), (150, 240, 0), (60, 80, 12000)]):
        img = synth.carla_depth_image(7400 + k, h=h, w=w)
        pts, sem = synth.carla_lidar_frame(max(n, 1), 7410 + k)
        lid = pts.copy(); lid[:, 1] *= -1; lid -= np.float32([1, 0, 2])
        frames.append((img, lid[:n], sem[:n]))
    got = voxelize_frames(frames, cfg)
    assert len(got) == len(frames)
    for f, g in zip(frames, got):
        want = voxelize_frame(*f, cfg)
        assert g.dtype == np.uint16 and np.array_equal(g, want)
    assert voxelize_frames([], cfg) == []


def test_opt_in_dataflow_kernel_equals_default_path(lib):
    """The single-launch dataflow kernel (points_mega.cu, opt-in with tuning key 3 = 2: measured slower than the multi-launch
    path) produces the same grids / range views as the default path, repeatedly on a reused workspace,
    in scan order and with shuffled points (every pixel / voxel contested out of order)."""
    from muvo_b200 import _lib
    remap = torch.from_numpy(synth.label_remap256())
    for seed, F, shuffle in ((2700, 5, False), (2710, 3, True), (2700, 5, False)):
        pts, sem, off = _batch(F, 4000, 12000, seed)
        if shuffle:
            rng = np.random.default_rng(seed)
            for f in range(F):
                perm = rng.permutation(off[f + 1] - off[f]) + off[f]
                pts[off[f]:off[f + 1]], sem[off[f]:off[f + 1]] = pts[perm], sem[perm]
        tp, ts = torch.from_numpy(pts).to(dev()), torch.from_numpy(sem).to(dev())
        kw = dict(grid=GridSpec(), range_spec=RangeSpec(lidar_position=tuple(LIDAR)), remap=remap, layout="xyzd")
        want = sensor_to_grid(tp, ts, off, **kw)
        assert lib.muvo_debug_set_tuning(3, 2) == 0
        try:
            with _lib.profile(_lib.current_stream(tp.device)) as prof:
                got = sensor_to_grid(tp, ts, off, **kw)
        finally:
            assert lib.muvo_debug_set_tuning(3, 0) == 0
        assert "k_points_tile" not in [k for k, _ in prof.kernels], prof.kernels      # it really was the other kernel
        for k in ("voxel", "n_occ", "range_xyzd", "range_sem"):
            assert torch.equal(got[k], want[k]), k


def test_randomised_configurations_vs_oracle(lib):
    """tools/fuzz_points.py: 40 random (grid, resolution, offset, image shape, field of view, sensor position, ragged batch,
    point order, dtype, layout, remap) configurations of stages (a) + (b), every output bit-equal to the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_points.py"), "40", "5"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "40 cases, 0 mismatches" in p.stdout


def test_randomised_next_rows_vs_oracle(lib):
    """tools/fuzz_next_rows.py: 40 random shapes each of the scal-loss sums, PointPillar scatter, label pyramids, densify (with
    duplicate rows), fused argmax + counts, fused lift-splat, against the oracle / the unfused path."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_next_rows.py"), "40", "2"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "40 cases, 0 mismatches" in p.stdout


def test_randomised_host_paths_vs_oracle(lib):
    """tools/fuzz_host_paths.py: HostPipeline with back-to-back ragged batches of very different sizes (sparse / dense, both
    layouts, device_out, 1-3 slots), merge_pcd on random images / sweeps, lidar_range_view on ragged raw sweeps: bit-equal."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_host_paths.py"), "10", "6"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "10 cases, 0 mismatches" in p.stdout

"""GPU parity for the PointPillar scatter reductions ("next" row N4; muvo/models/common.py:703, :731).

torch_scatter is not installable here, so the checks are the oracle's restatement of its documented semantics and
``torch.Tensor.scatter_reduce`` as an independent implementation.  scatter_max is exact; scatter_mean is a 64-bit
fixed-point sum (deterministic: bit-identical under any permutation of the rows), bar 1e-6 relative to the largest
magnitude against the float64 oracle."""
import numpy as np
import pytest
import torch

import oracle as O
from muvo_b200 import pillars

pytestmark = pytest.mark.gpu


def _case(N, F, M, seed, idx_dtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn((N, F), generator=g)
    src[torch.rand((N, F), generator=g) < 0.05] = 1.5          # exact ties inside pillars
    idx = torch.randint(0, max(M - 2, 1), (N,), generator=g).to(idx_dtype)   # the last two pillars stay empty
    return src, idx


@pytest.mark.parametrize("N,F,M,dt", [(5000, 3, 700, torch.int64), (4096, 32, 300, torch.int64), (1, 9, 4, torch.int32),
                                      (257, 1, 257, torch.int32)])
def test_scatter_mean_and_max_vs_oracle(lib, N, F, M, dt):
    src, idx = _case(N, F, M, 100 + N + F, dt)
    mean = pillars.scatter_mean(src.cuda(), idx.cuda(), dim=0, dim_size=M).cpu().numpy()
    want = O.scatter_mean(src.numpy(), idx.numpy(), M)
    assert np.abs(mean - want).max() <= 1e-6 * max(1.0, np.abs(src.numpy()).max())
    mx, arg = pillars.scatter_max(src.cuda(), idx.cuda(), dim=0, dim_size=M)
    wmx, warg = O.scatter_max(src.numpy(), idx.numpy(), M)
    assert np.array_equal(mx.cpu().numpy(), wmx) and np.array_equal(arg.cpu().numpy(), warg)
    # independent torch implementation
    ref = torch.zeros((M, F)).scatter_reduce(0, idx.long()[:, None].expand(-1, F), src, "amax", include_self=False)
    assert torch.equal(mx.cpu(), ref)
    pillars.check_indices()


def test_default_dim_size_and_gradients(lib):
    src, idx = _case(3000, 8, 200, 7)
    s = src.cuda().requires_grad_(True)
    mean = pillars.scatter_mean(s, idx.cuda())
    assert mean.shape[0] == int(idx.max()) + 1
    w = torch.randn(mean.shape, generator=torch.Generator().manual_seed(1)).cuda()
    (mean * w).sum().backward()
    cnt = torch.bincount(idx, minlength=mean.shape[0]).clamp(min=1).float()
    want = (w.cpu() / cnt[:, None])[idx]
    assert torch.allclose(s.grad.cpu(), want, rtol=1e-6, atol=1e-7)
    s2 = src.cuda().requires_grad_(True)
    mx, arg = pillars.scatter_max(s2, idx.cuda())
    (mx * w).sum().backward()
    want = torch.zeros_like(src)
    a = arg.cpu()
    for m in range(a.shape[0]):
        for f in range(a.shape[1]):
            if a[m, f] < src.shape[0]:
                want[a[m, f], f] = w[m, f].item()
    assert torch.equal(s2.grad.cpu(), want)


def test_empty_inputs_and_bad_index(lib):
    out = pillars.scatter_mean(torch.zeros((0, 3)).cuda(), torch.zeros((0,), dtype=torch.int64).cuda(), dim_size=5)
    assert out.shape == (5, 3) and torch.all(out == 0)
    mx, arg = pillars.scatter_max(torch.zeros((0, 3)).cuda(), torch.zeros((0,), dtype=torch.int64).cuda(), dim_size=5)
    assert torch.all(mx == 0) and torch.all(arg == 0)            # arg = N = 0 for empty rows
    pillars.check_indices()
    pillars.scatter_mean(torch.ones((4, 2)).cuda(), torch.tensor([0, 1, 9, 1]).cuda(), dim_size=3)
    with pytest.raises(IndexError):
        pillars.check_indices()
    with pytest.raises(NotImplementedError):
        pillars.scatter_mean(torch.ones((4, 2)).cuda(), torch.tensor([0, 1, 0, 1]).cuda(), dim=1)


def test_decorate_and_canvas_match_a_torch_restatement(lib):
    """PointPillarNet.grid_locations / pillar_generation / decorate / scatter_points (common.py:721-761) on one cloud."""
    g = torch.Generator().manual_seed(3)
    pts = torch.cat([torch.rand((6000, 2), generator=g) * torch.tensor([90.0, 90.0]) - torch.tensor([15.0, 45.0]),
                     torch.randn((6000, 2), generator=g)], 1).cuda()          # x, y, z, intensity
    kept, yx = pillars.pillar_grid_locations(pts)
    assert kept.shape[0] < pts.shape[0] and int(yx.min()) >= 0 and int(yx[:, 0].max()) < 320 and int(yx[:, 1].max()) < 320
    byx = torch.nn.functional.pad(yx, (1, 0), mode="constant", value=0)
    uniq, inv = byx.unique(return_inverse=True, dim=0)
    dec = pillars.pillar_decorate(kept, uniq, inv)
    assert dec.shape == (kept.shape[0], 4 + 3 + 2)
    mean = torch.zeros((uniq.shape[0], 3), device="cuda", dtype=torch.float64).index_add_(0, inv, kept[:, :3].double())
    mean = (mean / torch.bincount(inv, minlength=uniq.shape[0]).clamp(min=1)[:, None]).float()
    assert torch.allclose(dec[:, 4:7], kept[:, :3] - mean[inv], rtol=0, atol=1e-5)
    feat = torch.randn((kept.shape[0], 16), generator=torch.Generator(device="cuda").manual_seed(4), device="cuda")
    fmax, _ = pillars.scatter_max(feat, inv, dim_size=uniq.shape[0])
    canvas = pillars.pillar_scatter_points(fmax, uniq, 1, 320, 320)
    assert canvas.shape == (1, 16, 320, 320) and int((canvas.abs().sum(1) > 0).sum()) == uniq.shape[0]


def test_scatter_mean_is_deterministic_and_order_independent(lib):
    """north_star: no float atomics.  The fixed-point accumulation makes the result a function of the multiset of rows."""
    src, idx = _case(60000, 3, 9000, 11)
    src = src * 50.0                                                     # metres, like the xyz the reference averages (:731)
    a = pillars.scatter_mean(src.cuda(), idx.cuda(), dim_size=9000)
    b = pillars.scatter_mean(src.cuda(), idx.cuda(), dim_size=9000)
    perm = torch.randperm(src.shape[0], generator=torch.Generator().manual_seed(3))
    c = pillars.scatter_mean(src[perm].cuda(), idx[perm].cuda(), dim_size=9000)
    assert torch.equal(a, b) and torch.equal(a, c)
    want = O.scatter_mean(src.numpy().astype(np.float64), idx.numpy(), 9000)
    got = a.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= 6e-8 * np.abs(want).max() + 1e-30   # half a float32 ulp of the largest mean


def test_scatter_half_precision_and_special_values(lib):
    """under PRECISION='16-mixed' DynamicPointNet.forward feeds scatter_max float16 (common.py:702-703)."""
    src, idx = _case(2000, 16, 150, 21)
    for dt in (torch.float16, torch.bfloat16):
        s = src.to(dt).cuda().requires_grad_(True)
        mx, arg = pillars.scatter_max(s, idx.cuda(), dim_size=150)
        assert mx.dtype == dt
        wmx, warg = O.scatter_max(src.to(dt).float().numpy(), idx.numpy(), 150)
        assert np.array_equal(mx.detach().float().cpu().numpy(), wmx) and np.array_equal(arg.cpu().numpy(), warg)
        mx.float().sum().backward()
        assert s.grad.dtype == dt and float(s.grad.float().sum()) == float((warg < 2000).sum())
        mean = pillars.scatter_mean(src.to(dt).cuda(), idx.cuda(), dim_size=150)
        assert mean.dtype == dt
    with torch.autocast("cuda", dtype=torch.float16):
        h = torch.nn.functional.relu(torch.nn.Linear(16, 8).cuda()(src.cuda()))   # float16 under autocast
        out, _ = pillars.scatter_max(h, idx.cuda(), dim_size=150)
    assert h.dtype == torch.float16 and out.dtype == torch.float16
    # NaN / Inf propagate per output element; everything else is untouched
    s2 = src.clone()
    s2[0, 0], s2[1, 1], s2[2, 2], s2[3, 2] = float("nan"), float("inf"), float("inf"), float("-inf")
    idx2 = idx.clone(); idx2[2] = idx2[3]
    got = pillars.scatter_mean(s2.cuda(), idx2.cuda(), dim_size=150).cpu()
    assert torch.isnan(got[idx2[0], 0]) and got[idx2[1], 1] == float("inf") and torch.isnan(got[idx2[2], 2])
    clean = torch.ones_like(got, dtype=torch.bool)
    clean[idx2[0], 0] = clean[idx2[1], 1] = clean[idx2[2], 2] = False
    want = torch.from_numpy(O.scatter_mean(src.numpy(), idx2.numpy(), 150)).to(got.dtype)
    assert torch.allclose(got[clean], want[clean], rtol=0, atol=1e-6 * float(src.abs().max()))


def test_out_of_range_index_is_flagged_and_backward_is_safe(lib):
    src, idx = _case(500, 4, 50, 5)
    idx[7], idx[9] = 50, -1                                              # one past the end, negative
    s = src.cuda().requires_grad_(True)
    mean = pillars.scatter_mean(s, idx.cuda(), dim_size=50)
    mx, _ = pillars.scatter_max(s, idx.cuda(), dim_size=50)
    (mean.sum() + mx.sum()).backward()                                   # must not read out of bounds
    assert torch.all(s.grad[7] == 0) and torch.all(s.grad[9] == 0) and torch.isfinite(s.grad).all()
    with pytest.raises(IndexError):
        pillars.check_indices()

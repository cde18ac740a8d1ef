"""GPU parity for the PointPillar scatter reductions ("next" row N4; muvo/models/common.py:703, :731).

torch_scatter is not installable here, so the checks are the oracle's restatement of its documented semantics and
``torch.Tensor.scatter_reduce`` as an independent implementation.  scatter_max is exact; scatter_mean is a float
atomic sum (like torch_scatter's), bar 1e-6 relative to the largest magnitude in the row."""
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

"""Randomised shapes of stage (c) (bev_pool / bev_pool_masked forward + backward, all memory layouts and dtypes) against float64
sums, and of stage (d) (ssc_counts, all prediction dtypes, masks) against the oracle.  python tools/fuzz_bev_ssc.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (the checker)
from muvo_b200.frustum_pooling import bev_pool, bev_pool_masked  # noqa: E402
from muvo_b200.metrics import ssc_counts  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(seed)
g = torch.Generator().manual_seed(seed)
bad = 0
TOL = 1e-5
for case in range(n_cases):
    B, N = int(rng.integers(1, 5)), int(rng.choice([1, 1, 2]))
    D, H, W = int(rng.integers(1, 9)), int(rng.integers(1, 14)), int(rng.integers(1, 60))
    C = int(rng.choice([1, 3, 8, 16, 17, 64, 100]))
    n_cells = int(rng.choice([1, 7, 100, 2304, 2600, 3000, 4096, 9720, 9721, 12800, 12801, 20000]))
    dt = [torch.float32, torch.float16, torch.bfloat16][int(rng.integers(3))]
    layout = int(rng.integers(3))                  # 0 channels-last, 1 (B,N,C,D,H,W) memory, 2 sliced (non-contiguous) channels-last
    n_pts = N * D * H * W
    if layout == 1:
        base = torch.randn(B, N, C, D, H, W, generator=g).to(dt).cuda()
        x = base.permute(0, 1, 3, 4, 5, 2)
    elif layout == 2:
        base = torch.randn(B, N, D, H, W, C + 3, generator=g).to(dt).cuda()
        x = base[..., 1:C + 1]
    else:
        x = torch.randn(B, N, D, H, W, C, generator=g).to(dt).cuda()
    x = x.requires_grad_(True)
    cell = torch.randint(-1, n_cells, (B, n_pts), generator=g, dtype=torch.int32)
    if rng.random() < 0.5:                          # clustered cells (runs) like a frustum
        cell = (cell // max(1, int(rng.integers(1, 50)))).clamp(max=n_cells - 1)
    use_mask = rng.random() < 0.6
    mask = (torch.rand(B, n_pts, generator=g) < rng.random()) if use_mask else None
    try:
        out = bev_pool_masked(x, cell.cuda(), mask.cuda() if use_mask else None, n_cells)
        cc = cell.clone().long()
        if use_mask:
            cc[~mask] = -1
        xf = x.detach().reshape(B, n_pts, C).double().cpu()
        want = torch.zeros(B, n_cells, C, dtype=torch.float64)
        mag = torch.zeros(B, n_cells, C, dtype=torch.float64)
        for b in range(B):
            keep = cc[b] >= 0
            want[b].index_add_(0, cc[b][keep], xf[b][keep])
            mag[b].index_add_(0, cc[b][keep], xf[b][keep].abs())
        ok = out.shape == (B, C, n_cells) and bool(torch.all((out.detach().cpu().double() - want.permute(0, 2, 1)).abs() <= TOL * mag.permute(0, 2, 1) + 1e-30))
        gout = torch.randn(out.shape, generator=g).cuda()
        (gx,) = torch.autograd.grad(out, x, gout)
        exp = torch.zeros(B, n_pts, C)
        for b in range(B):
            keep = cc[b] >= 0
            exp[b][keep] = gout.cpu()[b].t()[cc[b][keep]]
        ok &= bool(torch.equal(gx.reshape(B, n_pts, C).cpu(), exp.to(dt)))
        ok &= bool(torch.equal(bev_pool(x, torch.where(cc >= 0, cc, torch.full_like(cc, -1)).int().cuda(), n_cells), out))
    except Exception as e:  # noqa: BLE001
        ok = False
        print("EXCEPTION", repr(e)[:300])
    if not ok:
        bad += 1
        print("BEV MISMATCH case", case, dict(B=B, N=N, D=D, H=H, W=W, C=C, n_cells=n_cells, dt=str(dt), layout=layout, mask=use_mask), flush=True)
    # ---- (d)
    F = int(rng.integers(1, 4))
    shape = (F, int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(1, 30)))
    Cn = int(rng.choice([1, 2, 9, 23, 40]))
    pdt = [torch.int64, torch.int32, torch.int16, torch.uint8][int(rng.integers(4))]
    yp = torch.randint(0, Cn, shape, generator=g).to(pdt)
    yt = torch.randint(0, Cn, shape, generator=g).to(torch.uint8)
    yt[torch.rand(shape, generator=g) < 0.05] = 255
    ne = (torch.rand(shape, generator=g) < 0.7) if rng.random() < 0.5 else None
    ns = (torch.rand(shape, generator=g) < 0.7) if rng.random() < 0.5 else None
    ig = bool(rng.random() < 0.5)
    try:
        got = ssc_counts(yp.cuda(), yt.cuda(), Cn, ne.cuda() if ne is not None else None, ns.cuda() if ns is not None else None, ig).cpu().numpy()
    except Exception as e:  # noqa: BLE001
        print("EXCEPTION", repr(e)[:300])
        got = None
    want = O.ssc_counts(yp.numpy(), yt.numpy(), Cn, ne.numpy() if ne is not None else None, ns.numpy() if ns is not None else None, ig)
    if got is None or not np.array_equal(got, want):
        bad += 1
        print("SSC MISMATCH case", case, dict(shape=shape, C=Cn, pdt=str(pdt), ne=ne is not None, ns=ns is not None, ig=ig), flush=True)
print(f"{n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)

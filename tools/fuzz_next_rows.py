"""Randomised shapes of the "next" rows against the oracle: scal-loss sums, PointPillar scatter mean / max,
label pyramids, sparse -> dense densify (incl. duplicate rows), fused argmax + IoU counts, fused lift-splat vs pooling the lifted
tensor.  python tools/fuzz_next_rows.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402  (the checker)
import muvo_b200  # noqa: E402
from muvo_b200 import pillars, synth  # noqa: E402
from muvo_b200.frustum_pooling import bev_pool, build_lift_splat_plan, fold_mask, lift_splat  # noqa: E402
from muvo_b200.losses import scal_sums  # noqa: E402
from muvo_b200.metrics import ssc_counts, ssc_counts_from_logits  # noqa: E402
from muvo_b200.points import densify_voxels  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
g = torch.Generator().manual_seed(seed)
bad = 0


def report(what, ok, **cfg):
    global bad
    if not ok:
        bad += 1
        print(what, "MISMATCH", cfg, flush=True)


for case in range(n_cases):
    try:
        # ---- scal sums
        C = int(rng.choice([2, 3, 9, 23]))
        shp = (int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.integers(1, 12)), int(rng.integers(1, 12)), int(rng.integers(1, 12)))
        dt = [torch.float32, torch.float16, torch.bfloat16][int(rng.integers(3))]
        pred = (torch.randn(shp[:2] + (C,) + shp[2:], generator=g) * 3).to(dt)
        tgt = torch.randint(0, C, shp, generator=g).to(torch.uint8)
        tgt[torch.rand(shp, generator=g) < 0.1] = 255
        got = scal_sums(pred.cuda(), tgt.cuda()).cpu().numpy()
        want = O.scal_sums(pred.float().numpy(), tgt.numpy())
        report("scal", np.array_equal(got[2 * C:], want[2 * C:]) and np.allclose(got[:2 * C], want[:2 * C], rtol=1e-6, atol=1e-9), C=C, shp=shp, dt=str(dt))
        # ---- fused argmax + counts == counts of torch.argmax
        lg = pred[:, 0].contiguous()                                 # (F, C, X, Y, Z)
        tt = tgt[:, 0].contiguous()
        a = ssc_counts_from_logits(lg.cuda(), tt.cuda()).cpu()
        b = ssc_counts(torch.argmax(lg.cuda(), 1), tt.cuda(), C, ignore255=True).cpu()
        report("argmax+counts", torch.equal(a, b), C=C, shp=shp, dt=str(dt))
        # ---- pillar scatter
        N, F, M = int(rng.choice([1, 33, 1000, 5000])), int(rng.choice([1, 3, 9, 32, 64])), int(rng.choice([1, 5, 300, 2000]))
        src = torch.randn(N, F, generator=g)
        src[torch.rand((N, F), generator=g) < 0.05] = 1.5
        idx = torch.randint(0, M, (N,), generator=g).to([torch.int64, torch.int32][int(rng.integers(2))])
        mean = pillars.scatter_mean(src.cuda(), idx.cuda(), dim=0, dim_size=M).cpu().numpy()
        ok = np.abs(mean - O.scatter_mean(src.numpy(), idx.numpy(), M)).max() <= 1e-6 * max(1.0, np.abs(src.numpy()).max())
        mx, arg = pillars.scatter_max(src.cuda(), idx.cuda(), dim=0, dim_size=M)
        ref = torch.zeros((M, F)).scatter_reduce(0, idx.long()[:, None].expand(-1, F), src, "amax", include_self=False)
        ok &= torch.equal(mx.cpu(), ref)
        if N <= 1000:
            wmx, warg = O.scatter_max(src.numpy(), idx.numpy(), M)
            ok &= np.array_equal(mx.cpu().numpy(), wmx) and np.array_equal(arg.cpu().numpy(), warg)
        report("pillar scatter", ok, N=N, F=F, M=M)
        # ---- pyramids
        Fp, H, W = int(rng.integers(1, 4)), int(rng.choice([4, 8, 10, 64, 30])), int(rng.choice([4, 16, 22, 64, 1024, 100]))
        X, Y, Z = int(rng.choice([4, 8, 13, 48, 192])), int(rng.choice([4, 9, 48, 192])), int(rng.choice([4, 6, 16, 64]))
        xyzd = torch.from_numpy(rng.normal(0, 30, (Fp, 4, H, W)).astype(np.float32)).cuda()
        sem = torch.from_numpy(rng.integers(0, 23, (Fp, H, W)).astype(np.uint8)).cuda()
        vox = torch.from_numpy(rng.integers(0, 3, (Fp, X, Y, Z)).astype(np.uint8)).cuda()
        gp = muvo_b200.label_pyramids(xyzd, sem, vox, scale=50.0)
        wp = O.label_pyramids(xyzd.cpu().numpy(), sem.cpu().numpy(), vox.cpu().numpy(), scale=50.0)
        report("pyramids", set(gp) == set(wp) and all(np.array_equal(gp[k].cpu().numpy(), wp[k]) for k in wp), F=Fp, H=H, W=W, X=X, Y=Y, Z=Z)
        # ---- densify (duplicates: the last row wins)
        size = (int(rng.choice([8, 48, 192])), int(rng.choice([8, 50, 192])), int(rng.choice([4, 64])))
        n = int(rng.choice([0, 1, 50, 4000]))
        rows = np.stack([rng.integers(0, size[0], n), rng.integers(0, size[1], n), rng.integers(0, size[2], n), rng.choice([0, 1, 6, 7, 13, 255], n)], 1).astype(np.uint16)
        if n > 10:
            rows[n // 2:n // 2 + 5, :3] = rows[:5, :3]               # repeated voxels
        remap = synth.label_remap256() if rng.random() < 0.5 else None
        gd = densify_voxels(rows, size, remap)
        gd = gd.cpu().numpy() if torch.is_tensor(gd) else np.asarray(gd)
        report("densify", np.array_equal(gd.reshape(size), O.densify_voxels(rows.copy(), size, remap)), size=size, n=n, remap=remap is not None)
        # ---- fused lift-splat == pooling the lifted tensor (same products, same order)
        B, Cc, D, Hh, Ww = int(rng.integers(1, 4)), int(rng.choice([4, 8, 12, 40, 64])), int(rng.integers(1, 7)), int(rng.integers(1, 9)), int(rng.integers(1, 30))
        n_cells = int(rng.choice([16, 300, 2304, 5000]))
        feat = torch.randn(B, Cc, Hh, Ww, generator=g).cuda().requires_grad_(True)
        dep = torch.rand(B, D, Hh, Ww, generator=g).cuda().requires_grad_(True)
        cell = torch.randint(-1, n_cells, (B, D * Hh * Ww), generator=g, dtype=torch.int32).cuda()
        out = lift_splat(feat, dep, cell, n_cells)
        xl = (dep.unsqueeze(1) * feat.unsqueeze(2)).unsqueeze(1).permute(0, 1, 3, 4, 5, 2)
        ref = bev_pool(xl, cell, n_cells)
        mag = bev_pool(xl.detach().abs(), cell, n_cells)                 # sum |x_i| per output: the fp32 summation-order bound
        ok = bool(torch.all((out - ref).abs() <= 2e-5 * mag + 1e-30))
        gout = torch.randn(out.shape, generator=g).cuda()
        gf, gd_ = torch.autograd.grad(out, (feat, dep), gout)
        rf, rd = torch.autograd.grad(ref, (feat, dep), gout)
        ok &= bool((gf - rf).abs().max() <= 2e-5 * rf.abs().max() + 1e-30) and bool((gd_ - rd).abs().max() <= 2e-5 * rd.abs().max() + 1e-30)
        report("lift-splat", bool(ok), B=B, C=Cc, D=D, H=Hh, W=Ww, n_cells=n_cells)
        # ---- the cached-plan forward (mask-independent sort + per-call mask filter) gives the same bits, masked and unmasked
        plan = build_lift_splat_plan(cell, n_cells)
        mk = torch.rand(cell.shape, generator=g).cuda() < float(rng.random())
        if rng.random() < 0.2:
            mk[int(rng.integers(B))] = False                             # a frame with everything masked
        folded = fold_mask(cell, mk)
        a = lift_splat(feat.detach(), dep.detach(), folded, n_cells, plan, mk)
        bref = lift_splat(feat.detach(), dep.detach(), folded, n_cells)
        a0 = lift_splat(feat.detach(), dep.detach(), cell, n_cells, plan, None)
        report("lift-splat plan", bool(torch.equal(a, bref) and torch.equal(a0, out.detach())), B=B, C=Cc, D=D, H=Hh, W=Ww, n_cells=n_cells)
    except Exception as e:  # noqa: BLE001
        bad += 1
        import traceback
        print("EXCEPTION in case", case, repr(e)[:300], flush=True)
        traceback.print_exc(limit=3)
print(f"{n_cases} cases, {bad} mismatches")
sys.exit(1 if bad else 0)

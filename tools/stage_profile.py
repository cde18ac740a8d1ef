"""Per-kernel CUDA-event breakdown of stages (c) and (d) at the BASELINE shapes (uses muvo_profile_begin/end)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muvo_b200  # noqa: E402
from muvo_b200 import _lib, synth  # noqa: E402
from muvo_b200.frustum_pooling import bev_pool  # noqa: E402
from muvo_b200.metrics import ssc_counts  # noqa: E402

dev = torch.device("cuda", 0)
stream = _lib.current_stream(dev)


def show(tag, fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    acc = {}
    for _ in range(reps):
        with _lib.profile(stream) as p:
            fn()
        for i, (k, ms) in enumerate(p.kernels):
            acc.setdefault((i, k), []).append(ms)
    tot = 0.0
    for (i, k), v in sorted(acc.items()):
        m = sum(v) / len(v)
        tot += m
        print(f"{tag:10s} {k:28s} {1e3 * m:9.1f} us")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{tag:10s} {'TOTAL':28s} {1e3 * tot:9.1f} us   (events around the call: {1e3 * e0.elapsed_time(e1) / reps:9.1f} us)")


B, C = 6, 384
feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000, device=dev)
fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
x = synth.lift(feat, depth).detach().requires_grad_(True)
fp.initialize_frustum(x)
cell = fp.cell_ids(fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None]), mask)
out = bev_pool(x, cell, 2304)
g = torch.ones_like(out)
show("bev_fwd", lambda: bev_pool(x.detach(), cell, 2304))
from muvo_b200.frustum_pooling import bev_pool_backward  # noqa: E402
show("bev_bwd", lambda: bev_pool_backward(g, cell, tuple(x.shape), x.dtype, 2304, x.stride()))
fl = feat.detach().requires_grad_(True)
dl = depth.detach().requires_grad_(True)
from muvo_b200.frustum_pooling import lift_splat  # noqa: E402
show("lift_splat_fwd", lambda: lift_splat(fl.detach(), dl.detach(), cell, 2304))
ol = lift_splat(fl, dl, cell, 2304)
show("lift_splat_bwd", lambda: torch.autograd.grad(ol, (fl, dl), g.view(ol.shape), retain_graph=True))
xc = x.detach().contiguous()
show("bev_fwd_cl", lambda: bev_pool(xc, cell, 2304))
for Cn in (2, 9, 23):
    yp, yt = synth.occupancy_pair(16, min(Cn, 23), 4000)
    tp, tt = torch.from_numpy(yp).to(dev), torch.from_numpy(yt).to(dev)
    show(f"ssc_C{Cn}", lambda: ssc_counts(tp, tt, Cn, ignore255=True))
del x, xc, out, g, fl, dl, ol
torch.cuda.empty_cache()
from muvo_b200.losses import scal_losses  # noqa: E402
for Cn, dt in ((2, torch.float32), (9, torch.float32), (2, torch.bfloat16), (5, torch.float32)):
    yp, yt = synth.occupancy_pair(16, min(Cn, 9), 4000)
    tl = torch.from_numpy(yt).to(dev).view(1, 16, 192, 192, 64)
    lg = torch.randn((1, 16, Cn, 192, 192, 64), device=dev).to(dt).requires_grad_(True)
    show(f"scal_C{Cn}_{str(dt)[6:]}", lambda: scal_losses(lg.detach(), tl))
    sem, geo = scal_losses(lg, tl)
    show(f"scalb_C{Cn}_{str(dt)[6:]}", lambda: torch.autograd.grad(sem + geo, lg, retain_graph=True))
    del lg, sem, geo

"""Per-kernel times of the fused lift-splat (N2) at cfg3 shapes, forward and backward.  python tools/lift_splat_time.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
import muvo_b200
from muvo_b200 import _lib, synth
dev = torch.device("cuda", 0)
feat, depth, mask, K, E = synth.bev_inputs(6, 384, 3000, device=dev)
fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
stream = _lib.current_stream(dev)
Kc, Ec = K[:, None].contiguous(), E[:, None].contiguous()
f, d = feat.requires_grad_(True), depth.requires_grad_(True)
for _ in range(3):
    out = fp.lift_splat(f, d, Kc, Ec, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = fp.lift_splat(f.detach(), d.detach(), Kc, Ec, mask)
e1.record(); torch.cuda.synchronize()
print("module lift_splat forward %.1f us" % (e0.elapsed_time(e1) * 50))
acc = {}
for _ in range(5):
    with _lib.profile(stream) as prof: fp.lift_splat(f.detach(), d.detach(), Kc, Ec, mask)
    for i, (k, v) in enumerate(prof.kernels): acc.setdefault((i, k), []).append(v * 1e3)
print({k: round(sum(v) / len(v), 1) for (i, k), v in sorted(acc.items())})

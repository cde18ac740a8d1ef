"""Streamed BEV pool at cfg3: kernel times with / without the consumer work (tuning key 3 bit 0 = skip the segment sums)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import muvo_b200
from muvo_b200 import _lib, synth
dev = torch.device("cuda", 0)
B, C = 6, 384
feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000, device=dev)
fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
x = synth.lift(feat, depth).detach()
Kc, Ec = K[:, None].contiguous(), E[:, None].contiguous()
lib = _lib.load()
stream = _lib.current_stream(dev)
for skip in (0, 1):
    lib.muvo_debug_set_tuning(3, skip)
    for m in (mask, torch.zeros(0, device=dev)):
        for _ in range(3): fp(x, Kc, Ec, m)
        torch.cuda.synchronize()
        acc = {}
        for _ in range(5):
            with _lib.profile(stream) as prof:
                fp(x, Kc, Ec, m)
            for k, v in prof.kernels: acc.setdefault(k, []).append(v * 1e3)
        print("skip", skip, "mask" if len(m) else "nomask", {k: round(sum(v) / len(v), 1) for k, v in acc.items()})
lib.muvo_debug_set_tuning(3, 0)

"""Streamed BEV forward at cfg3 for fp16 and fp32 inputs, with / without the consumer work (tuning key 3 = 1).  python tools/bev_stream_sweep16.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
import muvo_b200
from muvo_b200 import _lib, synth
dev = torch.device("cuda", 0)
feat, depth, mask, K, E = synth.bev_inputs(6, 384, 3000, device=dev)
fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
lib = _lib.load()
stream = _lib.current_stream(dev)
Kc, Ec = K[:, None].contiguous(), E[:, None].contiguous()
for dt in (torch.float16, torch.float32):
    x = synth.lift(feat.to(dt), depth.to(dt)).detach()
    for skip in (0, 1):
        lib.muvo_debug_set_tuning(3, skip)
        for _ in range(3): fp(x, Kc, Ec, mask)
        torch.cuda.synchronize()
        acc = {}
        for _ in range(5):
            with _lib.profile(stream) as prof: fp(x, Kc, Ec, mask)
            for k, v in prof.kernels: acc.setdefault(k, []).append(v * 1e3)
        print(dt, "skip", skip, {k: round(sum(v) / len(v), 1) for k, v in acc.items()})
lib.muvo_debug_set_tuning(3, 0)

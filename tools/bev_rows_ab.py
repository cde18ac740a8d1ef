"""A/B of the BEV-pool forward kernels at cfg3 through the module call FrustumPooling.forward(x, K, E, mask):
tuning key 2 = 0 -> streamed (bev_stream.cu, the default for (B,C,D,H,W) memory), 2 -> k_pool_rows_pipe, 1 -> k_pool_rows."""
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
nbytes = x.numel() * 4
for knob in (0, 2, 1):
    lib.muvo_debug_set_tuning(2, knob)
    for _ in range(3): fp(x, Kc, Ec, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fp(x, Kc, Ec, mask)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    with _lib.profile(stream) as prof:
        fp(x, Kc, Ec, mask)
    print("knob2", knob, "module fwd ms %.4f" % ms, "dense GB/s %.0f" % (nbytes / ms / 1e6), {k: round(v * 1e3, 1) for k, v in prof.kernels})
lib.muvo_debug_set_tuning(2, 0)
# unmasked
for _ in range(3): fp(x, Kc, Ec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fp(x, Kc, Ec)
e1.record(); torch.cuda.synchronize()
print("no mask, streamed: ms %.4f" % (e0.elapsed_time(e1) / 10))

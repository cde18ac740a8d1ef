"""cfg3 BEV backward through the module: streamed (TMA bulk stores, default) vs the gather kernel (tuning key 2 = 2)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import muvo_b200
from muvo_b200 import _lib, synth
dev = torch.device("cuda", 0)
lib = _lib.load()
stream = _lib.current_stream(dev)
for dt in (torch.float32, torch.float16):
    feat, depth, mask, K, E = synth.bev_inputs(6, 384, 3000, device=dev)
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
    x = synth.lift(feat.to(dt), depth.to(dt)).detach().requires_grad_(True)
    Kc, Ec = K[:, None].contiguous(), E[:, None].contiguous()
    gout = torch.randn((6, 384, 48, 48), device=dev)
    for knob in (0, 2):
        lib.muvo_debug_set_tuning(2, knob)
        out = fp(x, Kc, Ec, mask)
        g = gout.to(out.dtype)
        for _ in range(3): torch.autograd.grad(out, x, g, retain_graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): torch.autograd.grad(out, x, g, retain_graph=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        with _lib.profile(stream) as prof: torch.autograd.grad(out, x, g, retain_graph=True)
        print(dt, "knob2", knob, "module bwd ms %.4f" % ms, "GB/s %.0f" % (x.numel() * x.element_size() / ms / 1e6), [(k, round(v * 1e3, 1)) for k, v in prof.kernels])
    lib.muvo_debug_set_tuning(2, 0)

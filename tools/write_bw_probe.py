"""Plain write bandwidth of this GPU (torch zero_ / fill_ of 1.42 GB = the bytes of the BEV backward).  python tools/write_bw_probe.py"""
import torch
dev = torch.device("cuda", 0)
x = torch.empty(1418649600 // 4, dtype=torch.float32, device=dev)
for name, fn in (("zero_", lambda: x.zero_()), ("fill_(1)", lambda: x.fill_(1.0))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, "ms %.4f" % ms, "GB/s %.0f" % (x.numel() * 4 / ms / 1e6))

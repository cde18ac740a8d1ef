"""Label pyramids (N1) at 24 cfg2-sized frames under CUDA-graph replay (the host side of the wrapper is slower than the kernels)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import muvo_b200
from muvo_b200 import _lib
dev = torch.device("cuda", 0)
F = 24
xyzd = torch.randn(F, 4, 64, 1024, device=dev); sem = torch.randint(0, 20, (F, 64, 1024), dtype=torch.uint8, device=dev)
vox = torch.randint(0, 20, (F, 192, 192, 64), dtype=torch.uint8, device=dev)
def timed(fn, what):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(what, "%.1f us per call" % (e0.elapsed_time(e1) * 1e3 / 50))
timed(lambda: muvo_b200.label_pyramids(xyzd, sem, None, scale=50.0), "range pyramid (25 MB in, 35 MB out)")
timed(lambda: muvo_b200.label_pyramids(None, None, vox, scale=50.0), "voxel pyramid (14 MB in, 8 MB out)")

"""Where the end-to-end time goes: host staging copy, H2D, D2H rates on this box, and HostPipeline per-phase times."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muvo_b200 import _lib, synth  # noqa: E402
from muvo_b200.pipeline import HostPipeline  # noqa: E402
from muvo_b200.points import GridSpec, RangeSpec  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
print("cores", os.cpu_count())
n = 100 << 20
src = np.random.randint(0, 255, n, dtype=np.uint8)
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
for th in (1, 2, 4, 8, 16, 32):
    lib.muvo_host_copy(pin.data_ptr(), src.ctypes.data, n, th)
    t = time.perf_counter()
    for _ in range(5):
        lib.muvo_host_copy(pin.data_ptr(), src.ctypes.data, n, th)
    dt = (time.perf_counter() - t) / 5
    print(f"host_copy threads={th}: {n / dt / 1e9:.1f} GB/s")
for name, a, b in (("H2D", d, pin), ("D2H", pin, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        a.copy_(b, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
pin2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(pin, non_blocking=True)
    with torch.cuda.stream(s2): pin2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"duplex: {5 * n / dt / 1e9:.1f} GB/s each way")

pts, sem, off = synth.lidar_batch(96, 60000, 100000, 2000)
pipe = HostPipeline(dev, grid=GridSpec(), range_spec=RangeSpec(lidar_position=(1.0, 0.0, 2.0)), dense=False, sparse=True, layout="hwc")
pipe.warmup(pts, sem, off)
for _ in range(3):
    pipe.submit(pts, sem, off); pipe.result()
torch.cuda.synchronize()
t = time.perf_counter(); pipe.submit(pts, sem, off); t1 = time.perf_counter(); pipe.result(); t2 = time.perf_counter()
print(f"single batch: submit {1e3 * (t1 - t):.2f} ms, result wait {1e3 * (t2 - t1):.2f} ms")
t = time.perf_counter()
pipe.submit(pts, sem, off)
for _ in range(9):
    pipe.submit(pts, sem, off); pipe.result()
pipe.result(); torch.cuda.synchronize()
print(f"pipelined: {1e3 * (time.perf_counter() - t) / 10:.2f} ms/batch")

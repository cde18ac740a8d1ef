"""cfg2 step run as frame groups (CUDA-graph replay, one stream): do the winner tables stay in L2 when a group's kernels
reuse the same small workspace?  Prints ms/step per group size."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from muvo_b200 import synth
from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid
dev = torch.device("cuda", 0)
pts, sem, off = bench.make_batch(0)
F = len(off) - 1
d_pts, d_sem = torch.from_numpy(pts).to(dev), torch.from_numpy(sem).to(dev)
grid, rspec = GridSpec(), RangeSpec(lidar_position=bench.LIDAR)
remap = torch.from_numpy(synth.label_remap256()).to(dev)
out = {"voxel": torch.empty((F, 192, 192, 64), dtype=torch.uint8, device=dev), "n_occ": torch.empty((F,), dtype=torch.int64, device=dev),
       "range_xyzd": torch.empty((F, 4, 64, 1024), dtype=torch.float32, device=dev), "range_sem": torch.empty((F, 64, 1024), dtype=torch.uint8, device=dev)}
ref = None
for gs in (96, 48, 24, 12, 8, 4):
    groups = []
    for f0 in range(0, F, gs):
        f1 = min(F, f0 + gs)
        p0, p1 = int(off[f0]), int(off[f1])
        o = torch.from_numpy(off[f0:f1 + 1] - off[f0]).to(dev)
        groups.append((d_pts[p0:p1], d_sem[p0:p1], o, {k: v[f0:f1] for k, v in out.items()}))
    def step():
        for p, s, o, og in groups:
            sensor_to_grid(p, s, o, grid=grid, range_spec=rspec, dense=True, sparse=False, remap=remap, layout="xyzd", out=og)
    for _ in range(3): step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        step()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    cs = {k: int(v.view(torch.uint8).sum(dtype=torch.int64)) for k, v in out.items()}
    if ref is None: ref = cs
    print("group", gs, "ms/step %.4f" % (e0.elapsed_time(e1) / 20), "same" if cs == ref else "DIFF", flush=True)

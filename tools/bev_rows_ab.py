"""A/B of the two BEV-pool row kernels at cfg3: tuning key 2 = 1 -> k_pool_rows (one CTA per row), 0 -> k_pool_rows_pipe."""
import sys, torch
sys.path.insert(0, '/root/repo')
import muvo_b200
from muvo_b200 import _lib, synth
from muvo_b200.frustum_pooling import bev_pool
dev = torch.device("cuda", 0)
B, C = 6, 384
feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000, device=dev)
fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(dev)
x = synth.lift(feat, depth).detach()
fp.initialize_frustum(x)
cell = fp.cell_ids(fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None]), mask)
lib = _lib.load()
for knob in (1, 0):
    lib.muvo_debug_set_tuning(2, knob)
    for _ in range(3): bev_pool(x, cell, 2304)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): bev_pool(x, cell, 2304)
    e1.record(); torch.cuda.synchronize()
    print("knob2", knob, "ms", e0.elapsed_time(e1) / 10)

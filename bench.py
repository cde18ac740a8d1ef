#!/usr/bin/env python
"""Headline benchmark: LiDAR points/s voxelised + range-projected (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # ours   (torchrun launches N ranks)
    python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm (oracle port, rank 0 only)

A "step" = stages (a)+(b) over one batch of 8 x 12 = 96 ragged synthetic CARLA-style frames
(60-100 k points each): dense 192x192x64 uint8 occupancy grids + (4,64,1024) range views.
One JSON line on rank 0.  Extra stage numbers (BEV pool fwd/bwd GB/s, IoU counts) ride along
under "stages".  See DESIGN.md for the byte accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lidar_points_per_sec_voxelised_projected"
UNIT = "points/s"
F_BATCH, F_SEQ = 8, 12
N_MIN, N_MAX = 60000, 100000
G = 192 * 192 * 64
HW = 64 * 1024
LIDAR = (1.0, 0.0, 2.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(rank: int, n_frames: int = F_BATCH * F_SEQ, n_min: int = N_MIN, n_max: int = N_MAX):
    from muvo_b200 import synth
    return synth.lidar_batch(n_frames, n_min, n_max, 2000 + 100000 * rank)


# ----------------------------------------------------------------------------- CPU reference arm
_REF = {"tried": False, "ns": None}


def ref_namespace():
    """The UNMODIFIED reference functions (oracle/ref_import.py: $MUVO_REFERENCE_ROOT, /root/reference or baseline/_ref),
    or None -- then the oracle's restatement ("port") is timed instead."""
    if not _REF["tried"]:
        _REF["tried"] = True
        try:
            import warnings
            warnings.filterwarnings("ignore")
            import oracle.ref_import as R
            _REF["ns"] = R.load() if R.available() else None
        except Exception as e:                                  # a missing optional import of the reference, ...
            print(f"bench: reference not importable ({e}); timing the oracle port", file=sys.stderr)
            _REF["ns"] = None
    return _REF["ns"]


def cpu_kind():
    return "reference" if ref_namespace() is not None else "port"


def _cpu_frame(args):
    """(a) voxel_filter + densify and (b) do_range_projection + pack for one frame, on one core."""
    import warnings
    warnings.filterwarnings("ignore")
    import oracle as O
    from muvo_b200 import synth
    p, s = args
    ns = ref_namespace()
    if ns is not None:
        v, l = ns.voxel_filter(p, s, 0.5, [192, 192, 64], [0.0, 0, -10.0])           # data/data_preprocessing.py:172-228, unmodified
        d, x, sm = ns.PointCloud(64, 1024, -30, 10, list(LIDAR)).do_range_projection(p, s)   # geometry_utils.py:175-220, unmodified
    else:
        v, l = O.voxel_filter_loop(p, s, 0.5, [192, 192, 64], [0.0, 0, -10.0])
        d, x, sm = O.range_projection(p, s, lidar_position=list(LIDAR))
    data = np.concatenate([v, l[:, None].astype(np.uint16)], 1)
    grid = O.densify_voxels(data, (192, 192, 64), synth.label_remap256())             # dataset.py:317-327 (glue; muvo.data.dataset needs lightning)
    O.pack_range_view(d, x)                                                            # dataset.py:301-303
    return int(grid.sum()) + int(sm.sum())


def cpu_reference(frames, pool, cores):
    """One bounded sample: `frames` = list of (points, sem); returns (points, seconds)."""
    t0 = time.perf_counter()
    list(pool.map(_cpu_frame, frames, chunksize=1))
    dt = time.perf_counter() - t0
    return sum(len(p) for p, _ in frames), dt


def cpu_sample_frames(n_frames, cfg="cfg2"):
    if cfg == "cfg5":
        from muvo_b200 import synth
        return [synth.carla_lidar_frame(1_000_000, 5000 + i) for i in range(n_frames)]
    pts, sem, off = make_batch(0, n_frames)
    return [(pts[off[f]:off[f + 1]].copy(), sem[off[f]:off[f + 1]].copy()) for f in range(n_frames)]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ref_namespace()                                   # import once in the parent: forked workers inherit it
    if args.config == "cfg5":
        n_frames = 16                                 # the stated 16-frame subset of SURVEY.md 8(d) item 5 (a full pass is ~7700 core-seconds)
    else:
        n_frames = max(cores, 8)                      # one frame per worker per step: ~0.3-0.6 s of CPU work per frame
    frames = cpu_sample_frames(n_frames, args.config)
    ctx = mp.get_context("fork")
    steps = args.steps if args.config != "cfg5" else max(1, min(args.steps, 2))
    with ctx.Pool(min(cores, n_frames)) as pool:
        for _ in range(1 if args.warmup else 0):
            cpu_reference(frames[:min(cores, n_frames)], pool, cores)
        tot_pts, tot_s = 0, 0.0
        for _ in range(steps):
            n, dt = cpu_reference(frames, pool, cores)
            tot_pts += n
            tot_s += dt
    value = tot_pts / tot_s
    kind = cpu_kind()
    what = ("unmodified voxel_filter + PointCloud.do_range_projection of the reference" if kind == "reference"
            else "oracle *_loop port of voxel_filter + do_range_projection") + " (+ densify / pack glue)"
    workload, metric = ("cfg2", METRIC) if args.config != "cfg5" else ("cfg5", METRIC)
    sample = (f"{n_frames} frames/step of the {workload} generator ({tot_pts // max(steps, 1)} points/step), "
              f"multiprocessing.Pool({min(cores, n_frames)}) over frames like data/generate_voxels.py:134; {what}, numpy {np.__version__}")
    cfg = ({"workload": "cfg2: range-view projection + occupancy voxelisation (CPU reference, bounded sample)",
            "frames_per_step": n_frames, "points_per_frame": f"{N_MIN}-{N_MAX}"} if args.config != "cfg5" else
           {"workload": "cfg5: 1M-point frames, voxelise + project (CPU reference on a 16-frame subset; 768 frames extrapolate linearly: "
                        f"{768 * 1e6 / value:.0f} s on this box)", "frames_per_step": n_frames, "points_per_frame": 1_000_000})
    line = {"impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1 if args.warmup else 0, "ms_per_step": 1e3 * tot_s / max(steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(cores, n_frames), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)


def copy_ceiling(device, h2d_bytes, d2h_bytes, world, barrier, iters=8):
    """ms per step for moving h2d_bytes in and d2h_bytes out (pinned host <-> this rank's GPU) concurrently on two streams,
    all ranks at once: what the end-to-end loop would take if the kernels were free."""
    import torch
    import torch.distributed as dist
    hin = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8).pin_memory()
    hout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    din = torch.empty(max(h2d_bytes, 1), dtype=torch.uint8, device=device)
    dout = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def once():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    for _ in range(2):
        once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        once()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return 1e3 * float(t.item()) / iters


# ----------------------------------------------------------------------------- ours
def run_ours(args):
    if args.config == "cfg5":
        return run_cfg5(args)
    import torch
    import torch.distributed as dist
    import muvo_b200
    from muvo_b200 import _lib, synth
    from muvo_b200.distributed import init_distributed
    from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid
    from muvo_b200.pipeline import HostPipeline

    rank, world, device = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback for the muvo_b200 kernels)")
    hbm_peak, peak_src = peaks()
    n_frames = F_BATCH * F_SEQ
    pts, sem, off = make_batch(rank)
    P = int(pts.shape[0])
    d_pts, d_sem = torch.from_numpy(pts).to(device), torch.from_numpy(sem).to(device)
    d_off = torch.from_numpy(off).to(device)
    grid, rspec = GridSpec(), RangeSpec(lidar_position=LIDAR)
    remap = torch.from_numpy(synth.label_remap256()).to(device)
    out = {"voxel": torch.empty((n_frames, 192, 192, 64), dtype=torch.uint8, device=device),
           "n_occ": torch.empty((n_frames,), dtype=torch.int64, device=device),
           "range_xyzd": torch.empty((n_frames, 4, 64, 1024), dtype=torch.float32, device=device),
           "range_sem": torch.empty((n_frames, 64, 1024), dtype=torch.uint8, device=device)}

    def step():
        return sensor_to_grid(d_pts, d_sem, d_off, grid=grid, range_spec=rspec, dense=True, sparse=False, remap=remap,
                              layout="xyzd", out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # per-kernel durations, CUDA events on the launching stream, same step repeated
    stream = _lib.current_stream(device)
    per_kernel = {}
    n_launch_per_step = 0
    for _ in range(args.steps):
        with _lib.profile(stream) as prof:
            step()
        n_launch_per_step = len(prof.kernels)
        for name, ms in prof.kernels:
            per_kernel.setdefault(name, []).append(ms)
    kern = {k: float(np.mean(v)) for k, v in per_kernel.items()}
    n_total_pts = torch.tensor([P], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(n_total_pts)
    total_pts = float(n_total_pts.item())
    value = total_pts / (ms_step * 1e-3)

    # algorithmic bytes: 13 B/point + G + 64*1024*17 per frame
    alg_step = 13 * P + n_frames * (G + HW * 17)
    alg_kernel = {"k_points_tile": 13 * P, "k_emit_dense": n_frames * G,
                  "k_emit_range": n_frames * HW * 17, "k_bitmap_scan": n_frames * (G // 8)}
    top = max(kern, key=kern.get)
    achieved = alg_kernel.get(top, 0) / (kern[top] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")      # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if top in tj.get("kernels", {}):
            traffic = tj["kernels"][top]["dram_read_bytes"] + tj["kernels"][top]["dram_write_bytes"]
            traffic_src = tj.get("source")
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_kernel.get(top, 0), "ms_per_launch": kern[top],
                "step": {"algorithmic_bytes": alg_step, "achieved": alg_step / (ms_step * 1e-3) / 1e9,
                         "frac": alg_step / (ms_step * 1e-3) / 1e9 / hbm_peak},
                "kernels_ms": kern}

    # end to end through the host-buffer API: pinned staging, H2D, kernels, D2H of the reference-facing results
    pipe = HostPipeline(device, grid=grid, range_spec=rspec, dense=False, sparse=True, layout="hwc", depth=3)
    pipe.warmup(pts, sem, off)                 # every slot's pinned / device buffers allocated
    # the step's inputs in pinned host memory (filled once, outside the timed region), as the bench contract words it; the
    # same loop from ordinary pageable numpy arrays (one more staging copy per step) is reported next to it
    pin_pts, pin_sem, pin_off = HostPipeline.pinned_inputs(P, n_frames)
    pin_pts.numpy()[...] = pts; pin_sem.numpy()[...] = sem; pin_off.numpy()[...] = off
    e2e_steps = max(args.steps, 4)

    def e2e_loop(a_pts, a_sem, a_off):
        for _ in range(3):
            pipe.submit(a_pts, a_sem, a_off)
            pipe.result()
        barrier()
        t0 = time.perf_counter()
        submitted = done = 0
        while done < e2e_steps:
            while submitted < e2e_steps and submitted - done < len(pipe.slots):   # keep every slot busy
                pipe.submit(a_pts, a_sem, a_off)
                submitted += 1
            pipe.result()
            done += 1
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    e2e_pageable_s = e2e_loop(pts, sem, off)
    e2e_s = e2e_loop(pin_pts, pin_sem, pin_off)
    h2d_b, d2h_b = int(pipe.h2d_bytes), int(pipe.d2h_bytes)
    pipe.close()
    # the box's ceiling for exactly these copies: every rank moves the step's H2D and D2H bytes between pinned host memory and
    # its GPU, both directions at once, all ranks at the same time, no kernels (max over ranks, like the e2e figure)
    ceil_ms = copy_ceiling(device, h2d_b, d2h_b, world, barrier)
    # the same inputs with the results handed over ON THE DEVICE (what a training step needs; only the per-frame counts come back)
    pipe_d = HostPipeline(device, grid=grid, range_spec=rspec, dense=True, sparse=False, layout="xyzd", depth=3, device_out=True,
                          remap=synth.label_remap256())
    pipe_d.warmup(pin_pts, pin_sem, pin_off)
    pipe, pipe_h = pipe_d, pipe
    e2e_dev_s = e2e_loop(pin_pts, pin_sem, pin_off)
    h2d_d, d2h_d = int(pipe_d.h2d_bytes), int(pipe_d.d2h_bytes)
    ceil_dev_ms = copy_ceiling(device, h2d_d, max(d2h_d, 8), world, barrier)
    pipe_d.close()
    clocks = sampler.stop()                     # sampled across the device-timed and the end-to-end timed regions
    e2e = {"value": total_pts * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_b,
           "d2h_bytes_per_step": d2h_b, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "copy_ceiling": {"ms_per_step": ceil_ms, "frac": ceil_ms / (1e3 * e2e_s / e2e_steps),
                            "aggregate_GBps": world * (h2d_b + d2h_b) / ceil_ms / 1e6,
                            "what": "the same H2D + D2H bytes per step per rank, pinned <-> device, both directions and all ranks "
                                    "at once, no kernels, measured in this run: e2e / ceiling = frac"},
           "from_pageable_numpy": {"value": total_pts * e2e_steps / e2e_pageable_s, "ms_per_step": 1e3 * e2e_pageable_s / e2e_steps,
                                   "note": "same loop with ordinary numpy inputs: one more 98.6 MB staging copy per step on host threads"},
           "device_handoff": {"value": total_pts * e2e_steps / e2e_dev_s, "ms_per_step": 1e3 * e2e_dev_s / e2e_steps,
                              "h2d_bytes_per_step": h2d_d, "d2h_bytes_per_step": d2h_d,
                              "copy_ceiling_ms": ceil_dev_ms, "frac_of_ceiling": ceil_dev_ms / (1e3 * e2e_dev_s / e2e_steps),
                              "note": "HostPipeline(device_out=True): dense grids + (4,H,W) range views stay on the GPU for the "
                                      "training step (muvo/data/dataset.py:301-327 builds exactly these), only n_occ is read back"},
           "api": "muvo_b200.pipeline.HostPipeline.submit/result (inputs in pinned host memory, pinned host out: sparse voxel "
                  "lists + HWC range images = the returns of voxel_filter / do_range_projection; the device-timed `value` "
                  "writes dense grids + XYZD images instead, same points, same tables, different emit kernels)"}

    stages = {}
    if not args.no_stages:
        stages = other_stages(device, rank, world, hbm_peak, args)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_base = cpu_baseline_leg()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "cfg2: range-view projection + occupancy voxelisation, batch 8 x seq 12 frames per GPU",
                           "frames_per_gpu": n_frames, "points_per_gpu": P, "points_per_frame": f"{N_MIN}-{N_MAX} (ragged)",
                           "grid": "192x192x64 @0.5m uint8 dense", "range_image": "4x64x1024 f32 + 64x1024 u8",
                           "l2": "inputs+outputs %.0f MB per step > 126 MB L2 (no explicit flush)" % (alg_step / 1e6),
                           "sharding": "frames sharded by rank, no data-path collective"},
                "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks,
                "gpu_launches": n_launch_per_step * args.steps, "kernels_per_step": n_launch_per_step, "stages": stages}
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_cfg5(args):
    """BASELINE.json configs[4]: 768 frames (batch 64 x seq 12) of 1 M points, sharded frame % world == rank, streamed through
    each GPU in 24-frame chunks: (a)+(b) on every chunk, (c) forward + backward at C = 384 on 24-frame chunks of lifted
    features, host-CPU reference on a 16-frame subset with explicit extrapolation (data/generate_voxels.py:110-164 is the
    reference's own loop over frames).  A step = one sweep over the rank's shard.  Fixed total work: "scaling": "strong"."""
    import torch
    import torch.distributed as dist
    import muvo_b200
    from muvo_b200 import _lib, synth
    from muvo_b200.distributed import init_distributed, shard_frames
    from muvo_b200.points import GridSpec, RangeSpec, sensor_to_grid
    from muvo_b200.pipeline import HostPipeline

    rank, world, device = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback for the muvo_b200 kernels)")
    hbm_peak, peak_src = peaks()
    FRAMES, PTS, CHUNK, DISTINCT = 768, 1_000_000, 24, 4
    mine = len(shard_frames(FRAMES, rank, world))
    sizes = [CHUNK] * (mine // CHUNK) + ([mine % CHUNK] if mine % CHUNK else [])
    # chunk buffers are reused: DISTINCT different synthetic 1 M-point frames tiled to fill a chunk (768 distinct frames
    # would take ten minutes to generate on the host); every pass is a full run of the point kernels over its frames
    fr = [synth.carla_lidar_frame(PTS, 5000 + 17 * rank + i) for i in range(DISTINCT)]
    pts = np.concatenate([fr[i % DISTINCT][0] for i in range(CHUNK)])
    sem = np.concatenate([fr[i % DISTINCT][1] for i in range(CHUNK)])
    off = np.arange(CHUNK + 1, dtype=np.int64) * PTS
    d_pts, d_sem, d_off = torch.from_numpy(pts).to(device), torch.from_numpy(sem).to(device), torch.from_numpy(off).to(device)
    grid, rspec = GridSpec(), RangeSpec(lidar_position=LIDAR)
    remap = torch.from_numpy(synth.label_remap256()).to(device)
    out = {"voxel": torch.empty((CHUNK, 192, 192, 64), dtype=torch.uint8, device=device),
           "n_occ": torch.empty((CHUNK,), dtype=torch.int64, device=device),
           "range_xyzd": torch.empty((CHUNK, 4, 64, 1024), dtype=torch.float32, device=device),
           "range_sem": torch.empty((CHUNK, 64, 1024), dtype=torch.uint8, device=device)}

    def sweep_points():
        for n in sizes:
            sensor_to_grid(d_pts[:n * PTS], d_sem[:n * PTS], d_off[:n + 1], grid=grid, range_spec=rspec, dense=True, sparse=False,
                           remap=remap, layout="xyzd", out={k: v[:n] for k, v in out.items()})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    sampler = ClockSampler(device.index or 0)
    sampler.start()
    warm = max(args.warmup, 3)
    ms_step = timed(sweep_points, args.steps, warm)
    total_frames = torch.tensor([float(mine)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_frames)
    total_frames = float(total_frames.item())
    value = total_frames * PTS / (ms_step * 1e-3)
    stream = _lib.current_stream(device)
    per_kernel = {}
    with _lib.profile(stream) as prof:
        sensor_to_grid(d_pts, d_sem, d_off, grid=grid, range_spec=rspec, dense=True, sparse=False, remap=remap, layout="xyzd", out=out)
    kern = {k: float(v) for k, v in prof.kernels}
    top = max(kern, key=kern.get)
    alg_kernel = {"k_points_tile": 13 * CHUNK * PTS, "k_emit_dense": CHUNK * G, "k_emit_range": CHUNK * HW * 17}
    alg_step = mine * (13 * PTS + G + HW * 17)
    achieved = alg_kernel.get(top, 0) / (kern[top] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_kernel.get(top, 0),
                "ms_per_launch": kern[top], "kernels_ms_per_24_frame_chunk": kern,
                "step": {"algorithmic_bytes": alg_step, "achieved": alg_step / (ms_step * 1e-3) / 1e9,
                         "frac": alg_step / (ms_step * 1e-3) / 1e9 / hbm_peak}}
    # end to end: the rank's chunks through the host-buffer API from pinned inputs (sparse voxel lists + HWC images back)
    pipe = HostPipeline(device, grid=grid, range_spec=rspec, dense=False, sparse=True, layout="hwc", depth=2)
    pin_pts, pin_sem, pin_off = HostPipeline.pinned_inputs(CHUNK * PTS, CHUNK)
    pin_pts.numpy()[...] = pts; pin_sem.numpy()[...] = sem; pin_off.numpy()[...] = off
    pipe.warmup(pin_pts, pin_sem, pin_off)
    n_e2e = len(sizes)
    barrier()
    t0 = time.perf_counter()
    submitted = done = 0
    while done < n_e2e:
        while submitted < n_e2e and submitted - done < len(pipe.slots):
            pipe.submit(pin_pts, pin_sem, pin_off)
            submitted += 1
        pipe.result()
        done += 1
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d_b, d2h_b = int(pipe.h2d_bytes), int(pipe.d2h_bytes)
    pipe.close()
    del pipe, pin_pts, pin_sem
    e2e = {"value": world * n_e2e * CHUNK * PTS / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_b * n_e2e,
           "d2h_bytes_per_step": d2h_b * n_e2e, "ms_per_step": 1e3 * e2e_s,
           "api": "HostPipeline.submit/result per 24-frame chunk, pinned inputs, pinned sparse voxel lists + HWC range images out"}
    # (c) BEV pool forward + backward through the module, 24-frame chunks of lifted features (C = 384)
    stages = {}
    if not args.no_stages:
        Bc, C = 24, 384
        feat, depth, mask, K, E = synth.bev_inputs(Bc, C, 3000 + rank, device=device)
        fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(device)
        x = synth.lift(feat, depth).detach().requires_grad_(True)
        Kc, Ec = K[:, None].contiguous(), E[:, None].contiguous()
        gout = torch.randn((Bc, C, 48, 48), device=device)
        n_bev = -(-mine // Bc)

        def sweep_bev():
            for _ in range(n_bev):
                o = fp(x, Kc, Ec, mask)
                torch.autograd.grad(o, x, gout)
        ms_bev = timed(sweep_bev, max(1, min(args.steps, 3)), 1)
        n_pts_frame = 37 * 40 * 104
        stages["bev_pool_fwd_bwd"] = {"ms_per_sweep": ms_bev, "frames_per_s": total_frames / (ms_bev * 1e-3), "chunk_frames": Bc,
                                      "dense_GBps_per_gpu": n_bev * Bc * (2 * n_pts_frame * C * 4 + 2 * C * 2304 * 4) / ms_bev / 1e6,
                                      "frac": n_bev * Bc * (2 * n_pts_frame * C * 4 + 2 * C * 2304 * 4) / ms_bev / 1e6 / hbm_peak,
                                      "note": "FrustumPooling.forward + autograd backward, lifted tensor read once / gradient written once"}
        del x, feat, depth, gout
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_base = cpu_baseline_leg("cfg5")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "cfg5: scaling sweep, 1M-point frames, batch 64 x seq 12 = 768 frames, voxelise + project "
                                       "(+ BEV pool under stages), frames sharded frame % world == rank",
                           "frames_total": FRAMES, "frames_per_gpu": mine, "points_per_frame": PTS, "chunk_frames": CHUNK,
                           "l2": "a 24-frame chunk moves 312 MB in + 80 MB out > 126 MB L2 (no explicit flush)",
                           "inputs": f"{DISTINCT} distinct synthetic frames tiled per chunk, device resident, chunk buffers reused"},
                "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks,
                "gpu_launches": len(kern) * len(sizes) * args.steps, "kernels_per_step": len(kern) * len(sizes), "stages": stages}
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_leg(cfg="cfg2"):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ref_namespace()
    n_frames = max(cores, 8) if cfg != "cfg5" else 16
    frames = cpu_sample_frames(n_frames, cfg)
    ctx = mp.get_context("fork")
    with ctx.Pool(min(cores, n_frames)) as pool:
        if cfg != "cfg5":
            cpu_reference(frames[:cores], pool, cores)
        n, dt = cpu_reference(frames, pool, cores)
    t0 = time.perf_counter()
    _cpu_frame(frames[0])
    one = time.perf_counter() - t0
    kind = cpu_kind()
    what = "the reference's unmodified voxel_filter + do_range_projection" if kind == "reference" else "the oracle *_loop port"
    out = {"value": n / dt, "unit": UNIT, "cores": min(cores, n_frames), "kind": kind,
           "sample": f"{n_frames} {cfg} frames ({n} points) through {what} with multiprocessing.Pool({min(cores, n_frames)}); "
                     f"single core: {len(frames[0][0]) / one:.3g} points/s",
           "single_core_value": len(frames[0][0]) / one}
    if cfg == "cfg5":
        out["extrapolated_768_frames_s"] = 768 * 1e6 / (n / dt)
    return out


def other_stages(device, rank, world, hbm_peak, args):
    """BEV pool fwd/bwd at cfg3 shapes and the IoU count reduction at cfg4 per-rank shapes (+ NCCL all-reduce)."""
    import torch
    import torch.distributed as dist
    import muvo_b200
    from muvo_b200 import synth
    from muvo_b200.metrics import ssc_counts, all_reduce_counts
    res = {}
    steps = max(3, min(args.steps, 10))

    def timed(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # (c) cfg3: B_f = 6, C = 384
    B, C, D, H, W = 6, 384, 37, 40, 104
    feat, depth, mask, K, E = synth.bev_inputs(B, C, 3000 + rank, device=device)
    fp = muvo_b200.FrustumPooling(**synth.BEV_POOL_ARGS).to(device)
    x = synth.lift(feat, depth)                                  # materialised once; the pool reads it in place
    fp.initialize_frustum(x)
    geom = fp.get_geometry(E[:, None, :3, :3], E[:, None, :3, 3:], K[:, None])
    cell = fp.cell_ids(geom, mask)
    n_kept = int((cell >= 0).sum())
    n_pts = D * H * W
    from muvo_b200.frustum_pooling import bev_pool
    xg = x.detach().requires_grad_(True)
    ms_f = timed(lambda: bev_pool(x, cell, 2304), steps)
    out = bev_pool(xg, cell, 2304)
    gout = torch.randn_like(out)
    ms_b = timed(lambda: torch.autograd.grad(out, xg, gout, retain_graph=True), steps)
    bytes_f = n_kept * C * 4 + B * n_pts * 5 + B * C * 2304 * 4
    bytes_f_sector = B * n_pts * C * 4 + B * n_pts * 4 + B * C * 2304 * 4
    bytes_b = B * C * 2304 * 4 + B * n_pts * C * 4
    res["bev_pool_fwd"] = {"ms": ms_f, "algorithmic_GBps": bytes_f / ms_f / 1e6, "frac": bytes_f / ms_f / 1e6 / hbm_peak,
                           "dense_read_GBps": bytes_f_sector / ms_f / 1e6, "frames": B, "C": C, "n_kept": n_kept,
                           "note": "kernel only (cell ids precomputed); dense_read = whole lifted tensor, the sector-granular bound for a random top-k mask"}
    res["bev_pool_bwd"] = {"ms": ms_b, "algorithmic_GBps": bytes_b / ms_b / 1e6, "frac": bytes_b / ms_b / 1e6 / hbm_peak}
    # the drop-in boundary itself: FrustumPooling.forward(x, intrinsics, pose, mask) (frustum_pooling.py:189-209); geometry and
    # cell ids come from the per-camera cache, the mask is folded in per call
    Kc, Ec = K[:, None].contiguous(), E[:, None].contiguous()
    fp(x, Kc, Ec, mask)
    ms_m = timed(lambda: fp(x, Kc, Ec, mask), steps)
    res["bev_module_fwd"] = {"ms": ms_m, "algorithmic_GBps": bytes_f / ms_m / 1e6, "frac": bytes_f / ms_m / 1e6 / hbm_peak,
                             "dense_read_GBps": bytes_f_sector / ms_m / 1e6, "dense_read_frac": bytes_f_sector / ms_m / 1e6 / hbm_peak,
                             "cache_hits": int(fp._geom_cache["hits"]),
                             "note": "muvo_b200.FrustumPooling.forward(x, K, E, mask): cached cell ids + mask fold + index sort + pool kernel"}
    xm = x.detach().requires_grad_(True)
    om = fp(xm, Kc, Ec, mask)
    ms_mb = timed(lambda: torch.autograd.grad(om, xm, gout.view(om.shape), retain_graph=True), steps)
    res["bev_module_bwd"] = {"ms": ms_mb, "algorithmic_GBps": bytes_b / ms_mb / 1e6, "frac": bytes_b / ms_mb / 1e6 / hbm_peak}
    res["bev_module_fwd"]["note"] = ("muvo_b200.FrustumPooling.forward(x, K, E, mask): cached cell ids + per-chunk plan, mask compaction "
                                     "(k_chunk_compact) + TMA-streamed pool (k_pool_stream)")
    del xm, om
    # the same module under MUVO's default precision ('16-mixed', config.py:40: the lifted tensor is fp16, 0.71 GB): reported
    # beside the fp32 figures, not instead of them
    x16 = x.detach().to(torch.float16)
    fp(x16, Kc, Ec, mask)
    ms_m16 = timed(lambda: fp(x16, Kc, Ec, mask), steps)
    x16g = x16.requires_grad_(True)
    om16 = fp(x16g, Kc, Ec, mask)
    ms_mb16 = timed(lambda: torch.autograd.grad(om16, x16g, gout.view(om16.shape).to(om16.dtype), retain_graph=True), steps)
    by16 = B * n_pts * C * 2
    res["bev_module_fp16"] = {"fwd_ms": ms_m16, "bwd_ms": ms_mb16, "fwd_dense_GBps": by16 / ms_m16 / 1e6, "bwd_dense_GBps": by16 / ms_mb16 / 1e6,
                              "fwd_frac": by16 / ms_m16 / 1e6 / hbm_peak, "bwd_frac": by16 / ms_mb16 / 1e6 / hbm_peak,
                              "note": "fp16 lifted tensor (the reference's default precision); consumer bound, DESIGN.md section 7"}
    del x16, x16g, om16
    # the reference's own FrustumPooling on the same inputs: stock PyTorch on this B200, and on the host cores (rank 0)
    ns = ref_namespace()
    if ns is not None and rank == 0:
        try:
            rfp = ns.FrustumPooling(**synth.BEV_POOL_ARGS).to(device)
            xr = x.detach().clone().requires_grad_(True)
            outr = rfp(xr, Kc, Ec, mask)
            err = float((outr.detach() - fp(x, Kc, Ec, mask)).abs().max() / outr.detach().abs().max())
            ms_rf = timed(lambda: rfp(x, Kc, Ec, mask), max(2, steps // 2))
            ms_rb = timed(lambda: torch.autograd.grad(rfp(xr, Kc, Ec, mask), xr, gout.view(outr.shape)), max(2, steps // 2)) - ms_rf
            res["bev_reference_gpu"] = {"fwd_ms": ms_rf, "bwd_ms": ms_rb, "kind": "reference",
                                        "what": "unmodified muvo.models.frustum_pooling.FrustumPooling, CUDA tensors, same B200 (stock PyTorch)",
                                        "max_rel_diff_vs_ours": err,
                                        "vs_reference": {"fwd": ms_rf / ms_m, "bwd": ms_rb / ms_mb}}
            del rfp, xr, outr
            torch.cuda.empty_cache()
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            rfc = ns.FrustumPooling(**synth.BEV_POOL_ARGS)
            xc = x.detach().cpu().requires_grad_(True)
            Kh, Eh, mh, gh = Kc.cpu(), Ec.cpu(), mask.cpu(), gout.cpu()
            rfc(xc.detach(), Kh, Eh, mh)
            t0 = time.perf_counter(); oc = rfc(xc, Kh, Eh, mh); t1 = time.perf_counter()
            torch.autograd.grad(oc, xc, gh.view(oc.shape)); t2 = time.perf_counter()
            res["bev_reference_cpu"] = {"fwd_ms": 1e3 * (t1 - t0), "bwd_ms": 1e3 * (t2 - t1), "cores": cores, "kind": "reference",
                                        "sample": f"one forward + backward of the unmodified FrustumPooling at cfg3 shapes on {cores} torch threads",
                                        "vs_reference": {"fwd": 1e3 * (t1 - t0) / ms_m, "bwd": 1e3 * (t2 - t1) / ms_mb}}
            del rfc, xc, oc
        except Exception as e:                                    # never lose the bench line to a baseline
            res["bev_reference_error"] = repr(e)[:200]
    # N2: fused lift-splat on the same inputs (feat, depth -> BEV), forward and backward to feat / depth
    from muvo_b200.frustum_pooling import lift_splat
    fl, dl = feat.detach().requires_grad_(True), depth.detach().requires_grad_(True)
    fp.lift_splat(fl.detach(), dl.detach(), Kc, Ec, mask)           # builds the per-camera plan (mask-independent cell sort)
    ms_lf = timed(lambda: fp.lift_splat(fl.detach(), dl.detach(), Kc, Ec, mask), steps)
    ol = fp.lift_splat(fl, dl, Kc, Ec, mask).reshape(B, C, 2304)
    ms_lb = timed(lambda: torch.autograd.grad(ol, (fl, dl), gout, retain_graph=True), steps)
    bytes_l = B * (C * H * W * 4 + n_pts * 9 + C * 2304 * 4)       # feat + depth + cell ids + out, each once
    res["lift_splat_fused_fwd"] = {"ms": ms_lf, "algorithmic_GBps": bytes_l / ms_lf / 1e6, "frac": bytes_l / ms_lf / 1e6 / hbm_peak,
                                   "note": "N2: FrustumPooling.lift_splat(feat, depth, K, E, mask): same output as lift (mile.py:517-521) + "
                                           "FrustumPooling.forward, outer product never materialised; includes the channels-last copy of "
                                           "feat, the mask fold and the mask filter of the cached cell sort"}
    bytes_lb = B * (2 * C * H * W * 4 + n_pts * (4 + 4 + 4) + C * 2304 * 4)      # gout + feat in, grad_feat out, depth + cell in, grad_depth out
    res["lift_splat_fused_bwd"] = {"ms": ms_lb, "algorithmic_GBps": bytes_lb / ms_lb / 1e6, "frac": bytes_lb / ms_lb / 1e6 / hbm_peak,
                                   "note": "grad_feat + grad_depth, includes the [B,cells,C] copy of grad_out"}
    del x, xg, out, gout, feat, depth, fl, dl, ol
    torch.cuda.empty_cache()
    # (d) cfg4: 16 frames per rank, C = 2 and C = 9, counts all-reduced over the ranks and checked against the single-process
    # oracle over ALL ranks' frames (every rank's inputs are seeded, so rank 0 regenerates them)
    for Cn in (2, 9):
        yp, yt = synth.occupancy_pair(16, Cn, 4000 + rank + 100 * (Cn != 2))
        tp, tt = torch.from_numpy(yp).to(device), torch.from_numpy(yt).to(device)
        acc = torch.zeros(3 + 3 * Cn, dtype=torch.int64, device=device)

        def ssc():
            acc.zero_()
            ssc_counts(tp, tt, Cn, ignore255=True, out=acc)
            all_reduce_counts(acc)
        ms_d = timed(ssc, steps)
        acc.zero_()
        ssc_counts(tp, tt, Cn, ignore255=True, out=acc)
        ms_k = timed(lambda: ssc_counts(tp, tt, Cn, ignore255=True, out=acc), steps)      # kernel alone, no collective
        ssc()
        got = acc.cpu().numpy()
        # the collective taken off the per-batch path: SSCMetrics(sync_dist="epoch") accumulates on the device and reduces
        # once (trainer.py:515-567 reads the statistics at epoch end only); 32 batches + the flush, per batch
        md = muvo_b200.SSCMetrics(Cn, sync_dist="epoch")

        def ssc_epoch():
            for _ in range(32):
                md.add_batch(tp, tt)
            md.get_stats()
        ms_e = timed(ssc_epoch, max(2, steps // 2)) / 32
        bytes_d = tp.numel() * 9
        key = "ssc_counts" if Cn == 2 else "ssc_counts_c9"
        # headline = the mode the trainer's call pattern needs (add_batch per validation step, statistics read at epoch end,
        # trainer.py:483-490 / 515-567): SSCMetrics(sync_dist="epoch"), one all-reduce per epoch; the per-batch all-reduce
        # (sync_dist=True: every rank's running statistics global after EVERY batch) is reported beside it
        res[key] = {"ms": ms_e, "kernel_only_ms": ms_k, "algorithmic_GBps": bytes_d / ms_e / 1e6, "frac": bytes_d / ms_e / 1e6 / hbm_peak,
                    "kernel_only_frac": bytes_d / ms_k / 1e6 / hbm_peak, "n_classes": Cn,
                    "mode": "SSCMetrics(sync_dist='epoch').add_batch per batch, 32 batches + one flush (all-reduce + read-back)",
                    "per_batch_allreduce_ms": ms_d, "per_batch_allreduce_frac": bytes_d / ms_d / 1e6 / hbm_peak,
                    "voxels_per_s": tp.numel() * world / (ms_e * 1e-3), "frames_per_rank": 16,
                    "allreduce": f"nccl int64[{3 + 3 * Cn}]" if world > 1 else "none (1 rank)"}
        if rank == 0:
            import oracle as O                                     # the checker, not the thing measured
            want = np.zeros(3 + 3 * Cn, dtype=np.int64)
            for r in range(world):
                ypr, ytr = (yp, yt) if r == 0 else synth.occupancy_pair(16, Cn, 4000 + r + 100 * (Cn != 2))
                want += O.ssc_add_batch_counts(ypr, ytr, Cn)
            res[key]["allreduced_counts_equal_oracle"] = bool(np.array_equal(got, want))
            res[key]["frames_checked"] = 16 * world
            assert np.array_equal(got, want), f"cfg4 C={Cn}: all-reduced counts differ from the oracle over {16 * world} frames"
            ns = ref_namespace()
            if ns is not None:
                try:
                    m = ns.SSCMetrics(Cn)
                    m.add_batch(tp[:2], tt[:2])                      # warm
                    m.reset()
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    m.add_batch(tp, tt)
                    torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
                    cores = os.cpu_count() or 1
                    torch.set_num_threads(cores)
                    mc = ns.SSCMetrics(Cn)
                    t0 = time.perf_counter()
                    mc.add_batch(torch.from_numpy(yp), torch.from_numpy(yt))
                    t_cpu = time.perf_counter() - t0
                    ref_counts = np.r_[mc.completion_tp, mc.completion_fp, mc.completion_fn, mc.tps.numpy(), mc.fps.numpy(), mc.fns.numpy()]
                    mine = O.ssc_add_batch_counts(yp, yt, Cn)
                    res[key]["reference"] = {"kind": "reference", "what": "unmodified muvo.metrics.SSCMetrics.add_batch, 16 frames",
                                             "gpu_ms": 1e3 * t_gpu, "cpu_ms": 1e3 * t_cpu, "cores": cores,
                                             "counts_equal": bool(np.array_equal(ref_counts.astype(np.int64), mine)),
                                             "vs_reference": {"gpu": 1e3 * t_gpu / ms_k, "cpu": 1e3 * t_cpu / ms_k}}
                except Exception as e:
                    res[key]["reference_error"] = repr(e)[:200]
    yp, yt = synth.occupancy_pair(16, 2, 4000 + rank)
    tp, tt = torch.from_numpy(yp).to(device), torch.from_numpy(yt).to(device)
    # cfg1 (BASELINE.json configs[0]): ONE 100 k-point frame, voxel_filter on one CPU core vs the kernel behind the same signature
    if rank == 0:
        p1, s1 = synth.carla_lidar_frame(100_000, 1000)
        grid1 = (0.5, [192, 192, 64], [0.0, 0, -10.0])
        muvo_b200.voxel_filter(p1, s1, *grid1)
        t0 = time.perf_counter()
        for _ in range(5):
            v1, l1 = muvo_b200.voxel_filter(p1, s1, *grid1)
        t_ours = (time.perf_counter() - t0) / 5
        from muvo_b200.points import GridSpec as _G, sensor_to_grid as _s2g
        dp1, ds1 = torch.from_numpy(p1).to(device), torch.from_numpy(s1).to(device)
        ms_k1 = timed(lambda: _s2g(dp1, ds1, None, grid=_G(), dense=False, sparse=True), steps)
        ns = ref_namespace()
        import oracle as O
        t0 = time.perf_counter()
        v0, l0 = (ns.voxel_filter if ns is not None else O.voxel_filter_loop)(p1, s1, *grid1)
        t_ref = time.perf_counter() - t0
        res["cfg1_voxel_filter"] = {"points": 100_000, "kernel_ms": ms_k1, "numpy_in_numpy_out_ms": 1e3 * t_ours,
                                    "cpu_ms": 1e3 * t_ref, "cpu_kind": cpu_kind(), "cores": 1,
                                    "equal": bool(np.array_equal(v1, v0) and np.array_equal(l1, l0)),
                                    "points_per_s": {"kernel": 1e5 / (ms_k1 * 1e-3), "drop_in_call": 1e5 / t_ours, "cpu": 1e5 / t_ref},
                                    "vs_reference": {"drop_in_call": t_ref / t_ours, "kernel": 1e3 * t_ref / ms_k1},
                                    "note": "data/generate_voxels.py voxelisation of one CARLA-style frame: voxel_filter(pcd, sem, 0.5, "
                                            "[192,192,64], [0,0,-10]) NumPy in / NumPy out (H2D + kernels + D2H) vs one CPU core"}
        del dp1, ds1
    # N4: SemScalLoss + GeoScalLoss on the voxel logits of the same 16 frames (C = 2, fp32), forward and backward
    from muvo_b200.losses import scal_losses
    gen = torch.Generator(device=device).manual_seed(4100 + rank)
    logits = torch.randn((1, 16, 2, 192, 192, 64), generator=gen, device=device).requires_grad_(True)
    tl = tt.view(1, 16, 192, 192, 64)
    ms_sf = timed(lambda: scal_losses(logits.detach(), tl), steps)
    sem, geo = scal_losses(logits, tl)
    ms_sb = timed(lambda: torch.autograd.grad(sem + geo, logits, retain_graph=True), steps)
    bytes_sf = tl.numel() * (2 * 4 + 1)
    bytes_sb = tl.numel() * (4 * 4 + 1)
    res["scal_losses_fwd"] = {"ms": ms_sf, "algorithmic_GBps": bytes_sf / ms_sf / 1e6, "frac": bytes_sf / ms_sf / 1e6 / hbm_peak,
                              "note": "N4: both losses from one pass over the logits (losses.py:191-287), includes the scalar epilogue"}
    res["scal_losses_bwd"] = {"ms": ms_sb, "algorithmic_GBps": bytes_sb / ms_sb / 1e6, "frac": bytes_sb / ms_sb / 1e6 / hbm_peak}
    return res


_REAL_STDOUT = None


def emit_line(line: dict) -> None:
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL's version banner, ...) print to fd 1: keep stdout clean for the single JSON line
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-stages", action="store_true", help="skip the BEV-pool / IoU stage numbers")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 = BASELINE.json configs[1] (the headline); cfg5 = configs[4], the 1M-point-frame scaling sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

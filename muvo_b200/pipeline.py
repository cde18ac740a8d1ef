"""Host-buffer front end for stages (a)+(b): the call a data-loading process makes.

``HostPipeline.submit(points_np, sem_np, frame_offsets_np)`` stages the ragged batch through pinned host
memory (on a worker thread: ``submit`` returns at once, so the caller's thread is free while the 100 MB staging copy
runs), runs the fused point kernels and copies the reference-facing results back into pinned host
buffers: the sparse ``(n,4) uint16`` voxel lists (what ``voxel_filter`` returns) and/or the dense grids,
and the range images (what ``do_range_projection`` returns).  Two slots are double-buffered over three
streams (H2D / kernels / D2H), so the copies of batch i+1 overlap the kernels of batch i and the
read-back of batch i-1; PCIe is full duplex.  ``result()`` blocks on the oldest in-flight batch.

Two things keep the host side off the critical path: the pageable -> pinned staging copy is split over host
threads (``muvo_host_copy``), and the sparse voxel lists are written back to back (``sparse_start``) so that the
read-back moves only about as many rows as there are occupied voxels instead of one row per input point.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch

from . import _lib
from .points import GridSpec, RangeSpec, sensor_to_grid


def _pinned_view(a, dtype):
    """``a`` itself when it is a contiguous pinned CPU tensor of ``dtype`` (usable as a DMA source), else None."""
    if isinstance(a, torch.Tensor) and a.device.type == "cpu" and a.dtype == dtype and a.is_contiguous() and a.is_pinned():
        return a
    return None


class _Slot:
    def __init__(self):
        self.cap_pts = 0
        self.cap_frames = 0
        self.h_pts = self.h_sem = self.h_off = None
        self.d_pts = self.d_sem = self.d_off = None
        self.dev_out = {}
        self.host_out = {}
        self.done = None
        self.busy = False
        self.meta = None
        self.future = None


class HostPipeline:
    def __init__(self, device=None, grid: Optional[GridSpec] = GridSpec(), range_spec: Optional[RangeSpec] = RangeSpec(),
                 dense: bool = False, sparse: bool = True, layout: str = "hwc", remap: Optional[np.ndarray] = None,
                 depth: int = 2, host_threads: int = 0, device_out: bool = False):
        """``device_out=True``: the results stay on the device (a training step consumes them there, as MUVO's batches end
        up on the GPU anyway); ``result()`` then returns device tensors that are valid until the slot's next ``submit``,
        and only the per-frame counts (``n_occ`` / ``sparse_start``) are read back."""
        if not torch.cuda.is_available():
            raise _lib.MuvoError("muvo_b200 kernels need a CUDA device (sm_100a); no CPU fallback exists")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MuvoError("HostPipeline needs a CUDA device")
        self.grid, self.range_spec = grid, range_spec
        self.dense, self.sparse, self.layout = dense, sparse, layout
        self.device_out = bool(device_out)
        self.remap = torch.from_numpy(np.ascontiguousarray(remap)).to(self.device) if remap is not None else None
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.slots = [_Slot() for _ in range(depth)]
        self.queue = []
        self._worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="muvo_b200_stage")   # keeps batches in order
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        if int(host_threads) <= 0:              # default: the cores of this box shared by the ranks running on it, capped at 16
            import os
            local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
            host_threads = max(1, min(16, (os.cpu_count() or 1) // max(local, 1)))
        self.host_threads = int(host_threads)
        self.row_cap = None                     # sparse rows read back per batch: 1.25 x the largest total seen so far

    def _ensure(self, slot: _Slot, n_pts: int, n_frames: int):
        if n_pts > slot.cap_pts:
            cap = max(n_pts, int(slot.cap_pts * 1.25), 1024)
            slot.h_pts = torch.empty((cap, 3), dtype=torch.float32).pin_memory()
            slot.h_sem = torch.empty((cap,), dtype=torch.uint8).pin_memory()
            slot.d_pts = torch.empty((cap, 3), dtype=torch.float32, device=self.device)
            slot.d_sem = torch.empty((cap,), dtype=torch.uint8, device=self.device)
            slot.cap_pts = cap
            slot.host_out.pop("voxel_sparse", None)
            if self.sparse and self.grid is not None:      # sized for the slot, so ragged batches reuse it (and its pinned twin)
                slot.dev_out["voxel_sparse"] = torch.empty((cap, 4), dtype=torch.int16, device=self.device)
            else:
                slot.dev_out.pop("voxel_sparse", None)
        if n_frames > slot.cap_frames:
            slot.h_off = torch.empty((n_frames + 1,), dtype=torch.int64).pin_memory()
            slot.d_off = torch.empty((n_frames + 1,), dtype=torch.int64, device=self.device)
            slot.cap_frames = n_frames
            slot.dev_out = {k: v for k, v in slot.dev_out.items() if k == "voxel_sparse"}
            slot.host_out = {k: v for k, v in slot.host_out.items() if k == "voxel_sparse"}

    def submit(self, points: np.ndarray, semantics: np.ndarray, frame_offsets: np.ndarray):
        """Enqueue one ragged batch (host arrays).  Returns immediately; see :meth:`result`.

        The staging copy and the CUDA enqueues run on the pipeline's worker thread, in submission order; the input arrays
        must stay unchanged until ``result()`` of this batch has returned."""
        slot = next((s for s in self.slots if not s.busy), None)
        if slot is None:
            raise RuntimeError("all pipeline slots are in flight; call result() first")
        fo = frame_offsets.numpy() if isinstance(frame_offsets, torch.Tensor) else np.asarray(frame_offsets)
        if fo.ndim != 1 or len(fo) < 1 or int(fo[0]) != 0 or int(fo[-1]) != int(points.shape[0]) or np.any(np.diff(fo) < 0):
            raise ValueError("frame_offsets must be non-decreasing, start at 0 and end at the number of points")
        slot.busy = True
        slot.future = self._worker.submit(self._stage_and_launch, slot, points, semantics, frame_offsets)
        self.queue.append(slot)

    def _stage_and_launch(self, slot: _Slot, points: np.ndarray, semantics: np.ndarray, frame_offsets: np.ndarray):
        torch.cuda.set_device(self.device)       # the worker thread starts on device 0 (pinned allocations follow the current device)
        n_pts, n_frames = int(points.shape[0]), int(len(frame_offsets) - 1)
        self._ensure(slot, n_pts, n_frames)
        # host -> pinned staging for ordinary pageable arrays; inputs that already live in pinned memory
        # (torch CPU tensors with is_pinned(), e.g. from pinned_inputs()) are the DMA source themselves
        lib = _lib.load()
        src_pts, src_sem = _pinned_view(points, torch.float32), _pinned_view(semantics, torch.uint8)
        if src_pts is None:
            pts_c = np.ascontiguousarray(points.numpy() if isinstance(points, torch.Tensor) else points, dtype=np.float32)
            _lib.check(lib.muvo_host_copy(slot.h_pts.data_ptr(), pts_c.ctypes.data, pts_c.nbytes, self.host_threads), "muvo_host_copy")
            src_pts = slot.h_pts[:n_pts]
        if src_sem is None:
            sem_c = np.ascontiguousarray((semantics.numpy() if isinstance(semantics, torch.Tensor) else semantics).reshape(-1), dtype=np.uint8)
            _lib.check(lib.muvo_host_copy(slot.h_sem.data_ptr(), sem_c.ctypes.data, sem_c.nbytes, self.host_threads), "muvo_host_copy")
            src_sem = slot.h_sem[:n_pts]
        slot.h_off[:n_frames + 1].numpy()[...] = frame_offsets.numpy() if isinstance(frame_offsets, torch.Tensor) else frame_offsets
        with torch.cuda.device(self.device):
            if slot.done is not None:
                self.s_in.wait_event(slot.done)           # previous read-back of this slot finished
            with torch.cuda.stream(self.s_in):
                slot.d_pts[:n_pts].copy_(src_pts.view(n_pts, 3), non_blocking=True)
                slot.d_sem[:n_pts].copy_(src_sem.view(n_pts), non_blocking=True)
                slot.d_off[:n_frames + 1].copy_(slot.h_off[:n_frames + 1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            h2d_bytes = n_pts * 13 + (n_frames + 1) * 8
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in)
                # reuse the slot's output tensors whose leading dimension still fits: per-point rows for the sparse list,
                # F + 1 rows for the packed starts, F rows for everything else
                want_rows = {"voxel_sparse": slot.cap_pts, "sparse_start": n_frames + 1}
                out = {k: v for k, v in slot.dev_out.items() if v.shape[0] == want_rows.get(k, n_frames)}
                res = sensor_to_grid(slot.d_pts[:n_pts], slot.d_sem[:n_pts], slot.d_off[:n_frames + 1], grid=self.grid,
                                     range_spec=self.range_spec, dense=self.dense, sparse=self.sparse, remap=self.remap,
                                     layout=self.layout, out=out, packed_sparse=self.sparse)
                ev_run = torch.cuda.Event()
                ev_run.record(self.s_run)
            keys = [k for k in res if k not in ("frame_offsets", "diag")]
            slot.dev_out = {k: res[k] for k in keys}
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run)
                nbytes = 0
                rows = n_pts if self.row_cap is None else min(n_pts, self.row_cap)
                for k in keys:
                    t = res[k]
                    if self.device_out and k not in ("n_occ", "sparse_start"):
                        continue                          # stays on the device
                    h = slot.host_out.get(k)
                    if h is None or h.shape != t.shape:
                        h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                        slot.host_out[k] = h
                    if k == "voxel_sparse":          # packed: only the leading rows can be in use
                        h[:rows].copy_(t[:rows], non_blocking=True)
                        nbytes += rows * t.shape[1] * t.element_size()
                    else:
                        h.copy_(t, non_blocking=True)
                        nbytes += t.numel() * t.element_size()
                slot.done = torch.cuda.Event()
                slot.done.record(self.s_out)
        slot.meta = (n_pts, n_frames, rows, h2d_bytes, nbytes)      # published by result() on the caller's thread

    def result(self) -> dict:
        """Blocks until the oldest submitted batch is back in host memory; returns its pinned host tensors.

        ``voxel_sparse`` rows of frame f are ``[sparse_start[f], sparse_start[f] + n_occ[f])`` (int16 storage,
        view as uint16).  The buffers are reused by the next ``submit`` on the same slot.
        """
        slot = self.queue.pop(0)
        try:
            slot.future.result()                 # staging + enqueue finished (re-raises what the worker raised)
        except BaseException:
            slot.busy = False
            raise
        slot.done.synchronize()
        n_pts, n_frames, rows, self.h2d_bytes, self.d2h_bytes = slot.meta
        if self.device_out:
            slot.busy = False
            out = dict(slot.dev_out)
            out.update(slot.host_out)                # n_occ / sparse_start: host copies
            return out
        if self.sparse and "sparse_start" in slot.host_out:
            total = int(slot.host_out["sparse_start"][n_frames].item())
            if total > rows:                         # more occupied voxels than the rows read back so far: top up
                with torch.cuda.device(self.device), torch.cuda.stream(self.s_out):
                    slot.host_out["voxel_sparse"][rows:total].copy_(slot.dev_out["voxel_sparse"][rows:total], non_blocking=True)
                    self.s_out.synchronize()
                self.d2h_bytes += (total - rows) * 8
            cap = max(1024, int(total * 1.25))
            self.row_cap = cap if self.row_cap is None else max(self.row_cap, cap)       # (read by the worker: a stale value only costs a top-up)
        slot.busy = False
        return dict(slot.host_out)

    @staticmethod
    def pinned_inputs(n_points: int, n_frames: int):
        """Pinned host tensors ``(points (n,3) f32, semantics (n,) u8, frame_offsets (F+1,) i64)`` for the caller to fill
        (``.numpy()`` gives writable views): batches submitted from them skip the staging copy."""
        return (torch.empty((n_points, 3), dtype=torch.float32).pin_memory(), torch.empty((n_points,), dtype=torch.uint8).pin_memory(),
                torch.empty((n_frames + 1,), dtype=torch.int64).pin_memory())

    def warmup(self, points: np.ndarray, semantics: np.ndarray, frame_offsets: np.ndarray):
        """Run one batch through EVERY slot so that all pinned / device buffers exist before timing-sensitive use
        (cudaHostAlloc of a 100 MB staging buffer costs tens of milliseconds)."""
        self.drain()
        for _ in range(len(self.slots)):
            self.submit(points, semantics, frame_offsets)
        self.drain()

    def drain(self):
        while self.queue:
            self.result()

    def close(self):
        """Finish what is in flight and stop the worker thread."""
        self.drain()
        self._worker.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

"""ctypes binding of ``libmuvo_b200.so`` (the C ABI declared in ``include/muvo_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call
returns an error, a ``MuvoError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmuvo_b200.so")

ABI_VERSION = 5
# dtype codes (include/muvo_b200.h)
F32, F64, F16, BF16 = 0, 1, 2, 3
I64, I32, U8, I16 = 0, 1, 2, 3
RANGE_LAYOUT_HWC, RANGE_LAYOUT_XYZD = 0, 1
DIAG_DROPPED_NONFINITE, DIAG_NEAR_EDGE_W, DIAG_NEAR_EDGE_H, DIAG_IN_GRID, DIAG_COUNT = 0, 1, 2, 3, 8


class MuvoError(RuntimeError):
    pass


class MuvoGrid(C.Structure):
    _fields_ = [("res", C.c_double), ("offset", C.c_double * 3), ("upper", C.c_double * 3),
                ("size", C.c_int32 * 3), ("roadline_id", C.c_int32)]


class MuvoRangeCfg(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("fov_down_abs", C.c_double), ("fov", C.c_double),
                ("lidar_pos", C.c_double * 3)]


class MuvoLidarPrep(C.Structure):
    _fields_ = [("add", C.c_double * 3), ("box_lo", C.c_double * 3), ("box_hi", C.c_double * 3),
                ("use_ego_box", C.c_int32), ("reserved", C.c_int32), ("remap256", C.c_void_p)]


_P = C.c_void_p
_I32, _I64, _SZ = C.c_int32, C.c_int64, C.c_size_t

# name -> (restype, argtypes); mirrors include/muvo_b200.h one to one
SIGNATURES = {
    "muvo_abi_version": (C.c_int, []),
    "muvo_strerror": (C.c_char_p, [C.c_int]),
    "muvo_profile_begin": (C.c_int, [_P]),
    "muvo_profile_end": (C.c_int, [_P, _I32, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.POINTER(_I32)]),
    "muvo_points_workspace_bytes": (C.c_int, [_I64, _I32, C.POINTER(MuvoGrid), C.POINTER(MuvoRangeCfg), C.POINTER(_SZ)]),
    "muvo_ws_reset": (C.c_int, [_P, _SZ, _P]),
    "muvo_merge_pcd_workspace_bytes": (C.c_int, [_I32, _I32, _I64, C.POINTER(_SZ)]),
    "muvo_merge_pcd": (C.c_int, [_P, _I32, _I32, C.c_double, C.c_double, C.POINTER(C.c_double), _P, _P, _I64, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), _P, _P, _P, _P, _SZ, _P]),
    "muvo_merge_pcd_at": (C.c_int, [_P, _I32, _I32, C.c_double, C.c_double, C.POINTER(C.c_double), _P, _P, _I64, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), _P, _P, _I64, _P, _I32, _P, _SZ, _P]),
    "muvo_label_pyramids": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "muvo_voxelize": (C.c_int, [_P, _I32, _P, _P, _I32, _I64, C.POINTER(MuvoGrid), _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "muvo_range_project": (C.c_int, [_P, _P, _P, _I32, _I64, C.POINTER(MuvoRangeCfg), _I32, _P, _P, _P, _P, _P, _SZ, _P]),
    "muvo_densify_sparse": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "muvo_range_project_lidar": (C.c_int, [_P, _P, _P, _I32, _I64, C.POINTER(MuvoRangeCfg), C.POINTER(MuvoLidarPrep), _I32, _P, _P, _P,
                                           _P, _P, _SZ, _P]),
    "muvo_points_fused": (C.c_int, [_P, _P, _P, _I32, _I64, C.POINTER(MuvoGrid), _P, C.POINTER(MuvoRangeCfg), _I32,
                                    _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "muvo_host_copy": (C.c_int, [_P, _P, _SZ, _I32]),
    "muvo_bev_pool_workspace_bytes": (C.c_int, [_I32, _I64, _I32, C.POINTER(_SZ)]),
    "muvo_bev_pool_max_cells": (C.c_int, []),
    "muvo_bev_fold_mask": (C.c_int, [_P, _P, _I64, _P, _P]),
    "muvo_bev_pool_fwd": (C.c_int, [_P, _I32, _I64, _I64, _I64, _P, _I32, _I64, _I32, _I32, _P, _P, _SZ, _P]),
    "muvo_bev_pool_fwd_masked": (C.c_int, [_P, _I32, _I64, _I64, _I64, _P, _P, _P, _P, _I32, _I64, _I32, _I32, _P, _P, _SZ, _P]),
    "muvo_bev_plan_bytes": (C.c_int, [_I32, _I64, C.POINTER(_SZ)]),
    "muvo_bev_plan_build": (C.c_int, [_P, _I32, _I64, _I32, _P, _SZ, _P]),
    "muvo_bev_pool_is_streamed": (C.c_int, [_I32, _P, _I64, _I64, _I64, _I32, _I64, _I32, _I32]),
    "muvo_bev_pool_bwd_streamed": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _P, _I32, _I64, _I64, _I64, _P, _SZ, _P]),
    "muvo_bev_pool_bwd": (C.c_int, [_P, _P, _I32, _I64, _I32, _I32, _P, _I32, _I64, _I64, _I64, _P]),
    "muvo_lift_splat_fwd": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _SZ, _P]),
    "muvo_lift_splat_plan_bytes": (C.c_int, [_I32, _I64, _I32, C.POINTER(_SZ)]),
    "muvo_lift_splat_plan_build": (C.c_int, [_P, _I32, _I64, _I32, _P, _SZ, _P, _SZ, _P]),
    "muvo_lift_splat_fwd_planned": (C.c_int, [_P, _P, _P, _SZ, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _SZ, _P]),
    "muvo_lift_splat_bwd": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "muvo_segment_sum_workspace_bytes": (C.c_int, [_I64, C.POINTER(_SZ)]),
    "muvo_segment_sum_fwd": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P, _P, _SZ, _P]),
    "muvo_segment_sum_bwd": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "muvo_ssc_counts": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _I64, _I32, _P, _P]),
    "muvo_ssc_counts_from_logits": (C.c_int, [_P, _I32, _P, _I32, _I32, _I64, _I32, _P, _P]),
    "muvo_scal_workspace_bytes": (C.c_int, [_I32, C.POINTER(_SZ)]),
    "muvo_scal_sums_fwd": (C.c_int, [_P, _I32, _P, _I32, _I32, _I64, _I32, _P, _P, _P, _SZ, _P]),
    "muvo_scal_sums_bwd": (C.c_int, [_P, _I32, _P, _I32, _I32, _I64, _I32, _P, _P, _P, _P, _P]),
    "muvo_pillar_workspace_bytes": (C.c_int, [_I64, _I32, C.POINTER(_SZ)]),
    "muvo_pillar_scatter_mean": (C.c_int, [_P, _P, _I32, _I64, _I32, _I64, _P, _P, _P, _SZ, _P, _P]),
    "muvo_pillar_scatter_mean_bwd": (C.c_int, [_P, _P, _I32, _P, _I64, _I32, _I64, _P, _P]),
    "muvo_pillar_scatter_max": (C.c_int, [_P, _P, _I32, _I64, _I32, _I64, _P, _P, _P, _SZ, _P, _P]),
    "muvo_pillar_scatter_max_bwd": (C.c_int, [_P, _P, _I32, _P, _I64, _I32, _I64, _P, _P]),
    "muvo_debug_pixel_check": (C.c_int, [_P, _I64, C.POINTER(MuvoRangeCfg), _P, _P]),
    "muvo_debug_set_tuning": (C.c_int, [_I32, _I32]),
    "muvo_debug_mega_stats": (C.c_int, [_P, _I32, _I32]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load (once) and return the shared library; raises ``MuvoError`` when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MuvoError(f"{LIB_PATH} is missing: build it with `python -m muvo_b200.build` "
                            "(there is no CPU/PyTorch fallback for the muvo_b200 kernels)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError -> stale library
            fn.restype = res
            fn.argtypes = args
        if lib.muvo_abi_version() != ABI_VERSION:
            raise MuvoError(f"libmuvo_b200.so ABI {lib.muvo_abi_version()} != expected {ABI_VERSION}; rebuild")
        # MUVO_TUNING="key=value,..." : debug / benchmarking knobs of muvo_debug_set_tuning, applied once at load
        for kv in filter(None, os.environ.get("MUVO_TUNING", "").replace(" ", ",").split(",")):
            k, v = kv.split("=")
            lib.muvo_debug_set_tuning(int(k), int(v))
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().muvo_strerror(rc).decode()
        raise MuvoError(f"{what or 'muvo call'} failed: {msg} (code {rc})")


def ptr(t) -> int | None:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr() if t.numel() > 0 else None


def require_cuda(*tensors) -> None:
    import torch
    if not torch.cuda.is_available():
        raise MuvoError("muvo_b200 kernels need a CUDA device (sm_100a); no CPU fallback exists")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MuvoError("expected CUDA tensors")


def current_stream(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class profile:
    """``with _lib.profile(stream) as p: ...`` -> ``p.kernels`` = [(kernel name, milliseconds), ...] per launch."""

    def __init__(self, stream: int):
        self.stream = stream
        self.kernels = []

    def __enter__(self):
        check(load().muvo_profile_begin(self.stream), "muvo_profile_begin")
        return self

    def __exit__(self, *exc):
        cap = 64
        ms = (C.c_float * cap)()
        names = (C.c_char_p * cap)()
        n = _I32(0)
        rc = load().muvo_profile_end(self.stream, cap, ms, names, C.byref(n))
        if exc[0] is None:
            check(rc, "muvo_profile_end")
        self.kernels = [(names[i].decode(), float(ms[i])) for i in range(n.value)]
        return False

"""Self-cleaning device workspaces, one per (device, stream).

The point kernels leave their tables zeroed on success (include/muvo_b200.h), so a
workspace is zero-filled only when it is (re)allocated or after a failed call.
"""
from __future__ import annotations

import torch

from . import _lib

_cache: dict = {}


_signature: dict = {}


def get(nbytes: int, device: torch.device, stream: int, signature=None) -> torch.Tensor:
    """Workspace of at least ``nbytes``.  ``signature`` = everything the table layout depends on besides the
    number of points (frames, grid size, image size); when it changes the buffer is zero-filled again."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    buf = _cache.get(key)
    if buf is not None and buf.numel() >= nbytes and _signature.get(key) != signature:
        _lib.check(_lib.load().muvo_ws_reset(buf.data_ptr(), buf.numel(), stream), "muvo_ws_reset")
    _signature[key] = signature
    if buf is None or buf.numel() < nbytes:
        # grow geometrically so ragged batches do not reallocate every call
        size = max(int(nbytes), int(buf.numel() * 3 // 2) if buf is not None else 0, 1 << 20)
        buf = torch.empty(size, dtype=torch.uint8, device=device)
        _lib.check(_lib.load().muvo_ws_reset(buf.data_ptr(), buf.numel(), stream), "muvo_ws_reset")
        _cache[key] = buf
    return buf


def invalidate(device: torch.device, stream: int) -> None:
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    _cache.pop(key, None)
    _signature.pop(key, None)


def clear() -> None:
    _cache.clear()
    _signature.clear()

""""Next" row N4: ``SemScalLoss`` / ``GeoScalLoss`` with their reductions on the GPU.

Drop-in for ``muvo.losses.SemScalLoss`` and ``muvo.losses.GeoScalLoss`` (muvo/losses.py:191-287, called from
muvo/trainer.py:375-382): same constructor, same ``forward(prediction, target)`` with ``prediction (b,s,c,x,y,z)``
logits and ``target (b,s,[1,]x,y,z)`` labels, same scalar result, differentiable in ``prediction``.

The reference materialises the softmax and then, per class, boolean-indexes it and runs half a dozen full
reductions.  Here one streaming kernel reduces the grid to ``3C+1`` float64 scalars (softmax in registers), a
single-thread epilogue evaluates both losses and their derivatives with the reference's conditions, and one more
streaming kernel writes ``d loss / d prediction``.  :func:`scal_losses` returns both losses from ONE such pass;
the two module classes call it and pick theirs.  No CPU fallback: without the CUDA library the call raises.

Differences, stated: sums are float64 (the reference's are fp32 ``torch.sum``), so results agree to ~1e-6 relative,
not bitwise; where the reference raises (``ZeroDivisionError`` when no class is present, the BCE range assertion on a
NaN ratio) this returns NaN instead of synchronising with the device to find out.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.nn as nn

from . import _lib

_LOGIT_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _flatten(prediction: torch.Tensor, target: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int, int, int]:
    if prediction.dim() < 3:
        raise ValueError("prediction must be (b, s, c, ...)")
    b, s, c = (int(v) for v in prediction.shape[:3])
    S = 1
    for v in prediction.shape[3:]:
        S *= int(v)
    if target.numel() != b * s * S:
        raise ValueError("prediction / target shape mismatch")
    if prediction.dtype not in _LOGIT_DTYPES:
        raise TypeError(f"unsupported prediction dtype {prediction.dtype}")
    logits = prediction.contiguous().view(b * s, c, S)
    tgt = target
    if tgt.dtype == torch.bool:
        tgt = tgt.view(torch.uint8)
    elif tgt.dtype != torch.uint8:
        tgt = tgt.to(torch.uint8)                    # labels are 0..255 (dataset.py:317-327 stores uint8)
    return logits, tgt.contiguous().view(b * s, S), b * s, c, S


class _ScalLosses(torch.autograd.Function):
    """(logits [F,C,S], target [F,S] u8) -> (sem_scal, geo_scal, sums[3C+1] f64); backward = one streaming kernel."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        _lib.require_cuda(logits, target)
        F, n_cls, S = (int(v) for v in logits.shape)
        dev = logits.device
        lib = _lib.load()
        need = C.c_size_t(0)
        _lib.check(lib.muvo_scal_workspace_bytes(n_cls, C.byref(need)), "muvo_scal_workspace_bytes")
        ws = torch.empty(int(need.value), dtype=torch.uint8, device=dev)         # per-CTA partial sums (< 1 MB)
        sums = torch.empty(3 * n_cls + 1, dtype=torch.float64, device=dev)
        losses = torch.empty(2 + 4 * n_cls, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = lib.muvo_scal_sums_fwd(_lib.ptr(logits), _LOGIT_DTYPES[logits.dtype], _lib.ptr(target), F, n_cls, S, int(ignore_index),
                                        sums.data_ptr(), losses.data_ptr(), ws.data_ptr(), ws.numel(), _lib.current_stream(dev))
        _lib.check(rc, "muvo_scal_sums_fwd")
        ctx.save_for_backward(logits, target, losses)
        ctx.ignore_index = int(ignore_index)
        ctx.mark_non_differentiable(sums)
        return losses[0].float(), losses[1].float(), sums

    @staticmethod
    def backward(ctx, g_sem, g_geo, _g_sums):
        logits, target, losses = ctx.saved_tensors
        F, n_cls, S = (int(v) for v in logits.shape)
        dev = logits.device
        grad = torch.empty_like(logits)
        g_sem = g_sem.detach().float().contiguous() if g_sem is not None else None
        g_geo = g_geo.detach().float().contiguous() if g_geo is not None else None
        with torch.cuda.device(dev):
            rc = _lib.load().muvo_scal_sums_bwd(_lib.ptr(logits), _LOGIT_DTYPES[logits.dtype], _lib.ptr(target), F, n_cls, S,
                                                ctx.ignore_index, losses.data_ptr() + 16, _lib.ptr(g_sem), _lib.ptr(g_geo),
                                                grad.data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "muvo_scal_sums_bwd")
        return grad, None, None


def scal_losses(prediction: torch.Tensor, target: torch.Tensor, ignore_index: int = 255) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(SemScalLoss()(prediction, target), GeoScalLoss()(prediction, target))`` from one pass over the logits."""
    logits, tgt, _, _, _ = _flatten(prediction, target)
    ign = int(ignore_index) if 0 <= int(ignore_index) <= 255 else -1
    sem, geo, _ = _ScalLosses.apply(logits, tgt, ign)
    return sem, geo


def scal_sums(prediction: torch.Tensor, target: torch.Tensor, ignore_index: int = 255) -> torch.Tensor:
    """Device float64 ``[3C+1]``: sum_p[C], nom[C], cnt[C], n_valid (no autograd)."""
    logits, tgt, _, _, _ = _flatten(prediction.detach(), target)
    ign = int(ignore_index) if 0 <= int(ignore_index) <= 255 else -1
    return _ScalLosses.apply(logits, tgt, ign)[2]


class SemScalLoss(nn.Module):
    """muvo/losses.py:191-251."""

    def __init__(self, ignore_index=255):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, prediction, target):
        return scal_losses(prediction, target, self.ignore_index)[0]


class GeoScalLoss(nn.Module):
    """muvo/losses.py:254-287."""

    def __init__(self, ignore_index=255):
        super().__init__()
        self.ignore_index = ignore_index

    def forward(self, prediction, target):
        return scal_losses(prediction, target, self.ignore_index)[1]

"""Stage (d): ``SSCMetrics`` with the count reduction on the GPU.

Drop-in for ``muvo.metrics.SSCMetrics`` (muvo/metrics.py:47-216): same methods,
same return types (Python ints for completion, CPU ``int32[C]`` tensors per class),
same accumulate / compute / get_stats arithmetic.  The (3 + 3C) masked passes per
frame of the reference collapse into one kernel launch and ONE device->host copy of
``3 + 3C`` int64 per call.  Under ``torch.distributed`` the counts can be
all-reduced (NCCL over NVLink) before they are accumulated -- functionality the
reference lacks (its metric object is rank-local).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_PRED_DTYPES = {torch.int64: _lib.I64, torch.int32: _lib.I32, torch.uint8: _lib.U8, torch.int16: _lib.I16}
_LOGIT_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _mask_u8(m: Optional[torch.Tensor], like: torch.Tensor) -> Optional[torch.Tensor]:
    if m is None:
        return None
    m = m.to(device=like.device)
    if m.dtype == torch.bool:
        m = m.contiguous().view(torch.uint8)
    elif m.dtype != torch.uint8:
        m = (m == 1).contiguous().view(torch.uint8)      # the reference tests `nonempty_idx == 1` (:165, :203)
    m = m.reshape(-1).contiguous()
    if m.numel() != like.numel():
        raise ValueError("mask shape mismatch")
    return m


def ssc_counts(predict: torch.Tensor, target: torch.Tensor, n_classes: int, nonempty: Optional[torch.Tensor] = None,
               nonsurface: Optional[torch.Tensor] = None, ignore255: bool = False,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device int64 ``[3 + 3C]``: completion (tp,fp,fn), tp[C], fp[C], fn[C].  Asynchronous.

    ``nonempty`` restricts both count families, ``nonsurface`` completion only, ``ignore255`` drops
    ``target == 255`` voxels (``add_batch`` semantics, muvo/metrics.py:77-100).  ``out`` accumulates.
    """
    _lib.require_cuda(predict, target)
    if predict.dtype not in _PRED_DTYPES:
        raise TypeError(f"unsupported prediction dtype {predict.dtype}")
    if target.dtype != torch.uint8:
        raise TypeError("target must be uint8")
    if predict.numel() != target.numel():
        raise ValueError("predict / target shape mismatch")
    dev = predict.device
    p = predict.contiguous().reshape(-1)
    t = target.contiguous().reshape(-1)
    ne, ns = _mask_u8(nonempty, t), _mask_u8(nonsurface, t)
    C = int(n_classes)
    if out is None:
        out = torch.zeros(3 + 3 * C, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().muvo_ssc_counts(_lib.ptr(p), _PRED_DTYPES[p.dtype], _lib.ptr(t), _lib.ptr(ne), _lib.ptr(ns),
                                         1 if ignore255 else 0, p.numel(), C, out.data_ptr(), _lib.current_stream(dev))
    _lib.check(rc, "muvo_ssc_counts")
    return out


def ssc_counts_from_logits(logits: torch.Tensor, target: torch.Tensor, ignore255: bool = True,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused ``argmax(logits, 1)`` + counts (muvo/trainer.py:483-490) without the int64 prediction.

    logits ``(F, C, ...)`` float32/16/bf16, target ``(F, ...)`` uint8.
    """
    _lib.require_cuda(logits, target)
    if logits.dtype not in _LOGIT_DTYPES:
        raise TypeError(f"unsupported logits dtype {logits.dtype}")
    if target.dtype != torch.uint8:
        raise TypeError("target must be uint8")
    F, C = int(logits.shape[0]), int(logits.shape[1])
    S = logits[0, 0].numel()
    if target.numel() != F * S:
        raise ValueError("logits / target shape mismatch")
    dev = logits.device
    lg = logits.contiguous()
    t = target.contiguous().reshape(-1)
    if out is None:
        out = torch.zeros(3 + 3 * C, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().muvo_ssc_counts_from_logits(_lib.ptr(lg), _LOGIT_DTYPES[lg.dtype], _lib.ptr(t), F, C, S,
                                                     1 if ignore255 else 0, out.data_ptr(), _lib.current_stream(dev))
    _lib.check(rc, "muvo_ssc_counts_from_logits")
    return out


def all_reduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the int64 count vector over ranks (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class SSCMetrics:
    """Same interface as the reference class (muvo/metrics.py:47-216).

    ``sync_dist=True`` all-reduces the counts of every ``add_batch`` across ranks (NCCL), so that each rank's running
    statistics are the global ones after every batch, exactly as if one process had seen all frames.
    ``sync_dist="epoch"`` keeps the collective and the device->host copy off the per-batch path: counts accumulate in an
    int64 vector ON THE DEVICE (``add_batch`` neither synchronises nor communicates), and the first ``get_stats()`` /
    ``compute()`` afterwards does ONE all-reduce and ONE read-back -- the reference only looks at the statistics in
    ``on_validation_epoch_end`` (muvo/trainer.py:515-567).  In that mode the statistics come from the exact int64 totals
    (the reference's float32 running sums round once a count passes 2^24; ``counts_exact`` holds the integers).
    """

    def __init__(self, n_classes, sync_dist=False, process_group=None):
        self.n_classes = n_classes
        self.sync_dist = sync_dist
        self.process_group = process_group
        self._dev_acc = None            # "epoch" mode: per-rank int64[3+3C] on the device
        self._pending = False
        self.reset()

    @property
    def deferred(self) -> bool:
        return self.sync_dist == "epoch"

    # -- kernels -----------------------------------------------------------------
    def _counts(self, predict, target, nonempty, nonsurface, ignore255):
        c = ssc_counts(predict, target, self.n_classes, nonempty, nonsurface, ignore255)
        if self.sync_dist and not self.deferred:
            all_reduce_counts(c, self.process_group)
        return c.cpu()          # the single device->host copy of this call

    def get_score_completion(self, predict, target, nonempty=None):
        """(tp, fp, fn) Python ints, summed over frames (muvo/metrics.py:143-176)."""
        c = self._counts(predict, target, nonempty, None, False)
        return int(c[0]), int(c[1]), int(c[2])

    def get_score_semantic_and_completion(self, predict, target, nonempty=None):
        """Three CPU ``int32[C]`` tensors tp, fp, fn (muvo/metrics.py:178-216)."""
        c = self._counts(predict, target, nonempty, None, False)
        C = self.n_classes
        return (c[3:3 + C].to(torch.int32), c[3 + C:3 + 2 * C].to(torch.int32), c[3 + 2 * C:3 + 3 * C].to(torch.int32))

    # -- accumulate / report: arithmetic identical to the reference ---------------
    def add_batch(self, y_pred, y_true, nonempty=None, nonsurface=None):
        """muvo/metrics.py:77-100, one launch: completion uses nonempty & nonsurface, classes nonempty only."""
        self.count += 1
        if self.deferred:
            self._dev_acc = self._acc_on(y_pred.device)
            ssc_counts(y_pred, y_true, self.n_classes, nonempty, nonsurface, True, out=self._dev_acc)   # adds into out
            self._pending = True
            return
        c = self._counts(y_pred, y_true, nonempty, nonsurface, True)
        self._accumulate(c)

    def _acc_on(self, dev):
        if self._dev_acc is None or self._dev_acc.device != dev:
            self._dev_acc = torch.zeros(3 + 3 * self.n_classes, dtype=torch.int64, device=dev)
        return self._dev_acc

    def _flush(self):
        """"epoch" mode: one all-reduce + one read-back of everything accumulated since the last flush / reset."""
        if not self._pending:
            return
        c = self._dev_acc.clone()
        self._dev_acc.zero_()
        self._pending = False
        all_reduce_counts(c, self.process_group)
        c = c.cpu()
        C = self.n_classes
        self.counts_exact += c
        t = self.counts_exact
        self.completion_tp, self.completion_fp, self.completion_fn = int(t[0]), int(t[1]), int(t[2])
        self.tps = t[3:3 + C].to(torch.float32)
        self.fps = t[3 + C:3 + 2 * C].to(torch.float32)
        self.fns = t[3 + 2 * C:3 + 3 * C].to(torch.float32)

    def add_batch_from_logits(self, logits, y_true):
        """``add_batch(argmax(logits, 1), y_true)`` without materialising the prediction (trainer.py:483-490)."""
        self.count += 1
        if self.deferred:
            self._dev_acc = self._acc_on(logits.device)
            ssc_counts_from_logits(logits, y_true, True, out=self._dev_acc)
            self._pending = True
            return
        c = ssc_counts_from_logits(logits, y_true, True)
        if self.sync_dist:
            all_reduce_counts(c, self.process_group)
        self._accumulate(c.cpu())

    def _accumulate(self, c):
        C = self.n_classes
        self.completion_tp += int(c[0])
        self.completion_fp += int(c[1])
        self.completion_fn += int(c[2])
        self.tps += c[3:3 + C].to(torch.int32)            # int32 -> float32 running sums, as :96-98
        self.fps += c[3 + C:3 + 2 * C].to(torch.int32)
        self.fns += c[3 + 2 * C:3 + 3 * C].to(torch.int32)
        self.counts_exact += c                             # exact int64 totals (extra; the reference keeps float32)
        self.compute()

    def compute(self):
        if self.deferred:
            self._flush()
        if self.completion_tp != 0:
            self.precision = self.completion_tp / (self.completion_tp + self.completion_fp)
            self.recall = self.completion_tp / (self.completion_tp + self.completion_fn)
            self.iou = self.completion_tp / (self.completion_tp + self.completion_fp + self.completion_fn)
        else:
            self.precision, self.recall, self.iou = 0, 0, 0
        self.iou_ssc = self.tps / (self.tps + self.fps + self.fns + 1e-5)

    def get_stats(self):
        if self.deferred and self._pending:
            self.compute()
        return {
            "precision": self.precision,
            "recall": self.recall,
            "iou": self.iou,
            "iou_ssc": self.iou_ssc,
            "iou_ssc_mean": torch.mean(self.iou_ssc[1:]),
        }

    def reset(self):
        if self._dev_acc is not None:
            self._dev_acc.zero_()
        self._pending = False
        self.completion_tp = 0
        self.completion_fp = 0
        self.completion_fn = 0
        self.tps = torch.zeros(self.n_classes)
        self.fps = torch.zeros(self.n_classes)
        self.fns = torch.zeros(self.n_classes)
        self.counts_exact = torch.zeros(3 + 3 * self.n_classes, dtype=torch.int64)

        self.hist_ssc = torch.zeros((self.n_classes, self.n_classes))
        self.labeled_ssc = 0
        self.correct_ssc = 0

        self.precision = 0
        self.recall = 0
        self.iou = 0
        self.count = 1e-8
        self.iou_ssc = torch.zeros(self.n_classes, dtype=torch.float32)
        self.cnt_class = torch.zeros(self.n_classes, dtype=torch.float32)

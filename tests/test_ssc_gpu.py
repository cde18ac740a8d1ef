"""GPU parity for stage (d): SSC / IoU counts.  Integer results -> bit-exact."""
import numpy as np
import pytest
import torch

import muvo_b200
import oracle as O
from muvo_b200 import synth
from muvo_b200.metrics import ssc_counts, ssc_counts_from_logits

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_counts_and_stats(golden, lib):
    g = golden("ssc.npz")
    for C in (2, 9):
        yp, yt = g[f"c{C}_pred"].astype(np.int64), g[f"c{C}_true"]
        m = muvo_b200.SSCMetrics(C)
        tp, fp, fn = m.get_score_completion(cu(yp), cu(yt))
        assert isinstance(tp, int) and [tp, fp, fn] == g[f"c{C}_raw"][:3].tolist()
        a, b, c = m.get_score_semantic_and_completion(cu(yp), cu(yt))
        assert a.dtype == torch.int32 and a.device.type == "cpu" and a.shape == (C,)
        assert np.array_equal(np.r_[a.numpy(), b.numpy(), c.numpy()], g[f"c{C}_raw"][3:])
        ne, ns = g[f"c{C}_nonempty"], g[f"c{C}_nonsurface"]
        tp, fp, fn = m.get_score_completion(cu(yp), cu(yt), cu(ne))
        a, b, c = m.get_score_semantic_and_completion(cu(yp), cu(yt), cu(ne))
        assert np.array_equal(np.r_[tp, fp, fn, a.numpy(), b.numpy(), c.numpy()], g[f"c{C}_masked"])
        m.reset()
        m.add_batch(cu(yp), cu(yt))
        m.add_batch(cu(yp), cu(yt), cu(ne), cu(ns))
        st = m.get_stats()
        assert st["iou"] == float(g[f"c{C}_iou"]) and st["precision"] == float(g[f"c{C}_precision"])
        assert st["recall"] == float(g[f"c{C}_recall"])
        assert np.array_equal(st["iou_ssc"].numpy(), g[f"c{C}_iou_ssc"])
        acc = np.r_[m.completion_tp, m.completion_fp, m.completion_fn, m.tps.numpy(), m.fps.numpy(), m.fns.numpy()]
        assert np.array_equal(acc.astype(np.float64), g[f"c{C}_acc2"])


@pytest.mark.parametrize("C", [2, 9, 23, 40])
def test_counts_vs_oracle(C, lib):
    yp, yt = synth.occupancy_pair(2, min(C, 23), 4100 + C, size=(96, 96, 32))
    rng = np.random.default_rng(C)
    yp = rng.integers(-1, C + 2, yp.shape, dtype=np.int64)          # includes out-of-range predictions
    ne, ns = rng.random(yt.shape) < 0.9, rng.random(yt.shape) < 0.5
    for kw in (dict(), dict(ignore255=True), dict(nonempty=ne), dict(nonempty=ne, nonsurface=ns, ignore255=True)):
        want = O.ssc_counts(yp, yt, C, **kw)
        tkw = {k: (cu(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        got = ssc_counts(cu(yp), cu(yt), C, **tkw).cpu().numpy()
        assert np.array_equal(got, want), kw


def test_pred_dtypes_ragged_tail_and_accumulate(lib):
    rng = np.random.default_rng(3)
    n = 1_000_003                      # not a multiple of the 512-voxel tile, odd -> unaligned pair loads
    yt = rng.integers(0, 9, n).astype(np.uint8)
    yt[rng.random(n) < 0.01] = 255
    yp = rng.integers(0, 9, n)
    want = O.ssc_counts(yp, yt, 9, ignore255=True)
    for dt in (torch.int64, torch.int32, torch.int16, torch.uint8):
        got = ssc_counts(cu(yp).to(dt), cu(yt), 9, ignore255=True)
        assert np.array_equal(got.cpu().numpy(), want)
    acc = torch.zeros(3 + 27, dtype=torch.int64, device="cuda")
    ssc_counts(cu(yp)[1:], cu(yt)[1:], 9, ignore255=True, out=acc)      # misaligned views
    ssc_counts(cu(yp)[:1], cu(yt)[:1], 9, ignore255=True, out=acc)
    assert np.array_equal(acc.cpu().numpy(), want)
    assert ssc_counts(cu(yp)[:0], cu(yt)[:0], 9).sum() == 0             # empty input


def test_from_logits_matches_argmax_path(lib):
    g = torch.Generator(device="cuda").manual_seed(5)
    F, C, S = 3, 9, (48, 40, 16)
    logits = torch.randn((F, C) + S, generator=g, device="cuda")
    logits[0, 3] = logits[0, 5]                                         # ties -> first maximum
    yp, yt = synth.occupancy_pair(F, C, 4200, size=S)
    want = ssc_counts(torch.argmax(logits, 1), cu(yt), C, ignore255=True)
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        lg = logits.to(dt)
        want = ssc_counts(torch.argmax(lg, 1), cu(yt), C, ignore255=True)
        got = ssc_counts_from_logits(lg, cu(yt), ignore255=True)
        assert torch.equal(got, want)
    m1, m2 = muvo_b200.SSCMetrics(C), muvo_b200.SSCMetrics(C)
    m1.add_batch(torch.argmax(logits, 1), cu(yt))
    m2.add_batch_from_logits(logits, cu(yt))
    assert torch.equal(m1.counts_exact, m2.counts_exact)


def test_full_size_properties(lib):
    """cfg4 per-rank shape (16 frames of 192x192x64): oracle on the whole thing + additivity over frames."""
    C = 2
    yp, yt = synth.occupancy_pair(16, C, 4000)
    tp, tt = cu(yp), cu(yt)
    total = ssc_counts(tp, tt, C, ignore255=True)
    parts = sum(ssc_counts(tp[i:i + 4], tt[i:i + 4], C, ignore255=True) for i in range(0, 16, 4))
    assert torch.equal(total, parts)
    c = total.cpu().numpy()
    valid = yt != 255
    assert c[3:5].sum() + c[5:7].sum() == valid.sum()                   # every valid voxel is a tp or an fp of its prediction
    assert c[0] + c[2] == ((yt > 0) & valid).sum()                      # completion tp + fn = occupied ground truth
    assert np.array_equal(c, O.ssc_add_batch_counts(yp, yt, C))
    assert torch.equal(total, ssc_counts(tp, tt, C, ignore255=True))    # deterministic


def test_epoch_sync_accumulates_on_device_and_matches_exact_totals(lib):
    """``sync_dist="epoch"``: add_batch neither copies to the host nor communicates; get_stats() flushes once
    (muvo/trainer.py:515-567 reads the statistics at epoch end only).  Totals = the oracle's over all batches."""
    C = 9
    m = muvo_b200.SSCMetrics(C, sync_dist="epoch")
    want = np.zeros(3 + 3 * C, dtype=np.int64)
    for k in range(3):
        yp, yt = synth.occupancy_pair(2, C, 4200 + k, size=(48, 48, 16))
        m.add_batch(cu(yp), cu(yt))
        want += O.ssc_add_batch_counts(yp, yt, C)
    assert m._pending and int(m.counts_exact.sum()) == 0            # nothing has left the device yet
    st = m.get_stats()
    assert np.array_equal(m.counts_exact.numpy(), want)
    tp, fp, fn = want[0], want[1], want[2]
    assert st["iou"] == tp / (tp + fp + fn) and st["precision"] == tp / (tp + fp) and st["recall"] == tp / (tp + fn)
    tps, fps, fns = (torch.from_numpy(want[3 + i * C:3 + (i + 1) * C]).float() for i in range(3))
    assert torch.equal(st["iou_ssc"], tps / (tps + fps + fns + 1e-5))
    yp, yt = synth.occupancy_pair(1, C, 4300, size=(48, 48, 16))      # a later batch keeps accumulating after a flush
    m.add_batch(cu(yp), cu(yt))
    m.get_stats()
    assert np.array_equal(m.counts_exact.numpy(), want + O.ssc_add_batch_counts(yp, yt, C))
    m.reset()
    assert int(m.counts_exact.sum()) == 0 and not m._pending

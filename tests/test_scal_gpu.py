"""GPU parity for "next" row N4: SemScalLoss / GeoScalLoss (muvo/losses.py:191-287).

Floating point: the reference sums in fp32, the kernel in float64, so the bar is relative 1e-5 on the losses and
1e-5 of the largest gradient entry on d loss / d logits (golden = the unmodified reference on CPU), and 1e-9 on the
3C+1 sums against the float64 oracle."""
import numpy as np
import pytest
import torch

import muvo_b200
import oracle as O
from muvo_b200.losses import scal_losses, scal_sums

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_losses_and_grads(golden, lib):
    g = golden("scal.npz")
    for C in (2, 5, 9):                              # 2, 9: register kernels; 5: any-C kernel
        pred, tgt = g[f"c{C}_pred"], g[f"c{C}_target"]
        for name, cls in (("sem", muvo_b200.SemScalLoss), ("geo", muvo_b200.GeoScalLoss)):
            p = cu(pred).requires_grad_(True)
            loss = cls()(p, cu(tgt))
            assert loss.dtype == torch.float32 and loss.dim() == 0
            loss.backward()
            ref, gref = float(g[f"c{C}_{name}_loss"]), g[f"c{C}_{name}_grad"]
            assert abs(loss.item() - ref) <= RTOL * abs(ref), (C, name, loss.item(), ref)
            err = np.abs(p.grad.cpu().numpy() - gref).max()
            assert err <= RTOL * np.abs(gref).max(), (C, name, err)
        # the float64 oracle is the tighter check on the forward value
        sem, geo = scal_losses(cu(pred), cu(tgt))
        assert abs(sem.item() - O.sem_scal_loss(pred, tgt)) <= 2e-6 * abs(sem.item())
        assert abs(geo.item() - O.geo_scal_loss(pred, tgt)) <= 2e-6 * abs(geo.item())


def test_absent_class_is_skipped(golden, lib):
    g = golden("scal.npz")
    p = cu(g["absent_pred"]).requires_grad_(True)
    loss = muvo_b200.SemScalLoss()(p, cu(g["absent_target"]))
    loss.backward()
    ref = float(g["absent_sem_loss"])
    assert abs(loss.item() - ref) <= RTOL * abs(ref)
    gref = g["absent_sem_grad"]
    assert np.abs(p.grad.cpu().numpy() - gref).max() <= RTOL * np.abs(gref).max()


@pytest.mark.parametrize("C,shape", [(2, (1, 2, 16, 12, 8)), (9, (2, 1, 8, 8, 8)), (3, (1, 1, 7, 5, 3)), (2, (1, 3, 5, 3, 3)),
                                     (9, (1, 2, 3, 3, 5)), (23, (1, 1, 6, 6, 4))])
def test_sums_against_oracle(lib, C, shape):
    """Shapes with S % 4 != 0 take the scalar-load kernels; C = 3 / 23 the any-C kernels."""
    gen = torch.Generator().manual_seed(C * 100 + shape[-1])
    b, s = shape[:2]
    pred = torch.randn((b, s, C) + shape[2:], generator=gen) * 3
    tgt = torch.randint(0, C, (b, s) + shape[2:], generator=gen).to(torch.uint8)
    tgt[torch.rand(tgt.shape, generator=gen) < 0.05] = 255
    got = scal_sums(pred.cuda(), tgt.cuda()).cpu().numpy()
    want = O.scal_sums(pred.numpy(), tgt.numpy())
    assert np.array_equal(got[2 * C:], want[2 * C:])                        # counts are exact
    assert np.allclose(got[:2 * C], want[:2 * C], rtol=1e-6, atol=1e-9)     # fp32 softmax, float64 sums
    # no ignore index at all (-1): 255 is then an ordinary (out of range) label
    got = scal_sums(pred.cuda(), tgt.cuda(), ignore_index=-1).cpu().numpy()
    want = O.scal_sums(pred.numpy(), tgt.numpy(), ignore_index=-1)
    assert np.array_equal(got[2 * C:], want[2 * C:]) and np.allclose(got[:2 * C], want[:2 * C], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_logits(lib, dtype):
    gen = torch.Generator().manual_seed(5)
    pred = (torch.randn((1, 2, 2, 16, 16, 8), generator=gen) * 2).to(dtype)
    tgt = torch.randint(0, 2, (1, 2, 16, 16, 8), generator=gen).to(torch.uint8)
    p = pred.cuda().requires_grad_(True)
    sem, geo = scal_losses(p, tgt.cuda())
    (sem + geo).backward()
    assert p.grad.dtype == dtype
    p32 = pred.float().cuda().requires_grad_(True)                          # same values in fp32: the autocast behaviour
    sem32, geo32 = scal_losses(p32, tgt.cuda())
    (sem32 + geo32).backward()
    assert abs(sem.item() - sem32.item()) <= 1e-6 and abs(geo.item() - geo32.item()) <= 1e-6
    assert torch.allclose(p.grad.float(), p32.grad, rtol=1e-2, atol=2e-7)   # rounding (fp16 subnormals) of the stored gradient only


def test_gradient_matches_finite_differences(lib):
    """Independent of the golden file: central differences of the float64 oracle on a few logits."""
    gen = torch.Generator().manual_seed(11)
    pred = torch.randn((1, 1, 3, 4, 4, 4), generator=gen)
    tgt = torch.randint(0, 3, (1, 1, 4, 4, 4), generator=gen).to(torch.uint8)
    tgt[0, 0, 0, 0, 0] = 255
    p = pred.cuda().requires_grad_(True)
    sem, geo = scal_losses(p, tgt.cuda())
    (2.0 * sem + 0.5 * geo).backward()
    g = p.grad.cpu().numpy()
    assert np.all(g[0, 0, :, 0, 0, 0] == 0)                                  # ignored voxel
    base = pred.numpy().astype(np.float64)
    f = lambda x: 2.0 * O.sem_scal_loss(x, tgt.numpy()) + 0.5 * O.geo_scal_loss(x, tgt.numpy())
    for idx in [(0, 0, 0, 1, 2, 3), (0, 0, 1, 3, 0, 1), (0, 0, 2, 2, 2, 2)]:
        hi, lo = base.copy(), base.copy()
        hi[idx] += 1e-4; lo[idx] -= 1e-4
        fd = (f(hi) - f(lo)) / 2e-4
        assert abs(fd - g[idx]) <= 1e-4 * max(1.0, abs(fd)), (idx, fd, g[idx])


def test_full_size_properties(lib):
    """Full 192x192x64 grid, 4 frames: counts are exact, probabilities sum to the number of valid voxels."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn((1, 4, 2, 192, 192, 64), generator=gen, device="cuda")
    tgt = (torch.rand((1, 4, 192, 192, 64), generator=gen, device="cuda") < 0.05).to(torch.uint8)
    tgt[0, :, :2] = 255
    s = scal_sums(logits, tgt).cpu().numpy()
    n_valid = int((tgt != 255).sum().item())
    assert s[6] == n_valid and s[4] + s[5] == n_valid and s[5] == int((tgt == 1).sum().item())
    assert abs(s[0] + s[1] - n_valid) <= 1e-6 * n_valid                      # sum over classes of softmax = 1
    ref = torch.softmax(logits.view(4, 2, -1), 1)[:, 1][(tgt.view(4, -1) == 1)].double().sum().item()
    assert abs(s[3] - ref) <= 1e-6 * ref

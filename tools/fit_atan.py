"""Fit the odd polynomial used by the f32 fast path of the range-pixel computation:
atan(t)/pi ~= t * Q(t^2) on t in [0, 1]; prints coefficients and the worst error of an f32 Horner evaluation
(no FMA, i.e. pessimistic) against float64 arctan, for several degrees."""
import numpy as np

def cheb_nodes(n, a=0.0, b=1.0):
    k = np.arange(n)
    x = np.cos(np.pi * (2 * k + 1) / (2 * n))
    return 0.5 * (a + b) + 0.5 * (b - a) * x

def fit(deg_q, iters=30):
    # weighted least squares, iteratively re-weighted (Lawson) towards minimax of the ABSOLUTE error
    t = np.linspace(1e-6, 1.0, 20001)
    y = np.arctan(t) / np.pi
    A = np.stack([t ** (2 * k + 1) for k in range(deg_q + 1)], 1)
    w = np.ones_like(t)
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
        e = np.abs(A @ c - y)
        w = w * (e / e.max() + 1e-3)
        w /= w.max()
    return c

def eval_f32(c, t):
    t = t.astype(np.float32)
    t2 = (t * t).astype(np.float32)
    acc = np.float32(c[-1]) * np.ones_like(t2)
    for k in range(len(c) - 2, -1, -1):
        acc = (acc * t2).astype(np.float32)
        acc = (acc + np.float32(c[k])).astype(np.float32)
    return (acc * t).astype(np.float32)

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.random(4_000_000), np.linspace(0, 1, 1_000_001)])
    for dq in (4, 5, 6, 7):
        c = fit(dq)
        err = np.abs(eval_f32(c, t).astype(np.float64) - np.arctan(t.astype(np.float32).astype(np.float64)) / np.pi)
        print(dq, "max abs err (units of pi):", err.max(), " -> bins at W=1024:", err.max() * 512)
        print("   ", ", ".join(f"{np.float32(v)!r}" for v in c))

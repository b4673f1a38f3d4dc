"""Least-squares fit of the FMA-only GELU of gelu.cuh (gelu_poly2): Phi(x) - 1/2 is odd and saturates at +-1/2, so with
z = clamp(x / R, -1, 1), t = 2 z^2 - 1:  Phi(x) ~= 1/2 + z Q(t), Q of degree `deg`, constrained to z Q = 1/2 at z = 1 (the clamp
continues it with the exact limits).  Minimises the max error of GELU(x) = x Phi(x) RELATIVE to max(|x|, floor) by iterated
re-weighting; errors are then measured in fp32 Horner arithmetic as the kernel evaluates it."""
import sys
import numpy as np
from math import erf, sqrt

Phi = np.vectorize(lambda v: 0.5 * (1.0 + erf(v / sqrt(2.0))))


def fit(R, deg, iters=60):
    z = np.linspace(1e-4, 1.0, 8001)
    x = R * z
    t = 2 * z * z - 1
    target = Phi(x) - 0.5
    # basis: z * t^k, k = 0..deg; constraint sum_k c_k = 1/2 (t = 1 at z = 1)  ->  eliminate c_0 = 1/2 - sum_{k>=1} c_k
    A = np.stack([z * (t ** k - 1.0) for k in range(1, deg + 1)], axis=1)
    b = target - 0.5 * z
    w = np.ones_like(z)
    best = None
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(A * w[:, None], b * w, rcond=None)
        err = np.abs(A @ c - b)
        m = err.max()
        if best is None or m < best[0]:
            best = (m, c.copy())
        w = w * (1.0 + 4.0 * err / m) ; w /= w.mean()
    c = best[1]
    coef = np.concatenate([[0.5 - c.sum()], c])        # c_0 .. c_deg (powers of t)
    return coef


def check(R, coef):
    x = np.linspace(-8, 8, 400001).astype(np.float32)
    z = np.clip(x * np.float32(1.0 / R), -1, 1).astype(np.float32)
    t = (z * z * np.float32(2) - np.float32(1)).astype(np.float32)
    q = np.full_like(t, np.float32(coef[-1]))
    for ck in coef[-2::-1]:
        q = (q * t + np.float32(ck)).astype(np.float32)
    phi = (z * q + np.float32(0.5)).astype(np.float32)
    ref = Phi(x.astype(np.float64))
    e_phi = np.abs(phi - ref).max()
    g = x * phi
    gref = x.astype(np.float64) * ref
    e_abs = np.abs(g - gref).max()
    e_rel = (np.abs(g - gref) / np.maximum(np.abs(gref), 1e-3)).max()
    return e_phi, e_abs, e_rel


if __name__ == "__main__":
    for R in (3.5, 3.75, 4.0, 4.25, 4.5):
        for deg in (6, 7, 8, 9):
            coef = fit(R, deg)
            e_phi, e_abs, e_rel = check(R, coef)
            print(f"R={R} deg={deg}: max|dPhi|={e_phi:.2e}  max|dGELU|={e_abs:.2e}  max rel (floor 1e-3)={e_rel:.2e}")
            if len(sys.argv) > 1 and float(sys.argv[1]) == R and int(sys.argv[2]) == deg:
                print("  coef (t^0..):", ", ".join(f"{c:.9e}f" for c in coef))

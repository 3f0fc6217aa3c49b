"""Shared helpers for the parity tests (oracle = checker only)."""
import numpy as np


def plane_relerr(a, b, floor=1e-4):
    """max over planes of  |a-b|_inf / max(|b_plane|_inf, floor*|b|_inf)  -- SURVEY.md §8(d) parity gate."""
    a, b = np.asarray(a), np.asarray(b)
    fmax = np.max(np.abs(b))
    if fmax == 0.0:
        return float(np.max(np.abs(a)))
    worst = 0.0
    for pl in range(b.shape[0]):
        den = max(np.max(np.abs(b[pl])), floor * fmax)
        worst = max(worst, float(np.max(np.abs(a[pl] - b[pl])) / den))
    return worst


def smooth_field(rng, P, nr, dim, dr, amp=1.0):
    """analytic smooth modes a*r^k*exp(-r^2/s) with random amplitudes, all planes non-zero; guards included"""
    r = (np.arange(nr + 2) - 1) * dr
    f = np.zeros((P, nr + 2, dim))
    for pl in range(P):
        m = (pl + 1) // 2
        for c in range(dim):
            a, s, k = amp * rng.uniform(0.3, 1.0) * rng.choice([-1, 1]), rng.uniform(0.5, 3.0), rng.integers(0, 3)
            f[pl, :, c] = a * np.abs(r) ** (m + k) * np.exp(-r * r / s) + 0.05 * amp * rng.standard_normal(nr + 2)
    f[:, 0, :] = 0.0
    return f


def perturbed_lattice(O, rng, nr, dr, ppc1, ppc2, nth, pamp=0.5, jitter=0.3):
    x, p, g, psi, q = O.inject_uniform(nr, dr, ppc1, ppc2, nth)
    n = len(q)
    x = x + jitter * dr * rng.standard_normal((n, 2))
    rr = np.hypot(x[:, 0], x[:, 1])
    bad = rr >= (nr - 1e-3) * dr
    x[bad] *= ((nr - 1.0) * dr / rr[bad])[:, None]
    p = pamp * rng.standard_normal((n, 3))
    g = np.sqrt(1.0 + (p ** 2).sum(1))
    psi = rng.standard_normal(n) * 0.1
    return np.ascontiguousarray(x), np.ascontiguousarray(p), g, psi, q.copy()

"""Known answers for the sub-cycling / clamp variant of the oracle (proj_subcyc/ of the reference)."""
import ctypes as C

import numpy as np

from oracle import oracle as O

L = O.lib()


def test_subcycle_count_rule():
    """simulation_subcyc_class.f03:431-451"""
    def step(fac, fmax, dt, dtmin):
        d, n = C.c_double(), C.c_int()
        L.orc_subcyc_step(fac, fmax, dt, dtmin, C.byref(d), C.byref(n))
        return d.value, n.value
    assert step(1.2, 1.5, 0.02, 1e-3) == (0.02, 1)
    assert step(3.1, 1.5, 0.02, 1e-3) == (0.02 / 3, 3)                 # ceil(3.1 / 1.5) = 3
    d, n = step(100.0, 1.5, 0.02, 3e-3)                                  # 67 sub-steps would be shorter than dt_min: floor(dt / dt_min) = 6
    assert n == 6 and abs(d - 0.02 / 6) < 1e-18


def test_clamp_sets_the_expansion_factor_exactly():
    """part2d_subcyc_class.f03:48-66: a clamped particle ends at gamma / (gamma - p_z) = clamp with its direction unchanged"""
    rng = np.random.default_rng(4)
    n = 2000
    p = rng.standard_normal((n, 3)) * np.array([1.0, 1.0, 6.0])
    g = np.sqrt(1.0 + (p ** 2).sum(1))
    fac0 = g / (g - p[:, 2])
    assert abs(L.orc_exp_fac_max(p, g, n) - fac0.max()) == 0.0 and L.orc_exp_fac_max(p, g, 0) == 1.0
    clamp = 4.0
    p1, g1 = p.copy(), g.copy()
    L.orc_clamp_exp_fac(p1, g1, n, clamp)
    hit = fac0 > clamp
    assert 50 < hit.sum() < n - 50
    assert np.array_equal(p1[~hit], p[~hit]) and np.array_equal(g1[~hit], g[~hit])
    fac1 = g1 / (g1 - p1[:, 2])
    assert np.max(np.abs(fac1[hit] - clamp)) < 1e-10
    assert np.max(np.abs(g1 - np.sqrt(1 + (p1 ** 2).sum(1)))) < 1e-13
    cosang = (p1[hit] * p[hit]).sum(1) / np.linalg.norm(p1[hit], axis=1) / np.linalg.norm(p[hit], axis=1)
    assert np.max(np.abs(cosang - 1.0)) < 1e-12 and np.all(np.linalg.norm(p1[hit], axis=1) < np.linalg.norm(p[hit], axis=1))


def _blowout(**kw):
    from qpad_b200 import decks
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=3, ppc1=2, ppc2=2, num_theta=8)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C1"]["beam"]))
    sim = O.Sim(**cfg, **kw)
    sim.set_beam(*bm)
    sim.run_slices(24)
    return sim


def test_loop_without_subcycles_is_the_standard_loop():
    """thresholds that never trigger: one sub-step per slice with the full dxi and no clamping -- the sub-cycling loop
    (simulation_subcyc_class.f03:216-376) must then reproduce the standard loop (simulation_class.f03:342-469)"""
    ref = _blowout()
    sub = _blowout(subcyc_on=1, subcyc_exp_fac_max=1e9, subcyc_exp_fac_clamped=1e9, subcyc_dt_min=1e-6)
    assert sub.total_subcycles() == 24 and sub.total_iters() == ref.total_iters()
    for name in ("psi", "e", "b", "cu"):
        a, b = ref.field(name, 2)[:, :24], sub.field(name, 2)[:, :24]
        assert np.max(np.abs(a)) > 1e-3 and np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(a)), name


def test_subcycling_kicks_in_behind_the_beam():
    ref = _blowout()
    x, p, g, psi, q = ref.plasma()
    fac = (g / (g - p[:, 2])).max()
    assert fac > 1.3                                                   # sheath electrons with a forward momentum
    sub = _blowout(subcyc_on=1, subcyc_exp_fac_max=1.1, subcyc_exp_fac_clamped=50.0, subcyc_dt_min=1e-3)
    assert 24 < sub.total_subcycles() < 24 * 8
    a, b = ref.field("psi", 2)[:, :24], sub.field("psi", 2)[:, :24]
    assert np.all(np.isfinite(b)) and np.max(np.abs(a - b)) < 0.25 * np.max(np.abs(a))   # the same wake, resolved more finely

"""Parity cases of the per-routine kernels shared by the GPU tests (api = qpad_b200.capi, the C-ABI of libqpadb200.so) and the
host-emulation tests (api = tests.emu.emu, the same device sources compiled for the CPU): one body, two back ends."""
import ctypes as C

import numpy as np


def _mx(a):
    return float(np.max(np.abs(a))) if a.size else 0.0


def neutral_update(api, O, elem, mm, M):
    """ionize + add_particles on a given field, several updates in a row: levels to 1e-12, the released electrons bit-exact in
    number and order, positions / charges to 1e-14 (neutral_class.f03:600-837 vs oracle/qpad_oracle_neutral.c)"""
    L = O.lib()
    nr, nth, ppc = 48, 8, (2, 2)
    dr, dxi = 0.1, 0.02
    ctx = api.Ctx(nr, M, dr, dxi)
    rng = np.random.default_rng(elem)
    r = (np.arange(nr + 2) - 1) * dr
    wp = L.orc_plasma_frequency(1.0e17)
    amp = {1: 40.0, 2: 120.0, 3: 30.0}[elem] / (wp * 1.708e-12)
    e = np.zeros((2 * M + 1, nr + 2, 3))
    for pl in range(2 * M + 1):
        e[pl] = amp * (0.3 + rng.random(3))[None, :] * (np.exp(-((r - 2.0) / 1.2) ** 2) * (1 if pl == 0 else 0.3))[:, None]
    fe = api.Field(ctx, 3); fe.upload(e)
    ne = api.Neutral(ctx, elem, mm, ppc, nth, n0=1.0e17, dt_xi=dxi)
    adk = np.zeros(3 * mm); assert L.orc_adk_params(elem, mm, adk) == mm == ne.multi_max
    lev = np.zeros((mm + 2, nth, nr)); L.orc_neutral_reset(lev, nr, nth, mm)
    cap = nr * nth * 4 + 64
    x, p = np.zeros((cap, 2)), np.zeros((cap, 3))
    g, psi, q = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    xa, qa = np.zeros((cap, 2)), np.zeros(cap)
    npp = C.c_long(0)
    for step in range(6):
        old = lev[mm + 1].copy()
        L.orc_neutral_ionize(lev, adk, e, wp, dxi, ppc[0], ppc[1], nr, nth, M, mm)
        nadd = L.orc_neutral_add_particles(lev, old, nr, nth, mm, ppc[0], ppc[1], dr, -1.0, 1.0, 1e-10, x, p, g, psi, q, C.byref(npp), xa, qa)
        ne.update(fe)
        got = ne.levels()
        assert np.max(np.abs(got - lev)) < 1e-12, step
        gx, gp, gg, gpsi, gq = ne.part.download()
        assert len(gq) == npp.value
        assert _mx(gx - x[:npp.value]) < 1e-14 * 5 and _mx(gq - q[:npp.value]) < 1e-14
        assert np.all(gp == 0.0) and np.all(gg == 1.0) and np.all(gpsi == 0.0)
        ix, _, _, _, iq = ne.part_add.download()
        assert len(iq) == nadd and _mx(ix - xa[:nadd]) < 1e-14 * 5 and _mx(iq - qa[:nadd]) < 1e-14
    assert npp.value > 100
    ne.renew()
    assert ne.part.npp() == 0 and ne.part_add.npp() == 0
    lev0 = np.zeros_like(lev); L.orc_neutral_reset(lev0, nr, nth, mm)
    assert np.array_equal(ne.levels(), lev0)
    ne.close()

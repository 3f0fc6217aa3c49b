"""GPU parity of the field-ionisation neutral species (csrc/neutral.cu, qpad_b200/ionization.py) against the oracle
(oracle/qpad_oracle_neutral.c).  The device code was written at the end of round 1 after the round's GPU minutes were
spent: until it has had a first run on a GPU these tests are opt-in (QPG_TEST_NEUTRAL=1) so that an unvalidated path cannot
mask the state of the validated ones."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.environ.get("QPG_TEST_NEUTRAL"), reason="neutral path awaits its first GPU run (set QPG_TEST_NEUTRAL=1)")]


@pytest.fixture(scope="module")
def mods():
    from qpad_b200 import capi
    from oracle import oracle as O
    capi.load()
    return capi, O


@pytest.mark.parametrize("elem,mm,M", [(1, 1, 0), (3, 3, 1), (2, 2, 2)])
def test_neutral_update_matches_oracle(mods, elem, mm, M):
    """ionize + add_particles on a given field, several updates in a row: levels to 1e-12, the released electrons bit-exact in
    number and order, positions / charges to 1e-14"""
    capi, O = mods
    import ctypes as C
    L = O.lib()
    nr, nth, ppc = 48, 8, (2, 2)
    dr, dxi = 0.1, 0.02
    ctx = capi.Ctx(nr, M, dr, dxi)
    rng = np.random.default_rng(elem)
    r = (np.arange(nr + 2) - 1) * dr
    wp = L.orc_plasma_frequency(1.0e17)
    amp = {1: 40.0, 2: 120.0, 3: 30.0}[elem] / (wp * 1.708e-12)
    e = np.zeros((2 * M + 1, nr + 2, 3))
    for pl in range(2 * M + 1):
        e[pl] = amp * (0.3 + rng.random(3))[None, :] * (np.exp(-((r - 2.0) / 1.2) ** 2) * (1 if pl == 0 else 0.3))[:, None]
    fe = capi.Field(ctx, 3); fe.upload(e)
    ne = capi.Neutral(ctx, elem, mm, ppc, nth, n0=1.0e17, dt_xi=dxi)
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
        assert np.max(np.abs(gx - x[:npp.value])) < 1e-14 * 5 and np.max(np.abs(gq - q[:npp.value])) < 1e-14
        ix, _, _, _, iq = ne.part_add.download()
        assert len(iq) == nadd and np.max(np.abs(ix - xa[:nadd])) < 1e-14 * 5 and np.max(np.abs(iq - qa[:nadd])) < 1e-14
    assert npp.value > 100


def test_ionization_loop_matches_oracle(mods):
    """config 5 in small through the per-routine C-ABI (qpad_b200.ionization.IonizationStage) against the oracle's loop"""
    capi, O = mods
    from qpad_b200 import decks
    from qpad_b200.ionization import IonizationStage
    cfg = dict(nr=96, nz=64, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3, n0=1.0e17)
    beam = dict(decks.CONFIGS["C5"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8,
                **{k: v for k, v in cfg.items()})
    orc.set_beam(*bm)
    nsl = 48
    orc.run_slices(nsl)
    st = IonizationStage(cfg, dict(element=3, ion_max=1, ppc=(2, 2), num_theta=8), bm)
    st.step3d(nslices=nsl, beam_push=False)
    assert st.iters == orc.total_iters()
    lev = orc.levels(1)
    assert np.max(np.abs(st.neut.levels() - lev)) < 1e-10
    assert st.neut.part.npp() == len(orc.neutral()[4]) > 100
    for name, f in (("psi", st.psi), ("e", st.e), ("b", st.b)):
        got, want = f.download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
        assert np.max(np.abs(want)) > 1e-2 and np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), name
    st.close()

"""Known answers for the field-ionisation (ADK) neutral species of the oracle (oracle/qpad_oracle_neutral.c, SURVEY.md §8(f)
rank 2; reference: species/neutral_class.f03).  Nothing in the reference pins this path, so it is pinned by the closed form
of the rate equations in a constant field, by the tunnelling-rate formula the parameter table comes from, by the
conservation laws of the level populations and of charge, and by a small beam-driven run in a neutral lithium gas."""
import math

import numpy as np
import pytest

from oracle import oracle as O

L = O.lib()
XI_EV = {1: [13.598434], 2: [24.587387, 54.417760], 3: [5.391719, 75.6400, 122.45429]}     # NIST ASD ionisation energies


def _adk(elem, n):
    out = np.zeros(3 * n)
    assert L.orc_adk_params(elem, n, out) == n
    return out


@pytest.mark.parametrize("elem", [1, 2, 3])
def test_adk_table_is_the_tunnelling_rate_formula(elem):
    """w = r1 E^-r3 exp(-r2 / E): r2 = 6.83 xi^1.5, r3 = 2 n* - 1, r1 = 1.52e15 4^n* xi / (n* Gamma(2 n*)) (20.5 xi^1.5)^(2 n* - 1),
    n* = 3.69 Z / sqrt(xi) -- the table entries agree with the (rounded-constant) formula to 1 %"""
    xi = XI_EV[elem]
    tab = _adk(elem, len(xi)).reshape(-1, 3)
    for Z, (e, row) in enumerate(zip(xi, tab), 1):
        ns = 3.69 * Z / math.sqrt(e)
        r1 = 1.52e15 * 4 ** ns * e / (ns * math.gamma(2 * ns)) * (20.5 * e ** 1.5) ** (2 * ns - 1)
        assert abs(r1 / row[0] - 1) < 1.5e-2 and abs(6.83 * e ** 1.5 / row[1] - 1) < 3e-4 and abs((2 * ns - 1) / row[2] - 1) < 2e-3
    assert L.orc_adk_params(92, 3, np.zeros(9)) == 0                 # unsupported element
    assert abs(L.orc_plasma_frequency(1.0e17) - math.sqrt(1e17) * 5.641460231180626e4) < 1.0


def _uniform_field(nr, M, ez):
    e = O.zeros_f1(3, nr, M)
    e[0, :, 2] = ez
    return e


def test_single_level_in_a_constant_field():
    """hydrogen in a constant field: the neutral fraction after n updates is (1 - w dt (1 + w dt / 2))^n -- the reference's
    update rule (neutral_class.f03:661) -- and it tends to exp(-w t) as dt -> 0"""
    nr, nth, M, mm, ppc = 16, 4, 1, 1, (4, 4)
    adk = _adk(1, 1)
    wp = L.orc_plasma_frequency(1.0e17)
    e_gvm = 35.0                                                        # GV/m
    ez = e_gvm / (wp * 1.708e-12)
    w = adk[0] * e_gvm ** (-adk[2]) * math.exp(-adk[1] / e_gvm) / wp     # normalised rate
    assert 0.01 < w < 10.0
    for nsteps in (10, 40, 160):
        dt = 0.5 / (w * nsteps) * 4                                      # total w t = 2
        lev = np.zeros((mm + 2, nth, nr))
        L.orc_neutral_reset(lev, nr, nth, mm)
        for _ in range(nsteps):
            L.orc_neutral_ionize(lev, adk, _uniform_field(nr, M, ez), wp, dt, ppc[0], ppc[1], nr, nth, M, mm)
        want = (1.0 - w * dt * (1.0 + 0.5 * w * dt)) ** nsteps
        assert np.max(np.abs(lev[mm] - want)) < 1e-13
        assert np.max(np.abs(lev[0] + lev[mm] - 1.0)) < 1e-14
        # discrete total level: a multiple of ion_max / ppc, nearest to the continuous one
        q = mm / 16.0
        assert np.max(np.abs(lev[mm + 1] / q - np.round(lev[mm + 1] / q))) < 1e-12 and np.max(np.abs(lev[mm + 1] - lev[0])) <= 0.5 * q + 1e-12
        err = abs(want - math.exp(-2.0))
        if nsteps == 10:
            e10 = err
    assert err < e10 / 8                                                 # first-order convergence to exp(-w t)
    # below the field threshold nothing happens
    lev = np.zeros((mm + 2, nth, nr)); L.orc_neutral_reset(lev, nr, nth, mm)
    L.orc_neutral_ionize(lev, adk, _uniform_field(nr, M, 1e-7 / (wp * 1.708e-12) * 0.5), wp, 1.0, 4, 4, nr, nth, M, mm)
    assert np.all(lev[mm] == 1.0) and not lev[0].any()


def test_lithium_levels_conserve_population_and_cascade():
    nr, nth, M, mm = 8, 2, 0, 3
    adk = _adk(3, 3)
    wp = L.orc_plasma_frequency(1.0e17)
    lev = np.zeros((mm + 2, nth, nr)); L.orc_neutral_reset(lev, nr, nth, mm)
    e = _uniform_field(nr, M, 700.0 / (wp * 1.708e-12))                  # 700 GV/m: Li+ and Li2+ go at once, Li3+ at w dt ~ 0.05
    seen = []
    for _ in range(200):
        L.orc_neutral_ionize(lev, adk, e, wp, 0.02, 2, 2, nr, nth, M, mm)
        assert np.max(np.abs(lev[:mm + 1].sum(0) - 1.0)) < 1e-13 and lev[:mm + 1].min() >= -1e-15
        seen.append(lev[:mm + 1, 0, 0].copy())
    seen = np.array(seen)
    # the 5.4 eV electron goes within the first update (overshoot capped: the neutral residue is exactly 0), half of the Li+
    # made in that step moves on at once (the reference's time-centred overshoot rule, :676-690) ...
    assert seen[0, mm] == 0.0 and abs(seen[0, 0] - 0.5) < 1e-12
    assert seen[-1, 2] > 0.5 and np.all(np.diff(seen[:, 2]) >= -1e-15)  # ... and the deepest level fills up monotonically
    assert np.all(lev[mm + 1] <= mm + 1e-12)


def test_created_electrons_are_neutralised_by_the_ion_deposit():
    """add_particles + ion_deposit (neutral_class.f03:755, :904): the electrons released in a cell and the ion charge left
    behind deposit to exactly zero net charge; charges and positions follow the deterministic lattice rule"""
    nr, nth, M, mm, ppc = 32, 8, 1, 1, (2, 4)
    dr = 0.1
    lev = np.zeros((mm + 2, nth, nr)); L.orc_neutral_reset(lev, nr, nth, mm)
    old = lev[mm + 1].copy()
    rng = np.random.default_rng(0)
    lev[mm + 1] = rng.integers(0, 9, size=(nth, nr)) / 8.0               # discrete levels: k / ppc_tot
    cap = nr * nth * 8 + 8
    x, p = np.zeros((cap, 2)), np.ones((cap, 3))
    g, psi, q = np.zeros(cap), np.ones(cap), np.zeros(cap)
    xa, qa = np.zeros((cap, 2)), np.zeros(cap)
    import ctypes as C
    npp = C.c_long(0)
    nadd = L.orc_neutral_add_particles(lev, old, nr, nth, mm, ppc[0], ppc[1], dr, -1.0, 1.0, 1e-10, x, p, g, psi, q, C.byref(npp), xa, qa)
    assert nadd == npp.value == int(round(lev[mm + 1].sum() * 8))
    n = nadd
    assert np.array_equal(xa[:n], x[:n]) and np.array_equal(qa[:n], -q[:n]) and not p[:n].any() and np.all(g[:n] == 1) and not psi[:n].any()
    r = np.hypot(x[:n, 0], x[:n, 1]) / dr
    assert np.allclose(q[:n], -r * mm / (8.0 * nth), rtol=1e-14)
    # a cell that releases k electrons places them at (i - 1/2) / k inside the cell
    k0 = int(round(lev[mm + 1, 0, 5] * 8))
    cell = np.where((np.abs(np.arctan2(x[:n, 1], x[:n, 0])) < 1e-12) & (r >= 5) & (r < 6))[0]
    assert len(cell) == k0 and np.allclose(np.sort(r[cell]) - 5, (np.arange(1, k0 + 1) - 0.5) / max(k0, 1))
    qe, tot = O.zeros_f1(1, nr, M), O.zeros_f1(1, nr, M)
    L.orc_qdeposit(np.ascontiguousarray(x[:n]), np.ascontiguousarray(q[:n]), n, dr, nr, M, qe)
    rho_ion = O.zeros_f1(1, nr, M)
    L.orc_neutral_ion_deposit(xa, qa, n, dr, nr, M, rho_ion, tot)
    assert np.max(np.abs(qe)) > 0.1 and np.max(np.abs(qe + rho_ion)) < 1e-15 and np.array_equal(tot, rho_ion)


def test_beam_ionises_lithium_gas_and_drives_a_wake():
    """config 5 in small (input_file/ionization: nspecies 0, one Li neutral, one beam): the beam's field strips the 5.4 eV
    electron near the axis, the released electrons form a wake; far from the beam the gas stays neutral"""
    from qpad_b200 import decks
    cfg = dict(nr=96, nz=64, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    beam.update(center=(0.0, 0.0, 2.5), range3=(0.0, 5.0))
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    sim = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, n0=1.0e17, ppc1=2, ppc2=2,
                num_theta=8, **cfg)
    sim.set_beam(*bm)
    assert sim.plasma()[4].size == 0                                      # nspecies = 0
    sim.run_slices(48)
    lev = sim.levels(1)
    x, p, g, psi, q = sim.neutral()
    assert len(q) > 400 and np.all(q < 0)
    assert np.all(lev[2, :, 1:8] == 1.0)                                  # fully stripped around the beam (the cell on the axis sees E_r ~ 0)
    assert np.all(lev[1, :, 80:] == 1.0) and not lev[2, :, 80:].any()    # untouched far outside
    # every released macro-electron is accounted for by the discrete ion level (none has left the box yet)
    assert len(q) == int(round(lev[2].sum() * 4))
    # the released electrons are blown out of the beam's path: an ion column (psi > 0 on the axis) and an oscillating E_z
    fpsi, fez = sim.field("psi", 2)[0, :48, 1, 0], sim.field("e", 2)[0, :48, 1, 2]
    assert fpsi.max() > 0.1 and fpsi.min() > -1e-12 and fez.max() > 0.05 and fez.min() < -0.02
    r = np.hypot(x[:, 0], x[:, 1])
    assert 1.0 < r.max() < 6.0


def test_neutral_pipeline_stages_match_single_stage():
    """the neutral species across xi stages (psend / precv, neutral_class.f03:1025-1101: created electrons, the ions' position
    buffer, rho_ion and the ionisation levels travel forward) reproduces the single-stage run -- the ionization deck's nodes = [1, 2]"""
    from qpad_b200 import decks
    cfg = dict(nr=64, nz=48, max_mode=1, rmax=5.0, zmin=0.0, zmax=6.0, dt=10.0, iter_max=3, ppc1=2, ppc2=2, num_theta=8, sp_density=0.0,
               neut_on=1, neut_elem=3, neut_ion_max=2, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, n0=1.0e17)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    beam.update(center=(0.0, 0.0, 2.0), range3=(0.0, 4.0))
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    ref, pip = O.Sim(nstages=1, **cfg), O.Sim(nstages=3, **cfg)
    for s in (ref, pip):
        s.set_beam(*bm)
    # the state just before the renewal: run the slices of a step through the public stepping, compare the stored volumes
    u1, u2 = ref.step3d(1), pip.step3d(1)
    assert u1 == u2 > 1000                                     # particle-slice updates of the released electrons
    for name in ("psi", "e", "b"):
        full = ref.field(name, 2)[:, :-1]
        parts = np.concatenate([pip.field(name, 2, stage=k)[:, :-1] for k in range(3)], axis=1)
        assert np.max(np.abs(full)) > 1e-2 and np.max(np.abs(parts - full)) <= 1e-11 * np.max(np.abs(full)), name
    assert ref.total_iters() == pip.total_iters()

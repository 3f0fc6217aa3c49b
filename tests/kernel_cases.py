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


def subcyc_particles(api, O):
    """part2d_subcyc%get_exp_fac_max / clamp_exp_fac (proj_subcyc/part2d_subcyc_class.f03:28-66) and the sub-step rule
    (simulation_subcyc_class.f03:431-451): bit-exact against the oracle"""
    L = O.lib()
    ctx = api.Ctx(64, 1, 0.1, 0.02)
    rng = np.random.default_rng(7)
    n = 5003
    p = rng.normal(size=(n, 3)) * np.array([0.5, 0.5, 3.0])
    p[::17, 2] = np.abs(p[::17, 2]) * 40.0 + 5.0                    # a few strongly forward-moving particles: large gamma / (gamma - p_z)
    g = np.sqrt(1.0 + np.sum(p * p, axis=1))
    x = rng.random((n, 2)); psi = rng.random(n); q = -rng.random(n)
    pt = api.Part2d(ctx, -1.0, n + 100)
    assert pt.exp_fac_max() == 1.0                                   # no particles (:43)
    pt.upload(x, p, g, psi, q)
    want = L.orc_exp_fac_max(np.ascontiguousarray(p), g, n)
    got = pt.exp_fac_max()
    assert got == want and want > 20.0
    for clamp in (50.0, 4.0, 1.5):
        po, go = p.copy(), g.copy()
        L.orc_clamp_exp_fac(po, go, n, clamp)
        pt.upload(x, p, g, psi, q)
        pt.clamp_exp_fac(clamp)
        gx, gp, gg, gpsi, gq = pt.download()
        assert np.array_equal(gp, po) and np.array_equal(gg, go), clamp
        assert np.array_equal(gx, x) and np.array_equal(gq, q)
        assert np.any(po != p)
        assert pt.exp_fac_max() == L.orc_exp_fac_max(po, go, n) <= clamp * (1 + 1e-12)
    for ef, efm, dt, dtmin in ((1.0, 2.0, 0.02, 0.001), (7.3, 2.0, 0.02, 0.001), (900.0, 2.0, 0.02, 0.001), (2.0, 2.0, 0.02, 0.001), (4.0000001, 2.0, 0.02, 0.0)):
        dts, ns = C.c_double(), C.c_int()
        L.orc_subcyc_step(ef, efm, dt, dtmin, C.byref(dts), C.byref(ns))
        assert api.subcyc_step(ef, efm, dt, dtmin) == (dts.value, ns.value)
    pt.close(); ctx.close()


def vpot(api, O, M, bnd, nr=96, exact=False):
    """field_vpot%solve_vpotz / solve_vpott (fields/field_vpot_class.f03:354, :392) on a random current"""
    L = O.lib()
    dr = 0.05
    ctx = api.Ctx(nr, M, dr, 0.02, field_boundary=bnd)
    rng = np.random.default_rng(100 + M)
    r = (np.arange(nr + 2) - 1) * dr
    cu = rng.normal(size=(2 * M + 1, nr + 2, 3)) * np.exp(-((r - 1.5) / 0.8) ** 2)[None, :, None]
    fcu, fv = api.Field(ctx, 3), api.Field(ctx, 3)
    fcu.upload(cu)
    marker = rng.normal(size=cu.shape)                               # what a solve must leave alone: the other components and the guards
    fv.upload(marker)
    want = marker.copy()
    L.orc_solve_vpotz(cu, want, nr, M, dr, bnd)
    ctx.solve_vpotz(fcu, fv)
    got = fv.download()
    assert np.max(np.abs(want[:, 1:nr + 1, 2])) > 1e-3
    assert np.max(np.abs(got - want)) <= (0.0 if exact else 1e-13 * np.max(np.abs(want)))
    L.orc_solve_vpott(cu, want, nr, M, dr, bnd)
    ctx.solve_vpott(fcu, fv)
    got = fv.download()
    assert np.max(np.abs(want[:, 1:nr + 1, :2])) > 1e-3
    assert np.max(np.abs(got - want)) <= (0.0 if exact else 1e-13 * np.max(np.abs(want)))
    assert np.array_equal(got[:, 0], marker[:, 0]) and np.array_equal(got[:, nr + 1], marker[:, nr + 1])
    ctx.close()


def stage(api, O):
    """diagnostics staging (csrc/diag.cu): datasets in the layouts of hdf5io_class.f03 pwfield_pipe :591 (f2(dim, 1:nr, 1:nzp) per
    plane), pwpart_2d_r :1027 and pwpart_3d_pipe :1220 (tnpp = int(npp / dspl), every dspl-th particle, x3 + z0)"""
    nr, nzp, M = 40, 9, 1
    ctx = api.Ctx(nr, M, 0.1, 0.02)
    rng = np.random.default_rng(3)
    st = api.Stage(ctx, 3 * 3 * nzp * nr + 8 * 2000)
    for dim in (1, 3):
        f = api.Field(ctx, dim, nzp, True)
        vol = rng.normal(size=(2 * M + 1, nzp + 1, nr + 2, dim))
        f.upload_f2(vol)
        st.field(f)
        with pytest_raises_state(api):
            st.field(f)                                              # one transfer in flight per stage
        got = st.wait()
        assert got.shape == (2 * M + 1, dim, nzp, nr)
        assert np.array_equal(got, np.transpose(vol[:, :nzp, 1:nr + 1, :], (0, 3, 1, 2)))
    n = 1003
    x, p = rng.random((n, 2)), rng.normal(size=(n, 3))
    g, psi, q = rng.random(n) + 1, rng.random(n), -rng.random(n)
    pt = api.Part2d(ctx, -1.0, 1500)
    pt.upload(x, p, g, psi, q)
    for dspl in (1, 7, 2000):
        st.part2d(pt, dspl)
        got = st.wait()
        t = n // dspl
        sel = np.arange(t) * dspl
        assert got.shape == (6, t)
        assert np.array_equal(got, np.stack([x[sel, 0], x[sel, 1], p[sel, 0], p[sel, 1], p[sel, 2], q[sel]]))
    x3, p3 = rng.random((n, 3)), rng.normal(size=(n, 3))
    bm = api.Part3d(ctx, -1.0, 10.0, 1500, nzp, 0, nzp)
    bm.upload(x3, p3, q)
    st.part3d(bm, 5, z0=-3.25)
    got = st.wait()
    sel = np.arange(n // 5) * 5
    assert np.array_equal(got, np.stack([x3[sel, 0], x3[sel, 1], x3[sel, 2] + (-3.25), p3[sel, 0], p3[sel, 1], p3[sel, 2], q[sel]]))
    st.close(); ctx.close()


class pytest_raises_state:
    """the call inside must fail with the library's QPG_ERR_STATE (-5)"""

    def __init__(self, api): pass
    def __enter__(self): return self

    def __exit__(self, et, ev, tb):
        assert et is not None and "-5" in str(ev), "expected QPG_ERR_STATE"
        return True


def subcyc_loop(api, O):
    """qpad_b200.subcyc.SubcycStage on the GPU against the oracle's sub-cycling loop (the call sequence itself is pinned on the CPU
    by tests/test_subcyc_host_loop.py)"""
    from qpad_b200 import decks, subcyc
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C1"]["beam"]))
    nsl = 24
    for efm, clamp, dtmin in ((1.1, 50.0, 1e-3), (1.05, 1.6, 1e-3)):
        orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, subcyc_on=1, subcyc_exp_fac_max=efm, subcyc_exp_fac_clamped=clamp, subcyc_dt_min=dtmin, **cfg)
        orc.set_beam(*bm)
        orc.run_slices(nsl)
        st = subcyc.SubcycStage(dict(cfg, exp_fac_max=efm, exp_fac_clamped=clamp, dt_min=dtmin), O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8), bm)
        st.step3d(nslices=nsl)
        assert st.subcycles == orc.total_subcycles() > nsl and st.iters == orc.total_iters()
        for name, f in (("psi", st.psi), ("e", st.e), ("b", st.b)):
            got, want = f.download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
            assert np.max(np.abs(want)) > 1e-3 and np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), name
        st.close()


def ionization_loop(api, O, nsl=48):
    """config 5 in small through the per-routine C-ABI (qpad_b200.ionization.IonizationStage) against the oracle's loop"""
    from qpad_b200 import decks
    from qpad_b200.ionization import IonizationStage
    cfg = dict(nr=96, nz=64, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3, n0=1.0e17)
    beam = dict(decks.CONFIGS["C5"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8,
                **{k: v for k, v in cfg.items()})
    orc.set_beam(*bm)
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


def sim_neutral_loop(api, O, use_graph=0, nsl=40, ion_max=1, with_plasma=False):
    """the neutral species inside qpg_sim (qpg_sim_attach_neutral; per-slice launch paths) against the oracle's ionisation loop:
    levels, number / order of the released electrons, fields, and the sim's update and iteration counters"""
    from qpad_b200 import decks
    cfg = dict(nr=96, nz=64, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
    orc = O.Sim(sp_density=1.0 if with_plasma else 0.0, neut_on=1, neut_elem=3, neut_ion_max=ion_max, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8,
                ppc1=2, ppc2=2, num_theta=8, n0=1.0e17, **cfg)
    orc.set_beam(*bm)
    upd = orc.run_slices(nsl)
    lattice = O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8)
    if not with_plasma:
        lattice = tuple(a[:0] for a in lattice)                       # input_file/ionization: nspecies 0
    sim = api.Sim(sp_npmax=2 * max(len(lattice[4]), 32), beam_npmax=len(bm[2]) + 64, use_graph=use_graph, **cfg)
    sim.init_species(*lattice)
    ne = sim.attach_neutral(3, ion_max, (2, 2), 8, n0=1.0e17)
    sim.beam.upload(*bm)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, nsl // 2)
    sim.run_slices(nsl // 2 + 1, nsl)                                 # a second range: the look-ahead deposits are redone
    u, it, sl = sim.stats()
    assert sl == nsl and it == orc.total_iters() and u == upd > 1000, (u, upd, it, orc.total_iters())
    assert np.max(np.abs(ne.levels() - orc.levels(ion_max))) < 1e-10
    ox, op, og, opsi, oq = orc.neutral()
    gx, gp, gg, gpsi, gq = ne.part.download()
    assert len(gq) == len(oq) > 100 and np.max(np.abs(gq - oq)) < 1e-12 and np.max(np.abs(gx - ox)) < 1e-7 and np.max(np.abs(gp - op)) < 1e-7
    for name in ("psi", "e", "b", "cu", "q_spe"):
        got, want = sim.field(name).download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
        assert np.max(np.abs(want)) > 1e-2 and np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), name
    sim.close()


def sim_neutral_full_step(api, O, use_graph=0):
    """a complete 3D step with the neutral attached (beam push, renewal of the neutral: levels reset, electrons and ions cleared,
    rho_ion zeroed), then the first slices of the next step, against the oracle"""
    from qpad_b200 import decks
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
    orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8, n0=1.0e17, **cfg)
    orc.set_beam(*bm)
    orc.step3d(1)
    orc.run_slices(12)
    empty = tuple(a[:0] for a in O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8))
    sim = api.Sim(sp_npmax=64, beam_npmax=len(bm[2]) + 64, use_graph=use_graph, **cfg)
    sim.init_species(*empty)
    ne = sim.attach_neutral(3, 1, (2, 2), 8, n0=1.0e17)
    sim.beam.upload(*bm)
    sim.step3d()
    assert ne.part.npp() == 0 and ne.part_add.npp() == 0
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, 12)
    assert np.max(np.abs(ne.levels() - orc.levels(1))) < 1e-10 and ne.part.npp() == len(orc.neutral()[4]) > 20
    for name in ("psi", "e", "rho_ion"):
        got = sim.field(name).download_f2()[:, :12]
        want = orc.field(name, 2)[:, :12] if name != "rho_ion" else None
        if want is not None:
            assert np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), name
        else:
            assert np.max(np.abs(got)) > 0          # the ions' charge of this step only (zeroed by the renewal)
    sim.close()


def neutral_local_pipeline(api, O, S, nwaves=3, graph_unroll=None):
    """config 5 in small on the xi-pipeline (the deck is `nodes [1,2]`): S stages on one GPU, every stage with the neutral attached to its
    sim; the released electrons, the ions' buffer, rho_ion and the ionisation levels travel forward with the plasma hand-off (neut%psend /
    precv, neutral_class.f03:1025-1101).  `nwaves` 3D steps against the oracle's S-stage run: the wake of a downstream slab is driven by
    electrons released upstream, so its fields agree only if the neutral's state arrived."""
    from qpad_b200 import decks
    from qpad_b200.pipeline import LocalPipeline
    cfg = dict(nr=64, nz=36, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3, ppc1=2, ppc2=2, num_theta=8, n0=1.0e17,
               neutral=dict(element=3, ion_max=2))
    keys = ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
    orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=2, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8, n0=1.0e17, nstages=S,
                **{k: cfg[k] for k in keys})
    orc.set_beam(*bm)
    upd_o = sum(orc.step3d(k + 1) for k in range(nwaves))
    empty = tuple(a[:0] for a in O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8))
    lp = LocalPipeline(cfg, empty, bm, S, graph_unroll=graph_unroll)       # None: the pipeline's own choice (WHILE node below 6 stages)
    for _ in range(nwaves):
        lp.wave()
    lp.drain()
    upd, iters, slices = lp.stats()
    assert slices == nwaves * cfg["nz"] and iters == orc.total_iters() and upd == upd_o > 100, (upd, upd_o, iters, orc.total_iters())
    for r, sim in enumerate(lp.sims):
        nzp = sim.nzp
        for name in ("psi", "e", "b"):
            got, want = sim.field(name).download_f2()[:, :nzp], orc.field(name, 2, stage=r)[:, :nzp]
            assert np.max(np.abs(want)) > 1e-6, (r, name)
            assert np.max(np.abs(got - want)) < 1e-7 * np.max(np.abs(orc.field(name, 2, stage=0))) + 1e-7 * np.max(np.abs(want)), (r, name)
        gx, gp, gq = sim.beam.download()
        ox, op, oq = orc.beam(stage=r)
        assert len(gq) == len(oq) and np.array_equal(gq, oq)
        # the renewal at the end of the last step has emptied the neutral on both sides
        assert sim.neutral.part.npp() == 0 and len(orc.neutral(stage=r)[4]) == 0
    lp.close()


def beam_spin_pipeline(api, S=2, nwaves=3):
    """a beam with spin on the xi-pipeline: the spin vectors are pushed in every stage (split push included), follow their particles through
    update_bound and cross the stage boundary in the 10-real hand-off record -- the pipelined run must reproduce the one-stage run of the same
    3D steps particle by particle (particles are tagged through their charge; the one-stage spin push is held against the oracle in
    tests/test_gpu_parity.py::test_beam_spin_push_matches_oracle)"""
    from qpad_b200 import decks
    from qpad_b200.pipeline import LocalPipeline
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=2, ppc2=2, num_theta=8, iter_max=2, iter_reltol=1e-3, iter_abstol=1e-3)
    beam = dict(decks.CONFIGS["C1"]["beam"], gamma=40.0)               # a slow beam: particles slip backwards in xi and cross the slab edges
    bx, bp, bq = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    n = len(bq)
    bq = bq * (1.0 + 1e-9 * np.arange(n))                              # unique tags
    rng = np.random.default_rng(5)
    spin = rng.standard_normal((n, 3)); spin /= np.linalg.norm(spin, axis=1)[:, None]
    amm = 0.00115965
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    keys = ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")
    one = api.Sim(sp_npmax=2 * len(plasma[4]), beam_npmax=n + 1024, **{k: cfg[k] for k in keys})
    one.init_species(*plasma)
    one.beam.enable_spin(amm)
    one.beam.upload(bx, bp, bq); one.beam.upload_spin(spin)
    for _ in range(nwaves):
        one.step3d()
    ox, op, oq = one.beam.download()
    os_ = one.beam.download_spin()
    lp = LocalPipeline(cfg, plasma, (bx, bp, bq), S, beam_spin=(spin, amm))
    for _ in range(nwaves):
        lp.wave()
    lp.drain()
    parts = [(sim.beam.download(), sim.beam.download_spin()) for sim in lp.sims]
    gx = np.concatenate([b[0][0] for b in parts]); gq = np.concatenate([b[0][2] for b in parts]); gs = np.concatenate([b[1] for b in parts])
    crossed = sum(len(b[0][2]) for b in parts[1:]) - int(np.sum(bx[:, 2] >= lp.parts[1][0] * (cfg["zmax"] - cfg["zmin"]) / cfg["nz"]))
    assert len(gq) == len(oq) and crossed > 0, (len(gq), len(oq), crossed)          # particles did cross a stage boundary
    io, ig = np.argsort(oq), np.argsort(gq)
    assert np.array_equal(oq[io], gq[ig])
    assert np.max(np.abs(gx[ig] - ox[io])) < 1e-9 * np.max(np.abs(ox))
    assert np.max(np.abs(gs[ig] - os_[io])) < 1e-9 and np.max(np.abs(os_[io] - spin[np.argsort(bq)][np.isin(np.sort(bq), oq)])) > 1e-6
    assert np.max(np.abs(np.linalg.norm(gs, axis=1) - 1.0)) < 1e-11
    lp.close(); one.close()


def sim_subcyc_loop(api, O, with_neutral=False):
    """the sub-cycling variant inside qpg_sim (qpg_sim_set_subcyc) against the oracle's sub-cycled loop: number of sub-steps,
    iterations, update counter, fields, particle momenta (clamped)"""
    from qpad_b200 import decks
    if with_neutral:
        cfg = dict(nr=96, nz=64, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
        bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
        okw = dict(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, n0=1.0e17)
        nsl, cases = 48, ((1.002, 1.01, 1e-4),)
    else:
        cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
        bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C1"]["beam"]))
        okw = {}
        nsl, cases = 24, ((1.1, 50.0, 1e-3), (1.05, 1.6, 1e-3))
    for efm, clamp, dtmin in cases:
        orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, subcyc_on=1, subcyc_exp_fac_max=efm, subcyc_exp_fac_clamped=clamp, subcyc_dt_min=dtmin, **okw, **cfg)
        orc.set_beam(*bm)
        upd = orc.run_slices(nsl)
        lattice = O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8)
        if with_neutral:
            lattice = tuple(a[:0] for a in lattice)
        sim = api.Sim(sp_npmax=2 * max(len(lattice[4]), 32), beam_npmax=len(bm[2]) + 64, **cfg)
        sim.init_species(*lattice)
        ne = sim.attach_neutral(3, 1, (2, 2), 8, n0=1.0e17) if with_neutral else None
        sim.set_subcyc(efm, clamp, dtmin)
        sim.beam.upload(*bm)
        sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
        sim.run_slices(1, nsl)
        u, it, sl = sim.stats()
        assert sim.subcycles() == orc.total_subcycles() > nsl, (sim.subcycles(), orc.total_subcycles())
        assert sl == nsl and it == orc.total_iters() and u == upd, (u, upd, it, orc.total_iters())
        for name in ("psi", "e", "b", "cu"):
            got, want = sim.field(name).download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
            assert np.max(np.abs(want)) > 1e-3 and np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), name
        ox, op, og, opsi, oq = orc.neutral() if with_neutral else orc.plasma()
        gx, gp, gg, gpsi, gq = (ne.part if with_neutral else sim.species).download()
        assert len(gq) == len(oq) > 50 and np.max(np.abs(gp - op)) < 1e-7 and np.max(np.abs(gx - ox)) < 1e-7
        sim.close()


def fastmath_accuracy(api, max_ulp=2.0):
    """the MUFU-seeded reciprocal / square root of the momentum arithmetic (particles.cu fast_rcp / fast_sqrt) against IEEE over the
    range of their arguments (gamma-like, 1e-3 .. 1e6): error in units of the last place"""
    ctx = api.Ctx(32, 1, 0.1, 0.02)
    rng = np.random.default_rng(5)
    x = np.concatenate([10.0 ** rng.uniform(-3, 6, 200000), 1.0 + rng.random(100000) * 1e-3, np.array([1.0, 2.0, 4.0, 0.5, 3.0, 1e-3, 1e6])])
    r, q = ctx.debug_fastmath(x)
    want_r = (np.longdouble(1.0) / x.astype(np.longdouble))
    want_q = np.sqrt(x.astype(np.longdouble))
    ulp_r = np.abs((r.astype(np.longdouble) - want_r) / np.spacing(np.asarray(want_r, dtype=np.float64)).astype(np.longdouble))
    ulp_q = np.abs((q.astype(np.longdouble) - want_q) / np.spacing(np.asarray(want_q, dtype=np.float64)).astype(np.longdouble))
    assert float(ulp_r.max()) <= max_ulp and float(ulp_q.max()) <= max_ulp, (float(ulp_r.max()), float(ulp_q.max()))
    ctx.close()
    return float(ulp_r.max()), float(ulp_q.max())


def field_smooth(api, O, M, dim, kind, order, nr=64):
    """field_rho / field_jay / field_djdxi %smooth (fields/field_src_class.f03:102-271) = `order` passes of smooth_f1
    (fields/ufield_class.f03:274-339, stencil [1,2,1]) per mode plane with the per-component on-axis policy of the three source
    classes.  Bit-exact against the oracle's smooth_f1 except for the contraction of a multiply-add (<= 2 ulp per pass)."""
    L = O.lib()
    P = 2 * M + 1
    ctx = api.Ctx(nr, M, 0.07, 0.02)
    rng = np.random.default_rng(100 * kind + 10 * M + dim)
    f = rng.normal(size=(P, nr + 2, dim))
    fld = api.Field(ctx, dim); fld.upload(f)
    fld.smooth(order, kind)
    got = fld.download()
    want = f.copy()
    for _ in range(order):
        for pl in range(P):
            m = (pl + 1) // 2
            if kind == 0:
                ax = [m == 0]
            elif kind == 1:
                ax = [False, False, True] if m == 0 else ([True, True, False] if m == 1 else [False, False, False])
            else:
                ax = [m == 1, m == 1]
            plane = np.ascontiguousarray(want[pl])
            L.orc_smooth_f1(plane, dim, nr, np.asarray(ax, dtype=np.int32))
            want[pl] = plane
    scale = _mx(want)
    assert _mx(got[:, 1:nr + 1] - want[:, 1:nr + 1]) <= 4e-16 * order * scale
    assert np.array_equal(got[:, 0], f[:, 0]) and np.array_equal(got[:, nr + 1], f[:, nr + 1])      # guard cells untouched (copy_gc_f1 on one radial rank)
    if order == 0:
        assert np.array_equal(got, f)


def part2d_move(api, O):
    """move_part2d_comm (species/part2d_comm.f03:147) on a single radial partition: update_bound has already removed what left the
    box, nothing crosses a radial processor boundary -- the particle set must come back bit for bit, in order, count unchanged."""
    ctx = api.Ctx(64, 1, 0.1, 0.02)
    rng = np.random.default_rng(3)
    n = 4099
    x = rng.uniform(-4.0, 4.0, size=(n, 2)); p = rng.normal(size=(n, 3))
    g = np.sqrt(1.0 + np.sum(p * p, axis=1)); psi = rng.random(n); q = -rng.random(n)
    pt = api.Part2d(ctx, -1.0, n + 64)
    pt.move()
    assert pt.npp() == 0                                   # empty set
    pt.upload(x, p, g, psi, q)
    pt.update_bound()
    before = pt.download()
    pt.move()
    after = pt.download()
    assert pt.npp() == len(before[4]) <= n
    for a, b in zip(before, after):
        assert np.array_equal(a, b)


def neutral_overflow(api, O):
    """released electrons that do not fit the particle set must not vanish silently: the device latches the overflow and the next
    host synchronisation returns QPG_ERR_STATE (neutral.cu k_neutral_counts; the beam's wire buffer does the same)"""
    import pytest
    nr, nth = 48, 8
    ctx = api.Ctx(nr, 0, 0.1, 0.02)
    e = np.zeros((1, nr + 2, 3)); e[0, :, 0] = 400.0 / (O.lib().orc_plasma_frequency(1.0e17) * 1.708e-12)      # ionises everything at once
    fe = api.Field(ctx, 3); fe.upload(e)
    ne = api.Neutral(ctx, 1, 1, (2, 2), nth, n0=1.0e17, dt_xi=0.02)
    ne.update(fe)
    ctx.sync()                                              # fits: no error
    assert ne.part.npp() == nr * nth * 4
    ne.renew()
    ne.part.close()
    ne.part = api.Part2d(ctx, -1.0, 200)                    # far too small
    ne.update(fe)
    with pytest.raises(api.QpadError, match="did not fit"):
        ctx.sync()
    assert 200 <= ne.part.npp() <= 256 < nr * nth * 4       # clamped to the (alignment-rounded) capacity, and reported
    ne.close()

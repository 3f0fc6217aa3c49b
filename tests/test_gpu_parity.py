"""GPU parity: every C-ABI entry point against the CPU oracle on the same seeded inputs.

Gates (BASELINE.json north_star / SURVEY.md §8d): cell indices, sort permutation and compaction order bit-exact;
per-slice fields <= 1e-10 relative (inf-norm per plane); particle attributes after a push <= 1e-12.
"""
import numpy as np
import pytest
from util import plane_relerr, smooth_field, perturbed_lattice

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-10
CASES = [(64, 0), (64, 1), (250, 1), (96, 2), (1024, 1)]


@pytest.fixture(scope="module")
def mods():
    from qpad_b200 import capi
    from oracle import oracle as O
    capi.load()
    return capi, O


def _ctx(capi, nr, M, rmax=5.0, bnd=3):
    dr = rmax / nr
    return capi.Ctx(nr, M, dr, 0.02, bnd), dr


def _mk(capi, ctx, dim, arr=None, nzp=0):
    f = capi.Field(ctx, dim, nzp, nzp > 0)
    if arr is not None:
        f.upload(arr)
    return f


def test_field_roundtrip_and_arith(mods):
    capi, O = mods
    ctx, dr = _ctx(capi, 50, 2)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 52, 3))
    b = rng.standard_normal((5, 52, 3))
    fa, fb, fc = _mk(capi, ctx, 3, a), _mk(capi, ctx, 3, b), _mk(capi, ctx, 3)
    assert np.array_equal(fa.download(), a)
    capi.Field.add3(fa, fb, fc)
    assert np.array_equal(fc.download(), a + b)
    fa.add_to(fb)
    assert np.array_equal(fb.download(), b + a)
    fa.scale(0.37)
    assert np.array_equal(fa.download(), a * 0.37)
    q = rng.standard_normal((5, 52, 1))
    fq = _mk(capi, ctx, 1, q)
    fb.add_dim_to(fq, [3], [1])
    assert np.array_equal(fq.download()[..., 0], q[..., 0] + (b + a)[..., 2])
    fq.fill(0.0)
    assert not fq.download().any()
    # 2D layout + copy_slice
    f2 = capi.Field(ctx, 3, 4, True)
    vol = rng.standard_normal((5, 5, 52, 3))
    f2.upload_f2(vol)
    assert np.array_equal(f2.download_f2(), vol)
    f2.copy_slice(3, capi.COPY_2TO1)
    assert np.array_equal(f2.download(), vol[:, 2])
    f2.upload(a)
    f2.copy_slice(5, capi.COPY_1TO2)
    assert np.array_equal(f2.download_f2()[:, 4], a)


@pytest.mark.parametrize("nr,M", CASES)
@pytest.mark.parametrize("bnd", [3, 2])
def test_solves_match_oracle(mods, nr, M, bnd):
    capi, O = mods
    if bnd == 2 and nr != 64:
        pytest.skip("zero boundary covered at nr=64")
    ctx, dr = _ctx(capi, nr, M, bnd=bnd)
    L = O.lib()
    P = 2 * M + 1
    rng = np.random.default_rng(nr + M)
    relax = 1.0e-3 * (dr / 0.02) ** 2
    q = smooth_field(rng, P, nr, 1, dr)
    cu = smooth_field(rng, P, nr, 3, dr)
    dcu = smooth_field(rng, P, nr, 2, dr)
    amu = smooth_field(rng, P, nr, 3, dr)
    b0 = smooth_field(rng, P, nr, 3, dr)
    fq, fcu, fdcu, famu = _mk(capi, ctx, 1, q), _mk(capi, ctx, 3, cu), _mk(capi, ctx, 2, dcu), _mk(capi, ctx, 3, amu)

    # psi
    want = np.zeros_like(q); L.orc_solve_psi(q, want, nr, M, dr, bnd)
    fpsi = _mk(capi, ctx, 1); ctx.solve_psi(fq, fpsi); got = fpsi.download()
    assert plane_relerr(got, want) < FIELD_TOL, ("psi", plane_relerr(got, want))
    psi = want
    # bt (beam)
    want = np.zeros((P, nr + 2, 3)); L.orc_solve_bt(q, want, nr, M, dr, bnd)
    fb = _mk(capi, ctx, 3); ctx.solve_bt(fq, fb); got = fb.download()
    assert plane_relerr(got, want) < FIELD_TOL, ("bt", plane_relerr(got, want))
    # bz : writes only component 3, keeps the rest
    want = b0.copy(); L.orc_solve_bz(cu, want, nr, M, dr, bnd)
    fb.upload(b0); ctx.solve_bz(fcu, fb); got = fb.download()
    assert plane_relerr(got, want) < FIELD_TOL, ("bz", plane_relerr(got, want))
    # ez
    e0 = smooth_field(rng, P, nr, 3, dr)
    want = e0.copy(); L.orc_solve_ez(cu, want, nr, M, dr, bnd)
    fe = _mk(capi, ctx, 3, e0); ctx.solve_ez(fcu, fe); got = fe.download()
    assert plane_relerr(got, want) < FIELD_TOL, ("ez", plane_relerr(got, want))
    # bt_iter (uses the previous iterate in b)
    want = b0.copy(); L.orc_solve_bt_iter(dcu, cu, want, nr, M, dr, bnd, relax)
    fb.upload(b0); ctx.solve_bt_iter(fdcu, fcu, fb); got = fb.download()
    assert plane_relerr(got, want) < FIELD_TOL, ("bt_iter", plane_relerr(got, want))
    # et / et_beam
    want = e0.copy(); L.orc_solve_et(b0, psi, want, nr, M, dr)
    fb.upload(b0); fpsi.upload(psi); fe.upload(e0); ctx.solve_et(fb, fpsi, fe); got = fe.download()
    assert plane_relerr(got, want) < 1e-13, ("et", plane_relerr(got, want))
    want = e0.copy(); L.orc_solve_et_beam(b0, want, nr, M)
    fe.upload(e0); ctx.solve_et_beam(fb, fe)
    assert np.array_equal(fe.download(), want)
    # djdxi
    want = np.zeros((P, nr + 2, 2)); L.orc_solve_djdxi(dcu, amu, want, nr, M, dr)
    fout = _mk(capi, ctx, 2); ctx.solve_djdxi(fdcu, famu, fout); got = fout.download()
    assert plane_relerr(got, want) < 1e-13, ("djdxi", plane_relerr(got, want))


def test_solver_beats_thomas_accuracy(mods):
    """the scan solver should sit closer to the long-double solution than fp64 Thomas does (m=0, nr=1024)"""
    capi, O = mods
    nr, M = 1024, 0
    ctx, dr = _ctx(capi, nr, M)
    L = O.lib()
    r = (np.arange(nr + 2) - 1) * dr
    q = np.zeros((1, nr + 2, 1)); q[0, :, 0] = -0.5 * np.exp(-((r - 0.4) / 0.05) ** 2)   # test/TEST_field_psi.f03:55-65
    a, b, c = np.zeros(nr), np.zeros(nr), np.zeros(nr)
    L.orc_build_matrix(O.FK_PSI, 0, nr, dr, 3, 0.0, a, b, c)
    ld = -q[0, 1:nr + 1, 0].copy(); L.orc_tridiag_solve_ld(a, b, c, ld, nr)
    th = -q[0, 1:nr + 1, 0].copy(); L.orc_tridiag_solve(a, b, c, th, nr)
    fq, fpsi = _mk(capi, ctx, 1, q), _mk(capi, ctx, 1)
    ctx.solve_psi(fq, fpsi)
    got = fpsi.download()[0, 1:nr + 1, 0]
    e_gpu, e_th = np.max(np.abs(got - ld)) / np.max(np.abs(ld)), np.max(np.abs(th - ld)) / np.max(np.abs(ld))
    assert e_gpu < 1e-13 and e_gpu <= e_th + 1e-15, (e_gpu, e_th)


def test_convergence_tester(mods):
    capi, O = mods
    nr, M = 80, 2
    ctx, dr = _ctx(capi, nr, M)
    rng = np.random.default_rng(4)
    b1, b2 = smooth_field(rng, 5, nr, 3, dr), smooth_field(rng, 5, nr, 3, dr)
    f = _mk(capi, ctx, 3, b1)
    ctx.convergence_tester(f, 2, capi.CONV_RECORD)
    f.upload(b2)
    rel, ab = ctx.convergence_tester(f, 2, capi.CONV_COMPARE)
    sre = lambda b: np.abs(b[[0, 1, 3], 1:nr + 1, 1]).sum(0)
    sim = lambda b: np.abs(b[[2, 4], 1:nr + 1, 1]).sum(0)
    old = np.sqrt(np.max(sre(b1) ** 2 + sim(b1) ** 2))
    want_abs = np.sqrt(np.max((sre(b1) - sre(b2)) ** 2 + (sim(b1) - sim(b2)) ** 2))
    assert np.isclose(ab, want_abs, rtol=1e-13) and np.isclose(rel, want_abs / old, rtol=1e-13)


@pytest.mark.parametrize("nr,M", CASES)
def test_particle_kernels_match_oracle(mods, nr, M):
    capi, O = mods
    ctx, dr = _ctx(capi, nr, M)
    L = O.lib()
    P = 2 * M + 1
    rng = np.random.default_rng(100 + nr + M)
    ppc, nth = (2, 8) if nr < 1000 else (2, 16)
    x, p, g, psi, q = perturbed_lattice(O, rng, nr, dr, ppc, ppc, nth)
    n = len(q)
    part = capi.Part2d(ctx, -1.0, 2 * n)
    part.upload(x, p, g, psi, q)
    gx, gp, gg, gpsi, gq = part.download()
    assert all(np.array_equal(a, b) for a, b in ((gx, x), (gp, p), (gg, g), (gpsi, psi), (gq, q)))

    # qdeposit
    want = O.zeros_f1(1, nr, M); L.orc_qdeposit(x, q, n, dr, nr, M, want)
    fq = _mk(capi, ctx, 1); part.qdeposit(fq); got = fq.download()
    assert plane_relerr(got, want) < 1e-12, ("qdeposit", plane_relerr(got, want))

    # amjdeposit_robust
    e, b = smooth_field(rng, P, nr, 3, dr, 0.3), smooth_field(rng, P, nr, 3, dr, 0.3)
    fe, fb = _mk(capi, ctx, 3, e), _mk(capi, ctx, 3, b)
    cu, dcu, amu = O.zeros_f1(3, nr, M), O.zeros_f1(2, nr, M), O.zeros_f1(3, nr, M)
    g_o, psi_o = g.copy(), psi.copy()
    dt = 0.02
    L.orc_amjdeposit_robust(x, p, q, g_o, psi_o, n, dr, nr, M, -1.0, dt, e, b, cu, dcu, amu)
    fcu, fdcu, famu = _mk(capi, ctx, 3), _mk(capi, ctx, 2), _mk(capi, ctx, 3)
    part.amjdeposit_robust(fe, fb, fcu, famu, fdcu, dt)
    for name, f, w in (("cu", fcu, cu), ("dcu", fdcu, dcu), ("amu", famu, amu)):
        err = plane_relerr(f.download(), w)
        assert err < 1e-11, (name, err)
    _, _, gg, gpsi, _ = part.download()
    assert np.max(np.abs(gg - g_o)) < 1e-13 * np.max(np.abs(g_o))
    assert np.max(np.abs(gpsi - psi_o)) < 1e-12 * max(1.0, np.max(np.abs(psi_o)))

    # push_u_robust + push_x
    p_o, x_o = p.copy(), x.copy()
    L.orc_push_u_robust(x_o, p_o, g_o, n, dr, nr, M, -1.0, dt, e, b)
    part.push_u_robust(fe, fb, dt)
    _, gp, gg, _, _ = part.download()
    assert np.max(np.abs(gp - p_o)) < 1e-13 * np.max(np.abs(p_o)) and np.max(np.abs(gg - g_o)) < 1e-13 * np.max(g_o)
    L.orc_push_x(x_o, p_o, g_o, n, 40 * dt)
    part.push_x(40 * dt)
    gx = part.download()[0]
    assert np.max(np.abs(gx - x_o)) < 1e-13 * np.max(np.abs(x_o))

    # update_bound: exact compaction order on identical inputs (upload the oracle's positions)
    part.upload(x_o, p_o, g_o, psi_o, q)
    npp_o = L.orc_update_bound(x_o, p_o, g_o, psi_o, q_o := q.copy(), n, nr * dr)
    part.update_bound()
    gx, gp, gg, gpsi, gq = part.download()
    assert len(gq) == npp_o
    assert np.array_equal(gx, x_o[:npp_o]) and np.array_equal(gq, q_o[:npp_o]) and np.array_equal(gp, p_o[:npp_o])


@pytest.mark.parametrize("nr,M", [(64, 1), (96, 2), (250, 0)])
def test_std_pusher_kernels_match_oracle(mods, nr, M):
    """amjdeposit_std :478, push_u_std :1790 and interp_psi :2264 (with the reference's chunk quirk) against the oracle"""
    capi, O = mods
    ctx, dr = _ctx(capi, nr, M)
    L = O.lib()
    P = 2 * M + 1
    rng = np.random.default_rng(300 + nr + M)
    x, p, g, psi, q = perturbed_lattice(O, rng, nr, dr, 2, 2, 8)
    n = len(q)
    psi = 0.2 * rng.standard_normal(n) - 0.1            # 1 - qbm*psi stays away from 0 for qbm = -1
    part = capi.Part2d(ctx, -1.0, 2 * n)
    part.upload(x, p, g, psi, q)
    # interp_psi: only psi(first of each 1024-chunk) changes, to the value of the chunk's last particle
    psif = smooth_field(rng, P, nr, 1, dr, 0.3)
    fpsi = _mk(capi, ctx, 1, psif)
    psi_o = psi.copy()
    L.orc_interp_psi(x, psi_o, n, dr, nr, M, psif)
    part.interp_psi(fpsi)
    gpsi = part.download()[3]
    changed = np.nonzero(psi_o != psi)[0]
    assert len(changed) == (n + 1023) // 1024 and np.array_equal(changed, np.arange(0, n, 1024))
    assert np.max(np.abs(gpsi - psi_o)) < 1e-13 * max(1.0, np.max(np.abs(psi_o)))
    assert np.array_equal(np.delete(gpsi, changed), np.delete(psi, changed))
    # amjdeposit_std: psi untouched, gamma time-centred
    e, b = smooth_field(rng, P, nr, 3, dr, 0.3), smooth_field(rng, P, nr, 3, dr, 0.3)
    fe, fb = _mk(capi, ctx, 3, e), _mk(capi, ctx, 3, b)
    cu, dcu, amu = O.zeros_f1(3, nr, M), O.zeros_f1(2, nr, M), O.zeros_f1(3, nr, M)
    g_o, psi_in = g.copy(), psi_o.copy()
    dt = 0.02
    L.orc_amjdeposit_std(x, p, q, g_o, psi_o, n, dr, nr, M, -1.0, dt, e, b, cu, dcu, amu)
    assert np.array_equal(psi_o, psi_in)
    fcu, fdcu, famu = _mk(capi, ctx, 3), _mk(capi, ctx, 2), _mk(capi, ctx, 3)
    part.amjdeposit_std(fe, fb, fcu, famu, fdcu, dt)
    for name, f, w in (("cu", fcu, cu), ("dcu", fdcu, dcu), ("amu", famu, amu)):
        err = plane_relerr(f.download(), w)
        assert err < 1e-11, (name, err)
    _, _, gg, gpsi2, _ = part.download()
    assert np.max(np.abs(gg - g_o)) < 1e-13 * np.max(np.abs(g_o)) and np.array_equal(gpsi2, gpsi)
    # push_u_std uses the stored psi and gamma
    p_o = p.copy()
    L.orc_push_u_std(x, p_o, g_o, psi_o, n, dr, nr, M, -1.0, dt, e, b)
    part.push_u_std(fe, fb, dt)
    _, gp, gg, _, _ = part.download()
    assert np.max(np.abs(gp - p_o)) < 1e-13 * np.max(np.abs(p_o)) and np.max(np.abs(gg - g_o)) < 1e-13 * np.max(g_o)


@pytest.mark.parametrize("nr,M,std", [(64, 0, 0), (64, 1, 0), (96, 2, 0), (64, 1, 1)])
def test_pgc_pusher_kernels_match_oracle(mods, nr, M, std):
    """amjdeposit_{robust,std}_pgc :1310/:1012 and push_u_*_pgc :1967/:2094 with a given laser envelope on the grid"""
    capi, O = mods
    ctx, dr = _ctx(capi, nr, M)
    L = O.lib()
    P = 2 * M + 1
    rng = np.random.default_rng(500 + nr + M + std)
    x, p, g, psi, q = perturbed_lattice(O, rng, nr, dr, 2, 2, 8)
    n = len(q)
    psi = 0.2 * rng.standard_normal(n) - 0.1
    part = capi.Part2d(ctx, -1.0, 2 * n)
    part.upload(x, p, g, psi, q)
    e, b = smooth_field(rng, P, nr, 3, dr, 0.3), smooth_field(rng, P, nr, 3, dr, 0.3)
    las = [smooth_field(rng, P, nr, 1, dr, 0.8), smooth_field(rng, P, nr, 1, dr, 0.8), smooth_field(rng, P, nr, 3, dr, 0.5), smooth_field(rng, P, nr, 3, dr, 0.5)]
    fe, fb = _mk(capi, ctx, 3, e), _mk(capi, ctx, 3, b)
    fl = [_mk(capi, ctx, a.shape[2], a) for a in las]
    cu, dcu, amu = O.zeros_f1(3, nr, M), O.zeros_f1(2, nr, M), O.zeros_f1(3, nr, M)
    g_o, psi_o = g.copy(), psi.copy()
    dt = 0.02
    L.orc_amjdeposit_pgc(x, p, q, g_o, psi_o, n, dr, nr, M, -1.0, dt, e, b, *las, cu, dcu, amu, std)
    fcu, fdcu, famu = _mk(capi, ctx, 3), _mk(capi, ctx, 2), _mk(capi, ctx, 3)
    ptype = capi.PUSH2_STD_PGC if std else capi.PUSH2_ROBUST_PGC
    part.amjdeposit_pgc(ptype, fe, fb, fl, fcu, famu, fdcu, dt)
    for name, f, w in (("cu", fcu, cu), ("dcu", fdcu, dcu), ("amu", famu, amu)):
        err = plane_relerr(f.download(), w)
        assert err < 1e-11, (name, err)
    _, _, gg, gpsi, _ = part.download()
    assert np.max(np.abs(gg - g_o)) < 1e-13 * np.max(np.abs(g_o))
    assert np.max(np.abs(gpsi - psi_o)) < 1e-12 * max(1.0, np.max(np.abs(psi_o)))
    if std:
        assert np.array_equal(gpsi, psi)
    p_o = p.copy()
    L.orc_push_u_pgc(x, p_o, g_o, psi_o, n, dr, nr, M, -1.0, dt, e, b, *las)
    part.push_u_pgc(ptype, fe, fb, fl, dt)
    _, gp, gg, _, _ = part.download()
    assert np.max(np.abs(gp - p_o)) < 1e-13 * np.max(np.abs(p_o)) and np.max(np.abs(gg - g_o)) < 1e-13 * np.max(g_o)
    with pytest.raises(capi.QpadError):
        part.push_u_pgc(capi.PUSH2_ROBUST, fe, fb, fl, dt)


@pytest.mark.parametrize("use_graph", [1, 0])
def test_std_pusher_slice_loop(mods, use_graph):
    """the whole slice loop with push_type std (interp_psi after the psi solve, simulation_class.f03:357-359)"""
    capi, O = mods
    from qpad_b200 import decks
    cfg, beam = _deck(O, decks, M=1)
    nsl = 12
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, sp_push_type=0, **cfg)
    orc.set_beam(*beam)
    orc_upd = orc.run_slices(nsl)
    x, p, g, psi, q = O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8)
    sim = capi.Sim(sp_npmax=2 * len(q), beam_npmax=len(beam[2]) + 64, use_graph=use_graph, sp_push_std=1, **cfg)
    with pytest.raises(capi.QpadError):
        sim.set_sweep(1)                                   # the persistent kernel implements the robust pusher only
    sim.init_species(x, p, g, psi, q)
    sim.beam.upload(*beam)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, nsl)
    upd, iters, slices = sim.stats()
    assert slices == nsl and upd == orc_upd and iters == orc.total_iters()
    for name in ("psi", "e", "b", "cu"):
        got, want = sim.field(name).download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
        assert np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want)), name
    gx, gp, gg, gpsi, gq = sim.species.download()
    ox, op, og, opsi, oq = orc.plasma()
    assert np.array_equal(gq, oq) and np.max(np.abs(gx - ox)) < 1e-8 and np.max(np.abs(gpsi - opsi)) < 1e-8


def test_update_bound_heavy_loss(mods):
    capi, O = mods
    nr, M = 64, 1
    ctx, dr = _ctx(capi, nr, M)
    rng = np.random.default_rng(9)
    n = 5000
    r = rng.uniform(0.1, 7.0, n)            # ~30 % outside r = 5
    th = rng.uniform(0, 2 * np.pi, n)
    x = np.ascontiguousarray(np.stack([r * np.cos(th), r * np.sin(th)], 1))
    p = rng.standard_normal((n, 3)); g = rng.standard_normal(n); psi = rng.standard_normal(n); q = np.arange(n, dtype=float)
    for tail_out in (False, True):
        xx = x.copy()
        if tail_out:
            xx[-3:] = 9.0
        part = capi.Part2d(ctx, -1.0, n + 100)
        part.upload(xx, p, g, psi, q)
        xo, po, go, pso, qo = xx.copy(), p.copy(), g.copy(), psi.copy(), q.copy()
        npp = O.lib().orc_update_bound(xo, po, go, pso, qo, n, nr * dr)
        part.update_bound()
        gx, gp, gg, gpsi, gq = part.download()
        assert len(gq) == npp and np.array_equal(gq, qo[:npp]) and np.array_equal(gx, xo[:npp])
        part.update_bound()                  # idempotent
        assert np.array_equal(part.download()[4], qo[:npp])


@pytest.mark.parametrize("nr,ppc,nth", [(64, 2, 8), (250, 2, 16), (1024, 2, 16)])
def test_sort_bit_exact(mods, nr, ppc, nth):
    capi, O = mods
    ctx, dr = _ctx(capi, nr, 1)
    rng = np.random.default_rng(nr)
    x, p, g, psi, q = perturbed_lattice(O, rng, nr, dr, ppc, ppc, nth, jitter=3.0)
    n = len(q)
    q = np.arange(n, dtype=float)
    part = capi.Part2d(ctx, -1.0, n + 64)
    part.upload(x, p, g, psi, q)
    ix_o, ip_o = np.zeros(n, np.int32), np.zeros(n, np.int32)
    O.lib().orc_sort_idx(x, n, dr, nr, ix_o, ip_o)
    ix, ip = part.sort_index()
    assert np.array_equal(ix, ix_o), "cell keys differ"
    assert np.array_equal(ip, ip_o), "sort permutation differs"
    xo, po, go, pso, qo = x.copy(), p.copy(), g.copy(), psi.copy(), q.copy()
    O.lib().orc_sort_part2d(xo, po, go, pso, qo, n, dr, nr)
    part.sort()
    gx, gp, gg, gpsi, gq = part.download()
    assert np.array_equal(gq, qo) and np.array_equal(gx, xo) and np.array_equal(gp, po)
    keys = np.floor(np.sqrt(gx[:, 0] ** 2 + gx[:, 1] ** 2) * (1.0 / dr))
    assert np.all(np.diff(keys) >= 0)


@pytest.mark.parametrize("M,push", [(1, 1), (2, 2), (0, 1)])
def test_beam_kernels_match_oracle(mods, M, push):
    capi, O = mods
    nr, nz, nzp, noff2 = 64, 40, 16, 8
    ctx = capi.Ctx(nr, M, 5.0 / nr, 0.25)
    dr, dz = 5.0 / nr, 0.25
    L = O.lib()
    P = 2 * M + 1
    rng = np.random.default_rng(M * 7 + push)
    n = 6000
    x = np.stack([0.6 * rng.standard_normal(n), 0.6 * rng.standard_normal(n), rng.uniform(noff2 * dz, (noff2 + nzp) * dz * 0.999, n)], 1)
    p = np.stack([rng.standard_normal(n), rng.standard_normal(n), 50.0 + 5 * rng.standard_normal(n)], 1)
    q = -rng.uniform(0.5, 1.0, n) * 1e-3
    x, p = np.ascontiguousarray(x), np.ascontiguousarray(p)
    beam = capi.Part3d(ctx, -1.0, 4.0, n + 64, nz, noff2, nzp)
    beam.upload(x, p, q)
    # deposit
    guard_in = 0.01 * rng.standard_normal((P, nzp + 1, nr + 2, 1))
    want = guard_in.copy()
    L.orc_qdeposit3d(x, q, n, dr, dz, nr, nzp, noff2, M, want)
    fq = capi.Field(ctx, 1, nzp, True)
    fq.upload_f2(guard_in)
    beam.qdeposit(fq)
    got = fq.download_f2()
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) < 1e-12 * scale
    # push
    ef = np.stack([smooth_field(rng, P, nr, 3, dr, 0.2) for _ in range(nzp + 1)], 1)
    bf = np.stack([smooth_field(rng, P, nr, 3, dr, 0.2) for _ in range(nzp + 1)], 1)
    fe, fb = capi.Field(ctx, 3, nzp, True), capi.Field(ctx, 3, nzp, True)
    fe.upload_f2(ef); fb.upload_f2(bf)
    xo, po = x.copy(), p.copy()
    L.orc_push3d(xo, po, n, dr, dz, nr, nzp, noff2, M, -1.0, 4.0, push, np.ascontiguousarray(ef), np.ascontiguousarray(bf))
    beam.push(push, fe, fb)
    gx, gp, gq = beam.download()
    assert np.max(np.abs(gp - po)) < 1e-13 * np.max(np.abs(po))
    assert np.max(np.abs(gx - xo)) < 1e-13 * np.max(np.abs(xo))
    # the split push of a pipeline stage: interior (no guard slice needed) + edge == push, bit for bit; the interior pass must not have
    # touched a particle of the slab's last slice (the guard slice may not have arrived yet) and must have moved the others
    beam2 = capi.Part3d(ctx, -1.0, 4.0, n + 64, nz, noff2, nzp)
    beam2.upload(x, p, q)
    beam2.push_interior(push, fe, fb)
    hx, hp, _ = beam2.download()
    last = np.floor(x[:, 2] / dz).astype(int) - noff2 + 1 == nzp
    assert last.sum() > 100 and np.array_equal(hx[last], x[last]) and np.array_equal(hp[last], p[last])
    assert np.array_equal(hx[~last], gx[~last]) and np.array_equal(hp[~last], gp[~last])
    beam2.push_edge(push, fe, fb)
    hx, hp, _ = beam2.download()
    assert np.array_equal(hx, gx) and np.array_equal(hp, gp)
    # update_bound: exact order on identical inputs
    xo[::7, 0] = 9.0
    xo[5::11, 2] = nz * dz + 0.1
    beam.upload(xo, po, q)
    qo = q.copy()
    npp = L.orc_update_bound3d(xo, po, qo, n, nr * dr, nz * dz)
    beam.update_bound()
    gx, gp, gq = beam.download()
    assert len(gq) == npp and np.array_equal(gq, qo[:npp]) and np.array_equal(gx, xo[:npp])


@pytest.mark.parametrize("M,push", [(1, 1), (2, 2)])
def test_beam_spin_push_matches_oracle(mods, M, push):
    """a beam with spin (part3d%has_spin, amm): push_spin_part3d :578-638 inside both pushers (the Boris pusher calls it before it stores the
    new momentum, the reduced one after), the spin vectors following their particles through update_bound (:668-670), the split push of a
    pipeline stage and the 10-real hand-off record (part3d_comm.f03:683-694)"""
    capi, O = mods
    nr, nz, nzp, noff2 = 64, 40, 16, 8
    dr, dz = 5.0 / nr, 0.25
    ctx = capi.Ctx(nr, M, dr, dz)
    L = O.lib()
    P = 2 * M + 1
    rng = np.random.default_rng(100 + M * 7 + push)
    n, amm = 5000, 0.00115965
    x = np.ascontiguousarray(np.stack([0.6 * rng.standard_normal(n), 0.6 * rng.standard_normal(n), rng.uniform(noff2 * dz, (noff2 + nzp) * dz * 0.999, n)], 1))
    p = np.ascontiguousarray(np.stack([rng.standard_normal(n), rng.standard_normal(n), 50.0 + 5 * rng.standard_normal(n)], 1))
    q = -rng.uniform(0.5, 1.0, n) * 1e-3
    s = rng.standard_normal((n, 3))
    s /= np.linalg.norm(s, axis=1)[:, None]
    ef = np.ascontiguousarray(np.stack([smooth_field(rng, P, nr, 3, dr, 0.2) for _ in range(nzp + 1)], 1))
    bf = np.ascontiguousarray(np.stack([smooth_field(rng, P, nr, 3, dr, 0.2) for _ in range(nzp + 1)], 1))
    fe, fb = capi.Field(ctx, 3, nzp, True), capi.Field(ctx, 3, nzp, True)
    fe.upload_f2(ef); fb.upload_f2(bf)
    beam = capi.Part3d(ctx, -1.0, 4.0, n + 64, nz, noff2, nzp)
    assert not beam.has_spin() and beam.wire_count() == 7 * beam.wire_cap() + 1
    beam.enable_spin(amm)
    assert beam.has_spin() and beam.wire_count() == 10 * beam.wire_cap() + 1
    beam.upload(x, p, q); beam.upload_spin(s)
    assert np.array_equal(beam.download_spin(), s)
    xo, po, so = x.copy(), p.copy(), s.copy()
    L.orc_push3d_spin(xo, po, so, amm, n, dr, dz, nr, nzp, noff2, M, -1.0, 4.0, push, ef, bf)
    assert np.max(np.abs(so - s)) > 1e-3                                           # the spins have precessed
    assert np.max(np.abs(np.linalg.norm(so, axis=1) - 1.0)) < 1e-12                # a rotation
    beam.push(push, fe, fb)
    gx, gp, gq = beam.download()
    gs = beam.download_spin()
    assert np.max(np.abs(gp - po)) < 1e-13 * np.max(np.abs(po)) and np.max(np.abs(gx - xo)) < 1e-13 * np.max(np.abs(xo))
    assert np.max(np.abs(gs - so)) < 1e-12
    # the momenta and positions do not depend on the spin: same as the spin-less oracle push
    x2, p2 = x.copy(), p.copy()
    L.orc_push3d(x2, p2, n, dr, dz, nr, nzp, noff2, M, -1.0, 4.0, push, ef, bf)
    assert np.array_equal(x2, xo) and np.array_equal(p2, po)
    # split push of a pipeline stage: interior + edge == one pass, bit for bit, spins included
    beam2 = capi.Part3d(ctx, -1.0, 4.0, n + 64, nz, noff2, nzp)
    beam2.enable_spin(amm)
    beam2.upload(x, p, q); beam2.upload_spin(s)
    beam2.push_interior(push, fe, fb)
    last = np.floor(x[:, 2] / dz).astype(int) - noff2 + 1 == nzp
    hs = beam2.download_spin()
    assert last.sum() > 100 and np.array_equal(hs[last], s[last]) and np.array_equal(hs[~last], gs[~last])
    beam2.push_edge(push, fe, fb)
    assert np.array_equal(beam2.download_spin(), gs) and np.array_equal(beam2.download()[0], gx)
    # update_bound: the spin vectors move with their particles (exact order on identical inputs)
    xo[::7, 0] = 9.0
    xo[5::11, 2] = nz * dz + 0.1
    beam.upload(xo, po, q); beam.upload_spin(so)
    qo = q.copy()
    npp = L.orc_update_bound3d_spin(xo, po, qo, so, n, nr * dr, nz * dz)
    beam.update_bound()
    gx, gp, gq = beam.download()
    assert len(gq) == npp < n and np.array_equal(gq, qo[:npp]) and np.array_equal(gx, xo[:npp]) and np.array_equal(beam.download_spin(), so[:npp])
    # forward hand-off: 10 reals per particle, the receiving set appends particles and spins together
    c2 = capi.Ctx(nr, M, dr, 0.5)
    nb = 2000
    bx = np.ascontiguousarray(np.stack([0.5 * rng.standard_normal(nb), 0.5 * rng.standard_normal(nb), rng.uniform(0, 5.6, nb)], 1))
    bp = rng.standard_normal((nb, 3)); bq = np.arange(nb, dtype=float)
    bs = rng.standard_normal((nb, 3))
    b0 = capi.Part3d(c2, -1.0, 1.0, nb + 64, 20, 0, 10)
    b1 = capi.Part3d(c2, -1.0, 1.0, nb + 64, 20, 10, 10)
    b0.enable_spin(amm); b1.enable_spin(amm)
    b0.upload(bx, bp, bq); b0.upload_spin(bs)
    hb = _DevBuf(capi, b0.wire_count())
    b0.pack_forward(hb.ptr); c2.sync()
    go = np.nonzero(bx[:, 2] >= 10 * 0.5)[0]
    rec = hb.numpy()
    assert rec[0] == len(go) > 100
    recs = rec[1:1 + 10 * len(go)].reshape(-1, 10)
    assert np.array_equal(recs[:, 6], bq[go]) and np.array_equal(recs[:, 7:10], bs[go]) and np.array_equal(recs[:, 0:3], bx[go])
    kept_q, kept_s = b0.download()[2], b0.download_spin()
    assert np.array_equal(kept_s, bs[kept_q.astype(int)])                            # still aligned after the holes were filled
    b1.unpack(hb.ptr)
    got_q, got_s = b1.download()[2], b1.download_spin()
    assert np.array_equal(got_q, bq[go]) and np.array_equal(got_s, bs[go])


def test_beam_spin_on_the_pipeline(mods):
    import kernel_cases as K
    K.beam_spin_pipeline(mods[0])


def _deck(O, decks, nr=64, nz=40, M=1, iter_max=3, **kw):
    cfg = dict(nr=nr, nz=nz, max_mode=M, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=iter_max, iter_reltol=1e-3, iter_abstol=1e-3)
    cfg.update(kw)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    if M >= 2:
        beam["center"] = (0.0376, 0.0, -2.5)      # off-axis like the hosing deck, excites m >= 1
    bx, bp, bq = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    return cfg, (bx, bp, bq)


# slab drivers: the persistent sweep kernel (default), per-slice CUDA graph with cluster or op-list field programs,
# plain stream launches
PATHS = {"sweep": dict(sweep=1), "graph-cluster": dict(sweep=0, use_graph=1, fused=1), "graph-oplist": dict(sweep=0, use_graph=1, fused=0),
         "stream-cluster": dict(sweep=0, use_graph=0, fused=1)}


def _gpu_sim(capi, O, cfg, beam, ppc=2, nth=8, use_graph=0, sort_freq=0, fused=None, sweep=None):
    x, p, g, psi, q = O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], ppc, ppc, nth)
    sim = capi.Sim(sp_npmax=2 * len(q), beam_npmax=len(beam[2]) + 64, use_graph=use_graph, sort_freq=sort_freq, **cfg)
    if fused is not None:
        sim.set_fused(fused)
    if sweep is not None:
        sim.set_sweep(sweep)
    sim.init_species(x, p, g, psi, q)
    sim.beam.upload(*beam)
    return sim, len(q)


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("M", [1, 2, 0])
def test_slice_loop_matches_oracle(mods, M, path):
    capi, O = mods
    from qpad_b200 import decks
    cfg, beam = _deck(O, decks, M=M)
    nsl = 16
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, **cfg)
    orc.set_beam(*beam)
    orc_upd = orc.run_slices(nsl)
    sim, np0 = _gpu_sim(capi, O, cfg, beam, **PATHS[path])
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, nsl)
    upd, iters, slices = sim.stats()
    assert slices == nsl and upd == orc_upd
    assert iters == orc.total_iters(), (iters, orc.total_iters())
    for name in ("psi", "e", "b", "b_spe", "e_spe", "cu", "q_spe"):
        got = sim.field(name).download_f2()[:, :nsl]
        want = orc.field(name, 2)[:, :nsl]
        scale = np.max(np.abs(want))
        assert scale > 0
        assert np.max(np.abs(got - want)) < 1e-8 * scale, (name, np.max(np.abs(got - want)) / scale)
    gx, gp, gg, gpsi, gq = sim.species.download()
    ox, op, og, opsi, oq = orc.plasma()
    assert len(gq) == len(oq) and np.array_equal(gq, oq)
    assert np.max(np.abs(gx - ox)) < 1e-8 and np.max(np.abs(gp - op)) < 1e-8


@pytest.mark.parametrize("nr,M,ppc,nth", [(250, 1, 2, 8), (1024, 1, 2, 8), (300, 2, 2, 8), (1024, 0, 1, 8), (65, 1, 2, 8), (33, 2, 2, 8), (24, 1, 2, 8)])
def test_sweep_kernel_multi_cta_team(mods, nr, M, ppc, nth):
    """the persistent sweep kernel with a field team of several CTAs (one 32-node strip each): scan totals cross CTAs as
    flagged words, residual maxima through the exchange slab; whole-loop parity with the oracle incl. PC iteration
    counts.  nr = 65 / 33 end in a one-node strip (no halo shortcut: second team barrier path), nr = 24 is one strip."""
    capi, O = mods
    from qpad_b200 import decks
    cfg, beam = _deck(O, decks, nr=nr, nz=40, M=M, iter_max=4)
    nsl = 14
    orc = O.Sim(ppc1=ppc, ppc2=ppc, num_theta=nth, **cfg)
    orc.set_beam(*beam)
    orc_upd = orc.run_slices(nsl)
    sim, np0 = _gpu_sim(capi, O, cfg, beam, ppc=ppc, nth=nth, sweep=1)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, 5)          # two launches: the look-ahead deposit and the compaction carry over
    sim.run_slices(6, nsl)
    upd, iters, slices = sim.stats()
    prof = sim.sweep_profile()
    assert slices == nsl and upd == orc_upd and prof["slices"] == nsl
    assert iters == orc.total_iters() == prof["amj_phases"], (iters, orc.total_iters(), prof)
    for name in ("psi", "e", "b", "b_spe", "e_spe", "cu", "q_spe"):
        got = sim.field(name).download_f2()[:, :nsl]
        want = orc.field(name, 2)[:, :nsl]
        scale = np.max(np.abs(want))
        assert scale > 0
        assert np.max(np.abs(got - want)) < 1e-8 * scale, (name, np.max(np.abs(got - want)) / scale)
    gx, gp, gg, gpsi, gq = sim.species.download()
    ox, op, og, opsi, oq = orc.plasma()
    assert len(gq) == len(oq) and np.array_equal(gq, oq)
    assert np.max(np.abs(gx - ox)) < 1e-8 and np.max(np.abs(gp - op)) < 1e-8


@pytest.mark.parametrize("path", ["sweep", "graph-cluster", "graph-oplist"])
@pytest.mark.parametrize("M", [0, 1, 2])
def test_one_slice_from_identical_state(mods, M, path):
    """the north-star gate: <= 1e-10 relative per-slice field error after ONE slice from identical inputs, taken in
    the wake (slice 21 of 40) where every field is O(1)"""
    capi, O = mods
    from qpad_b200 import decks
    cfg, beam = _deck(O, decks, M=M)
    k = 20
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, **cfg)
    orc.set_beam(*beam)
    orc.run_slices(k)
    sim, np0 = _gpu_sim(capi, O, cfg, beam, **PATHS[path])
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.species.upload(*orc.plasma())                      # state carried between slices: particles, cu, b_spe
    sim.field("cu").upload(orc.field("cu", 1))
    sim.field("b_spe").upload(orc.field("b_spe", 1))
    orc.run_range(k + 1, k + 1)
    sim.run_slices(k + 1, k + 1)
    worst = {}
    for name in ("psi", "e", "b", "b_spe", "e_spe", "cu", "dcu", "amu", "acu", "q_spe"):
        worst[name] = plane_relerr(sim.field(name).download(), orc.field(name, 1))
    assert np.max(np.abs(orc.field("psi", 1))) > 1e-2
    assert max(worst.values()) < FIELD_TOL, worst
    gx, gp, gg, gpsi, gq = sim.species.download()
    ox, op, og, opsi, oq = orc.plasma()
    assert np.array_equal(gq, oq)
    assert np.max(np.abs(gx - ox)) < 1e-12 and np.max(np.abs(gp - op)) < 1e-11 * max(1.0, np.max(np.abs(op)))


@pytest.mark.parametrize("path", ["sweep", "graph-cluster"])
def test_full_3d_step_with_beam_push(mods, path):
    capi, O = mods
    from qpad_b200 import decks
    cfg, beam = _deck(O, decks, nr=64, nz=32, M=1, iter_max=2)
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, **cfg)
    orc.set_beam(*beam)
    sim, np0 = _gpu_sim(capi, O, cfg, beam, **PATHS[path])
    for step in range(2):
        orc.step3d(step + 1)
        sim.step3d()
    for name in ("psi", "e"):
        got = sim.field(name).download_f2()[:, :-1]
        want = orc.field(name, 2)[:, :-1]
        assert np.max(np.abs(got - want)) < 1e-6 * np.max(np.abs(want)), name
    gx, gp, gq = sim.beam.download()
    ox, op, oq = orc.beam()
    assert len(gq) == len(oq) and np.array_equal(gq, oq)
    assert np.max(np.abs(gx - ox)) < 1e-9 * np.max(np.abs(ox)) and np.max(np.abs(gp - op)) < 1e-9 * np.max(np.abs(op))
    # beam centroid / emittance (proj_popas/part3d_popas_class.f03:69-110), 1e-6 gate
    def emit(x, p, q):
        w = q / q.sum()
        mx, mp = (w * x[:, 0]).sum(), (w * p[:, 0]).sum()
        return mx, np.sqrt((w * (x[:, 0] - mx) ** 2).sum() * (w * (p[:, 0] - mp) ** 2).sum() - ((w * (x[:, 0] - mx) * (p[:, 0] - mp)).sum()) ** 2)
    (cg, eg), (co, eo) = emit(gx, gp, gq), emit(ox, op, oq)
    assert abs(eg - eo) < 1e-6 * eo and abs(cg - co) < 1e-6 * max(abs(co), 1e-3)


@pytest.mark.parametrize("path", ["sweep", "graph-cluster"])
def test_sorted_loop_still_matches(mods, path):
    """periodic counting sort changes the particle order, not the physics"""
    capi, O = mods
    from qpad_b200 import decks
    cfg, beam = _deck(O, decks, M=1)
    nsl = 12
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, sort_freq=4, **cfg)
    orc.set_beam(*beam)
    orc.run_slices(nsl)
    sim, np0 = _gpu_sim(capi, O, cfg, beam, sort_freq=4, **PATHS[path])
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, nsl)
    got, want = sim.field("psi").download_f2()[:, :nsl], orc.field("psi", 2)[:, :nsl]
    assert np.max(np.abs(got - want)) < 1e-8 * np.max(np.abs(want))
    gq, oq = sim.species.download()[4], orc.plasma()[4]
    assert np.array_equal(gq, oq)


class _DevBuf:
    """a zeroed fp64 wire buffer in device memory: torch on the GPU; in the host emulation (tests/test_emu_parity.py) device memory
    is the host heap, so a numpy array serves"""

    def __init__(self, capi, n):
        self.emu = hasattr(capi.load(), "emu_launches")
        if self.emu:
            self.a = np.zeros(n)
            self.ptr = self.a.ctypes.data
        else:
            import torch
            self.a = torch.zeros(n, dtype=torch.float64, device="cuda")
            self.ptr = self.a.data_ptr()

    def numpy(self): return self.a.copy() if self.emu else self.a.cpu().numpy()


def test_wire_formats(mods):
    """pipe_send/recv buffers: field slice (dim, nr+2, 2M+1), plasma 8 doubles/particle, beam 7 doubles/particle"""
    capi, O = mods
    nr, M = 40, 1
    ctx, dr = _ctx(capi, nr, M)
    rng = np.random.default_rng(2)
    a = rng.standard_normal((3, nr + 2, 3))
    f = capi.Field(ctx, 3, 3, True); f.upload(a)
    buf = _DevBuf(capi, f.wire_count())
    f.pack(0, buf.ptr); ctx.sync()
    assert np.array_equal(buf.numpy().reshape(3, nr + 2, 3), a)
    f.unpack(2, buf.ptr, add=False); f.unpack(2, buf.ptr, add=True); ctx.sync()
    assert np.array_equal(f.download_f2()[:, 1], 2 * a)
    x, p, g, psi, q = perturbed_lattice(O, rng, nr, dr, 2, 2, 8)
    n = len(q)
    pa, pb = capi.Part2d(ctx, -1.0, 2 * n), capi.Part2d(ctx, -1.0, 2 * n)
    pa.upload(x, p, g, psi, q)
    wb = _DevBuf(capi, pa.wire_count())
    pa.pack(wb.ptr); ctx.sync()
    rec = wb.numpy()
    assert rec[0] == n
    assert np.array_equal(rec[1:1 + 8 * n].reshape(n, 8), np.column_stack([x, p, g, psi, q]))
    pb.unpack(wb.ptr)
    assert all(np.array_equal(u, v) for u, v in zip(pb.download(), (x, p, g, psi, q)))
    # beam forward hand-off
    nz, nzp = 20, 10
    c2 = capi.Ctx(nr, M, dr, 0.5)
    nb = 3000
    bx = np.ascontiguousarray(np.stack([0.5 * rng.standard_normal(nb), 0.5 * rng.standard_normal(nb), rng.uniform(0, 5.6, nb)], 1))
    bp = rng.standard_normal((nb, 3)); bq = np.arange(nb, dtype=float)
    b0 = capi.Part3d(c2, -1.0, 1.0, nb + 64, nz, 0, nzp)
    b1 = capi.Part3d(c2, -1.0, 1.0, nb + 64, nz, nzp, nzp)
    b0.upload(bx, bp, bq)
    hb = _DevBuf(capi, 7 * b0.wire_cap() + 1)
    b0.pack_forward(hb.ptr); c2.sync()
    go = np.nonzero(bx[:, 2] >= nzp * 0.5)[0]
    rec = hb.numpy()
    assert rec[0] == len(go)
    assert np.array_equal(rec[1:1 + 7 * len(go)].reshape(-1, 7)[:, 6], bq[go])      # packed in ascending index order
    keep = b0.download()[2]
    # "fill the holes inversely" (part3d_comm.f03:733-745)
    exp = list(bq); npp = nb
    for h in go[::-1]:
        exp[h] = exp[npp - 1]; npp -= 1
    assert np.array_equal(keep, np.array(exp[:npp]))
    b1.unpack(hb.ptr)
    assert np.array_equal(np.sort(b1.download()[2]), np.sort(bq[go]))


def test_full_size_properties_c2(mods):
    """BASELINE.json configs[1] sizes (nr = 1024, 262 144 plasma particles per slice), checked through size-independent
    properties instead of the (slow) oracle: charge conservation of the deposits, sortedness + permutation of the
    counting sort, exact pack/unpack round trip, update counters and a quiet neutral plasma through the sweep kernel."""
    import torch
    capi, O = mods
    from qpad_b200 import decks
    cfg = {k: v for k, v in decks.CONFIGS["C2"].items() if k != "beam"}
    nr, M, dr = cfg["nr"], cfg["max_mode"], cfg["rmax"] / cfg["nr"]
    x, p, g, psi, q = decks.plasma_uniform(nr, cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    n = len(q)
    assert n == 262144
    rng = np.random.default_rng(11)
    xs = x + 0.3 * dr * rng.standard_normal(x.shape)                      # scrambled: particles change cells
    ctx, _ = _ctx(capi, nr, M)
    part = capi.Part2d(ctx, -1.0, 2 * n)
    part.upload(xs, p, g, psi, q)
    part.update_bound()                                                   # the scramble pushed a few particles past r_max
    xs, p, g, psi, q = part.download()
    n1 = len(q)
    assert 0 < n - n1 < 2000
    # (a) charge conservation: sum over nodes of (j-1) * rho_0(j) (the 1/(j-1) of the epilogue undone; node 1 carries a
    #     factor 8) returns the deposited charge, whatever the particle order
    fq = _mk(capi, ctx, 1); part.qdeposit(fq); rho = fq.download()[0, :, 0]
    j = np.arange(nr + 2, dtype=float)
    tot = np.sum(rho[2:] * (j[2:] - 1.0)) + rho[1] / 8.0
    assert abs(tot - q.sum()) < 1e-10 * abs(q.sum())
    # (b) counting sort: keys non-decreasing afterwards, the particle multiset is unchanged
    key = lambda xx: np.floor(np.hypot(xx[:, 0], xx[:, 1]) / dr).astype(np.int64)
    part.sort()
    sx, sp, sg, spsi, sq = part.download()
    assert np.all(np.diff(np.minimum(key(sx), nr - 1)) >= 0)
    o0, o1 = np.lexsort((xs[:, 1], xs[:, 0])), np.lexsort((sx[:, 1], sx[:, 0]))
    assert np.array_equal(xs[o0], sx[o1]) and np.array_equal(q[o0], sq[o1])
    # (c) pipeline wire record: pack -> unpack into a second particle object is the identity
    assert len(sq) == n1
    wb = torch.zeros(part.wire_count(), dtype=torch.float64, device="cuda")
    part.pack(wb.data_ptr())
    other = capi.Part2d(ctx, -1.0, 2 * n)
    other.unpack(wb.data_ptr())
    assert all(np.array_equal(u, v) for u, v in zip(other.download(), (sx, sp, sg, spsi, sq)))
    # (d) the sweep kernel at full size: a neutral plasma without beam stays quiet, every particle is updated every slice
    sim = capi.Sim(sp_npmax=2 * n, beam_npmax=64, use_graph=1, **{k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")})
    x, p, g, psi, q = decks.plasma_uniform(nr, cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    sim.init_species(x, p, g, psi, q)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    nsl = 16
    sim.run_slices(1, nsl)
    upd, iters, slices = sim.stats()
    assert upd == nsl * n and slices == nsl and iters == nsl
    for name in ("psi", "e", "b"):
        assert np.max(np.abs(sim.field(name).download_f2()[:, :nsl])) < 1e-9, name
    assert sim.species.npp() == n


@pytest.mark.parametrize("S", [2, 3])
def test_local_pipeline_matches_oracle(mods, S):
    """the xi-pipeline on ONE GPU (S sweep kernels on S streams, SM-partitioned, event-ordered hand-offs) reproduces the
    oracle's S-stage run: fields of every slab and the beam of every stage after the same number of 3D steps"""
    capi, O = mods
    from qpad_b200 import decks
    from qpad_b200.pipeline import LocalPipeline
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=2, ppc2=2, num_theta=8, iter_max=2,
               iter_reltol=1e-3, iter_abstol=1e-3)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    lp = LocalPipeline(cfg, plasma, bm, S)
    nwaves = 4
    for _ in range(nwaves):
        lp.wave()
    lp.drain()                                   # every stage has now finished 3D steps 0 .. nwaves-1
    kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, nstages=S, **kw)
    orc.set_beam(*bm)
    for k in range(nwaves):
        orc.step3d(k + 1)
    upd, iters, slices = lp.stats()
    assert slices == nwaves * cfg["nz"] and iters == orc.total_iters()
    nb = 0
    for r, sim in enumerate(lp.sims):
        nzp = sim.nzp
        for name in ("psi", "e", "b"):
            got, want = sim.field(name).download_f2()[:, :nzp], orc.field(name, 2, stage=r)[:, :nzp]
            assert np.max(np.abs(got - want)) < 1e-6 * np.max(np.abs(want)), (r, name)
        gx, gp, gq = sim.beam.download()
        ox, op, oq = orc.beam(stage=r)
        assert len(gq) == len(oq) and np.array_equal(gq, oq)
        if len(oq):
            assert np.max(np.abs(gx - ox)) < 1e-9 * np.max(np.abs(ox))
        nb += len(oq)
    assert nb > 0
    lp.close()


def test_local_pipeline_with_unequal_slabs(mods):
    """cost-balanced (unequal) xi slabs change nothing but the schedule: a 3-stage pipeline over slabs of 5 / 17 / 10
    slices reproduces the oracle's single-stage run, and probe_partition returns a valid tiling"""
    capi, O = mods
    from qpad_b200 import decks
    from qpad_b200.pipeline import LocalPipeline, probe_partition
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, ppc1=2, ppc2=2, num_theta=8, iter_max=2,
               iter_reltol=1e-3, iter_abstol=1e-3)
    beam = dict(decks.CONFIGS["C1"]["beam"])
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    parts = [(0, 5), (5, 17), (22, 10)]
    lp = LocalPipeline(cfg, plasma, bm, 3, partition=parts)
    nwaves = 4
    for _ in range(nwaves):
        lp.wave()
    lp.drain()
    kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
    orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, nstages=1, **kw)
    orc.set_beam(*bm)
    for k in range(nwaves):
        orc.step3d(k + 1)
    upd, iters, slices = lp.stats()
    assert slices == nwaves * cfg["nz"] and iters == orc.total_iters()
    got_q = []
    for (noff, nzp), sim in zip(parts, lp.sims):
        assert (sim.noff2, sim.nzp) == (noff, nzp)
        ns, it = sim.slice_trace()
        assert np.all(ns > 0) and np.all(it >= 1) and np.all(it <= cfg["iter_max"])
        for name in ("psi", "e", "b"):
            got, want = sim.field(name).download_f2()[:, :nzp], orc.field(name, 2, stage=0)[:, noff:noff + nzp]
            assert np.max(np.abs(got - want)) < 1e-6 * np.max(np.abs(want)), (noff, name)
        got_q.append(sim.beam.download()[2])
    ox, op, oq = orc.beam(stage=0)
    assert np.array_equal(np.sort(np.concatenate(got_q)), np.sort(oq)) and len(oq) > 0
    lp.close()
    auto = probe_partition(cfg, plasma, bm, 3, 3)
    assert auto[0][0] == 0 and sum(n for _, n in auto) == cfg["nz"] and all(n >= 2 for _, n in auto)
    assert all(a + n == b for (a, n), (b, _) in zip(auto[:-1], auto[1:]))

"""GPU parity of the laser-envelope path (csrc/laser.cu) against the oracle
(oracle/qpad_oracle_laser.c): slice images + gradients, susceptibility deposit, the envelope advance and the coupled
LWFA slice loop.  Tolerances: field-level 1e-10 relative (the north star's per-slice field gate), wake line-outs 1e-6."""
import numpy as np
import pytest
from util import perturbed_lattice

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from qpad_b200 import capi
    from oracle import oracle as O
    capi.load()
    return capi, O


def _random_envelope(rng, P, nz, nr, dr, dz):
    """smooth, all planes non-zero, guards (lower xi slices, radial guards) zero like a launched pulse"""
    r = (np.arange(nr + 2) - 1) * dr
    xi = (np.arange(nz + 3) - 2) * dz
    out = []
    for _ in range(2):
        v = np.zeros((P, nz + 3, nr + 2))
        for pl in range(P):
            m = (pl + 1) // 2
            a, k, s = rng.uniform(0.3, 1.0), rng.uniform(0.5, 2.0), rng.uniform(1.0, 3.0)
            v[pl] = a * np.sin(k * xi + rng.uniform(0, 6))[:, None] * (np.abs(r) ** m * np.exp(-r * r / s))[None, :]
        v[:, :2, :] = 0.0
        v[:, -1, :] = 0.0
        v[:, :, 0] = 0.0
        v[:, :, nr + 1] = 0.0
        if P > 1:
            v[1:, :, 1] = 0.0      # m > 0 vanish on the axis
        out.append(v)
    return out


@pytest.mark.parametrize("nr,nz,M", [(64, 12, 0), (96, 10, 1), (50, 9, 2)])
def test_laser_slice_images_match_oracle(mods, nr, nz, M):
    capi, O = mods
    rmax, dz = 6.0, 0.05
    ctx = capi.Ctx(nr, M, rmax / nr, dz)
    rng = np.random.default_rng(10 + nr + M)
    ar, ai = _random_envelope(rng, 2 * M + 1, nz, nr, rmax / nr, dz)
    las = capi.Laser(ctx, nz, 20.0, 2.0, 1)
    las.upload(ar, ai)
    gar, gai = las.download()
    assert np.array_equal(gar, ar) and np.array_equal(gai, ai)
    olas = O.Laser(nr, nz, M, rmax, 0.0, nz * dz, 2.0, 20.0)
    olas.ar[:], olas.ai[:] = ar, ai
    for j in (1, 2, nz // 2, nz):
        las.slice(j)
        want_gr, want_gi = olas.set_grad(j)
        got = [las.field(k).download() for k in range(4)]
        assert np.array_equal(got[0][..., 0], ar[:, j + 1, :]) and np.array_equal(got[1][..., 0], ai[:, j + 1, :])
        scale = max(np.max(np.abs(want_gr)), np.max(np.abs(want_gi)))
        assert np.max(np.abs(got[2] - want_gr)) < 1e-13 * scale and np.max(np.abs(got[3] - want_gi)) < 1e-13 * scale


@pytest.mark.parametrize("nr,M", [(64, 0), (96, 2)])
def test_deposit_chi_matches_oracle(mods, nr, M):
    capi, O = mods
    dr = 5.0 / nr
    ctx = capi.Ctx(nr, M, dr, 0.02)
    rng = np.random.default_rng(77 + M)
    x, p, g, psi, q = perturbed_lattice(O, rng, nr, dr, 2, 2, 8)
    psi = 0.3 * rng.standard_normal(len(q)) - 0.2
    part = capi.Part2d(ctx, -1.0, 2 * len(q))
    part.upload(x, p, g, psi, q)
    las = capi.Laser(ctx, 4, 20.0, 2.0, 1)
    corr = O.lib().orc_deposit_ax_corr(2)
    las.deposit_chi(part, 3, corr)
    want = O.zeros_f1(1, nr, M)
    O.lib().orc_deposit_chi(x, q, psi, len(q), dr, nr, M, -1.0, corr, want)
    got = las.field(4).download()
    assert np.max(np.abs(got - want)) < 1e-12 * np.max(np.abs(want))
    vol = las.field(4).download_f2()
    assert np.array_equal(vol[:, 2], got) and not vol[:, 0].any() and not vol[:, 3].any()
    las.deposit_chi(part, 0, corr)                   # raw sums were cleared: the same answer again, slice image only
    assert np.max(np.abs(las.field(4).download() - want)) < 1e-12 * np.max(np.abs(want))


@pytest.mark.parametrize("nr,nz,M,iters", [(128, 48, 0, 3), (100, 24, 1, 2), (64, 16, 2, 1), (33, 8, 1, 1)])
def test_envelope_advance_matches_oracle(mods, nr, nz, M, iters):
    """set_rhs + the xi-recurrent solve (block cyclic reduction on the GPU, banded elimination in the oracle), with a
    plasma susceptibility volume, two 3D steps"""
    capi, O = mods
    rmax, zmin, zmax, ds, k0 = 8.0, -2.0, 2.0, 2.0, 20.0
    P = 2 * M + 1
    dr, dz = rmax / nr, (zmax - zmin) / nz
    ctx = capi.Ctx(nr, M, dr, dz)
    rng = np.random.default_rng(5 + nr)
    olas = O.Laser(nr, nz, M, rmax, zmin, zmax, ds, k0, iters)
    if M == 0:
        olas.launch_gaussian(1.5, 2.0, 0.0, 0.0, 1.5, 0.5, 1.5)
    else:
        olas.ar[:], olas.ai[:] = _random_envelope(rng, P, nz, nr, dr, dz)
    r = (np.arange(nr + 2) - 1) * dr
    chi = np.zeros((P, nz + 1, nr + 2, 1))
    chi[0, :, :, 0] = -1.0 - 0.5 * np.exp(-r * r)[None, :] * rng.uniform(0.5, 1.0, size=(nz + 1, 1))
    for pl in range(1, P):
        chi[pl, :, :, 0] = 0.1 * (np.abs(r) * np.exp(-r * r))[None, :] * rng.standard_normal((nz + 1, 1))
    chi[:, :, 0, 0] = 0.0
    las = capi.Laser(ctx, nz, k0, ds, iters)
    las.upload(olas.ar, olas.ai)
    las.field(4).upload_f2(chi)
    for step in range(2):
        olas.advance(chi)
        las.advance()
        gar, gai = las.download()
        scale = max(np.max(np.abs(olas.ar)), np.max(np.abs(olas.ai)))
        err = max(np.max(np.abs(gar - olas.ar)), np.max(np.abs(gai - olas.ai))) / scale
        assert err < 1e-10, (step, err)
    assert scale > 0.1


@pytest.mark.parametrize("use_graph,sweep", [(0, 0), (1, 0), (0, 1)])
def test_lwfa_slice_loop_matches_oracle(mods, use_graph, sweep, nr=128, nz=96):
    """config 4 in small: robust_pgc plasma driven by a Gaussian laser pulse, envelope advanced with the deposited
    susceptibility, two 3D steps (simulation_class.f03:294-512 with nlasers = 1, nbeams = 0)"""
    capi, O = mods
    cfg = dict(nr=nr, nz=nz, max_mode=0, rmax=12.0, zmin=-3.0, zmax=6.0, dt=2.0, iter_max=6, iter_reltol=1e-3, iter_abstol=1e-6)
    ppc1, ppc2, nth, k0, iters = 4, 2, 8, 20.0, 3
    orc = O.Sim(ppc1=ppc1, ppc2=ppc2, num_theta=nth, sp_push_type=5, laser_on=1, laser_iter=iters, laser_k0=k0, beam_evol=0, **cfg)
    olas = O.Laser(nr, nz, 0, cfg["rmax"], cfg["zmin"], cfg["zmax"], cfg["dt"], k0, iters)
    olas.launch_gaussian(1.2, 2.5, 0.0, 0.0, 1.5, 0.0, 1.5)
    orc.set_laser(olas.ar, olas.ai)
    orc.set_beam(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    x, p, g, psi, q = O.inject_uniform(nr, cfg["rmax"] / nr, ppc1, ppc2, nth)
    sim = capi.Sim(sp_npmax=2 * len(q), beam_npmax=64, sp_push_pgc=1, laser_iter=iters, laser_k0=k0, sp_ppc_r=ppc1, beam_evol=0, use_graph=use_graph, **cfg)
    assert sim.laser is not None
    sim.set_sweep(sweep)       # 1: the laser hooks inside the persistent sweep kernel (k_sweep<M, PGC = true>); 0: per-slice launches
    sim.init_species(x, p, g, psi, q)
    sim.laser.upload(olas.ar, olas.ai)
    for step in range(2):
        orc.step3d(step + 1)
        sim.step3d()
        oar, oai, ochi = orc.laser()
        gar, gai = sim.laser.download()
        gchi = sim.laser.field(4).download_f2()[..., 0]
        assert np.max(np.abs(gchi[:, :nz] - ochi[:, :nz])) < 1e-9 * np.max(np.abs(ochi)), step
        for name in ("psi", "e", "b"):
            got, want = sim.field(name).download_f2()[:, :nz], orc.field(name, 2)[:, :nz]
            assert np.max(np.abs(want)) > 1e-3
            assert np.max(np.abs(got - want)) < 1e-7 * np.max(np.abs(want)), (step, name)
        assert max(np.max(np.abs(gar - oar)), np.max(np.abs(gai - oai))) < 1e-9 * np.max(np.abs(oar)), step
    upd, iters_done, slices = sim.stats()
    assert slices == 2 * nz and iters_done == orc.total_iters()


@pytest.mark.parametrize("S", [2, 3])
def test_lwfa_local_pipeline_matches_oracle(mods, S, nr=128, nz=96, nsteps=4):
    """the envelope on the xi-pipeline (sim_lasers_class.f03:197-222, the C4 deck is `nodes [1,4]`): S stages on one GPU, each with its slab
    of the envelope; a stage advances its slab after its sweep from the NEW last two slices of the upstream stage (guard hand-off through
    flag-ordered wire buffers, capi.Laser.set_handoff).  Four 3D steps against the oracle's S-stage run -- which equals its one-stage run."""
    capi, O = mods
    from qpad_b200.pipeline import LocalPipeline
    k0, iters = 20.0, 3
    cfg = dict(nr=nr, nz=nz, max_mode=0, rmax=12.0, zmin=-3.0, zmax=6.0, dt=2.0, iter_max=6, iter_reltol=1e-3, iter_abstol=1e-6, ppc1=4, ppc2=2, num_theta=8,
               laser=dict(k0=k0, iteration=iters))
    keys = ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")
    olas = O.Laser(nr, nz, 0, cfg["rmax"], cfg["zmin"], cfg["zmax"], cfg["dt"], k0, iters)
    olas.launch_gaussian(1.2, 2.5, 0.0, 0.0, 1.5, 0.0, 1.5)
    plasma = O.inject_uniform(nr, cfg["rmax"] / nr, 4, 2, 8)
    empty = (np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    orcs = []
    for nst in (S, 1):
        orc = O.Sim(ppc1=4, ppc2=2, num_theta=8, sp_push_type=5, laser_on=1, laser_iter=iters, laser_k0=k0, beam_evol=0, nstages=nst, **{k: cfg[k] for k in keys})
        orc.set_laser(olas.ar, olas.ai)
        orc.set_beam(*empty)
        orcs.append(orc)
    parts = [(0, 5 * nz // 12), (5 * nz // 12, nz - 5 * nz // 12)] if S == 2 else None      # unequal slabs too
    lp = LocalPipeline(cfg, plasma, empty, S, partition=parts, laser=(olas.ar.copy(), olas.ai.copy()))
    for _ in range(nsteps):
        lp.wave()
    lp.drain()
    for orc in orcs:
        for k in range(nsteps):
            orc.step3d(k + 1)
    oar, oai, ochi = orcs[0].laser()
    o1r, o1i, _ = orcs[1].laser()
    assert max(np.max(np.abs(oar - o1r)), np.max(np.abs(oai - o1i))) < 1e-12 * np.max(np.abs(o1r))      # the pipelined envelope IS the one-stage one
    assert np.max(np.abs(oar - olas.ar)) > 1e-3 * np.max(np.abs(oar))                                    # and it has moved
    upd, iters_done, slices = lp.stats()
    assert slices == nsteps * nz and iters_done == orcs[0].total_iters()
    for r, sim in enumerate(lp.sims):
        off, nzp = sim.noff2, sim.nzp
        gar, gai = sim.laser.download()
        want_r, want_i = oar[:, off + 2:off + 2 + nzp], oai[:, off + 2:off + 2 + nzp]
        assert max(np.max(np.abs(gar[:, 2:2 + nzp] - want_r)), np.max(np.abs(gai[:, 2:2 + nzp] - want_i))) < 1e-9 * np.max(np.abs(oar)), r
        if r > 0:       # the lower guards hold the upstream stage's new last two slices
            assert np.max(np.abs(gar[:, 0:2] - oar[:, off:off + 2])) < 1e-9 * np.max(np.abs(oar)), r
        gchi = sim.laser.field(4).download_f2()[..., 0]
        assert np.max(np.abs(gchi[:, :nzp] - ochi[:, off:off + nzp])) < 1e-9 * np.max(np.abs(ochi)), r
        for name in ("psi", "e", "b"):
            whole = orcs[1].field(name, 2)[:, :nz]                          # the one-stage oracle: any partition must reproduce it
            got, want = sim.field(name).download_f2()[:, :nzp], whole[:, off:off + nzp]
            assert np.max(np.abs(whole)) > 1e-3
            assert np.max(np.abs(got - want)) < 1e-7 * np.max(np.abs(whole)), (r, name)
    lp.close()

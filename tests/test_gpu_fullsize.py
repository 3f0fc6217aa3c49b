"""Full-size parity of the persistent sweep kernel path against the oracle on BASELINE.json's decks (north star: "the blowout-regime
and hosing decks reproduce the reference diagnostics within tolerance"): <= 1e-6 relative on the E_z / psi line-outs and on the beam
centroid / emittance after the deck's 3D steps, the predictor-corrector iteration count of EVERY slice equal to the oracle's, update
counters exact.  The oracle legs (tests/fullsize_cases.py, seconds to half a minute of one CPU core each) run in worker
processes while the GPU legs run."""
import concurrent.futures as cf
import multiprocessing as mp

import numpy as np
import pytest

import fullsize_cases as F

pytestmark = pytest.mark.gpu

TOL = 1e-6          # north star: E_z / psi line-outs and beam centroid / emittance after N 3D steps


@pytest.fixture(scope="module")
def oracle_runs():
    pool = cf.ProcessPoolExecutor(max_workers=4, mp_context=mp.get_context("spawn"))
    futs = {case: pool.submit(F.run_oracle, case) for case in ("C2c", "C2w", "C3", "C1")}
    yield futs
    pool.shutdown(wait=False, cancel_futures=True)


@pytest.fixture(scope="module")
def capi():
    from qpad_b200 import capi
    capi.load()
    return capi


def _gpu_sim(capi, cfg, plasma, bm):
    sim = capi.Sim(sp_npmax=2 * len(plasma[4]), beam_npmax=len(bm[2]) + 1024, use_graph=1, **{k: cfg[k] for k in F.KEYS})
    sim.init_species(*plasma)
    sim.beam.upload(*bm)
    return sim


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _compare_fields(sim, ref, nsl, tol=TOL):
    """on-axis E_z / psi line-outs at `tol` (the north star's diagnostics) and the WHOLE psi / e / b volumes (every mode plane, node,
    component, slice; max norm per slice relative to the field's maximum) at `tol` too -- except in slices the oracle itself marks as
    ill-conditioned (ref["conditioning"], fullsize_cases.run_oracle: its own response to a 1e-14 perturbation of the beam charge exceeds
    1e-9, i.e. rounding errors are amplified by more than 1e5), where the error may reach 30x that response instead"""
    worst = {}
    for name in ("psi", "e", "b"):
        got = sim.field(name).download_f2()[:, :nsl]
        want = ref[name]
        assert np.max(np.abs(want)) > 1e-3, name
        err = np.max(np.abs(got - want), axis=(0, 2, 3)) / np.max(np.abs(want))          # per slice
        allowed = np.full(nsl, tol)
        if "conditioning" in ref:
            cond = ref["conditioning"][name][:nsl]
            allowed = np.where(cond > 1e-9, np.maximum(tol, 30.0 * cond), tol)
            assert np.count_nonzero(cond > 1e-9) <= 8, "only the few slices of the closing bubble may be ill-conditioned"
        worst[name] = float(err.max())
        bad = np.nonzero(err >= allowed)[0]
        assert bad.size == 0, (name, bad[:8], err[bad[:8]], allowed[bad[:8]])
    # the diagnostics the north star names: on-axis line-outs of E_z (component 3 of e, m = 0) and psi along xi
    ez = sim.field("e").lineout(3, 0, 1)[:nsl]
    ps = sim.field("psi").lineout(1, 0, 1)[:nsl]
    assert _rel(ez, ref["e"][0, :, 1, 2]) < tol and _rel(ps, ref["psi"][0, :, 1, 0]) < tol
    return worst


def _compare_beam(sim, ref, tol=TOL):
    gx, gp, gq = sim.beam.download()
    ox, op, oq = ref["beam"]
    assert len(gq) == len(oq) and np.array_equal(gq, oq)          # same particles, same order
    assert _rel(gx, ox) < tol and _rel(gp, op) < tol
    mg, mo = F.beam_moments(gx, gp, gq), F.beam_moments(ox, op, oq)
    for ax in "xy":
        cg, sg, eg = mg[ax]; co, so, eo = mo[ax]
        assert abs(cg - co) < tol * so and abs(sg - so) < tol * so and abs(eg - eo) < tol * eo, (ax, mg[ax], mo[ax])
    assert abs(mg["pz"] - mo["pz"]) < tol * abs(mo["pz"])


def test_c1_blowout_deck_full_step(capi, oracle_runs):
    """configs[0]: input_file/blowout_regime/qpinput_tri-gaussian.json at its own size, one whole 3D step (500 slices + beam push)
    through qpg_sim's default path (k_sweep)"""
    cfg, plasma, bm, nsteps, _ = F.deck("C1")
    sim = _gpu_sim(capi, cfg, plasma, bm)
    sim.step3d()
    upd, iters, slices = sim.stats()
    ns, it = sim.slice_trace()
    ref = oracle_runs["C1"].result(timeout=600)
    assert slices == cfg["nz"] and upd == cfg["nz"] * len(plasma[4])              # nothing leaves the box in one step of this deck
    assert np.array_equal(it, ref["slice_iters"]), np.nonzero(it != ref["slice_iters"])[0][:10]
    assert iters == ref["total_iters"] and iters > 1.5 * slices                      # a real blow-out: several iterations per slice
    _compare_fields(sim, ref, cfg["nz"])
    _compare_beam(sim, ref)
    sim.close()


def test_c3_hosing_deck_two_steps(capi, oracle_runs):
    """configs[2]: input_file/hosing/qpinput.json -- max_mode 2, drive beam + off-axis witness beam, the deck's two 3D steps"""
    cfg, plasma, bm, nsteps, _ = F.deck("C3")
    assert cfg["max_mode"] == 2 and nsteps == 2
    sim = _gpu_sim(capi, cfg, plasma, bm)
    per_step = []
    for _ in range(nsteps):
        i0 = sim.stats()[1]
        sim.step3d()
        per_step.append(sim.stats()[1] - i0)
    upd, iters, slices = sim.stats()
    ns, it = sim.slice_trace()
    ref = oracle_runs["C3"].result(timeout=600)
    assert slices == nsteps * cfg["nz"]
    assert per_step == ref["iters_by_step"], (per_step, ref["iters_by_step"])
    assert np.array_equal(it, ref["slice_iters"])
    # Line-outs at 1e-6 as everywhere, volumes at 1e-6 in every slice but the last five: behind this deck's strong drive beam
    # (n_b = 93 n_0) the sheath electrons cross where the bubble closes (slices 433-437, up to 8 predictor-corrector passes; the
    # iteration counts equal the oracle's slice by slice, asserted above) and the DECK ITSELF is ill-conditioned there: the oracle run
    # with the beam charge scaled by (1 + 1e-14) differs from the unperturbed oracle by 1.6e-4 in e and b off the axis in those
    # slices (and by < 1e-12 elsewhere) -- measured by the oracle leg and used as the yardstick (see _compare_fields).  The on-axis
    # line-outs stay at 4e-10 even there.
    worst = _compare_fields(sim, ref, cfg["nz"])
    assert "conditioning" in ref
    # the m = 1, 2 planes carry the hosing signal: they must be there and agree as well
    e = sim.field("e").download_f2()[:, :cfg["nz"]]
    assert np.max(np.abs(e[1:3])) > 1e-4 and np.max(np.abs(e[3:5])) > 1e-6
    well = ref["conditioning"]["e"] <= 1e-9
    for pl in range(1, 5):
        assert np.max(np.abs(e[pl][well] - ref["e"][pl][well])) < TOL * np.max(np.abs(ref["e"][0])), pl
    _compare_beam(sim, ref)
    sim.close()


@pytest.mark.parametrize("case", ["C2w", "C2c"])
def test_c2_regime_through_the_wake(capi, oracle_runs, case):
    """configs[1]'s radial grid and particle load (nr = 1024, 262 144 plasma particles per slice) with the deck's beam density:
    C2w = 288 slices at C2's own d(xi) through the beam peak; C2c = the whole box at 4x coarser d(xi), i.e. the complete wake with
    the sheath crossing, where the oracle needs up to iter_max predictor-corrector passes"""
    cfg, plasma, bm, _, nsl = F.deck(case)
    assert cfg["nr"] == 1024 and len(plasma[4]) == 262144
    sim = _gpu_sim(capi, cfg, plasma, bm)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, nsl)
    upd, iters, slices = sim.stats()
    ns, it = sim.slice_trace()
    ref = oracle_runs[case].result(timeout=900)
    assert slices == nsl and upd == ref["updates"]
    assert np.array_equal(it[:nsl], ref["slice_iters"][:nsl]), np.nonzero(it[:nsl] != ref["slice_iters"][:nsl])[0][:10]
    if case == "C2c":
        assert ref["slice_iters"].max() >= 3                                        # the wake is non-linear in this case
    _compare_fields(sim, ref, nsl)
    gx, gp, gg, gpsi, gq = sim.species.download()
    ox, op, og, opsi, oq = ref["plasma"]
    assert len(gq) == len(oq) and np.array_equal(gq, oq)                            # same survivors in the same order
    assert np.max(np.abs(gx - ox)) < TOL * cfg["rmax"] and np.max(np.abs(gp - op)) < TOL * max(1.0, np.max(np.abs(op)))
    sim.close()

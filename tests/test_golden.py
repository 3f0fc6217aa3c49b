"""Committed fixtures (tests/golden/, made by tests/golden/make_golden.py from the oracle -- the reference ships none):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU) without the oracle in the loop."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def _close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300)


def test_oracle_reproduces_blowout_fixture():
    want = np.load(os.path.join(HERE, "golden", "blowout_small.npz"))
    got = mg.blowout()
    assert int(want["iters12"]) == got["iters12"] and int(want["beam_n"]) == got["beam_n"]
    for k in ("psi12", "e12", "b12", "ez_axis", "psi_axis"):
        assert np.max(np.abs(want[k])) > 0 and _close(got[k], want[k], 1e-12), k
    for k in ("beam_mean_x", "beam_mean_xi", "beam_rms_px"):
        assert abs(got[k] - float(want[k])) <= 1e-12 * max(abs(float(want[k])), 1e-3), k


def test_oracle_reproduces_lwfa_fixture():
    want = np.load(os.path.join(HERE, "golden", "lwfa_small.npz"))
    got = mg.lwfa()
    assert int(want["iters"]) == got["iters"]
    for k in ("psi_axis", "ez_axis", "psi_slices", "er_slices", "a_axis_r", "a_axis_i", "a_slice60_r", "chi_slice60"):
        assert np.max(np.abs(want[k])) > 0 and _close(got[k], want[k], 1e-11), k


@pytest.mark.gpu
def test_cuda_blowout_matches_fixture():
    """the smoke deck through the C-ABI (persistent sweep kernel) against the committed numbers: 1e-10 per-slice fields after
    12 slices, 1e-6 on the E_z / psi line-outs and the beam moments after a full 3D step (north-star gates)"""
    from qpad_b200 import capi, decks
    want = np.load(os.path.join(HERE, "golden", "blowout_small.npz"))
    cfg = mg.BLOWOUT
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C1"]["beam"]))
    pl = decks.plasma_uniform(cfg["nr"], cfg["rmax"], mg.BLOWOUT_PLASMA["ppc1"], mg.BLOWOUT_PLASMA["ppc2"], mg.BLOWOUT_PLASMA["num_theta"])
    sim = capi.Sim(sp_npmax=2 * len(pl[4]), beam_npmax=len(bm[2]) + 64, use_graph=1, **cfg)
    sim.init_species(*pl)
    sim.beam.upload(*bm)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    sim.run_slices(1, 12)
    for name, key in (("psi", "psi12"), ("e", "e12"), ("b", "b12")):
        assert _close(sim.field(name).download_f2()[:, :12], want[key], 1e-10), name
    sim.close()
    sim = capi.Sim(sp_npmax=2 * len(pl[4]), beam_npmax=len(bm[2]) + 64, use_graph=1, **cfg)
    sim.init_species(*pl)
    sim.beam.upload(*bm)
    sim.step3d()
    assert _close(sim.field("e").lineout(3, 0, 1), want["ez_axis"], 1e-6) and _close(sim.field("psi").lineout(1, 0, 1), want["psi_axis"], 1e-6)
    bx, bp, bq = sim.beam.download()
    w = bq / bq.sum()
    assert len(bq) == int(want["beam_n"])
    assert abs((w * bx[:, 2]).sum() - float(want["beam_mean_xi"])) < 1e-6 * abs(float(want["beam_mean_xi"]))
    assert abs(np.sqrt((w * bp[:, 0] ** 2).sum()) - float(want["beam_rms_px"])) < 1e-6 * float(want["beam_rms_px"])
    sim.close()


@pytest.mark.gpu
def test_cuda_lwfa_matches_fixture():
    from qpad_b200 import capi, decks
    want = np.load(os.path.join(HERE, "golden", "lwfa_small.npz"))
    cfg, las = mg.LWFA, dict(mg.LWFA_LASER)
    it = las.pop("iteration")
    pl = decks.plasma_uniform(cfg["nr"], cfg["rmax"], mg.LWFA_PLASMA["ppc1"], mg.LWFA_PLASMA["ppc2"], mg.LWFA_PLASMA["num_theta"])
    sim = capi.Sim(sp_npmax=2 * len(pl[4]), beam_npmax=64, sp_push_pgc=1, laser_iter=it, laser_k0=las["k0"], sp_ppc_r=mg.LWFA_PLASMA["ppc1"], beam_evol=0,
                   use_graph=1, **cfg)
    sim.init_species(*pl)
    sim.laser.upload(*decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las))
    for _ in range(2):
        sim.step3d()
    nz = cfg["nz"]
    assert _close(sim.field("psi").lineout(1, 0, 1), want["psi_axis"], 1e-6) and _close(sim.field("e").lineout(3, 0, 1), want["ez_axis"], 1e-6)
    psi = sim.field("psi").download_f2()
    assert _close(psi[0, [40, 60, 80], :, 0], want["psi_slices"], 1e-7)
    ar, ai = sim.laser.download()
    assert _close(ar[0, 2:nz + 2, 1], want["a_axis_r"], 1e-8) and _close(ai[0, 2:nz + 2, 1], want["a_axis_i"], 1e-8)
    sim.close()

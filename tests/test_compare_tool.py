"""tools/compare_reference.py (the script that compares this repository's fields with dumps of the reference itself).
No reference output exists in this image, so the tool is exercised on an .npz written in the REFERENCE'S dump layout from an oracle
run: identical data must pass, a perturbed dataset must be flagged, the deck reader must understand the reference's own qpinput.json."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import compare_reference as CR  # noqa: E402

REF_DECK = "/root/reference/input_file/blowout_regime/qpinput_tri-gaussian.json"


def _small_run():
    from qpad_b200 import decks
    cfg = dict(nr=48, nz=40, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3, ppc1=2, ppc2=2, num_theta=8)
    beam = dict(decks.CONFIGS["C1"]["beam"], center=(0.05, 0.0, -2.5))
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], 2, 2, 8)
    return cfg, bm, plasma, CR.run_ours(cfg, bm, plasma, 1, "oracle")


def test_mode_part_names():
    assert [CR.part_to_plane(p) for p in ("Re0", "Re1", "Im1", "Re2", "Im2")] == [0, 1, 2, 3, 4]      # diagnostics_class.f03:996-1003
    with pytest.raises(ValueError):
        CR.part_to_plane("Im0")


def test_roundtrip_through_the_reference_layout(tmp_path):
    cfg, bm, plasma, ours = _small_run()
    nr = cfg["nr"]
    dump = {}
    for name, (dset, fld, comp) in CR.FIELDS.items():
        for part in ("Re0", "Re1", "Im1"):
            a = ours[fld][CR.part_to_plane(part), :, 1:nr + 1, comp]
            dump[f"Fields/{name}/{part}"] = a.T.copy()        # (r, xi), the orientation of the Fortran array f2(1, 1:nr, 1:nz)
    f = tmp_path / "dump.npz"
    np.savez(f, **dump)
    ref = CR.load_reference(str(f), 1)
    assert len(ref) == 21
    assert CR.compare(ref, ours, cfg, 1e-6, 1e-5, out=open(os.devnull, "w"))
    ref[("Ez", "Re0")] = ref[("Ez", "Re0")] * (1.0 + 3e-6)     # a 3e-6 discrepancy must not pass a 1e-6 gate
    assert not CR.compare(ref, ours, cfg, 1e-6, 1e-5, out=open(os.devnull, "w"))


@pytest.mark.skipif(not os.path.exists(REF_DECK), reason="the reference tree is not on this machine")
def test_reads_the_reference_deck():
    cfg, beams = CR.deck_from_json(REF_DECK)
    from qpad_b200 import decks
    want = decks.CONFIGS["C1"]
    assert all(cfg[k] == want[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "ppc1", "ppc2", "num_theta"))
    b, w = beams[0], want["beam"]
    assert b["density"] == w["density"] and tuple(b["sigma"]) == tuple(w["sigma"]) and tuple(b["center"]) == tuple(w["center"]) and b["ppc"] == w["ppc"]


def test_laser_envelope_dumps_in_the_reference_layout(tmp_path):
    """a laser deck: the envelope dumps ./Lasers1/A_laser/{Re,Im}_<Part>/a_laser_*.h5 (diagnostics_class.f03:662-691, :1005-1019) beside the
    field dumps; identical data pass, a perturbed envelope is flagged"""
    from qpad_b200 import decks
    cfg = dict(nr=48, nz=40, max_mode=0, rmax=12.0, zmin=-3.0, zmax=6.0, dt=2.0, iter_max=4, iter_reltol=1e-3, iter_abstol=1e-6, ppc1=2, ppc2=2, num_theta=8,
               laser=dict(k0=20.0, a0=1.2, w0=2.5, focal_distance=0.0, lon_center=0.0, t_rise=1.5, t_flat=0.0, t_fall=1.5, iteration=2))
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], 2, 2, 8)
    ours = CR.run_ours_laser(cfg, plasma, 1, "oracle")
    nr, nz = cfg["nr"], cfg["nz"]
    assert np.max(np.abs(ours["a_r"])) > 0.5 and np.max(np.abs(ours["psi"])) > 1e-3
    dump = {"Lasers1/A_laser/Re_Re0": ours["a_r"][0, 2:nz + 2, 1:nr + 1].T.copy(), "Lasers1/A_laser/Im_Re0": ours["a_i"][0, 2:nz + 2, 1:nr + 1].T.copy(),
            "Fields/Psi/Re0": ours["psi"][0, :, 1:nr + 1, 0].T.copy(), "Fields/Ez/Re0": ours["e"][0, :, 1:nr + 1, 2].T.copy()}
    f = tmp_path / "lwfa.npz"
    np.savez(f, **dump)
    ref = CR.load_reference(str(f), 1)
    assert set(ref) == {("A_laser_Re", "Re0"), ("A_laser_Im", "Re0"), ("Psi", "Re0"), ("Ez", "Re0")}
    assert CR.compare(ref, ours, cfg, 1e-6, 1e-5, out=open(os.devnull, "w"))
    ref[("A_laser_Im", "Re0")] = ref[("A_laser_Im", "Re0")] * (1.0 + 5e-6)
    assert not CR.compare(ref, ours, cfg, 1e-6, 1e-5, out=open(os.devnull, "w"))


@pytest.mark.skipif(not os.path.exists("/root/reference/input_file/lwfa/qpinput.json"), reason="the reference tree is not on this machine")
def test_reads_the_reference_lwfa_deck():
    cfg, beams = CR.deck_from_json("/root/reference/input_file/lwfa/qpinput.json")
    from qpad_b200 import decks
    want = decks.CONFIGS["C4"]
    assert not beams and all(cfg[k] == want[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "ppc1", "ppc2", "num_theta"))
    assert all(cfg["laser"][k] == want["laser"][k] for k in ("k0", "a0", "w0", "focal_distance", "lon_center", "t_rise", "t_flat", "t_fall", "iteration"))

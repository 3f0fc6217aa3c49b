#!/usr/bin/env python
"""Compare this repository's results with OUTPUT OF THE REFERENCE ITSELF (a QPAD build run by a maintainer).

The reference cannot be built in this image (Fortran 2003 + MPI + HYPRE + HDF5), so the oracle of oracle/ is pinned by analytic
answers only ("parity unpinned").  This tool retires that caveat the day reference output is at hand:

    # on a machine with a QPAD build: run the deck with field dumps enabled (diag "psi_cyl_m", "ez_cyl_m", ...), then either
    python tools/compare_reference.py --ref /path/to/qpad/run --deck input_file/.../qpinput.json --step 1          (needs h5py)
    # or export the dumps to one .npz there (python + h5py, 6 lines, see export_npz below) and bring the file here
    python tools/compare_reference.py --ref dumps_step1.npz --deck C1 --step 1

Layout of the reference's field dumps (diagnostics_class.f03:583-601, :947-1044, hdf5io_class.f03:358-384):
    ./Fields/<Name>/<Part>/<name>_%08d.h5      dataset <name>, rank 2 = (r, xi) of one component of one mode part
    <Name>/<name> in Psi/psi, Er/er, Ephi/ephi, Ez/ez, Br/br, Bphi/bphi, Bz/bz;  <Part> = Re0, Re1, Im1, Re2, Im2, ...
    file number = 3D step at which the dump was written.
Laser decks (nlasers > 0) also dump the envelope (diagnostics_class.f03:662-691, :1005-1019): ./Lasers1/A_laser/<Cplx>_<Part>/a_laser_%08d.h5 with
    <Cplx> = Re | Im (a_r | a_i of the envelope) and <Part> as above -- compared with the envelope volumes of our run (slices 1..nz, nodes 1..nr).
Our side: the same deck through the CPU oracle (--impl oracle, default: runs anywhere) or the B200 library (--impl gpu), single
stage; psi / e / b volumes [plane][slice][node][component] with plane 0 = Re0, 2m-1 = Re m, 2m = Im m and node j <-> r = (j-1) dr.

Beam particles: the reference draws the thermal momenta from the compiler's random_number (math_module.f03:77-102), which cannot be
reproduced here; decks with uth = 0 (or a beam loaded from the reference's own Raw dump: --beam-npz) compare exactly, others to the
statistical level of the momentum spread's influence on the wake (none at the first step: the deposit uses positions only).

Exit status 0 iff every compared dataset agrees within --tol on its on-axis line-out (north star: 1e-6) and --vol-tol on the volume.
"""
import argparse
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FIELDS = {"Psi": ("psi", "psi", 0), "Er": ("er", "e", 0), "Ephi": ("ephi", "e", 1), "Ez": ("ez", "e", 2),
          "Br": ("br", "b", 0), "Bphi": ("bphi", "b", 1), "Bz": ("bz", "b", 2)}


def part_to_plane(part):
    m = re.fullmatch(r"(Re|Im)(\d+)", part)
    if not m:
        raise ValueError(f"unknown mode part {part!r}")
    k = int(m.group(2))
    if m.group(1) == "Re":
        return 0 if k == 0 else 2 * k - 1
    if k == 0:
        raise ValueError("Im0 does not exist")
    return 2 * k


def load_reference(path, step):
    """{(Name, Part): 2-D array} from a QPAD run directory (HDF5, needs h5py) or an .npz export with keys 'Fields/<Name>/<Part>'; envelope dumps of
    a laser deck come back under the names ("A_laser_Re", Part) / ("A_laser_Im", Part) (npz keys 'Lasers1/A_laser/Re_<Part>' ...)"""
    out = {}
    if os.path.isfile(path) and path.endswith(".npz"):
        z = np.load(path)
        for key in z.files:
            m = re.fullmatch(r"Fields/(\w+)/((?:Re|Im)\d+)", key)
            if m and m.group(1) in FIELDS:
                out[(m.group(1), m.group(2))] = np.asarray(z[key], dtype=np.float64)
            m = re.fullmatch(r"Lasers1/A_laser/(Re|Im)_((?:Re|Im)\d+)", key)
            if m:
                out[("A_laser_" + m.group(1), m.group(2))] = np.asarray(z[key], dtype=np.float64)
        return out
    try:
        import h5py
    except ImportError as exc:
        raise SystemExit("reading the reference's HDF5 dumps needs h5py (not in this image): export them to .npz where QPAD ran -- see "
                         "export_npz in this file -- and pass the .npz") from exc
    for name, (dset, _, _) in FIELDS.items():
        base = os.path.join(path, "Fields", name)
        if not os.path.isdir(base):
            continue
        for part in sorted(os.listdir(base)):
            f = os.path.join(base, part, f"{dset}_{step:08d}.h5")
            if os.path.exists(f):
                with h5py.File(f, "r") as h:
                    out[(name, part)] = np.asarray(h[dset], dtype=np.float64)
    base = os.path.join(path, "Lasers1", "A_laser")
    if os.path.isdir(base):
        for sub in sorted(os.listdir(base)):
            m = re.fullmatch(r"(Re|Im)_((?:Re|Im)\d+)", sub)
            f = os.path.join(base, sub, f"a_laser_{step:08d}.h5")
            if m and os.path.exists(f):
                with h5py.File(f, "r") as h:
                    out[("A_laser_" + m.group(1), m.group(2))] = np.asarray(h["a_laser"], dtype=np.float64)
    return out


def export_npz(run_dir, step, out_path):
    """the six lines to run where QPAD ran (python + h5py): all field dumps of one step -> one .npz"""
    np.savez_compressed(out_path, **{f"Fields/{n}/{p}": a for (n, p), a in load_reference(run_dir, step).items()})


def deck_from_json(path):
    """the subset of a qpinput.json this comparison needs (simulation / beam / species blocks of input_file/*/qpinput.json)"""
    from qpad_b200 import decks
    d = decks.load_deck(path)
    sim, box = d["simulation"], d["simulation"]["box"]
    cfg = dict(nr=sim["grid"][0], nz=sim["grid"][1], max_mode=sim["max_mode"], rmax=box["r"][1], zmin=box["z"][0], zmax=box["z"][1], dt=sim["dt"],
               iter_max=sim.get("iter_max", 1), iter_reltol=sim.get("iter_reltol", 1e-3), iter_abstol=sim.get("iter_abstol", 1e-3))
    sp = d["species"][0]
    cfg.update(ppc1=sp["ppc"][0], ppc2=sp["ppc"][1], num_theta=sp["num_theta"])
    if d.get("laser"):
        l0 = d["laser"][0]
        cfg["laser"] = {k: l0[k] for k in ("k0", "a0", "w0", "focal_distance", "lon_center", "t_rise", "t_flat", "t_fall", "iteration") if k in l0}
    beams = []
    for b in d.get("beam", []):
        beams.append(dict(ppc=tuple(b["ppc"]), num_theta=b["num_theta"], q=b["q"], m=b["m"], gamma=b["gamma"], density=b["density"], quiet=b.get("quiet_start", True),
                          center=(b["gauss_center"][0], b["gauss_center"][1], b["gauss_center"][2]), sigma=tuple(b["gauss_sigma"]),
                          range1=tuple(b["range1"]), range2=tuple(b["range2"]), range3=tuple(b["range3"]), uth=tuple(b["uth"]), den_min=b.get("den_min", 1e-10)))
    return cfg, beams


def run_ours_laser(cfg, plasma, nsteps, impl):
    """a laser deck (cfg["laser"]: the keys of the deck's laser block; robust_pgc plasma, no beam): fields + envelope volumes after nsteps"""
    from qpad_b200 import decks
    keys = ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")
    las = cfg["laser"]
    a_r, a_i = decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], max_mode=cfg["max_mode"], **las)
    nz = cfg["nz"]
    if impl == "oracle":
        from oracle import oracle as O
        sim = O.Sim(sp_push_type=5, laser_on=1, laser_iter=las["iteration"], laser_k0=las["k0"], beam_evol=0, **{k: cfg[k] for k in keys + ("ppc1", "ppc2", "num_theta")})
        sim.set_laser(a_r, a_i)
        sim.set_beam(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
        for k in range(nsteps):
            sim.step3d(k + 1)
        out = {n: sim.field(n, 2)[:, :nz] for n in ("psi", "e", "b")}
        out["a_r"], out["a_i"], _ = sim.laser()
        return out
    from qpad_b200 import capi
    sim = capi.Sim(sp_npmax=2 * len(plasma[4]), beam_npmax=64, beam_evol=0, sp_push_pgc=1, laser_iter=las["iteration"], laser_k0=las["k0"], sp_ppc_r=cfg["ppc1"], use_graph=1,
                   **{k: cfg[k] for k in keys})
    sim.init_species(*plasma)
    sim.laser.upload(a_r, a_i)
    for _ in range(nsteps):
        sim.step3d()
    out = {n: sim.field(n).download_f2()[:, :nz] for n in ("psi", "e", "b")}
    out["a_r"], out["a_i"] = sim.laser.download()
    return out


def run_ours(cfg, beam_arrays, plasma, nsteps, impl):
    keys = ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")
    if impl == "oracle":
        from oracle import oracle as O
        sim = O.Sim(**{k: cfg[k] for k in keys + ("ppc1", "ppc2", "num_theta")})
        sim.set_beam(*beam_arrays)
        for k in range(nsteps):
            sim.step3d(k + 1)
        return {n: sim.field(n, 2)[:, :cfg["nz"]] for n in ("psi", "e", "b")}
    from qpad_b200 import capi
    sim = capi.Sim(sp_npmax=2 * len(plasma[4]), beam_npmax=len(beam_arrays[2]) + 1024, use_graph=1, **{k: cfg[k] for k in keys})
    sim.init_species(*plasma)
    sim.beam.upload(*beam_arrays)
    for _ in range(nsteps):
        sim.step3d()
    return {n: sim.field(n).download_f2()[:, :cfg["nz"]] for n in ("psi", "e", "b")}


def compare(ref, ours, cfg, tol, vol_tol, out=sys.stdout):
    nr, nz = cfg["nr"], cfg["nz"]
    ok, rows = True, []
    for (name, part), a in sorted(ref.items()):
        pl = part_to_plane(part)
        if name.startswith("A_laser_"):                         # envelope volume (P, nz+3, nr+2): slice j at index j+1
            vol = ours.get("a_r" if name.endswith("Re") else "a_i")
            if vol is None or pl >= vol.shape[0]:
                continue
            mine = vol[pl, 2:nz + 2, 1:nr + 1]
        else:
            dset, fld, comp = FIELDS[name]
            if pl >= ours[fld].shape[0]:
                continue
            mine = ours[fld][pl, :, 1:nr + 1, comp]             # (xi, r)
        a = np.squeeze(a)
        if a.shape == (nr, nz):
            a = a.T
        if a.shape != (nz, nr):
            rows.append((name, part, "shape", a.shape)); ok = False
            continue
        scale = max(np.max(np.abs(a)), 1e-300)
        vol = float(np.max(np.abs(mine - a)) / scale)
        # the axis node of the higher mode parts is zero by the axis rules: the line-out is taken where the field lives
        jl = 0 if np.max(np.abs(a[:, 0])) > 1e-3 * scale else int(np.argmax(np.max(np.abs(a), axis=0)))
        line = float(np.max(np.abs(mine[:, jl] - a[:, jl])) / max(np.max(np.abs(a[:, jl])), 1e-300))
        good = line <= tol and vol <= vol_tol
        ok &= good
        rows.append((name, part, f"line-out(r index {jl}) {line:.3e}", f"volume {vol:.3e}", "ok" if good else "MISMATCH"))
    for r in rows:
        print("  ".join(str(x) for x in r), file=out)
    if not rows:
        print("no comparable datasets found", file=out)
        return False
    return ok


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref", required=True, help="QPAD run directory (HDF5 dumps; needs h5py) or an .npz export")
    ap.add_argument("--deck", required=True, help="C1 / C3 (decks.CONFIGS) or the path of the qpinput.json the reference ran")
    ap.add_argument("--step", type=int, default=1, help="number of the dump = 3D steps to run")
    ap.add_argument("--impl", default="oracle", choices=["oracle", "gpu"])
    ap.add_argument("--beam-npz", default=None, help="beam particles dumped by the reference (arrays x, p, q) instead of the PCG64 substitute")
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--vol-tol", type=float, default=1e-5)
    ap.add_argument("--export-npz", default=None, help="only convert the HDF5 dumps of --ref / --step to this .npz and exit")
    args = ap.parse_args(argv)
    if args.export_npz:
        export_npz(args.ref, args.step, args.export_npz)
        return 0
    import bench
    from qpad_b200 import decks
    if os.path.exists(args.deck):
        cfg, beams = deck_from_json(args.deck)
    else:
        cfg, beams = bench.deck_config(args.deck)
        beams = beams if isinstance(beams, list) else [beams]
    plasma = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
    if cfg.get("laser"):                                        # laser deck: no beam; fields + envelope
        ref = load_reference(args.ref, args.step)
        ok = compare(ref, run_ours_laser(cfg, plasma, args.step, args.impl), cfg, args.tol, args.vol_tol)
        print(json.dumps({"datasets": len(ref), "ok": bool(ok), "tol": args.tol, "vol_tol": args.vol_tol, "impl": args.impl}))
        return 0 if ok else 1
    if args.beam_npz:
        z = np.load(args.beam_npz)
        bm = (np.ascontiguousarray(z["x"]), np.ascontiguousarray(z["p"]), np.ascontiguousarray(z["q"]))
    else:
        parts = [decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(b, seed=10 + k)) for k, b in enumerate(beams)]
        bm = tuple(np.concatenate([p[a] for p in parts]) for a in range(3))
    ref = load_reference(args.ref, args.step)
    ours = run_ours(cfg, bm, plasma, args.step, args.impl)
    ok = compare(ref, ours, cfg, args.tol, args.vol_tol)
    print(json.dumps({"datasets": len(ref), "ok": bool(ok), "tol": args.tol, "vol_tol": args.vol_tol, "impl": args.impl}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

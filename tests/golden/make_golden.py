"""Regenerates the fixtures of tests/golden/ from the CPU oracle (oracle/, parity build -ffp-contract=off).

These are NOT outputs of the reference: QPAD (Fortran 2003 + MPI + HYPRE + HDF5) cannot be built in this image and ships
no golden vectors.  The fixtures freeze the oracle -- which is pinned by the analytic known answers of
tests/test_oracle_known_answers.py and tests/test_oracle_laser.py -- so that (a) a change of the oracle shows up as a diff
here, and (b) the CUDA path can be checked against committed numbers.  If reference HDF5 output for these decks becomes
available it replaces these files.

    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import oracle as O  # noqa: E402
from qpad_b200 import decks  # noqa: E402

BLOWOUT = dict(nr=64, nz=24, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=3)
BLOWOUT_PLASMA = dict(ppc1=2, ppc2=2, num_theta=8)
LWFA = dict(nr=128, nz=96, max_mode=0, rmax=12.0, zmin=-3.0, zmax=6.0, dt=2.0, iter_max=6, iter_reltol=1e-3, iter_abstol=1e-6)
LWFA_PLASMA = dict(ppc1=4, ppc2=2, num_theta=8)
LWFA_LASER = dict(k0=20.0, a0=1.2, w0=2.5, focal_distance=0.0, lon_center=0.0, t_rise=1.5, t_flat=0.0, t_fall=1.5, iteration=3)


def blowout():
    """the smoke deck: 12 slices of a small beam-driven blowout (max_mode 1), then one full 3D step with beam push"""
    beam = dict(decks.CONFIGS["C1"]["beam"])
    bm = decks.beam_std(BLOWOUT["nr"], BLOWOUT["nz"], BLOWOUT["rmax"], BLOWOUT["zmin"], BLOWOUT["zmax"], **beam)
    sim = O.Sim(**BLOWOUT_PLASMA, **BLOWOUT)
    sim.set_beam(*bm)
    sim.run_slices(12)
    out = dict(psi12=sim.field("psi", 2)[:, :12], e12=sim.field("e", 2)[:, :12], b12=sim.field("b", 2)[:, :12], iters12=sim.total_iters())
    sim2 = O.Sim(**BLOWOUT_PLASMA, **BLOWOUT)
    sim2.set_beam(*bm)
    sim2.step3d(1)
    bx, bp, bq = sim2.beam()
    w = bq / bq.sum()
    out.update(ez_axis=sim2.field("e", 2)[0, :BLOWOUT["nz"], 1, 2], psi_axis=sim2.field("psi", 2)[0, :BLOWOUT["nz"], 1, 0],
               beam_n=len(bq), beam_mean_x=float((w * bx[:, 0]).sum()), beam_mean_xi=float((w * bx[:, 2]).sum()),
               beam_rms_px=float(np.sqrt((w * bp[:, 0] ** 2).sum())))
    return out


def lwfa():
    """config 4 in small: two 3D steps of a laser-driven wake with envelope advance"""
    las = dict(LWFA_LASER)
    it = las.pop("iteration")
    sim = O.Sim(sp_push_type=5, laser_on=1, laser_iter=it, laser_k0=las["k0"], beam_evol=0, **LWFA_PLASMA, **LWFA)
    sim.set_laser(*decks.laser_gaussian(LWFA["nr"], LWFA["nz"], LWFA["rmax"], LWFA["zmin"], LWFA["zmax"], **las))
    sim.set_beam(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    for k in range(2):
        sim.step3d(k + 1)
    ar, ai, chi = sim.laser()
    nz = LWFA["nz"]
    psi, e = sim.field("psi", 2), sim.field("e", 2)
    return dict(psi_axis=psi[0, :nz, 1, 0], ez_axis=e[0, :nz, 1, 2], psi_slices=psi[0, [40, 60, 80], :, 0], er_slices=e[0, [40, 60, 80], :, 0],
                a_axis_r=ar[0, 2:nz + 2, 1], a_axis_i=ai[0, 2:nz + 2, 1], a_slice60_r=ar[0, 61, :], chi_slice60=chi[0, 59, :], iters=sim.total_iters())


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "blowout_small.npz"), **blowout())
    np.savez_compressed(os.path.join(HERE, "lwfa_small.npz"), **lwfa())
    for f in ("blowout_small.npz", "lwfa_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")

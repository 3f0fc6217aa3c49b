"""development aid: where does the C3 (hosing deck, 2 steps) GPU result differ from the oracle? per-slice max error of e / psi"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fullsize_cases as F
from qpad_b200 import capi
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg, plasma, bm, _, _ = F.deck("C3")
sim = capi.Sim(sp_npmax=2 * len(plasma[4]), beam_npmax=len(bm[2]) + 1024, use_graph=1, **{k: cfg[k] for k in F.KEYS})
sim.init_species(*plasma); sim.beam.upload(*bm)
from oracle import oracle as O
orc = O.Sim(**{k: cfg[k] for k in F.KEYS + ("ppc1", "ppc2", "num_theta")}); orc.set_beam(*bm)
for k in range(nsteps):
    sim.step3d(); orc.step3d(k + 1)
    for name in ("psi", "e", "b", "q_beam"):
        g = sim.field(name).download_f2()[:, :cfg["nz"]]; w = orc.field(name, 2)[:, :cfg["nz"]]
        err = np.max(np.abs(g - w), axis=(0, 2, 3)) / np.max(np.abs(w))
        top = np.argsort(err)[-5:][::-1]
        print(f"step {k + 1} {name}: max rel err {err.max():.3e} at slices {top.tolist()} ({[f'{err[t]:.1e}' for t in top]}); median {np.median(err):.1e}; first slice > 1e-9: {int(np.argmax(err > 1e-9)) if (err > 1e-9).any() else None}")
    gx, gp, gq = sim.beam.download(); ox, op, oq = orc.beam()
    print(f"step {k + 1} beam: n {len(gq)} {len(oq)} x err {np.max(np.abs(gx - ox)):.2e} p err {np.max(np.abs(gp - op)) / np.max(np.abs(op)):.2e}")
print("iters", sim.stats(), orc.total_iters())

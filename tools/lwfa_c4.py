"""config 4 (input_file/lwfa) on one B200: the robust_pgc slice loop + envelope advance through the C-ABI, K 3D steps timed
with CUDA events (the deck runs 5).  python tools/lwfa_c4.py [steps] -> one JSON line (also usable under ncu)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from qpad_b200 import capi, decks  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = dict(decks.CONFIGS["C4"])
las = cfg.pop("laser")
cfg.pop("nstep3d")
ppc1, ppc2, nth = cfg.pop("ppc1"), cfg.pop("ppc2"), cfg.pop("num_theta")
x, p, g, psi, q = decks.plasma_uniform(cfg["nr"], cfg["rmax"], ppc1, ppc2, nth)
stream = torch.cuda.Stream()
sim = capi.Sim(sp_npmax=2 * len(q), beam_npmax=64, beam_evol=0, sp_push_pgc=1, laser_iter=las["iteration"], laser_k0=las["k0"], sp_ppc_r=ppc1,
               stream=stream.cuda_stream, **cfg)
sim.init_species(x, p, g, psi, q)
sim.laser.upload(*decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las))
sim.step3d()                                             # warm-up step (also step 1 of the deck)
torch.cuda.synchronize()
u0, i0, s0 = sim.stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps - 1):
    sim.step3d()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
u1, i1, s1 = sim.stats()
ar, ai = sim.laser.download()
psi2 = sim.field("psi").download_f2()
print(json.dumps({"workload": "C4 lwfa: nr=512 nz=512 max_mode=0 Np/slice=%d robust_pgc, laser a0=2 k0=20 iteration 3" % len(q),
                  "steps_timed": steps - 1, "ms_per_step": ms / max(steps - 1, 1), "updates_per_s": (u1 - u0) / (ms * 1e-3),
                  "us_per_slice": ms * 1e3 / max(s1 - s0, 1), "pc_iters_per_slice": (i1 - i0) / max(s1 - s0, 1),
                  "max_abs_a": float(np.hypot(ar, ai).max()), "psi_min": float(psi2.min()), "psi_max": float(psi2.max()),
                  "path": "per-slice launches (field programs A/C/D + pgc particle kernels) + one persistent envelope-solve CTA per step"}))

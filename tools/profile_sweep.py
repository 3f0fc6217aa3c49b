"""Run a short stretch of the C2 slab sweep (for ncu): python tools/profile_sweep.py [first_slice] [nslices] [config]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qpad_b200.pipeline import SingleStage  # noqa: E402

j0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1100
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg, beam = bench.deck_config(sys.argv[3] if len(sys.argv) > 3 else "C2")
plasma, bm = bench.make_inputs(cfg, beam)
r = SingleStage(cfg, plasma, bm)
s = r.sim
if os.environ.get("QPG_SWEEP_CTAS"):
    s.set_sweep_ctas(int(os.environ["QPG_SWEEP_CTAS"]))
r.prepare_step()
if j0 > 1:
    s.run_slices(1, j0 - 1)
s.ctx.sync()
s.sweep_profile(reset=True)
s.run_slices(j0, j0 + n - 1)      # <- the launch to profile: the LAST k_sweep launch of the process
s.ctx.sync()
p = s.sweep_profile()
nspc = p["ns_total"] / p["cyc_total"]
print({k: round(v * nspc * 1e-3 / n, 2) for k, v in p.items() if k.startswith(("cyc_", "work_"))}, "us/slice; iters", p["amj_phases"] / n)

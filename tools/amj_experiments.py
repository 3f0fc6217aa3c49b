"""development aid: time the stand-alone particle kernels on a set streamed from HBM (4M particles)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qpad_b200.pipeline import SingleStage
cfg, beam = bench.deck_config("C2")
plasma, bm = bench.make_inputs(cfg, beam)
r = SingleStage(cfg, plasma, bm)
r.prepare_step(); r.sim.run_slices(1, 600)
out = r.kernel_microbench(6556.5)
n = out["particles"]
print({k: round(v, 1) for k, v in out.items() if k.endswith("GBs")}, "amj us/launch %.1f" % (64.0 * n / out["amjdeposit_GBs"] / 1e3), "push pair us %.1f" % (136.0 * n / out["push_u_plus_push_x_GBs"] / 1e3))

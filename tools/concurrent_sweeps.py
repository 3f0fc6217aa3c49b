"""development aid: do S cooperative sweep kernels with 148/S CTAs each overlap on one GPU?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from qpad_b200.pipeline import SingleStage
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nsl = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg, beam = bench.deck_config("C2")
plasma, bm = bench.make_inputs(cfg, beam)
stages = []
for k in range(S):
    st = torch.cuda.Stream()
    r = SingleStage(cfg, plasma, bm, stream=st)
    r.sim.set_sweep_ctas(148 // S)
    r.prepare_step()
    r.sim.run_slices(1, 1000)
    stages.append((st, r))
torch.cuda.synchronize()
for mode in ("serial", "concurrent"):
    t0 = time.perf_counter()
    for st, r in stages:
        r.sim.run_slices(1001, 1000 + nsl)
        if mode == "serial":
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"S={S} {mode}: {dt*1e3:.2f} ms for {S}x{nsl} slices -> {dt*1e6/(S*nsl):.2f} us per slice effective")

"""Timing probe: slice-loop wall time per slice for graph / stream modes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from qpad_b200 import capi, decks
from qpad_b200.pipeline import SingleStage

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
nsl = int(sys.argv[2]) if len(sys.argv) > 2 else 512
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["graph", "stream", "stream1"]
cfg = dict(decks.CONFIGS[name]); beam = cfg.pop("beam")
pl = decks.plasma_uniform(cfg["nr"], cfg["rmax"], cfg["ppc1"], cfg["ppc2"], cfg["num_theta"])
bm = decks.beam_std(min(cfg["nr"], 256), min(cfg["nz"], 512), cfg["rmax"], cfg["zmin"], cfg["zmax"], **beam)
for mode in modes:
    c = dict(cfg)
    if mode == "stream1":
        c["iter_max"] = 1
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        r = SingleStage(c, pl, bm, stream=st, use_graph=1 if mode == "graph" else 0)
        s = r.sim
        r.prepare_step()
        j0 = int(0.5 * cfg["nz"])
        s.run_slices(1, 8)
        torch.cuda.synchronize()
        i0 = s.stats()[1]
        t0 = time.perf_counter()
        s.run_slices(9, 8 + nsl)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        i1 = s.stats()[1]
        print(f"{name} {mode}: {1e6*(t2-t0)/nsl:.1f} us/slice (enqueue {1e6*(t1-t0)/nsl:.1f} us/slice), pc iters/slice {(i1-i0)/nsl:.2f}, npp {s.species.npp()}", flush=True)
        r.close()

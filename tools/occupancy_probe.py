"""development aid: throughput of the stand-alone amjdeposit kernel against resident blocks per SM (QPG_DEV_EXTRA_SMEM) and particle-set size
(4M particles = streamed from HBM, 262144 = L2 resident like a slice of C2): python tools/occupancy_probe.py"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import bench
    from qpad_b200.pipeline import SingleStage, kernel_microbench
    cfg, beam = bench.deck_config("C2")
    plasma, bm = bench.make_inputs(cfg, beam, xi_cells=(0, 8))
    r = SingleStage(cfg, plasma, bm)
    r.prepare_step(); r.sim.run_slices(1, 4)
    for n in (4 * 1024 * 1024, 262144):
        out = kernel_microbench(r.sim, cfg, 6556.5, n_big=n, reps=8)
        npart = out["particles"]
        us = 64.0 * npart / out["amjdeposit_GBs"] / 1e3
        print(json.dumps({"n": npart, "amj_us": round(us, 2), "tiles_per_us_per_sm": round(npart / 32 / us / 148, 2), "sm_cycles_per_tile": round(us * 1965 * 148 / (npart / 32), 1)}))
else:
    for blocks, extra in ((4, 0), (3, 44 * 1024), (2, 80 * 1024), (1, 160 * 1024)):
        env = dict(os.environ, QPG_DEV_EXTRA_SMEM=str(extra))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(f"blocks/SM <= {blocks} (8 warps each):", r.stdout.strip().replace("\n", "  |  "), r.stderr[-300:] if r.returncode else "", flush=True)

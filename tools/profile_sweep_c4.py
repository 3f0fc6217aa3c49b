"""Run a stretch of the C4 (laser wakefield) slab sweep for ncu: python tools/profile_sweep_c4.py [first_slice] [nslices]
(one xi stage; QPG_SWEEP_CTAS = CTAs of the sweep kernel, e.g. 28 = a stage of the 5-stage pipeline; set QPG_LASER_NO_OVERLAP=1 under a
profiler: it serialises kernels, so a sweep that follows the progress word of a concurrently running envelope solve would never end)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from qpad_b200 import capi, decks  # noqa: E402

j0 = int(sys.argv[1]) if len(sys.argv) > 1 else 129
n = int(sys.argv[2]) if len(sys.argv) > 2 else 102
cfg, _ = bench.deck_config("C4")
las = cfg["laser"]
plasma, _bm = bench.make_inputs(cfg, None)
npp0 = len(plasma[4])
stream = torch.cuda.Stream()
kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
sim = capi.Sim(sp_npmax=2 * npp0, beam_npmax=64, beam_evol=0, sp_push_pgc=1, laser_iter=las["iteration"], laser_k0=las["k0"], sp_ppc_r=cfg["ppc1"], use_graph=1,
               stream=stream.cuda_stream, **kw)
if os.environ.get("QPG_SWEEP_CTAS"):
    sim.set_sweep_ctas(int(os.environ["QPG_SWEEP_CTAS"]))
sim.init_species(*plasma)
sim.laser.upload(*decks.laser_gaussian(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **las))
sim.step3d()                      # one whole step first: chi volume and an advanced envelope
sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
if j0 > 1:
    sim.run_slices(1, j0 - 1)
sim.ctx.sync()
sim.sweep_profile(reset=True)
sim.run_slices(j0, j0 + n - 1)    # <- the launch to profile: the LAST k_sweep launch of the process
sim.ctx.sync()
p = sim.sweep_profile()
nspc = p["ns_total"] / p["cyc_total"]
print({k: round(v * nspc * 1e-3 / n, 2) for k, v in p.items() if k.startswith(("cyc_", "work_"))}, "us/slice; iters", p["amj_phases"] / n)

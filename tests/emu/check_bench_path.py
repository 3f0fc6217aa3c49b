"""TEST TOOL (host emulation, tests/emu): the default bench driver (probe_partition -> LocalPipeline with 4 SM-partitioned stages, cost-balanced slabs) on the corrected
beam inputs, at reduced size, in emulation, against the oracle's single-stage run"""
import os, sys, time
os.environ["QPAD_EMU_SWEEP"] = "1"; os.environ["QPAD_EMU_SMS"] = "48"
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from emu import emu, faketorch
sys.modules["torch"] = faketorch
import bench
from oracle import oracle as O
cfg, beam = bench.deck_config("C2")
cfg = dict(cfg, nr=320, nz=128, ppc1=2, ppc2=2, num_theta=8, iter_max=4)
plasma, bm = bench.make_inputs(cfg, beam)
print("plasma", len(plasma[4]), "beam", len(bm[2]), flush=True)
with emu.patched() as capi:
    from qpad_b200.pipeline import LocalPipeline, probe_partition
    S = 4
    t = time.time(); parts = probe_partition(cfg, plasma, bm, S, S); print("partition", parts, round(time.time() - t, 1), "s", flush=True)
    lp = LocalPipeline(cfg, plasma, bm, S, partition=parts)
    nwaves = 3
    t = time.time()
    lp.fill()
    for _ in range(nwaves): lp.wave()
    lp.drain()
    print("pipeline", round(time.time() - t, 1), "s; coop launches", emu.lib().emu_coop_launches(), flush=True)
    upd, iters, slices = lp.stats()
    nsteps = slices // cfg["nz"]
    kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol", "ppc1", "ppc2", "num_theta")}
    orc = O.Sim(nstages=1, **kw); orc.set_beam(*bm)
    for k in range(nsteps): orc.step3d(k + 1)
    print("steps", nsteps, "iters", iters, orc.total_iters(), "iters/slice", iters / slices)
    for (noff, nzp), sim in zip(parts, lp.sims):
        for name in ("psi", "e", "b"):
            got, want = sim.field(name).download_f2()[:, :nzp], orc.field(name, 2, stage=0)[:, noff:noff + nzp]
            print(noff, nzp, name, "rel err", np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300))
    q = np.concatenate([s.beam.download()[2] for s in lp.sims]); oq = orc.beam(stage=0)[2]
    print("beam particles", len(q), len(oq), np.array_equal(np.sort(q), np.sort(oq)))
    lp.close()

import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import bench
from emu import emu
from oracle import oracle as O
cfg, beam = bench.deck_config("C2")
cfg = dict(cfg, nr=512, nz=640, ppc1=2, ppc2=2, num_theta=8)
pl, bm = bench.make_inputs(cfg, beam)
print("plasma", len(pl[4]), "beam", len(bm[2]), flush=True)
nsl = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol", "ppc1", "ppc2", "num_theta")}
orc = O.Sim(**kw); orc.set_beam(*bm)
t = time.time(); upd = orc.run_slices(nsl); print("oracle", time.time() - t, "s, iters/slice", orc.total_iters() / nsl, flush=True)
with emu.patched() as capi:
    simkw = {k: cfg[k] for k in ("nr", "nz", "max_mode", "rmax", "zmin", "zmax", "dt", "iter_max", "iter_reltol", "iter_abstol")}
    sim = capi.Sim(sp_npmax=2 * len(pl[4]), beam_npmax=len(bm[2]) + 64, **simkw)
    sim.set_sweep(1)
    sim.init_species(*pl); sim.beam.upload(*bm)
    sim.beam_qdp_begin(); sim.beam_qdp_end(); sim.begin_step()
    t = time.time(); sim.run_slices(1, nsl); u, it, sl = sim.stats()
    print("emulated sweep", time.time() - t, "s; updates", u, upd, "iters", it, orc.total_iters(), "coop launches", emu.lib().emu_coop_launches(), flush=True)
    for name in ("psi", "e", "b", "cu"):
        got, want = sim.field(name).download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
        print(name, "max|want|", np.max(np.abs(want)), "rel err", np.max(np.abs(got - want)) / np.max(np.abs(want)), flush=True)
    gx, gp, gg, gpsi, gq = sim.species.download(); ox, op, og, opsi, oq = orc.plasma()
    print("npp", len(gq), len(oq), "max |dp|", np.max(np.abs(gp - op)) if len(gq) == len(oq) else None, "max |p|", np.max(np.abs(op)))

"""The call sequence of qpad_b200.subcyc.SubcycStage (the sub-cycling loop driven through the per-routine C-ABI) checked on the
CPU the same way as the ionisation loop (tests/test_ionization_host_loop.py): the C-ABI objects are replaced by adapters that
run the ORACLE's per-routine functions; the loop driven through them must reproduce the oracle's own integrated sub-cycling
loop (oracle slice_step_subcyc, proj_subcyc/simulation_subcyc_class.f03:216-376).  qpg_subcyc_step -- host arithmetic of the
real library, callable without a GPU -- is used as is."""
import types

import numpy as np

from oracle import oracle as O
from qpad_b200 import capi as real_capi
from qpad_b200 import decks, subcyc
from test_ionization_host_loop import L, OCtx, OField, OPart2d, OPart3d


class OPart2dS(OPart2d):
    def upload(self, x, p, gamma, psi, q):
        n = self.n = len(q)
        self.x[:n], self.p[:n], self.gamma[:n], self.psi[:n], self.q[:n] = x, p, gamma, psi, q

    def exp_fac_max(self): return L.orc_exp_fac_max(self.p, self.gamma, self.n)
    def clamp_exp_fac(self, clamp): L.orc_clamp_exp_fac(self.p, self.gamma, self.n, clamp)


class OFieldS(OField):
    def copy_to(self, dst): dst.f1[:] = self.f1


def test_subcyc_step_of_the_library_matches_the_oracle():
    import ctypes as C
    for ef, efm, dt, dtmin in ((1.0, 2.0, 0.02, 0.001), (7.3, 2.0, 0.02, 0.001), (900.0, 2.0, 0.02, 0.001), (2.0, 2.0, 0.02, 0.001)):
        dts, ns = C.c_double(), C.c_int()
        L.orc_subcyc_step(ef, efm, dt, dtmin, C.byref(dts), C.byref(ns))
        assert real_capi.subcyc_step(ef, efm, dt, dtmin) == (dts.value, ns.value)


def test_subcyc_call_sequence_reproduces_the_oracle_loop(monkeypatch):
    fake = types.SimpleNamespace(Ctx=OCtx, Field=OFieldS, Part2d=OPart2dS, Part3d=OPart3d, subcyc_step=real_capi.subcyc_step,
                                 COPY_1TO2=real_capi.COPY_1TO2, COPY_2TO1=real_capi.COPY_2TO1, CONV_RECORD=real_capi.CONV_RECORD,
                                 CONV_COMPARE=real_capi.CONV_COMPARE)
    monkeypatch.setattr(subcyc, "capi", fake)
    cfg = dict(nr=64, nz=32, max_mode=1, rmax=5.0, zmin=-5.0, zmax=5.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C1"]["beam"]))
    nsl = 24
    for efm, clamp, dtmin in ((1.1, 50.0, 1e-3), (1.05, 1.6, 1e-3), (1e9, 1e9, 1e-6)):
        orc = O.Sim(ppc1=2, ppc2=2, num_theta=8, subcyc_on=1, subcyc_exp_fac_max=efm, subcyc_exp_fac_clamped=clamp, subcyc_dt_min=dtmin, **cfg)
        orc.set_beam(*bm)
        upd = orc.run_slices(nsl)
        lattice = O.inject_uniform(cfg["nr"], cfg["rmax"] / cfg["nr"], 2, 2, 8)
        st = subcyc.SubcycStage(dict(cfg, exp_fac_max=efm, exp_fac_clamped=clamp, dt_min=dtmin), lattice, bm)
        st.step3d(nslices=nsl)
        assert st.subcycles == orc.total_subcycles() and st.iters == orc.total_iters() and st.updates == upd
        if efm < 2:
            assert st.subcycles > nsl
        for name, f in (("psi", st.psi), ("e", st.e), ("b", st.b), ("cu", st.cu)):
            got, want = f.download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
            assert np.max(np.abs(want)) > 1e-3 and np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), name
        x, p, g, psi, q = orc.plasma()
        assert st.part.n == len(q) and np.array_equal(st.part.p[:st.part.n], p)

"""The call sequence of qpad_b200.ionization.IonizationStage (the ionisation deck driven through the per-routine C-ABI, as the
Fortran host would) checked on CPU: the C-ABI objects are replaced by adapters that run the ORACLE's per-routine functions
(same entry-point names and semantics: deposits add into the field and apply the axis rules to the sum, `a.add_to(b)` is
b += a, ...).  The loop driven through those adapters must reproduce the oracle's own integrated loop -- which pins the ORDER
of the calls in ionization.py (simulation_class.f03:294-512 with nneutrals = 1) without a GPU.  The device kernels themselves
are covered by tests/test_gpu_neutral.py."""
import ctypes as C
import types

import numpy as np

from oracle import oracle as O
from qpad_b200 import capi as real_capi
from qpad_b200 import ionization

L = O.lib()


class OCtx:
    def __init__(self, nr, max_mode, dr, dxi, field_boundary=O.BND_OPEN, relax_fac=-1.0, device=0, **kw):
        self.nr, self.max_mode, self.P, self.dr, self.dxi, self.bnd = nr, max_mode, 2 * max_mode + 1, dr, dxi, field_boundary
        self.relax = relax_fac if relax_fac >= 0 else 1.0e-3 * ((dr / 0.02) * (dr / 0.02))        # sim_fields_class.f03:137
        self.conv = None

    def close(self): pass
    def solve_psi(self, q, psi): L.orc_solve_psi(q.f1, psi.f1, self.nr, self.max_mode, self.dr, self.bnd)
    def solve_bt(self, qb, b): L.orc_solve_bt(qb.f1, b.f1, self.nr, self.max_mode, self.dr, self.bnd)
    def solve_bz(self, cu, b): L.orc_solve_bz(cu.f1, b.f1, self.nr, self.max_mode, self.dr, self.bnd)
    def solve_bt_iter(self, dcu, cu, b): L.orc_solve_bt_iter(dcu.f1, cu.f1, b.f1, self.nr, self.max_mode, self.dr, self.bnd, self.relax)
    def solve_ez(self, cu, e): L.orc_solve_ez(cu.f1, e.f1, self.nr, self.max_mode, self.dr, self.bnd)
    def solve_et(self, b, psi, e): L.orc_solve_et(b.f1, psi.f1, e.f1, self.nr, self.max_mode, self.dr)
    def solve_djdxi(self, acu, amu, dcu): L.orc_solve_djdxi(acu.f1, amu.f1, dcu.f1, self.nr, self.max_mode, self.dr)

    def convergence_tester(self, fld, dim, op):
        """simulation_class.f03:522-606"""
        f = np.abs(fld.f1[:, 1:self.nr + 1, dim - 1])
        re = f[0] + sum(f[2 * m - 1] for m in range(1, self.max_mode + 1))
        im = sum((f[2 * m] for m in range(1, self.max_mode + 1)), np.zeros(self.nr))
        if op == real_capi.CONV_RECORD:
            self.conv = (re, im)
            return 0.0, 0.0
        ore, oim = self.conv
        old = np.sqrt(np.max(ore ** 2 + oim ** 2))
        ab = np.sqrt(np.max((ore - re) ** 2 + (oim - im) ** 2))
        return (ab / old if old > np.finfo(float).eps else np.finfo(float).max), ab


class OField:
    def __init__(self, ctx, dim, nzp=0, has_2d=False):
        self.ctx, self.dim, self.nzp = ctx, dim, nzp
        self.f1 = np.zeros((ctx.P, ctx.nr + 2, dim))
        self.f2 = np.zeros((ctx.P, nzp + 1, ctx.nr + 2, dim)) if has_2d else None

    def fill(self, v=0.0): self.f1[:] = v
    def fill_f2(self, v=0.0): self.f2[:] = v
    def add_to(self, b): b.f1 += self.f1
    def add_f2_to(self, b): b.f2 += self.f2
    def scale(self, s): self.f1 *= s
    def download_f2(self): return self.f2.copy()

    def copy_slice(self, idx, direction):
        if direction == real_capi.COPY_1TO2:
            self.f2[:, idx - 1] = self.f1
        else:
            self.f1[:] = self.f2[:, idx - 1]

    def add_dim_to(self, b, adim, bdim):
        for a, d in zip(adim, bdim):
            b.f1[..., d - 1] += self.f1[..., a - 1]

    @staticmethod
    def add3(a1, a2, a3): a3.f1[:] = a1.f1 + a2.f1


class OPart2d:
    def __init__(self, ctx, qbm, npmax):
        self.ctx, self.qbm, self.n = ctx, qbm, 0
        self.x, self.p = np.zeros((npmax, 2)), np.zeros((npmax, 3))
        self.gamma, self.psi, self.q = np.zeros(npmax), np.zeros(npmax), np.zeros(npmax)

    def _a(self): c = self.ctx; return c.dr, c.nr, c.max_mode
    def npp(self): return self.n
    def clear(self): self.n = 0
    def close(self): pass
    def qdeposit(self, qf): dr, nr, M = self._a(); L.orc_qdeposit(self.x, self.q, self.n, dr, nr, M, qf.f1)

    def amjdeposit_robust(self, e, b, cu, amu, dcu, dt):
        dr, nr, M = self._a()
        L.orc_amjdeposit_robust(self.x, self.p, self.q, self.gamma, self.psi, self.n, dr, nr, M, self.qbm, dt, e.f1, b.f1, cu.f1, dcu.f1, amu.f1)

    def push_u_robust(self, e, b, dt): dr, nr, M = self._a(); L.orc_push_u_robust(self.x, self.p, self.gamma, self.n, dr, nr, M, self.qbm, dt, e.f1, b.f1)
    def push_x(self, dt): L.orc_push_x(self.x, self.p, self.gamma, self.n, dt)
    def update_bound(self): self.n = L.orc_update_bound(self.x, self.p, self.gamma, self.psi, self.q, self.n, self.ctx.nr * self.ctx.dr)


class OPart3d:
    def __init__(self, ctx, qbm, dt, npmax, nz_total, noff2, nzp):
        self.ctx, self.qbm, self.dt, self.nz, self.noff2, self.nzp = ctx, qbm, dt, nz_total, noff2, nzp

    def upload(self, x, p, q): self.x, self.p, self.q, self.n = np.ascontiguousarray(x).copy(), np.ascontiguousarray(p).copy(), np.ascontiguousarray(q).copy(), len(q)

    def qdeposit(self, qf):
        c = self.ctx
        L.orc_qdeposit3d(self.x, self.q, self.n, c.dr, c.dxi, c.nr, self.nzp, self.noff2, c.max_mode, qf.f2)

    def push(self, push_type, e, b):
        c = self.ctx
        L.orc_push3d(self.x, self.p, self.n, c.dr, c.dxi, c.nr, self.nzp, self.noff2, c.max_mode, self.qbm, self.dt, push_type, e.f2, b.f2)

    def update_bound(self): self.n = L.orc_update_bound3d(self.x, self.p, self.q, self.n, self.ctx.nr * self.ctx.dr, self.nz * self.ctx.dxi)


class ONeutral:
    def __init__(self, ctx, element, ion_max, ppc, num_theta, q=-1.0, m=1.0, density=1.0, n0=1.0e17, dt_xi=None):
        self.ctx, self.ppc, self.nth, self.qm, self.density, self.dt = ctx, ppc, num_theta, q / m, density, dt_xi
        self.adk = np.zeros(60)
        self.multi_max = L.orc_adk_params(element, ion_max, self.adk)
        self.wp = L.orc_plasma_frequency(n0)
        self.lev = np.zeros((self.multi_max + 2, num_theta, ctx.nr))
        L.orc_neutral_reset(self.lev, ctx.nr, num_theta, self.multi_max)
        cap = ctx.nr * num_theta * ppc[0] * ppc[1] + 64
        self.part, self.part_add = OPart2d(ctx, q / m, cap), OPart2d(ctx, q / m, cap)

    def update(self, e):
        c, mm = self.ctx, self.multi_max
        old = self.lev[mm + 1].copy()
        L.orc_neutral_ionize(self.lev, self.adk, e.f1, self.wp, self.dt, self.ppc[0], self.ppc[1], c.nr, self.nth, c.max_mode, mm)
        npp = C.c_long(self.part.n)
        pt, pa = self.part, self.part_add
        nadd = L.orc_neutral_add_particles(self.lev, old, c.nr, self.nth, mm, self.ppc[0], self.ppc[1], c.dr, self.qm, self.density, 1e-10,
                                           pt.x, pt.p, pt.gamma, pt.psi, pt.q, C.byref(npp), pa.x, pa.q)
        pt.n, pa.n = npp.value, nadd

    def renew(self):
        L.orc_neutral_reset(self.lev, self.ctx.nr, self.nth, self.multi_max)
        self.part.clear(); self.part_add.clear()

    def levels(self): return self.lev.copy()
    def close(self): pass


def test_ionization_call_sequence_reproduces_the_oracle_loop(monkeypatch):
    from qpad_b200 import decks
    fake = types.SimpleNamespace(Ctx=OCtx, Field=OField, Part3d=OPart3d, Neutral=ONeutral, COPY_1TO2=real_capi.COPY_1TO2, COPY_2TO1=real_capi.COPY_2TO1,
                                 CONV_RECORD=real_capi.CONV_RECORD, CONV_COMPARE=real_capi.CONV_COMPARE, PUSH3_REDUCED=real_capi.PUSH3_REDUCED)
    monkeypatch.setattr(ionization, "capi", fake)
    cfg = dict(nr=96, nz=64, max_mode=1, rmax=6.0, zmin=0.0, zmax=8.0, dt=10.0, iter_max=3, iter_reltol=1e-3, iter_abstol=1e-3, n0=1.0e17)
    bm = decks.beam_std(cfg["nr"], cfg["nz"], cfg["rmax"], cfg["zmin"], cfg["zmax"], **dict(decks.CONFIGS["C5"]["beam"]))
    for ion_max, nsl in ((1, 48), (2, 40)):
        orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=ion_max, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8, **cfg)
        orc.set_beam(*bm)
        upd = orc.run_slices(nsl)
        st = ionization.IonizationStage(cfg, dict(element=3, ion_max=ion_max, ppc=(2, 2), num_theta=8), bm)
        st.step3d(nslices=nsl, beam_push=False)
        assert st.iters == orc.total_iters() and st.updates == upd > 1000
        assert np.array_equal(st.neut.levels(), orc.levels(ion_max))
        assert st.neut.part.npp() == len(orc.neutral()[4]) > 100
        for name, f in (("psi", st.psi), ("e", st.e), ("b", st.b), ("cu", st.cu)):
            got, want = f.download_f2()[:, :nsl], orc.field(name, 2)[:, :nsl]
            assert np.max(np.abs(want)) > 1e-2 and np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), name
    # a complete step incl. beam push and renewal, then the first slices of the next step
    orc = O.Sim(sp_density=0.0, neut_on=1, neut_elem=3, neut_ion_max=1, neut_ppc1=2, neut_ppc2=2, neut_num_theta=8, ppc1=2, ppc2=2, num_theta=8, **cfg)
    orc.set_beam(*bm)
    orc.step3d(1)
    orc.run_slices(20)
    st = ionization.IonizationStage(cfg, dict(element=3, ion_max=1, ppc=(2, 2), num_theta=8), bm)
    st.step3d()
    st.step3d(nslices=20, beam_push=False)
    got, want = st.psi.download_f2()[:, :20], orc.field("psi", 2)[:, :20]
    assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want))
    assert st.beam.n == len(orc.beam()[2])

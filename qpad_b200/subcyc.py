"""Host-side driver of the sub-cycling / clamp variant of the slice loop (proj_subcyc/simulation_subcyc_class.f03:216-376) on one
xi stage, written against the PER-ROUTINE C-ABI exactly as the Fortran host would call it: one plasma species (robust pusher),
one beam.  Per slice the largest expansion factor gamma / (gamma - p_z) of the plasma decides the number of sub-steps
(:229-236, :431-451); the whole deposit / solve / predictor-corrector / push sequence is repeated with dxi / n_subcyc, and pushed
particles are clamped to `expansion_fac_clamped` (:298-309).

NOT YET VALIDATED ON A GPU: the call sequence is checked on the CPU through oracle-backed adapters
(tests/test_subcyc_host_loop.py), the two new kernels in host emulation (tests/test_emu_kernels.py); the GPU comparison with the
oracle's loop is tests/test_gpu_extras.py::test_subcyc_loop_matches_oracle (QPG_TEST_EXTRAS=1)."""
from . import capi


class SubcycStage:
    def __init__(self, cfg, plasma, beam, device=0):
        """cfg: nr nz max_mode rmax zmin zmax dt iter_max iter_reltol iter_abstol exp_fac_max exp_fac_clamped dt_min ;
        plasma: (x, p, gamma, psi, q) of the injected lattice ; beam: (x, p, q) with xi measured from zmin"""
        self.cfg = cfg
        nr, nz, M = cfg["nr"], cfg["nz"], cfg["max_mode"]
        self.dr, self.dxi = cfg["rmax"] / nr, (cfg["zmax"] - cfg["zmin"]) / nz
        c = self.ctx = capi.Ctx(nr, M, self.dr, self.dxi, device=device)
        F = lambda dim, vol=False: capi.Field(c, dim, nz if vol else 0, vol)
        self.psi, self.e, self.b, self.e_spe, self.b_spe, self.b_beam = F(1, True), F(3, True), F(3, True), F(3, True), F(3, True), F(3)
        self.cu, self.amu, self.acu, self.dcu = F(3, True), F(3), F(2), F(2)
        self.q_spe, self.q_beam, self.beam_q = F(1, True), F(1, True), F(1, True)
        self.s_q, self.s_qn, self.s_cu, self.s_dcu, self.s_amu = F(1, True), F(1), F(3), F(2), F(3)     # the species' own fields
        x, p, g, psi, q = plasma
        self.part = capi.Part2d(c, cfg.get("sp_qbm", -1.0), 2 * len(q))
        self.part.upload(x, p, g, psi, q)
        self.s_q.fill(0.0); self.part.qdeposit(self.s_q)                                                 # species2d%new :118-133: qn = -q
        self.s_q.copy_to(self.s_qn); self.s_qn.scale(-1.0)
        bx, bp, bq = beam
        self.beam = capi.Part3d(c, -1.0, cfg["dt"], len(bq) + 1024, nz, 0, nz)
        self.beam.upload(bx, bp, bq)
        self.iters = self.subcycles = self.updates = 0

    def step3d(self, nslices=None):
        cfg, c, pt = self.cfg, self.ctx, self.part
        nz = cfg["nz"] if nslices is None else nslices
        self.q_beam.fill_f2(0.0); self.q_spe.fill_f2(0.0)                                                 # simulation_class.f03:299-331
        self.beam_q.fill_f2(0.0); self.beam.qdeposit(self.beam_q); self.beam_q.add_f2_to(self.q_beam)
        for f in (self.b, self.e, self.b_spe, self.e_spe, self.psi, self.cu, self.acu, self.amu):
            f.fill(0.0)
        for j in range(1, nz + 1):
            self.updates += pt.npp()
            self.q_beam.copy_slice(j, capi.COPY_2TO1); c.solve_bt(self.q_beam, self.b_beam)               # :218-219
            fac = max(1.0, pt.exp_fac_max())                                                              # :229-236
            dxi_sub, n_sub = capi.subcyc_step(fac, cfg["exp_fac_max"], self.dxi, cfg["dt_min"])           # :431-451
            self.subcycles += n_sub
            for _ in range(n_sub):                                                                        # :239-325
                self.q_spe.fill(0.0)
                self.s_q.fill(0.0); pt.qdeposit(self.s_q); self.s_q.add_to(self.q_spe); self.s_qn.add_to(self.q_spe)   # species2d%qdp
                c.solve_psi(self.q_spe, self.psi)
                c.solve_bz(self.cu, self.b_spe)
                for l in range(cfg["iter_max"]):
                    c.convergence_tester(self.b_spe, 2, capi.CONV_RECORD)
                    capi.Field.add3(self.b_spe, self.b_beam, self.b)
                    c.solve_ez(self.cu, self.e); c.solve_et(self.b, self.psi, self.e)
                    self.cu.fill(0.0); self.acu.fill(0.0); self.amu.fill(0.0)
                    self.s_cu.fill(0.0); self.s_dcu.fill(0.0); self.s_amu.fill(0.0)                        # species2d%amjdp
                    pt.amjdeposit_robust(self.e, self.b, self.s_cu, self.s_amu, self.s_dcu, dxi_sub)
                    self.s_cu.add_to(self.cu); self.s_dcu.add_to(self.acu); self.s_amu.add_to(self.amu)
                    c.solve_djdxi(self.acu, self.amu, self.dcu)
                    c.solve_bt_iter(self.dcu, self.cu, self.b_spe); c.solve_bz(self.cu, self.b_spe)
                    rel, ab = c.convergence_tester(self.b_spe, 2, capi.CONV_COMPARE)
                    self.iters += 1
                    if rel < cfg["iter_reltol"] or ab < cfg["iter_abstol"]:
                        break
                capi.Field.add3(self.b_spe, self.b_beam, self.b)                                          # :292-295
                c.solve_et(self.b_spe, self.psi, self.e_spe); c.solve_ez(self.cu, self.e); c.solve_et(self.b, self.psi, self.e)
                pt.push_u_robust(self.e, self.b, dxi_sub)                                                 # :298-309
                pt.clamp_exp_fac(cfg["exp_fac_clamped"])
                pt.push_x(dxi_sub); pt.update_bound()
            self.s_cu.add_dim_to(self.s_q, [3], [1]); self.s_q.copy_slice(j, capi.COPY_1TO2)              # :330-332 cbq
            self.cu.copy_slice(j, capi.COPY_1TO2)                                                         # :336
            self.cu.add_dim_to(self.q_spe, [3], [1]); self.q_spe.copy_slice(j, capi.COPY_1TO2)
            self.dcu.scale(self.dxi); self.dcu.add_dim_to(self.cu, [1, 2], [1, 2])                        # :347-348 (the full dxi)
            for f in (self.e, self.b, self.psi, self.b_spe, self.e_spe):                                  # :358-362
                f.copy_slice(j, capi.COPY_1TO2)

    def close(self):
        self.part.close()
        self.ctx.close()

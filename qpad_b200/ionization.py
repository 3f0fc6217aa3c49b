"""Host-side driver of the ionisation deck (input_file/ionization: nspecies 0, one ADK neutral species, one beam) on one
xi stage, written against the PER-ROUTINE C-ABI exactly as the Fortran host would call it (simulation_class.f03:294-512
with nneutrals = 1): every line below is one type-bound procedure of the reference.

NOT YET VALIDATED ON A GPU: the neutral kernels (csrc/neutral.cu) were written at the end of round 1 without GPU time left;
tests/test_gpu_neutral.py (QPG_TEST_NEUTRAL=1) compares this loop with the oracle's (oracle/qpad_oracle_neutral.c)."""
from . import capi


class IonizationStage:
    def __init__(self, cfg, neutral, beam, device=0):
        """cfg: nr nz max_mode rmax zmin zmax dt iter_max iter_reltol iter_abstol n0 ; neutral: element ion_max ppc num_theta q m
        density ; beam: (x, p, q) arrays with xi measured from zmin"""
        self.cfg = cfg
        nr, nz, M = cfg["nr"], cfg["nz"], cfg["max_mode"]
        self.dr, self.dxi = cfg["rmax"] / nr, (cfg["zmax"] - cfg["zmin"]) / nz
        c = self.ctx = capi.Ctx(nr, M, self.dr, self.dxi, device=device)
        F = lambda dim, vol=False: capi.Field(c, dim, nz if vol else 0, vol)
        self.psi, self.e, self.b, self.e_spe, self.b_spe, self.b_beam = F(1, True), F(3, True), F(3, True), F(3, True), F(3, True), F(3)
        self.cu, self.amu, self.acu, self.dcu = F(3, True), F(3), F(2), F(2)
        self.q_spe, self.q_beam, self.beam_q = F(1, True), F(1, True), F(1, True)
        self.n_q, self.n_cu, self.n_dcu, self.n_amu = F(1, True), F(3), F(2), F(3)          # the neutral's own deposit fields
        self.rho_ion, self.rho_ion_add = F(1, True), F(1)
        self.neut = capi.Neutral(c, neutral["element"], neutral["ion_max"], neutral["ppc"], neutral["num_theta"], neutral.get("q", -1.0),
                                 neutral.get("m", 1.0), neutral.get("density", 1.0), cfg.get("n0", 1.0e17), self.dxi)
        bx, bp, bq = beam
        self.beam = capi.Part3d(c, -1.0, cfg["dt"], len(bq) + 1024, nz, 0, nz)
        self.beam.upload(bx, bp, bq)
        self.iters = 0
        self.updates = 0

    def step3d(self, nslices=None, beam_push=True):
        cfg, c, ne = self.cfg, self.ctx, self.neut
        el, ions = ne.part, ne.part_add
        nz = cfg["nz"] if nslices is None else nslices
        # simulation_class.f03:299-331
        self.q_beam.fill_f2(0.0); self.q_spe.fill_f2(0.0)
        self.beam_q.fill_f2(0.0); self.beam.qdeposit(self.beam_q); self.beam_q.add_f2_to(self.q_beam)    # beam3d%qdp
        for f in (self.b, self.e, self.b_spe, self.e_spe, self.psi, self.cu, self.acu, self.amu):
            f.fill(0.0)
        for j in range(1, nz + 1):
            self.updates += el.npp()
            self.q_beam.copy_slice(j, capi.COPY_2TO1); c.solve_bt(self.q_beam, self.b_beam)               # :344-345
            self.q_spe.fill(0.0)                                                                            # :346
            self.n_q.fill(0.0); el.qdeposit(self.n_q); self.n_q.add_to(self.q_spe)                          # neut%qdp
            self.rho_ion_add.fill(0.0); ions.qdeposit(self.rho_ion_add)                                     # neut%ion_deposit
            self.rho_ion_add.add_to(self.rho_ion); self.rho_ion.add_to(self.q_spe); ions.clear()
            c.solve_psi(self.q_spe, self.psi)                                                               # :356
            c.solve_bz(self.cu, self.b_spe)                                                                 # :360
            for l in range(cfg["iter_max"]):                                                                # :370
                c.convergence_tester(self.b_spe, 2, capi.CONV_RECORD)
                capi.Field.add3(self.b_spe, self.b_beam, self.b)
                c.solve_ez(self.cu, self.e); c.solve_et(self.b, self.psi, self.e)
                self.cu.fill(0.0); self.acu.fill(0.0); self.amu.fill(0.0)
                self.n_cu.fill(0.0); self.n_dcu.fill(0.0); self.n_amu.fill(0.0)                              # neut%amjdp
                el.amjdeposit_robust(self.e, self.b, self.n_cu, self.n_amu, self.n_dcu, self.dxi)
                self.n_cu.add_to(self.cu); self.n_dcu.add_to(self.acu); self.n_amu.add_to(self.amu)
                c.solve_djdxi(self.acu, self.amu, self.dcu)                                                 # :390
                c.solve_bt_iter(self.dcu, self.cu, self.b_spe); c.solve_bz(self.cu, self.b_spe)             # :391-392
                rel, ab = c.convergence_tester(self.b_spe, 2, capi.CONV_COMPARE)
                self.iters += 1
                if rel < cfg["iter_reltol"] or ab < cfg["iter_abstol"]:
                    break
            self.n_cu.add_dim_to(self.n_q, [3], [1]); self.n_q.copy_slice(j, capi.COPY_1TO2)                # neut%cbq
            self.rho_ion.copy_slice(j, capi.COPY_1TO2)
            self.cu.copy_slice(j, capi.COPY_1TO2)                                                           # :409
            self.cu.add_dim_to(self.q_spe, [3], [1]); self.q_spe.copy_slice(j, capi.COPY_1TO2)              # :410-411
            capi.Field.add3(self.b_spe, self.b_beam, self.b)                                                # :413
            c.solve_et(self.b_spe, self.psi, self.e_spe); c.solve_ez(self.cu, self.e); c.solve_et(self.b, self.psi, self.e)
            self.dcu.scale(self.dxi); self.dcu.add_dim_to(self.cu, [1, 2], [1, 2])                          # :425-426
            ne.update(self.e)                                                                               # :445 neut%update
            el.push_u_robust(self.e, self.b, self.dxi); el.push_x(self.dxi); el.update_bound()              # :446-447
            for f in (self.e, self.b, self.psi, self.b_spe, self.e_spe):                                    # :452-456
                f.copy_slice(j, capi.COPY_1TO2)
        if beam_push:
            self.beam.push(capi.PUSH3_REDUCED, self.e, self.b); self.beam.update_bound()                    # :489-493
            ne.renew()                                                                                      # :504-510
            self.rho_ion.fill(0.0); self.n_q.fill(0.0); self.n_cu.fill(0.0)

    def close(self):
        self.neut.close()
        self.ctx.close()

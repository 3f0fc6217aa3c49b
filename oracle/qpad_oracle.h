/*
 * qpad_oracle.h -- CPU restatement (fp64, single thread) of QPAD's quasi-static
 * slice loop.  TEST INFRASTRUCTURE ONLY: nothing under qpad_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (Fortran 2003 + MPI + HYPRE + HDF5) cannot be
 * built in this image and ships no golden vectors or assertions (SURVEY.md §4,
 * §8c).  This file is pinned instead against analytic known answers
 * (tests/test_oracle_*.py) and is a line-by-line restatement of the cited
 * Fortran.  The tridiagonal systems the reference hands to HYPRE's cyclic
 * reduction (source/fields/field_solver_class.f03:172-177; HYPRE version not
 * pinned by the reference, docs recommend 2.11.x) are solved here by the
 * Thomas algorithm -- a direct solve of the same matrix.
 *
 * Layouts follow the reference:
 *   particles  x(2,np) p(3,np) AoS column-major, gamma/psi/q(np)
 *   field f1   per plane [0:nr+1][dim]  (Fortran f1(dim,0:nr+1)),
 *              planes ordered re0, re1, im1, re2, im2, ...
 *   field f2   per plane [1:nzp+1][0:nr+1][dim]
 */
#ifndef QPAD_ORACLE_H
#define QPAD_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* field-solver kinds, source/param.f03 p_fk_* */
enum { ORC_FK_PSI = 0, ORC_FK_EZ = 1, ORC_FK_BZ = 2, ORC_FK_BT = 3, ORC_FK_BPLUS = 4, ORC_FK_BMINUS = 5, ORC_FK_VPOTZ = 6, ORC_FK_VPOTP = 7, ORC_FK_VPOTM = 8 };
enum { ORC_BND_ZERO = 2, ORC_BND_OPEN = 3 };
enum { ORC_PUSH3_REDUCED = 1, ORC_PUSH3_BORIS = 2 };

/* ---- stand-alone routines (one per reference routine) ------------------- */
/* part2d_class.f03:231 */
void orc_qdeposit(const double *x, const double *q, long npp, double dr, int nr, int max_mode, double *f1);
/* part2d_class.f03:746 ; e,b dim 3, cu dim 3, dcu dim 2, amu dim 3 ; writes gamma, psi */
void orc_amjdeposit_robust(const double *x, const double *p, const double *q, double *gamma, double *psi, long npp,
                           double dr, int nr, int max_mode, double qbm, double dt, const double *ef, const double *bf,
                           double *cu, double *dcu, double *amu);
/* part2d_class.f03:1879 */
void orc_push_u_robust(const double *x, double *p, double *gamma, long npp, double dr, int nr, int max_mode,
                       double qbm, double dt, const double *ef, const double *bf);
/* part2d_class.f03:2221 */
void orc_push_x(double *x, const double *p, const double *gamma, long npp, double dt);
/* std pusher flavour: species/part2d_class.f03:478 amjdeposit_std, :1790 push_u_std, :2264 interp_psi (with its quirk) */
void orc_amjdeposit_std(const double *x, const double *p, const double *q, double *gamma, double *psi, long npp,
                        double dr, int nr, int max_mode, double qbm, double dt, const double *ef, const double *bf,
                        double *cu, double *dcu, double *amu);
void orc_push_u_std(const double *x, double *p, double *gamma, const double *psi, long npp, double dr, int nr, int max_mode,
                    double qbm, double dt, const double *ef, const double *bf);
void orc_interp_psi(const double *x, double *psi, long npp, double dr, int nr, int max_mode, const double *psif);
/* ponderomotive-guiding-centre flavours: :1012 amjdeposit_std_pgc, :1310 amjdeposit_robust_pgc, :1967/:2094 push_u_*_pgc */
void orc_amjdeposit_pgc(const double *x, const double *p, const double *q, double *gamma, double *psi, long npp,
                        double dr, int nr, int max_mode, double qbm, double dt, const double *ef, const double *bf,
                        const double *ar, const double *ai, const double *ar_grad, const double *ai_grad,
                        double *cu, double *dcu, double *amu, int push_std);
void orc_push_u_pgc(const double *x, double *p, double *gamma, const double *psi, long npp, double dr, int nr, int max_mode,
                    double qbm, double dt, const double *ef, const double *bf, const double *ar, const double *ai,
                    const double *ar_grad, const double *ai_grad);
/* part2d_class.f03:2307 ; returns new npp */
long orc_update_bound(double *x, double *p, double *gamma, double *psi, double *q, long npp, double edge);
/* sort_module.f03:11 + part2d_class.f03:2498 ; ip is 1-based like the reference */
void orc_sort_idx(const double *x, long npp, double dr, int nrp, int *ix, int *ip);
void orc_sort_part2d(double *x, double *p, double *gamma, double *psi, double *q, long npp, double dr, int nrp);
/* fdist2d_class.f03:289 (uniform profile, uth=0, ordered theta); returns npp */
long orc_inject_uniform(double *x, double *p, double *gamma, double *psi, double *q, int nr, double dr, int ppc1,
                        int ppc2, int num_theta, double qm, double density, double den_min);

/* field_solver_class.f03:256 ; writes a,b,c (nr each) */
void orc_build_matrix(int kind, int mode, int nr, double dr, int bnd, double relax_fac, double *a, double *b, double *c);
void orc_tridiag_solve(const double *a, const double *b, const double *c, double *d, int nr);
void orc_tridiag_solve_ld(const double *a, const double *b, const double *c, double *d, int nr); /* long double check */

/* field ops; all f1 arrays are full multi-plane fields */
void orc_solve_psi(const double *q, double *psi, int nr, int max_mode, double dr, int bnd);
void orc_solve_bt(const double *qb, double *b, int nr, int max_mode, double dr, int bnd);
void orc_solve_bz(const double *cu, double *b, int nr, int max_mode, double dr, int bnd);
void orc_solve_bt_iter(const double *dcu, const double *cu, double *b, int nr, int max_mode, double dr, int bnd,
                       double relax_fac);
void orc_solve_ez(const double *cu, double *e, int nr, int max_mode, double dr, int bnd);
void orc_solve_et(const double *b, const double *psi, double *e, int nr, int max_mode, double dr);
void orc_solve_et_beam(const double *b, double *e, int nr, int max_mode);
void orc_solve_djdxi(const double *acu, const double *amu, double *dcu, int nr, int max_mode, double dr);
void orc_smooth_f1(double *f1plane, int dim, int nr, const int *ax_smooth);
/* vector-potential diagnostics, fields/field_vpot_class.f03:354 / :392 (vpot: dim-3 field A_r, A_phi, A_z) */
void orc_solve_vpotz(const double *cu, double *vpot, int nr, int max_mode, double dr, int bnd);
void orc_solve_vpott(const double *cu, double *vpot, int nr, int max_mode, double dr, int bnd);

/* part3d_class.f03:221/477/358/640 ; x(3,np), p(3,np) ; f2 volumes with nzp+1 slices, noff2 = slab offset */
void orc_qdeposit3d(const double *x, const double *q, long npp, double dr, double dz, int nr, int nzp, int noff2,
                    int max_mode, double *f2);
void orc_push3d(double *x, double *p, long npp, double dr, double dz, int nr, int nzp, int noff2, int max_mode,
                double qbm, double dt, int push_type, const double *ef2, const double *bf2);
long orc_update_bound3d(double *x, double *p, double *q, long npp, double edge_r, double edge_z);
/* the same two with the spin vectors of a beam with has_spin (part3d_class.f03:53-63, :578-638 push_spin, :668-670); spin = NULL: none */
void orc_push3d_spin(double *x, double *p, double *spin, double amm, long npp, double dr, double dz, int nr, int nzp, int noff2, int max_mode,
                     double qbm, double dt, int push_type, const double *ef, const double *bf);
long orc_update_bound3d_spin(double *x, double *p, double *q, double *spin, long npp, double edge_r, double edge_z);

/* ---- whole simulation (simulation_class.f03:226-512), optionally as S xi-stages run in sequence ---- */
typedef struct orc_sim orc_sim;
typedef struct {
    int nr, nz, max_mode, bnd, iter_max, nstages;
    double rmax, zmin, zmax, dt, iter_reltol, iter_abstol, relax_fac; /* relax_fac<0 => default */
    /* one species (uniform) */
    int ppc1, ppc2, num_theta, sort_freq;
    double sp_q, sp_m, sp_density, sp_den_min;
    /* one beam */
    int beam_push_type, beam_evol;
    double beam_qbm;
    int sp_push_type;   /* p_push2_std = 0, p_push2_robust = 1, p_push2_std_pgc = 4, p_push2_robust_pgc = 5 (param.f03:63-64) */
    /* one laser envelope (nlasers = 1; single stage only): the envelope is handed in with orc_sim_set_laser */
    int laser_on, laser_iter;
    double laser_k0;
    /* one field-ionisation neutral species (nneutrals = 1; single stage only; uniform profile, robust pusher):
     * element = atomic number (1 H, 2 He, 3 Li), ion_max = deepest charge state followed, n0 = plasma density in cm^-3 */
    int neut_on, neut_elem, neut_ion_max, neut_ppc1, neut_ppc2, neut_num_theta;
    double neut_q, neut_m, neut_density, n0;
    /* "subcycling" algorithm (proj_subcyc/simulation_subcyc_class.f03): simulation.expansion_fac_max / expansion_fac_clamped / dt_2d_min */
    int subcyc_on;
    double subcyc_exp_fac_max, subcyc_exp_fac_clamped, subcyc_dt_min;
} orc_params;

orc_sim *orc_sim_create(const orc_params *prm);
void orc_sim_destroy(orc_sim *s);
/* hand the beam particles (global xi measured from zmin) to the stage that owns them */
void orc_sim_set_beam(orc_sim *s, const double *x, const double *p, const double *q, long np);
/* run one 3D step (all stages in pipeline order).  Returns total plasma particle-slice updates. */
long orc_sim_step3d(orc_sim *s, int istep);
/* run only the first `nslices` slices of stage 0 of a 3D step (parity at slice granularity; no beam push) */
long orc_sim_run_slices(orc_sim *s, int nslices);
/* continue stage 0 with slices j0..j1 (no re-initialisation) */
long orc_sim_run_range(orc_sim *s, int j0, int j1);
/* accessors (copy out) */
int orc_sim_nzp(const orc_sim *s, int stage);
long orc_sim_plasma_np(const orc_sim *s, int stage);
void orc_sim_get_plasma(const orc_sim *s, int stage, double *x, double *p, double *gamma, double *psi, double *q);
long orc_sim_beam_np(const orc_sim *s, int stage);
void orc_sim_get_beam(const orc_sim *s, int stage, double *x, double *p, double *q);
/* name in {psi,e,b,e_spe,b_spe,e_beam,b_beam,cu,amu,acu,dcu,q_spe,q_beam}; which=1 -> f1, 2 -> f2 (nzp+1 slices) */
long orc_sim_get_field(const orc_sim *s, int stage, const char *name, int which, double *out);
/* bench.py's CPU arm: save / restore the slice state of stage 0 (one snapshot per process) */
void orc_sim_snapshot(orc_sim *s);
void orc_sim_restore(orc_sim *s);
long orc_sim_total_iters(const orc_sim *s);
void orc_sim_get_slice_iters(const orc_sim *s, int stage, int *out);   /* PC iterations of each slice of the stage's last sweep (nzp ints) */
long orc_sim_total_subcycles(const orc_sim *s);   /* sub-steps taken so far (= slices when nothing was sub-cycled) */
/* proj_subcyc/part2d_subcyc_class.f03:28 / :48 ; simulation_subcyc_class.f03:431 */
double orc_exp_fac_max(const double *p, const double *gamma, long npp);
void orc_clamp_exp_fac(double *p, double *gamma, long npp, double exp_fac_clamped);
void orc_subcyc_step(double exp_fac, double exp_fac_max, double dt, double dt_min, double *dt_subcyc, int *n_subcyc);
/* laser envelope volumes a_r, a_i (layout of orc_laser_volume_size) of the single stage: copy in / copy out; chi = the
 * susceptibility volume deposited during the last 3D step, (P, nz+1, nr+2) */
void orc_sim_set_laser(orc_sim *s, const double *ar, const double *ai);
void orc_sim_get_laser(const orc_sim *s, double *ar, double *ai, double *chi);

/* ---- field-ionisation neutral species, qpad_oracle_neutral.c (species/neutral_class.f03) -----------------------------
 * level array lev[(i * n_theta + k) * nr + j]: i < multi_max charge states 1..multi_max, i = multi_max neutral residue,
 * i = multi_max + 1 total discrete ion level */
int orc_adk_params(int element, int max_e, double *out);
double orc_plasma_frequency(double n0);
void orc_neutral_reset(double *lev, int nr, int n_theta, int multi_max);
void orc_neutral_ionize(double *lev, const double *adk, const double *ef, double wp, double dt, int ppc1, int ppc2, int nr, int n_theta,
                        int max_mode, int multi_max);
long orc_neutral_add_particles(const double *lev, const double *ion_old, int nr, int n_theta, int multi_max, int ppc1, int ppc2, double dr,
                               double qm, double density, double den_min, double *x, double *p, double *gamma, double *psi, double *q, long *npp,
                               double *xa, double *qa);
void orc_neutral_ion_deposit(const double *xa, const double *qa, long nadd, double dr, int nr, int max_mode, double *rho_ion, double *q_tot);
/* whole-loop accessors of the neutral species of a stage: electrons created so far (plasma-particle layout), level array */
long orc_sim_neutral_np(const orc_sim *s, int stage);
void orc_sim_get_neutral(const orc_sim *s, int stage, double *x, double *p, double *gamma, double *psi, double *q);
void orc_sim_get_levels(const orc_sim *s, int stage, double *lev);

/* ---- laser envelope (ponderomotive guiding centre) field path, qpad_oracle_laser.c ---------------------------------
 * laser volumes v[plane][slice j = -1..nz+1 at index j+1][node 0..nr+1]; chi = dim-1 f2 volume; gradients = dim-3 f1 fields */
long orc_laser_volume_size(int nr, int nz, int max_mode);
/* species/part2d_class.f03:2581 ; :361 (chi = dim-1 multi-plane f1, accumulated into) */
double orc_deposit_ax_corr(int ppc_r);
void orc_deposit_chi(const double *x, const double *q, const double *psi, long npp, double dr, int nr, int max_mode, double qbm,
                     double ax_corr, double *chi);
/* laser/profile_laser_lib.f03:56 ; laser/profile_laser_class.f03:318 (gaussian x sin2) */
void orc_laser_gaussian_point(double r, double z, double k, double k0, double w0, double f_dist, double *ar, double *ai);
void orc_laser_launch_gaussian(double k0, double a0, double w0, double f_dist, double lon_center, double t_rise, double t_flat, double t_fall,
                               double z0, double dz, double dr, int nr, int nz, int max_mode, double *ar, double *ai);
/* laser/field_laser_class.f03:269 (A = [2nr][5] rows a,b,c,d,e) ; direct pentadiagonal solve (rhs -> x in place) */
void orc_laser_build_matrix(int m, int nr, double k0, double ds, double dr, double dz, double *A);
void orc_penta_solve(const double *A, double *x, int n);
/* laser/field_laser_class.f03:393 set_rhs, :752 solve, :637 set_grad */
void orc_laser_set_rhs(const double *ar, const double *ai, const double *chi, int nr, int nz, int max_mode, double k0, double ds, double dr,
                       double dz, double *sr, double *si);
void orc_laser_solve(double *ar, double *ai, const double *sr, const double *si, const double *chi, int nr, int nz, int max_mode,
                     double k0, double ds, double dr, double dz, int iter);
void orc_laser_set_grad(const double *ar, const double *ai, int slice, int nr, int nz, int max_mode, double dr, double dz, double *ar_grad,
                        double *ai_grad);

#ifdef __cplusplus
}
#endif
#endif

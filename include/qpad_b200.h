/*
 * qpad_b200.h -- C-ABI of libqpadb200.so: QPAD's quasi-static slice loop on one B200 (sm_100a).
 *
 * This is the drop-in boundary of SURVEY.md §8(b): every entry point replaces the body of one
 * Fortran type-bound procedure of the reference (cited per function, paths relative to
 * /root/reference/source/), and is what an ISO_C_BINDING shim binds (INTEGRATION.md shows the
 * `bind(C)` interface block a maintainer adds).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative qpg_status otherwise (the reference never
 *     returns errors: write_err logs and stops, sysutil_module.f03:165; the shim maps non-zero
 *     to write_err).  qpg_last_error() gives the message.
 *   - all arithmetic is fp64 (`real` == fp64 in every reference build config).
 *   - handles are opaque; memory they refer to lives in HBM and is owned by the library.
 *   - work is enqueued on the context's CUDA stream; only functions documented "synchronises"
 *     wait for the device.  One host thread per context.
 *   - HOST array layouts are the reference's:
 *        particles  x(2,np) p(3,np) column-major AoS, gamma(np) psi(np) q(np)
 *        field f1   per plane f1(dim, 0:nr+1); planes ordered re(0), re(1), im(1), re(2), im(2) ...
 *                   i.e. a C array [P][nr+2][dim], P = 2*max_mode+1  (same as the pipeline wire
 *                   buffer (dim, nr+2, 2M+1) of fields/field_class.f03:608-634)
 *        field f2   C array [P][nzp+1][nr+2][dim]
 *     The DEVICE layout is different (SoA particles; node-interleaved fields).
 *   - there is no CPU fallback: without a CUDA device every call fails with QPG_ERR_CUDA.
 */
#ifndef QPAD_B200_H
#define QPAD_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QPG_OK = 0, QPG_ERR_ARG = -1, QPG_ERR_CUDA = -2, QPG_ERR_ALLOC = -3, QPG_ERR_UNSUPPORTED = -4, QPG_ERR_STATE = -5
} qpg_status;

typedef struct qpg_ctx_s *qpg_ctx;
typedef struct qpg_field_s *qpg_field;
typedef struct qpg_part2d_s *qpg_part2d;
typedef struct qpg_part3d_s *qpg_part3d;
typedef struct qpg_sim_s *qpg_sim;
typedef struct qpg_laser_s *qpg_laser;
typedef struct qpg_neutral_s *qpg_neutral;
typedef struct qpg_stage_s *qpg_stage;

/* param.f03 constants mirrored 1:1 */
enum { QPG_BND_ZERO = 2, QPG_BND_OPEN = 3 };                              /* p_bnd_* */
enum { QPG_PUSH2_STD = 0, QPG_PUSH2_ROBUST = 1, QPG_PUSH2_STD_PGC = 4, QPG_PUSH2_ROBUST_PGC = 5 };   /* p_push2_* (param.f03:63-64) */
enum { QPG_PUSH3_REDUCED = 1, QPG_PUSH3_BORIS = 2 };                      /* p_push3_* */
enum { QPG_COPY_1TO2 = 0, QPG_COPY_2TO1 = 1 };                            /* ufield_class.f03 p_copy_* */
enum { QPG_CONV_RECORD = 0, QPG_CONV_COMPARE = 1 };

const char *qpg_last_error(void);
int qpg_version(void);

/* ------------------------------------------------------------------------------------------ */
/* context: grid + solver set.  Replaces options_class.f03 get_* values and the per-(kind,mode)   */
/* field_solver%new calls of fields/field_{psi,e,b}_class.f03 (HYPRE StructCycRed setup,       */
/* fields/field_solver_class.f03:58-95, 256-561).                                             */
/* cuda_stream: a cudaStream_t (0/NULL = library creates its own).  device: CUDA ordinal.      */
/* relax_fac < 0 selects the reference default 1e-3*(dr/0.02)^2 (sim_fields_class.f03:137).    */
/* ------------------------------------------------------------------------------------------ */
int qpg_ctx_create(qpg_ctx *out, int device, void *cuda_stream, int nr, int max_mode, double dr, double dxi,
                   int field_boundary, double relax_fac);
int qpg_ctx_destroy(qpg_ctx ctx);
int qpg_ctx_sync(qpg_ctx ctx);                                   /* synchronises */
/* named timers with the reference's event names (sysutil_module.f03:298-385 start/stop_tprof):
 * CUDA-event timing per kernel class.  qpg_tprof_get synchronises. */
int qpg_tprof_enable(qpg_ctx ctx, int on);
int qpg_tprof_reset(qpg_ctx ctx);
int qpg_tprof_get(qpg_ctx ctx, const char *event, double *ms_total, long *ncalls);
/* number of kernels this context has launched so far */
long qpg_launch_count(qpg_ctx ctx);

/* ------------------------------------------------------------------------------------------ */
/* field : fields/field_class.f03:50 `type field` + fields/ufield_class.f03:52 `type ufield`   */
/* ------------------------------------------------------------------------------------------ */
int qpg_field_create(qpg_field *out, qpg_ctx ctx, int dim, int nzp, int has_2d);      /* field%new  :190 */
int qpg_field_destroy(qpg_field f);                                                   /* field%del  :262 */
int qpg_field_dim(qpg_field f);
int qpg_field_fill(qpg_field f, double value);                                        /* assign_f1  :917 (`f = value`) */
int qpg_field_fill_f2(qpg_field f, double value);                                     /* assign_f2  :948 (`f%as(value)`) */
int qpg_field_copy(qpg_field src, qpg_field dst);                                     /* assign_f1, `dst = src` */
int qpg_field_copy_slice(qpg_field f, int idx, int dir);                              /* ufield copy_slice :341 (idx 1-based) */
int qpg_field_add(qpg_field a, qpg_field b);                                          /* add_f1(a,b): b += a  :1053 */
int qpg_field_add3(qpg_field a1, qpg_field a2, qpg_field a3);                         /* add_f1(a1,a2,a3): a3 = a1+a2 :1035 */
int qpg_field_add_dim(qpg_field a, qpg_field b, int ndim, const int *adim, const int *bdim); /* add_f1(a,b,adim,bdim) :1011 (1-based comps) */
int qpg_field_add_f2(qpg_field a, qpg_field b);                                       /* add_f2(a,b): b%f2 += a%f2 */
int qpg_field_scale(qpg_field f, double s);                                           /* dot_f1(s, f) :1159 */
int qpg_field_smooth(qpg_field f, int order, int kind);                               /* field_src_class.f03:102/169/243; kind 0 rho,1 jay,2 djdxi */
int qpg_field_upload_f1(qpg_field f, const double *host);                             /* host [P][nr+2][dim] */
int qpg_field_download_f1(qpg_field f, double *host);                                 /* synchronises */
int qpg_field_upload_f2(qpg_field f, const double *host);                             /* host [P][nzp+1][nr+2][dim] */
int qpg_field_download_f2(qpg_field f, double *host);                                 /* synchronises */
/* pipeline wire buffers (device pointers, reference wire layout [P][nr+2][dim]):
 * field_class.f03:560 pipe_send_f1 / :647 pipe_recv_f1 / :312 pipe_send_f2 / :420 pipe_recv_f2.
 * slice = 0 packs/unpacks f1, slice k>=1 the f2 slice k; add != 0 selects mode 'add'.
 * The transport itself (NCCL send/recv over NVLink) is done by the caller on these buffers. */
int qpg_field_pack(qpg_field f, int slice, double *dev_buf);
int qpg_field_unpack(qpg_field f, int slice, const double *dev_buf, int add);
long qpg_field_wire_count(qpg_field f);                                               /* doubles per wire buffer */
/* diagnostics staging: f2(comp, node, 1:nzp) of one plane -> host (nzp doubles); comp 1-based, node 0..nr+1.
 * The on-axis E_z / psi line-outs of the parity gate; synchronises. */
int qpg_field_lineout(qpg_field f, int comp, int plane, int node, double *host);

/* ------------------------------------------------------------------------------------------ */
/* field solves: one call per reference `solve` generic                                       */
/* ------------------------------------------------------------------------------------------ */
int qpg_solve_psi(qpg_ctx ctx, qpg_field q, qpg_field psi);                 /* field_psi_class.f03:218 solve_field_psi */
int qpg_solve_bt(qpg_ctx ctx, qpg_field q_beam, qpg_field b);              /* field_b_class.f03:798 solve_field_bt */
int qpg_solve_bz(qpg_ctx ctx, qpg_field cu, qpg_field b);                  /* field_b_class.f03:760 solve_field_bz */
int qpg_solve_bt_iter(qpg_ctx ctx, qpg_field dcu, qpg_field cu, qpg_field b); /* field_b_class.f03:836 solve_field_bt_iter */
int qpg_solve_ez(qpg_ctx ctx, qpg_field cu, qpg_field e);                  /* field_e_class.f03:298 solve_field_ez */
int qpg_solve_et(qpg_ctx ctx, qpg_field b, qpg_field psi, qpg_field e);    /* field_e_class.f03:412 solve_field_et */
int qpg_solve_et_beam(qpg_ctx ctx, qpg_field b, qpg_field e);              /* field_e_class.f03:516 solve_field_et_beam */
int qpg_solve_djdxi(qpg_ctx ctx, qpg_field acu, qpg_field amu, qpg_field dcu); /* field_src_class.f03:273 solve_field_djdxi */
/* simulation_class.f03:522 convergence_tester(fld, dim, op, rel_res, abs_res).
 * op = QPG_CONV_COMPARE synchronises and returns rel/abs. */
int qpg_bperp_residual(qpg_ctx ctx, qpg_field fld, int dim, int op, double *rel_res, double *abs_res);

/* ------------------------------------------------------------------------------------------ */
/* part2d : species/part2d_class.f03:24 `type part2d`                                          */
/* ------------------------------------------------------------------------------------------ */
int qpg_part2d_create(qpg_part2d *out, qpg_ctx ctx, double qbm, long npmax);                     /* init_part2d :98 */
int qpg_part2d_destroy(qpg_part2d p);                                                            /* end_part2d :144 */
int qpg_part2d_upload(qpg_part2d p, const double *x, const double *pm, const double *gamma, const double *psi,
                      const double *q, long npp);                                                /* host AoS -> device SoA (inject/renew) */
int qpg_part2d_download(qpg_part2d p, double *x, double *pm, double *gamma, double *psi, double *q, long *npp); /* synchronises; any ptr may be NULL */
int qpg_part2d_npp(qpg_part2d p, long *npp);                                                     /* synchronises */
/* renew from the device-resident copy of the lattice made by the first upload (renew_part2d :207: the
 * uniform/time-independent profiles of the decks re-inject the same particles every 3D step) */
int qpg_part2d_snapshot(qpg_part2d p);
int qpg_part2d_renew(qpg_part2d p);
int qpg_part2d_qdeposit(qpg_part2d p, qpg_field q);                                              /* qdeposit_part2d :231 */
int qpg_part2d_amjdeposit(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, qpg_field cu, qpg_field amu,
                          qpg_field dcu, double dt);                  /* amjdeposit_robust_part2d :746 / amjdeposit_std_part2d :478 */
int qpg_part2d_push_u(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, double dt);       /* push_u_robust_part2d :1879 / push_u_std_part2d :1790 */
/* ponderomotive-guiding-centre flavours (push_type = QPG_PUSH2_STD_PGC | QPG_PUSH2_ROBUST_PGC).  The laser envelope
 * arrives as four fields on the same grid: a_r, a_i (dim 1; field_laser get_cfr / get_cfi) and their gradients
 * (dim 3, cylindrical components; field_laser ar_grad / ai_grad).  amjdeposit_std_pgc_part2d :1012,
 * amjdeposit_robust_pgc_part2d :1310, push_u_robust_pgc_part2d :1967, push_u_std_pgc_part2d :2094.  The envelope
 * solver itself (laser/field_laser_class.f03) is not part of this library. */
int qpg_part2d_amjdeposit_pgc(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, qpg_field ar, qpg_field ai,
                              qpg_field ar_grad, qpg_field ai_grad, qpg_field cu, qpg_field amu, qpg_field dcu, double dt);
int qpg_part2d_push_u_pgc(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, qpg_field ar, qpg_field ai,
                          qpg_field ar_grad, qpg_field ai_grad, double dt);
int qpg_part2d_push_x(qpg_part2d p, double dt);                                                  /* push_x_part2d :2221 */
/* interp_psi_part2d :2264 (std pushers, simulation_class.f03:357-359), including the reference's quirk: `pp` is never
 * advanced (:2298-2301), so only the first particle of each 1024-particle chunk is written, with the last one's value */
int qpg_part2d_interp_psi(qpg_part2d p, qpg_field psi);
int qpg_part2d_update_bound(qpg_part2d p);                                                       /* update_bound_part2d :2307 */
int qpg_part2d_move(qpg_part2d p);   /* move_part2d_comm species/part2d_comm.f03:147: no radial decomposition on one GPU -> no-op */
int qpg_part2d_sort(qpg_part2d p);                                                               /* sort_part2d :2498 + sort_module.f03:11 */
int qpg_part2d_sort_index(qpg_part2d p, int *host_ix, int *host_ip);                             /* generate_sort_idx_1d output (1-based), synchronises */
/* pipesend_part2d :2355 / piperecv_part2d :2405 : dev_buf[0] = count, then 8 doubles per particle
 * (x1,x2,p1,p2,p3,gamma,psi,q) AoS.  Capacity = 8*npmax+1 doubles; the transport may move just the live prefix
 * 1 + 8*n (n >= count, e.g. the injected lattice size: plasma particles only ever leave). */
int qpg_part2d_pack(qpg_part2d p, double *dev_buf);
int qpg_part2d_unpack(qpg_part2d p, const double *dev_buf);
long qpg_part2d_wire_count(qpg_part2d p);

/* ------------------------------------------------------------------------------------------ */
/* part3d : beam/part3d_class.f03:24 `type part3d` (no spin)                                    */
/* x(3,np) = (x, y, xi - z0) like the reference; noff2/nzp = this stage's xi slab                */
/* ------------------------------------------------------------------------------------------ */
int qpg_part3d_create(qpg_part3d *out, qpg_ctx ctx, double qbm, double dt, long npmax, int nz_total, int noff2, int nzp); /* init_part3d :98 */
int qpg_part3d_destroy(qpg_part3d p);
int qpg_part3d_upload(qpg_part3d p, const double *x, const double *pm, const double *q, long npp);
int qpg_part3d_download(qpg_part3d p, double *x, double *pm, double *q, long *npp);              /* synchronises */
int qpg_part3d_qdeposit(qpg_part3d p, qpg_field q);                                              /* qdeposit_part3d :221 (into q%f2) */
int qpg_part3d_push(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf);                    /* push_reduced :477 / push_boris :358 */
/* push_* in two passes for a stage of the xi-pipeline (part3d_class.f03:358-576 per particle, unchanged): `interior` advances the particles
 * whose field gather does not touch the guard slice nzp + 1 -- it does not need the downstream stage's first-slice e / b
 * (simulation_class.f03:482-483) -- `edge` the others.  interior, then edge == qpg_part3d_push */
int qpg_part3d_push_interior(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf);
int qpg_part3d_push_edge(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf);
/* qdeposit_part3d :221-316 (the scatter) restricted to: part 1 = the particles advanced by qpg_part3d_push_interior, 2 = the others (call
 * before update_bound), 3 = the particles appended by the last qpg_part3d_unpack (call before qpg_part3d_pack_forward) */
int qpg_part3d_qdeposit_part(qpg_part3d p, qpg_field q, int part);
int qpg_part3d_update_bound(qpg_part3d p);                                                       /* update_bound_part3d :640 */
/* qpg_part3d_qdeposit in two halves (see qpg_sim_beam_qdp_raw / _fix) */
int qpg_part3d_qdeposit_raw(qpg_part3d p, qpg_field q);
int qpg_part3d_qdeposit_fix(qpg_part3d p, qpg_field q);
/* forward xi hand-off of beam/part3d_comm.f03:278-314 + pack_particles('pipeline') :685-745:
 * pack particles with xi >= upper slab edge into dev_buf (dev_buf[0] = count, then 7 doubles each), remove them
 * ("fill the holes inversely"); unpack appends.  Buffer size = 1 + 7*cap doubles, cap = qpg_part3d_wire_cap()
 * (default 0.1*npmax like the reference's nbmax; qpg_part3d_set_wire_cap shrinks the message).  More than cap
 * crossings in one step raise an overflow flag that qpg_part3d_download reports as QPG_ERR_STATE (the reference
 * stops with a buffer-overflow error in the same situation). */
int qpg_part3d_pack_forward(qpg_part3d p, double *dev_buf);
int qpg_part3d_unpack(qpg_part3d p, const double *dev_buf);
long qpg_part3d_wire_cap(qpg_part3d p);
int qpg_part3d_set_wire_cap(qpg_part3d p, long cap);
/* Beam with spin (part3d%has_spin, init_part3d :117-135 with `amm` = anomalous magnetic moment; array s(3, npmax) :53-57): after
 * qpg_part3d_enable_spin every qpg_part3d_push also advances the spin vectors (push_spin_part3d :578-638, called from push_boris :425-427
 * BEFORE the new momentum is stored and from push_reduced :550-557 after it), qpg_part3d_update_bound moves them with the particles
 * (:668-670) and the hand-off record has 10 instead of 7 reals (part3d_comm.f03:683-694): the buffer of qpg_part3d_pack_forward / _unpack
 * holds qpg_part3d_wire_count doubles.  upload_spin / download_spin: s[npp][3] in the particle order of qpg_part3d_upload / _download. */
int qpg_part3d_enable_spin(qpg_part3d p, double amm);
int qpg_part3d_has_spin(qpg_part3d p);
int qpg_part3d_upload_spin(qpg_part3d p, const double *s, long npp);
int qpg_part3d_download_spin(qpg_part3d p, double *s, long *npp_out);
long qpg_part3d_wire_count(qpg_part3d p);

/* ------------------------------------------------------------------------------------------ */
/* fused fast path: the whole `do j = 1, nstep2d` body of simulation_class.f03:342-469 on the  */
/* device (one species, one beam set), incl. the predictor-corrector loop with a device-side   */
/* convergence test.  Same arithmetic as the per-routine entry points above.                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int nr, nz_total, noff2, nzp, max_mode, field_boundary, iter_max, sort_freq;
    double dr, dxi, dt, iter_reltol, iter_abstol, relax_fac;
    double sp_qbm;            /* species q/m */
    long sp_npmax;
    int beam_push_type, beam_evol;
    double beam_qbm;
    long beam_npmax;
    int use_graph;            /* capture the slice body in a CUDA graph */
    int sp_push_std;          /* 0 = robust pusher (the decks' choice), 1 = std flavour (amjdeposit_std, push_u_std, interp_psi);
                                 std runs on the per-slice launch paths, not in the persistent sweep kernel */
    int sp_push_pgc;          /* 1 = ponderomotive-guiding-centre flavour of the chosen pusher (param.f03 p_push2_std_pgc / p_push2_robust_pgc):
                                 the sim owns ONE laser envelope (qpg_sim_laser: its slab of nzp slices + guards) and runs simulation_class.f03:361-366
                                 / :401 per slice -- inside the persistent sweep kernel (robust_pgc) or on the per-slice launch paths; on a
                                 xi-pipeline the stages' envelopes are linked with qpg_laser_set_handoff */
    int laser_iter;           /* laser.iteration: fixed-point passes of the envelope solve per slice (>= 1) */
    double laser_k0;          /* laser.k0 */
    int sp_ppc_r;             /* species ppc(1): the on-axis correction of the susceptibility deposit (part2d_class.f03:2581) */
} qpg_sim_params;

int qpg_sim_create(qpg_sim *out, int device, void *cuda_stream, const qpg_sim_params *prm);
int qpg_sim_destroy(qpg_sim s);
qpg_ctx qpg_sim_ctx(qpg_sim s);
/* object access so the host (Fortran shim / tests) can reach the same objects through the per-routine API.
 * names: psi e b e_spe b_spe e_beam b_beam cu amu acu dcu q_spe q_beam spe_q spe_qn beam_q */
qpg_field qpg_sim_field(qpg_sim s, const char *name);
qpg_part2d qpg_sim_species(qpg_sim s);
qpg_part3d qpg_sim_beam(qpg_sim s);
/* species2d%new / %renew (species2d_class.f03:73,154): upload lattice, deposit, build qn */
int qpg_sim_init_species(qpg_sim s, const double *x, const double *pm, const double *gamma, const double *psi,
                         const double *q, long npp);
/* beam3d%qdp (beam3d_class.f03:193-221) split around its two pipeline calls:
 *   qdp_begin : this%q%as(0)                        [caller: unpack upstream guard slice into beam_q slice 1, add]
 *   qdp_end   : part%qdeposit(this%q)               [caller: pack beam_q slice nzp+1 and send it downstream]      */
int qpg_sim_beam_qdp_begin(qpg_sim s);
int qpg_sim_beam_qdp_end(qpg_sim s);
/* simulation_class.f03:299-331 minus the MPI calls: q_beam = beam_q, q_spe = 0, zero b e b_spe e_spe psi cu acu amu.
 * [caller, stage > 0: unpack cu and b_spe (slice 0) and the plasma particles received from upstream] */
int qpg_sim_begin_step(qpg_sim s);
/* the same two calls in halves for a pipeline stage that overlaps them with the wait for its upstream neighbour:
 *   qpg_sim_beam_qdp_raw   : scatter-add of the stage's own beam particles (part3d_class.f03:221-316)
 *   qpg_sim_beam_qdp_fix   : axis rules and 1/(j-1) (:318-351) -- after the upstream guard slice has been added to slice 1
 *   qpg_sim_begin_step_zero: the zero fills of simulation_class.f03:299-331   qpg_sim_begin_step_add: q_beam += beam_q
 * (qpg_sim_beam_qdp_end = raw + fix, qpg_sim_begin_step = zero + add) */
int qpg_sim_beam_qdp_raw(qpg_sim s);
int qpg_sim_beam_qdp_fix(qpg_sim s);
int qpg_sim_begin_step_zero(qpg_sim s);
int qpg_sim_begin_step_add(qpg_sim s);
/* run slices j0..j1 (1-based, inclusive) of this slab: simulation_class.f03:342-469 */
int qpg_sim_run_slices(qpg_sim s, int j0, int j1);
/* backward hand-off of the xi-pipeline (simulation_class.f03:460-467) without leaving the sweep: the next
 * qpg_sim_run_slices call that starts at slice 1 writes b and e of that slice in wire layout ([P][nr+2][3] each) into
 * wire_b / wire_e -- device pointers, possibly the upstream GPU's memory mapped with qpg_wire_import -- as soon as the
 * slice is complete, then sets *flag = seq (system-scope release).  The consumer orders its stream behind it with
 * qpg_stream_wait(flag, seq).  One-shot: applies to one first slice. */
int qpg_sim_set_back_handoff(qpg_sim s, double *wire_b, double *wire_e, unsigned *flag, unsigned seq);
/* simulation_class.f03:489-493: beam push + update_bound (E,B guard slice nzp+1 already unpacked by the caller) */
int qpg_sim_beam_push(qpg_sim s);
/* qpg_sim_beam_push in two halves for a stage of the xi-pipeline, with the NEXT step's raw beam deposit riding on them (the beam charge
 * volume must have been zeroed: qpg_sim_beam_qdp_begin): _interior advances and deposits the particles that do not gather from the guard
 * slice (needs nothing from the downstream stage); _edge advances and deposits the rest, then update_bound.  Particles that arrive from the
 * upstream stage afterwards (qpg_part3d_unpack) are deposited with qpg_sim_beam_qdp_part(s, 3) before qpg_part3d_pack_forward; the
 * deposit is completed by qpg_sim_beam_qdp_fix as usual (no qpg_sim_beam_qdp_raw in that step). */
int qpg_sim_beam_push_interior(qpg_sim s);
int qpg_sim_beam_push_edge(qpg_sim s);
int qpg_sim_beam_qdp_part(qpg_sim s, int part);
/* the sim's laser envelope (NULL unless sp_push_pgc) and simulation_class.f03:486 lasers%advance (after the slab) */
qpg_laser qpg_sim_laser(qpg_sim s);
int qpg_sim_laser_advance(qpg_sim s);
/* CUDA-graph replay of the slice body (use_graph): on = capture all iter_max predictor-corrector iterations (the ones behind the converged
 * one skip themselves on the device) instead of a WHILE node around one.  Launching a graph with a conditional node costs ~100 us of host
 * time against ~4 us for a plain one: worth it when many short slabs are enqueued from one host thread (a pipeline of >= 6 stages on the
 * per-slice launch path), not for one slab (each skipped iteration costs two near-empty kernels on the device). */
int qpg_sim_set_graph_unroll(qpg_sim s, int on);
/* on = the envelope solve of a 3D step may run beside the sweep kernel of the next one also when the caller chose the sweep grid
 * (qpg_sim_set_sweep_ctas n > 0): the caller vouches that an SM stays free for the solve's CTA (it cannot share one with a sweep CTA) */
int qpg_sim_set_laser_overlap(qpg_sim s, int on);
/* simulation_class.f03:498-501 species%renew from the device snapshot of the injected lattice */
int qpg_sim_renew(qpg_sim s);
/* counters since creation: particle-slice updates, PC iterations, slices (synchronises) */
int qpg_sim_stats(qpg_sim s, long *updates, long *pc_iters, long *slices);
/* switch the slice body between CUDA-graph replay (1) and plain stream launches (0, needed for per-kernel tprof) */
int qpg_sim_set_graph(qpg_sim s, int use_graph);
/* field programs: 1 = thread-block-cluster kernels (default when nr <= 1024 and max_mode <= 2), 0 = generic op-list CTA */
int qpg_sim_set_fused(qpg_sim s, int on);
/* slab driver: 1 = ONE persistent cooperative kernel sweeps all slices of a qpg_sim_run_slices call (default when
 * max_mode <= 2 and nr <= 4096; phases separated by grid barriers, predictor-corrector loop on the device),
 * 0 = per-slice launches (CUDA graph or plain stream, see qpg_sim_set_graph) */
int qpg_sim_set_sweep(qpg_sim s, int on);
/* CTAs of the sweep kernel: n > 0 absolute; n <= 0 = one per SM minus |n| (SMs left free for kernels on other streams,
 * e.g. the NCCL send/recv kernels of the xi-pipeline, which otherwise could not start until the sweep ends) */
int qpg_sim_set_sweep_ctas(qpg_sim s, int n);
/* in-kernel clocks of the sweep kernel since the last reset (synchronises): out12 = SM cycles spent in phase
 * [0] A||update_bound, [1] amjdeposit, [2] C, [3] push+qdeposit||D; [4] total cycles, [5] total ns (globaltimer),
 * [6] slices, [7] amjdeposit phases, [8..11] CTA 0's own work cycles in the four phases (the rest of a phase is
 * barrier latency + waiting for the slowest CTA).  out12 holds 12 doubles.
 * Fails with QPG_ERR_STATE if the last sweep hit its barrier watchdog. */
int qpg_sim_sweep_profile(qpg_sim s, double *out12, int reset);
/* per-slice record of the sweep kernel's LAST pass over each slice of the slab (synchronises): device time spent in the
 * slice (ns, %globaltimer) and the predictor-corrector iterations it took (the n_it of simulation_class.f03:379-407).
 * nzp entries each.  The cost profile along xi that the pipeline's slab partition is balanced with. */
int qpg_sim_slice_trace(qpg_sim s, double *ns_per_slice, int *iters_per_slice);
/* test hook: raises the sweep kernel's (sticky) watchdog word as a timed-out barrier would.  The next qpg_sim_run_slices leaves at
 * once, still releases a pending backward hand-off flag (qpg_sim_set_back_handoff) so that the upstream stage does not hang, and
 * qpg_sim_stats / qpg_ctx_sync / qpg_sim_sweep_profile return QPG_ERR_STATE from then on. */
int qpg_sim_debug_abort(qpg_sim s);

/* ------------------------------------------------------------------------------------------ */
/* peer-memory transport of the xi-pipeline between the GPUs of one box (one process per GPU).
 * Replaces the mpi_isend / mpi_recv pairs of part2d_class.f03:2355-2450 (pipesend/piperecv), field_class.f03:560-700
 * (pipe_send/pipe_recv) and part3d_comm.f03:278-314: the producer's qpg_*_pack kernels write the wire record directly
 * into the consumer GPU's buffer (mapped through a CUDA IPC handle), then raise a flag in the consumer's memory; the
 * consumer's stream waits for the flag with a stream memory operation.  All calls are asynchronous on `cuda_stream`.
 *   qpg_wire_alloc/free    : wire buffer or flag block that can be exported (plain cudaMalloc, zero-filled)
 *   qpg_wire_export/import : 64-byte IPC handle of a buffer / map a peer's buffer into this process
 *   qpg_stream_signal      : *flag = value after all work enqueued on the stream so far (flag may be peer memory)
 *   qpg_stream_wait        : the stream's later work starts once (int)(*flag - value) >= 0 (flag in OWN memory)     */
int qpg_wire_alloc(void **dev_ptr, long bytes);
int qpg_wire_free(void *dev_ptr);
int qpg_wire_export(void *dev_ptr, unsigned char *handle64);
int qpg_wire_import(const unsigned char *handle64, void **dev_ptr);
int qpg_wire_unmap(void *dev_ptr);
int qpg_stream_signal(void *cuda_stream, unsigned *flag, unsigned value);
int qpg_stream_wait(void *cuda_stream, unsigned *flag, unsigned value);
/* like qpg_stream_wait, but the stream waits only if the int at dev_count is non-zero when the wait is reached (a polling kernel of one
 * thread): the backward e / b hand-off of the xi-pipeline (simulation_class.f03:460-467, :482-483) matters only to a stage that holds
 * beam particles -- dev_count = qpg_part3d_count_ptr(beam) */
int qpg_stream_wait_unless_empty(void *cuda_stream, const int *dev_count, unsigned *flag, unsigned value);
/* device address of the live particle count of a beam particle set (an int the kernels of this library keep up to date) */
const int *qpg_part3d_count_ptr(qpg_part3d p);
int qpg_stream_wait_is_memop(void);   /* 1 = cuStreamWaitValue32, 0 = fallback polling kernel */

/* ------------------------------------------------------------------------------------------ */
/* laser envelope (ponderomotive guiding centre), laser/field_laser_class.f03 + sim_lasers_class.f03; one xi stage (nz = the slab).
 * Envelope volumes a_r, a_i on the host: C arrays [P][nz+3][nr+2], xi slice j (1-based) at index j+1 -- two lower guard
 * slices for the 3-point backward xi difference (gc_num(1,2) = 2), one upper; radial guards at 0 and nr+1.
 *   qpg_laser_create      : init_field_laser :104 + init_solver :269 (nr <= 1024)
 *   qpg_laser_upload      : the launched profile (profile_laser%launch :318 stays on the host) ; qpg_laser_download for diagnostics
 *   qpg_laser_slice(j)    : copy_slice(j, 2to1) + set_grad(j) :637 + gather :929 -> the four slice images qpg_laser_field(0..3)
 *                           that qpg_part2d_amjdeposit_pgc / qpg_part2d_push_u_pgc take (a_r, a_i dim 1; grad a_r, grad a_i dim 3);
 *                           j = -1 (here and in deposit_chi): the slice counter a qpg_sim keeps on the device (CUDA-graph replay)
 *   qpg_laser_deposit_chi : part2d%deposit_chi :361 of one species into chi (qpg_laser_field(4)), slice j of its volume
 *                           (sim_lasers_class.f03:175-195; j = 0: slice image only); ax_corr = 12 ppc_r^2 / (1 + 2 ppc_r^2)
 *   qpg_laser_advance     : set_rhs :393 + solve :752 with the chi volume (sim_lasers_class.f03:197-222)                    */
int qpg_laser_create(qpg_laser *out, qpg_ctx ctx, int nz, double k0, double ds, int iter);
int qpg_laser_destroy(qpg_laser l);
long qpg_laser_volume_size(qpg_laser l);
int qpg_laser_upload(qpg_laser l, const double *ar, const double *ai);
int qpg_laser_download(qpg_laser l, double *ar, double *ai);
qpg_field qpg_laser_field(qpg_laser l, int which);
int qpg_laser_slice(qpg_laser l, int j);
int qpg_laser_deposit_chi(qpg_laser l, qpg_part2d p, int j, double ax_corr);
int qpg_laser_advance(qpg_laser l);
/* The envelope on a xi-pipeline (sim_lasers_class.f03:216-218 pipe_recv / pipe_send 'forward' tag 'guard'; init_field_laser :163-169): the
 * slab's lower guard slices 0, -1 are the upstream stage's last two slices.  With a hand-off set, every qpg_laser_advance (n-th call = message
 * n on both links) runs  set_rhs (old guards) -> wait *in_ready >= n, copy guard_in into the guards, *in_ack = n -> solve -> wait *out_ack >=
 * n-1, write the own new slices nz, nz-1 to guard_out, *out_ready = n  on the stream of the advance.  guard_in / in_ready / out_ack live in
 * THIS GPU's memory, guard_out / out_ready / in_ack may be peer memory (qpg_wire_map); a record has qpg_laser_guard_size doubles.  NULL
 * triples switch a link off (first / last stage).  The caller enqueues the upstream stage's advance of a 3D step before the downstream one's. */
int qpg_laser_sync(qpg_laser l);   /* host waits for everything enqueued for this envelope, an overlapped advance on its side stream included */
long qpg_laser_guard_size(qpg_laser l);
int qpg_laser_set_handoff(qpg_laser l, const double *guard_in, unsigned *in_ready, unsigned *in_ack, double *guard_out, unsigned *out_ready, unsigned *out_ack);

/* ------------------------------------------------------------------------------------------ */
/* field-ionisation (ADK) neutral species, species/neutral_class.f03 -- NOT YET VALIDATED ON A GPU (written at the end of
 * round 1 against oracle/qpad_oracle_neutral.c; tests/test_gpu_neutral.py, enabled with QPG_TEST_NEUTRAL=1).
 * A neutral owns the ionisation levels per (radial cell, theta sector); the released electrons and the ions' position
 * buffer are two ordinary qpg_part2d sets (`part`, `part_add` of the reference) that go through the part2d entry points:
 *   neut%qdp(q)         = qpg_part2d_qdeposit(electrons, q)                     (:880)
 *   neut%ion_deposit(q) = qpg_part2d_qdeposit(ions, rho_ion_add) + field adds   (:904)
 *   neut%amjdp / push_u / push_x = the part2d calls on `electrons`              (:932-1016)
 *   neut%update(e, ..)  = qpg_neutral_update(n, e, electrons, ions)             (:576: ionize :600 + add_particles :755)
 *   neut%renew          = qpg_neutral_reset(n) + qpg_part2d_clear(electrons) + qpg_part2d_clear(ions)   (:839)
 * element = atomic number (1 H, 2 He, 3 Li), n0 = plasma density [cm^-3] (omega_p of sim_plasma_class.f03:84), dt_xi = dxi. */
int qpg_neutral_create(qpg_neutral *out, qpg_ctx ctx, int element, int ion_max, int ppc1, int ppc2, int num_theta, double q, double m,
                       double density, double n0, double dt_xi);
int qpg_neutral_destroy(qpg_neutral n);
int qpg_neutral_reset(qpg_neutral n);
int qpg_neutral_multi_max(qpg_neutral n);
int qpg_neutral_update(qpg_neutral n, qpg_field e, qpg_part2d electrons, qpg_part2d ions);
int qpg_neutral_levels(qpg_neutral n, double *host);      /* [(multi_max + 2)][num_theta][nr], synchronises */
int qpg_part2d_clear(qpg_part2d p);                        /* npp = 0 on the device */
/* The neutral species inside the fast path: after this call qpg_sim_run_slices runs the hooks of simulation_class.f03:351-354
 * (neut%qdp, neut%ion_deposit), :386-388 (neut%amjdp), :404-407 (neut%cbq) and :444-450 (neut%update, push_u, push_x) per slice and
 * qpg_sim_renew the neutral's renewal (:504-510), on the per-slice launch paths (CUDA-graph replay or plain stream; the persistent
 * sweep kernel and the cluster programs are switched off).  `electrons` / `ions` are created on qpg_sim_ctx(sim) with npmax >=
 * nr * num_theta * ppc1 * ppc2.  Extra fields of qpg_sim_field: "neut_q" (the electrons' charge volume), "rho_ion".
 * Parity with the oracle's ionisation loop on the GPU and in host emulation (tests/test_gpu_neutral.py, tests/test_emu_parity.py). */
int qpg_sim_attach_neutral(qpg_sim sim, qpg_neutral n, qpg_part2d electrons, qpg_part2d ions);
/* neut%psend / precv (neutral_class.f03:1025-1101) for a sim that is one slab of a xi-pipeline: after the slab's last slice the released
 * electrons, the ions' position buffer, the rho_ion image and the ionisation levels go to the next stage in one wire record of
 * qpg_sim_neutral_wire_count doubles (device buffer, possibly peer memory); the next stage unpacks it after its qpg_sim_renew and before
 * its first slice.  Stream-ordered like qpg_part2d_pack / _unpack. */
long qpg_sim_neutral_wire_count(qpg_sim sim);
int qpg_sim_neutral_pack(qpg_sim sim, double *dev_buf);
int qpg_sim_neutral_unpack(qpg_sim sim, const double *dev_buf);
/* The sub-cycling variant inside the fast path (proj_subcyc/simulation_subcyc_class.f03:216-376; input-deck keys
 * expansion_fac_max, expansion_fac_clamped, dt_min of simulation_subcyc): per slice the largest expansion factor of the plasma (and
 * of an attached neutral's electrons) chooses n_subcyc, the slice body is repeated with dxi / n_subcyc, pushed particles are
 * clamped.  Plain per-slice launches with one host synchronisation per slice (the reference's allreduce); no graph, no sweep kernel.
 * qpg_sim_subcycles = sub-steps taken so far.  Parity with the oracle on the GPU and in host emulation (tests/test_gpu_extras.py). */
int qpg_sim_set_subcyc(qpg_sim sim, int on, double exp_fac_max, double exp_fac_clamped, double dt_min);
long qpg_sim_subcycles(qpg_sim sim);

/* ------------------------------------------------------------------------------------------ */
/* The three groups below are held against the oracle on the GPU by tests/test_gpu_extras.py and on the CPU
 * through the host emulation of tests/emu (tests/test_emu_kernels.py).
 *
 * Sub-cycling / clamp variant of the slice loop (proj_subcyc/):
 *   part2d_subcyc%get_exp_fac_max (part2d_subcyc_class.f03:28)  = qpg_part2d_exp_fac_max : max gamma / (gamma - p_z), 1 if empty;
 *                                                                  synchronises (the host chooses the number of sub-steps)
 *   part2d_subcyc%clamp_exp_fac   (:48)                          = qpg_part2d_clamp_exp_fac
 *   the sub-step rule of simulation_subcyc_class.f03:431-451     = qpg_subcyc_step (host arithmetic only)
 * The sub-cycled slice body (:216-376) is the standard per-routine sequence with dt = dxi / n_subcyc. */
int qpg_part2d_exp_fac_max(qpg_part2d p, double *exp_fac_max);
int qpg_part2d_clamp_exp_fac(qpg_part2d p, double exp_fac_clamped);
int qpg_subcyc_step(double exp_fac, double exp_fac_max, double dt, double dt_min, double *dt_subcyc, int *n_subcyc);

/* Vector-potential diagnostics (fields/field_vpot_class.f03): field_vpot%solve_vpotz(jay) :354 and %solve_vpott(jay) :392.
 * cu and vpot are dim-3 fields; vpotz writes component 3 (A_z), vpott components 1, 2 (A_r, A_phi) of the slice image. */
int qpg_solve_vpotz(qpg_ctx ctx, qpg_field cu, qpg_field vpot);
int qpg_solve_vpott(qpg_ctx ctx, qpg_field cu, qpg_field vpot);
int qpg_vpot_release(qpg_ctx ctx);                         /* frees the cached operators of a context (before qpg_ctx_destroy) */

/* Device-resident staging of diagnostics (replaces the host arrays handed to hdf5io_class.f03 pwfield_pipe :591,
 * pwpart_2d_r :1027, pwpart_3d_pipe :1220): a re-layout kernel on the context's stream, then the device-to-host copy into a
 * pinned buffer on the stage's own copy stream, so the next 3D step overlaps the transfer.  One transfer in flight per stage
 * (QPG_ERR_STATE otherwise); use two stages to double-buffer.  Layouts (fp64):
 *   field  : [plane][comp][slice 1..nzp][node 1..nr]  -- one (nzp, nr) dataset per plane and component (f2(dim, 1:nr, 1:nzp))
 *   part2d : [0] = tnpp = int(npp / dspl), then datasets x1 x2 p1 p2 p3 q of `stride` entries (first tnpp valid)
 *   part3d : [0] = tnpp, then datasets x1 x2 x3+z0 p1 p2 p3 q ; particle i of a dataset = particle 1 + i dspl of the set */
int qpg_stage_create(qpg_stage *out, qpg_ctx ctx, long capacity_doubles);
int qpg_stage_destroy(qpg_stage s);
int qpg_stage_field(qpg_stage s, qpg_field f, long *count);
int qpg_stage_part2d(qpg_stage s, qpg_part2d p, int dspl, long *stride);
int qpg_stage_part3d(qpg_stage s, qpg_part3d p, int dspl, double z0, long *stride);
int qpg_stage_wait(qpg_stage s, const double **host, long *count);   /* blocks the host until the copy has landed */

/* Accuracy probe (not part of the reference's interface): the MUFU-seeded reciprocal and square root of the momentum arithmetic
 * evaluated on the device for n host values; tests/test_gpu_extras.py::test_fastmath_accuracy bounds their error in ulp. */
int qpg_debug_fastmath(qpg_ctx ctx, long n, const double *host_x, double *host_rcp, double *host_sqrt);

#ifdef __cplusplus
}
#endif
#endif

// sim.cu -- fused fast path: the `do j = 1, nstep2d` body of simulation_class.f03:342-469 enqueued without host
// round trips.  Per slice the device runs
//     qdeposit -> A -> while(!done){ amjdeposit -> C } -> D -> push(+bound flags) -> compact [-> sort]
// where A, C, D are single-CTA field programs (fields.cu) and the predictor-corrector loop is a CUDA-graph WHILE
// node whose condition the C program sets on the device (convergence_tester, simulation_class.f03:522-606).
// Without graphs (use_graph = 0) the iter_max iterations are enqueued back to back and skip themselves on `done`.
#include "common.cuh"

int part2d_launch_qdeposit(qpg_part2d p);
int part2d_launch_amjdeposit(qpg_part2d p, qpg_field ef, qpg_field bf, double dt, const int *skip_flag, int std_flavour);
int part2d_launch_push(qpg_part2d p, qpg_field ef, qpg_field bf, double dt, int mode);
int part2d_launch_compact(qpg_part2d p, int *slice_flags);
int part2d_launch_amjdeposit_pgc(qpg_part2d p, qpg_field ef, qpg_field bf, qpg_field ar, qpg_field ai, qpg_field arg, qpg_field aig, double dt, const int *skip_flag, int std_flavour);

struct qpg_sim_s {
    qpg_sim_params prm;
    qpg_ctx ctx;
    qpg_field psi, e, b, e_spe, b_spe, e_beam, b_beam, cu, amu, acu, dcu, q_spe, q_beam, spe_q, spe_qn, spe_cu, spe_dcu, spe_amu, beam_q;
    qpg_part2d spe;
    qpg_part3d beam;
    qpg_laser laser;      // sp_push_pgc: the one laser envelope of the run
    bool split_deposit;   // the pipeline's split beam push is in use: the raw beam deposit rides on it (qpg_sim_beam_push_interior / _edge, qpg_sim_beam_qdp_part)
    unsigned *las_progress; unsigned las_base; bool las_overlap, las_overlap_req;   // overlapped envelope advance (qpg_sim_laser_advance): progress word of the running solve
    int cur_j;            // slice being enqueued (per-slice launch path)
    cudaGraph_t graph;
    cudaGraphExec_t gexec;
    bool graph_ready;
    bool graph_unroll;    // slice graph without the WHILE node (see build_graph)
    bool use_fused;       // cluster kernels of fused.cu instead of the op-list programs
    bool use_sweep;       // persistent cooperative slab-sweep kernel (sweep.cu)
    int sweep_grid;       // CTAs of the sweep kernel (0 = not yet queried)
    int sweep_ctas_req;   // requested CTA count: > 0 absolute, <= 0 = number of SMs + this (SMs left to other streams)
    unsigned *sw_bar;     // grid / team barrier counters + abort flag
    double *sw_xbuf;      // team exchange records
    void *sw_xll;         // flagged exchange words of the strip scans
    long long *sw_prof;   // in-kernel phase clocks
    double *back_b, *back_e; unsigned *back_flag; unsigned back_seq;   // one-shot backward hand-off of the next first slice (qpg_sim_set_back_handoff)
    long long *sw_trace;  // per-slice time and PC iteration count of the last sweep over each slice [nzp][2]
    double *phi;
    long host_updates, host_iters, host_slices;
    // field-ionisation neutral species attached with qpg_sim_attach_neutral (not owned): its released electrons and the position
    // buffer of the ions are two more particle sets of the per-slice launch paths (simulation_class.f03:351-354, :386-388, :444-450)
    // sub-cycling variant (qpg_sim_set_subcyc; proj_subcyc/simulation_subcyc_class.f03:216-376): plain per-slice launches only
    bool subcyc;
    double sc_exp_fac_max, sc_exp_fac_clamped, sc_dt_min;
    double cur_dt;        // xi step of the deposits / pushes being enqueued (dxi, or dxi / n_subcyc)
    long host_subcycles;
    qpg_neutral neut;
    qpg_part2d neut_e, neut_i;
    qpg_field neut_q, neut_cu, neut_dcu, neut_amu, rho_ion, rho_ion_add;
};

// flags: [0] done  [1] PC iterations (since last read)  [2] iteration inside the slice  [3] current slice j
//        [4] slices executed  ;  counters (long long) live in conv_out[4..] reinterpret: updates
static void prog_A(qpg_sim s, FProgBuilder &pb, int count_updates = 1)
{
    FOp *o;
    o = &pb.add(FOP_SLICE_2TO1); o->a = s->q_beam->f1; o->b = s->q_beam->f2; o->da = 1; o->i0 = -1;          // :344
    o = &pb.add(FOP_BT); o->a = s->q_beam->f1; o->b = s->b_beam->f1; o->da = 1;                               // :345
    o = &pb.add(FOP_ZERO); o->a = s->spe_q->f1; o->da = 1;                                                   // species2d qdp :198
    o = &pb.add(FOP_QFIX); o->a = s->spe->acc1; o->b = s->spe_q->f1; o->da = 1; o->c = (double *)s->spe->d_npp; o->i1 = count_updates;
    o = &pb.add(FOP_ADD3); o->a = s->spe_q->f1; o->b = s->spe_qn->f1; o->c = s->q_spe->f1; o->da = 1;        // q_spe = 0 + q + qn
    if (s->neut) {
        o = &pb.add(FOP_ZERO); o->a = s->neut_q->f1; o->da = 1;                                              // neut%qdp, neutral_class.f03:880
        o = &pb.add(FOP_QFIX); o->a = s->neut_e->acc1; o->b = s->neut_q->f1; o->da = 1; o->c = (double *)s->neut_e->d_npp; o->i1 = count_updates;
        o = &pb.add(FOP_ADD); o->a = s->neut_q->f1; o->b = s->q_spe->f1; o->da = 1;
        o = &pb.add(FOP_ZERO); o->a = s->rho_ion_add->f1; o->da = 1;                                         // neut%ion_deposit :904-930
        o = &pb.add(FOP_QFIX); o->a = s->neut_i->acc1; o->b = s->rho_ion_add->f1; o->da = 1; o->i1 = 0;
        o = &pb.add(FOP_ADD); o->a = s->rho_ion_add->f1; o->b = s->rho_ion->f1; o->da = 1;
        o = &pb.add(FOP_ADD); o->a = s->rho_ion->f1; o->b = s->q_spe->f1; o->da = 1;
    }
    o = &pb.add(FOP_PSI); o->a = s->q_spe->f1; o->b = s->psi->f1; o->da = 1;                                  // :356
    o = &pb.add(FOP_BZ); o->a = s->cu->f1; o->b = s->b_spe->f1; o->da = 3;                                    // :360
    o = &pb.add(FOP_PC_BEGIN);
    o = &pb.add(FOP_CONV_RECORD); o->a = s->b_spe->f1; o->da = 3; o->i0 = 1;                                  // :373
    o = &pb.add(FOP_ADD3); o->a = s->b_spe->f1; o->b = s->b_beam->f1; o->c = s->b->f1; o->da = 3;             // :375
    o = &pb.add(FOP_EZ); o->a = s->cu->f1; o->b = s->e->f1; o->da = 3;                                        // :376
    o = &pb.add(FOP_ET); o->a = s->b->f1; o->b = s->psi->f1; o->c = s->e->f1; o->da = 3;                      // :377
}
static void prog_C(qpg_sim s, FProgBuilder &pb)
{
    FOp *o;
    const int F = FOPF_SKIP_IF_DONE;
    o = &pb.add(FOP_ZERO); o->a = s->spe_cu->f1; o->da = 3; o->flags = F;                                     // species2d amjdp :250-252
    o = &pb.add(FOP_ZERO); o->a = s->spe_dcu->f1; o->da = 2; o->flags = F;
    o = &pb.add(FOP_ZERO); o->a = s->spe_amu->f1; o->da = 3; o->flags = F;
    o = &pb.add(FOP_AMJFIX); o->a = s->spe->acc8; o->b = s->spe_cu->f1; o->c = s->spe_dcu->f1; o->d = s->spe_amu->f1; o->da = 1; o->flags = F;
    o = &pb.add(FOP_COPY); o->a = s->spe_cu->f1; o->b = s->cu->f1; o->da = 3; o->flags = F;                   // cu = 0 + spe%cu :378,:274
    o = &pb.add(FOP_COPY); o->a = s->spe_dcu->f1; o->b = s->acu->f1; o->da = 2; o->flags = F;
    o = &pb.add(FOP_COPY); o->a = s->spe_amu->f1; o->b = s->amu->f1; o->da = 3; o->flags = F;
    if (s->neut) {                                                                                           // neut%amjdp :932, simulation_class.f03:386-388
        o = &pb.add(FOP_ZERO); o->a = s->neut_cu->f1; o->da = 3; o->flags = F;
        o = &pb.add(FOP_ZERO); o->a = s->neut_dcu->f1; o->da = 2; o->flags = F;
        o = &pb.add(FOP_ZERO); o->a = s->neut_amu->f1; o->da = 3; o->flags = F;
        o = &pb.add(FOP_AMJFIX); o->a = s->neut_e->acc8; o->b = s->neut_cu->f1; o->c = s->neut_dcu->f1; o->d = s->neut_amu->f1; o->da = 1; o->flags = F;
        o = &pb.add(FOP_ADD); o->a = s->neut_cu->f1; o->b = s->cu->f1; o->da = 3; o->flags = F;
        o = &pb.add(FOP_ADD); o->a = s->neut_dcu->f1; o->b = s->acu->f1; o->da = 2; o->flags = F;
        o = &pb.add(FOP_ADD); o->a = s->neut_amu->f1; o->b = s->amu->f1; o->da = 3; o->flags = F;
    }
    o = &pb.add(FOP_DJDXI); o->a = s->acu->f1; o->b = s->amu->f1; o->c = s->dcu->f1; o->da = 2; o->flags = F;  // :390
    o = &pb.add(FOP_BTITER); o->a = s->dcu->f1; o->b = s->cu->f1; o->c = s->b_spe->f1; o->da = 2; o->s0 = s->ctx->relax; o->flags = F; // :391
    o = &pb.add(FOP_BZ); o->a = s->cu->f1; o->b = s->b_spe->f1; o->da = 3; o->flags = F;                      // :392
    o = &pb.add(FOP_CONV_COMPARE); o->a = s->b_spe->f1; o->da = 3; o->i0 = 1; o->i1 = 1; o->i2 = s->prm.iter_max;
    o->s0 = s->prm.iter_reltol; o->s1 = s->prm.iter_abstol; o->flags = F;                                     // :395-396
    o = &pb.add(FOP_CONV_RECORD); o->a = s->b_spe->f1; o->da = 3; o->i0 = 1; o->flags = F;
    o = &pb.add(FOP_ADD3); o->a = s->b_spe->f1; o->b = s->b_beam->f1; o->c = s->b->f1; o->da = 3; o->flags = F; // :375 / :413
    o = &pb.add(FOP_EZ); o->a = s->cu->f1; o->b = s->e->f1; o->da = 3; o->flags = F;                          // :376 / :415
    o = &pb.add(FOP_ET); o->a = s->b->f1; o->b = s->psi->f1; o->c = s->e->f1; o->da = 3; o->flags = F;        // :377 / :416
}
static void prog_D(qpg_sim s, FProgBuilder &pb)
{
    FOp *o;
    o = &pb.add(FOP_ADD_DIM); o->a = s->spe_cu->f1; o->b = s->spe_q->f1; o->da = 3; o->db = 1; o->i0 = 2; o->i1 = 0;  // cbq, species2d :396
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->spe_q->f1; o->b = s->spe_q->f2; o->da = 1; o->i0 = -1;
    if (s->neut) {                                                                                             // neut%cbq :404-407
        o = &pb.add(FOP_ADD_DIM); o->a = s->neut_cu->f1; o->b = s->neut_q->f1; o->da = 3; o->db = 1; o->i0 = 2; o->i1 = 0;
        o = &pb.add(FOP_SLICE_1TO2); o->a = s->neut_q->f1; o->b = s->neut_q->f2; o->da = 1; o->i0 = -1;
        o = &pb.add(FOP_SLICE_1TO2); o->a = s->rho_ion->f1; o->b = s->rho_ion->f2; o->da = 1; o->i0 = -1;
    }
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->cu->f1; o->b = s->cu->f2; o->da = 3; o->i0 = -1;                    // :409
    o = &pb.add(FOP_ADD_DIM); o->a = s->cu->f1; o->b = s->q_spe->f1; o->da = 3; o->db = 1; o->i0 = 2; o->i1 = 0;  // :410
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->q_spe->f1; o->b = s->q_spe->f2; o->da = 1; o->i0 = -1;              // :411
    o = &pb.add(FOP_ET); o->a = s->b_spe->f1; o->b = s->psi->f1; o->c = s->e_spe->f1; o->da = 3;               // :414
    o = &pb.add(FOP_SCALE); o->a = s->dcu->f1; o->da = 2; o->s0 = s->prm.dxi;                                  // :425
    o = &pb.add(FOP_ADD_DIM); o->a = s->dcu->f1; o->b = s->cu->f1; o->da = 2; o->db = 3; o->i0 = 0; o->i1 = 0;  // :426
    o = &pb.add(FOP_ADD_DIM); o->a = s->dcu->f1; o->b = s->cu->f1; o->da = 2; o->db = 3; o->i0 = 1; o->i1 = 1;
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->e->f1; o->b = s->e->f2; o->da = 3; o->i0 = -1;                      // :452-456
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->b->f1; o->b = s->b->f2; o->da = 3; o->i0 = -1;
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->psi->f1; o->b = s->psi->f2; o->da = 1; o->i0 = -1;
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->b_spe->f1; o->b = s->b_spe->f2; o->da = 3; o->i0 = -1;
    o = &pb.add(FOP_SLICE_1TO2); o->a = s->e_spe->f1; o->b = s->e_spe->f2; o->da = 3; o->i0 = -1;
    o = &pb.add(FOP_SET_FLAG); o->i0 = 3; o->i1 = -1;  // flags[3] += 1 (next slice), flags[4] += 1
}

static FusedArgs fused_args(qpg_sim s)
{
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    qpg_ctx c = s->ctx;
    a.nr = c->nr; a.iter_max = s->prm.iter_max; a.dr = c->dr; a.dxi = s->prm.dxi; a.relax = c->relax;
    a.reltol = s->prm.iter_reltol; a.abstol = s->prm.iter_abstol;
    a.psi = s->psi->f1; a.e = s->e->f1; a.b = s->b->f1; a.e_spe = s->e_spe->f1; a.b_spe = s->b_spe->f1; a.b_beam = s->b_beam->f1;
    a.cu = s->cu->f1; a.amu = s->amu->f1; a.acu = s->acu->f1; a.dcu = s->dcu->f1; a.q_spe = s->q_spe->f1; a.q_beam = s->q_beam->f1;
    a.spe_q = s->spe_q->f1; a.spe_qn = s->spe_qn->f1; a.spe_cu = s->spe_cu->f1; a.spe_dcu = s->spe_dcu->f1; a.spe_amu = s->spe_amu->f1;
    a.q_beam2 = s->q_beam->f2; a.spe_q2 = s->spe_q->f2; a.cu2 = s->cu->f2; a.q_spe2 = s->q_spe->f2; a.e2 = s->e->f2; a.b2 = s->b->f2;
    a.psi2 = s->psi->f2; a.b_spe2 = s->b_spe->f2; a.e_spe2 = s->e_spe->f2;
    a.acc1 = s->spe->acc1; a.acc8 = s->spe->acc8; a.phi = s->phi; a.d_npp = s->spe->d_npp;
    a.ops = qpg_ctx_dev_ops(c); a.conv_old = c->conv_old; a.conv_out = c->conv_out; a.flags = c->flags; a.counters = c->counters;
    a.cond_handle = c->cond_handle;
    return a;
}
template <int M> static void l_fused(int which, cudaStream_t st, const FusedArgs &a)
{
    const int grid = FC;
    if (which == 0) k_fused_A<M><<<grid, FT, 0, st>>>(a);
    else if (which == 1) k_fused_C<M><<<grid, FT, 0, st>>>(a);
    else k_fused_D<M><<<((a.nr + 2) * (2 * M + 1) + FT - 1) / FT, FT, 0, st>>>(a);
}
static int launch_fused(qpg_sim s, int which)
{
    qpg_ctx c = s->ctx;
    FusedArgs a = fused_args(s);
    TprofScope tp(c, TP_FIELD_FUSED);
    switch (c->M) {
    case 0: l_fused<0>(which, c->stream, a); break;
    case 1: l_fused<1>(which, c->stream, a); break;
    default: l_fused<2>(which, c->stream, a); break;
    }
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_pc_iteration(qpg_sim s)
{
    int rc;
    if (s->prm.sp_push_pgc) {    // species2d_class.f03:256-259 amjdeposit_{std,robust}_pgc with laser_all's slice images
        qpg_laser l = s->laser;
        rc = part2d_launch_amjdeposit_pgc(s->spe, s->e, s->b, qpg_laser_field(l, 0), qpg_laser_field(l, 1), qpg_laser_field(l, 2), qpg_laser_field(l, 3),
                                          s->cur_dt, s->ctx->flags, s->prm.sp_push_std != 0);
    } else rc = part2d_launch_amjdeposit(s->spe, s->e, s->b, s->cur_dt, s->ctx->flags, s->prm.sp_push_std != 0);
    if (rc) return rc;
    if (s->neut && (rc = part2d_launch_amjdeposit(s->neut_e, s->e, s->b, s->cur_dt, s->ctx->flags, 0))) return rc;   // robust pusher (:958)
    if (s->use_fused) return launch_fused(s, 1);
    FProgBuilder pb(s->ctx);
    prog_C(s, pb);
    return pb.launch(TP_FIELD_FUSED);
}
static int enqueue_slice_head(qpg_sim s)
{
    int rc;
    if (s->use_fused) rc = launch_fused(s, 0);
    else { FProgBuilder pb(s->ctx); prog_A(s, pb); rc = pb.launch(TP_FIELD_FUSED); }
    if (rc) return rc;
    // simulation_class.f03:357-359: the std pushers read psi at the particle positions (program A has just solved psi;
    // nothing else in A depends on the particles' psi)
    if (s->prm.sp_push_std) rc = qpg_part2d_interp_psi(s->spe, s->psi);
    if (!rc && s->prm.sp_push_pgc) rc = qpg_laser_slice(s->laser, -1);    // simulation_class.f03:361-366, slice j = the device-side counter
    return rc;
}
static int enqueue_slice_tail(qpg_sim s)
{
    int rc;
    if (s->prm.sp_push_pgc) {   // simulation_class.f03:401 lasers%deposit_chi (psi of the particles is the converged iteration's)
        const int ppc = s->prm.sp_ppc_r > 0 ? s->prm.sp_ppc_r : 1;
        if ((rc = qpg_laser_deposit_chi(s->laser, s->spe, -1, (12.0 * ppc * ppc) / (1.0 + 2.0 * ppc * ppc)))) return rc;
    }
    if (s->use_fused) rc = launch_fused(s, 2);
    else { FProgBuilder pb(s->ctx); prog_D(s, pb); rc = pb.launch(TP_FIELD_FUSED); }
    if (rc) return rc;
    if (s->prm.sp_push_pgc) {   // push_u_{std,robust}_pgc (:438), then push_x + bound flags (:439)
        qpg_laser l = s->laser;
        rc = qpg_part2d_push_u_pgc(s->spe, s->prm.sp_push_std ? QPG_PUSH2_STD_PGC : QPG_PUSH2_ROBUST_PGC, s->e, s->b, qpg_laser_field(l, 0), qpg_laser_field(l, 1),
                                   qpg_laser_field(l, 2), qpg_laser_field(l, 3), s->prm.dxi);
        if (!rc) rc = part2d_launch_push(s->spe, s->e, s->b, s->prm.dxi, 6);
    } else rc = part2d_launch_push(s->spe, s->e, s->b, s->prm.dxi, s->prm.sp_push_std ? 7 | 8 : 7);  // push_u + push_x + bound flags :438-439
    if (rc) return rc;
    rc = part2d_launch_compact(s->spe, s->use_fused ? s->ctx->flags : nullptr);  // update_bound (+ slice counter)
    if (rc) return rc;
    if ((rc = part2d_launch_qdeposit(s->spe))) return rc;        // next slice's qdp (:346-349) on the advanced particles
    if (s->neut) {
        // simulation_class.f03:444-450: ionise with this slice's E and create the released electrons (they are pushed in the same
        // slice), push the neutral's electrons; then the look-ahead deposits of the next slice's neut%qdp and neut%ion_deposit
        if ((rc = qpg_neutral_update(s->neut, s->e, s->neut_e, s->neut_i))) return rc;
        if ((rc = part2d_launch_push(s->neut_e, s->e, s->b, s->prm.dxi, 7))) return rc;
        if ((rc = part2d_launch_compact(s->neut_e, nullptr))) return rc;
        if ((rc = part2d_launch_qdeposit(s->neut_e))) return rc;
        if ((rc = part2d_launch_qdeposit(s->neut_i))) return rc;
    }
    return 0;
}

// ---- sub-cycling variant of the slice body (proj_subcyc/simulation_subcyc_class.f03:216-376) -------------------------------
// The largest expansion factor gamma / (gamma - p_z) of the plasma decides the number of sub-steps (:229-236, :431-451: one host
// synchronisation per slice, as the reference's mpi_allreduce); the deposit / solve / predictor-corrector / push sequence is
// repeated with dxi / n_subcyc, pushed particles are clamped (:298-309), the slice is stored once (:330-362, the full dxi).
int qpg_part2d_exp_fac_max(qpg_part2d p, double *exp_fac_max);
int qpg_part2d_clamp_exp_fac(qpg_part2d p, double exp_fac_clamped);
int qpg_subcyc_step(double exp_fac, double exp_fac_max, double dt, double dt_min, double *dt_subcyc, int *n_subcyc);
static int enqueue_slice_subcyc(qpg_sim s)
{
    int rc, n_sub = 1;
    double fac = 1.0, f, dt_sub = s->prm.dxi;
    if ((rc = qpg_part2d_exp_fac_max(s->spe, &f))) return rc;
    if (f > fac) fac = f;
    if (s->neut) { if ((rc = qpg_part2d_exp_fac_max(s->neut_e, &f))) return rc; if (f > fac) fac = f; }
    if ((rc = qpg_subcyc_step(fac, s->sc_exp_fac_max, s->prm.dxi, s->sc_dt_min, &dt_sub, &n_sub))) return rc;
    s->host_subcycles += n_sub;
    s->cur_dt = dt_sub;
    for (int isub = 0; isub < n_sub && !rc; isub++) {
        { FProgBuilder pb(s->ctx); prog_A(s, pb, isub == 0); if ((rc = pb.launch(TP_FIELD_FUSED))) break; }
        for (int l = 0; l < s->prm.iter_max; l++) if ((rc = enqueue_pc_iteration(s))) break;
        if (rc) break;
        qpg_part2d sets[2] = {s->spe, s->neut ? s->neut_e : nullptr};
        for (qpg_part2d p : sets) {
            if (!p) continue;
            if (p == s->neut_e && (rc = qpg_neutral_update(s->neut, s->e, s->neut_e, s->neut_i))) break;   // :312-317, inside every sub-step
            if ((rc = part2d_launch_push(p, s->e, s->b, dt_sub, 1))) break;                                    // push_u
            if ((rc = qpg_part2d_clamp_exp_fac(p, s->sc_exp_fac_clamped))) break;
            if ((rc = part2d_launch_push(p, s->e, s->b, dt_sub, 6))) break;                                    // push_x + bound flags
            if ((rc = part2d_launch_compact(p, nullptr))) break;
            if ((rc = part2d_launch_qdeposit(p))) break;                                                       // the next sub-step's / slice's qdp
        }
        if (!rc && s->neut) rc = part2d_launch_qdeposit(s->neut_i);
    }
    s->cur_dt = s->prm.dxi;
    if (rc) return rc;
    FProgBuilder pb(s->ctx);
    prog_D(s, pb);
    return pb.launch(TP_FIELD_FUSED);
}

// ---- persistent slab sweep (sweep.cu) ---------------------------------------------------------------------
// robust pusher, with or without the ponderomotive-guiding-centre terms of a laser envelope (robust_pgc)
static bool sweep_supported(const qpg_sim_params &prm) { return prm.max_mode <= 2 && (prm.nr + ST_N - 1) / ST_N <= SW_MAX_TEAM && !prm.sp_push_std; }
template <int M> static constexpr size_t sweep_smem() { return (sizeof(StripSmem<M>) + 7) / 8 * 8 + sizeof(double) * DepTile<M>::doubles * (SW_T / 32); }
template <int M, bool PGC> static cudaError_t sweep_occupancy(int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(k_sweep<M, PGC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem<M>());
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_sweep<M, PGC>, SW_T, sweep_smem<M>());
}
template <int M, bool PGC> static cudaError_t sweep_launch(int grid, cudaStream_t st, SweepArgs &a)
{
    void *args[] = {(void *)&a};
    return cudaLaunchCooperativeKernel((const void *)k_sweep<M, PGC>, dim3(grid), dim3(SW_T), args, sweep_smem<M>(), st);
}
static int sweep_prepare(qpg_sim s)
{
    if (s->sweep_grid > 0) return 0;
    qpg_ctx c = s->ctx;
    int coop = 0, nsm = 0, per = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
    if (!coop) { qpg_set_error("device does not support cooperative launches"); return QPG_ERR_UNSUPPORTED; }
    const bool pgc = s->prm.sp_push_pgc != 0;
    switch (c->M) {
    case 0: CUDA_TRY(pgc ? (sweep_occupancy<0, true>(&per)) : (sweep_occupancy<0, false>(&per))); break;
    case 1: CUDA_TRY(pgc ? (sweep_occupancy<1, true>(&per)) : (sweep_occupancy<1, false>(&per))); break;
    default: CUDA_TRY(pgc ? (sweep_occupancy<2, true>(&per)) : (sweep_occupancy<2, false>(&per))); break;
    }
    const int nteam = (c->nr + ST_N - 1) / ST_N;
    if (per < 1 || nsm * per <= nteam) { qpg_set_error("sweep kernel: %d CTAs/SM x %d SMs cannot host a field team of %d", per, nsm, nteam); return QPG_ERR_UNSUPPORTED; }
    // the envelope solve of the previous step (one CTA that needs an SM of its own) runs beside the sweep: with the default grid one SM is left
    // to it; with a grid chosen by the caller (SM partitions of a pipeline) only if the caller vouches for the free SM (qpg_sim_set_laser_overlap)
    s->las_overlap = s->prm.sp_push_pgc && (s->sweep_ctas_req <= 0 || s->las_overlap_req) && !getenv("QPG_LASER_NO_OVERLAP");
    if (s->las_overlap && s->sweep_ctas_req == 0) s->sweep_ctas_req = -1;
    int g = s->sweep_ctas_req > 0 ? s->sweep_ctas_req : nsm * per + s->sweep_ctas_req;   // default: one CTA per SM
    if (g > nsm * per) g = nsm * per;
    if (g <= nteam) { qpg_set_error("sweep kernel: %d CTAs cannot host a field team of %d plus the update_bound CTA", g, nteam); return QPG_ERR_ARG; }
    // every check has passed: allocate (all or nothing, so that a failed attempt leaks nothing and can be repeated)
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, c->stream); } };
    alloc((void **)&s->sw_bar, sizeof(unsigned) * 128);                              // incl. the sticky abort word [64], never cleared again
    alloc((void **)&s->sw_xbuf, sizeof(double) * 3 * SW_MAX_TEAM * SW_XK);
    alloc((void **)&s->sw_xll, sizeof(uint4) * 2 * SW_MAX_TEAM * SW_XK);
    alloc((void **)&s->sw_trace, sizeof(long long) * 2 * s->prm.nzp);
    alloc((void **)&s->sw_prof, sizeof(long long) * 32);
    if (e != cudaSuccess) {
        cudaFree(s->sw_bar); cudaFree(s->sw_xbuf); cudaFree(s->sw_xll); cudaFree(s->sw_trace); cudaFree(s->sw_prof);
        s->sw_bar = nullptr; s->sw_xbuf = nullptr; s->sw_xll = nullptr; s->sw_trace = nullptr; s->sw_prof = nullptr;
        return qpg_cuda_fail(e, "sweep_prepare: cudaMalloc");
    }
    s->sweep_grid = g;
    return 0;
}
static double *const *part2d_plane_table(qpg_part2d p);
static int sweep_run(qpg_sim s, int j0, int j1)
{
    qpg_ctx c = s->ctx;
    int rc = sweep_prepare(s);
    if (rc) return rc;
    SweepArgs a;
    memset(&a, 0, sizeof(a));
    a.f = fused_args(s);
    qpg_part2d p = s->spe;
    a.pv = view_of(p);
    a.planes = part2d_plane_table(p);
    a.d_npp_w = p->d_npp; a.d_nout = p->d_nout; a.outmask = p->outmask; a.lists = p->lists;
    a.qbm = p->qbm; a.edge = (double)c->nr * c->dr;
    a.j0 = j0; a.j1 = j1; a.nteam = (c->nr + ST_N - 1) / ST_N;
    if (j0 == 1 && s->back_flag) {   // the launch that sweeps the slab's first slice publishes it (one-shot)
        a.back_b = s->back_b; a.back_e = s->back_e; a.back_flag = s->back_flag; a.back_seq = s->back_seq;
        s->back_flag = nullptr;
    }
    a.bar = s->sw_bar; a.xbuf = s->sw_xbuf; a.xll = (uint4 *)s->sw_xll; a.prof = s->sw_prof; a.trace = s->sw_trace;
    CUDA_TRY(cudaMemsetAsync(s->sw_bar, 0, sizeof(unsigned) * 64, c->stream));    // barrier counters only: the abort word [64] is sticky
    CUDA_TRY(cudaMemsetAsync(s->sw_xll, 0, sizeof(uint4) * 2 * SW_MAX_TEAM * SW_XK, c->stream));   // sequence numbers restart at 1 every launch
    TprofScope tp(c, TP_K_SWEEP);
    cudaError_t e;
    const bool pgc = s->prm.sp_push_pgc != 0;
    if (pgc) {   // the laser hooks of the slice loop run inside the kernel: slice images, pgc pushers, susceptibility deposit
        qpg_laser l = s->laser;
        a.lv = LaserView{l->f_ar->f1, l->f_ai->f1, l->f_gr->f1, l->f_gi->f1};
        a.las_far = l->f_ar->f1; a.las_fai = l->f_ai->f1; a.las_fgr = l->f_gr->f1; a.las_fgi = l->f_gi->f1;
        a.las_ar = l->ar; a.las_ai = l->ai; a.las_nz = l->nz; a.las_dz = l->dz;
        a.chi_acc = l->chi_acc; a.chi1 = l->chi->f1; a.chi2 = l->chi->f2;
        const int ppc = s->prm.sp_ppc_r > 0 ? s->prm.sp_ppc_r : 1;
        a.chi_ax = (12.0 * ppc * ppc) / (1.0 + 2.0 * ppc * ppc);
        if (j0 == 1) { a.las_progress = s->las_progress; a.las_base = s->las_base; }   // the step's first launch may overlap the previous step's advance ...
        else if (s->las_progress) {                                                     // ... any later one simply waits for it
            a.las_progress = s->las_progress; a.las_base = s->las_base;
        }
    }
    switch (c->M) {
    case 0: e = pgc ? sweep_launch<0, true>(s->sweep_grid, c->stream, a) : sweep_launch<0, false>(s->sweep_grid, c->stream, a); break;
    case 1: e = pgc ? sweep_launch<1, true>(s->sweep_grid, c->stream, a) : sweep_launch<1, false>(s->sweep_grid, c->stream, a); break;
    default: e = pgc ? sweep_launch<2, true>(s->sweep_grid, c->stream, a) : sweep_launch<2, false>(s->sweep_grid, c->stream, a); break;
    }
    if (e != cudaSuccess) return qpg_cuda_fail(e, "cudaLaunchCooperativeKernel(k_sweep)");
    count_launch(c);
    return 0;
}

static int build_graph(qpg_sim s)
{
    qpg_ctx c = s->ctx;
    cudaStream_t st = c->stream;
    c->capturing = true;
    int rc = 0;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { c->capturing = false; return qpg_cuda_fail(e, "cudaStreamBeginCapture"); }
    rc = enqueue_slice_head(s);
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    cudaStreamCaptureStatus stat;
    cudaGraphConditionalHandle handle = 0;
    // unrolled variant (qpg_sim_set_graph_unroll / QPG_GRAPH_UNROLL=1): no WHILE node -- all iter_max predictor-corrector iterations are captured
    // and the ones behind the converged one skip themselves on the device flag, as on the plain launch path.  cudaGraphLaunch of a graph
    // with a conditional node costs ~100 us of host time on this driver, of a plain graph ~4 us: a pipeline of many short slabs is paced by the
    // host with the WHILE node and by the device's kernel dispatch rate (iter_max x 2 mostly skipped kernels more per slice) without it
    const bool unroll = s->graph_unroll || getenv("QPG_GRAPH_UNROLL") != nullptr;
    if (unroll) {
        for (int l = 0; l < s->prm.iter_max && !rc; l++) rc = enqueue_pc_iteration(s);
        if (!rc) rc = enqueue_slice_tail(s);
        cudaGraph_t gout = nullptr;
        e = cudaStreamEndCapture(st, &gout);
        c->capturing = false;
        if (e != cudaSuccess && !rc) rc = qpg_cuda_fail(e, "cudaStreamEndCapture");
        if (rc) { if (gout) cudaGraphDestroy(gout); return rc; }
        s->graph = gout;
        e = cudaGraphInstantiate(&s->gexec, s->graph, 0);
        if (e != cudaSuccess) return qpg_cuda_fail(e, "cudaGraphInstantiate");
        s->graph_ready = true;
        return 0;
    }
    if (!rc) {
        e = cudaStreamGetCaptureInfo(st, &stat, nullptr, &g, &deps, &ndeps);
        if (e != cudaSuccess) rc = qpg_cuda_fail(e, "cudaStreamGetCaptureInfo");
    }
    if (!rc) {
        e = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault);
        if (e != cudaSuccess) rc = qpg_cuda_fail(e, "cudaGraphConditionalHandleCreate");
    }
    cudaGraphNode_t cnode;
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    if (!rc) {
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        e = cudaGraphAddNode(&cnode, g, deps, ndeps, &cp);
        if (e != cudaSuccess) rc = qpg_cuda_fail(e, "cudaGraphAddNode(conditional)");
    }
    if (!rc) {
        e = cudaStreamUpdateCaptureDependencies(st, &cnode, 1, cudaStreamSetCaptureDependencies);
        if (e != cudaSuccess) rc = qpg_cuda_fail(e, "cudaStreamUpdateCaptureDependencies");
    }
    if (!rc) {
        // body of the WHILE node, captured on a second stream into the child graph
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        cudaStream_t s2;
        cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
        e = cudaStreamBeginCaptureToGraph(s2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) rc = qpg_cuda_fail(e, "cudaStreamBeginCaptureToGraph");
        if (!rc) {
            cudaStream_t keep = c->stream;
            c->stream = s2;
            c->cond_handle = (unsigned long long)handle;
            rc = enqueue_pc_iteration(s);
            c->cond_handle = 0;
            c->stream = keep;
            e = cudaStreamEndCapture(s2, nullptr);
            if (e != cudaSuccess && !rc) rc = qpg_cuda_fail(e, "cudaStreamEndCapture(body)");
        }
        cudaStreamDestroy(s2);
    }
    if (!rc) rc = enqueue_slice_tail(s);
    cudaGraph_t gout = nullptr;
    e = cudaStreamEndCapture(st, &gout);
    c->capturing = false;
    if (e != cudaSuccess && !rc) rc = qpg_cuda_fail(e, "cudaStreamEndCapture");
    if (rc) { if (gout) cudaGraphDestroy(gout); return rc; }
    s->graph = gout;
    e = cudaGraphInstantiate(&s->gexec, s->graph, 0);
    if (e != cudaSuccess) return qpg_cuda_fail(e, "cudaGraphInstantiate");
    s->graph_ready = true;
    return 0;
}

extern "C" int qpg_sim_create(qpg_sim *out, int device, void *cuda_stream, const qpg_sim_params *prm)
{
    ARG_TRY(out && prm, "null arg");
    ARG_TRY(prm->nzp >= 1 && prm->noff2 >= 0 && prm->noff2 + prm->nzp <= prm->nz_total, "bad xi slab");
    ARG_TRY(prm->iter_max >= 1, "iter_max must be >= 1");
    qpg_sim s = new qpg_sim_s();
    memset(s, 0, sizeof(*s));
    s->prm = *prm;
    int rc = qpg_ctx_create(&s->ctx, device, cuda_stream, prm->nr, prm->max_mode, prm->dr, prm->dxi, prm->field_boundary, prm->relax_fac);
    if (rc) { delete s; return rc; }
    qpg_ctx c = s->ctx;
    const int nzp = prm->nzp;
    struct { qpg_field *f; int dim, has2d; } tbl[] = {
        {&s->psi, 1, 1}, {&s->e, 3, 1}, {&s->b, 3, 1}, {&s->e_spe, 3, 1}, {&s->b_spe, 3, 1}, {&s->e_beam, 3, 1}, {&s->b_beam, 3, 1},
        {&s->cu, 3, 1}, {&s->amu, 3, 0}, {&s->acu, 2, 0}, {&s->dcu, 2, 0}, {&s->q_spe, 1, 1}, {&s->q_beam, 1, 1},
        {&s->spe_q, 1, 1}, {&s->spe_qn, 1, 0}, {&s->spe_cu, 3, 0}, {&s->spe_dcu, 2, 0}, {&s->spe_amu, 3, 0}, {&s->beam_q, 1, 1}};
    for (auto &t : tbl) { rc = qpg_field_create(t.f, c, t.dim, nzp, t.has2d); if (rc) return rc; }
    s->cur_dt = prm->dxi;
    s->use_fused = (prm->nr <= FT * FC && prm->max_mode <= 2);
    s->use_sweep = sweep_supported(*prm);
    CUDA_TRY(cudaMalloc(&s->phi, sizeof(double) * (size_t)(prm->nr + 2) * c->P));
    CUDA_TRY(cudaMemsetAsync(s->phi, 0, sizeof(double) * (size_t)(prm->nr + 2) * c->P, c->stream));
    rc = qpg_part2d_create(&s->spe, c, prm->sp_qbm, prm->sp_npmax);
    if (rc) return rc;
    rc = qpg_part3d_create(&s->beam, c, prm->beam_qbm, prm->dt, prm->beam_npmax < 32 ? 32 : prm->beam_npmax, prm->nz_total, prm->noff2, nzp);
    if (rc) return rc;
    if (prm->sp_push_pgc) {
        // a slab of a xi-pipeline holds the envelope of its own slices + guards; the stages exchange them through qpg_laser_set_handoff
        rc = qpg_laser_create(&s->laser, c, nzp, prm->laser_k0, prm->dt, prm->laser_iter < 1 ? 1 : prm->laser_iter);
        if (rc) return rc;
    }
    *out = s;
    return 0;
}
extern "C" int qpg_sim_destroy(qpg_sim s)
{
    if (!s) return 0;
    cudaStreamSynchronize(s->ctx->stream);
    if (s->gexec) cudaGraphExecDestroy(s->gexec);
    if (s->graph) cudaGraphDestroy(s->graph);
    qpg_field all[] = {s->psi, s->e, s->b, s->e_spe, s->b_spe, s->e_beam, s->b_beam, s->cu, s->amu, s->acu, s->dcu, s->q_spe, s->q_beam,
                       s->spe_q, s->spe_qn, s->spe_cu, s->spe_dcu, s->spe_amu, s->beam_q};
    for (auto f : all) qpg_field_destroy(f);
    for (auto f : {s->neut_q, s->neut_cu, s->neut_dcu, s->neut_amu, s->rho_ion, s->rho_ion_add}) if (f) qpg_field_destroy(f);
    cudaFree(s->phi); cudaFree(s->sw_bar); cudaFree(s->sw_xbuf); cudaFree(s->sw_xll); cudaFree(s->sw_prof); cudaFree(s->sw_trace);
    qpg_laser_destroy(s->laser);
    qpg_part2d_destroy(s->spe);
    qpg_part3d_destroy(s->beam);
    qpg_ctx_destroy(s->ctx);
    delete s;
    return 0;
}
extern "C" qpg_ctx qpg_sim_ctx(qpg_sim s) { return s ? s->ctx : nullptr; }
extern "C" qpg_field qpg_sim_field(qpg_sim s, const char *name)
{
    if (!s || !name) return nullptr;
    struct { const char *n; qpg_field f; } tbl[] = {
        {"psi", s->psi}, {"e", s->e}, {"b", s->b}, {"e_spe", s->e_spe}, {"b_spe", s->b_spe}, {"e_beam", s->e_beam}, {"b_beam", s->b_beam},
        {"cu", s->cu}, {"amu", s->amu}, {"acu", s->acu}, {"dcu", s->dcu}, {"q_spe", s->q_spe}, {"q_beam", s->q_beam}, {"spe_q", s->spe_q},
        {"spe_qn", s->spe_qn}, {"spe_cu", s->spe_cu}, {"spe_dcu", s->spe_dcu}, {"spe_amu", s->spe_amu}, {"beam_q", s->beam_q},
        {"neut_q", s->neut_q}, {"rho_ion", s->rho_ion}};
    for (auto &t : tbl) if (t.f && !strcmp(t.n, name)) return t.f;
    qpg_set_error("unknown field '%s'", name);
    return nullptr;
}
extern "C" qpg_part2d qpg_sim_species(qpg_sim s) { return s ? s->spe : nullptr; }
extern "C" qpg_part3d qpg_sim_beam(qpg_sim s) { return s ? s->beam : nullptr; }

// species2d%new (species2d_class.f03:73-133): inject, q = deposit, qn = -q
extern "C" int qpg_sim_init_species(qpg_sim s, const double *x, const double *pm, const double *gamma, const double *psi, const double *q, long npp)
{
    ARG_TRY(s, "null sim");
    int rc = qpg_part2d_upload(s->spe, x, pm, gamma, psi, q, npp);
    if (rc) return rc;
    rc = qpg_part2d_snapshot(s->spe);
    if (rc) return rc;
    rc = qpg_field_fill(s->spe_q, 0.0);
    if (rc) return rc;
    rc = qpg_part2d_qdeposit(s->spe, s->spe_q);
    if (rc) return rc;
    rc = qpg_field_copy(s->spe_q, s->spe_qn);
    if (rc) return rc;
    return qpg_field_scale(s->spe_qn, -1.0);
}
extern "C" int qpg_sim_beam_qdp_begin(qpg_sim s) { ARG_TRY(s, "null sim"); return qpg_field_fill_f2(s->beam_q, 0.0); }   // beam3d_class.f03:207
extern "C" int qpg_sim_beam_qdp_end(qpg_sim s) { ARG_TRY(s, "null sim"); return qpg_part3d_qdeposit(s->beam, s->beam_q); } // :210
extern "C" int qpg_sim_beam_qdp_raw(qpg_sim s) { ARG_TRY(s, "null sim"); return qpg_part3d_qdeposit_raw(s->beam, s->beam_q); }
extern "C" int qpg_sim_beam_qdp_fix(qpg_sim s) { ARG_TRY(s, "null sim"); return qpg_part3d_qdeposit_fix(s->beam, s->beam_q); }
// the raw deposit of a pipeline stage in three parts (qpg_part3d_qdeposit_part): 1 behind qpg_sim_beam_push_interior, 2 behind the push of the
// rest, 3 behind the arrival of the upstream stage's particles
extern "C" int qpg_sim_beam_qdp_part(qpg_sim s, int part) { ARG_TRY(s, "null sim"); return qpg_part3d_qdeposit_part(s->beam, s->beam_q, part); }
// simulation_class.f03:299-331 (everything but the MPI calls), in two halves: _zero touches nothing a hand-off delivers
// (a pipeline stage runs it while it waits for the upstream stage), _add folds the finished beam charge into q_beam
extern "C" int qpg_sim_begin_step_zero(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    int rc;
    if ((rc = qpg_field_fill_f2(s->q_beam, 0.0))) return rc;
    if ((rc = qpg_field_fill_f2(s->q_spe, 0.0))) return rc;
    qpg_field z[] = {s->b, s->e, s->b_spe, s->e_spe, s->psi, s->cu, s->acu, s->amu};
    for (auto f : z) if ((rc = qpg_field_fill(f, 0.0))) return rc;
    return 0;
}
extern "C" int qpg_sim_begin_step_add(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    return qpg_field_add_f2(s->beam_q, s->q_beam);  // beam3d_class.f03:217 add_f2
}
extern "C" int qpg_sim_begin_step(qpg_sim s)
{
    int rc = qpg_sim_begin_step_zero(s);
    return rc ? rc : qpg_sim_begin_step_add(s);
}

extern "C" int qpg_sim_run_slices(qpg_sim s, int j0, int j1)
{
    ARG_TRY(s, "null sim");
    ARG_TRY(j0 >= 1 && j1 <= s->prm.nzp && j0 <= j1, "slice range out of the slab");
    qpg_ctx c = s->ctx;
    int rc;
    {   // slice counter on the device; stand-alone qdeposit for the first slice of the range
        FProgBuilder pb(c);
        FOp &o = pb.add(FOP_SET_FLAG); o.i0 = 3; o.i1 = j0;
        if ((rc = pb.launch(TP_ARITH))) return rc;
        // acc1 may hold the look-ahead deposit of a previous range / renewed particles: clear and redo
        CUDA_TRY(cudaMemsetAsync(s->spe->acc1, 0, sizeof(double) * (size_t)(c->nr + 2) * c->P, c->stream));
        if ((rc = part2d_launch_qdeposit(s->spe))) return rc;
        if (s->neut) {
            for (qpg_part2d p : {s->neut_e, s->neut_i}) {
                CUDA_TRY(cudaMemsetAsync(p->acc1, 0, sizeof(double) * (size_t)(c->nr + 2) * c->P, c->stream));
                if ((rc = part2d_launch_qdeposit(p))) return rc;
            }
        }
    }
    if (s->use_sweep) {
        // one persistent launch per stretch of slices between two sorts
        int j = j0;
        while (j <= j1) {
            int jend = j1;
            const int sf = s->prm.sort_freq;
            if (sf > 0) { const int k = sf - ((s->prm.noff2 + j - 1) % sf) - 1; jend = j + k < j1 ? j + k : j1; }
            if ((rc = sweep_run(s, j, jend))) return rc;
            if (sf > 0 && ((s->prm.noff2 + jend) % sf) == 0) {
                if ((rc = qpg_part2d_sort(s->spe))) return rc;
                CUDA_TRY(cudaMemsetAsync(s->spe->acc1, 0, sizeof(double) * (size_t)(c->nr + 2) * c->P, c->stream));
                if ((rc = part2d_launch_qdeposit(s->spe))) return rc;
            }
            j = jend + 1;
        }
        return 0;
    }
    if (s->prm.use_graph && !s->graph_ready) { if ((rc = build_graph(s))) return rc; }
    for (int j = j0; j <= j1; j++) {
        s->cur_j = j;
        if (s->prm.use_graph) {
            CUDA_TRY(cudaGraphLaunch(s->gexec, c->stream));
            c->launches += 5 + 2;  // head, tail x4, >= 1 PC iteration (exact count comes from the device iteration counter)
        } else if (s->subcyc) {
            if ((rc = enqueue_slice_subcyc(s))) return rc;
        } else {
            if ((rc = enqueue_slice_head(s))) return rc;
            for (int l = 0; l < s->prm.iter_max; l++) if ((rc = enqueue_pc_iteration(s))) return rc;
            if ((rc = enqueue_slice_tail(s))) return rc;
        }
        if (j == 1 && s->back_flag) {   // per-slice launch paths: pack kernels + flag write after the first slice
            if ((rc = qpg_field_pack(s->b, 1, s->back_b))) return rc;
            if ((rc = qpg_field_pack(s->e, 1, s->back_e))) return rc;
            if ((rc = qpg_stream_signal(c->stream, s->back_flag, s->back_seq))) return rc;
            s->back_flag = nullptr;
        }
        if (s->prm.sort_freq > 0 && ((s->prm.noff2 + j) % s->prm.sort_freq) == 0) {
            if ((rc = qpg_part2d_sort(s->spe))) return rc;
            CUDA_TRY(cudaMemsetAsync(s->spe->acc1, 0, sizeof(double) * (size_t)(c->nr + 2) * c->P, c->stream));
            if ((rc = part2d_launch_qdeposit(s->spe))) return rc;
        }
    }
    return 0;
}
extern "C" int qpg_sim_set_back_handoff(qpg_sim s, double *wire_b, double *wire_e, unsigned *flag, unsigned seq)
{
    ARG_TRY(s && wire_b && wire_e && flag, "null arg");
    s->back_b = wire_b; s->back_e = wire_e; s->back_flag = flag; s->back_seq = seq;
    return 0;
}
extern "C" int qpg_sim_beam_push(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    if (!s->prm.beam_evol) return 0;
    int rc = qpg_part3d_push(s->beam, s->prm.beam_push_type, s->e, s->b);
    if (rc) return rc;
    return qpg_part3d_update_bound(s->beam);
}
// the beam push of a pipeline stage in two halves (qpg_part3d_push_interior / _edge): `interior` needs nothing from the downstream stage
extern "C" int qpg_sim_beam_push_interior(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    if (!s->prm.beam_evol) {       // a frozen beam: nothing moves, the whole deposit can be done here
        s->split_deposit = false;
        return qpg_part3d_qdeposit_raw(s->beam, s->beam_q);
    }
    int rc = qpg_part3d_push_interior(s->beam, s->prm.beam_push_type, s->e, s->b);
    if (rc) return rc;
    s->split_deposit = true;       // the next step's beam charge is deposited in parts as the particles are advanced (the volume was zeroed before)
    return qpg_part3d_qdeposit_part(s->beam, s->beam_q, 1);
}
extern "C" int qpg_sim_beam_push_edge(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    if (!s->prm.beam_evol) return 0;
    int rc = qpg_part3d_push_edge(s->beam, s->prm.beam_push_type, s->e, s->b);
    if (rc) return rc;
    if (s->split_deposit && (rc = qpg_part3d_qdeposit_part(s->beam, s->beam_q, 2))) return rc;   // before update_bound: the bitmap follows the pre-compaction order
    return qpg_part3d_update_bound(s->beam);
}
// neut%renew (neutral_class.f03:839-878) + the zeroing of its fields; the particle sets keep npp_hi = npmax so that launches
// captured in a CUDA graph are sized for any number of released electrons (the kernels bound themselves by the device count)
static int neutral_renew(qpg_sim s)
{
    int rc;
    if ((rc = qpg_neutral_reset(s->neut))) return rc;
    for (qpg_part2d p : {s->neut_e, s->neut_i}) {
        if ((rc = qpg_part2d_clear(p))) return rc;
        p->npp_hi = p->npmax;
        CUDA_TRY(cudaMemsetAsync(p->acc1, 0, sizeof(double) * (size_t)(s->ctx->nr + 2) * s->ctx->P, s->ctx->stream));
    }
    for (qpg_field f : {s->rho_ion, s->neut_q, s->neut_cu}) if ((rc = qpg_field_fill(f, 0.0))) return rc;
    return 0;
}
// Attaches a neutral species (qpg_neutral_create) and its two particle sets to the slice loop of this sim.  Runs on the per-slice
// launch paths (CUDA-graph replay or plain stream): the persistent sweep kernel and the cluster programs are switched off.
// Call before the first qpg_sim_run_slices; the handles stay owned by the caller and must outlive the sim's use of them.
extern "C" int qpg_sim_attach_neutral(qpg_sim s, qpg_neutral n, qpg_part2d electrons, qpg_part2d ions)
{
    ARG_TRY(s && n && electrons && ions && electrons != ions, "null / identical handles");
    ARG_TRY(electrons->ctx == s->ctx && ions->ctx == s->ctx, "the particle sets must be created on qpg_sim_ctx(sim)");
    ARG_TRY(!s->neut, "a neutral species is already attached");
    ARG_TRY(!s->prm.sp_push_std && !s->prm.sp_push_pgc, "neutral species: robust pusher only");
    // one update can release ppc1 * ppc2 electrons in every (cell, sector): the ion buffer must hold that, the electron set at least that
    // (it accumulates over the slices of a step; an overflow later on latches QPG_ERR_STATE, neutral.cu k_neutral_counts)
    ARG_TRY(ions->npmax >= qpg_neutral_max_new_per_update(n) && electrons->npmax >= qpg_neutral_max_new_per_update(n),
            "particle sets too small: need npmax >= ppc1 * ppc2 * nr * num_theta");
    qpg_ctx c = s->ctx;
    int rc;
    struct { qpg_field *f; int dim, has2d; } tbl[] = {{&s->neut_q, 1, 1}, {&s->neut_cu, 3, 0}, {&s->neut_dcu, 2, 0}, {&s->neut_amu, 3, 0}, {&s->rho_ion, 1, 1}, {&s->rho_ion_add, 1, 0}};
    for (auto &t : tbl) if ((rc = qpg_field_create(t.f, c, t.dim, s->prm.nzp, t.has2d))) return rc;
    if (s->graph_ready) { cudaGraphExecDestroy(s->gexec); cudaGraphDestroy(s->graph); s->gexec = nullptr; s->graph = nullptr; s->graph_ready = false; }
    s->use_sweep = false; s->use_fused = false;
    s->neut = n; s->neut_e = electrons; s->neut_i = ions;
    return neutral_renew(s);
}
// Switches the sub-cycling variant of the slice loop on (off): plain per-slice launches with one host synchronisation per slice.
extern "C" int qpg_sim_set_subcyc(qpg_sim s, int on, double exp_fac_max, double exp_fac_clamped, double dt_min)
{
    ARG_TRY(s, "null sim");
    if (!on) { s->subcyc = false; return 0; }
    ARG_TRY(exp_fac_max > 1.0 && exp_fac_clamped > 1.0 && dt_min >= 0.0, "expansion factors must exceed 1");
    ARG_TRY(!s->prm.sp_push_std && !s->prm.sp_push_pgc, "sub-cycling: robust pusher only");
    if (s->graph_ready) { cudaGraphExecDestroy(s->gexec); cudaGraphDestroy(s->graph); s->gexec = nullptr; s->graph = nullptr; s->graph_ready = false; }
    s->subcyc = true; s->sc_exp_fac_max = exp_fac_max; s->sc_exp_fac_clamped = exp_fac_clamped; s->sc_dt_min = dt_min;
    s->use_sweep = false; s->use_fused = false; s->prm.use_graph = 0;
    return 0;
}
extern "C" long qpg_sim_subcycles(qpg_sim s) { return s ? s->host_subcycles : -1; }
extern "C" int qpg_sim_renew(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    int rc = qpg_part2d_renew(s->spe);  // species2d%renew: same lattice, q and qn unchanged for time-independent profiles
    if (rc || !s->neut) return rc;
    return neutral_renew(s);
}
extern "C" int qpg_sim_stats(qpg_sim s, long *updates, long *pc_iters, long *slices)
{
    ARG_TRY(s, "null sim");
    qpg_ctx c = s->ctx;
    int fl[8], nbeam = 0;
    long long cnt[2];
    CUDA_TRY(cudaMemcpyAsync(fl, c->flags, sizeof(fl), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(cnt, c->counters, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&nbeam, s->beam->d_npp, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->beam->npp_hi = nbeam;   // the host's upper bound of the beam count grows with every hand-off (qpg_part3d_unpack): resynchronise it here
    if (updates) *updates = (long)cnt[0];
    if (pc_iters) *pc_iters = (long)cnt[1];
    if (slices) *slices = fl[4];
    return qpg_ctx_check_latches(c, fl);
}
extern "C" int qpg_sim_set_fused(qpg_sim s, int on)
{
    ARG_TRY(s, "null sim");
    const bool can = s->prm.nr <= FT * FC && s->prm.max_mode <= 2;
    if (on && (s->neut || s->subcyc)) { qpg_set_error("a neutral species / sub-cycling runs on the op-list programs only"); return QPG_ERR_UNSUPPORTED; }
    if (on && !can) { qpg_set_error("fused cluster programs need nr <= %d and max_mode <= 2", FT * FC); return QPG_ERR_UNSUPPORTED; }
    if ((on != 0) != s->use_fused && s->graph_ready) {   // the captured graph holds the other variant
        cudaGraphExecDestroy(s->gexec); cudaGraphDestroy(s->graph);
        s->gexec = nullptr; s->graph = nullptr; s->graph_ready = false;
    }
    s->use_fused = on != 0;
    return 0;
}
extern "C" int qpg_sim_set_graph(qpg_sim s, int use_graph) { ARG_TRY(s, "null sim"); ARG_TRY(!(use_graph && s->subcyc), "sub-cycling needs the plain launch path"); s->prm.use_graph = use_graph != 0; return 0; }
extern "C" int qpg_sim_set_graph_unroll(qpg_sim s, int on)
{
    ARG_TRY(s, "null sim");
    if ((on != 0) != s->graph_unroll && s->graph_ready) { cudaGraphExecDestroy(s->gexec); cudaGraphDestroy(s->graph); s->gexec = nullptr; s->graph = nullptr; s->graph_ready = false; }
    s->graph_unroll = on != 0;
    return 0;
}
extern "C" qpg_laser qpg_sim_laser(qpg_sim s) { return s ? s->laser : nullptr; }
int qpg_laser_advance_overlapped(qpg_laser l, unsigned **progress, unsigned *base);   // laser.cu
extern "C" int qpg_sim_laser_advance(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    if (!s->laser) return 0;
    s->las_progress = nullptr;
    if (s->use_sweep && s->sweep_grid > 0 && s->las_overlap) {
        // the next sweep kernel runs beside the solve (it leaves one SM free, sweep_prepare) and follows its progress slice by slice
        int rc = qpg_laser_advance_overlapped(s->laser, &s->las_progress, &s->las_base);
        if (rc != QPG_ERR_UNSUPPORTED) return rc;
        s->las_progress = nullptr;
    }
    return qpg_laser_advance(s->laser);
}
extern "C" int qpg_sim_set_laser_overlap(qpg_sim s, int on)
{
    ARG_TRY(s, "null sim");
    s->las_overlap_req = on != 0;
    if (s->sweep_grid > 0) return qpg_sim_set_sweep_ctas(s, s->sweep_ctas_req);   // re-derive
    return 0;
}
extern "C" int qpg_sim_set_sweep(qpg_sim s, int on)
{
    ARG_TRY(s, "null sim");
    if (on && (s->neut || s->subcyc)) { qpg_set_error("a neutral species / sub-cycling runs on the per-slice launch paths only"); return QPG_ERR_UNSUPPORTED; }
    if (on && !sweep_supported(s->prm)) { qpg_set_error("the persistent sweep kernel needs max_mode <= 2, nr <= %d and the robust pusher", SW_MAX_TEAM * ST_N); return QPG_ERR_UNSUPPORTED; }
    s->use_sweep = on != 0;
    return 0;
}
extern "C" int qpg_sim_set_sweep_ctas(qpg_sim s, int n)
{
    ARG_TRY(s, "null sim");
    s->sweep_ctas_req = n;
    if (s->sweep_grid > 0) {   // already prepared: re-derive the grid
        cudaStreamSynchronize(s->ctx->stream);
        cudaFree(s->sw_bar); cudaFree(s->sw_xbuf); cudaFree(s->sw_xll); cudaFree(s->sw_prof); cudaFree(s->sw_trace);
        s->sw_bar = nullptr; s->sw_xbuf = nullptr; s->sw_xll = nullptr; s->sw_prof = nullptr; s->sw_trace = nullptr; s->sweep_grid = 0;
    }
    return 0;
}
extern "C" int qpg_sim_sweep_profile(qpg_sim s, double *out8, int reset)
{
    ARG_TRY(s && out8, "null arg");
    for (int k = 0; k < 12; k++) out8[k] = 0.0;
    if (!s->sw_prof) return 0;
    long long h[32];
    qpg_ctx c = s->ctx;
    CUDA_TRY(cudaMemcpyAsync(h, s->sw_prof, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    if (reset) CUDA_TRY(cudaMemsetAsync(s->sw_prof, 0, sizeof(h), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // the abort flag of the last launch: a watchdog exit must not pass silently
    unsigned ab = 0;
    CUDA_TRY(cudaMemcpyAsync(&ab, s->sw_bar + 64, sizeof(ab), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (ab) { qpg_set_error("sweep kernel aborted: a grid barrier timed out"); return QPG_ERR_STATE; }
    for (int k = 0; k < 12; k++) out8[k] = (double)h[k];
    if (getenv("QPG_SWEEP_STAMPS")) {   // development aid: stage stamps inside the field programs (cycles of CTA 0, thread 0)
        fprintf(stderr, "sweep stamps (cycles/slice):");
        for (int k = 16; k < 29; k++) fprintf(stderr, " %.0f", (double)h[k] / (h[6] > 0 ? (double)h[6] : 1.0));
        fprintf(stderr, "\n");
    }
    return 0;
}
extern "C" int qpg_sim_debug_abort(qpg_sim s)
{
    ARG_TRY(s, "null sim");
    if (!s->use_sweep) { qpg_set_error("the watchdog belongs to the persistent sweep kernel"); return QPG_ERR_UNSUPPORTED; }
    int rc = sweep_prepare(s);
    if (rc) return rc;
    const unsigned one = 1;
    CUDA_TRY(cudaMemcpyAsync(s->sw_bar + 64, &one, sizeof(one), cudaMemcpyHostToDevice, s->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    return 0;
}
extern "C" int qpg_sim_slice_trace(qpg_sim s, double *ns_per_slice, int *iters_per_slice)
{
    ARG_TRY(s && ns_per_slice && iters_per_slice, "null arg");
    const int n = s->prm.nzp;
    for (int k = 0; k < n; k++) { ns_per_slice[k] = 0.0; iters_per_slice[k] = 0; }
    if (!s->sw_trace) return 0;
    std::vector<long long> h(2 * (size_t)n);
    CUDA_TRY(cudaMemcpyAsync(h.data(), s->sw_trace, sizeof(long long) * 2 * n, cudaMemcpyDeviceToHost, s->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    for (int k = 0; k < n; k++) { ns_per_slice[k] = (double)h[2 * k]; iters_per_slice[k] = (int)h[2 * k + 1]; }
    return 0;
}

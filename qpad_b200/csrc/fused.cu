// fused.cu -- the three per-slice field programs of the fast path as thread-block-CLUSTER kernels.
//
// Why: the op-list interpreter (fields.cu, one CTA) is bound by the L2->SM bandwidth and latency of ONE SM: a slice
// moves ~1 MB of field data through it and every op exposes an L2 round trip (ncu: 89-138 us per program at
// nr=1024).  Here 8 CTAs x 128 threads form one cluster; thread <-> radial node, all azimuthal planes of a node live
// in registers, a program is 2-3 stages separated by cluster barriers, and the tridiagonal solves are block scans
// whose per-CTA totals are exchanged through distributed shared memory.  Arithmetic is identical to the op-list path
// (same right-hand-side functions, same Green's-function factors); tests compare both paths with the oracle.
//
//   A : simulation_class.f03:344-377  q_beam slice -> bt(beam), qdp epilogue, psi, bz, [record, b, ez, et]
//   C : :378-396 + :375-377           amjdp epilogue, djdxi, bt_iter, bz, compare, [record, b, ez, et]
//   D : :401-426, :452-456            cbq, rho, e_spe, J-perp predictor, slice -> volume copies
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define FT 128  // threads per CTA
#define FC 8    // CTAs per cluster  (FT*FC = 1024 radial nodes)

struct FusedArgs {
    int nr, iter_max;
    double dr, dxi, relax, reltol, abstol;
    double *psi, *e, *b, *e_spe, *b_spe, *b_beam, *cu, *amu, *acu, *dcu, *q_spe, *q_beam, *spe_q, *spe_qn, *spe_cu, *spe_dcu, *spe_amu;
    double *q_beam2, *spe_q2, *cu2, *q_spe2, *e2, *b2, *psi2, *b_spe2, *e_spe2;
    double *acc1, *acc8, *phi;  // phi: scratch potential of the beam solve [(nr+2)][P]
    const int *d_npp;
    const OpCoef *ops;
    double *conv_old, *conv_out;
    int *flags;
    long long *counters;
    unsigned long long cond_handle;
};

// ---- cluster-wide scans -------------------------------------------------------------------------------------
// fa/fb: one value per system for this thread's node.  Returns inclusive prefix (fa) and exclusive suffix (fb) over
// all FT*FC nodes; `red` values are summed over the cluster (NRED of them).
template <int NS, int NRED>
__device__ __forceinline__ void team_scan(double (&fa)[NS], double (&fb)[NS], double (&red)[NRED], double *sm /* see size below */, cg::cluster_group &cluster)
{
    constexpr int NW = FT / 32;
    // layout of sm: wA[NS][NW], wB[NS][NW], wR[NRED][NW], ex[2*NS+NRED] (exchanged), off[2*NS+NRED]
    double *wA = sm, *wB = wA + NS * NW, *wR = wB + NS * NW, *ex = wR + NRED * NW, *off = ex + (2 * NS + NRED);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double incl = fa[s];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wA[s * NW + warp] = incl;
        fa[s] = incl;
        double own = fb[s], sfx = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_down_sync(0xffffffffu, sfx, o); if (lane + o < 32) sfx += t; }
        if (lane == 0) wB[s * NW + warp] = sfx;
        double ex1 = __shfl_down_sync(0xffffffffu, sfx, 1);
        fb[s] = lane == 31 ? 0.0 : ex1;
    }
#pragma unroll
    for (int k = 0; k < NRED; k++) {
        double v = red[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wR[k * NW + warp] = v;
    }
    __syncthreads();
    // CTA totals -> exchange buffer
    if (threadIdx.x < 2 * NS + NRED) {
        const int k = threadIdx.x;
        const double *src = k < NS ? wA + k * NW : (k < 2 * NS ? wB + (k - NS) * NW : wR + (k - 2 * NS) * NW);
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) t += src[w];
        ex[k] = t;
    }
    cluster.sync();
    if (threadIdx.x < 2 * NS + NRED) {
        const int k = threadIdx.x;
        const unsigned me = cluster.block_rank();
        double t = 0.0;
        for (unsigned r = 0; r < FC; r++) {
            const double *rex = cluster.map_shared_rank(ex, r);
            const double v = rex[k];
            if (k < NS) { if (r < me) t += v; }
            else if (k < 2 * NS) { if (r > me) t += v; }
            else t += v;
        }
        off[k] = t;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double pa = off[s], pb = off[NS + s];
#pragma unroll
        for (int w = 0; w < NW; w++) { if (w < warp) pa += wA[s * NW + w]; if (w > warp) pb += wB[s * NW + w]; }
        fa[s] += pa;
        fb[s] += pb;
    }
#pragma unroll
    for (int k = 0; k < NRED; k++) red[k] = off[2 * NS + k];
}
template <int NS, int NRED> struct TeamScanSmem { static constexpr int doubles = (2 * NS + NRED) * (FT / 32) + 2 * (2 * NS + NRED); };

// cluster-wide max of two values
__device__ __forceinline__ void team_max2(double &a, double &b, double *sm /* [2*NW + 2 + 2] */, cg::cluster_group &cluster)
{
    constexpr int NW = FT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o)); }
    if (lane == 0) { sm[warp] = a; sm[NW + warp] = b; }
    __syncthreads();
    double *ex = sm + 2 * NW;
    if (threadIdx.x < 2) { double t = 0.0; for (int w = 0; w < NW; w++) t = fmax(t, sm[threadIdx.x * NW + w]); ex[threadIdx.x] = t; }
    cluster.sync();
    double ra = 0.0, rb = 0.0;
    for (unsigned r = 0; r < FC; r++) { const double *rex = cluster.map_shared_rank(ex, r); ra = fmax(ra, rex[0]); rb = fmax(rb, rex[1]); }
    a = ra; b = rb;
}

// field_e_class.f03:412-514 solve_field_et for one node / plane; b values passed in, psi read from the global image
template <int M>
__device__ __forceinline__ void et_node(const double *__restrict__ psi, int nr, double idr, int pl, int i, double b_r, double b_phi, double &er, double &ephi)
{
    constexpr int P = 2 * M + 1;
    const double idrh = 0.5 * idr;
    const int m = (pl + 1) >> 1;
    if (m == 0) {
        if (i == 1) { er = 0.0; ephi = 0.0; return; }
        if (i == nr) er = b_phi + idrh * (4.0 * FX(psi, 1, nr - 1, 0, 0) - FX(psi, 1, nr - 2, 0, 0) - 3.0 * FX(psi, 1, nr, 0, 0));
        else er = b_phi - idrh * (FX(psi, 1, i + 1, 0, 0) - FX(psi, 1, i - 1, 0, 0));
        ephi = -b_r;
        return;
    }
    const bool im = pl > 0 && (pl & 1) == 0;
    const int po = im ? pl - 1 : pl + 1;
    const double sg = im ? -1.0 : 1.0;
    if (i == 1) {
        if (m == 1) { er = b_phi - idr * FX(psi, 1, 2, pl, 0); ephi = -b_r + sg * idr * FX(psi, 1, 2, po, 0); }
        else { er = 0.0; ephi = 0.0; }
        return;
    }
    const double ir = idr / (double)(i - 1);
    if (i == nr) er = b_phi + idrh * (4.0 * FX(psi, 1, nr - 1, pl, 0) - FX(psi, 1, nr - 2, pl, 0) - 3.0 * FX(psi, 1, nr, pl, 0));
    else er = b_phi - idrh * (FX(psi, 1, i + 1, pl, 0) - FX(psi, 1, i - 1, pl, 0));
    ephi = -b_r + sg * ir * m * FX(psi, 1, i, po, 0);
}

// field_src_class.f03:273-405 solve_field_djdxi for one node / plane / component
template <int M>
__device__ __forceinline__ double djdxi_node(const double *__restrict__ acu, const double *__restrict__ amu, int nr, double idr, int pl, int c, int i)
{
    constexpr int P = 2 * M + 1;
    const double idrh = 0.5 * idr;
    const int m = (pl + 1) >> 1;
    if (m == 0) {
        if (i == 1) return 0.0;
        const double ir = idr / (double)(i - 1);
        if (i == nr) return FX(acu, 2, nr, 0, c) + idrh * (4.0 * FX(amu, 3, nr - 1, 0, c) - FX(amu, 3, nr - 2, 0, c) - 3.0 * FX(amu, 3, nr, 0, c)) - ir * FX(amu, 3, nr, 0, c);
        return FX(acu, 2, i, 0, c) - idrh * (FX(amu, 3, i + 1, 0, c) - FX(amu, 3, i - 1, 0, c)) - ir * FX(amu, 3, i, 0, c);
    }
    const bool im = pl > 0 && (pl & 1) == 0;
    const int po = im ? pl - 1 : pl + 1;
    const double sg = im ? -1.0 : 1.0;
    if (i == 1) {
        if (m == 1) return FX(acu, 2, 1, pl, c) - 2.0 * idr * FX(amu, 3, 2, pl, c) + sg * m * idr * FX(amu, 3, 2, po, c + 1);
        return 0.0;
    }
    if (i == 2 && m == 2) {
        const double ir = idr;
        return FX(acu, 2, 2, pl, c) - idr * (FX(amu, 3, 3, pl, c) - FX(amu, 3, 2, pl, c)) - ir * FX(amu, 3, 2, pl, c) + sg * m * ir * FX(amu, 3, 2, po, c + 1);
    }
    const double ir = idr / (double)(i - 1);
    if (i == nr)
        return FX(acu, 2, nr, pl, c) + idrh * (4.0 * FX(amu, 3, nr - 1, pl, c) - FX(amu, 3, nr - 2, pl, c) - 3.0 * FX(amu, 3, nr, pl, c)) - ir * FX(amu, 3, nr, pl, c) + sg * m * ir * FX(amu, 3, nr, po, c + 1);
    return FX(acu, 2, i, pl, c) - idrh * (FX(amu, 3, i + 1, pl, c) - FX(amu, 3, i - 1, pl, c)) - ir * FX(amu, 3, i, pl, c) + sg * m * ir * FX(amu, 3, i, po, c + 1);
}

// x = p*S + u*T (+ decoupled axis row)
__device__ __forceinline__ double green_apply(const OpCoef &oc, int t, double S, double T, double d_own)
{
    double x = __ldg(oc.pT + t) * S + __ldg(oc.uT + t) * T;
    if (t == 0 && oc.axis_inv != 0.0) x = d_own * oc.axis_inv;
    return x;
}

// own-node variant of rhs_bt_iter (field_b_class.f03:360-506): dcu of the node comes from registers
template <int M>
__device__ __forceinline__ double rhs_bt_iter_own(const SolveCtx &sc, const double (&dcu)[2 * M + 1][2], const double *cu, const double *b, double relax_idr2,
                                                  int which, int pl, int i)
{
    constexpr int P = 2 * M + 1;
    const int nr = sc.nr, m = (pl + 1) >> 1;
    const double idr = sc.idr, idrh = sc.idrh;
    if (m == 0) {
        if (i == 1) return 0.0;
        if (which == 0) return -dcu[0][1] - FX(b, 3, i, 0, 0) * relax_idr2;
        double dj;
        if (i == 2) dj = idr * (FX(cu, 3, 3, 0, 2) - FX(cu, 3, 2, 0, 2));
        else if (i == nr) dj = idrh * (3.0 * FX(cu, 3, nr, 0, 2) - 4.0 * FX(cu, 3, nr - 1, 0, 2) + FX(cu, 3, nr - 2, 0, 2));
        else dj = idrh * (FX(cu, 3, i + 1, 0, 2) - FX(cu, 3, i - 1, 0, 2));
        return dcu[0][0] + dj - FX(b, 3, i, 0, 1) * relax_idr2;
    }
    const int pr = 2 * m - 1, pi = 2 * m;
    double s1_re, s1_im, s2_re, s2_im;
    if (i == 1) {
        if (m == 1) {
            s1_re = -dcu[pr][1] + idr * m * FX(cu, 3, 2, pi, 2);
            s1_im = -dcu[pi][1] - idr * m * FX(cu, 3, 2, pr, 2);
            s2_re = dcu[pr][0] + idr * FX(cu, 3, 2, pr, 2);
            s2_im = dcu[pi][0] + idr * FX(cu, 3, 2, pi, 2);
        } else if ((m & 1) == 0) {
            s1_re = s1_im = s2_re = s2_im = 0.0;
        } else {
            s1_re = idr * m * FX(cu, 3, 2, pi, 2);
            s1_im = idr * m * FX(cu, 3, 2, pr, 2);
            s2_re = idr * FX(cu, 3, 2, pr, 2);
            s2_im = idr * FX(cu, 3, 2, pi, 2);
        }
    } else {
        const double ir = idr / (double)(i - 1);
        s1_re = -dcu[pr][1] + m * FX(cu, 3, i, pi, 2) * ir;
        s1_im = -dcu[pi][1] - m * FX(cu, 3, i, pr, 2) * ir;
        if (i == nr) {
            s2_re = dcu[pr][0] + idrh * (3.0 * FX(cu, 3, nr, pr, 2) - 4.0 * FX(cu, 3, nr - 1, pr, 2) + FX(cu, 3, nr - 2, pr, 2));
            s2_im = dcu[pi][0] + idrh * (3.0 * FX(cu, 3, nr, pi, 2) - 4.0 * FX(cu, 3, nr - 1, pi, 2) + FX(cu, 3, nr - 2, pi, 2));
        } else {
            s2_re = dcu[pr][0] + idrh * (FX(cu, 3, i + 1, pr, 2) - FX(cu, 3, i - 1, pr, 2));
            s2_im = dcu[pi][0] + idrh * (FX(cu, 3, i + 1, pi, 2) - FX(cu, 3, i - 1, pi, 2));
        }
    }
    const double brr = FX(b, 3, i, pr, 0), bri = FX(b, 3, i, pi, 0), bpr = FX(b, 3, i, pr, 1), bpi = FX(b, 3, i, pi, 1);
    const bool im = (pl & 1) == 0;
    if (which == 0) return im ? (s1_im + s2_re - (bri + bpr) * relax_idr2) : (s1_re - s2_im - (brr - bpi) * relax_idr2);
    return im ? (s1_im - s2_re - (bri - bpr) * relax_idr2) : (s1_re + s2_im - (brr + bpi) * relax_idr2);
}

// guard-inclusive node work of program A stage 1 (q_beam slice copy, qdp epilogue) for one node
template <int M>
__device__ __forceinline__ void a_stage1_node(const FusedArgs &a, int n, int j, size_t n1, double (&qs)[2 * M + 1], double (&qb)[2 * M + 1])
{
    constexpr int P = 2 * M + 1;
    double raw[P], qn[P];
#pragma unroll
    for (int pl = 0; pl < P; pl++) {
        const size_t k = (size_t)n * P + pl;
        qb[pl] = a.q_beam2[(size_t)(j - 1) * n1 + k];
        raw[pl] = a.acc1[k];
        qn[pl] = a.spe_qn[k];
    }
#pragma unroll
    for (int pl = 0; pl < P; pl++) {
        const size_t k = (size_t)n * P + pl;
        const double sq = axis_fix_q(n, pl, raw[pl]);
        qs[pl] = sq + qn[pl];
        a.q_beam[k] = qb[pl];       // copy_slice 2to1 :344
        a.acc1[k] = 0.0;
        a.spe_q[k] = sq;            // species2d qdp :198-204
        a.q_spe[k] = qs[pl];
    }
}
// guard-inclusive node work of program C stage 1 (amjdp epilogue, single species)
template <int M>
__device__ __forceinline__ void c_stage1_node(const FusedArgs &a, int n)
{
    constexpr int P = 2 * M + 1;
    double v[P][8];
#pragma unroll
    for (int pl = 0; pl < P; pl++)
#pragma unroll
        for (int c = 0; c < 8; c++) v[pl][c] = a.acc8[((size_t)n * P + pl) * 8 + c];
#pragma unroll
    for (int pl = 0; pl < P; pl++) {
        const size_t np = (size_t)n * P + pl;
#pragma unroll
        for (int c = 0; c < 8; c++) { v[pl][c] = axis_fix_amj(n, pl, c, v[pl][c]); a.acc8[np * 8 + c] = 0.0; }
#pragma unroll
        for (int c = 0; c < 3; c++) { a.spe_cu[np * 3 + c] = v[pl][c]; a.cu[np * 3 + c] = v[pl][c]; a.spe_amu[np * 3 + c] = v[pl][5 + c]; a.amu[np * 3 + c] = v[pl][5 + c]; }
#pragma unroll
        for (int c = 0; c < 2; c++) { a.spe_dcu[np * 2 + c] = v[pl][3 + c]; a.acu[np * 2 + c] = v[pl][3 + c]; }
    }
}

// ============================================================================================================
template <int M>
__global__ void __cluster_dims__(FC, 1, 1) __launch_bounds__(FT) k_fused_A(const __grid_constant__ FusedArgs a)
{
    constexpr int P = 2 * M + 1, NS = 4 * P;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double sm[TeamScanSmem<NS, 1>::doubles];
    const int nr = a.nr, tid = threadIdx.x, i = blockIdx.x * FT + tid + 1;
    const bool valid = i <= nr;
    const int j = a.flags[3];
    const size_t n1 = (size_t)(nr + 2) * P;
    const double idr = 1.0 / a.dr;
    SolveCtx sc; sc.nr = nr; sc.M = M; sc.P = P; sc.logC = 0; sc.stride = 1; sc.dr = a.dr; sc.idr = idr; sc.idrh = 0.5 * idr;
    if (i == 1) { a.counters[0] += (long long)*a.d_npp; }
    double d[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) d[s] = 0.0;
    if (valid) {
        double qs[P], qb[P];
#pragma unroll
        for (int pl = 0; pl < P; pl++) { d[2 * P + pl] = rhs_bz(sc, a.cu, pl, i); d[3 * P + pl] = rhs_ez(sc, a.cu, pl, i); }
        a_stage1_node<M>(a, i, j, n1, qs, qb);
#pragma unroll
        for (int pl = 0; pl < P; pl++) { d[pl] = -1.0 * qs[pl]; d[P + pl] = -1.0 * qb[pl]; }
        if (i == 1) a_stage1_node<M>(a, 0, j, n1, qs, qb);
        if (i == nr) a_stage1_node<M>(a, nr + 1, j, n1, qs, qb);
    }
    double red[1] = {0.0};
    if (valid && i >= 2 && i <= nr - 2) red[0] = d[3 * P] * (double)(i - 1);             // field_e_class.f03:189-197
    double fa[NS], fb[NS];
    const int t = i - 1;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int kind = s < P ? FK_PSI : (s < 2 * P ? FK_BT : (s < 3 * P ? FK_BZ : FK_EZ));
        const OpCoef &oc = a.ops[kind * (QPG_MAX_MODE + 1) + (((s % P) + 1) >> 1)];
        fa[s] = valid ? __ldg(oc.qT + t) * d[s] : 0.0;
        fb[s] = valid ? __ldg(oc.vT + t) * d[s] : 0.0;
    }
    team_scan<NS, 1>(fa, fb, red, sm, cluster);
    double x[NS];
    if (valid) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int kind = s < P ? FK_PSI : (s < 2 * P ? FK_BT : (s < 3 * P ? FK_BZ : FK_EZ));
            const OpCoef &oc = a.ops[kind * (QPG_MAX_MODE + 1) + (((s % P) + 1) >> 1)];
            x[s] = green_apply(oc, t, fa[s], fb[s], d[s]);
        }
        {   // E_z m=0: row 1 of the source is -8*(div - edge term)  (:199-209); by linearity x += rhs1 * G(:,1)
            const OpCoef &oc = a.ops[FK_EZ * (QPG_MAX_MODE + 1)];
            const double div = red[0] - sc.idrh * (FX(a.cu, 3, nr - 2, 0, 0) + FX(a.cu, 3, nr - 1, 0, 0)) * ((double)nr - 2.5);
            x[3 * P] += (-8.0 * div) * (__ldg(oc.pT + t) * __ldg(oc.qT));
        }
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            const bool ax0 = pl > 0 && i == 1;
            if (ax0) { x[pl] = 0.0; x[2 * P + pl] = 0.0; x[3 * P + pl] = 0.0; }
            FX(a.psi, 1, i, pl, 0) = x[pl];
            FX(a.phi, 1, i, pl, 0) = x[P + pl];
            FX(a.b_spe, 3, i, pl, 2) = x[2 * P + pl];
            FX(a.e, 3, i, pl, 2) = x[3 * P + pl];
        }
    }
    cluster.sync();
    if (i == 1) { a.flags[0] = 0; a.flags[2] = 0; }                                      // PC loop starts
    if (valid) {
        const double idrh = 0.5 * idr;
        double sre = 0.0, sim = 0.0;
        double ob[P][3], obb[P][2], oe[P][2];
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            const int m = (pl + 1) >> 1;
            // field_b_class.f03:545-701 get_solution_bt
            double bphi, br = 0.0;
            if (i == 1) bphi = (m == 1) ? -idr * FX(a.phi, 1, 2, pl, 0) : 0.0;
            else if (i == nr) bphi = -idrh * (3.0 * FX(a.phi, 1, nr, pl, 0) - 4.0 * FX(a.phi, 1, nr - 1, pl, 0) + FX(a.phi, 1, nr - 2, pl, 0));
            else bphi = -idrh * (FX(a.phi, 1, i + 1, pl, 0) - FX(a.phi, 1, i - 1, pl, 0));
            if (m > 0) {
                const bool im = (pl & 1) == 0;
                const int po = im ? pl - 1 : pl + 1;
                const double sg = im ? 1.0 : -1.0;
                if (i == 1) br = (m == 1) ? sg * idr * m * FX(a.phi, 1, 2, po, 0) : 0.0;
                else br = sg * (idr / (double)(i - 1)) * m * FX(a.phi, 1, i, po, 0);
            }
            obb[pl][0] = br; obb[pl][1] = bphi;
            const double bs_r = FX(a.b_spe, 3, i, pl, 0), bs_p = FX(a.b_spe, 3, i, pl, 1);
            const double v = fabs(bs_p);                                                  // convergence_tester 'record' :548-558
            if (pl > 0 && (pl & 1) == 0) sim += v; else sre += v;
            ob[pl][0] = bs_r + br; ob[pl][1] = bs_p + bphi;                               // b = b_spe + b_beam :375
            ob[pl][2] = x[2 * P + pl] + FX(a.b_beam, 3, i, pl, 2);
            et_node<M>(a.psi, nr, idr, pl, i, ob[pl][0], ob[pl][1], oe[pl][0], oe[pl][1]);  // :377
        }
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            FX(a.b_beam, 3, i, pl, 0) = obb[pl][0]; FX(a.b_beam, 3, i, pl, 1) = obb[pl][1];
            FX(a.b, 3, i, pl, 0) = ob[pl][0]; FX(a.b, 3, i, pl, 1) = ob[pl][1]; FX(a.b, 3, i, pl, 2) = ob[pl][2];
            FX(a.e, 3, i, pl, 0) = oe[pl][0]; FX(a.e, 3, i, pl, 1) = oe[pl][1];
        }
        a.conv_old[i] = sre; a.conv_old[nr + 2 + i] = sim;
    }
}

// ============================================================================================================
template <int M>
__global__ void __cluster_dims__(FC, 1, 1) __launch_bounds__(FT) k_fused_C(const __grid_constant__ FusedArgs a)
{
    constexpr int P = 2 * M + 1, NS = 4 * P;
    if (a.flags[0]) return;  // converged: the remaining pre-enqueued iterations are no-ops (uniform over the cluster)
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double sm[TeamScanSmem<NS, 1>::doubles];
    __shared__ double smx[2 * (FT / 32) + 4];
    const int nr = a.nr, tid = threadIdx.x, i = blockIdx.x * FT + tid + 1;
    const bool valid = i <= nr;
    const double idr = 1.0 / a.dr;
    SolveCtx sc; sc.nr = nr; sc.M = M; sc.P = P; sc.logC = 0; sc.stride = 1; sc.dr = a.dr; sc.idr = idr; sc.idrh = 0.5 * idr;
    // stage 1: deposit epilogue (part2d_class.f03:916-981) + species2d amjdp adds (:250-276), single species
    if (valid) {
        c_stage1_node<M>(a, i);
        if (i == 1) c_stage1_node<M>(a, 0);
        if (i == nr) c_stage1_node<M>(a, nr + 1);
    }
    cluster.sync();
    // stage 2: djdxi (:390), sources of bt_iter (:391), bz (:392), ez (:376 of the next pass / :415)
    double d[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) d[s] = 0.0;
    const double relax_idr2 = a.relax * (idr * idr);
    if (valid) {
        double dcu[P][2];
#pragma unroll
        for (int pl = 0; pl < P; pl++)
#pragma unroll
            for (int c = 0; c < 2; c++) dcu[pl][c] = djdxi_node<M>(a.acu, a.amu, nr, idr, pl, c, i);
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            d[pl] = rhs_bt_iter_own<M>(sc, dcu, a.cu, a.b_spe, relax_idr2, 0, pl, i);
            d[P + pl] = rhs_bt_iter_own<M>(sc, dcu, a.cu, a.b_spe, relax_idr2, 1, pl, i);
            d[2 * P + pl] = rhs_bz(sc, a.cu, pl, i);
            d[3 * P + pl] = rhs_ez(sc, a.cu, pl, i);
        }
#pragma unroll
        for (int pl = 0; pl < P; pl++) { FX(a.dcu, 2, i, pl, 0) = dcu[pl][0]; FX(a.dcu, 2, i, pl, 1) = dcu[pl][1]; }
    }
    double red[1] = {0.0};
    if (valid && i >= 2 && i <= nr - 2) red[0] = d[3 * P] * (double)(i - 1);
    double fa[NS], fb[NS];
    const int t = i - 1;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int kind = s < P ? FK_BPLUS : (s < 2 * P ? FK_BMINUS : (s < 3 * P ? FK_BZ : FK_EZ));
        const OpCoef &oc = a.ops[kind * (QPG_MAX_MODE + 1) + (((s % P) + 1) >> 1)];
        fa[s] = valid ? __ldg(oc.qT + t) * d[s] : 0.0;
        fb[s] = valid ? __ldg(oc.vT + t) * d[s] : 0.0;
    }
    team_scan<NS, 1>(fa, fb, red, sm, cluster);
    double mo = 0.0, mn = 0.0;
    if (valid) {
        double x[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int kind = s < P ? FK_BPLUS : (s < 2 * P ? FK_BMINUS : (s < 3 * P ? FK_BZ : FK_EZ));
            const OpCoef &oc = a.ops[kind * (QPG_MAX_MODE + 1) + (((s % P) + 1) >> 1)];
            x[s] = green_apply(oc, t, fa[s], fb[s], d[s]);
        }
        {
            const OpCoef &oc = a.ops[FK_EZ * (QPG_MAX_MODE + 1)];
            const double div = red[0] - sc.idrh * (FX(a.cu, 3, nr - 2, 0, 0) + FX(a.cu, 3, nr - 1, 0, 0)) * ((double)nr - 2.5);
            x[3 * P] += (-8.0 * div) * (__ldg(oc.pT + t) * __ldg(oc.qT));
        }
        // field_b_class.f03:703-758 get_solution_bt_iter, get_solution_bz, get_solution_ez ; compare ; b ; et
        const double ore = a.conv_old[i], oim = a.conv_old[nr + 2 + i];
        mo = ore * ore + oim * oim;
        double sre = 0.0, sim = 0.0;
        double obs[P][3], ob[P][3], oe[P][3];
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            const int m = (pl + 1) >> 1;
            double br, bp;
            if (m == 0) { br = (i == 1) ? 0.0 : x[0]; bp = (i == 1) ? 0.0 : x[P]; }
            else {
                const bool im = (pl & 1) == 0;
                const int po = im ? pl - 1 : pl + 1;
                br = 0.5 * (x[pl] + x[P + pl]);
                bp = im ? 0.5 * (-x[po] + x[P + po]) : 0.5 * (x[po] - x[P + po]);
                if (i == 1 && m != 1) { br = 0.0; bp = 0.0; }
            }
            double bz = x[2 * P + pl], ez = x[3 * P + pl];
            if (pl > 0 && i == 1) { bz = 0.0; ez = 0.0; }
            obs[pl][0] = br; obs[pl][1] = bp; obs[pl][2] = bz;
            oe[pl][2] = ez;
            const double v = fabs(bp);
            if (pl > 0 && (pl & 1) == 0) sim += v; else sre += v;
            ob[pl][0] = br + FX(a.b_beam, 3, i, pl, 0); ob[pl][1] = bp + FX(a.b_beam, 3, i, pl, 1); ob[pl][2] = bz + FX(a.b_beam, 3, i, pl, 2);
            et_node<M>(a.psi, nr, idr, pl, i, ob[pl][0], ob[pl][1], oe[pl][0], oe[pl][1]);
        }
#pragma unroll
        for (int pl = 0; pl < P; pl++)
#pragma unroll
            for (int c = 0; c < 3; c++) { FX(a.b_spe, 3, i, pl, c) = obs[pl][c]; FX(a.b, 3, i, pl, c) = ob[pl][c]; FX(a.e, 3, i, pl, c) = oe[pl][c]; }
        const double dre = ore - sre, dim = oim - sim;
        mn = dre * dre + dim * dim;
        a.conv_old[i] = sre; a.conv_old[nr + 2 + i] = sim;   // 'record' for the next pass (:373)
    }
    team_max2(mo, mn, smx, cluster);
    if (i == 1) {   // simulation_class.f03:560-599
        const double old_norm = sqrt(mo), abs_res = sqrt(mn);
        const double rel = old_norm > 2.220446049250313e-16 ? abs_res / old_norm : 1.7976931348623157e308;
        a.conv_out[0] = rel; a.conv_out[1] = abs_res;
        a.counters[1] += 1;
        const int it = a.flags[2] + 1;
        a.flags[2] = it;
        const bool fin = rel < a.reltol || abs_res < a.abstol || it >= a.iter_max;
        if (fin) a.flags[0] = 1;
        if (a.cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)a.cond_handle, fin ? 0u : 1u);
    }
    cluster.sync();  // keep every CTA's shared memory alive until all remote reads are done
}

// ============================================================================================================
// program D for one (node, plane) item, guards included; all loads first, then all stores
template <int M>
__device__ __forceinline__ void fused_D_item(const FusedArgs &a, int idx, int j)
{
    constexpr int P = 2 * M + 1;
    const int nr = a.nr;
    const int n = idx / P, pl = idx - n * P;
    const size_t np = (size_t)idx, n1 = (size_t)(nr + 2) * P, sl = (size_t)(j - 1) * n1 + np;
    const double idr = 1.0 / a.dr;
    double cu[3], dcu[2], bs[3], es[3], e[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { cu[c] = a.cu[np * 3 + c]; bs[c] = a.b_spe[np * 3 + c]; e[c] = a.e[np * 3 + c]; b[c] = a.b[np * 3 + c]; }
    es[2] = a.e_spe[np * 3 + 2]; es[0] = a.e_spe[np * 3]; es[1] = a.e_spe[np * 3 + 1];
    dcu[0] = a.dcu[np * 2]; dcu[1] = a.dcu[np * 2 + 1];
    const double psi = a.psi[np];
    const double sq = a.spe_q[np] + a.spe_cu[np * 3 + 2];                                 // cbq, species2d :396
    const double qs = a.q_spe[np] + cu[2];                                                // :410
    if (n >= 1 && n <= nr) et_node<M>(a.psi, nr, idr, pl, n, bs[0], bs[1], es[0], es[1]);   // e_spe%solve(b_spe, psi) :414
    dcu[0] *= a.dxi; dcu[1] *= a.dxi;                                                     // :425
    a.spe_q[np] = sq; a.spe_q2[sl] = sq;
    a.q_spe[np] = qs; a.q_spe2[sl] = qs;                                                  // :411
    a.dcu[np * 2] = dcu[0]; a.dcu[np * 2 + 1] = dcu[1];
    a.cu[np * 3] = cu[0] + dcu[0]; a.cu[np * 3 + 1] = cu[1] + dcu[1];                     // :426
    a.psi2[sl] = psi;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        a.cu2[sl * 3 + c] = cu[c];                                                        // :409
        a.e_spe[np * 3 + c] = es[c];
        a.e_spe2[sl * 3 + c] = es[c];                                                     // :452-456
        a.b_spe2[sl * 3 + c] = bs[c];
        a.e2[sl * 3 + c] = e[c];
        a.b2[sl * 3 + c] = b[c];
    }
}
template <int M>
__global__ void __launch_bounds__(FT) k_fused_D(const __grid_constant__ FusedArgs a)
{
    const int idx = blockIdx.x * FT + threadIdx.x;
    if (idx >= (a.nr + 2) * (2 * M + 1)) return;
    fused_D_item<M>(a, idx, a.flags[3]);
}

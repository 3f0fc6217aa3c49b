// laser.cu -- the laser-envelope (ponderomotive guiding centre) field path, SURVEY.md §8(f) rank 1.
//
//   laser/field_laser_class.f03:269-391  init_solver          -> laser_rows + pcr_factor (host, long double)
//   laser/field_laser_class.f03:393-635  set_rhs_field_laser   -> k_laser_set_rhs
//   laser/field_laser_class.f03:637-750  set_grad + copy_slice + gather (simulation_class.f03:361-366) -> k_laser_slice
//   laser/field_laser_class.f03:752-927  solve_field_laser     -> k_laser_solve (ONE persistent CTA sweeps all xi slices)
//   species/part2d_class.f03:361-476     deposit_chi_part2d    -> k_deposit_chi + k_chi_fix
//   sim_lasers_class.f03:175-222         deposit_chi / advance -> qpg_laser_deposit_chi / qpg_laser_advance
//
// The envelope a = a_r + i a_i obeys (i k0 + d/dxi) da/ds = lap_perp(a)/2 + chi a / 4 ...; per xi slice and azimuthal plane
// the reference solves a 2nr x 2nr pentadiagonal system (unknowns a_r, a_i interleaved) with a parallel cyclic reduction
// whose coefficients are generated once (pcr-fortran/fpcr_penta_class.f03:270).  Here the same system is written as a
// BLOCK-tridiagonal one (2x2 blocks coupling the (a_r, a_i) pairs of neighbouring radial nodes) and reduced by parallel
// cyclic reduction over the nodes: the operator never changes, so the elimination matrices of every level are
// pre-computed on the host in long double and a solve only transforms the right-hand side -- ceil(log2 nr) levels of
// two 2x2 mat-vecs per node, one thread per node, the vectors in shared memory.  The xi recurrence (slice j needs the
// NEW slices j-1, j-2) is inherently sequential: one persistent CTA walks the slices, so a 3D step costs one launch.
//
// Device layout of the envelope volumes (a_r, a_i, s_r, s_i): v[plane][slice j = -1..nz+1 at index j+1][node 0..nr+1]
// (node contiguous: one thread per node reads/writes coalesced).  Host layout of upload / download is the same.
#include "common.cuh"

struct qpg_laser_s {
    qpg_ctx ctx;
    int nz, iter, nsteps, nthreads;
    double k0, ds, dz;
    size_t nvol;
    double *ar, *ai, *sr, *si;
    qpg_field f_ar, f_ai, f_gr, f_gi, chi;
    double *chi_acc;     // raw susceptibility sums [(nr+2)][P]
    double *pcr;         // [M+1][nsteps][8][nr] elimination matrices (alpha 2x2 | gamma 2x2, row-major), then [M+1][4][nr] inverse diagonal blocks
    double *pcrc;        // the same as complex numbers (every block of this operator is [[x, -y], [y, x]]): [M+1][nsteps][4][nr] (alpha x, y | gamma x, y), then
                         // [M+1][2][nr] inverse diagonal (x, y); null if the structure check failed (then the general kernel runs)
    size_t pcrc_n;
    // overlapped advance (qpg_laser_advance_overlapped): the solve runs on a side stream and publishes the number of finished slices
    cudaStream_t side;           // created lazily
    cudaEvent_t ev_main, ev_done;
    unsigned *progress;          // device word: (advances launched - 1) * nz + slices finished by the running solve
    unsigned adv_count;          // overlapped advances launched so far
    bool pending;                // an overlapped advance may still be running: join before anybody else touches the envelope
    // xi-pipeline hand-off of the envelope (qpg_laser_set_handoff): wire buffers and message counters of the two links of this stage
    const double *g_in; unsigned *g_in_ready, *g_in_ack;
    double *g_out; unsigned *g_out_ready, *g_out_ack;
    unsigned g_seq;              // advances done with a hand-off = number of the last message on either link
};

#define LVI(pl, i, j) ((((size_t)(pl)) * (nz + 3) + (size_t)((j) + 1)) * (nr + 2) + (i))

// ---- host: operator rows and their cyclic-reduction factors ------------------------------------------------------
typedef long double ld;
struct M2 { ld a, b, c, d; };   // [[a, b], [c, d]]
static M2 m2mul(const M2 &x, const M2 &y) { return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
static M2 m2add(const M2 &x, const M2 &y) { return {x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
static M2 m2neg(const M2 &x) { return {-x.a, -x.b, -x.c, -x.d}; }
static M2 m2inv(const M2 &x) { const ld det = x.a * x.d - x.b * x.c; return {x.d / det, -x.b / det, -x.c / det, x.a / det}; }

// block rows of mode m (field_laser_class.f03:269-391): A couples to node i-1, B is the diagonal block, C couples to node i+1
static void laser_rows(int m, int nr, double k0, double ds, double dr, double dz, std::vector<M2> &A, std::vector<M2> &B, std::vector<M2> &C)
{
    const ld q = 0.25L * ds, h = 1.5L * dr * dr / dz, kap = (ld)k0 * dr * dr, m2 = (ld)m * m;
    A.assign(nr, {0, 0, 0, 0}); B.assign(nr, {0, 0, 0, 0}); C.assign(nr, {0, 0, 0, 0});
    for (int i = 0; i < nr; i++) {          // node i+1, radius i*dr
        if (i == 0) {
            if (m == 0) { B[i] = {ds + h, -kap, kap, ds + h}; C[i] = {-(ld)ds, 0, 0, -(ld)ds}; }   // :318-340
            else B[i] = {1, 0, 0, 1};                                                               // decoupled, :325-331
            continue;
        }
        const ld j = i, lo = -q * (1.0L - 0.5L / j), hi = -q * (1.0L + 0.5L / j), c = q * (2.0L + m2 / (j * j)) + h;
        A[i] = {lo, 0, 0, lo};
        B[i] = {c, -kap, kap, c};
        if (i < nr - 1) C[i] = {hi, 0, 0, hi};   // outer rows: e = 0 (:356-373)
    }
}
static void pcr_factor(int m, int nr, int nsteps, double k0, double ds, double dr, double dz, double *lvl /*[nsteps][8][nr]*/, double *binv /*[4][nr]*/)
{
    std::vector<M2> A, B, C;
    laser_rows(m, nr, k0, ds, dr, dz, A, B, C);
    for (int s = 0; s < nsteps; s++) {
        const int hstep = 1 << s;
        std::vector<M2> A2(nr), B2(nr), C2(nr);
        for (int i = 0; i < nr; i++) {
            M2 al = {0, 0, 0, 0}, ga = {0, 0, 0, 0};
            B2[i] = B[i]; A2[i] = {0, 0, 0, 0}; C2[i] = {0, 0, 0, 0};
            if (i - hstep >= 0) { al = m2neg(m2mul(A[i], m2inv(B[i - hstep]))); A2[i] = m2mul(al, A[i - hstep]); B2[i] = m2add(B2[i], m2mul(al, C[i - hstep])); }
            if (i + hstep < nr) { ga = m2neg(m2mul(C[i], m2inv(B[i + hstep]))); C2[i] = m2mul(ga, C[i + hstep]); B2[i] = m2add(B2[i], m2mul(ga, A[i + hstep])); }
            double *o = lvl + (size_t)s * 8 * nr + i;
            o[0 * nr] = (double)al.a; o[1 * nr] = (double)al.b; o[2 * nr] = (double)al.c; o[3 * nr] = (double)al.d;
            o[4 * nr] = (double)ga.a; o[5 * nr] = (double)ga.b; o[6 * nr] = (double)ga.c; o[7 * nr] = (double)ga.d;
        }
        A.swap(A2); B.swap(B2); C.swap(C2);
    }
    for (int i = 0; i < nr; i++) {
        const M2 bi = m2inv(B[i]);
        binv[0 * nr + i] = (double)bi.a; binv[1 * nr + i] = (double)bi.b; binv[2 * nr + i] = (double)bi.c; binv[3 * nr + i] = (double)bi.d;
    }
}

// ---- device ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lz_re(int m) { return m == 0 ? 0 : 2 * m - 1; }
__device__ __forceinline__ int lz_im(int m) { return 2 * m; }

// plasma-susceptibility couplings of one node (field_laser_class.f03:563-631 = :787-852): t += ds/4 dr^2 (chi * a)_m
__device__ __forceinline__ void chi_coupling(const double *ar, const double *ai, const double *chi, int M, double w, double *tr, double *ti)
{
    for (int m = 0; m <= M; m++) {
        double rr = 0.0, ri = 0.0, ir = 0.0, ii = 0.0;   // (t_r re, t_r im, t_i re, t_i im) of mode m
        for (int k = m - M; k <= M; k++) {
            const int ak = k < 0 ? -k : k, amk = m - k < 0 ? k - m : m - k;
            rr += w * chi[lz_re(ak)] * ar[lz_re(amk)];
            ir += w * chi[lz_re(ak)] * ai[lz_re(amk)];
            if (k == 0 || k == m) continue;
            const double sg = (k >= 0 && k <= m) ? 1.0 : -1.0;
            rr -= w * sg * chi[lz_im(ak)] * ar[lz_im(amk)];
            ir -= w * sg * chi[lz_im(ak)] * ai[lz_im(amk)];
        }
        tr[lz_re(m)] = rr; ti[lz_re(m)] = ir;
        if (m == 0) continue;
        for (int k = m - M; k <= M; k++) {
            const int ak = k < 0 ? -k : k, amk = m - k < 0 ? k - m : m - k;
            if (k != 0) {
                const double sg = k < 0 ? -1.0 : 1.0;
                ri += w * sg * chi[lz_im(ak)] * ar[lz_re(amk)];
                ii += w * sg * chi[lz_im(ak)] * ai[lz_re(amk)];
            }
            if (k != m) {
                const double sg = k > m ? -1.0 : 1.0;
                ri += w * sg * chi[lz_re(ak)] * ar[lz_im(amk)];
                ii += w * sg * chi[lz_re(ak)] * ai[lz_im(amk)];
            }
        }
        tr[lz_im(m)] = ri; ti[lz_im(m)] = ii;
    }
}

// explicit half from the OLD envelope, one thread per (node 1..nr, slice 1..nz)
__global__ void k_laser_set_rhs(const double *__restrict__ ar, const double *__restrict__ ai, const double *__restrict__ chi2, double *__restrict__ sr,
                                double *__restrict__ si, int nr, int nz, int M, double k0, double ds, double dr, double dz)
{
    const int P = 2 * M + 1;
    const long n = (long)nr * nz;
    const double dr2_idzh = 0.5 * dr * dr / dz, kappa = k0 * dr * dr, ds_qtr = 0.25 * ds, w = ds_qtr * dr * dr;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nr) + 1, j = (int)(t / nr) + 1;
        double a_r[2 * QPG_MAX_MODE + 1], a_i[2 * QPG_MAX_MODE + 1], ch[2 * QPG_MAX_MODE + 1], tr[2 * QPG_MAX_MODE + 1], ti[2 * QPG_MAX_MODE + 1];
        for (int pl = 0; pl < P; pl++) {
            a_r[pl] = ar[LVI(pl, i, j)]; a_i[pl] = ai[LVI(pl, i, j)];
            ch[pl] = chi2[(size_t)(j - 1) * (nr + 2) * P + (size_t)i * P + pl];
        }
        chi_coupling(a_r, a_i, ch, M, w, tr, ti);
        for (int pl = 0; pl < P; pl++) {
            const int m = (pl + 1) / 2;
            double beta_m, beta_p, alpha, vr, vi;
            if (i == 1 && m > 0) { vr = 0.0; vi = 0.0; }                                       // :507-512
            else {
                if (i == 1) { beta_m = 0.0; beta_p = ds; alpha = -ds; }                        // :439-444
                else {
                    const double ik = 1.0 / (double)(i - 1);
                    beta_m = ds_qtr * (1.0 - 0.5 * ik);
                    beta_p = i == nr ? 0.0 : ds_qtr * (1.0 + 0.5 * ik);
                    alpha = -ds_qtr * (2.0 + (double)(m * m) * ik * ik);
                }
                vr = dr2_idzh * (3.0 * a_r[pl] - 4.0 * ar[LVI(pl, i, j - 1)] + ar[LVI(pl, i, j - 2)]) - kappa * a_i[pl]
                     + beta_m * ar[LVI(pl, i - 1, j)] + alpha * a_r[pl] + beta_p * ar[LVI(pl, i + 1, j)];
                vi = dr2_idzh * (3.0 * a_i[pl] - 4.0 * ai[LVI(pl, i, j - 1)] + ai[LVI(pl, i, j - 2)]) + kappa * a_r[pl]
                     + beta_m * ai[LVI(pl, i - 1, j)] + alpha * a_i[pl] + beta_p * ai[LVI(pl, i + 1, j)];
            }
            sr[LVI(pl, i, j)] = vr + tr[pl];
            si[LVI(pl, i, j)] = vi + ti[pl];
        }
    }
}

// implicit half: ONE CTA, thread t <-> node t+1; slices in order, `iter` fixed-point passes over the chi coupling per slice
__global__ void __launch_bounds__(1024, 1) k_laser_solve(double *__restrict__ ar, double *__restrict__ ai, const double *__restrict__ sr, const double *__restrict__ si,
                                                        const double *__restrict__ chi2, const double *__restrict__ pcr, int nr, int nz, int M, int iter,
                                                        int nsteps, double ds, double dr, double dz)
{
    extern __shared__ double laser_sh[];   // [2 buffers][2 components][nthreads]
    const int P = 2 * M + 1, t = threadIdx.x, nt = blockDim.x, i = t + 1;
    const bool live = t < nr;
    const double dr2_idzh = 0.5 * dr * dr / dz, w = 0.25 * ds * dr * dr;
    const double *binv_all = pcr + (size_t)(M + 1) * nsteps * 8 * nr;
    int cur = 0;   // the two exchange buffers alternate over ALL levels of all solves: a buffer is rewritten only after the barrier of the level in between
    for (int j = 1; j <= nz; j++)
        for (int l = 0; l < iter; l++) {
            double a_r[2 * QPG_MAX_MODE + 1], a_i[2 * QPG_MAX_MODE + 1], ch[2 * QPG_MAX_MODE + 1], tr[2 * QPG_MAX_MODE + 1], ti[2 * QPG_MAX_MODE + 1];
            if (live) {
                for (int pl = 0; pl < P; pl++) {
                    a_r[pl] = ar[LVI(pl, i, j)]; a_i[pl] = ai[LVI(pl, i, j)];
                    ch[pl] = chi2[(size_t)(j - 1) * (nr + 2) * P + (size_t)i * P + pl];
                }
                chi_coupling(a_r, a_i, ch, M, w, tr, ti);
            }
            for (int pl = 0; pl < P; pl++) {
                const int m = (pl + 1) / 2;
                double d0 = 0.0, d1 = 0.0;
                if (live) {
                    d0 = sr[LVI(pl, i, j)] + tr[pl] + dr2_idzh * (4.0 * ar[LVI(pl, i, j - 1)] - ar[LVI(pl, i, j - 2)]);
                    d1 = si[LVI(pl, i, j)] + ti[pl] + dr2_idzh * (4.0 * ai[LVI(pl, i, j - 1)] - ai[LVI(pl, i, j - 2)]);
                }
                const double *lv = pcr + (size_t)m * nsteps * 8 * nr;
                for (int s = 0; s < nsteps; s++) {
                    double *buf = laser_sh + (size_t)cur * 2 * nt;
                    buf[t] = d0; buf[nt + t] = d1;
                    __syncthreads();
                    if (live) {
                        const int hs = 1 << s;
                        const double *c = lv + (size_t)s * 8 * nr + t;
                        if (t - hs >= 0) {
                            const double x0 = buf[t - hs], x1 = buf[nt + t - hs];
                            d0 += c[0] * x0 + c[(size_t)nr] * x1;
                            d1 += c[2 * (size_t)nr] * x0 + c[3 * (size_t)nr] * x1;
                        }
                        if (t + hs < nr) {
                            const double x0 = buf[t + hs], x1 = buf[nt + t + hs];
                            d0 += c[4 * (size_t)nr] * x0 + c[5 * (size_t)nr] * x1;
                            d1 += c[6 * (size_t)nr] * x0 + c[7 * (size_t)nr] * x1;
                        }
                    }
                    cur ^= 1;   // the next level writes the other buffer: no second barrier needed
                }
                if (live) {
                    const double *bi = binv_all + (size_t)m * 4 * nr + t;
                    double xr = bi[0] * d0 + bi[(size_t)nr] * d1, xi = bi[2 * (size_t)nr] * d0 + bi[3 * (size_t)nr] * d1;
                    if (m > 0 && i == 1) { xr = 0.0; xi = 0.0; }      // :913-918 (the decoupled axis rows give 0 anyway)
                    ar[LVI(pl, i, j)] = xr; ai[LVI(pl, i, j)] = xi;
                }
            }
            __syncthreads();   // slice j (and this pass) complete before anyone reads it as j-1 / in the next pass
        }
}

// The same solve with the elimination matrices as COMPLEX numbers (the 2x2 blocks of this operator all have the form [[x, -y], [y, x]]:
// a_r + i a_i is multiplied by x + i y) and resident in shared memory when they fit (nr = 512, one mode: 147 KB): the general kernel
// above re-read 8 doubles per node and level from L2 on every level of every pass (the 295 KB of one mode do not stay in L1) and spent
// ~1350 cycles per level on that latency -- 18.7 us per slice, 9.6 ms per 3D step of config 4 with the other 147 SMs idle.  Per slice the
// pass-invariant inputs are loaded once, a(j-1) and a(j-2) stay in registers from the previous slices, the iterate stays in registers
// between the fixed-point passes.
template <int MM>
__global__ void __launch_bounds__(1024, 1) k_laser_solve_c(double *__restrict__ ar, double *__restrict__ ai, const double *__restrict__ sr, const double *__restrict__ si,
                                                          const double *__restrict__ chi2, const double *__restrict__ pcrc, int nr, int nz, int iter, int nsteps,
                                                          double ds, double dr, double dz, int coef_in_smem, unsigned *progress, unsigned progress_base)
{
    constexpr int M = MM, P = 2 * M + 1;
    extern __shared__ double laser_sh[];   // [2 buffers][2 components][nthreads] exchange, then (optionally) the coefficients
    const int t = threadIdx.x, nt = blockDim.x, i = t + 1;
    const bool live = t < nr;
    const double dr2_idzh = 0.5 * dr * dr / dz, w = 0.25 * ds * dr * dr;
    const size_t nlv = (size_t)(M + 1) * nsteps * 4 * nr, nbi = (size_t)(M + 1) * 2 * nr;
    const double *coef = pcrc;
    if (coef_in_smem) {
        double *dst = laser_sh + 4 * (size_t)nt;
        for (size_t k = t; k < nlv + nbi; k += nt) dst[k] = pcrc[k];
        coef = dst;
        __syncthreads();
    }
    const double *binv_all = coef + nlv;
    double am1r[P], am1i[P], am2r[P], am2i[P];      // the new a(j-1), a(j-2) of this thread's node
#pragma unroll
    for (int pl = 0; pl < P; pl++) {
        am1r[pl] = live ? ar[LVI(pl, i, 0)] : 0.0; am1i[pl] = live ? ai[LVI(pl, i, 0)] : 0.0;
        am2r[pl] = live ? ar[LVI(pl, i, -1)] : 0.0; am2i[pl] = live ? ai[LVI(pl, i, -1)] : 0.0;
    }
    int cur = 0;
    for (int j = 1; j <= nz; j++) {
        double a_r[P], a_i[P], ch[P], s_r[P], s_i[P];
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            a_r[pl] = live ? ar[LVI(pl, i, j)] : 0.0; a_i[pl] = live ? ai[LVI(pl, i, j)] : 0.0;       // old envelope = first guess
            ch[pl] = live ? chi2[(size_t)(j - 1) * (nr + 2) * P + (size_t)i * P + pl] : 0.0;
            s_r[pl] = live ? sr[LVI(pl, i, j)] + dr2_idzh * (4.0 * am1r[pl] - am2r[pl]) : 0.0;
            s_i[pl] = live ? si[LVI(pl, i, j)] + dr2_idzh * (4.0 * am1i[pl] - am2i[pl]) : 0.0;
        }
        for (int l = 0; l < iter; l++) {
            double tr[P], ti[P];
            chi_coupling(a_r, a_i, ch, M, w, tr, ti);
#pragma unroll
            for (int pl = 0; pl < P; pl++) {
                const int m = (pl + 1) / 2;
                double d0 = s_r[pl] + tr[pl], d1 = s_i[pl] + ti[pl];
                const double *lv = coef + (size_t)m * nsteps * 4 * nr;
                for (int s = 0; s < nsteps; s++) {
                    double *buf = laser_sh + (size_t)cur * 2 * nt;
                    buf[t] = d0; buf[nt + t] = d1;
                    __syncthreads();
                    if (live) {
                        const int hs = 1 << s;
                        const double *c = lv + (size_t)s * 4 * nr + t;
                        if (t - hs >= 0) {
                            const double x0 = buf[t - hs], x1 = buf[nt + t - hs], cx = c[0], cy = c[(size_t)nr];
                            d0 += cx * x0 - cy * x1;
                            d1 += cy * x0 + cx * x1;
                        }
                        if (t + hs < nr) {
                            const double x0 = buf[t + hs], x1 = buf[nt + t + hs], cx = c[2 * (size_t)nr], cy = c[3 * (size_t)nr];
                            d0 += cx * x0 - cy * x1;
                            d1 += cy * x0 + cx * x1;
                        }
                    }
                    cur ^= 1;
                }
                if (live) {
                    const double *bi = binv_all + (size_t)m * 2 * nr + t;
                    const double bx = bi[0], by = bi[(size_t)nr];
                    double xr = bx * d0 - by * d1, xi = by * d0 + bx * d1;
                    if (m > 0 && i == 1) { xr = 0.0; xi = 0.0; }
                    a_r[pl] = xr; a_i[pl] = xi;
                }
            }
        }
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
            if (live) { ar[LVI(pl, i, j)] = a_r[pl]; ai[LVI(pl, i, j)] = a_i[pl]; }
            am2r[pl] = am1r[pl]; am2i[pl] = am1i[pl]; am1r[pl] = a_r[pl]; am1i[pl] = a_i[pl];
        }
        if (progress) {   // slice j is final: a sweep kernel of the NEXT 3D step that runs beside this solve may read it (sweep.cu sweep_laser_wait)
            __syncthreads();
            if (t == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(progress), "r"(progress_base + (unsigned)j) : "memory"); }
        }
    }
}
template <int MM> static void l_laser_solve_c(int nthreads, size_t smem, cudaStream_t st, double *ar, double *ai, const double *sr, const double *si, const double *chi2,
                                              const double *pcrc, int nr, int nz, int iter, int nsteps, double ds, double dr, double dz, int in_smem, unsigned *progress,
                                              unsigned progress_base)
{
    static size_t attr = 0;
    if (smem > attr) { cudaFuncSetAttribute(k_laser_solve_c<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = smem; }
    k_laser_solve_c<MM><<<1, nthreads, smem, st>>>(ar, ai, sr, si, chi2, pcrc, nr, nz, iter, nsteps, ds, dr, dz, in_smem, progress, progress_base);
}

// copy_slice(j, 2to1) + set_grad(j) + gather into the four slice images the pgc pushers read; one thread per (node, plane)
__global__ void k_laser_slice(const double *__restrict__ ar, const double *__restrict__ ai, int nr, int nz, int M, int j, const int *__restrict__ jflag, double dr,
                              double dz, double *__restrict__ f_ar, double *__restrict__ f_ai, double *__restrict__ f_gr, double *__restrict__ f_gi)
{
    if (jflag) j = jflag[3];   // the sim's device-side slice counter (CUDA-graph replay has no host-side j)
    if (j < 1 || j > nz) return;
    const int P = 2 * M + 1, n = (nr + 2) * P;
    const double idrh = 0.5 / dr, idzh = 0.5 / dz;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int i = k / P, pl = k % P, m = (pl + 1) / 2;
        const bool is_im = m > 0 && pl == 2 * m;
        const int other = m == 0 ? pl : (is_im ? pl - 1 : pl + 1);   // the partner plane (re <-> im) of the same mode
        f_ar[k] = ar[LVI(pl, i, j)];
        f_ai[k] = ai[LVI(pl, i, j)];
        double *gr = f_gr + (size_t)k * 3, *gi = f_gi + (size_t)k * 3;
        if (i >= 1 && i <= nr) {
            gr[2] = idzh * (3.0 * ar[LVI(pl, i, j)] - 4.0 * ar[LVI(pl, i, j - 1)] + ar[LVI(pl, i, j - 2)]);
            gi[2] = idzh * (3.0 * ai[LVI(pl, i, j)] - 4.0 * ai[LVI(pl, i, j - 1)] + ai[LVI(pl, i, j - 2)]);
        }
        if (i >= 2 && i <= nr) {
            gr[0] = idrh * (ar[LVI(pl, i + 1, j)] - ar[LVI(pl, i - 1, j)]);
            gi[0] = idrh * (ai[LVI(pl, i + 1, j)] - ai[LVI(pl, i - 1, j)]);
            if (m == 0) { gr[1] = 0.0; gi[1] = 0.0; }
            else {
                const double ir = 1.0 / ((double)(i - 1) * dr), sg = is_im ? 1.0 : -1.0;   // re: -(m/r) Im ; im: +(m/r) Re
                gr[1] = sg * ir * m * ar[LVI(other, i, j)];
                gi[1] = sg * ir * m * ai[LVI(other, i, j)];
            }
        }
        if (m == 0 && i == 1) { gr[0] = 0.0; gr[1] = 0.0; gi[0] = 0.0; gi[1] = 0.0; }
        // reference quirk (field_laser_class.f03:708-730): the m > 0 axis rules are written with the loop variable AFTER the
        // `do i = 2, nrp` loop, i.e. into node nrp+1; node 1 keeps components 1, 2 (zero since creation)
        if (m > 0 && i == nr + 1) {
            if (m % 2 == 1) {
                const double g1r = 2.0 * idrh * ar[LVI(pl, 2, j)], g1i = 2.0 * idrh * ai[LVI(pl, 2, j)];
                const double o1r = 2.0 * idrh * ar[LVI(other, 2, j)], o1i = 2.0 * idrh * ai[LVI(other, 2, j)];
                gr[0] = g1r; gi[0] = g1i;
                gr[1] = is_im ? m * o1r : -m * o1r;      // re: -m * grad_im(1) ; im: m * grad_re(1)
                gi[1] = is_im ? m * o1i : -m * o1i;
            } else { gr[0] = 0.0; gr[1] = 0.0; gi[0] = 0.0; gi[1] = 0.0; }
        }
    }
}

// species/part2d_class.f03:361-430: raw sums of -qbm q / (1 - qbm psi) e^{-im phi}, linear weights
template <int M>
__global__ void k_deposit_chi(PartView pv, double *__restrict__ acc, double idr, double qbm)
{
    constexpr int P = 2 * M + 1;
    const int npp = *pv.d_npp;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npp; p += gridDim.x * blockDim.x) {
        const double x1 = pv.x1[p], x2 = pv.x2[p];
        double pos = __dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2))), idr);
        const double c0 = x1 / pos * idr, s0 = -x2 / pos * idr;
        const int nn = (int)floor(pos);
        const double f = pos - (double)nn, w0 = 1.0 - f, w1 = f;
        double phr = -1.0 * qbm * pv.q[p] / (1.0 - qbm * pv.psi[p]), phi = 0.0;
        double *a0 = acc + (size_t)(nn + 1) * P, *a1 = a0 + P;
        atomicAdd(a0, w0 * phr); atomicAdd(a1, w1 * phr);
#pragma unroll
        for (int m = 1; m <= M; m++) {
            const double t = phr * c0 - phi * s0;
            phi = phr * s0 + phi * c0;
            phr = t;
            atomicAdd(a0 + 2 * m - 1, w0 * phr); atomicAdd(a1 + 2 * m - 1, w1 * phr);
            atomicAdd(a0 + 2 * m, w0 * phi); atomicAdd(a1 + 2 * m, w1 * phi);
        }
    }
}
// :432-452 axis rules (get_deposit_ax_corr instead of the charge deposit's 8) and 1/(j-1); chi = fix(raw), raw cleared;
// the slice image also goes into the chi volume (sim_lasers_class.f03:191 copy_slice 1to2)
__global__ void k_chi_fix(double *__restrict__ acc, double *__restrict__ chi1, double *__restrict__ chi2, int nz, int j, const int *__restrict__ jflag, int nr, int P,
                          double ax_corr)
{
    const int n = (nr + 2) * P;
    if (jflag) j = jflag[3];
    double *chi2_slice = (j >= 1 && j <= nz) ? chi2 + (size_t)(j - 1) * n : nullptr;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int j = k / P, pl = k % P;
        double v = acc[k];
        acc[k] = 0.0;
        if (j == 0) v = 0.0;
        else if (j == 1) v = pl == 0 ? v * ax_corr : 0.0;
        else v = v * (1.0 / (double)(j - 1));
        chi1[k] = v;
        if (chi2_slice) chi2_slice[k] = v;
    }
}
template <int M> static void l_deposit_chi(int grid, cudaStream_t st, PartView pv, double *acc, double idr, double qbm) { k_deposit_chi<M><<<grid, 256, 0, st>>>(pv, acc, idr, qbm); }

// ---- C-ABI -------------------------------------------------------------------------------------------------------
extern "C" int qpg_laser_create(qpg_laser *out, qpg_ctx ctx, int nz, double k0, double ds, int iter)
{
    ARG_TRY(out && ctx && nz >= 1 && iter >= 1, "bad arg");
    if (ctx->nr > 1024) { qpg_set_error("the envelope solver keeps one thread per radial node in one CTA: nr <= 1024"); return QPG_ERR_UNSUPPORTED; }
    qpg_laser l = new qpg_laser_s();
    memset(l, 0, sizeof(*l));
    l->ctx = ctx; l->nz = nz; l->iter = iter; l->k0 = k0; l->ds = ds; l->dz = ctx->dxi;
    const int nr = ctx->nr, P = ctx->P, M = ctx->M;
    l->nvol = (size_t)P * (nz + 3) * (nr + 2);
    double **vol[4] = {&l->ar, &l->ai, &l->sr, &l->si};
    for (auto v : vol) { CUDA_TRY(cudaMalloc(v, sizeof(double) * l->nvol)); CUDA_TRY(cudaMemsetAsync(*v, 0, sizeof(double) * l->nvol, ctx->stream)); }
    int rc;
    if ((rc = qpg_field_create(&l->f_ar, ctx, 1, 0, 0)) || (rc = qpg_field_create(&l->f_ai, ctx, 1, 0, 0)) || (rc = qpg_field_create(&l->f_gr, ctx, 3, 0, 0)) ||
        (rc = qpg_field_create(&l->f_gi, ctx, 3, 0, 0)) || (rc = qpg_field_create(&l->chi, ctx, 1, nz, 1))) return rc;
    CUDA_TRY(cudaMalloc(&l->chi_acc, sizeof(double) * (size_t)(nr + 2) * P));
    CUDA_TRY(cudaMemsetAsync(l->chi_acc, 0, sizeof(double) * (size_t)(nr + 2) * P, ctx->stream));
    l->nsteps = 0;
    while ((1 << l->nsteps) < nr) l->nsteps++;
    l->nthreads = (nr + 31) / 32 * 32;
    const size_t nlv = (size_t)(M + 1) * l->nsteps * 8 * nr, nbi = (size_t)(M + 1) * 4 * nr;
    std::vector<double> h(nlv + nbi);
    for (int m = 0; m <= M; m++) pcr_factor(m, nr, l->nsteps, k0, ds, ctx->dr, l->dz, h.data() + (size_t)m * l->nsteps * 8 * nr, h.data() + nlv + (size_t)m * 4 * nr);
    CUDA_TRY(cudaMalloc(&l->pcr, sizeof(double) * (nlv + nbi)));
    CUDA_TRY(cudaMemcpy(l->pcr, h.data(), sizeof(double) * (nlv + nbi), cudaMemcpyHostToDevice));
    {   // the complex form of the same factors (k_laser_solve_c): valid iff every block is [[x, -y], [y, x]]
        const size_t nlc = (size_t)(M + 1) * l->nsteps * 4 * nr, nbc = (size_t)(M + 1) * 2 * nr;
        std::vector<double> hc(nlc + nbc);
        bool ok = true;
        double scale = 0.0;
        for (double v : h) scale = fabs(v) > scale ? fabs(v) : scale;
        for (int m = 0; m <= M && ok; m++) {
            for (int sstep = 0; sstep < l->nsteps && ok; sstep++)
                for (int half = 0; half < 2; half++)
                    for (int i = 0; i < nr; i++) {
                        const double *o = h.data() + (size_t)m * l->nsteps * 8 * nr + (size_t)sstep * 8 * nr + (size_t)half * 4 * nr + i;
                        const double a = o[0], b = o[(size_t)nr], cc = o[2 * (size_t)nr], d = o[3 * (size_t)nr];
                        if (fabs(a - d) > 1e-13 * scale || fabs(b + cc) > 1e-13 * scale) ok = false;
                        double *q = hc.data() + (size_t)m * l->nsteps * 4 * nr + (size_t)sstep * 4 * nr + (size_t)half * 2 * nr + i;
                        q[0] = a; q[(size_t)nr] = cc;
                    }
            for (int i = 0; i < nr; i++) {
                const double *o = h.data() + nlv + (size_t)m * 4 * nr + i;
                if (fabs(o[0] - o[3 * (size_t)nr]) > 1e-13 * scale || fabs(o[(size_t)nr] + o[2 * (size_t)nr]) > 1e-13 * scale) ok = false;
                double *q = hc.data() + nlc + (size_t)m * 2 * nr + i;
                q[0] = o[0]; q[(size_t)nr] = o[2 * (size_t)nr];
            }
        }
        if (ok && M <= 2) {
            CUDA_TRY(cudaMalloc(&l->pcrc, sizeof(double) * (nlc + nbc)));
            CUDA_TRY(cudaMemcpy(l->pcrc, hc.data(), sizeof(double) * (nlc + nbc), cudaMemcpyHostToDevice));
            l->pcrc_n = nlc + nbc;
        }
    }
    *out = l;
    return 0;
}
extern "C" int qpg_laser_destroy(qpg_laser l)
{
    if (!l) return 0;
    cudaStreamSynchronize(l->ctx->stream);
    cudaFree(l->ar); cudaFree(l->ai); cudaFree(l->sr); cudaFree(l->si); cudaFree(l->chi_acc); cudaFree(l->pcr); cudaFree(l->pcrc);
    if (l->progress) { cudaStreamSynchronize(l->side); cudaStreamDestroy(l->side); cudaEventDestroy(l->ev_main); cudaEventDestroy(l->ev_done); cudaFree(l->progress); }
    qpg_field_destroy(l->f_ar); qpg_field_destroy(l->f_ai); qpg_field_destroy(l->f_gr); qpg_field_destroy(l->f_gi); qpg_field_destroy(l->chi);
    delete l;
    return 0;
}
extern "C" long qpg_laser_volume_size(qpg_laser l) { return l ? (long)l->nvol : -1; }
// an overlapped advance may still be running on the side stream: everything else that touches the envelope orders itself behind it
static int laser_join(qpg_laser l)
{
    if (l && l->pending) { CUDA_TRY(cudaStreamWaitEvent(l->ctx->stream, l->ev_done, 0)); l->pending = false; }
    return 0;
}
extern "C" int qpg_laser_sync(qpg_laser l)
{
    ARG_TRY(l, "null arg");
    { int rc = laser_join(l); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(l->ctx->stream));
    return 0;
}
extern "C" int qpg_laser_upload(qpg_laser l, const double *ar, const double *ai)
{
    ARG_TRY(l && ar && ai, "null arg");
    { int rc = laser_join(l); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(l->ar, ar, sizeof(double) * l->nvol, cudaMemcpyHostToDevice, l->ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(l->ai, ai, sizeof(double) * l->nvol, cudaMemcpyHostToDevice, l->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(l->ctx->stream));
    return 0;
}
extern "C" int qpg_laser_download(qpg_laser l, double *ar, double *ai)
{
    ARG_TRY(l && ar && ai, "null arg");
    { int rc = laser_join(l); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(ar, l->ar, sizeof(double) * l->nvol, cudaMemcpyDeviceToHost, l->ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ai, l->ai, sizeof(double) * l->nvol, cudaMemcpyDeviceToHost, l->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(l->ctx->stream));
    return 0;
}
// which: 0 a_r, 1 a_i (dim 1), 2 grad a_r, 3 grad a_i (dim 3) -- the slice images of qpg_laser_slice; 4 chi (dim 1, with volume)
extern "C" qpg_field qpg_laser_field(qpg_laser l, int which)
{
    if (!l) return nullptr;
    switch (which) { case 0: return l->f_ar; case 1: return l->f_ai; case 2: return l->f_gr; case 3: return l->f_gi; case 4: return l->chi; }
    qpg_set_error("qpg_laser_field: which must be 0..4");
    return nullptr;
}
extern "C" int qpg_laser_slice(qpg_laser l, int j)
{
    ARG_TRY(l && (j == -1 || (j >= 1 && j <= l->nz)), "slice out of range");
    { int rc = laser_join(l); if (rc) return rc; }
    qpg_ctx c = l->ctx;
    const int n = (c->nr + 2) * c->P;
    k_laser_slice<<<(n + 255) / 256, 256, 0, c->stream>>>(l->ar, l->ai, c->nr, l->nz, c->M, j, j == -1 ? c->flags : nullptr, c->dr, l->dz, l->f_ar->f1, l->f_ai->f1,
                                                          l->f_gr->f1, l->f_gi->f1);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_laser_deposit_chi(qpg_laser l, qpg_part2d p, int j, double ax_corr)
{
    ARG_TRY(l && p && p->ctx == l->ctx && j >= -1 && j <= l->nz, "bad arg");
    { int rc = laser_join(l); if (rc) return rc; }
    qpg_ctx c = l->ctx;
    if (p->npp_hi > 0) {
        const int grid = (int)((p->npp_hi + 255) / 256);
        DISPATCH_M(c->M, l_deposit_chi, grid, c->stream, view_of(p), l->chi_acc, 1.0 / c->dr, p->qbm);
        count_launch(c);
    }
    const int n = (c->nr + 2) * c->P;
    k_chi_fix<<<(n + 255) / 256, 256, 0, c->stream>>>(l->chi_acc, l->chi->f1, l->chi->f2, l->nz, j, j == -1 ? c->flags : nullptr, c->nr, c->P, ax_corr);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
// The envelope on a xi-pipeline (sim_lasers_class.f03:197-222, field_laser_class.f03:163-169): a stage's slab has the two lower guard slices
// 0 and -1 = the last two slices of the upstream stage.  The explicit half of an advance (set_rhs) reads them OLD; then the upstream stage's
// NEW last two slices arrive (pipe_recv 'forward' 'guard': the upstream stage advanced its slab of this 3D step before) and the implicit half
// marches from them.  Wire record: [a_r | a_i][plane][g = 0: slice nz -> guard 0, g = 1: slice nz-1 -> guard -1][node 0..nr+1].
// Message n of a link belongs to the n-th advance of both of its ends; flag words count messages (ready: producer -> consumer, written after
// the record; ack: consumer -> producer, written after the record was copied out, so that one wire buffer per link is enough).
__global__ void k_laser_guard_unpack(double *__restrict__ ar, double *__restrict__ ai, const double *__restrict__ src, int nr, int nz, int P)
{
    const int n = 4 * P * (nr + 2);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int i = t % (nr + 2), g = (t / (nr + 2)) & 1, pl = (t / (2 * (nr + 2))) % P, im = t / (2 * P * (nr + 2));
        (im ? ai : ar)[LVI(pl, i, -g)] = src[t];
    }
}
__global__ void k_laser_guard_pack(const double *__restrict__ ar, const double *__restrict__ ai, double *__restrict__ dst, int nr, int nz, int P)
{
    const int n = 4 * P * (nr + 2);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int i = t % (nr + 2), g = (t / (nr + 2)) & 1, pl = (t / (2 * (nr + 2))) % P, im = t / (2 * P * (nr + 2));
        dst[t] = (im ? ai : ar)[LVI(pl, i, nz - g)];
    }
}
extern "C" long qpg_laser_guard_size(qpg_laser l) { return l ? 4L * l->ctx->P * (l->ctx->nr + 2) : -1; }
extern "C" int qpg_laser_set_handoff(qpg_laser l, const double *guard_in, unsigned *in_ready, unsigned *in_ack, double *guard_out, unsigned *out_ready, unsigned *out_ack)
{
    ARG_TRY(l, "null arg");
    ARG_TRY((guard_in == nullptr) == (in_ready == nullptr) && (guard_in == nullptr) == (in_ack == nullptr), "upstream link: buffer, ready and ack words go together");
    ARG_TRY((guard_out == nullptr) == (out_ready == nullptr) && (guard_out == nullptr) == (out_ack == nullptr), "downstream link: buffer, ready and ack words go together");
    ARG_TRY(l->nz >= 2 || !guard_out, "a stage that hands its last two slices on needs a slab of at least two");
    { int rc = laser_join(l); if (rc) return rc; }
    l->g_in = guard_in; l->g_in_ready = in_ready; l->g_in_ack = in_ack;
    l->g_out = guard_out; l->g_out_ready = out_ready; l->g_out_ack = out_ack;
    l->g_seq = 0;
    return 0;
}
static int laser_launch_advance(qpg_laser l, cudaStream_t st, unsigned *progress, unsigned base)
{
    qpg_ctx c = l->ctx;
    const int nr = c->nr, nz = l->nz;
    const long n = (long)nr * nz;
    const unsigned seq = (l->g_in || l->g_out) ? ++l->g_seq : 0;
    const int ng = 4 * c->P * (nr + 2);
    k_laser_set_rhs<<<(int)((n + 255) / 256), 256, 0, st>>>(l->ar, l->ai, l->chi->f2, l->sr, l->si, nr, nz, c->M, l->k0, l->ds, c->dr, l->dz);
    if (l->g_in) {   // the upstream stage's NEW last two slices replace the guards between the explicit and the implicit half
        int rc = qpg_stream_wait((void *)st, l->g_in_ready, seq);
        if (rc) return rc;
        k_laser_guard_unpack<<<(ng + 255) / 256, 256, 0, st>>>(l->ar, l->ai, l->g_in, nr, nz, c->P);
        count_launch(c);
        if ((rc = qpg_stream_signal((void *)st, l->g_in_ack, seq))) return rc;
    }
    const size_t smem = sizeof(double) * 4 * l->nthreads;
    if (l->pcrc && !getenv("QPG_LASER_GENERAL_SOLVE")) {
        const size_t with_coef = smem + sizeof(double) * l->pcrc_n;
        const int in_smem = with_coef <= 220 * 1024;
        const size_t sm = in_smem ? with_coef : smem;
        switch (c->M) {
        case 0: l_laser_solve_c<0>(l->nthreads, sm, st, l->ar, l->ai, l->sr, l->si, l->chi->f2, l->pcrc, nr, nz, l->iter, l->nsteps, l->ds, c->dr, l->dz, in_smem, progress, base); break;
        case 1: l_laser_solve_c<1>(l->nthreads, sm, st, l->ar, l->ai, l->sr, l->si, l->chi->f2, l->pcrc, nr, nz, l->iter, l->nsteps, l->ds, c->dr, l->dz, in_smem, progress, base); break;
        default: l_laser_solve_c<2>(l->nthreads, sm, st, l->ar, l->ai, l->sr, l->si, l->chi->f2, l->pcrc, nr, nz, l->iter, l->nsteps, l->ds, c->dr, l->dz, in_smem, progress, base); break;
        }
    } else {
        if (progress) { qpg_set_error("the overlapped advance needs the complex-factor solve"); return QPG_ERR_UNSUPPORTED; }
        k_laser_solve<<<1, l->nthreads, smem, st>>>(l->ar, l->ai, l->sr, l->si, l->chi->f2, l->pcr, nr, nz, c->M, l->iter, l->nsteps, l->ds, c->dr, l->dz);
    }
    count_launch(c, 2);
    CUDA_TRY(cudaGetLastError());
    if (l->g_out) {   // pipe_send of the own NEW last two slices, into the downstream stage's wire buffer once it has consumed the previous record
        int rc = seq > 1 ? qpg_stream_wait((void *)st, l->g_out_ack, seq - 1) : 0;
        if (rc) return rc;
        k_laser_guard_pack<<<(ng + 255) / 256, 256, 0, st>>>(l->ar, l->ai, l->g_out, nr, nz, c->P);
        count_launch(c);
        CUDA_TRY(cudaGetLastError());
        if ((rc = qpg_stream_signal((void *)st, l->g_out_ready, seq))) return rc;
    }
    return 0;
}
extern "C" int qpg_laser_advance(qpg_laser l)
{
    ARG_TRY(l, "null arg");
    int rc = laser_join(l);
    if (rc) return rc;
    return laser_launch_advance(l, l->ctx->stream, nullptr, 0);
}
// The advance on a SIDE stream, overlapped with what follows on the context's stream: a sweep kernel of the next 3D step (sweep.cu) may run
// beside the solve and follows its progress word slice by slice (slice j of the new envelope is needed by slice j of the next sweep; the
// solve takes ~8 us per slice, the sweep ~27).  Everything else that touches the envelope joins first (laser_join).  Returns the progress
// word and the value it will hold once slice 0 ... i.e. the caller waits for  *progress - base >= j  (wrap-safe).
int qpg_laser_advance_overlapped(qpg_laser l, unsigned **progress, unsigned *base)
{
    ARG_TRY(l && progress && base, "null arg");
    if (!l->pcrc || getenv("QPG_LASER_GENERAL_SOLVE")) return QPG_ERR_UNSUPPORTED;
    int rc = laser_join(l);      // at most one overlapped advance in flight
    if (rc) return rc;
    qpg_ctx c = l->ctx;
    if (!l->progress) {
        CUDA_TRY(cudaStreamCreateWithFlags(&l->side, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&l->ev_main, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&l->ev_done, cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&l->progress, 128));
        CUDA_TRY(cudaMemsetAsync(l->progress, 0, 128, c->stream));
    }
    CUDA_TRY(cudaEventRecord(l->ev_main, c->stream));            // behind the sweep that deposited chi
    CUDA_TRY(cudaStreamWaitEvent(l->side, l->ev_main, 0));
    *base = l->adv_count * (unsigned)l->nz;
    l->adv_count++;
    rc = laser_launch_advance(l, l->side, l->progress, *base);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(l->ev_done, l->side));
    l->pending = true;
    *progress = l->progress;
    return 0;
}

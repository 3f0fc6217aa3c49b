/*
 * qpad_oracle_laser.c -- CPU restatement of the laser-envelope (ponderomotive guiding centre) field path of QPAD:
 * SURVEY.md §8(f) rank 1.  TEST INFRASTRUCTURE ONLY (see qpad_oracle.h).  PARITY UNPINNED like the rest of the oracle; pinned
 * by the analytic vacuum diffraction of a Gaussian pulse (tests/test_oracle_laser.py).
 *
 * Restated routines (source/ relative to /root/reference):
 *   laser/field_laser_class.f03:269-391  init_solver        -> orc_laser_build_matrix
 *   laser/field_laser_class.f03:393-635  set_rhs_field_laser -> orc_laser_set_rhs
 *   laser/field_laser_class.f03:637-750  set_grad_field_laser -> orc_laser_set_grad   (with the reference's index quirk, below)
 *   laser/field_laser_class.f03:752-927  solve_field_laser   -> orc_laser_solve
 *   laser/profile_laser_class.f03:318-378 launch + profile_laser_lib.f03:56-96 (gaussian), :472-502 (sin2) -> orc_laser_launch_gaussian
 *   species/part2d_class.f03:361-476     deposit_chi_part2d  -> orc_deposit_chi ;  :2581 get_deposit_ax_corr
 * The pentadiagonal systems the reference hands to its parallel cyclic reduction (pcr-fortran/fpcr_penta_class.f03; rows
 * a,b,c,d,e = the five diagonals, sub-sub .. super-super, `set_values_matrix_byrow` :570) are solved here by banded
 * Gaussian elimination -- a direct solve of the same matrix (it is diagonally dominant: |c| > |a|+|b|+|d|+|e|).  The
 * epsilon(1.0) entries the reference writes into the corner rows couple to rows outside the matrix and do not exist here.
 *
 * Laser volume layout (one real array per quantity a_r, a_i, s_r, s_i):
 *   v[plane][slice][node],  plane = re0, re1, im1, ... ; slice index = j + 1 for xi slice j = -1 .. nz+1 (two lower guard
 *   slices as gc_num(1,2) = 2 needs for the 3-point backward xi difference, one upper); node = 0 .. nr+1 (radial guards).
 * chi volume: the ordinary field f2 layout of qpad_oracle.h, [plane][slice 1..nz+1][node][dim=1].
 */
#include "qpad_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int lpl_re(int m) { return m == 0 ? 0 : 2 * m - 1; }
static inline int lpl_im(int m) { return 2 * m; }
#define LV(v, nr, nz, pl, i, j) ((v)[(((size_t)(pl)) * ((nz) + 3) + (size_t)((j) + 1)) * ((nr) + 2) + (i)])
#define CHI(f, nr, nz, pl, i, j) ((f)[(((size_t)(pl)) * ((nz) + 1) + (size_t)((j)-1)) * ((nr) + 2) + (i)])
#define G1(f, nr, pl, c, j) ((f)[(((size_t)(pl)) * ((nr) + 2) + (j)) * 3 + ((c)-1)])

long orc_laser_volume_size(int nr, int nz, int max_mode) { return (long)(2 * max_mode + 1) * (nz + 3) * (nr + 2); }

/* species/part2d_class.f03:2581-2586 */
double orc_deposit_ax_corr(int ppc_r) { return (12.0 * ppc_r * ppc_r) / (1.0 + 2.0 * ppc_r * ppc_r); }

/* species/part2d_class.f03:361-476 (noff == 0 branch :432-452); chi is a dim-1 multi-plane f1 [plane][0:nr+1] */
void orc_deposit_chi(const double *x, const double *q, const double *psi, long npp, double dr, int nr, int max_mode, double qbm,
                     double ax_corr, double *chi)
{
#define C1(pl, j) chi[((size_t)(pl)) * (nr + 2) + (j)]
    const double idr = 1.0 / dr;
    for (long pp = 0; pp < npp; pp++) {
        double pos = sqrt(x[2 * pp] * x[2 * pp] + x[2 * pp + 1] * x[2 * pp + 1]) * idr;
        const double p0r = x[2 * pp] / pos * idr, p0i = -x[2 * pp + 1] / pos * idr;
        int nn = (int)floor(pos);
        pos = pos - (double)nn;
        nn = nn + 1;
        const double wt[2] = {1.0 - pos, pos};
        double phr = -1.0 * qbm * q[pp] / (1.0 - qbm * psi[pp]), phi = 0.0;   /* :409 */
        for (int j = 0; j < 2; j++) C1(0, nn + j) += wt[j] * phr;
        for (int m = 1; m <= max_mode; m++) {
            const double t = phr * p0r - phi * p0i;
            phi = phr * p0i + phi * p0r;
            phr = t;
            for (int j = 0; j < 2; j++) { C1(lpl_re(m), nn + j) += wt[j] * phr; C1(lpl_im(m), nn + j) += wt[j] * phi; }
        }
    }
    C1(0, 0) = 0.0;
    C1(0, 1) = C1(0, 1) * ax_corr;                                           /* :436 */
    for (int j = 2; j <= nr + 1; j++) C1(0, j) = C1(0, j) * (1.0 / (double)(j - 1));
    for (int m = 1; m <= max_mode; m++) {
        const int pr = lpl_re(m), pi = lpl_im(m);
        C1(pr, 0) = 0.0; C1(pi, 0) = 0.0; C1(pr, 1) = 0.0; C1(pi, 1) = 0.0;
        for (int j = 2; j <= nr + 1; j++) { const double ir = 1.0 / (double)(j - 1); C1(pr, j) = C1(pr, j) * ir; C1(pi, j) = C1(pi, j) * ir; }
    }
#undef C1
}

/* laser/profile_laser_lib.f03:472-502 */
static double prof_lon_sin2(double z, double t_rise, double t_flat, double t_fall)
{
    const double pih = 1.570796326794897;
    const double flat_start = -0.5 * t_flat, flat_end = 0.5 * t_flat;
    double env;
    if (z < flat_start - t_rise) env = 0.0;
    else if (z < flat_start) { env = cos((z - flat_start) / t_rise * pih); env = env * env; }
    else if (z < flat_end) env = 1.0;
    else if (z < flat_end + t_fall) { env = cos((z - flat_end) / t_fall * pih); env = env * env; }
    else env = 0.0;
    return env;
}
/* laser/profile_laser_lib.f03:56-96 (mode 0; higher modes of a Gaussian are zero) */
void orc_laser_gaussian_point(double r, double z, double k, double k0, double w0, double f_dist, double *ar, double *ai)
{
    const double z_shift = -1.0 * (z + f_dist), r2 = r * r, z2 = z_shift * z_shift, zr = 0.5 * k * w0 * w0, zr2 = zr * zr;
    const double curv = z_shift / (z2 + zr2), w = w0 * sqrt(1.0 + z2 / zr2), gouy = atan2(z_shift, zr);
    const double phase = 0.5 * k * r2 * curv - gouy - (k - k0) * z, amp = w0 / w * exp(-r2 / (w * w));
    *ar = amp * cos(phase);
    *ai = -amp * sin(phase);
}
/* laser/profile_laser_class.f03:318-378 launch (gaussian x sin2, no chirp), one stage owning the whole box; guards stay 0 */
void orc_laser_launch_gaussian(double k0, double a0, double w0, double f_dist, double lon_center, double t_rise, double t_flat, double t_fall,
                               double z0, double dz, double dr, int nr, int nz, int max_mode, double *ar, double *ai)
{
    memset(ar, 0, sizeof(double) * (size_t)orc_laser_volume_size(nr, nz, max_mode));
    memset(ai, 0, sizeof(double) * (size_t)orc_laser_volume_size(nr, nz, max_mode));
    for (int j = 1; j <= nz; j++) {
        const double z = (double)(j - 1) * dz + z0 - lon_center;   /* "z" is xi = t - z */
        const double env = prof_lon_sin2(z, t_rise, t_flat, t_fall) * a0;
        for (int i = 1; i <= nr; i++) {
            double arr, air;
            orc_laser_gaussian_point((double)(i - 1) * dr, z, k0, k0, w0, f_dist, &arr, &air);
            LV(ar, nr, nz, 0, i, j) = env * arr;
            LV(ai, nr, nz, 0, i, j) = env * air;
        }
    }
}

/* laser/field_laser_class.f03:269-391: rows of the 2nr x 2nr pentadiagonal operator of mode m, A[row][0..4] = a,b,c,d,e;
 * unknowns interleaved x[2(i-1)] = a_r(node i), x[2(i-1)+1] = a_i(node i), node i <-> radius (i-1) dr */
void orc_laser_build_matrix(int m, int nr, double k0, double ds, double dr, double dz, double *A)
{
    const double ds_qtr = 0.25 * ds, dr2_idz_1hf = 1.5 * dr * dr / dz, m2 = (double)m * m;
    for (int i = 1; i <= 2 * nr; i += 2) {
        const int j = (i + 1) / 2 - 1;
        double *r0 = A + (size_t)(i - 1) * 5, *r1 = r0 + 5;
        if (j == 0) {                                   /* axial rows, :318-351 */
            if (m == 0) {
                r0[0] = 0.0; r0[1] = 0.0; r0[2] = ds + dr2_idz_1hf; r0[3] = -k0 * dr * dr; r0[4] = -ds;
                r1[0] = 0.0; r1[1] = k0 * dr * dr; r1[2] = ds + dr2_idz_1hf; r1[3] = 0.0; r1[4] = -ds;
            } else {                                    /* decoupled (the reference puts epsilon(1.0) off the diagonal; the axis
                                                           values are overwritten with 0 after the solve, :913-918) */
                r0[0] = r0[1] = r0[3] = r0[4] = 0.0; r0[2] = 1.0;
                r1[0] = r1[1] = r1[3] = r1[4] = 0.0; r1[2] = 1.0;
            }
            continue;
        }
        const double j2 = (double)j * j;
        r0[0] = -ds_qtr * (1.0 - 0.5 / j); r0[1] = 0.0; r0[2] = ds_qtr * (2.0 + m2 / j2) + dr2_idz_1hf; r0[3] = -k0 * dr * dr; r0[4] = -ds_qtr * (1.0 + 0.5 / j);
        r1[0] = -ds_qtr * (1.0 - 0.5 / j); r1[1] = k0 * dr * dr; r1[2] = ds_qtr * (2.0 + m2 / j2) + dr2_idz_1hf; r1[3] = 0.0; r1[4] = -ds_qtr * (1.0 + 0.5 / j);
        if (j == nr - 1) { r0[4] = 0.0; r1[4] = 0.0; }  /* outer rows, :356-373 */
    }
}

/* direct solve of the pentadiagonal system (rows a,b,c,d,e; entries reaching outside the matrix are ignored); rhs -> x */
void orc_penta_solve(const double *A, double *x, int n)
{
    /* band storage with fill-in confined to the band (no pivoting: the operator is diagonally dominant) */
    double *w = (double *)malloc(sizeof(double) * 5 * (size_t)n);
    memcpy(w, A, sizeof(double) * 5 * (size_t)n);
#define W(r, k) w[(size_t)(r)*5 + (k)]   /* k = 0..4 <-> column r-2..r+2 */
    for (int r = 0; r < n; r++) {
        if (r < 1) { W(r, 1) = 0.0; }
        if (r < 2) { W(r, 0) = 0.0; }
        if (r > n - 2) { W(r, 3) = 0.0; }
        if (r > n - 3) { W(r, 4) = 0.0; }
    }
    for (int r = 0; r < n; r++) {
        const double piv = W(r, 2);
        for (int t = 1; t <= 2 && r + t < n; t++) {           /* eliminate column r from rows r+1, r+2 */
            const double f = W(r + t, 2 - t) / piv;
            if (f == 0.0) continue;
            W(r + t, 2 - t) = 0.0;
            for (int c = 1; c <= 2; c++) if (2 - t + c <= 4) W(r + t, 2 - t + c) -= f * W(r, 2 + c);
            x[r + t] -= f * x[r];
        }
    }
    for (int r = n - 1; r >= 0; r--) {
        double s = x[r];
        if (r + 1 < n) s -= W(r, 3) * x[r + 1];
        if (r + 2 < n) s -= W(r, 4) * x[r + 2];
        x[r] = s / W(r, 2);
    }
#undef W
    free(w);
}

/* the plasma-susceptibility couplings shared by set_rhs (:563-631) and solve (:787-852): adds ds/4 dr^2 (chi * a)_m of slice j
 * into tr_re/ti_re (and _im for m > 0), each [max_mode+1][nr+2] */
static void chi_coupling(const double *ar, const double *ai, const double *chi, int nr, int nz, int max_mode, int j, double ds_qtr_dr2,
                         double *tr_re, double *tr_im, double *ti_re, double *ti_im)
{
#define T(t, m, i) t[(size_t)(m) * (nr + 2) + (i)]
    for (int m = 0; m <= max_mode; m++) {
        for (int k = m - max_mode; k <= max_mode; k++) {
            const int ak = abs(k), amk = abs(m - k);
            double sign_pm = (k >= 0 && k <= m) ? 1.0 : -1.0;
            for (int i = 1; i <= nr; i++) {
                T(tr_re, m, i) += ds_qtr_dr2 * CHI(chi, nr, nz, lpl_re(ak), i, j) * LV(ar, nr, nz, lpl_re(amk), i, j);
                T(ti_re, m, i) += ds_qtr_dr2 * CHI(chi, nr, nz, lpl_re(ak), i, j) * LV(ai, nr, nz, lpl_re(amk), i, j);
            }
            if (k == 0 || k == m) continue;
            for (int i = 1; i <= nr; i++) {
                T(tr_re, m, i) -= ds_qtr_dr2 * sign_pm * CHI(chi, nr, nz, lpl_im(ak), i, j) * LV(ar, nr, nz, lpl_im(amk), i, j);
                T(ti_re, m, i) -= ds_qtr_dr2 * sign_pm * CHI(chi, nr, nz, lpl_im(ak), i, j) * LV(ai, nr, nz, lpl_im(amk), i, j);
            }
        }
        if (m == 0) continue;
        for (int k = m - max_mode; k <= max_mode; k++) {
            const int ak = abs(k), amk = abs(m - k);
            if (k != 0) {
                const double sign_pm = k < 0 ? -1.0 : 1.0;
                for (int i = 1; i <= nr; i++) {
                    T(tr_im, m, i) += ds_qtr_dr2 * sign_pm * CHI(chi, nr, nz, lpl_im(ak), i, j) * LV(ar, nr, nz, lpl_re(amk), i, j);
                    T(ti_im, m, i) += ds_qtr_dr2 * sign_pm * CHI(chi, nr, nz, lpl_im(ak), i, j) * LV(ai, nr, nz, lpl_re(amk), i, j);
                }
            }
            if (k != m) {
                const double sign_pm = k > m ? -1.0 : 1.0;
                for (int i = 1; i <= nr; i++) {
                    T(tr_im, m, i) += ds_qtr_dr2 * sign_pm * CHI(chi, nr, nz, lpl_re(ak), i, j) * LV(ar, nr, nz, lpl_im(amk), i, j);
                    T(ti_im, m, i) += ds_qtr_dr2 * sign_pm * CHI(chi, nr, nz, lpl_re(ak), i, j) * LV(ai, nr, nz, lpl_im(amk), i, j);
                }
            }
        }
    }
#undef T
}

/* laser/field_laser_class.f03:393-635: explicit half of the Crank-Nicolson step from the OLD envelope: sr, si volumes (same
 * layout as ar, ai; only slices 1..nz, nodes 1..nr are written) */
void orc_laser_set_rhs(const double *ar, const double *ai, const double *chi, int nr, int nz, int max_mode, double k0, double ds, double dr,
                       double dz, double *sr, double *si)
{
    const double dr2_idzh = 0.5 * dr * dr / dz, kappa = k0 * dr * dr, ds_qtr = 0.25 * ds, ds_qtr_dr2 = ds_qtr * dr * dr;
    for (int m = 0; m <= max_mode; m++) {
        const double m2 = (double)m * m;
        const int npl = m == 0 ? 1 : 2;
        for (int h = 0; h < npl; h++) {
            const int pl = h == 0 ? lpl_re(m) : lpl_im(m);
            for (int j = 1; j <= nz; j++)
                for (int i = 1; i <= nr; i++) {
                    double beta_m, beta_p, alpha;
                    if (i == 1) {
                        if (m == 0) { beta_m = 0.0; beta_p = ds; alpha = -ds; }                   /* :439-444 */
                        else { LV(sr, nr, nz, pl, i, j) = 0.0; LV(si, nr, nz, pl, i, j) = 0.0; continue; }   /* :507-512 */
                    } else {
                        const double ik = 1.0 / (double)(i - 1);
                        beta_m = ds_qtr * (1.0 - 0.5 * ik);
                        beta_p = i == nr ? 0.0 : ds_qtr * (1.0 + 0.5 * ik);                       /* :459-465, :533-539 */
                        alpha = -ds_qtr * (2.0 + m2 * ik * ik);
                    }
#define AR(ii, jj) LV(ar, nr, nz, pl, ii, jj)
#define AI(ii, jj) LV(ai, nr, nz, pl, ii, jj)
                    LV(sr, nr, nz, pl, i, j) = dr2_idzh * (3.0 * AR(i, j) - 4.0 * AR(i, j - 1) + AR(i, j - 2)) - kappa * AI(i, j)
                                               + beta_m * AR(i - 1, j) + alpha * AR(i, j) + beta_p * AR(i + 1, j);
                    LV(si, nr, nz, pl, i, j) = dr2_idzh * (3.0 * AI(i, j) - 4.0 * AI(i, j - 1) + AI(i, j - 2)) + kappa * AR(i, j)
                                               + beta_m * AI(i - 1, j) + alpha * AI(i, j) + beta_p * AI(i + 1, j);
#undef AR
#undef AI
                }
        }
    }
    /* :563-631 contribution of the plasma susceptibility (old envelope) */
    const size_t nt = (size_t)(max_mode + 1) * (nr + 2);
    double *t = (double *)calloc(4 * nt, sizeof(double));
    for (int j = 1; j <= nz; j++) {
        memset(t, 0, sizeof(double) * 4 * nt);
        chi_coupling(ar, ai, chi, nr, nz, max_mode, j, ds_qtr_dr2, t, t + nt, t + 2 * nt, t + 3 * nt);
        for (int m = 0; m <= max_mode; m++)
            for (int i = 1; i <= nr; i++) {
                LV(sr, nr, nz, lpl_re(m), i, j) += t[(size_t)m * (nr + 2) + i];
                LV(si, nr, nz, lpl_re(m), i, j) += t[2 * nt + (size_t)m * (nr + 2) + i];
                if (m == 0) continue;
                LV(sr, nr, nz, lpl_im(m), i, j) += t[nt + (size_t)m * (nr + 2) + i];
                LV(si, nr, nz, lpl_im(m), i, j) += t[3 * nt + (size_t)m * (nr + 2) + i];
            }
    }
    free(t);
}

/* laser/field_laser_class.f03:752-927: implicit half, slice after slice in xi (slice j needs the NEW slices j-1, j-2: lower
 * guard slices 0, -1 are the upstream stage's or zero), `iter` fixed-point passes over the chi coupling per slice */
void orc_laser_solve(double *ar, double *ai, const double *sr, const double *si, const double *chi, int nr, int nz, int max_mode,
                     double k0, double ds, double dr, double dz, int iter)
{
    const double dr2_idzh = 0.5 * dr * dr / dz, ds_qtr_dr2 = 0.25 * ds * dr * dr;
    const size_t nt = (size_t)(max_mode + 1) * (nr + 2);
    double *t = (double *)calloc(4 * nt, sizeof(double));
    double *A = (double *)malloc(sizeof(double) * 5 * 2 * (size_t)nr * (max_mode + 1));
    double *x = (double *)malloc(sizeof(double) * 2 * (size_t)nr);
    for (int m = 0; m <= max_mode; m++) orc_laser_build_matrix(m, nr, k0, ds, dr, dz, A + (size_t)m * 10 * nr);
    for (int j = 1; j <= nz; j++)
        for (int l = 1; l <= iter; l++) {
            memset(t, 0, sizeof(double) * 4 * nt);
            chi_coupling(ar, ai, chi, nr, nz, max_mode, j, ds_qtr_dr2, t, t + nt, t + 2 * nt, t + 3 * nt);
            for (int m = 0; m <= max_mode; m++) {
                const int npl = m == 0 ? 1 : 2;
                for (int h = 0; h < npl; h++) {
                    const int pl = h == 0 ? lpl_re(m) : lpl_im(m);
                    const double *tr = t + (h == 0 ? 0 : nt) + (size_t)m * (nr + 2), *ti = t + (h == 0 ? 2 * nt : 3 * nt) + (size_t)m * (nr + 2);
                    for (int i = 1; i <= nr; i++) {
                        x[2 * i - 2] = LV(sr, nr, nz, pl, i, j) + tr[i] + dr2_idzh * (4.0 * LV(ar, nr, nz, pl, i, j - 1) - LV(ar, nr, nz, pl, i, j - 2));
                        x[2 * i - 1] = LV(si, nr, nz, pl, i, j) + ti[i] + dr2_idzh * (4.0 * LV(ai, nr, nz, pl, i, j - 1) - LV(ai, nr, nz, pl, i, j - 2));
                    }
                    orc_penta_solve(A + (size_t)m * 10 * nr, x, 2 * nr);
                    for (int i = 1; i <= nr; i++) { LV(ar, nr, nz, pl, i, j) = x[2 * i - 2]; LV(ai, nr, nz, pl, i, j) = x[2 * i - 1]; }
                }
                if (m > 0) {   /* :913-918 on-axis values are zeros for m > 0 (after BOTH planes were solved) */
                    LV(ar, nr, nz, lpl_re(m), 1, j) = 0.0; LV(ai, nr, nz, lpl_re(m), 1, j) = 0.0;
                    LV(ar, nr, nz, lpl_im(m), 1, j) = 0.0; LV(ai, nr, nz, lpl_im(m), 1, j) = 0.0;
                }
            }
        }
    free(t); free(A); free(x);
}

/* laser/field_laser_class.f03:637-750: gradients of slice `slice` into dim-3 multi-plane f1 fields (comp 1 = d/dr, 2 = the
 * azimuthal term (im/r), 3 = d/dxi by the 3-point backward difference).
 * Quirk kept from the reference (:708-730): for m > 0 the on-axis rules are written with the loop variable `i` AFTER the
 * `do i = 2, nrp` loop, i.e. into node nrp+1 (the outer guard), not into node 1; node 1 keeps components 1, 2 of the previous
 * call (0 after init).  copy_gc_f1 (:739-745) is a no-op for a single radial owner. */
void orc_laser_set_grad(const double *ar, const double *ai, int slice, int nr, int nz, int max_mode, double dr, double dz, double *ar_grad,
                        double *ai_grad)
{
    const double idrh = 0.5 / dr, idzh = 0.5 / dz;
    const int s = slice;
    for (int i = 1; i <= nr; i++) {
        G1(ar_grad, nr, 0, 3, i) = idzh * (3.0 * LV(ar, nr, nz, 0, i, s) - 4.0 * LV(ar, nr, nz, 0, i, s - 1) + LV(ar, nr, nz, 0, i, s - 2));
        G1(ai_grad, nr, 0, 3, i) = idzh * (3.0 * LV(ai, nr, nz, 0, i, s) - 4.0 * LV(ai, nr, nz, 0, i, s - 1) + LV(ai, nr, nz, 0, i, s - 2));
    }
    for (int i = 2; i <= nr; i++) {
        G1(ar_grad, nr, 0, 1, i) = idrh * (LV(ar, nr, nz, 0, i + 1, s) - LV(ar, nr, nz, 0, i - 1, s));
        G1(ai_grad, nr, 0, 1, i) = idrh * (LV(ai, nr, nz, 0, i + 1, s) - LV(ai, nr, nz, 0, i - 1, s));
        G1(ar_grad, nr, 0, 2, i) = 0.0;
        G1(ai_grad, nr, 0, 2, i) = 0.0;
    }
    G1(ar_grad, nr, 0, 1, 1) = 0.0; G1(ai_grad, nr, 0, 1, 1) = 0.0; G1(ar_grad, nr, 0, 2, 1) = 0.0; G1(ai_grad, nr, 0, 2, 1) = 0.0;
    for (int m = 1; m <= max_mode; m++) {
        const int pr = lpl_re(m), pi = lpl_im(m);
        for (int i = 1; i <= nr; i++) {
            G1(ar_grad, nr, pr, 3, i) = idzh * (3.0 * LV(ar, nr, nz, pr, i, s) - 4.0 * LV(ar, nr, nz, pr, i, s - 1) + LV(ar, nr, nz, pr, i, s - 2));
            G1(ar_grad, nr, pi, 3, i) = idzh * (3.0 * LV(ar, nr, nz, pi, i, s) - 4.0 * LV(ar, nr, nz, pi, i, s - 1) + LV(ar, nr, nz, pi, i, s - 2));
            G1(ai_grad, nr, pr, 3, i) = idzh * (3.0 * LV(ai, nr, nz, pr, i, s) - 4.0 * LV(ai, nr, nz, pr, i, s - 1) + LV(ai, nr, nz, pr, i, s - 2));
            G1(ai_grad, nr, pi, 3, i) = idzh * (3.0 * LV(ai, nr, nz, pi, i, s) - 4.0 * LV(ai, nr, nz, pi, i, s - 1) + LV(ai, nr, nz, pi, i, s - 2));
        }
        int i;
        for (i = 2; i <= nr; i++) {
            const double ir = 1.0 / ((double)(i - 1) * dr);
            G1(ar_grad, nr, pr, 1, i) = idrh * (LV(ar, nr, nz, pr, i + 1, s) - LV(ar, nr, nz, pr, i - 1, s));
            G1(ar_grad, nr, pi, 1, i) = idrh * (LV(ar, nr, nz, pi, i + 1, s) - LV(ar, nr, nz, pi, i - 1, s));
            G1(ar_grad, nr, pr, 2, i) = -ir * m * LV(ar, nr, nz, pi, i, s);
            G1(ar_grad, nr, pi, 2, i) = ir * m * LV(ar, nr, nz, pr, i, s);
            G1(ai_grad, nr, pr, 1, i) = idrh * (LV(ai, nr, nz, pr, i + 1, s) - LV(ai, nr, nz, pr, i - 1, s));
            G1(ai_grad, nr, pi, 1, i) = idrh * (LV(ai, nr, nz, pi, i + 1, s) - LV(ai, nr, nz, pi, i - 1, s));
            G1(ai_grad, nr, pr, 2, i) = -ir * m * LV(ai, nr, nz, pi, i, s);
            G1(ai_grad, nr, pi, 2, i) = ir * m * LV(ai, nr, nz, pr, i, s);
        }
        /* here i == nr + 1 : the reference's axis block lands on the outer guard node */
        if (m % 2 == 1) {
            G1(ar_grad, nr, pr, 1, i) = 2.0 * idrh * LV(ar, nr, nz, pr, 2, s);
            G1(ar_grad, nr, pi, 1, i) = 2.0 * idrh * LV(ar, nr, nz, pi, 2, s);
            G1(ar_grad, nr, pr, 2, i) = -m * G1(ar_grad, nr, pi, 1, i);
            G1(ar_grad, nr, pi, 2, i) = m * G1(ar_grad, nr, pr, 1, i);
            G1(ai_grad, nr, pr, 1, i) = 2.0 * idrh * LV(ai, nr, nz, pr, 2, s);
            G1(ai_grad, nr, pi, 1, i) = 2.0 * idrh * LV(ai, nr, nz, pi, 2, s);
            G1(ai_grad, nr, pr, 2, i) = -m * G1(ai_grad, nr, pi, 1, i);
            G1(ai_grad, nr, pi, 2, i) = m * G1(ai_grad, nr, pr, 1, i);
        } else {
            for (int c = 1; c <= 2; c++) { G1(ar_grad, nr, pr, c, i) = 0.0; G1(ar_grad, nr, pi, c, i) = 0.0; G1(ai_grad, nr, pr, c, i) = 0.0; G1(ai_grad, nr, pi, c, i) = 0.0; }
        }
    }
}

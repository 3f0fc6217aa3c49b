/*
 * qpad_oracle_neutral.c -- CPU restatement of QPAD's field-ionisation (ADK) neutral species: SURVEY.md §8(f) rank 2.
 * TEST INFRASTRUCTURE ONLY (see qpad_oracle.h).  PARITY UNPINNED like the rest of the oracle; pinned by the closed form of the
 * rate equations in a constant field (tests/test_oracle_neutral.py).
 *
 * Restated routines (source/species/neutral_class.f03 of /root/reference):
 *   :600-753  ionize_neutral   -> orc_neutral_ionize      (per (radial cell, theta sector) ionisation levels, ADK rates)
 *   :755-837  add_particles    -> orc_neutral_add_particles (deterministic electron creation, position buffer for the ions)
 *   :904-930  ion_deposit      -> orc_neutral_ion_deposit
 *   :839-878  renew            -> orc_neutral_reset
 * ADK rate of level i in a field E [GV/m] (:654-656):  w_i = r1_i E^(-r3_i) exp(-r2_i / E)  [1/s], the three parameters per
 * level being the usual tunnelling-rate constants  r2 = 6.83 xi^1.5,  r3 = 2 n* - 1,
 * r1 = 1.52e15 4^n* xi / (n* Gamma(2 n*)) (20.5 xi^1.5)^(2 n* - 1),  n* = 3.69 Z / sqrt(xi)  (xi = ionisation energy in eV, NIST
 * ASD).  The numeric values below are the reference's constants for H, He and Li (:39-52, :177-180) -- the elements the
 * shipped decks use (input_file/ionization: Li, ion_max 3); other elements return an error.
 *
 * Quirks of the reference that are kept (they define its results):
 *   - :636-637  the imaginary plane of the m > 0 field modes is read from the REAL plane (`e_im => e%rf_re(m)`);
 *   - :661, :688  the "2nd order Runge-Kutta" update n w dt (1 + w dt / 2) is the reference's (its own comment calls it wrong);
 *   - :733-734  the total ion level is rounded to a multiple of ion_max / ppc so that whole macro-electrons are released.
 * Level array layout: lev[(i * n_theta + k) * nr + j], i = 0..multi_max+1 (0-based level index: 0..multi_max-1 = charge states
 * 1..multi_max, multi_max = neutral residue, multi_max+1 = total discrete ion level), k = theta sector, j = radial cell.
 */
#include "qpad_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double H_param[1][3] = {{8.522542995398661e19, 342.53947239007687, 1.0005337056631487}};
static const double He_param[2][3] = {{7.2207661763501e18, 832.809878216992, 0.48776427204592254}, {2.7226733893691e21, 2742.1316798375965, 1.0000920088118899}};
static const double Li_param[3][3] = {{3.460272990838495e21, 85.51998980232813, 2.1770706138013733},
                                      {3.6365138642921554e20, 4493.713340713575, 0.6964625952167312},
                                      {2.0659396971422902e22, 9256.32561931876, 0.9999745128918196}};

/* element = atomic number (param.f03:135-148); returns the number of levels available (0: unsupported element) and copies
 * min(max_e, available) rows of (r1, r2, r3) */
int orc_adk_params(int element, int max_e, double *out)
{
    const double(*tab)[3] = NULL;
    int n = 0;
    switch (element) {
    case 1: tab = H_param; n = 1; break;
    case 2: tab = He_param; n = 2; break;
    case 3: tab = Li_param; n = 3; break;
    default: return 0;
    }
    if (max_e < n) n = max_e;
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) out[3 * i + c] = tab[i][c];
    return n;
}
/* sim_plasma_class.f03:84  omega_p = sqrt(n0) * 5.641460231180626e4  (n0 in cm^-3) */
double orc_plasma_frequency(double n0) { return sqrt(n0) * 5.641460231180626e4; }

#define LEV(i, k, j) lev[((size_t)(i) * n_theta + (k)) * nr + (j)]
#define EF(pl, c, j) ef[(((size_t)(pl)) * (nr + 2) + (j)) * 3 + ((c)-1)]

void orc_neutral_reset(double *lev, int nr, int n_theta, int multi_max)
{
    memset(lev, 0, sizeof(double) * (size_t)(multi_max + 2) * n_theta * nr);
    for (int k = 0; k < n_theta; k++) for (int j = 0; j < nr; j++) LEV(multi_max, k, j) = 1.0;   /* :562-563, :864-865 */
}

/* neutral_class.f03:600-753; ef = dim-3 multi-plane f1 of E; dt = dxi (:424) */
void orc_neutral_ionize(double *lev, const double *adk, const double *ef, double wp, double dt, int ppc1, int ppc2, int nr, int n_theta,
                        int max_mode, int multi_max)
{
    const double pi = 3.14159265358979323846;
    const int ppc_tot = ppc1 * ppc2, idx_neut = multi_max, idx_ion = multi_max + 1;
    const double pi2_ntheta = 2.0 * pi / (double)n_theta;
    double *e1 = (double *)malloc(sizeof(double) * 3 * (size_t)nr), *e2 = e1 + nr, *e3 = e2 + nr;
    double w_ion[20];
    for (int k = 0; k < n_theta; k++) {
        const double theta = pi2_ntheta * (double)k;
        const double incr = cos(theta), inci = sin(theta);
        double phr = 1.0, phi = 0.0;
        for (int j = 1; j <= nr; j++) {                                     /* in-cell values, m = 0 (:622-627) */
            e1[j - 1] = 0.5 * (EF(0, 1, j) + EF(0, 1, j + 1));
            e2[j - 1] = 0.5 * (EF(0, 2, j) + EF(0, 2, j + 1));
            e3[j - 1] = 0.5 * (EF(0, 3, j) + EF(0, 3, j + 1));
        }
        for (int m = 1; m <= max_mode; m++) {
            const int pr = 2 * m - 1, pi_ = pr;                             /* :636-637: BOTH pointers are the real plane */
            const double t = phr * incr - phi * inci;
            phi = phr * inci + phi * incr;
            phr = t;
            for (int j = 1; j <= nr; j++) {
                e1[j - 1] = e1[j - 1] + (EF(pr, 1, j) + EF(pr, 1, j + 1)) * phr - (EF(pi_, 1, j) + EF(pi_, 1, j + 1)) * phi;
                e2[j - 1] = e2[j - 1] + (EF(pr, 2, j) + EF(pr, 2, j + 1)) * phr - (EF(pi_, 2, j) + EF(pi_, 2, j + 1)) * phi;
                e3[j - 1] = e3[j - 1] + (EF(pr, 3, j) + EF(pr, 3, j + 1)) * phr - (EF(pi_, 3, j) + EF(pi_, 3, j + 1)) * phi;
            }
        }
        for (int j = 0; j < nr; j++) {
            const double eij = sqrt(e1[j] * e1[j] + e2[j] * e2[j] + e3[j] * e3[j]) * wp * 1.708e-12;   /* GV/m */
            if (!(eij > 1.0e-6)) continue;
            if (!(LEV(idx_ion, k, j) < (double)multi_max)) continue;
            int shoot = 0;
            for (int i = 0; i < multi_max; i++) w_ion[i] = adk[3 * i] * pow(eij, -adk[3 * i + 2]) * exp(-adk[3 * i + 1] / eij) / wp;
            double cons = LEV(idx_neut, k, j) * w_ion[0] * dt * (1.0 + 0.5 * w_ion[0] * dt);          /* :661 */
            if (cons > LEV(idx_neut, k, j)) { shoot = 1; cons = LEV(idx_neut, k, j); }
            LEV(idx_neut, k, j) = LEV(idx_neut, k, j) - cons;
            for (int i = 0; i < multi_max - 1; i++) {                       /* the levels in the middle (:672-700) */
                const double inj = cons;
                double dens_temp = 0.0;
                if (shoot) { dens_temp = inj * 0.5; shoot = 0; }
                cons = (LEV(i, k, j) + dens_temp) * w_ion[i + 1] * dt * (1.0 + 0.5 * w_ion[i + 1] * dt);
                if (cons > LEV(i, k, j) + dens_temp) { shoot = 1; cons = LEV(i, k, j) + dens_temp; }
                LEV(i, k, j) = fmin(LEV(i, k, j) - cons + inj, 1.0);
            }
            LEV(multi_max - 1, k, j) = fmin(LEV(multi_max - 1, k, j) + cons, 1.0);                    /* :703 */
            double tot = 0.0;
            for (int i = 0; i < multi_max; i++) tot = tot + (double)(i + 1) * LEV(i, k, j);
            /* :733-734 discrete total: whole macro-electrons, released at the 'half' step */
            LEV(idx_ion, k, j) = (double)multi_max / (double)ppc_tot * (double)(int)(tot * (double)ppc_tot / (double)multi_max + 0.5);
        }
    }
    free(e1);
}

/* neutral_class.f03:755-837 for the uniform / uniform profile (den_lon = den_perp = 1): appends the new electrons to the particle
 * arrays (room for ppc_tot * nr * n_theta more must exist) and fills the position buffer of the ions (xa (2,n), qa (n)).
 * Returns the number of created electrons; *npp is advanced. */
long orc_neutral_add_particles(const double *lev, const double *ion_old, int nr, int n_theta, int multi_max, int ppc1, int ppc2, double dr,
                               double qm, double density, double den_min, double *x, double *p, double *gamma, double *psi, double *q, long *npp,
                               double *xa, double *qa)
{
    const double pi = 3.14159265358979323846;
    const int ppc_tot = ppc1 * ppc2, idx_ion = multi_max + 1;
    const double dtheta = 2.0 * pi / (double)n_theta;
    const double coef = (double)multi_max * (qm < 0 ? -1.0 : 1.0) / ((double)ppc_tot * (double)n_theta);
    long nadd = 0, pp1 = *npp;
    for (int k = 0; k < n_theta; k++)
        for (int j = 0; j < nr; j++) {
            const int ppc_add = (int)((LEV(idx_ion, k, j) - ion_old[(size_t)k * nr + j]) / (double)multi_max * (double)ppc_tot + 0.5);   /* :793 */
            long a1 = pp1, a2 = nadd;
            for (int i = 1; i <= ppc_add; i++) {
                const double rn = (double)j + ((double)i - 0.5) / (double)ppc_add;                     /* :803 (j is 0-based here) */
                const double theta = (double)k * dtheta;
                const double x1 = rn * dr * cos(theta), x2 = rn * dr * sin(theta);
                if (1.0 * 1.0 * density < den_min) continue;
                x[2 * a1] = x1; x[2 * a1 + 1] = x2;
                q[a1] = rn * 1.0 * 1.0 * density * coef;
                p[3 * a1] = 0.0; p[3 * a1 + 1] = 0.0; p[3 * a1 + 2] = 0.0;
                gamma[a1] = 1.0; psi[a1] = 0.0;
                xa[2 * a2] = x1; xa[2 * a2 + 1] = x2;
                qa[a2] = -q[a1];                                                                       /* note the sign (:826) */
                a1++; a2++;
            }
            /* :830-831 the counters advance by ppc_add even if the density cut skipped particles; with the uniform profile the two agree */
            pp1 += ppc_add; nadd += ppc_add;
        }
    *npp = pp1;
    return nadd;
}

/* neutral_class.f03:904-930: rho_ion_add = deposit(position buffer); rho_ion += rho_ion_add; q_tot += rho_ion.  All dim-1 f1. */
void orc_neutral_ion_deposit(const double *xa, const double *qa, long nadd, double dr, int nr, int max_mode, double *rho_ion, double *q_tot)
{
    const size_t n1 = (size_t)(2 * max_mode + 1) * (nr + 2);
    double *add = (double *)calloc(n1, sizeof(double));
    if (nadd > 0) orc_qdeposit(xa, qa, nadd, dr, nr, max_mode, add);
    for (size_t i = 0; i < n1; i++) rho_ion[i] = rho_ion[i] + add[i];
    for (size_t i = 0; i < n1; i++) q_tot[i] = q_tot[i] + rho_ion[i];
    free(add);
}

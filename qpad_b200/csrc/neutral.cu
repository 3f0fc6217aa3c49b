// neutral.cu -- field-ionisation (ADK) neutral species, SURVEY.md §8(f) rank 2.  species/neutral_class.f03 of the reference:
//
//   :600-753  ionize_neutral   -> k_neutral_ionize   (one thread per (radial cell, theta sector): rate equations of the levels)
//   :755-837  add_particles    -> k_neutral_scan + k_neutral_add (the released electrons are appended in the reference's order
//                                 -- sector, cell, index -- through an exclusive scan over the cells; the same positions with
//                                 the opposite charge fill the ions' position buffer)
//   :839-878  renew            -> qpg_neutral_reset
//   :880-930  qdeposit / ion_deposit, :932-1016 amjdeposit / push_u / push_x reuse the part2d kernels on the two particle sets.
//
// STATUS: parity with oracle/qpad_oracle_neutral.c on the GPU (tests/test_gpu_neutral.py: levels and released electrons bit-exact in
// order and count, the ionisation slice loop of config 5 to 1e-7) and in the host emulation; the neutral's state also travels
// along the xi-pipeline (qpg_sim_neutral_pack / _unpack below, neut%psend / precv :1018-1101).
//
// Level array on the device: lev[(i * n_theta + k) * nr + j], i = 0..multi_max-1 charge states 1..multi_max, i = multi_max
// neutral residue, i = multi_max + 1 total discrete ion level (the layout of the oracle).
#include "common.cuh"

struct qpg_neutral_s {
    qpg_ctx ctx;
    int multi_max, n_theta, ppc1, ppc2;
    double wp, dt, qm, density, den_min;
    double *lev, *ion_old;
    int *cnt, *off, *d_nadd;
    double adk[60];
};
struct AdkTable { double v[60]; };

// ADK rate parameters (r1 [1/s], r2 [GV/m], r3) per charge state: r2 = 6.83 xi^1.5, r3 = 2 n* - 1,
// r1 = 1.52e15 4^n* xi / (n* Gamma(2 n*)) (20.5 xi^1.5)^(2 n* - 1), n* = 3.69 Z / sqrt(xi), xi = ionisation energy in eV (NIST ASD).
// Values as the reference tabulates them (neutral_class.f03:39-52, :177-180) for the elements its decks use.
static const double ADK_H[3] = {8.522542995398661e19, 342.53947239007687, 1.0005337056631487};
static const double ADK_HE[6] = {7.2207661763501e18, 832.809878216992, 0.48776427204592254, 2.7226733893691e21, 2742.1316798375965, 1.0000920088118899};
static const double ADK_LI[9] = {3.460272990838495e21, 85.51998980232813, 2.1770706138013733, 3.6365138642921554e20, 4493.713340713575, 0.6964625952167312,
                                 2.0659396971422902e22, 9256.32561931876, 0.9999745128918196};

// neutral_class.f03:600-753.  ef = node-interleaved dim-3 image of E.  Keeps the reference's quirks: the imaginary plane of the
// m > 0 modes is read from the REAL plane (:636-637); the level update is its "2nd order Runge-Kutta" (:661, :688).
// (no __restrict__ on the arrays the steps hand to each other: inside k_neutral_update one step's output is the next one's input)
__device__ __forceinline__ void neutral_ionize(double *lev, double *ion_old, int *cnt, const AdkTable &adk,
                                               const double *__restrict__ ef, double wp, double dt, int ppc_tot, int nr, int n_theta, int M, int multi_max)
{
    const int P = 2 * M + 1, idx_neut = multi_max, idx_ion = multi_max + 1;
    const int n = nr * n_theta;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int k = t / nr, j = t % nr;                         // sector, 0-based cell (nodes j+1, j+2)
#define NLEV(i) lev[((size_t)(i) * n_theta + k) * nr + j]
#define NEF(pl, c, node) ef[((size_t)(node) * P + (pl)) * 3 + (c)]
        const double old = NLEV(idx_ion);
        ion_old[(size_t)k * nr + j] = old;
        const double theta = 2.0 * 3.14159265358979323846 / (double)n_theta * (double)k;
        const double incr = cos(theta), inci = sin(theta);
        double e1 = 0.5 * (NEF(0, 0, j + 1) + NEF(0, 0, j + 2)), e2 = 0.5 * (NEF(0, 1, j + 1) + NEF(0, 1, j + 2)), e3 = 0.5 * (NEF(0, 2, j + 1) + NEF(0, 2, j + 2));
        double phr = 1.0, phi = 0.0;
        for (int m = 1; m <= M; m++) {
            const int pr = 2 * m - 1;                             // both the "real" and the "imaginary" pointer of the reference
            const double tt = phr * incr - phi * inci;
            phi = phr * inci + phi * incr;
            phr = tt;
            e1 = e1 + (NEF(pr, 0, j + 1) + NEF(pr, 0, j + 2)) * phr - (NEF(pr, 0, j + 1) + NEF(pr, 0, j + 2)) * phi;
            e2 = e2 + (NEF(pr, 1, j + 1) + NEF(pr, 1, j + 2)) * phr - (NEF(pr, 1, j + 1) + NEF(pr, 1, j + 2)) * phi;
            e3 = e3 + (NEF(pr, 2, j + 1) + NEF(pr, 2, j + 2)) * phr - (NEF(pr, 2, j + 1) + NEF(pr, 2, j + 2)) * phi;
        }
        const double eij = sqrt(e1 * e1 + e2 * e2 + e3 * e3) * wp * 1.708e-12;   // GV/m
        if (eij > 1.0e-6 && NLEV(idx_ion) < (double)multi_max) {
            double w_ion[20];
            for (int i = 0; i < multi_max; i++) w_ion[i] = adk.v[3 * i] * pow(eij, -adk.v[3 * i + 2]) * exp(-adk.v[3 * i + 1] / eij) / wp;
            bool shoot = false;
            double cons = NLEV(idx_neut) * w_ion[0] * dt * (1.0 + 0.5 * w_ion[0] * dt);
            if (cons > NLEV(idx_neut)) { shoot = true; cons = NLEV(idx_neut); }
            NLEV(idx_neut) = NLEV(idx_neut) - cons;
            for (int i = 0; i < multi_max - 1; i++) {
                const double inj = cons;
                double dens_temp = 0.0;
                if (shoot) { dens_temp = inj * 0.5; shoot = false; }
                cons = (NLEV(i) + dens_temp) * w_ion[i + 1] * dt * (1.0 + 0.5 * w_ion[i + 1] * dt);
                if (cons > NLEV(i) + dens_temp) { shoot = true; cons = NLEV(i) + dens_temp; }
                NLEV(i) = fmin(NLEV(i) - cons + inj, 1.0);
            }
            NLEV(multi_max - 1) = fmin(NLEV(multi_max - 1) + cons, 1.0);
            double tot = 0.0;
            for (int i = 0; i < multi_max; i++) tot = tot + (double)(i + 1) * NLEV(i);
            NLEV(idx_ion) = (double)multi_max / (double)ppc_tot * (double)(int)(tot * (double)ppc_tot / (double)multi_max + 0.5);
        }
        // add_particles :793: macro-electrons this cell releases now
        cnt[t] = (int)((NLEV(idx_ion) - old) / (double)multi_max * (double)ppc_tot + 0.5);
#undef NLEV
#undef NEF
    }
}

__global__ void k_neutral_ionize(double *__restrict__ lev, double *__restrict__ ion_old, int *__restrict__ cnt, AdkTable adk, const double *__restrict__ ef,
                                 double wp, double dt, int ppc_tot, int nr, int n_theta, int M, int multi_max)
{ neutral_ionize(lev, ion_old, cnt, adk, ef, wp, dt, ppc_tot, nr, n_theta, M, multi_max); }

// exclusive scan of cnt[0..n) in index order (sector-major, cell-minor = the reference's loop order) by ONE CTA of 1024 threads
__device__ __forceinline__ void neutral_scan(const int *cnt, int *off, int *d_nadd, int n)
{
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (n + 1023) / 1024, beg = min(t * per, n), end = min(beg + per, n);
    int s = 0;
    for (int i = beg; i < end; i++) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {        // Hillis-Steele inclusive scan of the chunk sums
        const int v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int i = beg; i < end; i++) { off[i] = run; run += cnt[i]; }
    if (t == 1023) *d_nadd = part[1023];
}
__global__ void __launch_bounds__(1024, 1) k_neutral_scan(const int *__restrict__ cnt, int *__restrict__ off, int *__restrict__ d_nadd, int n) { neutral_scan(cnt, off, d_nadd, n); }

// neutral_class.f03:795-829: the electrons of cell (k, j) at r = (j + (i + 1/2) / cnt) dr, theta = k dtheta, at rest;
// the ions' buffer gets the same positions with the opposite charge
__device__ __forceinline__ void neutral_add(const int *cnt, const int *off, const PartView &pe, const PartView &pi, long cap_e, long cap_i,
                                            double dr, double density, double den_min, double coef, int nr, int n_theta)
{
    const int n = nr * n_theta, base = *pe.d_npp;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int c = cnt[t];
        if (c <= 0) continue;
        const int k = t / nr, j = t % nr;
        const double theta = (double)k * (2.0 * 3.14159265358979323846 / (double)n_theta);
        const double ct = cos(theta), st = sin(theta);
        for (int i = 0; i < c; i++) {
            const double rn = (double)j + ((double)(i + 1) - 0.5) / (double)c;
            const double x1 = rn * dr * ct, x2 = rn * dr * st, q = rn * 1.0 * 1.0 * density * coef;
            const long a = (long)base + off[t] + i, b = (long)off[t] + i;
            if (density < den_min) continue;     // uniform / uniform profile: den_lon = den_perp = 1
            if (a < cap_e) { pe.x1[a] = x1; pe.x2[a] = x2; pe.p1[a] = 0.0; pe.p2[a] = 0.0; pe.p3[a] = 0.0; pe.gamma[a] = 1.0; pe.psi[a] = 0.0; pe.q[a] = q; }
            if (b < cap_i) { pi.x1[b] = x1; pi.x2[b] = x2; pi.p1[b] = 0.0; pi.p2[b] = 0.0; pi.p3[b] = 0.0; pi.gamma[b] = 1.0; pi.psi[b] = 0.0; pi.q[b] = -q; }
        }
    }
}
__global__ void k_neutral_add(const int *__restrict__ cnt, const int *__restrict__ off, PartView pe, PartView pi, long cap_e, long cap_i, double dr, double density,
                              double den_min, double coef, int nr, int n_theta)
{ neutral_add(cnt, off, pe, pi, cap_e, cap_i, dr, density, den_min, coef, nr, n_theta); }
// flags[7] of the context is latched when released electrons (or ion positions) did not fit the particle sets: the host reports
// QPG_ERR_STATE at its next synchronisation point instead of silently losing charge (beam.cu does the same for its wire buffer)
__device__ __forceinline__ void neutral_counts(int *npp_e, int *npp_i, const int *d_nadd, long cap_e, long cap_i, int *ctx_flags)
{
    const int add = *d_nadd;
    if ((long)*npp_e + add > cap_e || (long)add > cap_i) ctx_flags[7] = 1;
    *npp_e = (int)min((long)*npp_e + add, cap_e);
    *npp_i = (int)min((long)add, cap_i);
}
__global__ void k_neutral_counts(int *npp_e, int *npp_i, const int *d_nadd, long cap_e, long cap_i, int *ctx_flags) { neutral_counts(npp_e, npp_i, d_nadd, cap_e, cap_i, ctx_flags); }
// The whole update as ONE kernel of one CTA (the four steps above with CTA barriers in between): on the per-slice launch path a slice is a chain
// of ~30 small dependent kernels and a pipeline of many stages is paced by the device's kernel dispatch -- three launches less
// per slice count; 16 000 (cell, sector) pairs are 16 per thread
__global__ void __launch_bounds__(1024, 1) k_neutral_update(double *lev, double *ion_old, int *cnt, int *off, int *d_nadd, AdkTable adk, const double *__restrict__ ef, double wp, double dt, int ppc_tot, int nr,
                                                           int n_theta, int M, int multi_max, PartView pe, PartView pi, long cap_e, long cap_i, double dr, double density,
                                                           double den_min, double coef, int *npp_e, int *npp_i, int *ctx_flags)
{
    neutral_ionize(lev, ion_old, cnt, adk, ef, wp, dt, ppc_tot, nr, n_theta, M, multi_max);
    __syncthreads();
    neutral_scan(cnt, off, d_nadd, nr * n_theta);
    __syncthreads();
    neutral_add(cnt, off, pe, pi, cap_e, cap_i, dr, density, den_min, coef, nr, n_theta);
    __syncthreads();
    if (threadIdx.x == 0) neutral_counts(npp_e, npp_i, d_nadd, cap_e, cap_i, ctx_flags);
}
// most electrons one update can release: every (cell, sector) ionises all its ppc particles at once
long qpg_neutral_max_new_per_update(qpg_neutral ne) { return ne ? (long)ne->ppc1 * ne->ppc2 * ne->ctx->nr * ne->n_theta : 0; }

extern "C" int qpg_neutral_create(qpg_neutral *out, qpg_ctx ctx, int element, int ion_max, int ppc1, int ppc2, int num_theta, double q, double m, double density,
                                  double n0, double dt_xi)
{
    ARG_TRY(out && ctx && ion_max >= 1 && ppc1 >= 1 && ppc2 >= 1 && num_theta >= 1 && m != 0.0 && n0 > 0.0, "bad arg");
    const double *tab = nullptr;
    int nlev = 0;
    switch (element) {      // param.f03:135-148 (atomic numbers)
    case 1: tab = ADK_H; nlev = 1; break;
    case 2: tab = ADK_HE; nlev = 2; break;
    case 3: tab = ADK_LI; nlev = 3; break;
    default: qpg_set_error("Invalid neutral gas species! (supported: H, He, Li)"); return QPG_ERR_UNSUPPORTED;
    }
    qpg_neutral ne = new qpg_neutral_s();
    memset(ne, 0, sizeof(*ne));
    ne->ctx = ctx; ne->multi_max = ion_max < nlev ? ion_max : nlev; ne->n_theta = num_theta; ne->ppc1 = ppc1; ne->ppc2 = ppc2;
    ne->wp = sqrt(n0) * 5.641460231180626e4;                     // sim_plasma_class.f03:84
    ne->dt = dt_xi; ne->qm = q / m; ne->density = density; ne->den_min = 1e-10;
    for (int i = 0; i < 3 * ne->multi_max; i++) ne->adk[i] = tab[i];
    const size_t nc = (size_t)ctx->nr * num_theta;
    CUDA_TRY(cudaMalloc(&ne->lev, sizeof(double) * (ne->multi_max + 2) * nc));
    CUDA_TRY(cudaMalloc(&ne->ion_old, sizeof(double) * nc));
    CUDA_TRY(cudaMalloc(&ne->cnt, sizeof(int) * nc));
    CUDA_TRY(cudaMalloc(&ne->off, sizeof(int) * nc));
    CUDA_TRY(cudaMalloc(&ne->d_nadd, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(ne->d_nadd, 0, sizeof(int), ctx->stream));
    *out = ne;
    return qpg_neutral_reset(ne);
}
__global__ void k_neutral_reset(double *lev, size_t nc, int multi_max)
{
    const size_t n = nc * (multi_max + 2);
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) lev[t] = (t / nc == (size_t)multi_max) ? 1.0 : 0.0;
}
extern "C" int qpg_neutral_reset(qpg_neutral ne)
{
    ARG_TRY(ne, "null arg");
    const size_t nc = (size_t)ne->ctx->nr * ne->n_theta;
    k_neutral_reset<<<64, 256, 0, ne->ctx->stream>>>(ne->lev, nc, ne->multi_max);      // neutral_class.f03:562-563, :864-865
    count_launch(ne->ctx);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_neutral_destroy(qpg_neutral ne)
{
    if (!ne) return 0;
    cudaStreamSynchronize(ne->ctx->stream);
    cudaFree(ne->lev); cudaFree(ne->ion_old); cudaFree(ne->cnt); cudaFree(ne->off); cudaFree(ne->d_nadd);
    delete ne;
    return 0;
}
extern "C" int qpg_neutral_multi_max(qpg_neutral ne) { return ne ? ne->multi_max : -1; }
// neutral%update (:576-598): ionize with the slice's E, then create the released electrons (appended to `electrons`) and
// fill `ions` (replaced) with their positions and the opposite charge for the next slice's ion_deposit
extern "C" int qpg_neutral_update(qpg_neutral ne, qpg_field e, qpg_part2d electrons, qpg_part2d ions)
{
    ARG_TRY(ne && e && electrons && ions && e->dim == 3 && e->ctx == ne->ctx && electrons->ctx == ne->ctx && ions->ctx == ne->ctx, "bad handle");
    qpg_ctx c = ne->ctx;
    const int nc = c->nr * ne->n_theta;
    AdkTable tab;
    memcpy(tab.v, ne->adk, sizeof(tab.v));
    const double coef = (double)ne->multi_max * (ne->qm < 0 ? -1.0 : 1.0) / ((double)(ne->ppc1 * ne->ppc2) * (double)ne->n_theta);
    static const bool split = getenv("QPG_NEUTRAL_SPLIT_UPDATE") != nullptr;     // A/B: the four kernels one by one
    if (!split && nc <= 64 * 1024) {
        k_neutral_update<<<1, 1024, 0, c->stream>>>(ne->lev, ne->ion_old, ne->cnt, ne->off, ne->d_nadd, tab, e->f1, ne->wp, ne->dt, ne->ppc1 * ne->ppc2, c->nr, ne->n_theta,
                                                    c->M, ne->multi_max, view_of(electrons), view_of(ions), electrons->npmax, ions->npmax, c->dr, ne->density,
                                                    ne->den_min, coef, electrons->d_npp, ions->d_npp, c->flags);
        count_launch(c);
    } else {
        k_neutral_ionize<<<(nc + 127) / 128, 128, 0, c->stream>>>(ne->lev, ne->ion_old, ne->cnt, tab, e->f1, ne->wp, ne->dt, ne->ppc1 * ne->ppc2, c->nr, ne->n_theta, c->M,
                                                                 ne->multi_max);
        k_neutral_scan<<<1, 1024, 0, c->stream>>>(ne->cnt, ne->off, ne->d_nadd, nc);
        k_neutral_add<<<(nc + 127) / 128, 128, 0, c->stream>>>(ne->cnt, ne->off, view_of(electrons), view_of(ions), electrons->npmax, ions->npmax, c->dr, ne->density,
                                                             ne->den_min, coef, c->nr, ne->n_theta);
        k_neutral_counts<<<1, 1, 0, c->stream>>>(electrons->d_npp, ions->d_npp, ne->d_nadd, electrons->npmax, ions->npmax, c->flags);
        count_launch(c, 4);
    }
    CUDA_TRY(cudaGetLastError());
    electrons->npp_hi = electrons->npmax;     // unknown until the next sync; the kernels bound themselves by the device count
    ions->npp_hi = ions->npmax;
    return 0;
}
extern "C" int qpg_neutral_levels(qpg_neutral ne, double *host)
{
    ARG_TRY(ne && host, "null arg");
    const size_t n = (size_t)(ne->multi_max + 2) * ne->n_theta * ne->ctx->nr;
    CUDA_TRY(cudaMemcpyAsync(host, ne->lev, sizeof(double) * n, cudaMemcpyDeviceToHost, ne->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ne->ctx->stream));
    return 0;
}
// empties a particle set on the device (neutral%renew :873-874: part%npp = 0, part_add%npp = 0)
extern "C" int qpg_part2d_clear(qpg_part2d p)
{
    ARG_TRY(p, "null arg");
    CUDA_TRY(cudaMemsetAsync(p->d_npp, 0, 2 * sizeof(int), p->ctx->stream));
    p->npp_hi = 0;
    return 0;
}

// ---- the neutral on a xi-pipeline: neut%psend / precv (neutral_class.f03:1025-1101) ----------------------------------------------
// What travels to the next stage after a slab: the released electrons (`part`), the ions' position buffer of the last update (`part_add`),
// the ion density image rho_ion (pipe_send 'forward' / pipe_recv 'replace') and the level array multi_ion.  One wire record at fixed
// offsets: [electrons: qpg_part2d_wire_count][ions: qpg_part2d_wire_count][rho_ion f1: (nr+2) P][levels: (multi_max+2) num_theta nr];
// the particle records copy their live prefix only.  The look-ahead deposits of the slab's first slice are redone by qpg_sim_run_slices.
extern "C" long qpg_sim_neutral_wire_count(qpg_sim s)
{
    if (!s || !s->neut) return -1;
    return qpg_part2d_wire_count(s->neut_e) + qpg_part2d_wire_count(s->neut_i) + (long)s->rho_ion->n1 + (long)(s->neut->multi_max + 2) * s->neut->n_theta * s->ctx->nr;
}
extern "C" int qpg_sim_neutral_pack(qpg_sim s, double *dev_buf)
{
    ARG_TRY(s && s->neut && dev_buf, "no neutral attached / null buffer");
    int rc;
    double *b = dev_buf;
    if ((rc = qpg_part2d_pack(s->neut_e, b))) return rc;
    b += qpg_part2d_wire_count(s->neut_e);
    if ((rc = qpg_part2d_pack(s->neut_i, b))) return rc;
    b += qpg_part2d_wire_count(s->neut_i);
    cudaStream_t st = s->ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(b, s->rho_ion->f1, sizeof(double) * s->rho_ion->n1, cudaMemcpyDeviceToDevice, st));
    b += s->rho_ion->n1;
    CUDA_TRY(cudaMemcpyAsync(b, s->neut->lev, sizeof(double) * (size_t)(s->neut->multi_max + 2) * s->neut->n_theta * s->ctx->nr, cudaMemcpyDeviceToDevice, st));
    return 0;
}
extern "C" int qpg_sim_neutral_unpack(qpg_sim s, const double *dev_buf)
{
    ARG_TRY(s && s->neut && dev_buf, "no neutral attached / null buffer");
    int rc;
    const double *b = dev_buf;
    if ((rc = qpg_part2d_unpack(s->neut_e, b))) return rc;
    b += qpg_part2d_wire_count(s->neut_e);
    if ((rc = qpg_part2d_unpack(s->neut_i, b))) return rc;
    b += qpg_part2d_wire_count(s->neut_i);
    cudaStream_t st = s->ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(s->rho_ion->f1, b, sizeof(double) * s->rho_ion->n1, cudaMemcpyDeviceToDevice, st));
    b += s->rho_ion->n1;
    CUDA_TRY(cudaMemcpyAsync(s->neut->lev, b, sizeof(double) * (size_t)(s->neut->multi_max + 2) * s->neut->n_theta * s->ctx->nr, cudaMemcpyDeviceToDevice, st));
    return 0;
}

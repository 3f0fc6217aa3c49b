// vpot.cu -- the vector-potential diagnostic fields of QPAD, SURVEY.md §8(f) rank 4 (fields/field_vpot_class.f03):
//
//   :354-390  solve_field_vpotz  -> qpg_solve_vpotz : lap_m A_z = -J_z per mode (set_source :159, get_solution :260: A_z of the
//                                                     m > 0 modes vanishes on the axis)
//   :392-431  solve_field_vpott  -> qpg_solve_vpott : A_+ = A_r + i A_phi obeys lap_(m+1), A_- = A_r - i A_phi obeys lap_(m-1)
//                                                     (sources :207-258, recombination and axis rules :297-352)
//   operator rows: fields/field_solver_class.f03:418-484 (p_fk_vpotz / p_fk_vpotp / p_fk_vpotm), outer boundary :532-540
//
// These solves run on dump steps only (a diagnostic), never in the slice loop, so they are kept OUT of the operator table of
// the sweep kernel (common.cuh FK_*): ONE CTA builds the right-hand sides of all (plane, kind) systems, one thread per system
// runs the Thomas recurrence with the elimination coefficients pre-computed on the host (the sequential chain of 2 nr steps is
// ~20 us at nr = 1024, irrelevant at dump cadence), and the CTA recombines the solutions.  The recurrence uses non-contracted
// IEEE operations in the order of the oracle's Thomas solve, so the result is bit-identical to oracle/qpad_oracle.c
// (orc_solve_vpotz / orc_solve_vpott); HYPRE's cyclic reduction of the reference agrees to O(cond * eps) like every other solve.
//
// STATUS: parity with the oracle on the GPU (tests/test_gpu_extras.py, run by default since round 2) and in the host emulation of tests/emu.
#include "common.cuh"
#include <cmath>

enum { VK_Z = 0, VK_P = 1, VK_M = 2 };
struct VpotOps {
    int nr, M, bnd;
    double dr;
    double *coef;      // device: [(kind * (M + 1) + m) * 3 + {a, cp, den}][nr]
    double *rhs;       // device scratch: [2 + 4 M or P systems][nr]
};
static std::map<qpg_ctx, VpotOps> g_vpot;

// fields/field_solver_class.f03:418-484 + :532-540, rows in units of 1 / dr^2; then the Thomas elimination coefficients
static void vpot_rows(int kind, int m, int nr, double dr, int bnd, std::vector<double> &a, std::vector<double> &cp, std::vector<double> &den)
{
    std::vector<double> b(nr), c(nr);
    a.assign(nr, 0.0); cp.assign(nr, 0.0); den.assign(nr, 0.0);
    const int k = kind == VK_Z ? m : (kind == VK_P ? m + 1 : m - 1);
    double j = 0.0;
    for (int i = 1; i < nr; i++) {
        j = j + 1.0;
        a[i] = 1.0 - 0.5 / j;
        c[i] = 1.0 + 0.5 / j;
        b[i] = kind == VK_Z ? -2.0 - (double)(m * m) / (j * j) : -2.0 - ((double)k / j) * ((double)k / j);
    }
    const bool axis_coupled = kind == VK_Z ? (m == 0) : (kind == VK_M ? (m == 1) : false);
    if (axis_coupled) { a[0] = 0.0; b[0] = -4.0; c[0] = 4.0; }
    else { a[0] = 0.0; b[0] = 1.0; c[0] = 0.0; a[1] = 0.0; }
    if (bnd == QPG_BND_ZERO) c[nr - 1] = 0.0;
    else {
        const double jmax = (double)nr;
        if (kind == VK_Z) { if (m != 0) b[nr - 1] = b[nr - 1] + (1.0 - (double)m / jmax) * c[nr - 1]; }
        else b[nr - 1] = b[nr - 1] + (1.0 - (double)(m + 1) / jmax) * c[nr - 1];        // both vpotp and vpotm use m + 1 (:532-540)
        c[nr - 1] = 0.0;
    }
    const double dr2 = dr * dr;
    for (int i = 0; i < nr; i++) { a[i] = a[i] / dr2; b[i] = b[i] / dr2; c[i] = c[i] / dr2; }
    den[0] = b[0];
    cp[0] = c[0] / b[0];
    for (int i = 1; i < nr; i++) { den[i] = b[i] - a[i] * cp[i - 1]; cp[i] = c[i] / den[i]; }
}

static int vpot_ops(qpg_ctx c, VpotOps **out)
{
    VpotOps &o = g_vpot[c];
    if (o.coef && (o.nr != c->nr || o.M != c->M || o.bnd != c->bnd || o.dr != c->dr)) {       // a recycled context pointer
        cudaFree(o.coef); cudaFree(o.rhs); o.coef = nullptr; o.rhs = nullptr;
    }
    if (!o.coef) {
        o.nr = c->nr; o.M = c->M; o.bnd = c->bnd; o.dr = c->dr;
        const size_t nr = c->nr, nset = (size_t)3 * (c->M + 1);
        std::vector<double> h(nset * 3 * nr), a, cp, den;
        for (int kind = 0; kind < 3; kind++) for (int m = 0; m <= c->M; m++) {
            vpot_rows(kind, m, c->nr, c->dr, c->bnd, a, cp, den);
            double *dst = h.data() + ((size_t)kind * (c->M + 1) + m) * 3 * nr;
            memcpy(dst, a.data(), 8 * nr); memcpy(dst + nr, cp.data(), 8 * nr); memcpy(dst + 2 * nr, den.data(), 8 * nr);
        }
        CUDA_TRY(cudaMalloc(&o.coef, sizeof(double) * h.size()));
        CUDA_TRY(cudaMalloc(&o.rhs, sizeof(double) * (size_t)(4 * c->M + 2 + c->P) * nr));
        CUDA_TRY(cudaMemcpyAsync(o.coef, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));     // h goes out of scope
    }
    *out = &o;
    return 0;
}
extern "C" int qpg_vpot_release(qpg_ctx c)
{
    auto it = g_vpot.find(c);
    if (it != g_vpot.end()) { cudaFree(it->second.coef); cudaFree(it->second.rhs); g_vpot.erase(it); }
    return 0;
}

#define VF(f, j, pl, c) (f)[((size_t)(j) * P + (pl)) * 3 + (c)]
__device__ __forceinline__ void vpot_thomas(const double *__restrict__ co, double *__restrict__ d, int nr)
{
    const double *a = co, *cp = co + nr, *den = co + 2 * nr;
    double prev = __ddiv_rn(d[0], den[0]);
    d[0] = prev;
    for (int i = 1; i < nr; i++) { prev = __ddiv_rn(__dsub_rn(d[i], __dmul_rn(a[i], prev)), den[i]); d[i] = prev; }
    for (int i = nr - 2; i >= 0; i--) { prev = __dsub_rn(d[i], __dmul_rn(cp[i], prev)); d[i] = prev; }
}

// which = 0: A_z (component 3) ; which = 1: A_r, A_phi (components 1, 2).  One CTA.
__global__ void __launch_bounds__(256, 1) k_vpot(const double *__restrict__ cu, double *__restrict__ vp, const double *__restrict__ coef, double *__restrict__ rhs,
                                                 int nr, int M, int which)
{
    const int P = 2 * M + 1, tid = threadIdx.x, nt = blockDim.x;
    const int nsys = which == 0 ? P : 2 + 4 * M;
    // 1. sources (:159-258)
    for (int idx = tid; idx < nsys * nr; idx += nt) {
        const int s = idx / nr, i = idx % nr + 1;
        double v;
        if (which == 0) v = -1.0 * VF(cu, i, s, 2);
        else if (s < 2) v = -VF(cu, i, 0, s);                                   // m = 0: buf1 = -J_r (vpotp), buf2 = -J_phi (vpotm)
        else {
            const int m = (s - 2) / 4 + 1, t = (s - 2) % 4, pr = 2 * m - 1, pi = 2 * m;
            if (t == 0) v = __dadd_rn(-VF(cu, i, pr, 0), VF(cu, i, pi, 1));      // b1r = -J_r,re + J_phi,im
            else if (t == 1) v = __dsub_rn(-VF(cu, i, pi, 0), VF(cu, i, pr, 1)); // b1i = -J_r,im - J_phi,re
            else if (t == 2) v = __dsub_rn(-VF(cu, i, pr, 0), VF(cu, i, pi, 1)); // b2r = -J_r,re - J_phi,im
            else v = __dadd_rn(-VF(cu, i, pi, 0), VF(cu, i, pr, 1));             // b2i = -J_r,im + J_phi,re
        }
        rhs[(size_t)s * nr + (i - 1)] = v;
    }
    __syncthreads();
    // 2. one thread per system
    if (tid < nsys) {
        int kind, m;
        if (which == 0) { kind = VK_Z; m = (tid + 1) / 2; }
        else if (tid < 2) { kind = tid == 0 ? VK_P : VK_M; m = 0; }
        else { m = (tid - 2) / 4 + 1; kind = ((tid - 2) % 4) < 2 ? VK_P : VK_M; }
        vpot_thomas(coef + ((size_t)kind * (M + 1) + m) * 3 * nr, rhs + (size_t)tid * nr, nr);
    }
    __syncthreads();
    // 3. solutions (:260-352)
    if (which == 0) {
        for (int idx = tid; idx < P * nr; idx += nt) {
            const int pl = idx / nr, i = idx % nr + 1;
            VF(vp, i, pl, 2) = (pl > 0 && i == 1) ? 0.0 : rhs[(size_t)pl * nr + (i - 1)];
        }
        return;
    }
    for (int idx = tid; idx < (M + 1) * nr; idx += nt) {
        const int m = idx / nr, i = idx % nr + 1;
        if (m == 0) {
            VF(vp, i, 0, 0) = i == 1 ? 0.0 : rhs[i - 1];
            VF(vp, i, 0, 1) = i == 1 ? 0.0 : rhs[(size_t)nr + (i - 1)];
            continue;
        }
        const int pr = 2 * m - 1, pi = 2 * m;
        const double *b = rhs + (size_t)(2 + 4 * (m - 1)) * nr;
        const double b1r = b[i - 1], b1i = b[nr + i - 1], b2r = b[2 * nr + i - 1], b2i = b[3 * nr + i - 1];
        const bool ax = (i == 1 && m != 1);
        VF(vp, i, pr, 0) = ax ? 0.0 : __dmul_rn(0.5, __dadd_rn(b1r, b2r));
        VF(vp, i, pi, 0) = ax ? 0.0 : __dmul_rn(0.5, __dadd_rn(b1i, b2i));
        VF(vp, i, pr, 1) = ax ? 0.0 : __dmul_rn(0.5, __dsub_rn(b1i, b2i));
        VF(vp, i, pi, 1) = ax ? 0.0 : __dmul_rn(0.5, __dadd_rn(-b1r, b2r));
    }
}
#undef VF

static int vpot_solve(qpg_ctx ctx, qpg_field cu, qpg_field vpot, int which)
{
    ARG_TRY(ctx && cu && vpot && cu->dim == 3 && vpot->dim == 3 && cu->ctx == ctx && vpot->ctx == ctx, "field handle has wrong dim or context");
    VpotOps *o = nullptr;
    int rc = vpot_ops(ctx, &o);
    if (rc) return rc;
    k_vpot<<<1, 256, 0, ctx->stream>>>(cu->f1, vpot->f1, o->coef, o->rhs, ctx->nr, ctx->M, which);
    count_launch(ctx);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_solve_vpotz(qpg_ctx ctx, qpg_field cu, qpg_field vpot) { return vpot_solve(ctx, cu, vpot, 0); }
extern "C" int qpg_solve_vpott(qpg_ctx ctx, qpg_field cu, qpg_field vpot) { return vpot_solve(ctx, cu, vpot, 1); }

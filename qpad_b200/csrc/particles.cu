// particles.cu -- part2d: the per-slice plasma particle kernels (species/part2d_class.f03, interp_part2d.f03).
//
// Design:
//  * SoA fp64 planes, one particle per thread, fully coalesced 8-byte loads/stores.
//  * gathers read the node-interleaved field images through the read-only path; particles that are close in the
//    array are close in r (lattice order / counting sort), so a warp reads one or two 72-byte node records.
//  * deposits never issue one atomic per particle: the sums over the lanes of a warp that share a radial cell are
//    small contractions (node weight x mode phase) (x) (momentum moments), done on the fp64 tensor pipe with
//    mma.sync.m8n8k4 through a per-warp shared-memory tile, then ONE fp64 RED per (cell, plane, component) from
//    distinct lanes.  Only warps scattered over more than 8 cells fall back to per-lane REDs (charge deposit).
//  * cell index / boundary tests use non-contracted IEEE ops so indices match the reference bit for bit
//    (pos = sqrt(x1*x1 + x2*x2) * (1/dr), interp_part2d.f03:44-59).
#include "common.cuh"

#ifndef FULL
#define FULL 0xffffffffu
#endif
// Defaults since round 2 (A/B on a B200, profiles/r02_ab_variants.txt: -3.3 % step time; measured max error of the short Newton
// chains on the device: 0.51 ulp reciprocal, 0.50 ulp square root -- tests/test_gpu_extras.py::test_fastmath_accuracy, gate 2 ulp).
// -DQPG_LEGACY_MATH restores round 1's longer sequences.
#ifndef QPG_LEGACY_MATH
#define QPG_GATHER_FOLDED 1
#define QPG_FASTMATH_SHORT 1
#endif
#define PT_BLOCK 256

struct PartView {
    double *x1, *x2, *p1, *p2, *p3, *gamma, *psi, *q;
    const int *d_npp;
};
static PartView view_of(qpg_part2d p) { PartView v{p->x1, p->x2, p->p1, p->p2, p->p3, p->gamma, p->psi, p->q, p->d_npp}; return v; }

struct Interp { double c, s, w0, w1; int idx; };

// Reciprocal / square root for the momentum arithmetic: MUFU seed (2^-22) + two Newton steps, no denormal / special
// paths (arguments are gamma-like, >= O(1e-300)).  Within 1-2 ulp of the IEEE result at a third of the instructions
// and latency of the compiler's division.  NOT used for anything that decides a cell index or a boundary test: those
// keep __dsqrt_rn / __dmul_rn (bit-exact with the reference).
__device__ __forceinline__ double fast_rcp(double y)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    double e = fma(-y, r, 1.0);
#ifdef QPG_FASTMATH_SHORT
    // one cubic step r (1 + e + e^2) instead of two quadratic ones -- seed error eps ->
    // eps^3 (2^-60 or better for a 2^-20 seed) in 3 instead of 4 dependent FMAs.  qpg_debug_fastmath measures the ulp error on the GPU.
    e = fma(e, e, e);
    return fma(r, e, r);
#else
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    return fma(r, e, r);
#endif
}
__device__ __forceinline__ double rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double fast_sqrt(double x)
{
    double y = rsqrt_seed(x);
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
#ifndef QPG_FASTMATH_SHORT      // the short variant drops this second coupled step: eps -> 1.5 eps^2 (first step) -> ~eps^4 (the final Heron correction)
    r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
#endif
    return fma(fma(-g, g, x), h, g);
}

// accuracy probe of the two routines above on the device (tests/test_gpu_extras.py::test_fastmath_accuracy): out_rcp[i] =
// fast_rcp(x[i]), out_sqrt[i] = fast_sqrt(x[i])
__global__ void k_debug_fastmath(const double *__restrict__ x, double *__restrict__ out_rcp, double *__restrict__ out_sqrt, long n)
{
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { out_rcp[i] = fast_rcp(x[i]); out_sqrt[i] = fast_sqrt(x[i]); }
}
extern "C" int qpg_debug_fastmath(qpg_ctx ctx, long n, const double *host_x, double *host_rcp, double *host_sqrt)
{
    ARG_TRY(ctx && n > 0 && host_x && host_rcp && host_sqrt, "bad arg");
    double *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(double) * 3 * n));
    CUDA_TRY(cudaMemcpyAsync(d, host_x, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    k_debug_fastmath<<<148, 256, 0, ctx->stream>>>(d, d + n, d + 2 * n, n);
    count_launch(ctx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(host_rcp, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(host_sqrt, d + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaFree(d));
    return 0;
}

// fire-and-forget reductions.  Written as PTX `red` because ptxas keeps `atomicAdd` as a returning ATOMG (a ~320-cycle
// round trip per instruction) inside the persistent sweep kernel, where fences / volatile loads are present.
#ifdef QPG_EXP_NO_RED   // bottleneck experiment only (results are wrong): deposits never reach memory
__device__ __forceinline__ void red_add(double *p, double v) { if (v == 1.2345e-300) *p = v; }
#else
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
#endif
__device__ __forceinline__ void red_add(int *p, int v) { asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// species/interp_part2d.f03:28-65 gen_interp_info
__device__ __forceinline__ Interp interp_info(double x1, double x2, double idr)
{
    Interp it;
#ifdef QPG_EXP_CHEAP_INTERP   // bottleneck experiment only (results are wrong in the last bits -> different cells now and then): what would the particle
    // phases gain if cos / sin / r/dr came from stored planes instead of an IEEE square root + reciprocal per pass?
    const double rinv = rsqrt_seed(fma(x1, x1, x2 * x2));
    double r = (x1 * x1 + x2 * x2) * rinv;
#else
    double r = __dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2)));
    const double rinv = fast_rcp(r);
#endif
    it.c = x1 * rinv;
    it.s = x2 * rinv;
    double pos = __dmul_rn(r, idr);
    int ip = (int)pos;
    it.idx = ip + 1;
    double f = pos - (double)ip;
    it.w0 = 1.0 - f;
    it.w1 = f;
    return it;
}

// species/interp_part2d.f03:67-109 interp_field (3-vector) on the node-interleaved image f[(node*P + pl)*3 + c]
template <int M>
__device__ __forceinline__ void gather3(const double *f, const Interp &it, double out[3])
{
    constexpr int P = 2 * M + 1;
#ifdef QPG_EXP_FIXED_NODE   // bottleneck experiment only: every particle gathers from node 1 (one address per warp, always an L1 hit)
    const double *n0 = f + (size_t)(it.idx > 100000 ? it.idx : 1) * (P * 3);
#else
    const double *n0 = f + (size_t)it.idx * (P * 3);
#endif
    const double *n1 = n0 + P * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = n0[c] * it.w0;
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] = fma(n1[c], it.w1, out[c]);
    double phr = 1.0, phi = 0.0;
#pragma unroll
    for (int m = 1; m <= M; m++) {
        double t = phr * it.c - phi * it.s;
        phi = phr * it.s + phi * it.c;
        phr = t;
        const double pr2 = 2.0 * phr, pi2 = 2.0 * phi;
#ifdef QPG_GATHER_FOLDED
        // node weight x mode phase once per particle -- the four
        // products are common to every gather of the particle (the compiler shares them between the e and b gathers) -- then two
        // FMAs per (node, component, mode) instead of a multiply and two FMAs: -8 fp64 instructions per particle at M = 1 in
        // amjdeposit and in push (tools/sass_mix.py).  Same sums in a different association (~1e-16 relative).
        const double a0 = it.w0 * pr2, b0 = -(it.w0 * pi2), a1 = it.w1 * pr2, b1 = -(it.w1 * pi2);
#pragma unroll
        for (int c = 0; c < 3; c++) out[c] = fma(n0[(2 * m) * 3 + c], b0, fma(n0[(2 * m - 1) * 3 + c], a0, out[c]));
#pragma unroll
        for (int c = 0; c < 3; c++) out[c] = fma(n1[(2 * m) * 3 + c], b1, fma(n1[(2 * m - 1) * 3 + c], a1, out[c]));
#else
#pragma unroll
        for (int c = 0; c < 3; c++) out[c] = fma(n0[(2 * m - 1) * 3 + c] * pr2 - n0[(2 * m) * 3 + c] * pi2, it.w0, out[c]);
#pragma unroll
        for (int c = 0; c < 3; c++) out[c] = fma(n1[(2 * m - 1) * 3 + c] * pr2 - n1[(2 * m) * 3 + c] * pi2, it.w1, out[c]);
#endif
    }
}

// ---- segmented warp reduction of the amjdeposit sums on the fp64 tensor pipe -----------------------------------
// Per particle the 2*P*8 deposited values are an outer product  alpha (x) beta :
//     alpha[j*P + pl] = w_j * q*ipsi*{Re,Im}(e^{-im phi})     (2P numbers: node weight x mode phase)
//     beta[k]         = (u_r, u_phi, u_z, du_r, du_phi, u_r u_r, u_r u_phi, u_phi u_phi)/...   (8 numbers)
// so the sum over the particles of one cell is the small contraction  C[2P x 8] = sum_p alpha_p beta_p^T, which is
// exactly what mma.sync.m8n8k4.f64 computes: 8 DMMAs contract the 32 particles of a warp (k = 4 particles each) for
// 8 alpha rows.  alpha/beta go through a per-warp shared-memory tile ([row][lane], row stride DEP_LD = 36 doubles
// -> conflict-free stores and fragment loads); lanes that do not belong to the cell are masked out of the A
// fragment, so a warp spanning several cells repeats only the 8 DMMAs + 2 REDs per lane, not a shuffle butterfly
// (~45 instructions per cell instead of ~620).
#define DEP_LD 36
template <int M> struct DepTile { static constexpr int rows = 2 * (2 * M + 1) + 8, doubles = rows * DEP_LD; };

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int M>
__device__ __forceinline__ void warp_deposit_mma(const double (&alpha)[2 * (2 * M + 1)], const double (&beta)[8], int key, double *acc8, double *tile,
                                                 int lane)
{
    constexpr int P = 2 * M + 1, R = 2 * P, NTILE = (R + 7) / 8;
#ifdef QPG_EXP_NO_DEPOSIT   // bottleneck experiment only
    { double sacc = 0.0; for (int r = 0; r < R; r++) sacc += alpha[r]; for (int k = 0; k < 8; k++) sacc += beta[k]; if (sacc == 1.2345e-300) acc8[0] = sacc; return; }
#endif
    __syncwarp();                                   // the previous tile's fragment loads are done
#pragma unroll
    for (int r = 0; r < R; r++) tile[r * DEP_LD + lane] = alpha[r];
#pragma unroll
    for (int k = 0; k < 8; k++) tile[(R + k) * DEP_LD + lane] = beta[k];
    __syncwarp();
    const int row = lane >> 2, kk = lane & 3;
    double bfr[8];
#pragma unroll
    for (int s = 0; s < 8; s++) bfr[s] = tile[(R + row) * DEP_LD + 4 * s + kk];
    const unsigned same = __match_any_sync(FULL, key);
    const bool leader = key >= 0 && (__ffs(same) - 1 == lane);
    unsigned leaders = __ballot_sync(FULL, leader);
    while (leaders) {
        const int l = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        const int cell = __shfl_sync(FULL, key, l);
        const unsigned memb = __ballot_sync(FULL, key == cell);            // warp-uniform
        const unsigned members = memb >> kk;                               // bit 4s <-> particle 4s + kk
#pragma unroll
        for (int t = 0; t < NTILE; t++) {
            const int r = t * 8 + row;
            const bool live = r < R;
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int s = 0; s < 8; s++) {
                if ((memb >> (4 * s)) & 0xfu) {   // k-steps without a member of this cell contribute nothing: skip (cell-sorted
                    double a = 0.0;               // tiles then need 8 DMMAs in total, not 8 per cell)
                    if (live && ((members >> (4 * s)) & 1u)) a = tile[r * DEP_LD + 4 * s + kk];
                    dmma884(c0, c1, a, bfr[s]);
                }
            }
            if (live) {
                const int j = r >= P ? 1 : 0, pl = r - j * P;
                double *dst = acc8 + ((size_t)(cell + j) * P + pl) * 8 + 2 * kk;
                red_add(dst, c0);
                red_add(dst + 1, c1);
            }
        }
    }
}

// Charge deposit (2P sums per cell): all cells of the warp in ONE pass of 8 DMMAs.  A = alpha rows as above, B = the
// one-hot membership matrix  B[p][g] = (particle p belongs to the g-th distinct cell of the warp), so
// C[row][g] = sum over the particles of cell g of alpha[row].  Up to 8 cells per warp; more scattered warps fall back
// to per-lane REDs.
template <int M>
__device__ __forceinline__ void warp_deposit_q_mma(const double (&alpha)[2 * (2 * M + 1)], int key, double *acc1, double *tile, int lane)
{
    constexpr int P = 2 * M + 1, R = 2 * P, NTILE = (R + 7) / 8;
    const unsigned same = __match_any_sync(FULL, key);
    const int myleader = __ffs(same) - 1;
    const unsigned leaders = __ballot_sync(FULL, key >= 0 && myleader == lane);
    const int ng = __popc(leaders);
    if (ng == 0) return;
    if (ng > 8) {
        if (key >= 0) {
            double *a = acc1 + (size_t)key * P;
#pragma unroll
            for (int r = 0; r < R; r++) red_add(a + r, alpha[r]);
        }
        return;
    }
    const int gidx = key >= 0 ? __popc(leaders & ((1u << myleader) - 1u)) : 8;
    // cell of the c-th group, held by lane c
    const unsigned lsrc = __fns(leaders, 0, lane + 1);
    const int kc = __shfl_sync(FULL, key, lsrc & 31);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; r++) tile[r * DEP_LD + lane] = alpha[r];
    __syncwarp();
    const int row = lane >> 2, kk = lane & 3;
    double bfr[8];
#pragma unroll
    for (int s = 0; s < 8; s++) bfr[s] = (__shfl_sync(FULL, gidx, 4 * s + kk) == row) ? 1.0 : 0.0;
    const int cell0 = __shfl_sync(FULL, kc, 2 * kk), cell1 = __shfl_sync(FULL, kc, 2 * kk + 1);
#pragma unroll
    for (int t = 0; t < NTILE; t++) {
        const int r = t * 8 + row;
        const bool live = r < R;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const double a = live ? tile[r * DEP_LD + 4 * s + kk] : 0.0;
            dmma884(c0, c1, a, bfr[s]);
        }
        if (live) {
            const int j = r >= P ? 1 : 0, pl = r - j * P;
            if (2 * kk < ng) red_add(acc1 + (size_t)(cell0 + j) * P + pl, c0);
            if (2 * kk + 1 < ng) red_add(acc1 + (size_t)(cell1 + j) * P + pl, c1);
        }
    }
}

// ---- qdeposit: species/part2d_class.f03:231-359 (accumulation part; axis rules live in FOP_QFIX) --------
// per-particle charge products (part2d_class.f03:277-289): X[pl] = Re/Im(q * phase0^m), key = cell (1-based), weights
template <int M>
__device__ __forceinline__ void qdep_products(double x1, double x2, double q, double idr, double (&X)[2 * M + 1], double &w0, double &w1, int &key)
{
    double pos = __dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2))), idr);
    // phase0 = cmplx(x1, -x2) / pos * idr   (part2d_class.f03:278)
    const double rpos = fast_rcp(pos) * idr;
    const double c0 = x1 * rpos, s0 = -x2 * rpos;
    int nn = (int)floor(pos);
    double f = pos - (double)nn;
    key = nn + 1;
    w0 = 1.0 - f; w1 = f;
    double phr = q, phi = 0.0;
    X[0] = phr;
#pragma unroll
    for (int m = 1; m <= M; m++) {
        double t = phr * c0 - phi * s0;
        phi = phr * s0 + phi * c0;
        phr = t;
        X[2 * m - 1] = phr;
        X[2 * m] = phi;
    }
}
// one warp-tile of qdeposit: particle i (whole warp participates)
template <int M>
__device__ __forceinline__ void qdep_alpha(double x1, double x2, double q, double idr, double (&alpha)[2 * (2 * M + 1)], int &key)
{
    constexpr int P = 2 * M + 1;
    double X[P], w0, w1;
    qdep_products<M>(x1, x2, q, idr, X, w0, w1, key);
#pragma unroll
    for (int k = 0; k < P; k++) { alpha[k] = w0 * X[k]; alpha[P + k] = w1 * X[k]; }
}
template <int M>
__device__ __forceinline__ void qdep_body(const PartView &pv, double *acc1, double idr, int npp, int i, int lane, double *tile)
{
    constexpr int P = 2 * M + 1;
    double alpha[2 * P];
    int key = -1;
    if (i < npp) qdep_alpha<M>(pv.x1[i], pv.x2[i], pv.q[i], idr, alpha, key);
    else {
#pragma unroll
        for (int k = 0; k < 2 * P; k++) alpha[k] = 0.0;
    }
    warp_deposit_q_mma<M>(alpha, key, acc1, tile, lane);
}
// STRIDE: grid-stride over the tiles under a capped grid (pt_grid) for a set whose live count only the device knows (npp_hi = capacity: a
// neutral's electrons, a set just unpacked), which would otherwise pay for thousands of empty blocks.  A separate instantiation: the loop
// costs registers (amjdeposit 64 -> 72, qdeposit 32 -> 48: one resident block less per SM), which the exactly sized launches must not pay.
template <int M, bool STRIDE = false>
__global__ void __launch_bounds__(PT_BLOCK) k_qdeposit(PartView pv, double *__restrict__ acc1, double idr)
{
    const int npp = *pv.d_npp, lane = threadIdx.x & 31;
    extern __shared__ double dep_tiles[];
    if constexpr (STRIDE) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < npp; i += gridDim.x * blockDim.x)
            qdep_body<M>(pv, acc1, idr, npp, i, lane, dep_tiles + (threadIdx.x >> 5) * DepTile<M>::doubles);
    } else {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if ((i & ~31) >= npp) return;
        qdep_body<M>(pv, acc1, idr, npp, i, lane, dep_tiles + (threadIdx.x >> 5) * DepTile<M>::doubles);
    }
}

// particle planes of one warp-tile in registers, so a tile loop can fetch the next tile while it works on the current one
struct PartRegs { double x1, x2, p1, p2, p3, q; };
__device__ __forceinline__ PartRegs part_load(const PartView &pv, int i, int npp)
{
    PartRegs r{1.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < npp) { r.x1 = pv.x1[i]; r.x2 = pv.x2[i]; r.p1 = pv.p1[i]; r.p2 = pv.p2[i]; r.p3 = pv.p3[i]; r.q = pv.q[i]; }
    return r;
}

// ---- amjdeposit_robust: species/part2d_class.f03:746-1010 ------------------------------------------------
// STD = amjdeposit_std_part2d (:478-744): field normalisation from the stored psi (:562-601), psi left untouched
// the per-particle arithmetic of amjdeposit: gathers, the implicit Boris step, gamma / psi stores -> the factors alpha (x) beta of the
// particle's 2P x 8 contributions and its cell (key = -1: no particle).  Free of warp-level synchronisation, so that a caller can
// interleave the arithmetic of two independent tiles (the chains of one particle are long and serial: sweep.cu QPG_SWEEP_ILP2)
template <int M, bool STD = false>
__device__ __forceinline__ void amj_math(const PartView &pv, const PartRegs &pr, const double *ef, const double *bf, double qbm, double dt, double idr, int npp, int i,
                                         double (&alpha)[2 * (2 * M + 1)], double (&beta)[8], int &key)
{
    constexpr int P = 2 * M + 1;
    const bool valid = i < npp;
    key = -1;
    if (valid) {
        const double x1 = pr.x1, x2 = pr.x2;
        const double pp1 = pr.p1, pp2 = pr.p2, pp3 = pr.p3, q = pr.q;
        const Interp it = interp_info(x1, x2, idr);
        double ep[3], bp[3];
        gather3<M>(ef, it, ep);
        gather3<M>(bf, it, bp);
        const double idt = 1.0 / dt, qtmh = 0.5 * qbm * dt, rqbm = 1.0 / qbm;
        const double wp0 = ep[0] - bp[1], wp1 = ep[1] + bp[0], wp2 = ep[2];
        const double u00 = pp1 * it.c + pp2 * it.s, u01 = pp2 * it.c - pp1 * it.s, u02 = pp3;
        double gam = fast_sqrt(1.0 + u00 * u00 + u01 * u01 + u02 * u02);
        double ut0, ut1, ut2, std_q2 = 0.0, std_ipsi = 0.0;
        if constexpr (STD) {
            std_ipsi = 1.0 / (1.0 - qbm * pv.psi[i]);
            std_q2 = qtmh * std_ipsi;                               // qtmh2 = qtmh / (1 - qbm psi)
            const double q1 = std_q2 * gam;
            bp[0] *= std_q2; bp[1] *= std_q2; bp[2] *= std_q2;
            ut0 = u00 + ep[0] * q1; ut1 = u01 + ep[1] * q1; ut2 = u02 + ep[2] * q1;
        } else {
            const double qtmh1 = qtmh * gam * fast_rcp(gam - u02);
            ep[0] *= qtmh1; ep[1] *= qtmh1; ep[2] *= qtmh1;
            ut0 = u00 + ep[0]; ut1 = u01 + ep[1]; ut2 = u02 + ep[2];
            gam = fast_sqrt(1.0 + ut0 * ut0 + ut1 * ut1 + ut2 * ut2);
            const double qtmh2 = qtmh * fast_rcp(gam - ut2);
            bp[0] *= qtmh2; bp[1] *= qtmh2; bp[2] *= qtmh2;
        }
        double u0 = ut0 + ut1 * bp[2] - ut2 * bp[1];
        double u1 = ut1 + ut2 * bp[0] - ut0 * bp[2];
        double u2 = ut2 + ut0 * bp[1] - ut1 * bp[0];
        const double ostq = 2.0 * fast_rcp(1.0 + bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]);
        bp[0] *= ostq; bp[1] *= ostq; bp[2] *= ostq;
        ut0 = ut0 + u1 * bp[2] - u2 * bp[1];
        ut1 = ut1 + u2 * bp[0] - u0 * bp[2];
        ut2 = ut2 + u0 * bp[1] - u1 * bp[0];
        if constexpr (STD) {                                        // second half kick with the new gamma (:587-590)
            const double q1 = std_q2 * fast_sqrt(1.0 + ut0 * ut0 + ut1 * ut1 + ut2 * ut2);
            u0 = ut0 + ep[0] * q1; u1 = ut1 + ep[1] * q1; u2 = ut2 + ep[2] * q1;
        } else { u0 = ut0 + ep[0]; u1 = ut1 + ep[1]; u2 = ut2 + ep[2]; }
        double du0 = idt * (u0 - u00), du1 = idt * (u1 - u01);
        u0 = 0.5 * (u0 + u00); u1 = 0.5 * (u1 + u01); u2 = 0.5 * (u2 + u02);
        const double g = fast_sqrt(1.0 + u0 * u0 + u1 * u1 + u2 * u2);
        double ipsi;
        pv.gamma[i] = g;
        if constexpr (STD) ipsi = std_ipsi;                         // :607
        else {
            const double gmu = g - u2;
            ipsi = fast_rcp(gmu);
            pv.psi[i] = (1.0 - gmu) * rqbm;                  // (1 - 1/ipsi)/qbm  :864
        }
        const double dpsi = qbm * (wp2 - (wp0 * u0 + wp1 * u1) * ipsi);
        du0 = du0 + u0 * dpsi * ipsi;
        du1 = du1 + u1 * dpsi * ipsi;
        beta[0] = u0; beta[1] = u1; beta[2] = u2; beta[3] = du0; beta[4] = du1;
        beta[5] = u0 * u0 * ipsi; beta[6] = u0 * u1 * ipsi; beta[7] = u1 * u1 * ipsi;
        // phase = q*ipsi*(cos - i sin)^m ; alpha = node weight x phase
        double phr = q * ipsi, phi = 0.0;
        alpha[0] = it.w0 * phr; alpha[P] = it.w1 * phr;
#pragma unroll
        for (int m = 1; m <= M; m++) {
            double t = phr * it.c + phi * it.s;
            phi = phi * it.c - phr * it.s;
            phr = t;
            alpha[2 * m - 1] = it.w0 * phr; alpha[2 * m] = it.w0 * phi;
            alpha[P + 2 * m - 1] = it.w1 * phr; alpha[P + 2 * m] = it.w1 * phi;
        }
        key = it.idx;
    } else {
#pragma unroll
        for (int k = 0; k < 2 * P; k++) alpha[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) beta[k] = 0.0;
    }
}
template <int M, bool STD = false>
__device__ __forceinline__ void amj_core(const PartView &pv, const PartRegs &pr, const double *ef, const double *bf, double *acc8, double qbm, double dt,
                                         double idr, int npp, int i, int lane, double *tile)
{
    constexpr int P = 2 * M + 1;
    double alpha[2 * P], beta[8];
    int key;
    amj_math<M, STD>(pv, pr, ef, bf, qbm, dt, idr, npp, i, alpha, beta, key);
    warp_deposit_mma<M>(alpha, beta, key, acc8, tile, lane);
}
template <int M, bool STD = false>
__device__ __forceinline__ void amj_body(const PartView &pv, const double *ef, const double *bf, double *acc8, double qbm, double dt, double idr,
                                         int npp, int i, int lane, double *tile)
{
    amj_core<M, STD>(pv, part_load(pv, i, npp), ef, bf, acc8, qbm, dt, idr, npp, i, lane, tile);
}
template <int M, bool STD = false, bool STRIDE = false>
__global__ void __launch_bounds__(PT_BLOCK) k_amjdeposit(PartView pv, const double *__restrict__ ef, const double *__restrict__ bf,
                                                        double *__restrict__ acc8, double qbm, double dt, double idr,
                                                        const int *__restrict__ skip_flag)
{
    if (skip_flag && *skip_flag) return;
    const int npp = *pv.d_npp, lane = threadIdx.x & 31;
    extern __shared__ double dep_tiles[];
    if constexpr (STRIDE) {      // see k_qdeposit
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < npp; i += gridDim.x * blockDim.x)
            amj_body<M, STD>(pv, ef, bf, acc8, qbm, dt, idr, npp, i, lane, dep_tiles + (threadIdx.x >> 5) * DepTile<M>::doubles);
    } else {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if ((i & ~31) >= npp) return;
        amj_body<M, STD>(pv, ef, bf, acc8, qbm, dt, idr, npp, i, lane, dep_tiles + (threadIdx.x >> 5) * DepTile<M>::doubles);
    }
}

// ---- push: push_u_robust :1879-1965, push_x :2221-2262, bound test of update_bound :2323-2348 --------------
// mode bit0: push_u, bit1: push_x, bit2: flag particles with r >= edge in the bitmap, bit3: push_u is the std flavour
// acc1 != nullptr: additionally deposit the charge of the advanced, still in-bounds particle (the next slice's qdeposit
// :346-349 fused into the push; out-of-bounds particles are removed by update_bound before the reference deposits)
// the per-particle arithmetic of the push (no warp-level synchronisation: two tiles can be interleaved, sweep.cu QPG_SWEEP_ILP2);
// returns the advanced position and the out-of-bounds flag
template <int M>
__device__ __forceinline__ void push_math(const PartView &pv, const PartRegs &pr, const double *ef, const double *bf, double qbm, double dt, double idr,
                                          double edge, int mode, int npp, int i, double &xn1, double &xn2, bool &out)
{
    const bool valid = i < npp;
    out = false;
    xn1 = 0.0; xn2 = 0.0;
    if (valid) {
        double x1 = pr.x1, x2 = pr.x2;
        double p1 = pr.p1, p2 = pr.p2, p3 = pr.p3, g;
        if (mode & 1) {
            const Interp it = interp_info(x1, x2, idr);
            double ep[3], bp[3];
            gather3<M>(ef, it, ep);
            gather3<M>(bf, it, bp);
            double t = ep[0] * it.c - ep[1] * it.s; ep[1] = ep[0] * it.s + ep[1] * it.c; ep[0] = t;
            t = bp[0] * it.c - bp[1] * it.s; bp[1] = bp[0] * it.s + bp[1] * it.c; bp[0] = t;
            const double qtmh = qbm * dt * 0.5;
            double qtmh1, qtmh2;
            if (mode & 8) {                                         // push_u_std_part2d :1841-1844: stored psi and gamma
                qtmh1 = qtmh / (1.0 - qbm * pv.psi[i]);
                qtmh2 = qtmh1 * pv.gamma[i];
            } else {
                const double gam = fast_sqrt(1.0 + p1 * p1 + p2 * p2 + p3 * p3);
                qtmh1 = qtmh * fast_rcp(gam - p3);
                qtmh2 = qtmh1 * gam;
            }
            ep[0] *= qtmh2; ep[1] *= qtmh2; ep[2] *= qtmh2;
            bp[0] *= qtmh1; bp[1] *= qtmh1; bp[2] *= qtmh1;
            double ut0 = p1 + ep[0], ut1 = p2 + ep[1], ut2 = p3 + ep[2];
            p1 = ut0 + ut1 * bp[2] - ut2 * bp[1];
            p2 = ut1 + ut2 * bp[0] - ut0 * bp[2];
            p3 = ut2 + ut0 * bp[1] - ut1 * bp[0];
            const double ostq = 2.0 * fast_rcp(1.0 + bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]);
            bp[0] *= ostq; bp[1] *= ostq; bp[2] *= ostq;
            ut0 = ut0 + p2 * bp[2] - p3 * bp[1];
            ut1 = ut1 + p3 * bp[0] - p1 * bp[2];
            ut2 = ut2 + p1 * bp[1] - p2 * bp[0];
            p1 = ut0 + ep[0]; p2 = ut1 + ep[1]; p3 = ut2 + ep[2];
            g = fast_sqrt(1.0 + p1 * p1 + p2 * p2 + p3 * p3);
            pv.p1[i] = p1; pv.p2[i] = p2; pv.p3[i] = p3; pv.gamma[i] = g;
        } else {
            g = pv.gamma[i];
        }
        if (mode & 2) {
            const double dtc = dt * fast_rcp(g - p3);
            x1 = x1 + p1 * dtc;
            x2 = x2 + p2 * dtc;
            pv.x1[i] = x1; pv.x2[i] = x2;
        }
        xn1 = x1; xn2 = x2;
        if (mode & 4) {
            const double pos = __dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2)));
            out = pos >= edge;
        }
    }
}
// the warp-level part of the push: bound-flag bitmap and the fused charge deposit of the advanced, in-bounds particle
template <int M>
__device__ __forceinline__ void push_finish(const PartRegs &pr, double xn1, double xn2, bool out, double idr, int mode, unsigned *outmask, int *d_nout, double *acc1,
                                            int npp, int i, int lane, double *tile)
{
    const bool valid = i < npp;
    double qv = 0.0;
    if (mode & 4) {
        const unsigned bal = __ballot_sync(FULL, out);
        if (lane == 0) {
            outmask[i >> 5] = bal;
            if (bal) red_add(d_nout, __popc(bal));
        }
    }
    if (acc1) {
        constexpr int P = 2 * M + 1;
        double alpha[2 * P];
        int key = -1;
        if (valid && !out) { qv = pr.q; qdep_alpha<M>(xn1, xn2, qv, idr, alpha, key); }
        else {
#pragma unroll
            for (int k = 0; k < 2 * P; k++) alpha[k] = 0.0;
        }
        warp_deposit_q_mma<M>(alpha, key, acc1, tile, lane);
    }
}
template <int M>
__device__ __forceinline__ void push_core(const PartView &pv, const PartRegs &pr, const double *ef, const double *bf, double qbm, double dt, double idr,
                                          double edge, int mode, unsigned *outmask, int *d_nout, double *acc1, int npp, int i, int lane, double *tile)
{
    double xn1, xn2;
    bool out;
    push_math<M>(pv, pr, ef, bf, qbm, dt, idr, edge, mode, npp, i, xn1, xn2, out);
    push_finish<M>(pr, xn1, xn2, out, idr, mode, outmask, d_nout, acc1, npp, i, lane, tile);
}
template <int M>
__device__ __forceinline__ void push_body(const PartView &pv, const double *ef, const double *bf, double qbm, double dt, double idr, double edge,
                                          int mode, unsigned *outmask, int *d_nout, double *acc1, int npp, int i, int lane, double *tile)
{
    PartRegs pr{1.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < npp) {   // the position-only modes must not touch momenta they do not need
        pr.x1 = pv.x1[i]; pr.x2 = pv.x2[i]; pr.p1 = pv.p1[i]; pr.p2 = pv.p2[i]; pr.p3 = pv.p3[i];
        if (acc1) pr.q = pv.q[i];
    }
    push_core<M>(pv, pr, ef, bf, qbm, dt, idr, edge, mode, outmask, d_nout, acc1, npp, i, lane, tile);
}
template <int M, bool STRIDE = false>
__global__ void __launch_bounds__(PT_BLOCK) k_push(PartView pv, const double *__restrict__ ef, const double *__restrict__ bf, double qbm,
                                                  double dt, double idr, double edge, int mode, unsigned *__restrict__ outmask,
                                                  int *__restrict__ d_nout)
{
    const int npp = *pv.d_npp, lane = threadIdx.x & 31;
    if constexpr (STRIDE) {      // see k_qdeposit
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; (i & ~31) < npp; i += gridDim.x * blockDim.x)
            push_body<M>(pv, ef, bf, qbm, dt, idr, edge, mode, outmask, d_nout, nullptr, npp, i, lane, nullptr);
    } else {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if ((i & ~31) >= npp) return;
        push_body<M>(pv, ef, bf, qbm, dt, idr, edge, mode, outmask, d_nout, nullptr, npp, i, lane, nullptr);
    }
}

// ---- ponderomotive-guiding-centre flavours (laser envelope a = a_r + i a_i given on the grid) ---------------------
// amjdeposit_std_pgc :1012-1308 (STD), amjdeposit_robust_pgc :1310-1602, push_u_robust_pgc :1967-2092 (= push_u_std_pgc
// :2094-2219 arithmetically).  Same gather / deposit machinery as the plain flavours; plain IEEE division and sqrt.
struct LaserView { const double *ar, *ai, *arg, *aig; };   // a_r, a_i: dim 1; grad a_r, grad a_i: dim 3 (cylindrical)

template <int M>
__device__ __forceinline__ double gather1(const double *f, const Interp &it)
{
    constexpr int P = 2 * M + 1;
    const double *n0 = f + (size_t)it.idx * P, *n1 = n0 + P;
    double v = n0[0] * it.w0;
    v = fma(n1[0], it.w1, v);
    double phr = 1.0, phi = 0.0;
#pragma unroll
    for (int m = 1; m <= M; m++) {
        const double t = phr * it.c - phi * it.s;
        phi = phr * it.s + phi * it.c;
        phr = t;
        v = fma(n0[2 * m - 1] * (2.0 * phr) - n0[2 * m] * (2.0 * phi), it.w0, v);
        v = fma(n1[2 * m - 1] * (2.0 * phr) - n1[2 * m] * (2.0 * phi), it.w1, v);
    }
    return v;
}

// per-particle arithmetic of amjdeposit_{std,robust}_pgc (no warp-level synchronisation; shared by the stand-alone kernel and the
// persistent sweep kernel)
template <int M, bool STD>
__device__ __forceinline__ void amj_math_pgc(const PartView &pv, const double *ef, const double *bf, const LaserView &lv, double qbm, double dt, double idr, int npp, int i,
                                             double (&alpha)[2 * (2 * M + 1)], double (&beta)[8], int &key)
{
    constexpr int P = 2 * M + 1;
    key = -1;
    if (i < npp) {
        const Interp it = interp_info(pv.x1[i], pv.x2[i], idr);
        double ep[3], bp[3], agr[3], agi[3];
        gather3<M>(ef, it, ep);
        gather3<M>(bf, it, bp);
        const double apr = gather1<M>(lv.ar, it), api = gather1<M>(lv.ai, it);
        gather3<M>(lv.arg, it, agr);
        gather3<M>(lv.aig, it, agi);
        const double idt = 1.0 / dt, qtmh = 0.5 * qbm * dt;
        const double gam_corr = 0.5 * qbm * qbm * (apr * apr + api * api);                     // :1402
        const double pp1 = pv.p1[i], pp2 = pv.p2[i];
        const double u00 = pp1 * it.c + pp2 * it.s, u01 = pp2 * it.c - pp1 * it.s, u02 = pv.p3[i];
        double gam = fast_sqrt(1.0 + u00 * u00 + u01 * u01 + u02 * u02 + gam_corr);
        const double wp0 = ep[0] - bp[1], wp1 = ep[1] + bp[0], wp2 = ep[2];
        const double tmp = 0.5 * qbm * fast_rcp(gam);                                                   // ponderomotive force :1420-1423
        ep[0] -= tmp * (apr * agr[0] + api * agi[0]);
        ep[1] -= tmp * (apr * agr[1] + api * agi[1]);
        ep[2] += tmp * (apr * agr[2] + api * agi[2]);
        double qe, qb, ut0, ut1, ut2;
        if constexpr (STD) {
            qb = qtmh * fast_rcp(1.0 - qbm * pv.psi[i]);
            qe = qb * gam;
            ut0 = u00 + ep[0] * qe; ut1 = u01 + ep[1] * qe; ut2 = u02 + ep[2] * qe;
        } else {
            qe = qtmh * gam * fast_rcp(gam - u02);
            ut0 = u00 + ep[0] * qe; ut1 = u01 + ep[1] * qe; ut2 = u02 + ep[2] * qe;
            gam = fast_sqrt(1.0 + ut0 * ut0 + ut1 * ut1 + ut2 * ut2 + gam_corr);
            qb = qtmh * fast_rcp(gam - ut2);
        }
        bp[0] *= qb; bp[1] *= qb; bp[2] *= qb;
        double u0 = ut0 + ut1 * bp[2] - ut2 * bp[1];
        double u1 = ut1 + ut2 * bp[0] - ut0 * bp[2];
        double u2 = ut2 + ut0 * bp[1] - ut1 * bp[0];
        const double ostq = 2.0 * fast_rcp(1.0 + bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]);
        bp[0] *= ostq; bp[1] *= ostq; bp[2] *= ostq;
        ut0 = ut0 + u1 * bp[2] - u2 * bp[1];
        ut1 = ut1 + u2 * bp[0] - u0 * bp[2];
        ut2 = ut2 + u0 * bp[1] - u1 * bp[0];
        gam = fast_sqrt(1.0 + ut0 * ut0 + ut1 * ut1 + ut2 * ut2 + gam_corr);
        qe = STD ? qb * gam : qtmh * gam * fast_rcp(gam - ut2);                                        // second half kick re-normalised
        u0 = ut0 + ep[0] * qe; u1 = ut1 + ep[1] * qe; u2 = ut2 + ep[2] * qe;
        double du0 = idt * (u0 - u00), du1 = idt * (u1 - u01);
        u0 = 0.5 * (u0 + u00); u1 = 0.5 * (u1 + u01); u2 = 0.5 * (u2 + u02);
        const double g = fast_sqrt(1.0 + u0 * u0 + u1 * u1 + u2 * u2 + gam_corr);
        pv.gamma[i] = g;
        double ipsi;
        if constexpr (STD) ipsi = fast_rcp(1.0 - qbm * pv.psi[i]);
        else { const double gmu = g - u2; ipsi = fast_rcp(gmu); pv.psi[i] = (1.0 - gmu) * (1.0 / qbm); }
        const double dpsi = qbm * (wp2 - (wp0 * u0 + wp1 * u1) * ipsi);
        du0 = du0 + u0 * dpsi * ipsi;
        du1 = du1 + u1 * dpsi * ipsi;
        beta[0] = u0; beta[1] = u1; beta[2] = u2; beta[3] = du0; beta[4] = du1;
        beta[5] = u0 * u0 * ipsi; beta[6] = u0 * u1 * ipsi; beta[7] = u1 * u1 * ipsi;
        double phr = pv.q[i] * ipsi, phi = 0.0;
        alpha[0] = it.w0 * phr; alpha[P] = it.w1 * phr;
#pragma unroll
        for (int m = 1; m <= M; m++) {
            const double t = phr * it.c + phi * it.s;
            phi = phi * it.c - phr * it.s;
            phr = t;
            alpha[2 * m - 1] = it.w0 * phr; alpha[2 * m] = it.w0 * phi;
            alpha[P + 2 * m - 1] = it.w1 * phr; alpha[P + 2 * m] = it.w1 * phi;
        }
        key = it.idx;
    } else {
#pragma unroll
        for (int k = 0; k < 2 * P; k++) alpha[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) beta[k] = 0.0;
    }
}
template <int M, bool STD>
__global__ void __launch_bounds__(PT_BLOCK) k_amjdeposit_pgc(PartView pv, const double *__restrict__ ef, const double *__restrict__ bf, LaserView lv,
                                                            double *__restrict__ acc8, double qbm, double dt, double idr, const int *__restrict__ skip)
{
    constexpr int P = 2 * M + 1;
    if (skip && *skip) return;   // predictor-corrector loop already converged (per-slice launch path of the sim)
    const int npp = *pv.d_npp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    if ((i & ~31) >= npp) return;
    extern __shared__ double dep_tiles[];
    double alpha[2 * P], beta[8];
    int key;
    amj_math_pgc<M, STD>(pv, ef, bf, lv, qbm, dt, idr, npp, i, alpha, beta, key);
    warp_deposit_mma<M>(alpha, beta, key, acc8, dep_tiles + (threadIdx.x >> 5) * DepTile<M>::doubles, lane);
}

// per-particle arithmetic of push_u_{std,robust}_pgc :1967-2219 (shared by the stand-alone kernel and the persistent sweep kernel)
template <int M>
__device__ __forceinline__ void push_u_pgc_math(const PartView &pv, const double *ef, const double *bf, const LaserView &lv, double qbm, double dt, double idr, int i)
{
    const Interp it = interp_info(pv.x1[i], pv.x2[i], idr);
    double ep[3], bp[3], agr[3], agi[3];
    gather3<M>(ef, it, ep);
    gather3<M>(bf, it, bp);
    const double apr = gather1<M>(lv.ar, it), api = gather1<M>(lv.ai, it);
    gather3<M>(lv.arg, it, agr);
    gather3<M>(lv.aig, it, agi);
    double t = ep[0] * it.c - ep[1] * it.s; ep[1] = ep[0] * it.s + ep[1] * it.c; ep[0] = t;       // transform_to_cartesian
    t = bp[0] * it.c - bp[1] * it.s; bp[1] = bp[0] * it.s + bp[1] * it.c; bp[0] = t;
    t = agr[0] * it.c - agr[1] * it.s; agr[1] = agr[0] * it.s + agr[1] * it.c; agr[0] = t;
    t = agi[0] * it.c - agi[1] * it.s; agi[1] = agi[0] * it.s + agi[1] * it.c; agi[0] = t;
    const double qtmh = qbm * dt * 0.5, qbm2_hf = qbm * qbm * 0.5;
    double p1 = pv.p1[i], p2 = pv.p2[i], p3 = pv.p3[i];
    const double g0 = pv.gamma[i];
    double gam_corr = qbm2_hf * (apr * apr + api * api);
    const double tmp = 0.5 * qbm * fast_rcp(g0);
    ep[0] -= tmp * (apr * agr[0] + api * agi[0]);
    ep[1] -= tmp * (apr * agr[1] + api * agi[1]);
    ep[2] += tmp * (apr * agr[2] + api * agi[2]);
    const double qb = qtmh * fast_rcp(1.0 - qbm * pv.psi[i]), qe = qb * g0;
    ep[0] *= qe; ep[1] *= qe; ep[2] *= qe;
    double ut0 = p1 + ep[0], ut1 = p2 + ep[1], ut2 = p3 + ep[2];
    bp[0] *= qb; bp[1] *= qb; bp[2] *= qb;
    p1 = ut0 + ut1 * bp[2] - ut2 * bp[1];
    p2 = ut1 + ut2 * bp[0] - ut0 * bp[2];
    p3 = ut2 + ut0 * bp[1] - ut1 * bp[0];
    const double ostq = 2.0 * fast_rcp(1.0 + bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]);
    bp[0] *= ostq; bp[1] *= ostq; bp[2] *= ostq;
    ut0 = ut0 + p2 * bp[2] - p3 * bp[1];
    ut1 = ut1 + p3 * bp[0] - p1 * bp[2];
    ut2 = ut2 + p1 * bp[1] - p2 * bp[0];
    p1 = ut0 + ep[0]; p2 = ut1 + ep[1]; p3 = ut2 + ep[2];
    double tt = agr[2] * dt; gam_corr = gam_corr + qbm2_hf * (apr + 0.25 * tt) * tt;                 // :2080-2082
    tt = agi[2] * dt; gam_corr = gam_corr + qbm2_hf * (api + 0.25 * tt) * tt;
    pv.p1[i] = p1; pv.p2[i] = p2; pv.p3[i] = p3;
    pv.gamma[i] = fast_sqrt(1.0 + p1 * p1 + p2 * p2 + p3 * p3 + gam_corr);
}
template <int M>
__global__ void __launch_bounds__(PT_BLOCK) k_push_u_pgc(PartView pv, const double *__restrict__ ef, const double *__restrict__ bf, LaserView lv, double qbm,
                                                        double dt, double idr)
{
    const int npp = *pv.d_npp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npp) return;
    push_u_pgc_math<M>(pv, ef, bf, lv, qbm, dt, idr, i);
}

// ---- interp_psi: species/part2d_class.f03:2264-2305 + interp_part2d.f03:111-153 (std pushers only) ---------------
// The reference never advances `pp` inside its chunk loop (:2298-2301): of every p_cache_size = 1024 chunk only the
// FIRST particle's psi is written, with the value interpolated for the LAST particle of the chunk.  Reproduced as is
// (one thread per chunk).
template <int M>
__global__ void k_interp_psi(PartView pv, const double *__restrict__ psif, double idr)
{
    constexpr int P = 2 * M + 1;
    const int npp = *pv.d_npp;
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const long first = (long)chunk * 1024;
    if (first >= npp) return;
    const int last = (int)min((long)npp, first + 1024) - 1;
    const Interp it = interp_info(pv.x1[last], pv.x2[last], idr);
    const double *n0 = psif + (size_t)it.idx * P, *n1 = n0 + P;
    double v = n0[0] * it.w0;
    v = fma(n1[0], it.w1, v);
    double phr = 1.0, phi = 0.0;
#pragma unroll
    for (int m = 1; m <= M; m++) {
        const double t = phr * it.c - phi * it.s;
        phi = phr * it.s + phi * it.c;
        phr = t;
        v = fma(n0[2 * m - 1] * (2.0 * phr) - n0[2 * m] * (2.0 * phi), it.w0, v);
        v = fma(n1[2 * m - 1] * (2.0 * phr) - n1[2 * m] * (2.0 * phi), it.w1, v);
    }
    pv.psi[first] = v;
}

// ---- compaction (update_bound_part2d :2307-2353 / pack_particles "fill the holes inversely") ---------------
// One CTA.  K = n - nout survivors.  Head holes (flagged, index < K) are filled from tail survivors (not flagged,
// index >= K).  The sequential swap-with-last loop of the reference yields: head holes in ASCENDING order receive
// tail survivors in DESCENDING order (desc_holes = 0); pack_particles' inverse fill pairs DESCENDING holes with
// DESCENDING tail survivors (desc_holes = 1).
__device__ int block_excl_scan_int(int v, int *sm, int *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) sm[w] = incl;
    __syncthreads();
    if (w == 0) {
        int t = lane < nw ? sm[lane] : 0, ti = t;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(FULL, ti, o); if (lane >= o) ti += u; }
        sm[lane] = ti - t;
        if (lane == 31) sm[32] = ti;
    }
    __syncthreads();
    int res = sm[w] + incl - v;
    *total = sm[32];
    __syncthreads();
    return res;
}

// all threads of one CTA (any multiple of 32 up to 1024); sm = 40 ints of shared memory
__device__ void compact_body(double *const *planes, int nplanes, int *d_npp, int *d_nout, unsigned *outmask, int *lists, int desc_holes, int *sm)
{
    const int n = *d_npp, nout = *d_nout;
    if (nout == 0) return;
    const int K = n - nout, tid = threadIdx.x, nt = blockDim.x;
    const int nwords = (n + 31) >> 5;
    int *holes = lists, *surv = lists + nout;
    // words are split into contiguous per-thread ranges so ranks are ordered
    const int wpt = (nwords + nt - 1) / nt;
    const int wbeg = min(tid * wpt, nwords), wend = min(wbeg + wpt, nwords);
    // mode 0: holes = flagged indices < K (ascending), surv = unflagged indices >= K (ascending)
    // mode 1: holes = ALL flagged indices (ascending); sources are found by chasing (see below)
    int ch = 0, cs = 0;
    for (int w = wbeg; w < wend; w++) {
        unsigned bits = outmask[w];
        const int lo = w << 5;
        unsigned inrange = (lo + 32 <= n) ? FULL : ((1u << (n - lo)) - 1u);
        bits &= inrange;
        unsigned headmask = (lo + 32 <= K) ? FULL : (lo >= K ? 0u : ((1u << (K - lo)) - 1u));
        ch += __popc(desc_holes ? bits : (bits & headmask));
        cs += __popc(~bits & inrange & ~headmask);
    }
    int toth, tots;
    int oh = block_excl_scan_int(ch, sm, &toth);
    int os = block_excl_scan_int(cs, sm, &tots);
    for (int w = wbeg; w < wend; w++) {
        unsigned bits = outmask[w];
        const int lo = w << 5;
        unsigned inrange = (lo + 32 <= n) ? FULL : ((1u << (n - lo)) - 1u);
        bits &= inrange;
        unsigned headmask = (lo + 32 <= K) ? FULL : (lo >= K ? 0u : ((1u << (K - lo)) - 1u));
        unsigned hb = desc_holes ? bits : (bits & headmask), sb = ~bits & inrange & ~headmask;
        while (hb) { int b = __ffs(hb) - 1; hb &= hb - 1; holes[oh++] = lo + b; }
        if (!desc_holes) while (sb) { int b = __ffs(sb) - 1; sb &= sb - 1; surv[os++] = lo + b; }
    }
    __syncthreads();
    if (!desc_holes) {
        // sequential swap-with-last (update_bound): ascending head holes <- descending tail survivors
        for (int r = tid; r < toth; r += nt) {
            const int dst = holes[r], src = surv[tots - 1 - r];
            for (int a = 0; a < nplanes; a++) planes[a][dst] = planes[a][src];
        }
    } else {
        // "fill the holes inversely" (part3d_comm.f03:733-745): the t-th largest hole receives the CURRENT content of
        // position n-t, which may itself be a larger hole filled earlier -> chase until a surviving particle is found.
        for (int r = tid; r < nout; r += nt) {
            const int dst = holes[r];
            if (dst >= K) continue;  // dropped with the tail
            int pos = n - (nout - r);
            while (pos != dst && ((outmask[pos >> 5] >> (pos & 31)) & 1u)) {
                int lo = 0, hi = nout;  // index of pos in holes[]
                while (lo < hi) { int mid = (lo + hi) >> 1; if (holes[mid] < pos) lo = mid + 1; else hi = mid; }
                const int np2 = n - (nout - lo);
                if (np2 == pos) break;
                pos = np2;
            }
            for (int a = 0; a < nplanes; a++) planes[a][dst] = planes[a][pos];
        }
    }
    __syncthreads();
    for (int w = wbeg; w < wend; w++) outmask[w] = 0u;
    if (tid == 0) { *d_npp = K; *d_nout = 0; }
}
__global__ void __launch_bounds__(1024, 1) k_compact(double *const *planes, int nplanes, int *d_npp, int *d_nout, unsigned *outmask, int *lists,
                                                    int desc_holes, int *slice_flags)
{
    __shared__ int sm[40];
    if (slice_flags && threadIdx.x == 0) { slice_flags[3] += 1; slice_flags[4] += 1; }  // fused path: slice j is complete
    compact_body(planes, nplanes, d_npp, d_nout, outmask, lists, desc_holes, sm);
}

// ---- wire format (part2d_class.f03:2381-2388) ----------------------------------------------------------
__global__ void k_pack2d(PartView pv, double *__restrict__ buf, long cap)
{
    const int npp = *pv.d_npp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npp; i += gridDim.x * blockDim.x) {
        double *r = buf + 1 + (size_t)8 * i;
        r[0] = pv.x1[i]; r[1] = pv.x2[i]; r[2] = pv.p1[i]; r[3] = pv.p2[i]; r[4] = pv.p3[i]; r[5] = pv.gamma[i]; r[6] = pv.psi[i]; r[7] = pv.q[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) buf[0] = (double)npp;   // count first: the transport may send just the live prefix
}
__global__ void k_unpack2d(PartView pv, int *d_npp_w, const double *__restrict__ buf, long cap)
{
    const int npp = (int)min((long)buf[0], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npp; i += gridDim.x * blockDim.x) {
        const double *r = buf + 1 + (size_t)8 * i;
        pv.x1[i] = r[0]; pv.x2[i] = r[1]; pv.p1[i] = r[2]; pv.p2[i] = r[3]; pv.p3[i] = r[4]; pv.gamma[i] = r[5]; pv.psi[i] = r[6]; pv.q[i] = r[7];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *d_npp_w = npp;
}

// ---- counting sort (sort_module.f03:11-42 + part2d_class.f03:2498-2579) ---------------------------------
// Reference: ip(i) = counter(ix(i)); counter(ix(i)) -= 1 in input order, i.e. within a cell the FIRST particle
// gets the LAST slot (reverse-stable).  new_pos(i) = end(cell) - rank(i), rank = # earlier particles in that cell.
// Tiles of SORT_TILE particles: (1) per-tile histogram, (2) scan over (cell-major, tile-minor), (3) ranks + scatter.
#define SORT_TILE 1024
__global__ void __launch_bounds__(256) k_sort_hist(PartView pv, double idr, int nr, int *__restrict__ keys, int *__restrict__ hist, int ntiles)
{
    extern __shared__ int sh[];  // nr+1 counters
    const int npp = *pv.d_npp, tile = blockIdx.x;
    for (int k = threadIdx.x; k <= nr; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    const int beg = tile * SORT_TILE, end = min(beg + SORT_TILE, npp);
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const double x1 = pv.x1[i], x2 = pv.x2[i];
        double pos = __dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2))), idr);
        int key = (int)floor(pos) + 1;
        key = min(max(key, 1), nr);
        keys[i] = key;
        atomicAdd(&sh[key], 1);
    }
    __syncthreads();
    for (int k = 1 + threadIdx.x; k <= nr; k += blockDim.x) hist[(size_t)(k - 1) * ntiles + tile] = sh[k];
}
// exclusive scan of hist (length n) in place; one CTA
__global__ void __launch_bounds__(1024, 1) k_sort_scan(int *hist, long n)
{
    __shared__ int sm[40];
    const int tid = threadIdx.x, nt = blockDim.x;
    const long per = (n + nt - 1) / nt;
    const long beg = min((long)tid * per, n), end = min(beg + per, n);
    int s = 0;
    for (long k = beg; k < end; k++) s += hist[k];
    int tot;
    int off = block_excl_scan_int(s, sm, &tot);
    for (long k = beg; k < end; k++) { int v = hist[k]; hist[k] = off; off += v; }
}
// ranks inside the tile in input order, then destination = (start of next cell-tile bucket) - 1 - rank
__global__ void __launch_bounds__(256) k_sort_pos(const int *__restrict__ d_npp, int nr, const int *__restrict__ keys, const int *__restrict__ hist,
                                                 int ntiles, int *__restrict__ pos)
{
    extern __shared__ int sh[];  // running counters per key
    const int npp = *d_npp, tile = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k <= nr; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    const int beg = tile * SORT_TILE, end = min(beg + SORT_TILE, npp);
    // process the tile in order: chunks of blockDim, inside a chunk warp after warp
    for (int c0 = beg; c0 < end; c0 += blockDim.x) {
        const int i = c0 + threadIdx.x;
        const bool valid = i < end;
        const int key = valid ? keys[i] : -1 - lane;
        const unsigned same = __match_any_sync(FULL, key);
        const int before = __popc(same & ((1u << lane) - 1u));  // earlier lanes with my key
        int rank = 0;
        for (int ww = 0; ww < nw; ww++) {
            if (w == ww) {                      // the whole warp: every lane of a key group reads the count BEFORE its highest lane updates it
                const int cur = valid ? sh[key] : 0;
                __syncwarp();
                if (valid) {
                    rank = cur + before;
                    if ((same >> lane) == 1u) sh[key] = rank + 1;  // highest lane of the group publishes the new count
                }
            }
            __syncthreads();
        }
        if (valid) {
            // bucket (key, tile) spans [hist[key-1][tile], next bucket); the cell's END is the start of cell key+1, but
            // reverse-stable order over the whole cell means: pos = cell_end - 1 - global_rank, global_rank = (# in
            // earlier tiles) + rank = (hist[key-1][tile] - cell_start) + rank
            const int cell_start = hist[(size_t)(key - 1) * ntiles];
            const int cell_end = (key < nr) ? hist[(size_t)key * ntiles] : npp;
            const int grank = hist[(size_t)(key - 1) * ntiles + tile] - cell_start + rank;
            pos[i] = cell_end - 1 - grank;
        }
    }
}
__global__ void k_sort_scatter(const int *__restrict__ d_npp, const int *__restrict__ pos, const double *__restrict__ src, double *__restrict__ dst,
                               long npmax)
{
    const int npp = *d_npp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npp) return;
    const int d = pos[i];
#pragma unroll
    for (int a = 0; a < 8; a++) dst[(size_t)a * npmax + d] = src[(size_t)a * npmax + i];
}

// ------------------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------------------
static void set_planes(qpg_part2d p, double *slab)
{
    p->slab = slab;
    p->x1 = slab; p->x2 = slab + p->npmax; p->p1 = slab + 2 * p->npmax; p->p2 = slab + 3 * p->npmax; p->p3 = slab + 4 * p->npmax;
    p->gamma = slab + 5 * p->npmax; p->psi = slab + 6 * p->npmax; p->q = slab + 7 * p->npmax;
}
static double **plane_table(qpg_part2d p) { return (double **)(p->lists + 2 * p->npmax); }  // 64 ints of tail room
static double *const *part2d_plane_table(qpg_part2d p) { return plane_table(p); }

extern "C" int qpg_part2d_create(qpg_part2d *out, qpg_ctx ctx, double qbm, long npmax)
{
    ARG_TRY(out && ctx, "null arg");
    ARG_TRY(npmax >= 32 && npmax < (1L << 31) - 64, "npmax out of range");
    ARG_TRY(qbm != 0.0, "qbm must be non-zero");
    npmax = (npmax + 31) & ~31L;
    qpg_part2d p = new qpg_part2d_s();
    memset(p, 0, sizeof(*p));
    p->ctx = ctx; p->qbm = qbm; p->npmax = npmax; p->npp_hi = 0;
    double *slab;
    CUDA_TRY(cudaMalloc(&slab, sizeof(double) * 8 * npmax));
    CUDA_TRY(cudaMemsetAsync(slab, 0, sizeof(double) * 8 * npmax, ctx->stream));
    set_planes(p, slab);
    CUDA_TRY(cudaMalloc(&p->d_npp, sizeof(int) * 4));
    CUDA_TRY(cudaMemsetAsync(p->d_npp, 0, sizeof(int) * 4, ctx->stream));
    p->d_nout = p->d_npp + 1;
    CUDA_TRY(cudaMalloc(&p->outmask, sizeof(unsigned) * (npmax / 32 + 1)));
    CUDA_TRY(cudaMemsetAsync(p->outmask, 0, sizeof(unsigned) * (npmax / 32 + 1), ctx->stream));
    CUDA_TRY(cudaMalloc(&p->lists, sizeof(int) * (2 * npmax + 64)));
    {
        double *h[8] = {p->x1, p->x2, p->p1, p->p2, p->p3, p->gamma, p->psi, p->q};
        CUDA_TRY(cudaMemcpy(plane_table(p), h, sizeof(h), cudaMemcpyHostToDevice));
    }
    size_t nacc = (size_t)(ctx->nr + 2) * ctx->P;
    CUDA_TRY(cudaMalloc(&p->acc1, sizeof(double) * nacc));
    CUDA_TRY(cudaMemsetAsync(p->acc1, 0, sizeof(double) * nacc, ctx->stream));
    CUDA_TRY(cudaMalloc(&p->acc8, sizeof(double) * nacc * 8));
    CUDA_TRY(cudaMemsetAsync(p->acc8, 0, sizeof(double) * nacc * 8, ctx->stream));
    *out = p;
    return 0;
}
extern "C" int qpg_part2d_destroy(qpg_part2d p)
{
    if (!p) return 0;
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->slab); cudaFree(p->alt); cudaFree(p->snap); cudaFree(p->d_npp); cudaFree(p->outmask); cudaFree(p->lists);
    cudaFree(p->acc1); cudaFree(p->acc8); cudaFree(p->sort_keys); cudaFree(p->sort_pos); cudaFree(p->sort_hist);
    delete p;
    return 0;
}
extern "C" int qpg_part2d_upload(qpg_part2d p, const double *x, const double *pm, const double *gamma, const double *psi, const double *q, long npp)
{
    ARG_TRY(p && x && pm && gamma && psi && q, "null arg");
    ARG_TRY(npp >= 0 && npp <= p->npmax, "npp exceeds npmax");
    std::vector<double> h((size_t)5 * npp);
    for (long i = 0; i < npp; i++) {
        h[i] = x[2 * i]; h[npp + i] = x[2 * i + 1];
        h[2 * npp + i] = pm[3 * i]; h[3 * npp + i] = pm[3 * i + 1]; h[4 * npp + i] = pm[3 * i + 2];
    }
    cudaStream_t st = p->ctx->stream;
    double *dst[5] = {p->x1, p->x2, p->p1, p->p2, p->p3};
    for (int a = 0; a < 5; a++) CUDA_TRY(cudaMemcpyAsync(dst[a], h.data() + (size_t)a * npp, sizeof(double) * npp, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->gamma, gamma, sizeof(double) * npp, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->psi, psi, sizeof(double) * npp, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(p->q, q, sizeof(double) * npp, cudaMemcpyHostToDevice, st));
    int cnt[2] = {(int)npp, 0};
    CUDA_TRY(cudaMemcpyAsync(p->d_npp, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(p->outmask, 0, sizeof(unsigned) * (p->npmax / 32 + 1), st));
    CUDA_TRY(cudaStreamSynchronize(st));  // h goes out of scope
    p->npp_hi = npp;
    return 0;
}
extern "C" int qpg_part2d_npp(qpg_part2d p, long *npp)
{
    ARG_TRY(p && npp, "null arg");
    int n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, p->d_npp, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
    *npp = n;
    p->npp_hi = n;
    return 0;
}
extern "C" int qpg_part2d_download(qpg_part2d p, double *x, double *pm, double *gamma, double *psi, double *q, long *npp_out)
{
    ARG_TRY(p, "null arg");
    long npp;
    int rc = qpg_part2d_npp(p, &npp);
    if (rc) return rc;
    if (npp_out) *npp_out = npp;
    cudaStream_t st = p->ctx->stream;
    std::vector<double> h((size_t)5 * npp);
    if (x || pm) {
        double *src[5] = {p->x1, p->x2, p->p1, p->p2, p->p3};
        for (int a = 0; a < 5; a++) CUDA_TRY(cudaMemcpyAsync(h.data() + (size_t)a * npp, src[a], sizeof(double) * npp, cudaMemcpyDeviceToHost, st));
    }
    if (gamma) CUDA_TRY(cudaMemcpyAsync(gamma, p->gamma, sizeof(double) * npp, cudaMemcpyDeviceToHost, st));
    if (psi) CUDA_TRY(cudaMemcpyAsync(psi, p->psi, sizeof(double) * npp, cudaMemcpyDeviceToHost, st));
    if (q) CUDA_TRY(cudaMemcpyAsync(q, p->q, sizeof(double) * npp, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (long i = 0; i < npp; i++) {
        if (x) { x[2 * i] = h[i]; x[2 * i + 1] = h[npp + i]; }
        if (pm) { pm[3 * i] = h[2 * npp + i]; pm[3 * i + 1] = h[3 * npp + i]; pm[3 * i + 2] = h[4 * npp + i]; }
    }
    return 0;
}
extern "C" int qpg_part2d_snapshot(qpg_part2d p)
{
    ARG_TRY(p, "null arg");
    long npp;
    int rc = qpg_part2d_npp(p, &npp);
    if (rc) return rc;
    if (!p->snap) CUDA_TRY(cudaMalloc(&p->snap, sizeof(double) * 8 * p->npmax));
    CUDA_TRY(cudaMemcpyAsync(p->snap, p->slab, sizeof(double) * 8 * p->npmax, cudaMemcpyDeviceToDevice, p->ctx->stream));
    p->snap_np = npp;
    return 0;
}
extern "C" int qpg_part2d_renew(qpg_part2d p)
{
    ARG_TRY(p, "null arg");
    if (!p->snap) { qpg_set_error("qpg_part2d_renew: no snapshot taken"); return QPG_ERR_STATE; }
    cudaStream_t st = p->ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(p->slab, p->snap, sizeof(double) * 8 * p->npmax, cudaMemcpyDeviceToDevice, st));
    int cnt[2] = {(int)p->snap_np, 0};
    CUDA_TRY(cudaMemcpyAsync(p->d_npp, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(p->outmask, 0, sizeof(unsigned) * (p->npmax / 32 + 1), st));
    p->npp_hi = p->snap_np;
    return 0;
}

#define DISPATCH_M(M, FN, ...)                 \
    switch (M) {                               \
    case 0: FN<0>(__VA_ARGS__); break;         \
    case 1: FN<1>(__VA_ARGS__); break;         \
    case 2: FN<2>(__VA_ARGS__); break;         \
    case 3: FN<3>(__VA_ARGS__); break;         \
    default: FN<4>(__VA_ARGS__); break;        \
    }
template <int M> static void l_qdeposit(int grid, cudaStream_t st, PartView pv, double *acc1, double idr)
{
    constexpr size_t smem = sizeof(double) * DepTile<M>::doubles * (PT_BLOCK / 32);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_qdeposit<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_qdeposit<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (grid < 0) k_qdeposit<M, true><<<-grid, PT_BLOCK, smem, st>>>(pv, acc1, idr);      // pt_grid: a negative grid = the capped, striding launch
    else k_qdeposit<M, false><<<grid, PT_BLOCK, smem, st>>>(pv, acc1, idr);
}
// development aid (tools/occupancy_probe.py): QPG_DEV_EXTRA_SMEM=<bytes> pads the dynamic shared memory of the amjdeposit launches to
// cap the resident blocks per SM
static size_t dev_extra_smem() { static long v = -1; if (v < 0) { const char *e = getenv("QPG_DEV_EXTRA_SMEM"); v = e ? atol(e) : 0; } return (size_t)v; }
template <int M> static void l_amjdeposit(int grid, cudaStream_t st, PartView pv, const double *ef, const double *bf, double *acc8, double qbm, double dt, double idr, const int *skip, int std_flavour)
{
    const size_t smem = sizeof(double) * DepTile<M>::doubles * (PT_BLOCK / 32) + dev_extra_smem();
    static bool attr_set = false;   // > 48 KB of dynamic shared memory needs the opt-in (M >= 3)
    if (!attr_set) {
        cudaFuncSetAttribute(k_amjdeposit<M, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_amjdeposit<M, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_amjdeposit<M, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_amjdeposit<M, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (grid < 0) {
        if (std_flavour) k_amjdeposit<M, true, true><<<-grid, PT_BLOCK, smem, st>>>(pv, ef, bf, acc8, qbm, dt, idr, skip);
        else k_amjdeposit<M, false, true><<<-grid, PT_BLOCK, smem, st>>>(pv, ef, bf, acc8, qbm, dt, idr, skip);
    } else if (std_flavour) k_amjdeposit<M, true, false><<<grid, PT_BLOCK, smem, st>>>(pv, ef, bf, acc8, qbm, dt, idr, skip);
    else k_amjdeposit<M, false, false><<<grid, PT_BLOCK, smem, st>>>(pv, ef, bf, acc8, qbm, dt, idr, skip);
}
template <int M> static void l_interp_psi(int grid, cudaStream_t st, PartView pv, const double *psif, double idr)
{ k_interp_psi<M><<<grid, 128, 0, st>>>(pv, psif, idr); }
template <int M> static void l_push(int grid, cudaStream_t st, PartView pv, const double *ef, const double *bf, double qbm, double dt, double idr, double edge, int mode, unsigned *outmask, int *d_nout)
{
    if (grid < 0) k_push<M, true><<<-grid, PT_BLOCK, 0, st>>>(pv, ef, bf, qbm, dt, idr, edge, mode, outmask, d_nout);
    else k_push<M, false><<<grid, PT_BLOCK, 0, st>>>(pv, ef, bf, qbm, dt, idr, edge, mode, outmask, d_nout);
}

// grid of the tile kernels above: one block per PT_BLOCK particles of the host's upper bound npp_hi.  When that bound is the capacity itself --
// the state "live count known to the device only" after a hand-off or a neutral's update -- the grid is capped at PT_GRID_CAP blocks and the
// kernels stride: the bound is then typically far above the live count (config 5: 3 M slots, 1e5 electrons) and thousands of empty blocks
// per launch cost more than the work.  A set whose count the host knows keeps its exact grid (measured on 4 M particles streaming from HBM:
// the capped, striding grid is 10 % slower in the deposits, 6 % faster in the push).
#define PT_GRID_CAP (148 * 8)
static inline int pt_grid(qpg_part2d p)
{
    const long g = (p->npp_hi + PT_BLOCK - 1) / PT_BLOCK;
    return (int)((p->npp_hi >= p->npmax && g > PT_GRID_CAP) ? -PT_GRID_CAP : g);      // negative: the launchers pick the striding instantiation
}
int part2d_launch_qdeposit(qpg_part2d p)
{
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    const int grid = pt_grid(p);
    PartView pv = view_of(p);
    TprofScope tp(c, TP_K_QDEP);
    DISPATCH_M(c->M, l_qdeposit, grid, c->stream, pv, p->acc1, 1.0 / c->dr);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int part2d_launch_amjdeposit(qpg_part2d p, qpg_field ef, qpg_field bf, double dt, const int *skip_flag, int std_flavour)
{
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    const int grid = pt_grid(p);
    PartView pv = view_of(p);
    TprofScope tp(c, TP_K_AMJ);
    DISPATCH_M(c->M, l_amjdeposit, grid, c->stream, pv, ef->f1, bf->f1, p->acc8, p->qbm, dt, 1.0 / c->dr, skip_flag, std_flavour);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int part2d_launch_push(qpg_part2d p, qpg_field ef, qpg_field bf, double dt, int mode)
{
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    const int grid = pt_grid(p);
    PartView pv = view_of(p);
    TprofScope tp(c, TP_K_PUSH);
    const double edge = (double)c->nr * c->dr;
    const double *e1 = ef ? ef->f1 : nullptr, *b1 = bf ? bf->f1 : nullptr;
    DISPATCH_M(c->M, l_push, grid, c->stream, pv, e1, b1, p->qbm, dt, 1.0 / c->dr, edge, mode, p->outmask, p->d_nout);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int part2d_launch_compact(qpg_part2d p, int *slice_flags)
{
    qpg_ctx c = p->ctx;
    double **tbl = plane_table(p);
    TprofScope tp(c, TP_K_COMPACT);
    k_compact<<<1, 1024, 0, c->stream>>>(tbl, 8, p->d_npp, p->d_nout, p->outmask, p->lists, 0, slice_flags);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int part2d_epilogue_q(qpg_part2d p, qpg_field q)
{
    FProgBuilder pb(p->ctx);
    FOp &o = pb.add(FOP_QFIX); o.a = p->acc1; o.b = q->f1; o.da = 1;
    return pb.launch(TP_DEPOSIT2D);
}

extern "C" int qpg_part2d_qdeposit(qpg_part2d p, qpg_field q)
{
    ARG_TRY(p && q && q->dim == 1 && q->ctx == p->ctx, "bad field handle");
    int rc = part2d_launch_qdeposit(p);
    if (rc) return rc;
    return part2d_epilogue_q(p, q);
}
extern "C" int qpg_part2d_amjdeposit(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, qpg_field cu, qpg_field amu, qpg_field dcu, double dt)
{
    ARG_TRY(p && ef && bf && cu && amu && dcu, "null arg");
    ARG_TRY(ef->dim == 3 && bf->dim == 3 && cu->dim == 3 && amu->dim == 3 && dcu->dim == 2, "field dims must be e3 b3 cu3 amu3 dcu2");
    if (push_type != QPG_PUSH2_ROBUST && push_type != QPG_PUSH2_STD) { qpg_set_error("push_type must be std (0) or robust (1); the pgc flavours need the laser path (SURVEY.md §8f)"); return QPG_ERR_UNSUPPORTED; }
    int rc = part2d_launch_amjdeposit(p, ef, bf, dt, nullptr, push_type == QPG_PUSH2_STD);
    if (rc) return rc;
    FProgBuilder pb(p->ctx);
    FOp &o = pb.add(FOP_AMJFIX); o.a = p->acc8; o.b = cu->f1; o.c = dcu->f1; o.d = amu->f1; o.da = 1;
    return pb.launch(TP_DEPOSIT2D);
}
extern "C" int qpg_part2d_push_u(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, double dt)
{
    ARG_TRY(p && ef && bf && ef->dim == 3 && bf->dim == 3, "bad field handles");
    if (push_type != QPG_PUSH2_ROBUST && push_type != QPG_PUSH2_STD) { qpg_set_error("push_type must be std (0) or robust (1)"); return QPG_ERR_UNSUPPORTED; }
    return part2d_launch_push(p, ef, bf, dt, push_type == QPG_PUSH2_STD ? 1 | 8 : 1);
}
extern "C" int qpg_part2d_interp_psi(qpg_part2d p, qpg_field psi)
{
    ARG_TRY(p && psi && psi->dim == 1 && psi->ctx == p->ctx, "bad field handle");
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    const int nchunks = (int)((p->npp_hi + 1023) / 1024), grid = (nchunks + 127) / 128;
    TprofScope tp(c, TP_PUSH2D);
    DISPATCH_M(c->M, l_interp_psi, grid, c->stream, view_of(p), psi->f1, 1.0 / c->dr);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
template <int M> static void l_amj_pgc(int grid, cudaStream_t st, PartView pv, const double *ef, const double *bf, LaserView lv, double *acc8, double qbm, double dt, double idr, int std_flavour, const int *skip = nullptr)
{
    constexpr size_t smem = sizeof(double) * DepTile<M>::doubles * (PT_BLOCK / 32);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_amjdeposit_pgc<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_amjdeposit_pgc<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (std_flavour) k_amjdeposit_pgc<M, true><<<grid, PT_BLOCK, smem, st>>>(pv, ef, bf, lv, acc8, qbm, dt, idr, skip);
    else k_amjdeposit_pgc<M, false><<<grid, PT_BLOCK, smem, st>>>(pv, ef, bf, lv, acc8, qbm, dt, idr, skip);
}
template <int M> static void l_push_pgc(int grid, cudaStream_t st, PartView pv, const double *ef, const double *bf, LaserView lv, double qbm, double dt, double idr)
{ k_push_u_pgc<M><<<grid, PT_BLOCK, 0, st>>>(pv, ef, bf, lv, qbm, dt, idr); }

static int laser_view(qpg_part2d p, qpg_field ar, qpg_field ai, qpg_field arg, qpg_field aig, LaserView *lv)
{
    ARG_TRY(ar && ai && arg && aig, "null laser field");
    ARG_TRY(ar->dim == 1 && ai->dim == 1 && arg->dim == 3 && aig->dim == 3, "laser fields must be a_r(1) a_i(1) grad a_r(3) grad a_i(3)");
    ARG_TRY(ar->ctx == p->ctx && ai->ctx == p->ctx && arg->ctx == p->ctx && aig->ctx == p->ctx, "laser fields belong to another context");
    lv->ar = ar->f1; lv->ai = ai->f1; lv->arg = arg->f1; lv->aig = aig->f1;
    return 0;
}
// raw variants for the sim's per-slice path: sums stay in acc8 for program C, the push is push_u only
int part2d_launch_amjdeposit_pgc(qpg_part2d p, qpg_field ef, qpg_field bf, qpg_field ar, qpg_field ai, qpg_field arg, qpg_field aig, double dt, const int *skip_flag, int std_flavour)
{
    if (p->npp_hi == 0) return 0;
    LaserView lv;
    int rc = laser_view(p, ar, ai, arg, aig, &lv);
    if (rc) return rc;
    qpg_ctx c = p->ctx;
    const int grid = (int)((p->npp_hi + PT_BLOCK - 1) / PT_BLOCK);
    TprofScope tp(c, TP_K_AMJ);
    DISPATCH_M(c->M, l_amj_pgc, grid, c->stream, view_of(p), ef->f1, bf->f1, lv, p->acc8, p->qbm, dt, 1.0 / c->dr, std_flavour, skip_flag);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part2d_amjdeposit_pgc(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, qpg_field ar, qpg_field ai, qpg_field ar_grad,
                                         qpg_field ai_grad, qpg_field cu, qpg_field amu, qpg_field dcu, double dt)
{
    ARG_TRY(p && ef && bf && cu && amu && dcu, "null arg");
    ARG_TRY(ef->dim == 3 && bf->dim == 3 && cu->dim == 3 && amu->dim == 3 && dcu->dim == 2, "field dims must be e3 b3 cu3 amu3 dcu2");
    ARG_TRY(push_type == QPG_PUSH2_STD_PGC || push_type == QPG_PUSH2_ROBUST_PGC, "push_type must be std_pgc (4) or robust_pgc (5)");
    LaserView lv;
    int rc = laser_view(p, ar, ai, ar_grad, ai_grad, &lv);
    if (rc) return rc;
    qpg_ctx c = p->ctx;
    if (p->npp_hi > 0) {
        const int grid = (int)((p->npp_hi + PT_BLOCK - 1) / PT_BLOCK);
        TprofScope tp(c, TP_K_AMJ);
        DISPATCH_M(c->M, l_amj_pgc, grid, c->stream, view_of(p), ef->f1, bf->f1, lv, p->acc8, p->qbm, dt, 1.0 / c->dr, push_type == QPG_PUSH2_STD_PGC);
        count_launch(c);
        CUDA_TRY(cudaGetLastError());
    }
    FProgBuilder pb(c);
    FOp &o = pb.add(FOP_AMJFIX); o.a = p->acc8; o.b = cu->f1; o.c = dcu->f1; o.d = amu->f1; o.da = 1;
    return pb.launch(TP_DEPOSIT2D);
}
extern "C" int qpg_part2d_push_u_pgc(qpg_part2d p, int push_type, qpg_field ef, qpg_field bf, qpg_field ar, qpg_field ai, qpg_field ar_grad,
                                     qpg_field ai_grad, double dt)
{
    ARG_TRY(p && ef && bf && ef->dim == 3 && bf->dim == 3, "bad field handles");
    ARG_TRY(push_type == QPG_PUSH2_STD_PGC || push_type == QPG_PUSH2_ROBUST_PGC, "push_type must be std_pgc (4) or robust_pgc (5)");
    LaserView lv;
    int rc = laser_view(p, ar, ai, ar_grad, ai_grad, &lv);
    if (rc) return rc;
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    const int grid = (int)((p->npp_hi + PT_BLOCK - 1) / PT_BLOCK);
    TprofScope tp(c, TP_K_PUSH);
    DISPATCH_M(c->M, l_push_pgc, grid, c->stream, view_of(p), ef->f1, bf->f1, lv, p->qbm, dt, 1.0 / c->dr);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part2d_push_x(qpg_part2d p, double dt)
{
    ARG_TRY(p, "null arg");
    return part2d_launch_push(p, nullptr, nullptr, dt, 2);
}
extern "C" int qpg_part2d_update_bound(qpg_part2d p)
{
    ARG_TRY(p, "null arg");
    int rc = part2d_launch_push(p, nullptr, nullptr, 0.0, 4);
    if (rc) return rc;
    return part2d_launch_compact(p, nullptr);
}

// species/part2d_comm.f03:147 move_part2d_comm: the radial relay between the ranks of one stage.  One GPU owns the whole
// radial extent (nodes(1) = 1), so after update_bound there is nobody to hand particles to: kept as an entry point so
// the species2d%push_x call sequence (push_x -> update_bound -> move) maps one to one.
extern "C" int qpg_part2d_move(qpg_part2d p)
{
    ARG_TRY(p, "null arg");
    return 0;
}

extern "C" long qpg_part2d_wire_count(qpg_part2d p) { return p ? 8 * p->npmax + 1 : -1; }
extern "C" int qpg_part2d_pack(qpg_part2d p, double *dev_buf)
{
    ARG_TRY(p && dev_buf, "null arg");
    TprofScope tp(p->ctx, TP_PIPELINE);
    k_pack2d<<<296, 256, 0, p->ctx->stream>>>(view_of(p), dev_buf, p->npmax);
    count_launch(p->ctx);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part2d_unpack(qpg_part2d p, const double *dev_buf)
{
    ARG_TRY(p && dev_buf, "null arg");
    TprofScope tp(p->ctx, TP_PIPELINE);
    k_unpack2d<<<296, 256, 0, p->ctx->stream>>>(view_of(p), p->d_npp, dev_buf, p->npmax);
    count_launch(p->ctx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemsetAsync(p->d_nout, 0, sizeof(int), p->ctx->stream));
    p->npp_hi = p->npmax;  // unknown until the next sync; kernels bound themselves by the device count
    return 0;
}

static int sort_prepare(qpg_part2d p)
{
    qpg_ctx c = p->ctx;
    long ntiles = (p->npmax + SORT_TILE - 1) / SORT_TILE;
    if (!p->sort_keys) {
        CUDA_TRY(cudaMalloc(&p->sort_keys, sizeof(int) * p->npmax));
        CUDA_TRY(cudaMalloc(&p->sort_pos, sizeof(int) * p->npmax));
        CUDA_TRY(cudaMalloc(&p->sort_hist, sizeof(int) * ((size_t)c->nr * ntiles + 1)));
        p->sort_tiles = ntiles;
    }
    return 0;
}
static int sort_index_device(qpg_part2d p, int *ntiles_out)
{
    qpg_ctx c = p->ctx;
    int rc = sort_prepare(p);
    if (rc) return rc;
    int ntiles = (int)((p->npp_hi + SORT_TILE - 1) / SORT_TILE);
    if (ntiles < 1) ntiles = 1;
    *ntiles_out = ntiles;
    PartView pv = view_of(p);
    size_t sh = sizeof(int) * (c->nr + 1);
    CUDA_TRY(cudaMemsetAsync(p->sort_hist, 0, sizeof(int) * ((size_t)c->nr * ntiles + 1), c->stream));
    k_sort_hist<<<ntiles, 256, sh, c->stream>>>(pv, 1.0 / c->dr, c->nr, p->sort_keys, p->sort_hist, ntiles);
    k_sort_scan<<<1, 1024, 0, c->stream>>>(p->sort_hist, (long)c->nr * ntiles);
    k_sort_pos<<<ntiles, 256, sh, c->stream>>>(p->d_npp, c->nr, p->sort_keys, p->sort_hist, ntiles, p->sort_pos);
    count_launch(c, 3);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part2d_sort(qpg_part2d p)
{
    ARG_TRY(p, "null arg");
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    ARG_TRY(sizeof(int) * (c->nr + 1) <= 48 * 1024, "nr too large for the sort histogram");
    TprofScope tp(c, TP_SORT2D);
    int ntiles;
    int rc = sort_index_device(p, &ntiles);
    if (rc) return rc;
    if (!p->alt) CUDA_TRY(cudaMalloc(&p->alt, sizeof(double) * 8 * p->npmax));
    const int grid = (int)((p->npp_hi + 255) / 256);
    k_sort_scatter<<<grid, 256, 0, c->stream>>>(p->d_npp, p->sort_pos, p->slab, p->alt, p->npmax);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    // pointers stay fixed (CUDA-graph friendly): copy the sorted planes back
    CUDA_TRY(cudaMemcpyAsync(p->slab, p->alt, sizeof(double) * 8 * p->npmax, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
extern "C" int qpg_part2d_sort_index(qpg_part2d p, int *host_ix, int *host_ip)
{
    ARG_TRY(p && host_ix && host_ip, "null arg");
    long npp;
    int rc = qpg_part2d_npp(p, &npp);
    if (rc) return rc;
    if (npp == 0) return 0;
    int ntiles;
    rc = sort_index_device(p, &ntiles);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(host_ix, p->sort_keys, sizeof(int) * npp, cudaMemcpyDeviceToHost, p->ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(host_ip, p->sort_pos, sizeof(int) * npp, cudaMemcpyDeviceToHost, p->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
    for (long i = 0; i < npp; i++) host_ip[i] += 1;  // 1-based like generate_sort_idx_1d
    return 0;
}

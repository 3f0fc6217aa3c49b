// fields.cu -- context, field storage and the single-CTA "field program" kernel.
//
// All per-slice field work of QPAD (fields/field_{psi,e,b,src}_class.f03, fields/field_class.f03 arithmetic,
// simulation_class.f03:522 convergence_tester) is O(nr * planes) and strictly sequential between stages, so it is
// latency- not bandwidth-bound.  It runs as ONE CTA of 1024 threads that executes a short program of field ops
// back to back (one launch per program, __syncthreads between ops, data stays in L1/L2).
//
// Tridiagonal solves: the reference calls HYPRE cyclic reduction (field_solver_class.f03:172-177).  The operators
// are constant, so at context creation we factor each inverse in its semiseparable (Green's function) form
//     x_i = p_i * sum_{j<=i} q_j d_j + u_i * sum_{j>i} v_j d_j
// (u, w = homogeneous solutions meeting the inner / outer boundary row, computed in long double) and a solve is
// two scans.  One warp owns one system: each lane scans a contiguous chunk serially, the 32 lane totals are
// combined with warp shuffles.  4*nr coefficients per operator instead of 2*nr*log2(nr) for pre-factored PCR, and
// the result is accurate to ~1e-15 (Thomas/PCR sit at 1e-12..1e-11 for the m=0 operators at nr=1024).
#include "common.cuh"
#include <cmath>
#include <cstdarg>

// ------------------------------------------------------------------------------------------------
// error handling / tprof
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void qpg_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int qpg_cuda_fail(cudaError_t e, const char *what)
{
    qpg_set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return QPG_ERR_CUDA;
}
extern "C" const char *qpg_last_error(void) { return g_err; }
extern "C" int qpg_version(void) { return 100; }

const char *const qpg_tprof_names[TP_COUNT] = {
    "deposit 2D particles", "push 2D particles", "move 2D particles", "sort 2D particles", "solve psi", "solve bz",
    "solve ez", "solve plasma bt", "solve beam bt", "solve plasma et", "solve beam et", "set source", "arithmetics",
    "pipeline", "deposit 3D particles", "push 3D particles", "move 3D particles", "fused field program",
    "kernel qdeposit", "kernel amjdeposit", "kernel push", "kernel compact", "kernel sweep"};

TprofScope::TprofScope(qpg_ctx c, int e) : ctx(c), ev(e), a(nullptr), b(nullptr), on(false)
{
    if (!c->tprof_on || c->capturing) return;
    auto get = [&]() {
        cudaEvent_t x;
        if (!c->ev_pool.empty()) { x = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(&x);
        return x;
    };
    a = get(); b = get();
    cudaEventRecord(a, c->stream);
    on = true;
}
TprofScope::~TprofScope()
{
    if (!on) return;
    cudaEventRecord(b, ctx->stream);
    ctx->tp_pending.push_back({ev, {a, b}});
    if (ctx->tp_pending.size() > 4096) qpg_tprof_flush(ctx);
}
int qpg_tprof_flush(qpg_ctx ctx)
{
    if (ctx->tp_pending.empty()) return 0;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto &pe : ctx->tp_pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pe.second.first, pe.second.second);
        ctx->tp_ms[pe.first] += ms;
        ctx->tp_calls[pe.first] += 1;
        ctx->ev_pool.push_back(pe.second.first);
        ctx->ev_pool.push_back(pe.second.second);
    }
    ctx->tp_pending.clear();
    return 0;
}
extern "C" int qpg_tprof_enable(qpg_ctx ctx, int on) { ARG_TRY(ctx, "null ctx"); qpg_tprof_flush(ctx); ctx->tprof_on = on != 0; return 0; }
extern "C" int qpg_tprof_reset(qpg_ctx ctx)
{
    ARG_TRY(ctx, "null ctx");
    qpg_tprof_flush(ctx);
    for (int i = 0; i < TP_COUNT; i++) { ctx->tp_ms[i] = 0; ctx->tp_calls[i] = 0; }
    return 0;
}
extern "C" int qpg_tprof_get(qpg_ctx ctx, const char *event, double *ms_total, long *ncalls)
{
    ARG_TRY(ctx && event, "null arg");
    int rc = qpg_tprof_flush(ctx);
    if (rc) return rc;
    for (int i = 0; i < TP_COUNT; i++)
        if (!strcmp(event, qpg_tprof_names[i])) {
            if (ms_total) *ms_total = ctx->tp_ms[i];
            if (ncalls) *ncalls = ctx->tp_calls[i];
            return 0;
        }
    qpg_set_error("unknown tprof event '%s'", event);
    return QPG_ERR_ARG;
}
extern "C" long qpg_launch_count(qpg_ctx ctx) { return ctx ? ctx->launches : -1; }

#define NT_FIELD 1024
// ------------------------------------------------------------------------------------------------
// operator construction (fields/field_solver_class.f03:256-561 set_struct_matrix, single r-owner) and its
// semiseparable factorisation
// ------------------------------------------------------------------------------------------------
static void build_matrix(int kind, int m, int nr, double dr, int bnd, double relax, std::vector<long double> &a,
                         std::vector<long double> &b, std::vector<long double> &c)
{
    a.assign(nr, 0.0L); b.assign(nr, 0.0L); c.assign(nr, 0.0L);
    // The reference evaluates the entries in fp64; reproduce the fp64 values exactly, then promote.
    std::vector<double> A(nr), B(nr), Cc(nr);
    double m2 = (double)(m * m);
    double j = 0.0;
    for (int i = 1; i < nr; i++) {
        j = j + 1.0;
        A[i] = 1.0 - 0.5 / j;
        Cc[i] = 1.0 + 0.5 / j;
        if (kind == FK_BPLUS) { double k = (double)(m + 1) / j; B[i] = -2.0 - k * k - relax; }
        else if (kind == FK_BMINUS) { double k = (double)(m - 1) / j; B[i] = -2.0 - k * k - relax; }
        else B[i] = -2.0 - m2 / (j * j);
    }
    bool coupled = (kind == FK_BPLUS) ? false : (kind == FK_BMINUS ? (m == 1) : (m == 0));
    if (coupled) { A[0] = 0.0; B[0] = (kind == FK_BMINUS) ? -4.0 - relax : -4.0; Cc[0] = 4.0; }
    else { A[0] = 0.0; B[0] = 1.0; Cc[0] = 0.0; A[1] = 0.0; }
    if (bnd == QPG_BND_ZERO) Cc[nr - 1] = 0.0;
    else {
        double jmax = (double)nr, fold;
        if (kind == FK_PSI || kind == FK_EZ || kind == FK_BZ) fold = (m == 0) ? 0.0 : 1.0 - (double)m / jmax;
        else if (kind == FK_BT) fold = (m == 0) ? 1.0 + 1.0 / (jmax * log(jmax * dr)) : 1.0 - (double)m / jmax;
        else fold = 1.0 - (double)(m + 1) / jmax;  // b+ and b- both, field_solver_class.f03:532-540
        if (!((kind == FK_PSI || kind == FK_EZ || kind == FK_BZ) && m == 0)) B[nr - 1] = B[nr - 1] + fold * Cc[nr - 1];
        Cc[nr - 1] = 0.0;
    }
    double dr2 = dr * dr;
    for (int i = 0; i < nr; i++) { a[i] = A[i] / dr2; b[i] = B[i] / dr2; c[i] = Cc[i] / dr2; }
}

static void factor_operator(int kind, int m, int nr, double dr, int bnd, double relax, int C, double *hq, double *hv,
                            double *hp, double *hu, double *axis_inv)
{
    std::vector<long double> a, b, c;
    build_matrix(kind, m, nr, dr, bnd, relax, a, b, c);
    int off = (c[0] == 0.0L && a[1] == 0.0L) ? 1 : 0;
    *axis_inv = off ? (double)(1.0L / b[0]) : 0.0;
    std::vector<long double> u(nr, 0.0L), w(nr, 0.0L), D(nr, 1.0L);
    u[off] = 1.0L;
    u[off + 1] = -b[off] * u[off] / c[off];
    for (int i = off + 1; i <= nr - 2; i++) u[i + 1] = -(a[i] * u[i - 1] + b[i] * u[i]) / c[i];
    w[nr - 1] = 1.0L;
    w[nr - 2] = -b[nr - 1] * w[nr - 1] / a[nr - 1];
    for (int i = nr - 2; i >= off + 1; i--) w[i - 1] = -(b[i] * w[i] + c[i] * w[i + 1]) / a[i];
    for (int jn = off; jn < nr; jn++) {
        long double t = b[jn] * u[jn] * w[jn];
        if (jn > off) t += a[jn] * u[jn - 1] * w[jn];
        if (jn < nr - 1) t += c[jn] * w[jn + 1] * u[jn];
        D[jn] = t;
    }
    int len = NT_FIELD * C;  // C = nodes per thread (IPT); natural node order, zero padded
    for (int i = 0; i < len; i++) hq[i] = hv[i] = hp[i] = hu[i] = 0.0;
    for (int t = off; t < nr; t++) {
        int pos = t;
        hp[pos] = (double)w[t];
        hq[pos] = (double)(u[t] / D[t]);
        hu[pos] = (double)u[t];
        hv[pos] = (double)(w[t] / D[t]);
    }
}

struct CtxDev { OpCoef ops[FK_NKIND][QPG_MAX_MODE + 1]; };
static std::map<qpg_ctx, CtxDev *> g_devops;
OpCoef *qpg_ctx_dev_ops(qpg_ctx ctx) { return &g_devops[ctx]->ops[0][0]; }

__global__ void k_field_prog(const __grid_constant__ FProg pg);

extern "C" int qpg_ctx_create(qpg_ctx *out, int device, void *cuda_stream, int nr, int max_mode, double dr, double dxi,
                              int field_boundary, double relax_fac)
{
    ARG_TRY(out, "null out");
    ARG_TRY(nr >= 8 && nr <= 65536, "nr out of range [8, 65536]");
    ARG_TRY(max_mode >= 0 && max_mode <= QPG_MAX_MODE, "max_mode out of range");
    ARG_TRY(field_boundary == QPG_BND_ZERO || field_boundary == QPG_BND_OPEN, "field_boundary must be zero(2) or open(3)");
    ARG_TRY(dr > 0 && dxi > 0, "dr, dxi must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        qpg_set_error("no CUDA device available (%s): libqpadb200 has no CPU fallback", cudaGetErrorString(e));
        return QPG_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    qpg_ctx c = new qpg_ctx_s();
    c->device = device;
    c->own_stream = (cuda_stream == nullptr);
    if (c->own_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    else c->stream = (cudaStream_t)cuda_stream;
    c->nr = nr; c->M = max_mode; c->P = 2 * max_mode + 1;
    c->dr = dr; c->dxi = dxi; c->bnd = field_boundary;
    c->relax = relax_fac >= 0.0 ? relax_fac : 1.0e-3 * ((dr / 0.02) * (dr / 0.02));
    int need = (nr + NT_FIELD - 1) / NT_FIELD, C = 1, logC = 0;  // nodes per thread of the field CTA (1, 2 or 4)
    while (C < need) { C <<= 1; logC++; }
    if (C > 4) { qpg_set_error("nr=%d exceeds the 4096 radial nodes the field kernel supports", nr); return QPG_ERR_UNSUPPORTED; }
    c->C = C; c->logC = logC;
    c->launches = 0; c->tprof_on = false; c->capturing = false;
    for (int i = 0; i < TP_COUNT; i++) { c->tp_ms[i] = 0; c->tp_calls[i] = 0; }
    // coefficient pool
    size_t len = (size_t)NT_FIELD * C, nops = (size_t)FK_NKIND * (max_mode + 1);
    std::vector<double> host(nops * 4 * len);
    CUDA_TRY(cudaMalloc(&c->coef_pool, host.size() * sizeof(double)));
    CtxDev hostdev;
    memset(&hostdev, 0, sizeof(hostdev));
    size_t o = 0;
    for (int kind = 0; kind < FK_NKIND; kind++)
        for (int m = 0; m <= max_mode; m++) {
            double ax;
            double *hq = &host[o], *hv = &host[o + len], *hp = &host[o + 2 * len], *hu = &host[o + 3 * len];
            factor_operator(kind, m, nr, dr, field_boundary, c->relax, C, hq, hv, hp, hu, &ax);
            OpCoef oc;
            oc.qT = c->coef_pool + o; oc.vT = c->coef_pool + o + len; oc.pT = c->coef_pool + o + 2 * len; oc.uT = c->coef_pool + o + 3 * len;
            oc.axis_inv = ax;
            c->ops[kind][m] = oc;
            hostdev.ops[kind][m] = oc;
            o += 4 * len;
        }
    for (double v : host)
        if (!std::isfinite(v)) { qpg_set_error("operator factorisation overflowed (nr=%d, max_mode=%d)", nr, max_mode); return QPG_ERR_UNSUPPORTED; }
    CUDA_TRY(cudaMemcpy(c->coef_pool, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
    CtxDev *dd;
    CUDA_TRY(cudaMalloc(&dd, sizeof(CtxDev)));
    CUDA_TRY(cudaMemcpy(dd, &hostdev, sizeof(CtxDev), cudaMemcpyHostToDevice));
    g_devops[c] = dd;
    CUDA_TRY(cudaMalloc(&c->conv_old, sizeof(double) * 2 * (nr + 2)));
    CUDA_TRY(cudaMemset(c->conv_old, 0, sizeof(double) * 2 * (nr + 2)));
    CUDA_TRY(cudaMalloc(&c->conv_out, sizeof(double) * 8));
    CUDA_TRY(cudaMemset(c->conv_out, 0, sizeof(double) * 8));
    CUDA_TRY(cudaMalloc(&c->flags, sizeof(int) * 16));
    CUDA_TRY(cudaMemset(c->flags, 0, sizeof(int) * 16));
    CUDA_TRY(cudaMalloc(&c->counters, sizeof(long long) * 4));
    CUDA_TRY(cudaMemset(c->counters, 0, sizeof(long long) * 4));
    c->cond_handle = 0;
    // field kernel shared memory: as much as the device allows (<= 227 KB)
    int maxsm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    int per_sys = NT_FIELD * C * (int)sizeof(double);
    int want = (2 * c->P > 4 ? 2 * c->P : 4) * per_sys + 4096;
    if (want > maxsm) want = (maxsm / 1024) * 1024;
    if (want < 4 * per_sys + 4096) { qpg_set_error("nr=%d needs %d B shared memory per system, device offers %d", nr, per_sys, maxsm); return QPG_ERR_UNSUPPORTED; }
    c->smem_field = want;
    CUDA_TRY(cudaFuncSetAttribute(k_field_prog, cudaFuncAttributeMaxDynamicSharedMemorySize, want));
    *out = c;
    return 0;
}

extern "C" int qpg_ctx_destroy(qpg_ctx c)
{
    if (!c) return 0;
    cudaStreamSynchronize(c->stream);
    for (auto &pe : c->tp_pending) { cudaEventDestroy(pe.second.first); cudaEventDestroy(pe.second.second); }
    for (auto ev : c->ev_pool) cudaEventDestroy(ev);
    cudaFree(c->scratch); cudaFree(c->coef_pool); cudaFree(c->conv_old); cudaFree(c->conv_out); cudaFree(c->flags); cudaFree(c->counters);
    auto it = g_devops.find(c);
    if (it != g_devops.end()) { cudaFree(it->second); g_devops.erase(it); }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}
int qpg_ctx_check_latches(qpg_ctx c, const int *fl)
{
    (void)c;
    if (fl[6]) { qpg_set_error("sweep kernel aborted (a grid barrier or strip exchange timed out): results on this context are invalid"); return QPG_ERR_STATE; }
    if (fl[7]) { qpg_set_error("neutral species: released electrons did not fit the particle set (npmax too small): charge was lost"); return QPG_ERR_STATE; }
    return 0;
}
extern "C" int qpg_ctx_sync(qpg_ctx c)
{
    ARG_TRY(c, "null ctx");
    int fl[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // sticky error latches written by kernels: see qpg_ctx_check_latches
    CUDA_TRY(cudaMemcpyAsync(fl, c->flags, sizeof(fl), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return qpg_ctx_check_latches(c, fl);
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#define NT 1024
#define FX(f, dim, j, pl, c) (f)[((size_t)(j) * P + (pl)) * (dim) + (c)]

__device__ __forceinline__ int pl_re(int m) { return m == 0 ? 0 : 2 * m - 1; }
__device__ __forceinline__ int pl_im(int m) { return 2 * m; }
__device__ __forceinline__ int mode_of(int pl) { return (pl + 1) >> 1; }
__device__ __forceinline__ bool is_im(int pl) { return pl > 0 && (pl & 1) == 0; }

__device__ double block_sum(double v, double *red)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ double block_max(double v, double *red)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// smem index of radial node i (1-based) inside one system's array (natural order)
__device__ __forceinline__ int sidx(int i, int, int) { return i - 1; }

// Block-wide forward / reverse scans for NSR systems at once; every thread owns IPT consecutive radial nodes.
// in : fa[j][s] = q*d, fb[j][s] = v*d of the thread's nodes
// out: fa[j][s] = sum over nodes <= own (inclusive prefix), fb[j][s] = sum over nodes > own (exclusive suffix)
template <int IPT, int NSR>
__device__ __forceinline__ void scan_round(double (&fa)[IPT][NSR], double (&fb)[IPT][NSR], double *sA, double *sB)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NSR; s++) {
        double run = 0.0;
#pragma unroll
        for (int j = 0; j < IPT; j++) { run += fa[j][s]; fa[j][s] = run; }
        double incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        if (lane == 31) sA[s * 32 + warp] = incl;
#pragma unroll
        for (int j = 0; j < IPT; j++) fa[j][s] += excl;
        run = 0.0;
#pragma unroll
        for (int j = IPT - 1; j >= 0; j--) { double t = fb[j][s]; fb[j][s] = run; run += t; }
        incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += t; }
        excl = __shfl_down_sync(0xffffffffu, incl, 1);
        if (lane == 31) excl = 0.0;
        if (lane == 0) sB[s * 32 + warp] = incl;
#pragma unroll
        for (int j = 0; j < IPT; j++) fb[j][s] += excl;
    }
    __syncthreads();
    if (warp < NSR) {            // warp s: exclusive prefix of the 32 warp totals of system s
        double v = sA[warp * 32 + lane], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        sA[warp * 32 + lane] = lane == 0 ? 0.0 : excl;
    } else if (warp >= 16 && warp < 16 + NSR) {  // warp 16+s: exclusive suffix
        const int s = warp - 16;
        double v = sB[s * 32 + lane], incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += t; }
        double excl = __shfl_down_sync(0xffffffffu, incl, 1);
        sB[s * 32 + lane] = lane == 31 ? 0.0 : excl;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NSR; s++) {
        const double oa = sA[s * 32 + warp], ob = sB[s * 32 + warp];
#pragma unroll
        for (int j = 0; j < IPT; j++) { fa[j][s] += oa; fb[j][s] += ob; }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// right-hand sides and post-processing of the five solve families.  `i` = radial node 1..nr, `pl` = plane.
// Formulas follow the cited Fortran line by line (single r-owner branches: idproc == 0 == nvp-1).
// ------------------------------------------------------------------------------------------------
struct SolveCtx { int nr, M, P, logC, stride; double dr, idr, idrh; };

// field_b_class.f03:205-303 set_source_bz ; field_e_class.f03:148-263 set_source_ez (m=0 row 1 handled by caller)
__device__ double rhs_bz(const SolveCtx &sc, const double *cu, int pl, int i)
{
    const int P = sc.P, nr = sc.nr, m = mode_of(pl);
    const double idr = sc.idr, idrh = sc.idrh;
    if (m == 0) {
        if (i == 1) return -2.0 * idr * FX(cu, 3, 2, 0, 1);
        double ir = idr / (double)(i - 1);
        if (i == nr) return -idrh * (3.0 * FX(cu, 3, nr, 0, 1) - 4.0 * FX(cu, 3, nr - 1, 0, 1) + FX(cu, 3, nr - 2, 0, 1)) - ir * FX(cu, 3, nr, 0, 1);
        return -idrh * (FX(cu, 3, i + 1, 0, 1) - FX(cu, 3, i - 1, 0, 1)) - ir * FX(cu, 3, i, 0, 1);
    }
    const bool im = is_im(pl);
    const int ps = pl, po = im ? pl_re(m) : pl_im(m);  // same plane / partner plane
    const double sg = im ? 1.0 : -1.0;                 // re: - m ir J_r,im ; im: + m ir J_r,re
    if (i == 1) {
        if ((m & 1) == 0) return -2.0 * idr * FX(cu, 3, 2, ps, 1) + sg * m * idr * FX(cu, 3, 2, po, 0);
        return 0.0;
    }
    if (i == 2 && m == 1) {
        double ir = idr;
        return -idr * (FX(cu, 3, 3, ps, 1) - FX(cu, 3, 2, ps, 1)) - ir * FX(cu, 3, 2, ps, 1) + sg * m * ir * FX(cu, 3, 2, po, 0);
    }
    double ir = idr / (double)(i - 1);
    if (i == nr)
        return -idrh * (3.0 * FX(cu, 3, nr, ps, 1) - 4.0 * FX(cu, 3, nr - 1, ps, 1) + FX(cu, 3, nr - 2, ps, 1)) - ir * FX(cu, 3, nr, ps, 1) + sg * m * ir * FX(cu, 3, nr, po, 0);
    return -idrh * (FX(cu, 3, i + 1, ps, 1) - FX(cu, 3, i - 1, ps, 1)) - ir * FX(cu, 3, i, ps, 1) + sg * m * ir * FX(cu, 3, i, po, 0);
}

__device__ double rhs_ez(const SolveCtx &sc, const double *cu, int pl, int i)
{
    const int P = sc.P, nr = sc.nr, m = mode_of(pl);
    const double idr = sc.idr, idrh = sc.idrh;
    if (m == 0) {
        if (i == 1) return 0.0;  // filled in by the divergence reduction
        if (i == nr) { double ir = idr / (double)(nr - 1); return idr * (FX(cu, 3, nr, 0, 0) - FX(cu, 3, nr - 1, 0, 0)) + ir * FX(cu, 3, nr, 0, 0); }
        double ir = idr / (double)(i - 1);
        return idrh * (FX(cu, 3, i + 1, 0, 0) - FX(cu, 3, i - 1, 0, 0)) + ir * FX(cu, 3, i, 0, 0);
    }
    const bool im = is_im(pl);
    const int ps = pl, po = im ? pl_re(m) : pl_im(m);
    const double sg = im ? 1.0 : -1.0;
    if (i == 1) {
        if ((m & 1) == 0) return 2.0 * idr * FX(cu, 3, 2, ps, 0) + sg * m * idr * FX(cu, 3, 2, po, 1);
        return 0.0;
    }
    if (i == 2 && m == 1) {
        double ir = idr;
        return idr * (FX(cu, 3, 3, ps, 0) - FX(cu, 3, 2, ps, 0)) + ir * FX(cu, 3, 2, ps, 0) + sg * m * ir * FX(cu, 3, 2, po, 1);
    }
    double ir = idr / (double)(i - 1);
    if (i == nr)
        return idrh * (3.0 * FX(cu, 3, nr, ps, 0) - 4.0 * FX(cu, 3, nr - 1, ps, 0) + FX(cu, 3, nr - 2, ps, 0)) + ir * FX(cu, 3, nr, ps, 0) + sg * m * ir * FX(cu, 3, nr, po, 1);
    return idrh * (FX(cu, 3, i + 1, ps, 0) - FX(cu, 3, i - 1, ps, 0)) + ir * FX(cu, 3, i, ps, 0) + sg * m * ir * FX(cu, 3, i, po, 1);
}

// field_b_class.f03:360-506 set_source_bt_iter.  which = 0: B+ (buf1), 1: B- (buf2).
__device__ double rhs_bt_iter(const SolveCtx &sc, const double *dcu, const double *cu, const double *b, double relax_idr2,
                              int which, int pl, int i)
{
    const int P = sc.P, nr = sc.nr, m = mode_of(pl);
    const double idr = sc.idr, idrh = sc.idrh;
    if (m == 0) {
        if (i == 1) return 0.0;
        if (which == 0) return -FX(dcu, 2, i, 0, 1) - FX(b, 3, i, 0, 0) * relax_idr2;
        double dj;
        if (i == 2) dj = idr * (FX(cu, 3, 3, 0, 2) - FX(cu, 3, 2, 0, 2));
        else if (i == nr) dj = idrh * (3.0 * FX(cu, 3, nr, 0, 2) - 4.0 * FX(cu, 3, nr - 1, 0, 2) + FX(cu, 3, nr - 2, 0, 2));
        else dj = idrh * (FX(cu, 3, i + 1, 0, 2) - FX(cu, 3, i - 1, 0, 2));
        return FX(dcu, 2, i, 0, 0) + dj - FX(b, 3, i, 0, 1) * relax_idr2;
    }
    const int pr = pl_re(m), pi = pl_im(m);
    double s1_re, s1_im, s2_re, s2_im;
    if (i == 1) {
        if (m == 1) {
            s1_re = -FX(dcu, 2, 1, pr, 1) + idr * m * FX(cu, 3, 2, pi, 2);
            s1_im = -FX(dcu, 2, 1, pi, 1) - idr * m * FX(cu, 3, 2, pr, 2);
            s2_re = FX(dcu, 2, 1, pr, 0) + idr * FX(cu, 3, 2, pr, 2);
            s2_im = FX(dcu, 2, 1, pi, 0) + idr * FX(cu, 3, 2, pi, 2);
        } else if ((m & 1) == 0) {
            s1_re = s1_im = s2_re = s2_im = 0.0;
        } else {  // :453-458 as written upstream
            s1_re = idr * m * FX(cu, 3, 2, pi, 2);
            s1_im = idr * m * FX(cu, 3, 2, pr, 2);
            s2_re = idr * FX(cu, 3, 2, pr, 2);
            s2_im = idr * FX(cu, 3, 2, pi, 2);
        }
    } else {
        double ir = idr / (double)(i - 1);
        s1_re = -FX(dcu, 2, i, pr, 1) + m * FX(cu, 3, i, pi, 2) * ir;
        s1_im = -FX(dcu, 2, i, pi, 1) - m * FX(cu, 3, i, pr, 2) * ir;
        if (i == nr) {
            s2_re = FX(dcu, 2, nr, pr, 0) + idrh * (3.0 * FX(cu, 3, nr, pr, 2) - 4.0 * FX(cu, 3, nr - 1, pr, 2) + FX(cu, 3, nr - 2, pr, 2));
            s2_im = FX(dcu, 2, nr, pi, 0) + idrh * (3.0 * FX(cu, 3, nr, pi, 2) - 4.0 * FX(cu, 3, nr - 1, pi, 2) + FX(cu, 3, nr - 2, pi, 2));
        } else {
            s2_re = FX(dcu, 2, i, pr, 0) + idrh * (FX(cu, 3, i + 1, pr, 2) - FX(cu, 3, i - 1, pr, 2));
            s2_im = FX(dcu, 2, i, pi, 0) + idrh * (FX(cu, 3, i + 1, pi, 2) - FX(cu, 3, i - 1, pi, 2));
        }
    }
    const double brr = FX(b, 3, i, pr, 0), bri = FX(b, 3, i, pi, 0), bpr = FX(b, 3, i, pr, 1), bpi = FX(b, 3, i, pi, 1);
    const bool im = is_im(pl);
    if (which == 0) return im ? (s1_im + s2_re - (bri + bpr) * relax_idr2) : (s1_re - s2_im - (brr - bpi) * relax_idr2);
    return im ? (s1_im - s2_re - (bri - bpr) * relax_idr2) : (s1_re + s2_im - (brr + bpi) * relax_idr2);
}

// ------------------------------------------------------------------------------------------------
// the program kernel
// ------------------------------------------------------------------------------------------------
template <int IPT>
__device__ void op_solve(const FProg &pg, const FOp &op, double *smem, double *red, const OpCoef **sysop_s, double *scanbuf)
{
    constexpr int NSR = IPT == 1 ? 6 : (IPT == 2 ? 3 : 2);  // systems scanned together (register budget)
    const int nr = pg.nr, M = pg.M, P = pg.P, logC = 0, stride = 1;
    const int per = NT_FIELD * IPT;
    SolveCtx sc; sc.nr = nr; sc.M = M; sc.P = P; sc.logC = 0; sc.stride = 1; sc.dr = pg.dr; sc.idr = 1.0 / pg.dr; sc.idrh = 0.5 * sc.idr;
    const int fac = (op.code == FOP_BTITER) ? 2 : 1;  // systems per plane
    const int maxsys = op.i3;                          // systems whose solutions fit in shared memory
    const double relax_idr2 = op.s0 * (sc.idr * sc.idr);
    const int tid = threadIdx.x, nt = blockDim.x;
    double *X = smem;
    double *sA = scanbuf, *sB = scanbuf + NSR * 32;
    // batches hold whole modes, so the (re, im) planes of a mode and its B+/B- systems are resident together
    int m0 = 0;
    while (m0 <= M) {
        int m1 = m0, npl = 0;
        while (m1 <= M) { int add = (m1 == 0) ? 1 : 2; if (fac * (npl + add) > maxsys) break; npl += add; m1++; }
        if (m1 == m0) { m1 = m0 + 1; npl = (m0 == 0) ? 1 : 2; }  // cannot happen (host check); avoid an endless loop
        const int pl0 = (m0 == 0) ? 0 : 2 * m0 - 1;
        const int ns = fac * npl;
        __syncthreads();
        if (tid < ns) {
            int which = tid / npl, pl = pl0 + tid % npl, m = mode_of(pl), kind;
            switch (op.code) {
            case FOP_PSI: kind = FK_PSI; break;
            case FOP_BT: kind = FK_BT; break;
            case FOP_BZ: kind = FK_BZ; break;
            case FOP_EZ: kind = FK_EZ; break;
            default: kind = which ? FK_BMINUS : FK_BPLUS; break;
            }
            sysop_s[tid] = pg.ops + kind * (QPG_MAX_MODE + 1) + m;
        }
        __syncthreads();
        for (int r0 = 0; r0 < ns; r0 += NSR) {
            double fa[IPT][NSR], fb[IPT][NSR];
            // right-hand sides of this thread's nodes
#pragma unroll
            for (int ss = 0; ss < NSR; ss++) {
                const int sl = r0 + ss;
#pragma unroll
                for (int j = 0; j < IPT; j++) {
                    const int i = tid * IPT + j + 1;
                    double v = 0.0;
                    if (sl < ns && i <= nr) {
                        const int which = sl / npl, pl = pl0 + sl % npl;
                        switch (op.code) {
                        case FOP_PSI: case FOP_BT: v = -1.0 * FX(op.a, 1, i, pl, 0); break;
                        case FOP_BZ: v = rhs_bz(sc, op.a, pl, i); break;
                        case FOP_EZ: v = rhs_ez(sc, op.a, pl, i); break;
                        default: v = rhs_bt_iter(sc, op.a, op.b, op.c, relax_idr2, which, pl, i); break;
                        }
                    }
                    fa[j][ss] = v;
                }
            }
            if (op.code == FOP_EZ && m0 == 0 && r0 == 0) {
                // m=0 divergence fix, field_e_class.f03:189-209: row 1 = -8 * ( sum_{i=2}^{nr-2} rhs_i (i-1) - edge term )
                double part = 0.0;
#pragma unroll
                for (int j = 0; j < IPT; j++) { const int i = tid * IPT + j + 1; if (i >= 2 && i <= nr - 2) part += fa[j][0] * (double)(i - 1); }
                double div = block_sum(part, red);
                if (tid == 0) {
                    const double *cu = op.a;
                    div = div - sc.idrh * (FX(cu, 3, nr - 2, 0, 0) + FX(cu, 3, nr - 1, 0, 0)) * ((double)nr - 2.5);
                    fa[0][0] = -8.0 * div;
                }
            }
            double d0[NSR];  // node-1 right-hand sides (decoupled axis rows)
#pragma unroll
            for (int ss = 0; ss < NSR; ss++) d0[ss] = fa[0][ss];
#pragma unroll
            for (int ss = 0; ss < NSR; ss++) {
                const int sl = r0 + ss;
                if (sl < ns) {
                    const OpCoef *oc = sysop_s[sl];
#pragma unroll
                    for (int j = 0; j < IPT; j++) { const int t = tid * IPT + j; const double d = fa[j][ss]; fa[j][ss] = __ldg(oc->qT + t) * d; fb[j][ss] = __ldg(oc->vT + t) * d; }
                } else {
#pragma unroll
                    for (int j = 0; j < IPT; j++) { fa[j][ss] = 0.0; fb[j][ss] = 0.0; }
                }
            }
            scan_round<IPT, NSR>(fa, fb, sA, sB);
#pragma unroll
            for (int ss = 0; ss < NSR; ss++) {
                const int sl = r0 + ss;
                if (sl < ns) {
                    const OpCoef *oc = sysop_s[sl];
#pragma unroll
                    for (int j = 0; j < IPT; j++) {
                        const int t = tid * IPT + j;
                        double x = __ldg(oc->pT + t) * fa[j][ss] + __ldg(oc->uT + t) * fb[j][ss];
                        if (t == 0 && oc->axis_inv != 0.0) x = d0[ss] * oc->axis_inv;
                        X[(size_t)sl * per + t] = x;
                    }
                }
            }
        }
        __syncthreads();
        if (op.code == FOP_PSI) {
            double *psi = op.b;
            for (int k = tid; k < ns * nr; k += nt) {
                int sl = k / nr, i = k % nr + 1, pl = pl0 + sl;
                double v = X[(size_t)sl * per + sidx(i, logC, stride)];
                if (pl > 0 && i == 1) v = 0.0;
                FX(psi, 1, i, pl, 0) = v;
            }
        } else if (op.code == FOP_BZ || op.code == FOP_EZ) {
            double *f = op.b;
            for (int k = tid; k < ns * nr; k += nt) {
                int sl = k / nr, i = k % nr + 1, pl = pl0 + sl;
                double v = X[(size_t)sl * per + sidx(i, logC, stride)];
                if (pl > 0 && i == 1) v = 0.0;
                FX(f, 3, i, pl, 2) = v;
            }
        } else if (op.code == FOP_BT) {
            // field_b_class.f03:545-701 get_solution_bt
            double *b = op.b;
            for (int k = tid; k < ns * nr; k += nt) {
                int sl = k / nr, i = k % nr + 1, pl = pl0 + sl, m = mode_of(pl);
                const double *Xs = X + (size_t)sl * per;
                double bphi;
                if (i == 1) bphi = (m == 1) ? -sc.idr * Xs[sidx(2, logC, stride)] : 0.0;
                else if (i == nr) bphi = -sc.idrh * (3.0 * Xs[sidx(nr, logC, stride)] - 4.0 * Xs[sidx(nr - 1, logC, stride)] + Xs[sidx(nr - 2, logC, stride)]);
                else bphi = -sc.idrh * (Xs[sidx(i + 1, logC, stride)] - Xs[sidx(i - 1, logC, stride)]);
                FX(b, 3, i, pl, 1) = bphi;
                double br = 0.0;
                if (m > 0) {
                    // B_r,re = -(m/r) Phi_im ; B_r,im = +(m/r) Phi_re  (partner plane is adjacent: re = 2m-1, im = 2m)
                    const bool im = is_im(pl);
                    const double *Xo = X + (size_t)(im ? sl - 1 : sl + 1) * per;
                    const double sg = im ? 1.0 : -1.0;
                    if (i == 1) br = (m == 1) ? sg * sc.idr * m * Xo[sidx(2, logC, stride)] : 0.0;
                    else { double ir = sc.idr / (double)(i - 1); br = sg * ir * m * Xo[sidx(i, logC, stride)]; }
                }
                FX(b, 3, i, pl, 0) = br;
            }
        } else {  // FOP_BTITER : field_b_class.f03:703-758 get_solution_bt_iter ; slots [0,npl) = B+, [npl,2npl) = B-
            double *b = op.c;
            for (int k = tid; k < npl * nr; k += nt) {
                int sl = k / nr, i = k % nr + 1, pl = pl0 + sl, m = mode_of(pl);
                const double *Xp = X + (size_t)sl * per, *Xm = X + (size_t)(npl + sl) * per;
                int ix = sidx(i, logC, stride);
                if (m == 0) {
                    FX(b, 3, i, 0, 0) = (i == 1) ? 0.0 : Xp[ix];
                    FX(b, 3, i, 0, 1) = (i == 1) ? 0.0 : Xm[ix];
                } else {
                    const bool im = is_im(pl);
                    const int so = im ? sl - 1 : sl + 1;
                    const double *Xpo = X + (size_t)so * per, *Xmo = X + (size_t)(npl + so) * per;
                    double br = 0.5 * (Xp[ix] + Xm[ix]);
                    // Re(Bphi) = 0.5*(B+_im - B-_im) ; Im(Bphi) = 0.5*(-B+_re + B-_re)
                    double bp = im ? 0.5 * (-Xpo[ix] + Xmo[ix]) : 0.5 * (Xpo[ix] - Xmo[ix]);
                    if (i == 1 && m != 1) { br = 0.0; bp = 0.0; }
                    FX(b, 3, i, pl, 0) = br;
                    FX(b, 3, i, pl, 1) = bp;
                }
            }
        }
        m0 = m1;
    }
    __syncthreads();
}

// field_e_class.f03:412-514 solve_field_et
__device__ void op_et(const FProg &pg, const FOp &op)
{
    const int nr = pg.nr, P = pg.P;
    const double idr = 1.0 / pg.dr, idrh = idr * 0.5;
    const double *b = op.a, *psi = op.b;
    double *e = op.c;
    for (int k = threadIdx.x; k < P * nr; k += blockDim.x) {
        int pl = k % P, i = k / P + 1, m = mode_of(pl);
        double er, ephi;
        if (m == 0) {
            if (i == 1) { er = 0.0; ephi = 0.0; }
            else {
                if (i == nr) er = FX(b, 3, nr, 0, 1) + idrh * (4.0 * FX(psi, 1, nr - 1, 0, 0) - FX(psi, 1, nr - 2, 0, 0) - 3.0 * FX(psi, 1, nr, 0, 0));
                else er = FX(b, 3, i, 0, 1) - idrh * (FX(psi, 1, i + 1, 0, 0) - FX(psi, 1, i - 1, 0, 0));
                ephi = -FX(b, 3, i, 0, 0);
            }
        } else {
            const bool im = is_im(pl);
            const int po = im ? pl - 1 : pl + 1;
            const double sg = im ? -1.0 : 1.0;  // E_phi,re = -B_r,re + (m/r) psi_im ; E_phi,im = -B_r,im - (m/r) psi_re
            if (i == 1) {
                if (m == 1) {
                    er = FX(b, 3, 1, pl, 1) - idr * FX(psi, 1, 2, pl, 0);
                    ephi = -FX(b, 3, 1, pl, 0) + sg * idr * FX(psi, 1, 2, po, 0);
                } else { er = 0.0; ephi = 0.0; }
            } else {
                double ir = idr / (double)(i - 1);
                if (i == nr) er = FX(b, 3, nr, pl, 1) + idrh * (4.0 * FX(psi, 1, nr - 1, pl, 0) - FX(psi, 1, nr - 2, pl, 0) - 3.0 * FX(psi, 1, nr, pl, 0));
                else er = FX(b, 3, i, pl, 1) - idrh * (FX(psi, 1, i + 1, pl, 0) - FX(psi, 1, i - 1, pl, 0));
                ephi = -FX(b, 3, i, pl, 0) + sg * ir * m * FX(psi, 1, i, po, 0);
            }
        }
        FX(e, 3, i, pl, 0) = er;
        FX(e, 3, i, pl, 1) = ephi;
    }
}

// field_src_class.f03:273-405 solve_field_djdxi
__device__ void op_djdxi(const FProg &pg, const FOp &op)
{
    const int nr = pg.nr, P = pg.P;
    const double idr = 1.0 / pg.dr, idrh = idr * 0.5;
    const double *acu = op.a, *amu = op.b;
    double *dcu = op.c;
    for (int k = threadIdx.x; k < P * nr * 2; k += blockDim.x) {
        int c = k % 2, pl = (k / 2) % P, i = k / (2 * P) + 1, m = mode_of(pl);
        double v;
        if (m == 0) {
            if (i == 1) v = 0.0;
            else {
                double ir = idr / (double)(i - 1);
                if (i == nr) v = FX(acu, 2, nr, 0, c) + idrh * (4.0 * FX(amu, 3, nr - 1, 0, c) - FX(amu, 3, nr - 2, 0, c) - 3.0 * FX(amu, 3, nr, 0, c)) - ir * FX(amu, 3, nr, 0, c);
                else v = FX(acu, 2, i, 0, c) - idrh * (FX(amu, 3, i + 1, 0, c) - FX(amu, 3, i - 1, 0, c)) - ir * FX(amu, 3, i, 0, c);
            }
        } else {
            const bool im = is_im(pl);
            const int po = im ? pl - 1 : pl + 1;
            const double sg = im ? -1.0 : 1.0;  // re: + m ir amu_im(c+1) ; im: - m ir amu_re(c+1)
            if (i == 1) {
                if (m == 1) v = FX(acu, 2, 1, pl, c) - 2.0 * idr * FX(amu, 3, 2, pl, c) + sg * m * idr * FX(amu, 3, 2, po, c + 1);
                else v = 0.0;
            } else if (i == 2 && m == 2) {
                double ir = idr;
                v = FX(acu, 2, 2, pl, c) - idr * (FX(amu, 3, 3, pl, c) - FX(amu, 3, 2, pl, c)) - ir * FX(amu, 3, 2, pl, c) + sg * m * ir * FX(amu, 3, 2, po, c + 1);
            } else {
                double ir = idr / (double)(i - 1);
                if (i == nr)
                    v = FX(acu, 2, nr, pl, c) + idrh * (4.0 * FX(amu, 3, nr - 1, pl, c) - FX(amu, 3, nr - 2, pl, c) - 3.0 * FX(amu, 3, nr, pl, c)) - ir * FX(amu, 3, nr, pl, c) + sg * m * ir * FX(amu, 3, nr, po, c + 1);
                else
                    v = FX(acu, 2, i, pl, c) - idrh * (FX(amu, 3, i + 1, pl, c) - FX(amu, 3, i - 1, pl, c)) - ir * FX(amu, 3, i, pl, c) + sg * m * ir * FX(amu, 3, i, po, c + 1);
            }
        }
        FX(dcu, 2, i, pl, c) = v;
    }
}

// axis rules of part2d_class.f03:312-333 (q) and :916-981 (cu, dcu, amu):  value = fix( old + raw ), raw cleared
__device__ __forceinline__ double axis_fix_q(int j, int pl, double v)
{
    if (j == 0) return 0.0;
    if (j == 1) return pl == 0 ? 8.0 * v : 0.0;
    return v * (1.0 / (double)(j - 1));
}
// comp8: 0..2 cu, 3..4 dcu, 5..7 amu
__device__ __forceinline__ double axis_fix_amj(int j, int pl, int comp8, double v)
{
    if (j == 0) return 0.0;
    if (j == 1) {
        int m = (pl + 1) >> 1;
        if (m == 0) return comp8 == 2 ? 8.0 * v : 0.0;
        if (m == 1) return (comp8 == 0 || comp8 == 1 || comp8 == 3 || comp8 == 4) ? 8.0 * v : 0.0;
        if (m == 2) return comp8 >= 5 ? 8.0 * v : 0.0;
        return 0.0;
    }
    return v * (1.0 / (double)(j - 1));
}

// fields/ufield_class.f03:274-339 smooth_f1 with stencil (1,2,1)/4 and the per-component axis policy of
// field_src_class.f03:102-271 (kind 0 rho, 1 jay, 2 djdxi).  One pass; uses smem as the temporary.
__device__ bool ax_smooth_policy(int kind, int m, int c)
{
    if (kind == 0) return m == 0;
    if (kind == 1) return m == 0 ? (c == 2) : (m == 1 ? (c < 2) : false);
    return m == 1;
}
__device__ void op_smooth(const FProg &pg, const FOp &op, double *smem)
{
    const int nr = pg.nr, P = pg.P, dim = op.da, kind = op.i0;
    double *f = op.a;
    const double km1 = 0.25, k0 = 0.5, kp1 = 0.25;
    const int n = P * dim;
    if (op.b) smem = op.b;     // image larger than the shared-memory scratch: a global temporary (qpg_field_smooth)
    for (int k = threadIdx.x; k < nr * n; k += blockDim.x) {
        int j = k / n + 1, r = k % n, pl = r / dim, c = r % dim, m = mode_of(pl);
        double v;
        if (j == 1) v = ax_smooth_policy(kind, m, c) ? (k0 + kp1) * FX(f, dim, 1, pl, c) + 8.0 * kp1 * FX(f, dim, 2, pl, c) : 0.0;
        else if (j == 2) v = k0 * FX(f, dim, 2, pl, c) + 0.125 * km1 * FX(f, dim, 1, pl, c) + 2.0 * kp1 * FX(f, dim, 3, pl, c);
        else { int ri = j - 1; v = k0 * FX(f, dim, j, pl, c) + (1.0 - 1.0 / ri) * km1 * FX(f, dim, j - 1, pl, c) + (1.0 + 1.0 / ri) * kp1 * FX(f, dim, j + 1, pl, c); }
        smem[k] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nr * n; k += blockDim.x) { int j = k / n + 1, r = k % n; f[(size_t)j * n + r] = smem[k]; }
}

__global__ void __launch_bounds__(NT, 1) k_field_prog(const __grid_constant__ FProg pg)
{
    extern __shared__ double smem_all[];
    __shared__ double red[40];
    __shared__ const OpCoef *sysop_s[64];
    __shared__ double scanbuf[2 * 6 * 32];
    const int nr = pg.nr, P = pg.P, tid = threadIdx.x, nt = blockDim.x;
    double *smem = smem_all;
    const bool done = pg.flags[0] != 0;
    for (int io = 0; io < pg.nops; io++) {
        const FOp &op = pg.op[io];
        if ((op.flags & FOPF_SKIP_IF_DONE) && done) continue;
        const size_t n1a = (size_t)(nr + 2) * P * op.da;
        switch (op.code) {
        case FOP_ZERO: for (size_t k = tid; k < n1a; k += nt) op.a[k] = 0.0; break;
        case FOP_COPY: for (size_t k = tid; k < n1a; k += nt) op.b[k] = op.a[k]; break;
        case FOP_ADD: for (size_t k = tid; k < n1a; k += nt) op.b[k] = op.b[k] + op.a[k]; break;
        case FOP_ADD3: for (size_t k = tid; k < n1a; k += nt) op.c[k] = op.a[k] + op.b[k]; break;
        case FOP_SCALE: for (size_t k = tid; k < n1a; k += nt) op.a[k] = op.a[k] * op.s0; break;
        case FOP_ADD_DIM:  // b(:, i1) += a(:, i0) over all nodes and planes
            for (int k = tid; k < (nr + 2) * P; k += nt) op.b[(size_t)k * op.db + op.i1] = op.b[(size_t)k * op.db + op.i1] + op.a[(size_t)k * op.da + op.i0];
            break;
        case FOP_SLICE_1TO2: { const int sl = op.i0 > 0 ? op.i0 : pg.flags[3]; double *dst = op.b + (size_t)(sl - 1) * n1a; for (size_t k = tid; k < n1a; k += nt) dst[k] = op.a[k]; } break;
        case FOP_SLICE_2TO1: { const int sl = op.i0 > 0 ? op.i0 : pg.flags[3]; const double *src = op.b + (size_t)(sl - 1) * n1a; for (size_t k = tid; k < n1a; k += nt) op.a[k] = src[k]; } break;
        case FOP_ZERO_F2: { size_t n = n1a * (size_t)op.i0; for (size_t k = tid; k < n; k += nt) op.a[k] = 0.0; } break;
        case FOP_ADD_F2: { size_t n = n1a * (size_t)op.i0; for (size_t k = tid; k < n; k += nt) op.b[k] = op.b[k] + op.a[k]; } break;
        case FOP_QFIX:  // a = raw acc1 [(nr+2)][P], b = q field (dim 1)
            for (int k = tid; k < (nr + 2) * P; k += nt) { int j = k / P, pl = k % P; op.b[k] = axis_fix_q(j, pl, op.b[k] + op.a[k]); op.a[k] = 0.0; }
            if (op.i1 && tid == 0) pg.counters[0] += (long long)*(const int *)op.c;  // particle-slice updates
            break;
        case FOP_AMJFIX:  // a = raw acc8, b = cu, c = dcu, d = amu
            for (int k = tid; k < (nr + 2) * P * 8; k += nt) {
                int c8 = k % 8, np = k / 8, j = np / P, pl = np % P;
                double raw = op.a[k];
                op.a[k] = 0.0;
                if (c8 < 3) { double *t = &op.b[(size_t)np * 3 + c8]; *t = axis_fix_amj(j, pl, c8, *t + raw); }
                else if (c8 < 5) { double *t = &op.c[(size_t)np * 2 + (c8 - 3)]; *t = axis_fix_amj(j, pl, c8, *t + raw); }
                else { double *t = &op.d[(size_t)np * 3 + (c8 - 5)]; *t = axis_fix_amj(j, pl, c8, *t + raw); }
            }
            break;
        case FOP_PSI: case FOP_BT: case FOP_BZ: case FOP_EZ: case FOP_BTITER:
            if (pg.logC == 0) op_solve<1>(pg, op, smem, red, sysop_s, scanbuf);
            else if (pg.logC == 1) op_solve<2>(pg, op, smem, red, sysop_s, scanbuf);
            else op_solve<4>(pg, op, smem, red, sysop_s, scanbuf);
            break;
        case FOP_ET: op_et(pg, op); break;
        case FOP_ETBEAM:
            for (int k = tid; k < nr * P; k += nt) { int i = k / P + 1, pl = k % P; FX(op.b, 3, i, pl, 0) = FX(op.a, 3, i, pl, 1); FX(op.b, 3, i, pl, 1) = -FX(op.a, 3, i, pl, 0); }
            break;
        case FOP_DJDXI: op_djdxi(pg, op); break;
        case FOP_SMOOTH: op_smooth(pg, op, smem); break;
        case FOP_CONV_RECORD:  // simulation_class.f03:548-558 ; i0 = component (0-based)
            for (int i = 1 + tid; i <= nr; i += nt) {
                double sre = 0.0, sim = 0.0;
                for (int pl = 0; pl < P; pl++) { double v = fabs(FX(op.a, op.da, i, pl, op.i0)); if (is_im(pl)) sim += v; else sre += v; }
                pg.conv_old[i] = sre; pg.conv_old[nr + 2 + i] = sim;
            }
            break;
        case FOP_CONV_COMPARE: {  // :560-599 ; s0 = reltol, s1 = abstol ; sets flags[0] (done) unless i1 == 0
            double mo = 0.0, mn = 0.0;
            for (int i = 1 + tid; i <= nr; i += nt) {
                double ore = pg.conv_old[i], oim = pg.conv_old[nr + 2 + i];
                mo = fmax(mo, ore * ore + oim * oim);
                double sre = ore, sim = oim;
                for (int pl = 0; pl < P; pl++) { double v = fabs(FX(op.a, op.da, i, pl, op.i0)); if (is_im(pl)) sim -= v; else sre -= v; }
                pg.conv_old[i] = sre; pg.conv_old[nr + 2 + i] = sim;
                mn = fmax(mn, sre * sre + sim * sim);
            }
            double old_norm = sqrt(block_max(mo, red));
            double abs_res = sqrt(block_max(mn, red));
            if (tid == 0) {
                double rel = old_norm > 2.220446049250313e-16 ? abs_res / old_norm : 1.7976931348623157e308;
                pg.conv_out[0] = rel; pg.conv_out[1] = abs_res;
                if (op.i1) {
                    pg.counters[1] += 1;  // PC iterations executed
                    int it = pg.flags[2] + 1;
                    pg.flags[2] = it;
                    const bool fin = rel < op.s0 || abs_res < op.s1 || it >= op.i2;
                    if (fin) pg.flags[0] = 1;
                    if (pg.cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)pg.cond_handle, fin ? 0u : 1u);
                }
            }
        } break;
        case FOP_PC_BEGIN: if (tid == 0) { pg.flags[0] = 0; pg.flags[2] = 0; } break;
        case FOP_SET_FLAG: if (tid == 0) { if (op.i1 >= 0) pg.flags[op.i0] = op.i1; else { pg.flags[3] += 1; pg.flags[4] += 1; } } break;
        case FOP_PACK: {  // node-interleaved -> wire [P][nr+2][dim] ; a = source image, b = wire buffer
            const int dim = op.da;
            for (size_t k = tid; k < n1a; k += nt) { int c = k % dim; size_t r = k / dim; int j = r % (nr + 2), pl = r / (nr + 2); op.b[k] = FX(op.a, dim, j, pl, c); }
        } break;
        case FOP_UNPACK: {
            const int dim = op.da;
            for (size_t k = tid; k < n1a; k += nt) {
                int c = k % dim; size_t r = k / dim; int j = r % (nr + 2), pl = r / (nr + 2);
                if (op.i0) FX(op.a, dim, j, pl, c) = FX(op.a, dim, j, pl, c) + op.b[k]; else FX(op.a, dim, j, pl, c) = op.b[k];
            }
        } break;
        default: break;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// host side: program builder
// ------------------------------------------------------------------------------------------------
FProgBuilder::FProgBuilder(qpg_ctx c) : ctx(c)
{
    memset(&prog, 0, sizeof(prog));
    prog.nr = c->nr; prog.M = c->M; prog.P = c->P; prog.logC = c->logC; prog.dr = c->dr;
    prog.ops = qpg_ctx_dev_ops(c);
    prog.conv_old = c->conv_old; prog.conv_out = c->conv_out; prog.flags = c->flags;
    prog.counters = c->counters; prog.cond_handle = c->cond_handle;
}
FOp &FProgBuilder::add(int code)
{
    if (prog.nops >= QPG_MAX_FOPS) { fprintf(stderr, "qpad_b200: field program overflow\n"); abort(); }
    FOp &o = prog.op[prog.nops++];
    memset(&o, 0, sizeof(o));
    o.code = code;
    int per_sys = NT_FIELD * ctx->C * (int)sizeof(double);
    int maxsys = (ctx->smem_field - 1024) / per_sys;
    if (maxsys > 2 * ctx->P) maxsys = 2 * ctx->P;
    if (maxsys > 60) maxsys = 60;
    o.i3 = maxsys;
    return o;
}
int FProgBuilder::launch(int tp_event)
{
    if (prog.nops == 0) return 0;
    TprofScope tp(ctx, tp_event);
    k_field_prog<<<1, NT, ctx->smem_field, ctx->stream>>>(prog);
    count_launch(ctx);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// field objects
// ------------------------------------------------------------------------------------------------
extern "C" int qpg_field_create(qpg_field *out, qpg_ctx ctx, int dim, int nzp, int has_2d)
{
    ARG_TRY(out && ctx, "null arg");
    ARG_TRY(dim >= 1 && dim <= 3, "dim must be 1..3");
    ARG_TRY(!has_2d || nzp >= 1, "nzp must be >= 1 for a field with 2D layout");
    qpg_field f = new qpg_field_s();
    f->ctx = ctx; f->dim = dim; f->nzp = nzp; f->has2d = has_2d;
    f->n1 = (size_t)(ctx->nr + 2) * ctx->P * dim;
    f->f2 = nullptr;
    CUDA_TRY(cudaMalloc(&f->f1, f->n1 * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(f->f1, 0, f->n1 * sizeof(double), ctx->stream));
    if (has_2d) {
        CUDA_TRY(cudaMalloc(&f->f2, f->n1 * (size_t)(nzp + 1) * sizeof(double)));
        CUDA_TRY(cudaMemsetAsync(f->f2, 0, f->n1 * (size_t)(nzp + 1) * sizeof(double), ctx->stream));
    }
    *out = f;
    return 0;
}
extern "C" int qpg_field_destroy(qpg_field f)
{
    if (!f) return 0;
    cudaStreamSynchronize(f->ctx->stream);
    cudaFree(f->f1); cudaFree(f->f2);
    delete f;
    return 0;
}
extern "C" int qpg_field_dim(qpg_field f) { return f ? f->dim : QPG_ERR_ARG; }

// fills of any value go through the program kernel only when non-zero; zero uses memset nodes
__global__ void k_fill(double *a, size_t n, double v) { for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) a[k] = v; }
__global__ void k_axpy(const double *a, double *b, size_t n) { for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) b[k] = b[k] + a[k]; }

extern "C" int qpg_field_fill(qpg_field f, double value)
{
    ARG_TRY(f, "null field");
    TprofScope tp(f->ctx, TP_ARITH);
    if (value == 0.0) CUDA_TRY(cudaMemsetAsync(f->f1, 0, f->n1 * sizeof(double), f->ctx->stream));
    else { k_fill<<<8, 256, 0, f->ctx->stream>>>(f->f1, f->n1, value); count_launch(f->ctx); CUDA_TRY(cudaGetLastError()); }
    return 0;
}
extern "C" int qpg_field_fill_f2(qpg_field f, double value)
{
    ARG_TRY(f && f->has2d, "field has no 2D layout");
    int rc = qpg_field_fill(f, value);
    if (rc) return rc;
    size_t n = f->n1 * (size_t)(f->nzp + 1);
    if (value == 0.0) CUDA_TRY(cudaMemsetAsync(f->f2, 0, n * sizeof(double), f->ctx->stream));
    else { k_fill<<<592, 256, 0, f->ctx->stream>>>(f->f2, n, value); count_launch(f->ctx); CUDA_TRY(cudaGetLastError()); }
    return 0;
}
extern "C" int qpg_field_copy(qpg_field src, qpg_field dst)
{
    ARG_TRY(src && dst && src->dim == dst->dim && src->ctx == dst->ctx, "fields do not match");
    CUDA_TRY(cudaMemcpyAsync(dst->f1, src->f1, src->n1 * sizeof(double), cudaMemcpyDeviceToDevice, src->ctx->stream));
    return 0;
}
extern "C" int qpg_field_copy_slice(qpg_field f, int idx, int dir)
{
    ARG_TRY(f && f->has2d, "The field has no 2D layout.");
    ARG_TRY(idx >= 1 && idx <= f->nzp + 1, "slice index out of range");
    double *s2 = f->f2 + (size_t)(idx - 1) * f->n1;
    if (dir == QPG_COPY_1TO2) CUDA_TRY(cudaMemcpyAsync(s2, f->f1, f->n1 * sizeof(double), cudaMemcpyDeviceToDevice, f->ctx->stream));
    else if (dir == QPG_COPY_2TO1) CUDA_TRY(cudaMemcpyAsync(f->f1, s2, f->n1 * sizeof(double), cudaMemcpyDeviceToDevice, f->ctx->stream));
    else { qpg_set_error("invalid copy direction"); return QPG_ERR_ARG; }
    return 0;
}
#define ONE_OP_PROLOGUE(ctxexpr) FProgBuilder pb(ctxexpr)
extern "C" int qpg_field_add(qpg_field a, qpg_field b)
{
    ARG_TRY(a && b && a->dim == b->dim, "guard cells / dims not matched!");
    ONE_OP_PROLOGUE(a->ctx);
    FOp &o = pb.add(FOP_ADD); o.a = a->f1; o.b = b->f1; o.da = a->dim; o.db = b->dim;
    return pb.launch(TP_ARITH);
}
extern "C" int qpg_field_add3(qpg_field a1, qpg_field a2, qpg_field a3)
{
    ARG_TRY(a1 && a2 && a3 && a1->dim == a2->dim && a1->dim == a3->dim, "dims not matched!");
    ONE_OP_PROLOGUE(a1->ctx);
    FOp &o = pb.add(FOP_ADD3); o.a = a1->f1; o.b = a2->f1; o.c = a3->f1; o.da = a1->dim;
    return pb.launch(TP_ARITH);
}
extern "C" int qpg_field_add_dim(qpg_field a, qpg_field b, int ndim, const int *adim, const int *bdim)
{
    ARG_TRY(a && b && adim && bdim && ndim >= 1 && ndim <= 3, "bad arguments");
    ONE_OP_PROLOGUE(a->ctx);
    for (int k = 0; k < ndim; k++) {
        ARG_TRY(adim[k] >= 1 && adim[k] <= a->dim && bdim[k] >= 1 && bdim[k] <= b->dim, "component out of range");
        FOp &o = pb.add(FOP_ADD_DIM); o.a = a->f1; o.b = b->f1; o.da = a->dim; o.db = b->dim; o.i0 = adim[k] - 1; o.i1 = bdim[k] - 1;
    }
    return pb.launch(TP_ARITH);
}
extern "C" int qpg_field_add_f2(qpg_field a, qpg_field b)
{
    ARG_TRY(a && b && a->has2d && b->has2d && a->dim == b->dim && a->nzp == b->nzp, "fields do not match");
    size_t n = a->n1 * (size_t)(a->nzp + 1);
    TprofScope tp(a->ctx, TP_ARITH);
    k_axpy<<<592, 256, 0, a->ctx->stream>>>(a->f2, b->f2, n);
    count_launch(a->ctx);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_field_scale(qpg_field f, double s)
{
    ARG_TRY(f, "null field");
    ONE_OP_PROLOGUE(f->ctx);
    FOp &o = pb.add(FOP_SCALE); o.a = f->f1; o.da = f->dim; o.s0 = s;
    return pb.launch(TP_ARITH);
}
extern "C" int qpg_field_smooth(qpg_field f, int order, int kind)
{
    ARG_TRY(f && order >= 0 && kind >= 0 && kind <= 2, "bad arguments");
    qpg_ctx c = f->ctx;
    const size_t need = (size_t)c->nr * c->P * f->dim;
    double *tmp = nullptr;
    if (need * sizeof(double) > (size_t)c->smem_field) {
        if (c->scratch_n < need) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(c->scratch); c->scratch = nullptr; c->scratch_n = 0;
            CUDA_TRY(cudaMalloc(&c->scratch, sizeof(double) * need));
            c->scratch_n = need;
        }
        tmp = c->scratch;
    }
    for (int k = 0; k < order; k++) {
        ONE_OP_PROLOGUE(f->ctx);
        FOp &o = pb.add(FOP_SMOOTH); o.a = f->f1; o.da = f->dim; o.i0 = kind; o.b = tmp;
        int rc = pb.launch(TP_ARITH);
        if (rc) return rc;
    }
    return 0;
}

// host <-> device in the reference layout goes through the wire (pack/unpack) form
static int field_pack_image(qpg_field f, const double *image, double *wire)
{
    ONE_OP_PROLOGUE(f->ctx);
    FOp &o = pb.add(FOP_PACK); o.a = const_cast<double *>(image); o.b = wire; o.da = f->dim;
    return pb.launch(TP_PIPELINE);
}
static int field_unpack_image(qpg_field f, double *image, const double *wire, int add)
{
    ONE_OP_PROLOGUE(f->ctx);
    FOp &o = pb.add(FOP_UNPACK); o.a = image; o.b = const_cast<double *>(wire); o.da = f->dim; o.i0 = add;
    return pb.launch(TP_PIPELINE);
}
__global__ void k_lineout(const double *f2, size_t n1, int nzp, size_t off, double *out)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nzp; k += gridDim.x * blockDim.x) out[k] = f2[(size_t)k * n1 + off];
}
extern "C" int qpg_field_lineout(qpg_field f, int comp, int plane, int node, double *host)
{
    ARG_TRY(f && host && f->has2d, "null arg / no 2D layout");
    ARG_TRY(comp >= 1 && comp <= f->dim && plane >= 0 && plane < f->ctx->P && node >= 0 && node <= f->ctx->nr + 1, "index out of range");
    double *tmp;
    CUDA_TRY(cudaMalloc(&tmp, sizeof(double) * f->nzp));
    size_t off = ((size_t)node * f->ctx->P + plane) * f->dim + (comp - 1);
    k_lineout<<<8, 256, 0, f->ctx->stream>>>(f->f2, f->n1, f->nzp, off, tmp);
    count_launch(f->ctx);
    cudaError_t e = cudaMemcpyAsync(host, tmp, sizeof(double) * f->nzp, cudaMemcpyDeviceToHost, f->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(f->ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return qpg_cuda_fail(e, "qpg_field_lineout");
    return 0;
}
extern "C" long qpg_field_wire_count(qpg_field f) { return f ? (long)f->n1 : -1; }
extern "C" int qpg_field_pack(qpg_field f, int slice, double *dev_buf)
{
    ARG_TRY(f && dev_buf, "null arg");
    ARG_TRY(slice == 0 || (f->has2d && slice >= 1 && slice <= f->nzp + 1), "slice out of range");
    return field_pack_image(f, slice == 0 ? f->f1 : f->f2 + (size_t)(slice - 1) * f->n1, dev_buf);
}
extern "C" int qpg_field_unpack(qpg_field f, int slice, const double *dev_buf, int add)
{
    ARG_TRY(f && dev_buf, "null arg");
    ARG_TRY(slice == 0 || (f->has2d && slice >= 1 && slice <= f->nzp + 1), "slice out of range");
    return field_unpack_image(f, slice == 0 ? f->f1 : f->f2 + (size_t)(slice - 1) * f->n1, dev_buf, add);
}
static int field_xfer(qpg_field f, double *host, int nslices, double *image0, bool upload)
{
    double *tmp;
    CUDA_TRY(cudaMalloc(&tmp, f->n1 * sizeof(double)));
    int rc = 0;
    const int P = f->ctx->P, nr2 = f->ctx->nr + 2, dim = f->dim;
    // host layout [P][nslices][nr+2][dim]; wire layout per slice [P][nr+2][dim]
    for (int k = 0; k < nslices && !rc; k++) {
        double *img = image0 + (size_t)k * f->n1;
        if (upload) {
            for (int pl = 0; pl < P; pl++)
                if (cudaMemcpyAsync(tmp + (size_t)pl * nr2 * dim, host + ((size_t)pl * nslices + k) * nr2 * dim, sizeof(double) * nr2 * dim, cudaMemcpyHostToDevice, f->ctx->stream) != cudaSuccess) rc = QPG_ERR_CUDA;
            if (!rc) rc = field_unpack_image(f, img, tmp, 0);
        } else {
            rc = field_pack_image(f, img, tmp);
            for (int pl = 0; pl < P && !rc; pl++)
                if (cudaMemcpyAsync(host + ((size_t)pl * nslices + k) * nr2 * dim, tmp + (size_t)pl * nr2 * dim, sizeof(double) * nr2 * dim, cudaMemcpyDeviceToHost, f->ctx->stream) != cudaSuccess) rc = QPG_ERR_CUDA;
        }
        if (cudaStreamSynchronize(f->ctx->stream) != cudaSuccess) rc = QPG_ERR_CUDA;
    }
    cudaFree(tmp);
    if (rc == QPG_ERR_CUDA) qpg_set_error("field transfer failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}
extern "C" int qpg_field_upload_f1(qpg_field f, const double *host) { ARG_TRY(f && host, "null arg"); return field_xfer(f, const_cast<double *>(host), 1, f->f1, true); }
extern "C" int qpg_field_download_f1(qpg_field f, double *host) { ARG_TRY(f && host, "null arg"); return field_xfer(f, host, 1, f->f1, false); }
extern "C" int qpg_field_upload_f2(qpg_field f, const double *host) { ARG_TRY(f && host && f->has2d, "null arg / no 2D layout"); return field_xfer(f, const_cast<double *>(host), f->nzp + 1, f->f2, true); }
extern "C" int qpg_field_download_f2(qpg_field f, double *host) { ARG_TRY(f && host && f->has2d, "null arg / no 2D layout"); return field_xfer(f, host, f->nzp + 1, f->f2, false); }

// ------------------------------------------------------------------------------------------------
// solves (one program op each)
// ------------------------------------------------------------------------------------------------
#define CHK_F(f, d) ARG_TRY((f) && (f)->dim == (d) && (f)->ctx == ctx, "field handle has wrong dim or context")
extern "C" int qpg_solve_psi(qpg_ctx ctx, qpg_field q, qpg_field psi)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(q, 1); CHK_F(psi, 1);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_PSI); o.a = q->f1; o.b = psi->f1; o.da = 1;
    return pb.launch(TP_SOLVE_PSI);
}
extern "C" int qpg_solve_bt(qpg_ctx ctx, qpg_field qb, qpg_field b)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(qb, 1); CHK_F(b, 3);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_BT); o.a = qb->f1; o.b = b->f1; o.da = 1;
    return pb.launch(TP_SOLVE_BBT);
}
extern "C" int qpg_solve_bz(qpg_ctx ctx, qpg_field cu, qpg_field b)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(cu, 3); CHK_F(b, 3);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_BZ); o.a = cu->f1; o.b = b->f1; o.da = 3;
    return pb.launch(TP_SOLVE_BZ);
}
extern "C" int qpg_solve_bt_iter(qpg_ctx ctx, qpg_field dcu, qpg_field cu, qpg_field b)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(dcu, 2); CHK_F(cu, 3); CHK_F(b, 3);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_BTITER); o.a = dcu->f1; o.b = cu->f1; o.c = b->f1; o.da = 2; o.s0 = ctx->relax;
    return pb.launch(TP_SOLVE_PBT);
}
extern "C" int qpg_solve_ez(qpg_ctx ctx, qpg_field cu, qpg_field e)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(cu, 3); CHK_F(e, 3);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_EZ); o.a = cu->f1; o.b = e->f1; o.da = 3;
    return pb.launch(TP_SOLVE_EZ);
}
extern "C" int qpg_solve_et(qpg_ctx ctx, qpg_field b, qpg_field psi, qpg_field e)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(b, 3); CHK_F(psi, 1); CHK_F(e, 3);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_ET); o.a = b->f1; o.b = psi->f1; o.c = e->f1; o.da = 3;
    return pb.launch(TP_SOLVE_PET);
}
extern "C" int qpg_solve_et_beam(qpg_ctx ctx, qpg_field b, qpg_field e)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(b, 3); CHK_F(e, 3);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_ETBEAM); o.a = b->f1; o.b = e->f1; o.da = 3;
    return pb.launch(TP_SOLVE_BET);
}
extern "C" int qpg_solve_djdxi(qpg_ctx ctx, qpg_field acu, qpg_field amu, qpg_field dcu)
{
    ARG_TRY(ctx, "null ctx"); CHK_F(acu, 2); CHK_F(amu, 3); CHK_F(dcu, 2);
    FProgBuilder pb(ctx);
    FOp &o = pb.add(FOP_DJDXI); o.a = acu->f1; o.b = amu->f1; o.c = dcu->f1; o.da = 2;
    return pb.launch(TP_SET_SOURCE);
}
extern "C" int qpg_bperp_residual(qpg_ctx ctx, qpg_field fld, int dim, int op, double *rel_res, double *abs_res)
{
    ARG_TRY(ctx && fld && fld->ctx == ctx, "bad field handle");
    ARG_TRY(dim >= 1 && dim <= fld->dim, "component out of range");
    FProgBuilder pb(ctx);
    if (op == QPG_CONV_RECORD) {
        FOp &o = pb.add(FOP_CONV_RECORD); o.a = fld->f1; o.da = fld->dim; o.i0 = dim - 1;
        return pb.launch(TP_ARITH);
    }
    ARG_TRY(op == QPG_CONV_COMPARE, "Invalid operation mode!");
    ARG_TRY(rel_res && abs_res, "Parameter 'rel_res' and 'abs_res' must be given for 'compare' operation.");
    FOp &o = pb.add(FOP_CONV_COMPARE); o.a = fld->f1; o.da = fld->dim; o.i0 = dim - 1; o.i1 = 0;
    int rc = pb.launch(TP_ARITH);
    if (rc) return rc;
    double out[2];
    CUDA_TRY(cudaMemcpyAsync(out, ctx->conv_out, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *rel_res = out[0]; *abs_res = out[1];
    return 0;
}

// common.cuh -- shared declarations of libqpadb200 (sm_100a only).
// Device data layouts:
//   particles : SoA fp64 planes x1,x2,p1,p2,p3,gamma,psi,q of length npmax; live count `npp` is DEVICE resident
//               (update_bound changes it without a host round trip); the host keeps an upper bound `npp_hi`.
//   field f1  : node-interleaved  f1[(j*P + pl)*dim + c],  j = 0..nr+1 (guard, axis .. guard), pl = plane,
//               c = component.  One radial node's values for all azimuthal planes are contiguous, so a particle
//               gather / a stencil row touches one or two cache lines.
//   field f2  : slice-major stack of f1 images, slice k (1-based, nzp+1 slices incl. guard) at (k-1)*n1.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include "../../include/qpad_b200.h"

#define QPG_MAX_MODE 4

void qpg_set_error(const char *fmt, ...);
int qpg_cuda_fail(cudaError_t e, const char *what);
#define CUDA_TRY(expr)                                             \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return qpg_cuda_fail(_e, #expr);    \
    } while (0)
#define ARG_TRY(cond, msg)                                         \
    do {                                                           \
        if (!(cond)) { qpg_set_error("%s: %s", __func__, msg); return QPG_ERR_ARG; } \
    } while (0)

// solver kinds (param.f03 p_fk_*)
enum { FK_PSI = 0, FK_EZ = 1, FK_BZ = 2, FK_BT = 3, FK_BPLUS = 4, FK_BMINUS = 5, FK_NKIND = 6 };

// Semiseparable (Green's function) factors of one tridiagonal operator, stored lane-major transposed for the
// one-warp-per-system scan: coefT[k*32 + lane] = coef[lane*C + k].  x_i = p_i*sum_{j<=i} q_j d_j + u_i*sum_{j>i} v_j d_j.
struct OpCoef {
    const double *qT, *vT, *pT, *uT;
    double axis_inv;  // != 0: axis row decoupled (0,1,0)/dr^2 -> x_1 = d_1*axis_inv
};

// tprof event ids (names follow sysutil_module.f03:298-334)
enum {
    TP_DEPOSIT2D = 0, TP_PUSH2D, TP_MOVE2D, TP_SORT2D, TP_SOLVE_PSI, TP_SOLVE_BZ, TP_SOLVE_EZ, TP_SOLVE_PBT, TP_SOLVE_BBT,
    TP_SOLVE_PET, TP_SOLVE_BET, TP_SET_SOURCE, TP_ARITH, TP_PIPELINE, TP_DEPOSIT3D, TP_PUSH3D, TP_MOVE3D, TP_FIELD_FUSED,
    TP_K_QDEP, TP_K_AMJ, TP_K_PUSH, TP_K_COMPACT, TP_K_SWEEP,
    TP_COUNT
};
extern const char *const qpg_tprof_names[TP_COUNT];

struct qpg_ctx_s {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int nr, M, P;
    double dr, dxi, relax;
    int bnd;
    int logC, C;  // scan chunk per lane
    OpCoef ops[FK_NKIND][QPG_MAX_MODE + 1];
    double *coef_pool;     // device pool backing all OpCoef arrays
    double *conv_old;      // [2][nr+2] record of sum|B| (re, im)
    double *conv_out;      // [0]=rel [1]=abs (device)
    int *flags;            // device ints: [0] done [2] iteration in slice [3] current slice j [4] slices executed [6] sweep kernel aborted (sticky) [7] neutral particle set overflow (sticky)
    long long *counters;   // device: [0] particle-slice updates [1] PC iterations
    unsigned long long cond_handle;  // CUDA-graph WHILE handle while capturing the PC-loop body (else 0)
    long launches;
    bool tprof_on;
    double tp_ms[TP_COUNT];
    long tp_calls[TP_COUNT];
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> tp_pending;
    std::vector<cudaEvent_t> ev_pool;
    int smem_field;        // dynamic smem bytes for the field kernel
    double *scratch; size_t scratch_n;   // lazily allocated global temporary (smooth_f1 of images that do not fit the shared-memory scratch)
    bool capturing;        // inside stream capture: no event timing
};

struct qpg_field_s {
    qpg_ctx ctx;
    int dim, nzp, has2d;
    size_t n1;  // doubles in one f1 image
    double *f1, *f2;
};

struct qpg_part2d_s {
    qpg_ctx ctx;
    double qbm;
    long npmax;
    long npp_hi;  // host upper bound of the live count
    double *x1, *x2, *p1, *p2, *p3, *gamma, *psi, *q;  // SoA planes (one slab allocation)
    double *slab;
    double *alt;                 // second slab for sort scatter (lazy)
    double *snap; long snap_np;  // device copy of the injected lattice for renew
    int *d_npp;                  // live count (device)
    int *d_nout;                 // particles flagged out of bounds
    unsigned *outmask;           // 1 bit per particle
    int *lists;                  // compaction lists (2*npmax)
    double *acc1;                // raw charge accumulators [(nr+2)][P]
    double *acc8;                // raw cu(3) dcu(2) amu(3) accumulators [(nr+2)][P][8]
    int *sort_keys, *sort_pos, *sort_hist; long sort_tiles;
};

struct qpg_part3d_s {
    qpg_ctx ctx;
    double qbm, dt;
    long npmax, npp_hi;
    int nz_total, noff2, nzp;
    double *x1, *x2, *x3, *p1, *p2, *p3, *q, *slab;
    int *d_npp, *d_nout;     // d_npp[2] = wire-buffer overflow flag (more particles crossed the slab edge than wire_cap)
    unsigned *outmask;
    unsigned *pushed;        // bitmap of the particles the interior pass of the split push has advanced (qpg_part3d_push_interior / _edge)
    int *lists;
    long wire_cap;           // particles per forward hand-off message (0 = default 0.1 npmax, part3d_class.f03:127)
    double *s1, *s2, *s3;    // spin vector planes, null unless qpg_part3d_enable_spin (has_spin, part3d_class.f03:53-63)
    double amm;              // anomalous magnetic moment of the spin push
};

// ---- launch bookkeeping -------------------------------------------------------------------
struct TprofScope {
    qpg_ctx ctx; int ev; cudaEvent_t a, b; bool on;
    TprofScope(qpg_ctx c, int e);
    ~TprofScope();
};
int qpg_tprof_flush(qpg_ctx ctx);
static inline void count_launch(qpg_ctx c, int n = 1) { c->launches += n; }

// ---- field program (fields.cu) ------------------------------------------------------------
enum {
    FOP_NOP = 0, FOP_ZERO, FOP_COPY, FOP_ADD, FOP_ADD3, FOP_ADD_DIM, FOP_SCALE, FOP_SLICE_1TO2, FOP_SLICE_2TO1,
    FOP_QFIX, FOP_AMJFIX, FOP_PSI, FOP_BT, FOP_BZ, FOP_BTITER, FOP_EZ, FOP_ET, FOP_ETBEAM, FOP_DJDXI,
    FOP_CONV_RECORD, FOP_CONV_COMPARE, FOP_SMOOTH, FOP_PACK, FOP_UNPACK, FOP_ZERO_F2, FOP_ADD_F2, FOP_SET_FLAG,
    FOP_PC_BEGIN
};
enum { FOPF_SKIP_IF_DONE = 1 };
struct FOp {
    int code, flags;
    int i0, i1, i2, i3;
    double s0, s1;
    double *a, *b, *c, *d;
    int da, db, dc, dd;  // dims of a..d
};
#define QPG_MAX_FOPS 40
struct FProg {
    int nops;
    int nr, M, P, logC;
    double dr;
    const OpCoef *ops;   // device copy of ctx->ops, [FK_NKIND][QPG_MAX_MODE+1]
    double *conv_old, *conv_out;
    int *flags;
    long long *counters;
    unsigned long long cond_handle;
    FOp op[QPG_MAX_FOPS];
};
struct FProgBuilder {
    qpg_ctx ctx;
    FProg prog;
    explicit FProgBuilder(qpg_ctx c);
    FOp &add(int code);
    int launch(int tp_event);
};
OpCoef *qpg_ctx_dev_ops(qpg_ctx ctx);

// helpers used across translation units
int part2d_epilogue_q(qpg_part2d p, qpg_field q);
long qpg_neutral_max_new_per_update(qpg_neutral ne);   // neutral.cu
int qpg_ctx_check_latches(qpg_ctx c, const int *host_flags);   // fields.cu: flags[6] sweep watchdog, flags[7] neutral overflow

// sweep.cu -- the slab sweep as ONE persistent cooperative kernel.
//
// Why: a slice of the quasi-static loop (simulation_class.f03:342-469) is a strictly sequential chain of small
// steps (deposit -> field solves -> {deposit -> field solves}* -> push) whose device time at C2 is ~5-12 us each; as
// separate kernel launches -- even replayed from a CUDA graph with a device-side WHILE node -- the chain costs
// ~94 us per slice, one third of it launch / dependency latency.  Here one CTA per SM (of the kernel's SM share)
// stays resident for the whole slab; the phases of a slice are separated by a hand-rolled grid barrier (one L2 RED +
// one polled line, ~1.3 us) and the predictor-corrector loop is an ordinary loop whose exit every CTA derives from
// the residual maxima the field team publishes.
//
//   per slice j:   [A on the field team  ||  update_bound compaction on the last CTA]          -- barrier
//                  { amjdeposit on all CTAs -- barrier -- C on the field team -- barrier }  x n_it
//                  [push_u + push_x + bound flags + next slice's qdeposit on all CTAs  ||  D items, 1 warp per CTA] -- barrier
//
// The field team is the first ceil(nr/32) CTAs; each owns a strip of 32 radial nodes and works on it with all its
// threads (see "field team: strip decomposition" below).  Strip totals of the tridiagonal scans cross CTAs as
// self-validating flagged words, so the programs need no team barrier and no cluster co-scheduling, and several sweep
// kernels (one per xi slab, pipeline.LocalPipeline) can share a GPU.  Every barrier has a watchdog: a CTA that waits
// longer than ~2 s raises the abort flag, all CTAs leave, and the host reports QPG_ERR_STATE instead of hanging.
#include "common.cuh"

#ifndef SW_T
#define SW_T 512           // threads per CTA (16 warps, <= 128 registers per thread)
#endif
#define SW_XK 64           // slots of one exchange record; layout [slot][strip] so a warp reads one slot of all strips coalesced
#define SW_MAX_TEAM 128

struct SweepArgs {
    FusedArgs f;
    PartView pv;
    double *const *planes;
    int *d_npp_w, *d_nout;
    unsigned *outmask;
    int *lists;
    double qbm, edge;
    int j0, j1, nteam;
    unsigned *bar;          // [0] grid arrivals, [32] team arrivals, [64] abort flag (one 128-byte line each)
    double *xbuf;           // [3][SW_XK][SW_MAX_TEAM]: slab 2 = residual maxima of program C (read after a grid barrier)
    uint4 *xll;             // [2][SW_XK][SW_MAX_TEAM] flagged 16-byte exchange words of the strip scans (cleared before each launch)
    double *back_b, *back_e; // wire buffers [P][nr+2][3] (possibly in the upstream GPU's memory) that receive b and e of the slab's FIRST slice
    unsigned *back_flag;    // ... and the flag word raised once they are complete (null: no backward hand-off from this launch)
    unsigned back_seq;
    // ponderomotive-guiding-centre path (PGC = true): the laser envelope of the run (laser.cu)
    LaserView lv;           // slice images a_r, a_i (dim 1) and their gradients (dim 3) the pgc pushers gather from
    double *las_far, *las_fai, *las_fgr, *las_fgi;   // ... the same images, writable: refreshed for every slice in phase A
    const double *las_ar, *las_ai;                   // envelope volumes [plane][slice -1..nz+1][node]
    double *chi_acc, *chi1, *chi2;                   // raw susceptibility sums, slice image, volume
    double las_dz, chi_ax;
    int las_nz;
    const unsigned *las_progress;   // != null: an envelope advance runs beside this sweep (laser.cu); slice j may be read once *las_progress - las_base >= j
    unsigned las_base;
    long long *trace;       // per slice of the slab: [2*(j-1)] ns spent in slice j (globaltimer), [2*(j-1)+1] PC iterations it took
    long long *prof;        // [0..3] cycles in phase A / amj / C / push, [4] total cycles, [5] total ns, [6] slices, [7] amj phases, [8..11] CTA 0's own work cycles per phase (thread 0's arrival at the barrier)
};

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// spin until *ctr reaches target (wrap-safe); false on abort / watchdog
__device__ __forceinline__ bool spin_until(const unsigned *ctr, unsigned target, unsigned *abort_flag)
{
    long long t0 = 0;
    for (unsigned n = 1;; n++) {
        if ((int)(ld_volatile_u32(ctr) - target) >= 0) return true;
        if ((n & 63u) == 0) {
            if (ld_volatile_u32(abort_flag)) return false;
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) { atomicExch(abort_flag, 1u); return false; }
        }
    }
}

// all SW_T threads of every CTA.  Returns false when the sweep must be abandoned.
__device__ __forceinline__ bool grid_barrier(unsigned *bar, unsigned &epoch, int *sm_i)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        const bool ok = spin_until(bar, epoch, bar + 64);
        __threadfence();
        sm_i[46] = ok ? 1 : 0;
    }
    __syncthreads();
    return sm_i[46] != 0;
}

// ---- field team: strip decomposition ---------------------------------------------------------------------------
// Team CTA `rank` owns the ST_N = 32 radial nodes i0 = 32*rank+1 .. i0+31 and ALL its 512 threads work on them:
//   * source terms / epilogues: one thread per (node, plane, kind) item
//   * tridiagonal solves: one WARP per system, lane <-> node, so the in-strip scans are pure warp shuffles; the strip
//     totals of all systems go to the exchange record, and after ONE team barrier every warp folds the other strips'
//     totals with a masked butterfly
//   * post-processing: one thread per (node, plane)
// A thread executes a few hundred instructions per program instead of the ~3000 of the thread-per-node layout, which
// is what bounds these latency-critical phases (one warp per scheduler cannot hide its own dependency chains).
#define ST_N 32
#define ST_H 2             // halo nodes each side of the strip (centred + one-sided 3-point stencils)
#define ST_W (ST_N + 2 * ST_H)

struct Team {
    unsigned *ctr, *abort_flag;
    unsigned epoch;          // team-barrier arrivals expected so far
    unsigned xep;            // sequence number of the current strip exchange (flag value of its words)
    int n, rank, xpar;
    double *xbuf;
    uint4 *xll;
};
// Strip totals travel as self-validating 16-byte words {lo32, seq, hi32, seq} (each 8-byte half is written atomically,
// the protocol NCCL calls LL): a reader polls the word itself, so the exchange needs no barrier and no fence -- its
// latency is one store + one load through L2 instead of fence + atomic + poll + fence + load.
__device__ __forceinline__ void ll_store(uint4 *p, double v, unsigned seq)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)__double2loint(v)), "r"(seq), "r"((unsigned)__double2hiint(v)), "r"(seq)
                 : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4 *p, unsigned seq, double &v)
{
    unsigned a, b, c, d;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    v = __hiloint2double((int)c, (int)a);
    return b == seq && d == seq;
}
// all SW_T threads of the team CTAs
__device__ __forceinline__ void team_barrier(Team &tm)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        tm.epoch += (unsigned)tm.n;
        __threadfence();
        atomicAdd(tm.ctr, 1u);
        spin_until(tm.ctr, tm.epoch, tm.abort_flag);
        __threadfence();
    }
    __syncthreads();
}

#define SW_STAMP(k) do { if (stamp) { const long long _t = clock64(); stamp[k] += _t - tlast; tlast = _t; } } while (0)
template <int M> struct StripSmem {
    static constexpr int P = 2 * M + 1, NS = 4 * P;
    double d[NS][ST_N];            // right-hand sides, then solutions
    double cu[ST_W * P * 3];       // program C: deposit epilogue values of the strip + halo (node-interleaved like f1)
    double amu[ST_W * P * 3];
    double acu[ST_W * P * 2];
    double dcu[ST_N][P][2];
    double t[P][ST_N];             // |B_phi| partials of the convergence test
    // Green's-function factors of this strip's nodes, resident for the whole sweep (the L1 is invalidated by every
    // barrier fence, so re-reading them from global memory costs an L2 round trip per use): [0 = program A, 1 = C]
    double fq[2][NS][ST_N], fv[2][NS][ST_N], fp[2][NS][ST_N], fu[2][NS][ST_N];
    double fax[2][NS];             // axis_inv of the system's operator
    double ez_q0;                  // q[0] of the E_z m=0 operator (divergence correction)
    // program A, psi and beam-potential systems (s < 2P): factors of the halo nodes i0-1 and i0+32, the right halo's
    // source term, and the solutions of strip + halo (so the r-derivatives need no second exchange)
    double hl_p[2 * P], hl_u[2 * P], hr_q[2 * P], hr_v[2 * P], hr_p[2 * P], hr_u[2 * P], dh[2 * P];
    double psit[(ST_N + 2) * P], phit[(ST_N + 2) * P];
    long long stamps[16];          // stage clocks of CTA 0 (kept on chip; flushed to prof[16..] when the kernel ends)
};

template <int M> __device__ __forceinline__ int strip_kind(int prog, int s)
{
    constexpr int P = 2 * M + 1;
    const int g = s / P;
    if (prog == 0) return g == 0 ? FK_PSI : (g == 1 ? FK_BT : (g == 2 ? FK_BZ : FK_EZ));
    return g == 0 ? FK_BPLUS : (g == 1 ? FK_BMINUS : (g == 2 ? FK_BZ : FK_EZ));
}
// once per launch: operator factors -> shared memory (all threads of a team CTA)
template <int M>
__device__ void strip_load_factors(const FusedArgs &a, int rank, StripSmem<M> &sm)
{
    constexpr int P = 2 * M + 1, NS = 4 * P;
    const int i0 = rank * ST_N + 1, nr = a.nr;
    for (int it = threadIdx.x; it < 2 * NS * ST_N; it += SW_T) {
        const int ln = it % ST_N, s = (it / ST_N) % NS, prog = it / (ST_N * NS), t = i0 - 1 + ln;
        const OpCoef &oc = a.ops[strip_kind<M>(prog, s) * (QPG_MAX_MODE + 1) + (((s % P) + 1) >> 1)];
        const bool ok = t < nr;
        sm.fq[prog][s][ln] = ok ? oc.qT[t] : 0.0; sm.fv[prog][s][ln] = ok ? oc.vT[t] : 0.0;
        sm.fp[prog][s][ln] = ok ? oc.pT[t] : 0.0; sm.fu[prog][s][ln] = ok ? oc.uT[t] : 0.0;
        if (ln == 0) sm.fax[prog][s] = oc.axis_inv;
    }
    for (int s = threadIdx.x; s < 2 * P; s += SW_T) {
        const OpCoef &oc = a.ops[strip_kind<M>(0, s) * (QPG_MAX_MODE + 1) + (((s % P) + 1) >> 1)];
        const int tl = i0 - 2, tr = i0 + ST_N - 1;   // 0-based rows of the nodes i0-1 and i0+32
        sm.hl_p[s] = tl >= 0 ? oc.pT[tl] : 0.0; sm.hl_u[s] = tl >= 0 ? oc.uT[tl] : 0.0;
        const bool ok = tr < nr;
        sm.hr_q[s] = ok ? oc.qT[tr] : 0.0; sm.hr_v[s] = ok ? oc.vT[tr] : 0.0; sm.hr_p[s] = ok ? oc.pT[tr] : 0.0; sm.hr_u[s] = ok ? oc.uT[tr] : 0.0;
    }
    if (threadIdx.x == 0) sm.ez_q0 = a.ops[FK_EZ * (QPG_MAX_MODE + 1)].qT[0];
    __syncthreads();
}

// strip-local scans of one system (one warp): inclusive prefix of q*d and inclusive suffix of v*d; strip totals -> exchange record
__device__ __forceinline__ void strip_scan(double q, double v, double d, int lane, double &incl, double &sfx, uint4 *xrec, unsigned seq, int slot_a, int slot_b)
{
    incl = q * d;
    sfx = v * d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double ta = __shfl_up_sync(0xffffffffu, incl, o), tb = __shfl_down_sync(0xffffffffu, sfx, o);
        if (lane >= o) incl += ta;
        if (lane + o < 32) sfx += tb;
    }
    if (lane == 31) ll_store(xrec + (size_t)slot_a * SW_MAX_TEAM, incl, seq);
    if (lane == 0) ll_store(xrec + (size_t)slot_b * SW_MAX_TEAM, sfx, seq);
}
// totals of the strips before / after mine, and (slot_r >= 0) the sum of an all-strip reduction slot; every lane gets
// all.  Polls the flagged words until every strip has published this exchange.
__device__ __forceinline__ void strip_fold(const uint4 *xb, unsigned seq, unsigned *abort_flag, int nteam, int rank, int lane, int slot_a, int slot_b,
                                           int slot_r, double &pa, double &pb, double &tot)
{
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r0 = 0; r0 < nteam; r0 += 32) {
        const int r = r0 + lane;
        double va = 0.0, vb = 0.0, vc = 0.0;
        bool ok = r >= nteam;
        for (unsigned n = 1; !__all_sync(0xffffffffu, ok); n++) {
            if (!ok) {
                ok = ll_load(xb + (size_t)slot_a * SW_MAX_TEAM + r, seq, va) & ll_load(xb + (size_t)slot_b * SW_MAX_TEAM + r, seq, vb);
                if (slot_r >= 0) ok &= ll_load(xb + (size_t)slot_r * SW_MAX_TEAM + r, seq, vc);
            }
            if ((n & 1023u) == 0 && (ld_volatile_u32(abort_flag) || n > (1u << 22))) { atomicExch(abort_flag, 1u); break; }
        }
        if (r < rank) a += va;
        if (r > rank && r < nteam) b += vb;
        if (r < nteam) c += vc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
        if (slot_r >= 0) c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    pa = a; pb = b; tot = c;
}

// ============================================================================================================
// program A (simulation_class.f03:344-377): q_beam slice -> bt(beam), qdp epilogue, psi, bz, record, b, ez, et
// `halo`: derive psi / phi of the two halo nodes from the exchanged strip totals instead of a second team barrier
template <int M>
#ifdef QPG_SWEEP_NOINLINE
__noinline__
#endif
__device__ void sweep_field_A(const FusedArgs &a, int j, Team &tm, StripSmem<M> &sm, bool halo, long long *stamp)
{
    constexpr int P = 2 * M + 1, NS = 4 * P, NWARP = SW_T / 32, SPW = (NS + NWARP - 1) / NWARP;
    static_assert(2 * NS + 1 <= SW_XK, "exchange record too small");
    static_assert(ST_N * P <= SW_T, "one thread per (node, plane) in the post-processing stage");
    long long tlast = stamp ? clock64() : 0;
    const int nr = a.nr, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = tm.rank * ST_N + 1;
    const size_t n1 = (size_t)(nr + 2) * P;
    const double idr = 1.0 / a.dr, idrh = 0.5 * idr;
    SolveCtx sc; sc.nr = nr; sc.M = M; sc.P = P; sc.logC = 0; sc.stride = 1; sc.dr = a.dr; sc.idr = idr; sc.idrh = idrh;
    if (tm.rank == 0 && tid == 0) { a.flags[0] = 0; a.flags[2] = 0; }
    const int nlo = (tm.rank == 0) ? 0 : i0, nhi = (i0 + ST_N - 1 >= nr) ? nr + 1 : i0 + ST_N - 1;   // nodes stored by this strip (+ guards)
    // inputs of the post-processing stage that do not depend on the solves: fetched now, used after the exchange
    const int p_ln = tid % ST_N, p_pl = tid / ST_N, p_i = i0 + p_ln;
    const bool p_on = tid < ST_N * P && p_i <= nr;
    double p_bsr = 0.0, p_bsp = 0.0, p_bbz = 0.0;
    if (p_on) { p_bsr = FX(a.b_spe, 3, p_i, p_pl, 0); p_bsp = FX(a.b_spe, 3, p_i, p_pl, 1); p_bbz = FX(a.b_beam, 3, p_i, p_pl, 2); }
    // ---- S1: sources.  threads 0..127: qdp epilogue + beam charge slice (species2d qdp :198-204, copy_slice :344);
    //          threads 128..511: bz / ez right-hand sides from the predicted current
    if (tid < 128) {
        for (int it = tid; it < (nhi - nlo + 1) * P; it += 128) {
            const int n = nlo + it / P, pl = it % P;
            const size_t k = (size_t)n * P + pl;
            const double qb = a.q_beam2[(size_t)(j - 1) * n1 + k];
            const double sq = axis_fix_q(n, pl, a.acc1[k]);
            const double qs = sq + a.spe_qn[k];
            a.q_beam[k] = qb; a.spe_q[k] = sq; a.q_spe[k] = qs;
            if (n >= i0 && n <= nr && n < i0 + ST_N) { sm.d[pl][n - i0] = -1.0 * qs; sm.d[P + pl][n - i0] = -1.0 * qb; }
        }
        if (tid >= 96 && tid < 96 + 2 * P) {   // right halo node of the psi / phi systems
            const int s = tid - 96, pl = s % P, ih = i0 + ST_N;
            double v = 0.0;
            if (ih <= nr) {
                const size_t k = (size_t)ih * P + pl;
                v = s < P ? -1.0 * (axis_fix_q(ih, pl, a.acc1[k]) + a.spe_qn[k]) : -1.0 * a.q_beam2[(size_t)(j - 1) * n1 + k];
            }
            sm.dh[s] = v;
        }
    } else {
        for (int it = tid - 128; it < ST_N * P * 2; it += SW_T - 128) {
            const int ln = it % ST_N, pl = (it / ST_N) % P, kind = it / (ST_N * P), i = i0 + ln;
            double v = 0.0;
            if (i <= nr) v = kind == 0 ? rhs_bz(sc, a.cu, pl, i) : rhs_ez(sc, a.cu, pl, i);
            sm.d[(2 + kind) * P + pl][ln] = v;
        }
    }
    const double cu_e1 = FX(a.cu, 3, nr - 2, 0, 0), cu_e2 = FX(a.cu, 3, nr - 1, 0, 0);   // edge term of the E_z divergence sum
    __syncthreads();
    SW_STAMP(0);
    // ---- S2: strip scans, one warp per system
    const int i = i0 + lane, t = i - 1;
    const bool valid = i <= nr;
    tm.xep++;
    uint4 *xrec = tm.xll + (size_t)tm.xpar * SW_MAX_TEAM * SW_XK + tm.rank;   // slot k of this strip: xrec[k * SW_MAX_TEAM]
    double incl[SPW], sfx[SPW], dd[SPW];
#pragma unroll
    for (int q = 0; q < SPW; q++) {
        const int s = warp + q * NWARP;
        if (s < NS) {
            dd[q] = valid ? sm.d[s][lane] : 0.0;
            strip_scan(sm.fq[0][s][lane], sm.fv[0][s][lane], dd[q], lane, incl[q], sfx[q], xrec, tm.xep, s, NS + s);
            if (s == 3 * P) {   // E_z m=0 divergence sum, field_e_class.f03:189-197
                double r = (valid && i >= 2 && i <= nr - 2) ? dd[q] * (double)(i - 1) : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                if (lane == 0) ll_store(xrec + (size_t)(2 * NS) * SW_MAX_TEAM, r, tm.xep);
            }
        }
    }
    SW_STAMP(1);
    SW_STAMP(2);
    // ---- S3: fold the other strips (polling their flagged words), apply the Green's-function factors, store
    const uint4 *xb = tm.xll + (size_t)tm.xpar * SW_MAX_TEAM * SW_XK;
#pragma unroll
    for (int q = 0; q < SPW; q++) {
        const int s = warp + q * NWARP;
        if (s < NS) {
            const int kind = strip_kind<M>(0, s), pl = s % P;
            double pa, pb, tot;
            strip_fold(xb, tm.xep, tm.abort_flag, tm.n, tm.rank, lane, s, NS + s, s == 3 * P ? 2 * NS : -1, pa, pb, tot);
            const double fp = sm.fp[0][s][lane], fu = sm.fu[0][s][lane], fax = sm.fax[0][s];
            const double ex = __shfl_down_sync(0xffffffffu, sfx[q], 1);
            double x = fp * (incl[q] + pa) + fu * ((lane == 31 ? 0.0 : ex) + pb);
            if (t == 0 && fax != 0.0) x = dd[q] * fax;
            if (s == 3 * P) {   // row 1 of the E_z m=0 source is -8*(div - edge term) (:199-209); by linearity x += rhs1 * G(:,1)
                const double div = tot - idrh * (cu_e1 + cu_e2) * ((double)nr - 2.5);
                x = (-8.0 * div) * (fp * sm.ez_q0) + x;
            }
            if (pl > 0 && i == 1 && kind != FK_BT) x = 0.0;
            if (valid) {
                sm.d[s][lane] = x;
                if (kind == FK_PSI) { FX(a.psi, 1, i, pl, 0) = x; sm.psit[(lane + 1) * P + pl] = x; }
                else if (kind == FK_BT) { FX(a.phi, 1, i, pl, 0) = x; sm.phit[(lane + 1) * P + pl] = x; }
                else if (kind == FK_BZ) FX(a.b_spe, 3, i, pl, 2) = x;
                else FX(a.e, 3, i, pl, 2) = x;
            }
            if (halo && s < 2 * P) {   // x = p*S + u*T at the nodes i0-1 and i0+32 from the same prefix / suffix sums
                const double tot_a = __shfl_sync(0xffffffffu, incl[q], 31), tot_b = __shfl_sync(0xffffffffu, sfx[q], 0);
                double *tile = s < P ? sm.psit : sm.phit;
                if (lane == 0 && i0 > 1) tile[pl] = sm.hl_p[s] * pa + sm.hl_u[s] * (tot_b + pb);
                if (lane == 31 && i0 + ST_N <= nr) {
                    const double dhv = sm.dh[s];
                    tile[(ST_N + 1) * P + pl] = sm.hr_p[s] * (pa + tot_a + sm.hr_q[s] * dhv) + sm.hr_u[s] * (pb - sm.hr_v[s] * dhv);
                }
            }
        }
    }
    tm.xpar ^= 1;
    SW_STAMP(3);
    if (halo) __syncthreads();
    else team_barrier(tm);   // psi / phi of the neighbouring strips through global memory
    SW_STAMP(4);
    // raw charge sums are consumed: the neighbouring strip read our first node's sum in its S1, i.e. before it published
    // the totals some warp of this CTA has just seen
    for (int it = tid; it < (nhi - nlo + 1) * P; it += SW_T) a.acc1[(size_t)nlo * P + it] = 0.0;
    // ---- S4: beam B-perp from phi, b = b_spe + b_beam, E-perp, convergence 'record'   (one thread per node, plane)
    const double *psi_s = halo ? sm.psit - (ptrdiff_t)(i0 - 1) * P : a.psi, *phi_s = halo ? sm.phit - (ptrdiff_t)(i0 - 1) * P : a.phi;
    if (tid < ST_N * P) {
        const int ln = p_ln, pl = p_pl, ii = p_i, m = (pl + 1) >> 1;
        if (!p_on) sm.t[pl][ln] = 0.0;
        else {
            // field_b_class.f03:545-701 get_solution_bt
            double bphi, br = 0.0;
            if (ii == 1) bphi = (m == 1) ? -idr * FX(phi_s, 1, 2, pl, 0) : 0.0;
            else if (ii == nr) bphi = -idrh * (3.0 * FX(phi_s, 1, nr, pl, 0) - 4.0 * FX(phi_s, 1, nr - 1, pl, 0) + FX(phi_s, 1, nr - 2, pl, 0));
            else bphi = -idrh * (FX(phi_s, 1, ii + 1, pl, 0) - FX(phi_s, 1, ii - 1, pl, 0));
            if (m > 0) {
                const bool im = (pl & 1) == 0;
                const int po = im ? pl - 1 : pl + 1;
                const double sg = im ? 1.0 : -1.0;
                if (ii == 1) br = (m == 1) ? sg * idr * m * FX(phi_s, 1, 2, po, 0) : 0.0;
                else br = sg * (idr / (double)(ii - 1)) * m * FX(phi_s, 1, ii, po, 0);
            }
            sm.t[pl][ln] = fabs(p_bsp);                                                       // convergence_tester 'record' :548-558
            const double b0 = p_bsr + br, b1 = p_bsp + bphi;                                  // b = b_spe + b_beam :375
            const double b2 = sm.d[2 * P + pl][ln] + p_bbz;
            double er, ephi;
            et_node<M>(psi_s, nr, idr, pl, ii, b0, b1, er, ephi);                              // :377
            FX(a.b_beam, 3, ii, pl, 0) = br; FX(a.b_beam, 3, ii, pl, 1) = bphi;
            FX(a.b, 3, ii, pl, 0) = b0; FX(a.b, 3, ii, pl, 1) = b1; FX(a.b, 3, ii, pl, 2) = b2;
            FX(a.e, 3, ii, pl, 0) = er; FX(a.e, 3, ii, pl, 1) = ephi;
        }
    }
    __syncthreads();
    if (tid < ST_N && i0 + tid <= nr) {
        double sre = 0.0, sim = 0.0;
#pragma unroll
        for (int pl = 0; pl < P; pl++) { if (pl > 0 && (pl & 1) == 0) sim += sm.t[pl][tid]; else sre += sm.t[pl][tid]; }
        a.conv_old[i0 + tid] = sre; a.conv_old[nr + 2 + i0 + tid] = sim;
    }
    SW_STAMP(5);
}

// program C (:378-396 + :375-377): amjdp epilogue, djdxi, bt_iter, bz, compare, record, b, ez, et.
// The two per-CTA residual maxima go to the exchange buffer; sweep_conv_decide() combines them after the grid barrier.
template <int M>
#ifdef QPG_SWEEP_NOINLINE
__noinline__
#endif
__device__ void sweep_field_C(const FusedArgs &a, Team &tm, StripSmem<M> &sm, long long *stamp)
{
    constexpr int P = 2 * M + 1, NS = 4 * P, NWARP = SW_T / 32, SPW = (NS + NWARP - 1) / NWARP;
    long long tlast = stamp ? clock64() : 0;
    const int nr = a.nr, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = tm.rank * ST_N + 1;
    const double idr = 1.0 / a.dr, idrh = 0.5 * idr;
    SolveCtx sc; sc.nr = nr; sc.M = M; sc.P = P; sc.logC = 0; sc.stride = 1; sc.dr = a.dr; sc.idr = idr; sc.idrh = idrh;
    // inputs of the post-processing stage that do not depend on this pass: beam field, psi terms of E-perp, old record
    const int p_ln = tid % ST_N, p_pl = tid / ST_N, p_i = i0 + p_ln;
    const bool p_on = tid < ST_N * P && p_i <= nr;
    double p_bb0 = 0.0, p_bb1 = 0.0, p_bb2 = 0.0, p_er0 = 0.0, p_ephi0 = 0.0, p_ore = 0.0, p_oim = 0.0;
    if (p_on) {
        p_bb0 = FX(a.b_beam, 3, p_i, p_pl, 0); p_bb1 = FX(a.b_beam, 3, p_i, p_pl, 1); p_bb2 = FX(a.b_beam, 3, p_i, p_pl, 2);
        et_node<M>(a.psi, nr, idr, p_pl, p_i, 0.0, 0.0, p_er0, p_ephi0);   // E-perp is affine in B-perp: er = b_phi + er0, ephi = -b_r + ephi0
    }
    if (tid < ST_N && i0 + tid <= nr) { p_ore = a.conv_old[i0 + tid]; p_oim = a.conv_old[nr + 2 + i0 + tid]; }
    // ---- S1: deposit epilogue (part2d_class.f03:916-981) of the strip + halo into shared tiles; the owner also stores
    //          the species2d amjdp results (:250-276, single species).  acc8 is cleared in S4, after the team barrier,
    //          because the neighbouring strips read the halo nodes' raw sums here.
    const int tlo = i0 - ST_H;   // node of tile slot 0
    const int own_lo = (tm.rank == 0) ? 0 : i0, own_hi = (i0 + ST_N - 1 >= nr) ? nr + 1 : i0 + ST_N - 1;
    for (int it = tid; it < ST_W * P * 8; it += SW_T) {
        const int c = it & 7, pl = (it >> 3) % P, sl = (it >> 3) / P, n = tlo + sl;
        if (n < 0 || n > nr + 1) continue;
        const size_t np = (size_t)n * P + pl;
        const double v = axis_fix_amj(n, pl, c, a.acc8[np * 8 + c]);
        const int tp = sl * P + pl;
        const bool own = n >= own_lo && n <= own_hi;
        if (c < 3) { sm.cu[tp * 3 + c] = v; if (own) { a.spe_cu[np * 3 + c] = v; a.cu[np * 3 + c] = v; } }
        else if (c < 5) { sm.acu[tp * 2 + c - 3] = v; if (own) { a.spe_dcu[np * 2 + c - 3] = v; a.acu[np * 2 + c - 3] = v; } }
        else { sm.amu[tp * 3 + c - 5] = v; if (own) { a.spe_amu[np * 3 + c - 5] = v; a.amu[np * 3 + c - 5] = v; } }
    }
    __syncthreads();
    SW_STAMP(6);
    // tile pointers addressed with absolute node numbers through FX()
    const double *cu_t = sm.cu - (ptrdiff_t)tlo * P * 3, *amu_t = sm.amu - (ptrdiff_t)tlo * P * 3, *acu_t = sm.acu - (ptrdiff_t)tlo * P * 2;
    // ---- S2a: dcu = djdxi(acu, amu) (:390)
    for (int it = tid; it < ST_N * P * 2; it += SW_T) {
        const int c = it & 1, pl = (it >> 1) % P, ln = (it >> 1) / P, ii = i0 + ln;
        double v = 0.0;
        if (ii <= nr) { v = djdxi_node<M>(acu_t, amu_t, nr, idr, pl, c, ii); FX(a.dcu, 2, ii, pl, c) = v; }
        sm.dcu[ln][pl][c] = v;
    }
    __syncthreads();
    SW_STAMP(7);
    // ---- S2b: sources of bt_iter (:391), bz (:392), ez (:376 of the next pass / :415)
    const double relax_idr2 = a.relax * (idr * idr);
    for (int it = tid; it < ST_N * NS; it += SW_T) {
        const int ln = it % ST_N, s = it / ST_N, kind = s / P, pl = s % P, ii = i0 + ln;
        double v = 0.0;
        if (ii <= nr) {
            if (kind < 2) {
                double dcu[P][2];
#pragma unroll
                for (int q = 0; q < P; q++) { dcu[q][0] = sm.dcu[ln][q][0]; dcu[q][1] = sm.dcu[ln][q][1]; }
                v = rhs_bt_iter_own<M>(sc, dcu, cu_t, a.b_spe, relax_idr2, kind, pl, ii);
            } else v = kind == 2 ? rhs_bz(sc, cu_t, pl, ii) : rhs_ez(sc, cu_t, pl, ii);
        }
        sm.d[s][ln] = v;
    }
    __syncthreads();
    SW_STAMP(8);
    // ---- S3: strip scans -> exchange -> solutions
    const int i = i0 + lane, t = i - 1;
    const bool valid = i <= nr;
    tm.xep++;
    uint4 *xrec = tm.xll + (size_t)tm.xpar * SW_MAX_TEAM * SW_XK + tm.rank;   // slot k of this strip: xrec[k * SW_MAX_TEAM]
    double incl[SPW], sfx[SPW], dd[SPW];
#pragma unroll
    for (int q = 0; q < SPW; q++) {
        const int s = warp + q * NWARP;
        if (s < NS) {
            dd[q] = valid ? sm.d[s][lane] : 0.0;
            strip_scan(sm.fq[1][s][lane], sm.fv[1][s][lane], dd[q], lane, incl[q], sfx[q], xrec, tm.xep, s, NS + s);
            if (s == 3 * P) {
                double r = (valid && i >= 2 && i <= nr - 2) ? dd[q] * (double)(i - 1) : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
                // the last strip holds cu(nr-2), cu(nr-1) of this pass: it folds the edge term of the divergence sum in
                if (tm.rank == tm.n - 1) r -= idrh * (FX(cu_t, 3, nr - 2, 0, 0) + FX(cu_t, 3, nr - 1, 0, 0)) * ((double)nr - 2.5);
                if (lane == 0) ll_store(xrec + (size_t)(2 * NS) * SW_MAX_TEAM, r, tm.xep);
            }
        }
    }
    SW_STAMP(9);
    SW_STAMP(10);
    const uint4 *xb = tm.xll + (size_t)tm.xpar * SW_MAX_TEAM * SW_XK;
#pragma unroll
    for (int q = 0; q < SPW; q++) {
        const int s = warp + q * NWARP;
        if (s < NS) {
            double pa, pb, tot;
            strip_fold(xb, tm.xep, tm.abort_flag, tm.n, tm.rank, lane, s, NS + s, s == 3 * P ? 2 * NS : -1, pa, pb, tot);
            const double fp = sm.fp[1][s][lane], fu = sm.fu[1][s][lane], fax = sm.fax[1][s];
            const double ex = __shfl_down_sync(0xffffffffu, sfx[q], 1);
            double x = fp * (incl[q] + pa) + fu * ((lane == 31 ? 0.0 : ex) + pb);
            if (t == 0 && fax != 0.0) x = dd[q] * fax;
            if (s == 3 * P) x = (-8.0 * tot) * (fp * sm.ez_q0) + x;
            sm.d[s][lane] = valid ? x : 0.0;
        }
    }
    tm.xpar ^= 1;
    __syncthreads();
    SW_STAMP(11);
    // ---- S4: get_solution_bt_iter (field_b_class.f03:703-758), bz, ez axis rules; b = b_spe + b_beam; E-perp; compare
    if (tid < ST_N * P) {
        const int ln = p_ln, pl = p_pl, ii = p_i, m = (pl + 1) >> 1;
        if (!p_on) sm.t[pl][ln] = 0.0;
        else {
            double br, bp;
            if (m == 0) { br = (ii == 1) ? 0.0 : sm.d[0][ln]; bp = (ii == 1) ? 0.0 : sm.d[P][ln]; }
            else {
                const bool im = (pl & 1) == 0;
                const int po = im ? pl - 1 : pl + 1;
                br = 0.5 * (sm.d[pl][ln] + sm.d[P + pl][ln]);
                bp = im ? 0.5 * (-sm.d[po][ln] + sm.d[P + po][ln]) : 0.5 * (sm.d[po][ln] - sm.d[P + po][ln]);
                if (ii == 1 && m != 1) { br = 0.0; bp = 0.0; }
            }
            double bz = sm.d[2 * P + pl][ln], ez = sm.d[3 * P + pl][ln];
            if (pl > 0 && ii == 1) { bz = 0.0; ez = 0.0; }
            sm.t[pl][ln] = fabs(bp);
            const double b0 = br + p_bb0, b1 = bp + p_bb1, b2 = bz + p_bb2;
            const bool et_zero = ii == 1 && m != 1;                  // solve_field_et axis rows (field_e_class.f03:450-470)
            const double er = et_zero ? 0.0 : b1 + p_er0, ephi = et_zero ? 0.0 : -b0 + p_ephi0;
            FX(a.b_spe, 3, ii, pl, 0) = br; FX(a.b_spe, 3, ii, pl, 1) = bp; FX(a.b_spe, 3, ii, pl, 2) = bz;
            FX(a.b, 3, ii, pl, 0) = b0; FX(a.b, 3, ii, pl, 1) = b1; FX(a.b, 3, ii, pl, 2) = b2;
            FX(a.e, 3, ii, pl, 0) = er; FX(a.e, 3, ii, pl, 1) = ephi; FX(a.e, 3, ii, pl, 2) = ez;
        }
    }
    // clear the raw deposit sums of the strip (+ guards); every strip read its halo nodes' sums before it published its totals
    for (int it = tid; it < (own_hi - own_lo + 1) * P * 8; it += SW_T) a.acc8[(size_t)own_lo * P * 8 + it] = 0.0;
    __syncthreads();
    double mo = 0.0, mn = 0.0;
    if (tid < ST_N && i0 + tid <= nr) {
        double sre = 0.0, sim = 0.0;
#pragma unroll
        for (int pl = 0; pl < P; pl++) { if (pl > 0 && (pl & 1) == 0) sim += sm.t[pl][tid]; else sre += sm.t[pl][tid]; }
        mo = p_ore * p_ore + p_oim * p_oim;
        const double dre = p_ore - sre, dim = p_oim - sim;
        mn = dre * dre + dim * dim;
        a.conv_old[i0 + tid] = sre; a.conv_old[nr + 2 + i0 + tid] = sim;   // 'record' for the next pass (:373)
    }
    if (warp == 0) {   // strip maxima -> third exchange slab (read by every CTA after the grid barrier)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mo = fmax(mo, __shfl_xor_sync(0xffffffffu, mo, o)); mn = fmax(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
        if (lane == 0) {
            double *xm = tm.xbuf + (size_t)2 * SW_MAX_TEAM * SW_XK + tm.rank;
            __stcg(xm, mo); __stcg(xm + SW_MAX_TEAM, mn);
        }
    }
    SW_STAMP(12);
}

// simulation_class.f03:560-599 on the per-CTA maxima published by program C.  Called by one thread per CTA after the
// grid barrier; every CTA derives the same decision.  `it` = iterations done in this slice including this one.
__device__ __forceinline__ bool sweep_conv_decide(const SweepArgs &a, int it, bool writer, int lane)
{
    const double *xb = a.xbuf + (size_t)2 * SW_MAX_TEAM * SW_XK;
    double mo = 0.0, mn = 0.0;
    for (int r = lane; r < a.nteam; r += 32) { mo = fmax(mo, __ldcg(xb + r)); mn = fmax(mn, __ldcg(xb + SW_MAX_TEAM + r)); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mo = fmax(mo, __shfl_xor_sync(0xffffffffu, mo, o)); mn = fmax(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
    const double old_norm = sqrt(mo), abs_res = sqrt(mn);
    const double rel = old_norm > 2.220446049250313e-16 ? abs_res / old_norm : 1.7976931348623157e308;
    const bool fin = rel < a.f.reltol || abs_res < a.f.abstol || it >= a.f.iter_max;
    if (writer && lane == 0) {
        a.f.conv_out[0] = rel; a.f.conv_out[1] = abs_res;
        a.f.counters[1] += 1;
        a.f.flags[2] = it;
        if (fin) a.f.flags[0] = 1;
    }
    return fin;
}

// ============================================================================================================
// Backward hand-off of the xi-pipeline from inside the sweep (simulation_class.f03:460-467: e and b of the slab's first slice
// go to the upstream stage, which needs them as guard slice nzp+1 for its beam push).  One CTA that idles during phase A
// copies the two slice images into the wire layout [P][nr+2][3] -- the destination may be mapped peer memory -- and raises
// the flag the consumer's stream waits for (csrc/p2p.cu).  The upstream stage gets its guard slice one slice into the
// sweep instead of after a separate first-slice launch + pack kernels.
template <int M>
__device__ void sweep_publish_back(const SweepArgs &a)
{
    constexpr int P = 2 * M + 1;
    const int n = (a.f.nr + 2) * P * 3;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int c = k % 3, r = k / 3, j = r % (a.f.nr + 2), pl = r / (a.f.nr + 2);
        const size_t src = ((size_t)j * P + pl) * 3 + c;       // slice 1 = the first image of the f2 volume
        a.back_b[k] = __ldcg(a.f.b2 + src);
        a.back_e[k] = __ldcg(a.f.e2 + src);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(a.back_flag, a.back_seq);
}

// ---- particle phases of a slice: the warps of a CTA take the CTA's tiles (32 particles) round-robin ---------------------------
// QPG_SWEEP_NOINLINE: the phases are separate functions (own register allocation: the particle arithmetic fits 64 registers on its
// own, tools/ab_bench.py); QPG_SWEEP_PREFETCH: the next tile's particle planes are fetched while the current tile is worked on
// (one L2 round trip per tile saved at the price of 12 registers)
#ifdef QPG_SWEEP_NOINLINE
#define SW_PHASE_FN __device__ __noinline__
#else
#define SW_PHASE_FN __device__ __forceinline__
#endif
#ifdef QPG_SWEEP_ILP2
#undef QPG_SWEEP_PREFETCH
#define QPG_SWEEP_PREFETCH 0
#endif
#ifndef QPG_SWEEP_PREFETCH
#define QPG_SWEEP_PREFETCH 1
#endif
template <int M>
SW_PHASE_FN void sweep_amj_phase(const SweepArgs &a, int npp, int tile0, int tile1, double *dep_tiles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FusedArgs &f = a.f;
    const double idr = 1.0 / f.dr;
    double *tile = dep_tiles + warp * DepTile<M>::doubles;
#if QPG_SWEEP_PREFETCH
    PartRegs cur = part_load(a.pv, (tile0 + warp) * 32 + lane, tile0 + warp < tile1 ? npp : 0);
    for (int tl = tile0 + warp; tl < tile1; tl += SW_T / 32) {
        const int tn = tl + SW_T / 32;
        const PartRegs nxt = part_load(a.pv, tn * 32 + lane, tn < tile1 ? npp : 0);
        amj_core<M>(a.pv, cur, f.e, f.b, f.acc8, a.qbm, f.dxi, idr, npp, tl * 32 + lane, lane, tile);
        cur = nxt;
    }
#elif defined(QPG_SWEEP_ILP2)
    // two tiles per warp and pass: the arithmetic of tile A and tile B is independent, the scheduler interleaves the two dependency chains
    // (a particle's chain is long and serial and only 4 warps share a scheduler); the reductions follow one after the other
    constexpr int P = 2 * M + 1;
    for (int tl = tile0 + 2 * warp; tl < tile1; tl += 2 * (SW_T / 32)) {
        const int ia = tl * 32 + lane, ib = ia + 32, nb = tl + 1 < tile1 ? npp : 0;
        const PartRegs pa = part_load(a.pv, ia, npp), pb = part_load(a.pv, ib, nb);
        double al_a[2 * P], be_a[8], al_b[2 * P], be_b[8];
        int key_a, key_b;
        amj_math<M>(a.pv, pa, f.e, f.b, a.qbm, f.dxi, idr, npp, ia, al_a, be_a, key_a);
        amj_math<M>(a.pv, pb, f.e, f.b, a.qbm, f.dxi, idr, nb, ib, al_b, be_b, key_b);
        warp_deposit_mma<M>(al_a, be_a, key_a, f.acc8, tile, lane);
        if (tl + 1 < tile1) warp_deposit_mma<M>(al_b, be_b, key_b, f.acc8, tile, lane);
    }
#else
    for (int tl = tile0 + warp; tl < tile1; tl += SW_T / 32)
        amj_core<M>(a.pv, part_load(a.pv, tl * 32 + lane, npp), f.e, f.b, f.acc8, a.qbm, f.dxi, idr, npp, tl * 32 + lane, lane, tile);
#endif
}
template <int M>
SW_PHASE_FN void sweep_push_phase(const SweepArgs &a, int npp, int tile0, int tile1, double *dep_tiles)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef QPG_EXP_NO_PUSH_QDEP   // bottleneck experiment only (results are wrong): the push phase without the fused charge deposit
    FusedArgs f = a.f; f.acc1 = nullptr;
#else
    const FusedArgs &f = a.f;
#endif
    const double idr = 1.0 / f.dr;
    double *tile = dep_tiles + warp * DepTile<M>::doubles;
#if QPG_SWEEP_PREFETCH
    PartRegs cur = part_load(a.pv, (tile0 + warp) * 32 + lane, tile0 + warp < tile1 ? npp : 0);
    for (int tl = tile0 + warp; tl < tile1; tl += SW_T / 32) {
        const int tn = tl + SW_T / 32;
        const PartRegs nxt = part_load(a.pv, tn * 32 + lane, tn < tile1 ? npp : 0);
        push_core<M>(a.pv, cur, f.e, f.b, a.qbm, f.dxi, idr, a.edge, 7, a.outmask, a.d_nout, f.acc1, npp, tl * 32 + lane, lane, tile);
        cur = nxt;
    }
#elif defined(QPG_SWEEP_ILP2)
    for (int tl = tile0 + 2 * warp; tl < tile1; tl += 2 * (SW_T / 32)) {   // two tiles per warp and pass, see sweep_amj_phase
        const int ia = tl * 32 + lane, ib = ia + 32, nb = tl + 1 < tile1 ? npp : 0;
        const PartRegs pa = part_load(a.pv, ia, npp), pb = part_load(a.pv, ib, nb);
        double xa1, xa2, xb1, xb2;
        bool oa, ob;
        push_math<M>(a.pv, pa, f.e, f.b, a.qbm, f.dxi, idr, a.edge, 7, npp, ia, xa1, xa2, oa);
        push_math<M>(a.pv, pb, f.e, f.b, a.qbm, f.dxi, idr, a.edge, 7, nb, ib, xb1, xb2, ob);
        push_finish<M>(pa, xa1, xa2, oa, idr, 7, a.outmask, a.d_nout, f.acc1, npp, ia, lane, tile);
        if (tl + 1 < tile1) push_finish<M>(pb, xb1, xb2, ob, idr, 7, a.outmask, a.d_nout, f.acc1, nb, ib, lane, tile);
    }
#else
    for (int tl = tile0 + warp; tl < tile1; tl += SW_T / 32)
        push_core<M>(a.pv, part_load(a.pv, tl * 32 + lane, npp), f.e, f.b, a.qbm, f.dxi, idr, a.edge, 7, a.outmask, a.d_nout, f.acc1, npp, tl * 32 + lane, lane, tile);
#endif
}

// ---- laser hooks of the slice loop inside the sweep (simulation_class.f03:361-366, :401): one helper CTA --------------------------
// set_grad + copy_slice + gather of slice j (laser/field_laser_class.f03:637-750; the stand-alone version is k_laser_slice of laser.cu)
__device__ void sweep_laser_slice(const SweepArgs &a, int M, int j)
{
    const int nr = a.f.nr, nz = a.las_nz, P = 2 * M + 1, n = (nr + 2) * P;
    if (j < 1 || j > nz) return;
    const double *ar = a.las_ar, *ai = a.las_ai;
    const double idrh = 0.5 / a.f.dr, idzh = 0.5 / a.las_dz, dr = a.f.dr;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int i = k / P, pl = k % P, m = (pl + 1) / 2;
        const bool is_im = m > 0 && pl == 2 * m;
        const int other = m == 0 ? pl : (is_im ? pl - 1 : pl + 1);
        a.las_far[k] = __ldcg(ar + LVI(pl, i, j));
        a.las_fai[k] = __ldcg(ai + LVI(pl, i, j));
        double *gr = a.las_fgr + (size_t)k * 3, *gi = a.las_fgi + (size_t)k * 3;
        if (i >= 1 && i <= nr) {
            gr[2] = idzh * (3.0 * __ldcg(ar + LVI(pl, i, j)) - 4.0 * __ldcg(ar + LVI(pl, i, j - 1)) + __ldcg(ar + LVI(pl, i, j - 2)));
            gi[2] = idzh * (3.0 * __ldcg(ai + LVI(pl, i, j)) - 4.0 * __ldcg(ai + LVI(pl, i, j - 1)) + __ldcg(ai + LVI(pl, i, j - 2)));
        }
        if (i >= 2 && i <= nr) {
            gr[0] = idrh * (__ldcg(ar + LVI(pl, i + 1, j)) - __ldcg(ar + LVI(pl, i - 1, j)));
            gi[0] = idrh * (__ldcg(ai + LVI(pl, i + 1, j)) - __ldcg(ai + LVI(pl, i - 1, j)));
            if (m == 0) { gr[1] = 0.0; gi[1] = 0.0; }
            else {
                const double ir = 1.0 / ((double)(i - 1) * dr), sg = is_im ? 1.0 : -1.0;
                gr[1] = sg * ir * m * __ldcg(ar + LVI(other, i, j));
                gi[1] = sg * ir * m * __ldcg(ai + LVI(other, i, j));
            }
        }
        if (m == 0 && i == 1) { gr[0] = 0.0; gr[1] = 0.0; gi[0] = 0.0; gi[1] = 0.0; }
        if (m > 0 && i == nr + 1) {   // the reference's quirk, see k_laser_slice
            if (m % 2 == 1) {
                const double g1r = 2.0 * idrh * __ldcg(ar + LVI(pl, 2, j)), g1i = 2.0 * idrh * __ldcg(ai + LVI(pl, 2, j));
                const double o1r = 2.0 * idrh * __ldcg(ar + LVI(other, 2, j)), o1i = 2.0 * idrh * __ldcg(ai + LVI(other, 2, j));
                gr[0] = g1r; gi[0] = g1i;
                gr[1] = is_im ? m * o1r : -m * o1r;
                gi[1] = is_im ? m * o1i : -m * o1i;
            } else { gr[0] = 0.0; gr[1] = 0.0; gi[0] = 0.0; gi[1] = 0.0; }
        }
    }
}
// axis rules + 1/(j-1) of the susceptibility deposited during slice j's push phase, stored as slice j of the chi volume (k_chi_fix)
__device__ void sweep_chi_fix(const SweepArgs &a, int M, int j)
{
    const int nr = a.f.nr, P = 2 * M + 1, n = (nr + 2) * P;
    double *chi2_slice = (j >= 1 && j <= a.las_nz) ? a.chi2 + (size_t)(j - 1) * n : nullptr;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int jj = k / P, pl = k % P;
        double v = __ldcg(a.chi_acc + k);
        a.chi_acc[k] = 0.0;
        if (jj == 0) v = 0.0;
        else if (jj == 1) v = pl == 0 ? v * a.chi_ax : 0.0;
        else v = v * (1.0 / (double)(jj - 1));
        a.chi1[k] = v;
        if (chi2_slice) chi2_slice[k] = v;
    }
}
// pgc particle phases (robust_pgc pusher; one tile per warp and pass)
template <int M>
SW_PHASE_FN void sweep_amj_phase_pgc(const SweepArgs &a, int npp, int tile0, int tile1, double *dep_tiles)
{
    constexpr int P = 2 * M + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FusedArgs &f = a.f;
    const double idr = 1.0 / f.dr;
    double *tile = dep_tiles + warp * DepTile<M>::doubles;
    for (int tl = tile0 + warp; tl < tile1; tl += SW_T / 32) {
        double alpha[2 * P], beta[8];
        int key;
        amj_math_pgc<M, false>(a.pv, f.e, f.b, a.lv, a.qbm, f.dxi, idr, npp, tl * 32 + lane, alpha, beta, key);
        warp_deposit_mma<M>(alpha, beta, key, f.acc8, tile, lane);
    }
}
template <int M>
SW_PHASE_FN void sweep_push_phase_pgc(const SweepArgs &a, int npp, int tile0, int tile1, double *dep_tiles)
{
    constexpr int P = 2 * M + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FusedArgs &f = a.f;
    const double idr = 1.0 / f.dr;
    double *tile = dep_tiles + warp * DepTile<M>::doubles;
    for (int tl = tile0 + warp; tl < tile1; tl += SW_T / 32) {
        const int i = tl * 32 + lane;
        const bool valid = i < npp;
        // lasers%deposit_chi (simulation_class.f03:401, part2d_class.f03:361-430) with the converged psi, at the position before the push
        double alpha[2 * P];
        int key = -1;
        if (valid) {
            const double x1 = a.pv.x1[i], x2 = a.pv.x2[i];
            const double pos = __dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2))), idr);
            const double rpos = fast_rcp(pos) * idr;
            const double c0 = x1 * rpos, s0 = -x2 * rpos;
            const int nn = (int)floor(pos);
            const double fr = pos - (double)nn, w0 = 1.0 - fr, w1 = fr;
            double phr = -1.0 * a.qbm * a.pv.q[i] * fast_rcp(1.0 - a.qbm * a.pv.psi[i]), phi = 0.0;
            alpha[0] = w0 * phr; alpha[P] = w1 * phr;
#pragma unroll
            for (int m = 1; m <= M; m++) {
                const double t = phr * c0 - phi * s0;
                phi = phr * s0 + phi * c0;
                phr = t;
                alpha[2 * m - 1] = w0 * phr; alpha[P + 2 * m - 1] = w1 * phr;
                alpha[2 * m] = w0 * phi; alpha[P + 2 * m] = w1 * phi;
            }
            key = nn + 1;
        } else {
#pragma unroll
            for (int k = 0; k < 2 * P; k++) alpha[k] = 0.0;
        }
        warp_deposit_q_mma<M>(alpha, key, a.chi_acc, tile, lane);
        if (valid) push_u_pgc_math<M>(a.pv, f.e, f.b, a.lv, a.qbm, f.dxi, idr, i);          // push_u_robust_pgc :1967 (stores p, gamma)
        const PartRegs pr = part_load(a.pv, i, npp);                                       // the advanced momenta
        double xn1, xn2;
        bool out;
        push_math<M>(a.pv, pr, f.e, f.b, a.qbm, f.dxi, idr, a.edge, 6, npp, i, xn1, xn2, out);   // push_x + bound test
        push_finish<M>(pr, xn1, xn2, out, idr, 6, a.outmask, a.d_nout, f.acc1, npp, i, lane, tile);
    }
}

template <int M, bool PGC = false>
__global__ void __launch_bounds__(SW_T, 1) k_sweep(const __grid_constant__ SweepArgs a)
{
    constexpr int P = 2 * M + 1;
    __shared__ int sm_i[48];
    extern __shared__ double sm_dyn[];      // [StripSmem<M>][per-warp alpha/beta tiles of the DMMA deposits]
    StripSmem<M> &sm_f = *reinterpret_cast<StripSmem<M> *>(sm_dyn);
    double *dep_tiles = sm_dyn + (sizeof(StripSmem<M>) + 7) / 8;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const FusedArgs &f = a.f;
    unsigned gep = 0;
    Team tm;
    tm.ctr = a.bar + 32; tm.abort_flag = a.bar + 64; tm.epoch = 0; tm.xep = 0; tm.n = a.nteam; tm.rank = b; tm.xpar = 0; tm.xbuf = a.xbuf; tm.xll = a.xll;
    const bool in_team = b < a.nteam;
    const int las_cta = G - 2 >= a.nteam ? G - 2 : G - 1;   // PGC: the CTA that prepares the laser slice images (an idle one if there is one)
    // the halo shortcut of program A needs the one-sided stencil nodes nr-1, nr-2 inside the last strip's tile
    const bool halo = (a.f.nr - ((a.nteam - 1) * ST_N + 1)) >= 1;
    if (in_team) strip_load_factors<M>(f, b, sm_f);
    long long prof[4] = {0, 0, 0, 0}, work[4] = {0, 0, 0, 0}, tprev = 0, tstart = 0, nstart = 0, namj = 0;
    const bool timer = (b == 0 && tid == 0);
    if (timer) { for (int k = 0; k < 16; k++) sm_f.stamps[k] = 0; tstart = tprev = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(nstart)); }
    // the abort word is sticky (the host never clears it between launches): a sweep that follows an aborted one leaves at once,
    // and the host reports QPG_ERR_STATE at its next synchronisation point (ctx->flags[6], qpg_sim_stats / qpg_ctx_sync)
    bool ok = ld_volatile_u32(a.bar + 64) == 0;
    long long nprev = nstart;
    for (int j = a.j0; j <= a.j1 && ok; j++) {
        const long long namj0 = namj;
        // ---- phase A || compaction -------------------------------------------------------------------------
        if (in_team) sweep_field_A<M>(f, j, tm, sm_f, halo, timer ? sm_f.stamps : nullptr);
        else if (b == G - 1) compact_body(a.planes, 8, a.d_npp_w, a.d_nout, a.outmask, a.lists, 0, sm_i);
        if (PGC && b == las_cta) {   // the helper CTA: chi of the previous slice -> volume, laser slice images of this slice
            if (a.las_progress) {    // the envelope advance of the previous step may still be running on another SM: follow its progress
                if (tid == 0) {
                    long long t0 = 0;
                    for (unsigned n = 1;; n++) {
                        unsigned v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.las_progress) : "memory");
                        if ((int)(v - a.las_base) >= min(j, a.las_nz)) break;
                        if ((n & 63u) == 0) {
                            if (ld_volatile_u32(a.bar + 64)) break;
                            const long long now = clock64();
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > 4000000000LL) { atomicExch(a.bar + 64, 1u); break; }
                        }
                    }
                }
                __syncthreads();
            }
            if (j > a.j0) sweep_chi_fix(a, M, j - 1);
            sweep_laser_slice(a, M, j);
        }
        if (timer) work[0] += clock64() - tprev;
        ok = grid_barrier(a.bar, gep, sm_i);
        if (timer) { const long long t = clock64(); prof[0] += t - tprev; tprev = t; }
        if (!ok) break;
        const int npp = *(volatile const int *)a.pv.d_npp;
        if (timer) f.counters[0] += (long long)npp;
        const int ntiles = (npp + 31) >> 5, per = (ntiles + G - 1) / G;
        const int tile0 = b * per, tile1 = min(tile0 + per, ntiles);
        // ---- predictor-corrector loop -----------------------------------------------------------------------
        for (int it = 1; it <= f.iter_max; it++) {
            if (PGC) sweep_amj_phase_pgc<M>(a, npp, tile0, tile1, dep_tiles);
            else sweep_amj_phase<M>(a, npp, tile0, tile1, dep_tiles);
            if (timer) work[1] += clock64() - tprev;
            ok = grid_barrier(a.bar, gep, sm_i);
            if (timer) { const long long t = clock64(); prof[1] += t - tprev; tprev = t; namj++; }
            if (!ok) break;
            if (in_team) sweep_field_C<M>(f, tm, sm_f, timer ? sm_f.stamps : nullptr);
            if (timer) work[2] += clock64() - tprev;
            ok = grid_barrier(a.bar, gep, sm_i);
            if (timer) { const long long t = clock64(); prof[2] += t - tprev; tprev = t; }
            if (!ok) break;
            if (warp == 0) { const bool fin = sweep_conv_decide(a, it, b == 0, lane); if (lane == 0) sm_i[45] = fin ? 1 : 0; }   // one warp per CTA, same decision everywhere
            __syncthreads();
            if (sm_i[45] != 0) break;
        }
        if (!ok) break;
        // ---- push_u + push_x + bound flags + next slice's qdeposit || D ---------------------------------------
        if (warp == 0) {
            const int items = (f.nr + 2) * P, ipc = (items + G - 1) / G;
            for (int k = b * ipc + lane; k < min((b + 1) * ipc, items); k += 32) fused_D_item<M>(f, k, j);
        }
        if (PGC) sweep_push_phase_pgc<M>(a, npp, tile0, tile1, dep_tiles);
        else sweep_push_phase<M>(a, npp, tile0, tile1, dep_tiles);
        if (timer) work[3] += clock64() - tprev;
        ok = grid_barrier(a.bar, gep, sm_i);
        if (timer) {
            const long long t = clock64(); prof[3] += t - tprev; tprev = t;
            long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            a.trace[2 * (j - 1)] = now - nprev; a.trace[2 * (j - 1) + 1] = namj - namj0;
            nprev = now;
        }
        if (ok && j == 1 && a.back_flag && b == min(a.nteam, G - 1)) sweep_publish_back<M>(a);
    }
    if (!ok && tid == 0) {
        f.flags[6] = 1;   // latched for the host: this sim's state is not to be trusted any more
        // the upstream stage waits for this launch's backward hand-off with a stream memory operation: release it (poisoned
        // data, the error surfaces on the host) instead of leaving its stream hanging
        if (a.back_flag && b == min(a.nteam, G - 1)) st_release_sys(a.back_flag, a.back_seq);
    }
    if (PGC && ok && b == las_cta) sweep_chi_fix(a, M, a.j1);   // the last slice's susceptibility (the deposits are behind the slice's final barrier)
    // update_bound of the last slice (the next launch / the host expects compacted particles)
    if (ok && b == G - 1) compact_body(a.planes, 8, a.d_npp_w, a.d_nout, a.outmask, a.lists, 0, sm_i);
    if (timer) {
        long long nend;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(nend));
        for (int k = 0; k < 4; k++) { a.prof[k] += prof[k]; a.prof[8 + k] += work[k]; }
        for (int k = 0; k < 16; k++) a.prof[16 + k] += sm_f.stamps[k];
        a.prof[4] += clock64() - tstart;
        a.prof[5] += nend - nstart;
        a.prof[6] += a.j1 - a.j0 + 1;
        a.prof[7] += namj;
        if (ok) { f.flags[3] = a.j1 + 1; f.flags[4] += a.j1 - a.j0 + 1; }
    }
}

// beam.cu -- part3d: once-per-3D-step beam kernels (beam/part3d_class.f03, beam/part3d_comm.f03 xi hand-off).
// The beam is sharded by xi with the slab (noff2, nzp); x3 is xi measured from the box start like the reference.
#include "common.cuh"

#ifndef FULL
#define FULL 0xffffffffu
#endif
#define B3_BLOCK 256



struct Part3View {
    double *x1, *x2, *x3, *p1, *p2, *p3, *q;
    const int *d_npp;
    double *s1, *s2, *s3;   // spin planes (null: no spin)
    double amm;
};
static Part3View view3(qpg_part3d p) { Part3View v{p->x1, p->x2, p->x3, p->p1, p->p2, p->p3, p->q, p->d_npp, p->s1, p->s2, p->s3, p->amm}; return v; }
static inline int nplanes3(qpg_part3d p) { return p->s1 ? 10 : 7; }   // planes the compaction moves = reals of a wire record (part3d_comm.f03:683-694)

// beam/part3d_class.f03:578-638 push_spin_part3d: T-BMT precession of the spin vector s by the rotation vector omega (Boris-like rotation:
// |s| is conserved); ep = E q dt / 2m, bp = B q dt / (2 m gamma), v = the "time-centred velocity" (p_old + p_now) / (2 gamma) with p_now =
// this%p at the time of the call (see k_push3d)
__device__ __forceinline__ void push_spin3(double &s0, double &s1, double &s2, const double *ep, const double *bp, const double *v, double gam, double a)
{
    double coef = a + 1.0 / gam;
    double o0 = coef * bp[0] * gam, o1 = coef * bp[1] * gam, o2 = coef * bp[2] * gam;
    coef = -1.0 * (a + 1.0 / (1.0 + gam));
    o0 = o0 + coef * (v[1] * ep[2] - v[2] * ep[1]);
    o1 = o1 + coef * (v[2] * ep[0] - v[0] * ep[2]);
    o2 = o2 + coef * (v[0] * ep[1] - v[1] * ep[0]);
    const double vdotb = v[0] * bp[0] + v[1] * bp[1] + v[2] * bp[2];
    coef = -1.0 * (a * (gam * gam) / (1.0 + gam) * vdotb);
    o0 = o0 + coef * v[0]; o1 = o1 + coef * v[1]; o2 = o2 + coef * v[2];
    const double t0 = s0 + (s1 * o2 - s2 * o1), t1 = s1 + (s2 * o0 - s0 * o2), t2 = s2 + (s0 * o1 - s1 * o0);
    coef = 2.0 / (1.0 + o0 * o0 + o1 * o1 + o2 * o2);
    const double n0 = s0 + coef * (t1 * o2 - t2 * o1), n1 = s1 + coef * (t2 * o0 - t0 * o2), n2 = s2 + coef * (t0 * o1 - t1 * o0);
    s0 = n0; s1 = n1; s2 = n2;
}
static double **plane_table3(qpg_part3d p) { return (double **)(p->lists + 2 * p->npmax); }

// beam/part3d_class.f03:221-356 qdeposit_part3d (accumulation); f2 image layout [slice][node][P] (dim 1).
// Beam particles are created slice by slice, sector by sector, cell by cell (fdist3d_std_class.f03:439-613), so the lanes of a warp
// share a few (slice, cell) pairs: the 4 P sums a particle contributes to its four nodes are reduced over the warp with the one-hot
// DMMA reduction of the plasma charge deposit (particles.cu warp_deposit_q_mma, key = flattened node index of the f2 volume) -- one
// RED per (node, plane) and warp instead of one atomic per lane: the scatter was atomic-throughput bound (12 atomics per particle
// onto ~50 radial cells).  Launched as a grid-stride loop over warp tiles: the host only knows an upper bound of the live count.
#define B3_MAX_GRID (148 * 8)
static inline int b3_grid(long n) { const long g = (n + B3_BLOCK - 1) / B3_BLOCK; return (int)(g < B3_MAX_GRID ? g : B3_MAX_GRID); }
template <int M>
__global__ void __launch_bounds__(B3_BLOCK) k_qdeposit3d(Part3View pv, double *__restrict__ f2, double idr, double idz, int nr, int noff2, int nzp, int mode,
                                                        const unsigned *__restrict__ pushed)
{
    // mode 0: every particle.  The split deposit of a pipeline stage (qpg_part3d_qdeposit_part): 1 = the particles the interior pass of the
    // split push has advanced (bit set in `pushed`), 2 = the others, 3 = the particles appended since the last hand-off (index >= d_npp[3])
    constexpr int P = 2 * M + 1;
    extern __shared__ double dep_tiles[];
    const int npp = *pv.d_npp, lane = threadIdx.x & 31;
    const int first = mode == 3 ? pv.d_npp[3] : 0;
    double *tile = dep_tiles + (threadIdx.x >> 5) * DepTile<M>::doubles;
    for (long base = (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < npp; base += (long)gridDim.x * blockDim.x) {   // warp-uniform
        const long i = base + lane;
        bool ok = i < npp && i >= first;
        if (mode == 1 || mode == 2) {
            const unsigned w = pushed[base >> 5];
            if ((mode == 1 && w == 0u) || (mode == 2 && w == 0xffffffffu)) continue;      // warp-uniform: nothing to do in this tile
            ok = ok && (((w >> lane) & 1u) == (mode == 1 ? 1u : 0u));
        } else if (base + 32 <= first) continue;
        double ph[P], wr[2] = {0.0, 0.0}, wz[2] = {0.0, 0.0};
        int nn = 0, mm = 0;
#pragma unroll
        for (int pl = 0; pl < P; pl++) ph[pl] = 0.0;
        if (ok) {
            const double x1 = pv.x1[i], x2 = pv.x2[i], x3 = pv.x3[i], q = pv.q[i];
            double pos_r = __dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2))), idr);
            double pos_z = __dmul_rn(x3, idz);
            const double c0 = x1 / pos_r * idr, s0 = -x2 / pos_r * idr;
            nn = (int)floor(pos_r); mm = (int)floor(pos_z);
            const double fr = pos_r - (double)nn, fz = pos_z - (double)mm;
            nn = nn + 1;
            mm = mm - noff2 + 1;
            ok = !(mm < 1 || mm > nzp || nn < 1 || nn > nr);  // not ours (hand-off pending) -- never deposit out of bounds
            wr[0] = 1.0 - fr; wr[1] = fr; wz[0] = 1.0 - fz; wz[1] = fz;
            double phr = q, phi = 0.0;
            ph[0] = phr;
#pragma unroll
            for (int m = 1; m <= M; m++) {
                double t = phr * c0 - phi * s0;
                phi = phr * s0 + phi * c0;
                phr = t;
                ph[2 * m - 1] = phr;
                ph[2 * m] = phi;
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {   // slice mm + k - 1: nodes nn, nn + 1 are the pair (key, key + 1) of the flattened volume
            double alpha[2 * P];
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int pl = 0; pl < P; pl++) alpha[j * P + pl] = ok ? (wr[j] * wz[k]) * ph[pl] : 0.0;
            const int key = ok ? (mm + k - 1) * (nr + 2) + nn : -1;
#ifdef QPG_BEAM_DEPOSIT_PLAIN   // A/B aid: one RED per lane and value instead of the warp reduction
            if (ok) for (int r = 0; r < 2 * P; r++) red_add(f2 + (size_t)key * P + r, alpha[r]);
#else
            warp_deposit_q_mma<M>(alpha, key, f2, tile, lane);
#endif
        }
    }
}
// axis rules + 1/(j-1) for slices 1..nzp, part3d_class.f03:318-351
__global__ void k_qdep3d_fix(double *__restrict__ f2, int nr, int P, int nzp)
{
    const size_t n1 = (size_t)(nr + 2) * P, n = n1 * nzp;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(k % n1), j = r / P, pl = r % P;
        double v = f2[k];
        if (j == 0) v = 0.0;
        else if (j == 1) v = pl == 0 ? 8.0 * v : 0.0;
        else v = v * (1.0 / (double)(j - 1));
        f2[k] = v;
    }
}

// beam/part3d_class.f03:691-790 interp_emf_part3d + :477-576 push_reduced / :358-475 push_boris ; also flags
// r >= edge_r or xi >= edge_z (update_bound_part3d :640-689) when flag != 0
template <int M>
__global__ void __launch_bounds__(B3_BLOCK) k_push3d(Part3View pv, const double *__restrict__ ef2, const double *__restrict__ bf2, double idr, double idz,
                                                    int nr, int noff2, int nzp, double qbm, double dt, int push_type, int pass, unsigned *__restrict__ pushed)
{
    // pass 0: every particle of the slab.  The xi-pipeline splits the push (pipeline.LocalPipeline._tail): pass 1 = the particles whose
    // gather does not touch the guard slice nzp + 1 (slice index < nzp) -- it runs BEFORE the downstream stage's first-slice e / b has
    // arrived -- and records them in the bitmap `pushed`; pass 2 = the rest (the last slice of the slab), after the message.
    constexpr int P = 2 * M + 1;
    const int npp = *pv.d_npp, lane = threadIdx.x & 31;
    for (long base = (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < npp; base += (long)gridDim.x * blockDim.x) {   // warp-uniform; the host knows an upper bound only
    const long i = base + lane;
    const bool live = i < npp;
    if (pass == 2 && pushed[base >> 5] == 0xffffffffu) continue;      // warp-uniform: the interior pass has advanced the whole tile
    double x1 = 1.0, x2 = 0.0, x3 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    if (live) { x1 = pv.x1[i]; x2 = pv.x2[i]; x3 = pv.x3[i]; p1 = pv.p1[i]; p2 = pv.p2[i]; p3 = pv.p3[i]; }
    double pos_r = __dmul_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2))), idr);
    double pos_z = __dmul_rn(x3, idz);
    const double cc = x1 / pos_r * idr, ss = x2 / pos_r * idr;
    int nn = (int)pos_r, mm = (int)pos_z;
    const double fr = pos_r - (double)nn, fz = pos_z - (double)mm;
    nn = nn + 1;
    mm = mm - noff2 + 1;
    bool go = live && !(mm < 1 || mm > nzp || nn < 1 || nn > nr);
    if (pass == 1) {
        go = go && mm < nzp;
        const unsigned bal = __ballot_sync(FULL, go);
        if (lane == 0) pushed[base >> 5] = bal;
    } else if (pass == 2) go = go && !((pushed[base >> 5] >> lane) & 1u);
    if (!go) continue;
    const size_t n1 = (size_t)(nr + 2) * P * 3;
    const double wr[2] = {1.0 - fr, fr}, wz[2] = {1.0 - fz, fz};
    double ep[3] = {0, 0, 0}, bp[3] = {0, 0, 0};
    double pr2[M + 1], pi2[M + 1];
    {
        double phr = 1.0, phi = 0.0;
        pr2[0] = 1.0; pi2[0] = 0.0;
#pragma unroll
        for (int m = 1; m <= M; m++) { double t = phr * cc - phi * ss; phi = phr * ss + phi * cc; phr = t; pr2[m] = 2.0 * phr; pi2[m] = 2.0 * phi; }
    }
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const size_t off = (size_t)(mm + k - 1) * n1 + (size_t)(nn + j) * P * 3;
            const double *e = ef2 + off, *b = bf2 + off;
            const double wt = wr[j] * wz[k];
#pragma unroll
            for (int c = 0; c < 3; c++) { ep[c] = fma(__ldg(e + c), wt, ep[c]); bp[c] = fma(__ldg(b + c), wt, bp[c]); }
#pragma unroll
            for (int m = 1; m <= M; m++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    ep[c] = fma(__ldg(e + (2 * m - 1) * 3 + c) * pr2[m] - __ldg(e + (2 * m) * 3 + c) * pi2[m], wt, ep[c]);
                    bp[c] = fma(__ldg(b + (2 * m - 1) * 3 + c) * pr2[m] - __ldg(b + (2 * m) * 3 + c) * pi2[m], wt, bp[c]);
                }
        }
    double t = ep[0] * cc - ep[1] * ss; ep[1] = ep[0] * ss + ep[1] * cc; ep[0] = t;
    t = bp[0] * cc - bp[1] * ss; bp[1] = bp[0] * ss + bp[1] * cc; bp[0] = t;
    const double qtmh = qbm * dt * 0.5;
    if (push_type == QPG_PUSH3_REDUCED) {
#pragma unroll
        for (int c = 0; c < 3; c++) { ep[c] *= qtmh; bp[c] *= qtmh; }
        const double w0 = ep[0] - bp[1], w1 = ep[1] + bp[0], w2 = ep[2];
        const double po[3] = {p1, p2, p3};
        p1 = p1 + w0; p2 = p2 + w1; p3 = p3 + w2;
        const double gam = sqrt(1.0 + p1 * p1 + p2 * p2 + p3 * p3);            // :536, of the half-advanced momentum
        p1 = p1 + w0; p2 = p2 + w1; p3 = p3 + w2;
        if (pv.s1) {                                                            // :550-557: after both half advances
            const double igam = 1.0 / gam;
            bp[0] *= igam; bp[1] *= igam; bp[2] *= igam;
            const double v[3] = {0.5 * (po[0] + p1) / gam, 0.5 * (po[1] + p2) / gam, 0.5 * (po[2] + p3) / gam};
            double s0 = pv.s1[i], s1 = pv.s2[i], s2 = pv.s3[i];
            push_spin3(s0, s1, s2, ep, bp, v, gam, pv.amm);
            pv.s1[i] = s0; pv.s2[i] = s1; pv.s3[i] = s2;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) ep[c] *= qtmh;
        double ut0 = p1 + ep[0], ut1 = p2 + ep[1], ut2 = p3 + ep[2];
        const double gam = sqrt(1.0 + (ut0 * ut0 + ut1 * ut1 + ut2 * ut2));
        const double gq = qtmh / gam;
        bp[0] *= gq; bp[1] *= gq; bp[2] *= gq;
        if (pv.s1) {                                                            // :425-427: this%p still holds the OLD momentum, so v = p_old / gamma
            const double v[3] = {0.5 * (p1 + p1) / gam, 0.5 * (p2 + p2) / gam, 0.5 * (p3 + p3) / gam};
            double s0 = pv.s1[i], s1 = pv.s2[i], s2 = pv.s3[i];
            push_spin3(s0, s1, s2, ep, bp, v, gam, pv.amm);
            pv.s1[i] = s0; pv.s2[i] = s1; pv.s3[i] = s2;
        }
        p1 = ut0 + ut1 * bp[2] - ut2 * bp[1];
        p2 = ut1 + ut2 * bp[0] - ut0 * bp[2];
        p3 = ut2 + ut0 * bp[1] - ut1 * bp[0];
        const double ostq = 2.0 / (1.0 + bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]);
        bp[0] *= ostq; bp[1] *= ostq; bp[2] *= ostq;
        ut0 = ut0 + p2 * bp[2] - p3 * bp[1];
        ut1 = ut1 + p3 * bp[0] - p1 * bp[2];
        ut2 = ut2 + p1 * bp[1] - p2 * bp[0];
        p1 = ut0 + ep[0]; p2 = ut1 + ep[1]; p3 = ut2 + ep[2];
    }
    const double dt_gam = dt / sqrt(1.0 + p1 * p1 + p2 * p2 + p3 * p3);
    x1 = x1 + p1 * dt_gam;
    x2 = x2 + p2 * dt_gam;
    x3 = x3 - p3 * dt_gam + dt;
    pv.x1[i] = x1; pv.x2[i] = x2; pv.x3[i] = x3;
    pv.p1[i] = p1; pv.p2[i] = p2; pv.p3[i] = p3;
    }
}

// flag kernel: kind 0 -> out of the box (r >= edge_r or xi >= edge_z); kind 1 -> xi >= zhi (forward hand-off)
__global__ void __launch_bounds__(B3_BLOCK) k_flag3d(Part3View pv, double edge_r, double edge_z, int kind, unsigned *__restrict__ outmask, int *__restrict__ d_nout)
{
    const int npp = *pv.d_npp, lane = threadIdx.x & 31;
    for (long base = (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < npp; base += (long)gridDim.x * blockDim.x) {   // warp-uniform
        const long i = base + lane;
        bool out = false;
        if (i < npp) {
            const double x1 = pv.x1[i], x2 = pv.x2[i], x3 = pv.x3[i];
            if (kind == 0) {
                const double pos = __dsqrt_rn(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2)));
                out = (pos >= edge_r) || (x3 >= edge_z);
            } else out = x3 >= edge_z;
        }
        const unsigned bal = __ballot_sync(FULL, out);
        if (lane == 0) { outmask[i >> 5] = bal; if (bal) atomicAdd(d_nout, __popc(bal)); }
    }
}

// ordered pack of the flagged particles (ascending index like pack_particles :685-745): one CTA
__global__ void __launch_bounds__(1024, 1) k_pack3d(Part3View pv, const int *d_nout, const unsigned *__restrict__ outmask, double *__restrict__ buf, long cap,
                                                   int *pv_overflow)
{
    const int nrec = pv.s1 ? 10 : 7;     // reals of a wire record: x, p, q (+ s)
    __shared__ int sm[40];
    const int n = *pv.d_npp, nout = *d_nout, tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { buf[0] = (double)(nout > cap ? cap : nout); if (nout > cap) pv_overflow[0] = 1; }
    if (nout == 0) return;
    const int nwords = (n + 31) >> 5, wpt = (nwords + nt - 1) / nt;
    const int wbeg = min(tid * wpt, nwords), wend = min(wbeg + wpt, nwords);
    int cnt = 0;
    for (int w = wbeg; w < wend; w++) {
        const int lo = w << 5;
        const unsigned inrange = (lo + 32 <= n) ? FULL : ((1u << (n - lo)) - 1u);
        cnt += __popc(outmask[w] & inrange);
    }
    // exclusive scan of cnt
    const int lane = tid & 31, wp = tid >> 5, nw = nt >> 5;
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sm[wp] = incl;
    __syncthreads();
    if (wp == 0) {
        int t = lane < nw ? sm[lane] : 0, ti = t;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(FULL, ti, o); if (lane >= o) ti += u; }
        sm[lane] = ti - t;
    }
    __syncthreads();
    int off = sm[wp] + incl - cnt;
    for (int w = wbeg; w < wend; w++) {
        const int lo = w << 5;
        const unsigned inrange = (lo + 32 <= n) ? FULL : ((1u << (n - lo)) - 1u);
        unsigned bits = outmask[w] & inrange;
        while (bits) {
            const int b = __ffs(bits) - 1; bits &= bits - 1;
            const int i = lo + b;
            if (off < cap) {
                double *r = buf + 1 + (size_t)nrec * off;
                r[0] = pv.x1[i]; r[1] = pv.x2[i]; r[2] = pv.x3[i]; r[3] = pv.p1[i]; r[4] = pv.p2[i]; r[5] = pv.p3[i]; r[6] = pv.q[i];
                if (pv.s1) { r[7] = pv.s1[i]; r[8] = pv.s2[i]; r[9] = pv.s3[i]; }
            }
            off++;
        }
    }
}
__global__ void k_unpack3d(Part3View pv, int *d_npp_w, const double *__restrict__ buf, long cap, long npmax)
{
    const int add = (int)min((long)buf[0], cap);
    const int n0 = *pv.d_npp;
    const int room = (int)min((long)add, npmax - n0);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < room; k += gridDim.x * blockDim.x) {
        const double *r = buf + 1 + (size_t)(pv.s1 ? 10 : 7) * k;
        const int i = n0 + k;
        pv.x1[i] = r[0]; pv.x2[i] = r[1]; pv.x3[i] = r[2]; pv.p1[i] = r[3]; pv.p2[i] = r[4]; pv.p3[i] = r[5]; pv.q[i] = r[6];
        if (pv.s1) { pv.s1[i] = r[7]; pv.s2[i] = r[8]; pv.s3[i] = r[9]; }
    }
    // the count is bumped by a follow-up single-thread kernel so every block sees the old n0
}
__global__ void k_bump_npp(int *d_npp, const double *__restrict__ buf, long cap, long npmax)
{
    const int add = (int)min((long)buf[0], cap);
    const int n0 = *d_npp;
    d_npp[3] = n0;                                   // first index of the appended particles (qpg_part3d_qdeposit_part, part 3)
    *d_npp = n0 + (int)min((long)add, npmax - n0);
}

template <int M> static void l_qdep3d(int grid, cudaStream_t st, Part3View pv, double *f2, double idr, double idz, int nr, int noff2, int nzp, int mode, const unsigned *pushed)
{
    constexpr size_t smem = sizeof(double) * DepTile<M>::doubles * (B3_BLOCK / 32);
    static bool attr_set = false;   // > 48 KB of dynamic shared memory needs the opt-in (M >= 3)
    if (!attr_set) { cudaFuncSetAttribute(k_qdeposit3d<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    k_qdeposit3d<M><<<grid, B3_BLOCK, smem, st>>>(pv, f2, idr, idz, nr, noff2, nzp, mode, pushed);
}
template <int M> static void l_push3d(int grid, cudaStream_t st, Part3View pv, const double *e, const double *b, double idr, double idz, int nr, int noff2, int nzp, double qbm, double dt, int pt,
                                      int pass, unsigned *pushed)
{ k_push3d<M><<<grid, B3_BLOCK, 0, st>>>(pv, e, b, idr, idz, nr, noff2, nzp, qbm, dt, pt, pass, pushed); }

extern "C" int qpg_part3d_create(qpg_part3d *out, qpg_ctx ctx, double qbm, double dt, long npmax, int nz_total, int noff2, int nzp)
{
    ARG_TRY(out && ctx, "null arg");
    ARG_TRY(npmax >= 32 && npmax < (1L << 31) - 64, "npmax out of range");
    ARG_TRY(nzp >= 1 && noff2 >= 0 && noff2 + nzp <= nz_total, "bad xi slab");
    npmax = (npmax + 31) & ~31L;
    qpg_part3d p = new qpg_part3d_s();
    memset(p, 0, sizeof(*p));
    p->ctx = ctx; p->qbm = qbm; p->dt = dt; p->npmax = npmax; p->npp_hi = 0;
    p->nz_total = nz_total; p->noff2 = noff2; p->nzp = nzp;
    CUDA_TRY(cudaMalloc(&p->slab, sizeof(double) * 7 * npmax));
    CUDA_TRY(cudaMemsetAsync(p->slab, 0, sizeof(double) * 7 * npmax, ctx->stream));
    p->x1 = p->slab; p->x2 = p->slab + npmax; p->x3 = p->slab + 2 * npmax; p->p1 = p->slab + 3 * npmax; p->p2 = p->slab + 4 * npmax;
    p->p3 = p->slab + 5 * npmax; p->q = p->slab + 6 * npmax;
    CUDA_TRY(cudaMalloc(&p->d_npp, sizeof(int) * 4));
    CUDA_TRY(cudaMemsetAsync(p->d_npp, 0, sizeof(int) * 4, ctx->stream));
    p->d_nout = p->d_npp + 1;
    CUDA_TRY(cudaMalloc(&p->outmask, sizeof(unsigned) * (npmax / 32 + 1)));
    CUDA_TRY(cudaMemsetAsync(p->outmask, 0, sizeof(unsigned) * (npmax / 32 + 1), ctx->stream));
    CUDA_TRY(cudaMalloc(&p->pushed, sizeof(unsigned) * (npmax / 32 + 1)));
    CUDA_TRY(cudaMemsetAsync(p->pushed, 0, sizeof(unsigned) * (npmax / 32 + 1), ctx->stream));
    CUDA_TRY(cudaMalloc(&p->lists, sizeof(int) * (2 * npmax + 64)));
    double *h[7] = {p->x1, p->x2, p->x3, p->p1, p->p2, p->p3, p->q};
    CUDA_TRY(cudaMemcpy(plane_table3(p), h, sizeof(h), cudaMemcpyHostToDevice));
    *out = p;
    return 0;
}
extern "C" int qpg_part3d_destroy(qpg_part3d p)
{
    if (!p) return 0;
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->slab); cudaFree(p->d_npp); cudaFree(p->outmask); cudaFree(p->pushed); cudaFree(p->lists); cudaFree(p->s1);
    delete p;
    return 0;
}
// has_spin (init_part3d :117-135 with `amm` present): three more planes that follow the particles through the push (push_spin :578), the
// removal of particles (update_bound :668-670) and the hand-off to the next stage (10-real wire record, part3d_comm.f03:683-694)
extern "C" int qpg_part3d_enable_spin(qpg_part3d p, double amm)
{
    ARG_TRY(p, "null arg");
    if (!p->s1) {
        CUDA_TRY(cudaMalloc(&p->s1, sizeof(double) * 3 * p->npmax));
        CUDA_TRY(cudaMemsetAsync(p->s1, 0, sizeof(double) * 3 * p->npmax, p->ctx->stream));
        p->s2 = p->s1 + p->npmax; p->s3 = p->s1 + 2 * p->npmax;
        double *h[10] = {p->x1, p->x2, p->x3, p->p1, p->p2, p->p3, p->q, p->s1, p->s2, p->s3};
        CUDA_TRY(cudaMemcpy(plane_table3(p), h, sizeof(h), cudaMemcpyHostToDevice));
    }
    p->amm = amm;
    return 0;
}
extern "C" int qpg_part3d_has_spin(qpg_part3d p) { return p && p->s1 ? 1 : 0; }
// s[npp][3] for the particles of the last qpg_part3d_upload (same order) / of the current set
extern "C" int qpg_part3d_upload_spin(qpg_part3d p, const double *s, long npp)
{
    ARG_TRY(p && p->s1 && (npp == 0 || s), "no spin planes (qpg_part3d_enable_spin) / null arg");
    ARG_TRY(npp >= 0 && npp <= p->npmax, "npp exceeds npmax");
    std::vector<double> h((size_t)3 * npp);
    for (long i = 0; i < npp; i++)
        for (int c = 0; c < 3; c++) h[(size_t)c * npp + i] = s[3 * i + c];
    double *dst[3] = {p->s1, p->s2, p->s3};
    for (int a = 0; a < 3 && npp; a++) CUDA_TRY(cudaMemcpyAsync(dst[a], h.data() + (size_t)a * npp, sizeof(double) * npp, cudaMemcpyHostToDevice, p->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(p->ctx->stream));
    return 0;
}
extern "C" int qpg_part3d_download_spin(qpg_part3d p, double *s, long *npp_out)
{
    ARG_TRY(p && p->s1, "no spin planes (qpg_part3d_enable_spin)");
    int n = 0;
    cudaStream_t st = p->ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(&n, p->d_npp, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (npp_out) *npp_out = n;
    if (!s || !n) return 0;
    std::vector<double> h((size_t)3 * n);
    double *src[3] = {p->s1, p->s2, p->s3};
    for (int a = 0; a < 3; a++) CUDA_TRY(cudaMemcpyAsync(h.data() + (size_t)a * n, src[a], sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (long i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) s[3 * i + c] = h[(size_t)c * n + i];
    return 0;
}
// doubles of a forward hand-off message: the count + wire_cap records of 7 (10 with spin) reals
extern "C" long qpg_part3d_wire_count(qpg_part3d p) { return p ? 1 + (long)nplanes3(p) * qpg_part3d_wire_cap(p) : -1; }
extern "C" int qpg_part3d_upload(qpg_part3d p, const double *x, const double *pm, const double *q, long npp)
{
    ARG_TRY(p && (npp == 0 || (x && pm && q)), "null arg");
    ARG_TRY(npp >= 0 && npp <= p->npmax, "npp exceeds npmax");
    std::vector<double> h((size_t)6 * npp);
    for (long i = 0; i < npp; i++)
        for (int c = 0; c < 3; c++) { h[(size_t)c * npp + i] = x[3 * i + c]; h[(size_t)(3 + c) * npp + i] = pm[3 * i + c]; }
    cudaStream_t st = p->ctx->stream;
    double *dst[6] = {p->x1, p->x2, p->x3, p->p1, p->p2, p->p3};
    for (int a = 0; a < 6 && npp; a++) CUDA_TRY(cudaMemcpyAsync(dst[a], h.data() + (size_t)a * npp, sizeof(double) * npp, cudaMemcpyHostToDevice, st));
    if (npp) CUDA_TRY(cudaMemcpyAsync(p->q, q, sizeof(double) * npp, cudaMemcpyHostToDevice, st));
    int cnt[2] = {(int)npp, 0};
    CUDA_TRY(cudaMemcpyAsync(p->d_npp, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(p->outmask, 0, sizeof(unsigned) * (p->npmax / 32 + 1), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->npp_hi = npp;
    return 0;
}
extern "C" int qpg_part3d_download(qpg_part3d p, double *x, double *pm, double *q, long *npp_out)
{
    ARG_TRY(p, "null arg");
    int nn[3] = {0, 0, 0};
    cudaStream_t st = p->ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(nn, p->d_npp, sizeof(nn), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (nn[2]) { qpg_set_error("beam hand-off overflow: more than %ld particles crossed the slab edge in one step (qpg_part3d_set_wire_cap)", qpg_part3d_wire_cap(p)); return QPG_ERR_STATE; }
    const int n = nn[0];
    const long npp = n;
    p->npp_hi = npp;
    if (npp_out) *npp_out = npp;
    if (!x && !pm && !q) return 0;
    std::vector<double> h((size_t)6 * npp);
    double *src[6] = {p->x1, p->x2, p->x3, p->p1, p->p2, p->p3};
    for (int a = 0; a < 6 && npp; a++) CUDA_TRY(cudaMemcpyAsync(h.data() + (size_t)a * npp, src[a], sizeof(double) * npp, cudaMemcpyDeviceToHost, st));
    if (q && npp) CUDA_TRY(cudaMemcpyAsync(q, p->q, sizeof(double) * npp, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (long i = 0; i < npp; i++)
        for (int c = 0; c < 3; c++) { if (x) x[3 * i + c] = h[(size_t)c * npp + i]; if (pm) pm[3 * i + c] = h[(size_t)(3 + c) * npp + i]; }
    return 0;
}
// the deposit in two halves, so that a pipeline stage can scatter its own particles BEFORE the upstream guard slice
// arrives: raw = the scatter-add of part3d_class.f03:221-316, fix = the axis rules and 1/(j-1) of :318-351 (which must see
// the upstream stage's guard-slice contribution in slice 1)
extern "C" int qpg_part3d_qdeposit_raw(qpg_part3d p, qpg_field q)
{
    ARG_TRY(p && q && q->dim == 1 && q->has2d && q->nzp == p->nzp && q->ctx == p->ctx, "q must be a dim-1 field with this slab's 2D layout");
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_DEPOSIT3D);
    if (p->npp_hi > 0) {
        const int grid = b3_grid(p->npp_hi);
        DISPATCH_M(c->M, l_qdep3d, grid, c->stream, view3(p), q->f2, 1.0 / c->dr, 1.0 / c->dxi, c->nr, p->noff2, p->nzp, 0, p->pushed);
        count_launch(c);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
/* the raw deposit in three parts for a stage of the xi-pipeline, so that only a sliver of it stays behind the backward hand-off:
 * part 1 = the particles advanced by qpg_part3d_push_interior, 2 = the others (after qpg_part3d_push_edge, BEFORE update_bound: the bitmap is
 * indexed by the pre-compaction order), 3 = the particles appended by the last qpg_part3d_unpack (before qpg_part3d_pack_forward compacts).
 * Particles that have left the slab or the box deposit nothing, exactly as in the full deposit after update_bound / the hand-off. */
extern "C" int qpg_part3d_qdeposit_part(qpg_part3d p, qpg_field q, int part)
{
    ARG_TRY(p && q && q->dim == 1 && q->has2d && q->nzp == p->nzp && q->ctx == p->ctx && part >= 1 && part <= 3, "bad arg");
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_DEPOSIT3D);
    if (p->npp_hi > 0) {
        const int grid = b3_grid(p->npp_hi);
        DISPATCH_M(c->M, l_qdep3d, grid, c->stream, view3(p), q->f2, 1.0 / c->dr, 1.0 / c->dxi, c->nr, p->noff2, p->nzp, part, p->pushed);
        count_launch(c);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part3d_qdeposit_fix(qpg_part3d p, qpg_field q)
{
    ARG_TRY(p && q && q->dim == 1 && q->has2d && q->nzp == p->nzp && q->ctx == p->ctx, "q must be a dim-1 field with this slab's 2D layout");
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_DEPOSIT3D);
    k_qdep3d_fix<<<592, 256, 0, c->stream>>>(q->f2, c->nr, c->P, p->nzp);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part3d_qdeposit(qpg_part3d p, qpg_field q)
{
    int rc = qpg_part3d_qdeposit_raw(p, q);
    return rc ? rc : qpg_part3d_qdeposit_fix(p, q);
}
static int part3d_push_pass(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf, int pass)
{
    ARG_TRY(p && ef && bf && ef->dim == 3 && bf->dim == 3 && ef->has2d && bf->has2d && ef->nzp == p->nzp && bf->nzp == p->nzp, "e, b must be dim-3 fields with this slab's 2D layout");
    ARG_TRY(push_type == QPG_PUSH3_REDUCED || push_type == QPG_PUSH3_BORIS, "Invalid pusher type! Only \"reduced\" and \"boris\" are supported currently.");
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_PUSH3D);
    const int grid = b3_grid(p->npp_hi);
    DISPATCH_M(c->M, l_push3d, grid, c->stream, view3(p), ef->f2, bf->f2, 1.0 / c->dr, 1.0 / c->dxi, c->nr, p->noff2, p->nzp, p->qbm, p->dt, push_type, pass, p->pushed);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part3d_push(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf) { return part3d_push_pass(p, push_type, ef, bf, 0); }
/* the push in two passes for a pipeline stage: `interior` = the particles that do not gather from the guard slice nzp + 1 (may run before
 * the downstream stage's first-slice e / b has arrived), `edge` = the others; interior must precede edge, together they are qpg_part3d_push */
extern "C" int qpg_part3d_push_interior(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf) { return part3d_push_pass(p, push_type, ef, bf, 1); }
extern "C" int qpg_part3d_push_edge(qpg_part3d p, int push_type, qpg_field ef, qpg_field bf) { return part3d_push_pass(p, push_type, ef, bf, 2); }
extern "C" int qpg_part3d_update_bound(qpg_part3d p)
{
    ARG_TRY(p, "null arg");
    if (p->npp_hi == 0) return 0;
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_PUSH3D);
    const int grid = b3_grid(p->npp_hi);
    k_flag3d<<<grid, B3_BLOCK, 0, c->stream>>>(view3(p), (double)c->nr * c->dr, (double)p->nz_total * c->dxi, 0, p->outmask, p->d_nout);
    k_compact<<<1, 1024, 0, c->stream>>>(plane_table3(p), nplanes3(p), p->d_npp, p->d_nout, p->outmask, p->lists, 0, nullptr);
    count_launch(c, 2);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" const int *qpg_part3d_count_ptr(qpg_part3d p) { return p ? p->d_npp : nullptr; }
extern "C" long qpg_part3d_wire_cap(qpg_part3d p)
{
    if (!p) return -1;
    if (p->wire_cap > 0) return p->wire_cap;
    return p->npmax / 10 > 1024 ? p->npmax / 10 : 1024;   /* nbmax = 0.1 npmax, part3d_class.f03:127 */
}
extern "C" int qpg_part3d_set_wire_cap(qpg_part3d p, long cap)
{
    ARG_TRY(p && cap >= 0 && cap <= p->npmax, "wire cap out of range");
    p->wire_cap = cap;
    return 0;
}
extern "C" int qpg_part3d_pack_forward(qpg_part3d p, double *dev_buf)
{
    ARG_TRY(p && dev_buf, "null arg");
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_MOVE3D);
    const long cap = qpg_part3d_wire_cap(p);
    const int grid = b3_grid(p->npp_hi);
    if (grid > 0) k_flag3d<<<grid, B3_BLOCK, 0, c->stream>>>(view3(p), 0.0, (double)(p->noff2 + p->nzp) * c->dxi, 1, p->outmask, p->d_nout);
    k_pack3d<<<1, 1024, 0, c->stream>>>(view3(p), p->d_nout, p->outmask, dev_buf, cap, p->d_npp + 2);
    k_compact<<<1, 1024, 0, c->stream>>>(plane_table3(p), nplanes3(p), p->d_npp, p->d_nout, p->outmask, p->lists, 1, nullptr);
    count_launch(c, 3);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_part3d_unpack(qpg_part3d p, const double *dev_buf)
{
    ARG_TRY(p && dev_buf, "null arg");
    qpg_ctx c = p->ctx;
    TprofScope tp(c, TP_MOVE3D);
    const long cap = qpg_part3d_wire_cap(p);
    k_unpack3d<<<64, 256, 0, c->stream>>>(view3(p), p->d_npp, dev_buf, cap, p->npmax);
    k_bump_npp<<<1, 1, 0, c->stream>>>(p->d_npp, dev_buf, cap, p->npmax);
    count_launch(c, 2);
    CUDA_TRY(cudaGetLastError());
    p->npp_hi = p->npp_hi + cap < p->npmax ? p->npp_hi + cap : p->npmax;
    return 0;
}

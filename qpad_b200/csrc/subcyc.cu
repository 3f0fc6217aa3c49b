// subcyc.cu -- the particle side of QPAD's sub-cycling / clamp variant of the slice loop, SURVEY.md §8(f) rank 4.
//
//   proj_subcyc/part2d_subcyc_class.f03:28-46   get_exp_fac_max  -> k_exp_fac_max   (max of gamma / (gamma - p_z), 1 if empty)
//   proj_subcyc/part2d_subcyc_class.f03:48-66   clamp_exp_fac    -> k_clamp_exp_fac (slow the particle down along its momentum
//                                                                   until its expansion factor equals the clamp)
//   proj_subcyc/simulation_subcyc_class.f03:431-451  the sub-step rule -> qpg_subcyc_step (host arithmetic, no device work)
//
// The sub-cycled slice body itself (simulation_subcyc_class.f03:216-376) is the standard per-routine sequence repeated n_subcyc
// times with dxi / n_subcyc: every other routine it calls already exists (qpg_part2d_qdeposit / amjdeposit / push_u / push_x
// take dt as an argument); the host loop is qpad_b200/subcyc.py.
//
// STATUS: both kernels are bit-exact against the oracle on the GPU (tests/test_gpu_extras.py, run by default since round 2) and in
// the host emulation of tests/emu: the expansion factor is one IEEE division, the clamp uses non-contracted IEEE operations in
// the reference's order.
#include "common.cuh"

// bit pattern order == numeric order for positive doubles (the expansion factor is > 0 because gamma > p_z)
__global__ void __launch_bounds__(256) k_exp_fac_max(const double *__restrict__ gamma, const double *__restrict__ p3, const int *__restrict__ d_npp,
                                                     unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long red[256];
    const int npp = *d_npp;
    unsigned long long best = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npp; i += gridDim.x * blockDim.x) {
        const double f = __ddiv_rn(gamma[i], __dsub_rn(gamma[i], p3[i]));
        const unsigned long long b = (unsigned long long)__double_as_longlong(f);
        if (f > 0.0 && b > best) best = b;
    }
    red[threadIdx.x] = best;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s && red[threadIdx.x + s] > red[threadIdx.x]) red[threadIdx.x] = red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && red[0] != 0ull) atomicMax(out, red[0]);
}

__global__ void __launch_bounds__(256) k_clamp_exp_fac(double *__restrict__ p1, double *__restrict__ p2, double *__restrict__ p3, double *__restrict__ gamma,
                                                       const int *__restrict__ d_npp, double clamp)
{
    const int npp = *d_npp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npp; i += gridDim.x * blockDim.x) {
        const double g = gamma[i], pz = p3[i];
        const double exp_fac = __ddiv_rn(g, __dsub_rn(g, pz));
        if (exp_fac > clamp) {
            const double cm1 = __dsub_rn(clamp, 1.0);
            double s = __dmul_rn(cm1, cm1);
            // s / (clamp^2 p_z^2 - s (gamma^2 - 1)), evaluated left to right like the reference (:58-59)
            const double den = __dsub_rn(__dmul_rn(__dmul_rn(__dmul_rn(clamp, clamp), pz), pz), __dmul_rn(s, __dsub_rn(__dmul_rn(g, g), 1.0)));
            s = __dsqrt_rn(__ddiv_rn(s, den));
            const double a = __dmul_rn(p1[i], s), b = __dmul_rn(p2[i], s), c = __dmul_rn(pz, s);
            p1[i] = a; p2[i] = b; p3[i] = c;
            gamma[i] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(a, a)), __dmul_rn(b, b)), __dmul_rn(c, c)));
        }
    }
}

// part2d_subcyc%get_exp_fac_max: synchronises (the host needs the value to choose the number of sub-steps)
extern "C" int qpg_part2d_exp_fac_max(qpg_part2d p, double *exp_fac_max)
{
    ARG_TRY(p && exp_fac_max, "null arg");
    qpg_ctx c = p->ctx;
    *exp_fac_max = 1.0;
    if (p->npp_hi == 0) return 0;
    unsigned long long *d_out = (unsigned long long *)(p->d_npp + 2);   // d_npp is a 4-int block: [0] npp, [1] scratch, [2..3] this word
    CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), c->stream));
    const long work = p->npp_hi < p->npmax ? p->npp_hi : p->npmax;
    int grid = (int)((work + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    k_exp_fac_max<<<grid, 256, 0, c->stream>>>(p->gamma, p->p3, p->d_npp, d_out);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    unsigned long long bits = 0ull;
    CUDA_TRY(cudaMemcpyAsync(&bits, d_out, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (bits != 0ull) memcpy(exp_fac_max, &bits, sizeof(double));       // no live particle: 1.0 (:43)
    return 0;
}

// part2d_subcyc%clamp_exp_fac
extern "C" int qpg_part2d_clamp_exp_fac(qpg_part2d p, double exp_fac_clamped)
{
    ARG_TRY(p && exp_fac_clamped > 1.0, "null handle or clamp <= 1");
    qpg_ctx c = p->ctx;
    if (p->npp_hi == 0) return 0;
    const long work = p->npp_hi < p->npmax ? p->npp_hi : p->npmax;
    int grid = (int)((work + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    k_clamp_exp_fac<<<grid, 256, 0, c->stream>>>(p->p1, p->p2, p->p3, p->gamma, p->d_npp, exp_fac_clamped);
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// simulation_subcyc_class.f03:431-451: number of sub-steps and their length for a slice whose largest expansion factor is exp_fac
extern "C" int qpg_subcyc_step(double exp_fac, double exp_fac_max, double dt, double dt_min, double *dt_subcyc, int *n_subcyc)
{
    ARG_TRY(dt_subcyc && n_subcyc && exp_fac_max > 0.0 && dt > 0.0, "bad arg");
    if (exp_fac > exp_fac_max) {
        *n_subcyc = (int)ceil(exp_fac / exp_fac_max);
        *dt_subcyc = dt / *n_subcyc;
        if (*dt_subcyc < dt_min) { *n_subcyc = (int)floor(dt / dt_min); *dt_subcyc = dt / *n_subcyc; }
    } else { *n_subcyc = 1; *dt_subcyc = dt; }
    return 0;
}

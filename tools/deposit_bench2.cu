// deposit_bench2.cu -- development tool: which part of warp_deposit_mma costs what (knock-out variants of a local copy)
#include <cuda_runtime.h>
#include <cstdio>
#define FULL 0xffffffffu
#define DEP_LD 36
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
enum { NO_MATCH = 1, NO_RED = 2, NO_DMMA = 4, NO_STS = 8, NO_ALDS = 16, NO_BLDS = 32, NO_SKIP = 64 };
template <int KO>
__device__ __forceinline__ void dep(const double (&alpha)[6], const double (&beta)[8], int key, double *acc8, double *tile, int lane)
{
    constexpr int P = 3, R = 6;
    __syncwarp();
    if (!(KO & NO_STS)) {
#pragma unroll
        for (int r = 0; r < R; r++) tile[r * DEP_LD + lane] = alpha[r];
#pragma unroll
        for (int k = 0; k < 8; k++) tile[(R + k) * DEP_LD + lane] = beta[k];
    }
    __syncwarp();
    const int row = lane >> 2, kk = lane & 3;
    double bfr[8];
#pragma unroll
    for (int s = 0; s < 8; s++) bfr[s] = (KO & NO_BLDS) ? beta[s] : tile[(R + row) * DEP_LD + 4 * s + kk];
    unsigned leaders;
    if (KO & NO_MATCH) leaders = 1u;
    else {
        const unsigned same = __match_any_sync(FULL, key);
        const bool leader = key >= 0 && (__ffs(same) - 1 == lane);
        leaders = __ballot_sync(FULL, leader);
    }
    while (leaders) {
        const int l = __ffs(leaders) - 1;
        leaders &= leaders - 1;
        const int cell = __shfl_sync(FULL, key, l);
        const unsigned memb = __ballot_sync(FULL, key == cell);
        const unsigned members = memb >> kk;
        const int r = row;
        const bool live = r < R;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if ((KO & NO_SKIP) || ((memb >> (4 * s)) & 0xfu)) {
                double a = 0.0;
                if (live && ((members >> (4 * s)) & 1u)) a = (KO & NO_ALDS) ? alpha[s % 6] : tile[r * DEP_LD + 4 * s + kk];
                if (KO & NO_DMMA) { c0 = fma(a, bfr[s], c0); } else dmma884(c0, c1, a, bfr[s]);
            }
        }
        if (live) {
            const int j = r >= P ? 1 : 0, pl = r - j * P;
            double *dst = acc8 + ((size_t)(cell + j) * P + pl) * 8 + 2 * kk;
            if (KO & NO_RED) { if (c0 + c1 == 1.2345e-300) dst[0] = c0; }
            else { red_add(dst, c0); red_add(dst + 1, c1); }
        }
    }
}
template <int KO>
__global__ void k_dep(double *acc8, int iters, int cells)
{
    extern __shared__ double tiles[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *tile = tiles + warp * 14 * DEP_LD;
    double alpha[6], beta[8];
    for (int k = 0; k < 6; k++) alpha[k] = 1e-3 * (lane + k);
    for (int k = 0; k < 8; k++) beta[k] = 1e-2 * (lane - k);
    const int base = 1 + (blockIdx.x * 37 + warp * 5) % 900;
    const int key = base + (lane * cells) / 32;
    for (int it = 0; it < iters; it++) {
        dep<KO>(alpha, beta, key, acc8, tile, lane);
        alpha[0] += 1e-9; beta[3] += 1e-9;
    }
}
template <int KO> void run(const char *name, double *acc8)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 2000;
    for (int cells : {1, 2}) {
        printf("%-40s cells %d:", name, cells);
        for (int warps : {8, 16, 32}) {
            const size_t smem = sizeof(double) * 14 * DEP_LD * warps;
            cudaFuncSetAttribute(k_dep<KO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_dep<KO><<<148, warps * 32, smem>>>(acc8, 10, cells);
            cudaEventRecord(a); k_dep<KO><<<148, warps * 32, smem>>>(acc8, iters, cells); cudaEventRecord(b); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("  %2dw %6.1f", warps, ms * 1e-3 * 1.965e9 / ((double)iters * warps));
        }
        printf("  SM-cycles/tile (%s)\n", cudaGetErrorString(cudaGetLastError()));
    }
}
int main()
{
    double *acc8; cudaMalloc(&acc8, sizeof(double) * 1100 * 3 * 8); cudaMemset(acc8, 0, sizeof(double) * 1100 * 3 * 8);
    run<0>("full", acc8);
    run<NO_MATCH>("no match (1 cell assumed)", acc8);
    run<NO_RED>("no RED", acc8);
    run<NO_DMMA>("DFMA instead of DMMA", acc8);
    run<NO_STS>("no STS", acc8);
    run<NO_ALDS>("no A-fragment LDS", acc8);
    run<NO_BLDS>("no B-fragment LDS", acc8);
    run<NO_SKIP>("no k-step skipping", acc8);
    run<NO_STS | NO_ALDS | NO_BLDS>("no shared memory at all", acc8);
    run<NO_STS | NO_ALDS | NO_BLDS | NO_RED>("DMMA + match only", acc8);
    run<NO_STS | NO_ALDS | NO_BLDS | NO_RED | NO_MATCH>("DMMA only", acc8);
    run<NO_DMMA | NO_RED | NO_MATCH>("shared memory only", acc8);
    return 0;
}

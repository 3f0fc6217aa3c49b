// pipe_bench.cu -- development tool: SM-wide throughput of the instruction classes the particle kernels are made of (B200):
// DFMA, DMMA.8x8x4, LDS.64, STS.64, LDG.64 (L1-resident, warp-uniform and per-lane addresses), SHFL, and pairwise mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_bench.cu -o tools/pipe_bench
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
enum { DFMA = 1, DMMA = 2, LDS = 4, STS = 8, LDGU = 16, LDGL = 32, SHFL = 64 };
// one "unit" per iteration: 32 DFMA (8 chains) | 8 DMMA (4 chains) | 16 LDS.64 | 16 STS.64 | 16 LDG.64 | 16 SHFL.32
template <int MIX>
__global__ void k_mix(double *out, const double *gbuf, int iters)
{
    __shared__ double sm[32 * 33 * 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *mine = sm + (warp % 2) * 32 * 33;
    double f[8], c[8], acc = 0.0;
    for (int k = 0; k < 8; k++) { f[k] = 1.0 + 1e-9 * (threadIdx.x + k); c[k] = 0.0; }
    const double a = 1.0 + 1e-12 * lane, b = 1e-7;
    int idx = lane;
    for (int i = 0; i < iters; i++) {
        if (MIX & DFMA) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int k = 0; k < 8; k++) f[k] = fma(f[k], a, b);
        }
        if (MIX & DMMA) {
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int k = 0; k < 4; k++) dmma(c[2 * k], c[2 * k + 1], f[k], a);
        }
        if (MIX & LDS) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc += mine[k * 33 + ((lane + idx) & 31)];
        }
        if (MIX & STS) {
#pragma unroll
            for (int k = 0; k < 16; k++) mine[k * 33 + lane] = f[k & 7];
        }
        if (MIX & LDGU) {   // warp-uniform address per instruction (every lane reads the same double), L1 resident
#pragma unroll
            for (int k = 0; k < 16; k++) acc += __ldg(gbuf + ((idx >> 5) & 63) * 16 + k);
        }
        if (MIX & LDGL) {   // consecutive lanes read consecutive doubles (one 256-byte run), L1 resident
#pragma unroll
            for (int k = 0; k < 16; k++) acc += __ldg(gbuf + k * 32 + ((lane + idx) & 31));
        }
        if (MIX & SHFL) {
#pragma unroll
            for (int k = 0; k < 16; k++) idx += __shfl_xor_sync(0xffffffffu, idx, 1 + (k & 15));
        }
        idx += (int)(acc != 1.2345);    // keeps the loads in the loop
    }
    double s = acc + idx;
    for (int k = 0; k < 8; k++) s += f[k] + c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MIX> void run(const char *name, double units, double *out, double *gbuf)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000;
    printf("%-28s", name);
    for (int warps : {4, 8, 16, 32}) {
        k_mix<MIX><<<148, warps * 32>>>(out, gbuf, 10);
        cudaEventRecord(a); k_mix<MIX><<<148, warps * 32>>>(out, gbuf, iters); cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * warps);   // SM cycles per warp-iteration
        printf("  %2dw: %6.1f cyc/iter", warps, cyc);
    }
    printf("   (%s)\n", cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    double *out, *gbuf; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&gbuf, 8192 * 8); cudaMemset(gbuf, 0, 8192 * 8);
    printf("SM cycles per warp-iteration (per SM); one iteration = 32 DFMA | 8 DMMA | 16 LDS.64 | 16 STS.64 | 16 LDG.64 | 16 SHFL\n");
    run<DFMA>("32 DFMA", 32, out, gbuf);
    run<DMMA>("8 DMMA", 8, out, gbuf);
    run<LDS>("16 LDS.64", 16, out, gbuf);
    run<STS>("16 STS.64", 16, out, gbuf);
    run<LDGU>("16 LDG.64 uniform addr", 16, out, gbuf);
    run<LDGL>("16 LDG.64 coalesced", 16, out, gbuf);
    run<SHFL>("16 SHFL.32", 16, out, gbuf);
    run<DFMA | DMMA>("32 DFMA + 8 DMMA", 0, out, gbuf);
    run<DFMA | LDS>("32 DFMA + 16 LDS", 0, out, gbuf);
    run<DMMA | LDS>("8 DMMA + 16 LDS", 0, out, gbuf);
    run<DMMA | STS>("8 DMMA + 16 STS", 0, out, gbuf);
    run<DMMA | LDGL>("8 DMMA + 16 LDG", 0, out, gbuf);
    run<LDS | LDGL>("16 LDS + 16 LDG", 0, out, gbuf);
    run<LDS | SHFL>("16 LDS + 16 SHFL", 0, out, gbuf);
    run<DMMA | SHFL>("8 DMMA + 16 SHFL", 0, out, gbuf);
    run<DFMA | DMMA | LDS | STS | LDGL>("all but SHFL", 0, out, gbuf);
    return 0;
}

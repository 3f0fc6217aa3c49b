// dfma_operands_bench.cu -- development tool: DFMA issue rate on B200 against the operand pattern.
//   REUSE: c = fma(c, a, b) with a, b fixed (two operands come from the reuse cache)            -- what peak-FLOP microbenchmarks measure
//   MIX3:  r[k] = fma(r[k+1], r[k+2], r[k+3]) over a ring of 16 registers (three distinct, changing register operands per instruction,
//          each result consumed 13 instructions later: no dependency stall)                      -- what real arithmetic looks like
//   MIX2:  r[k] = fma(r[k+1], a, r[k+3])  (one fixed operand)
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(double *out, int iters)
{
    double r[16];
    for (int i = 0; i < 16; i++) r[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (MODE == 0) r[k] = fma(r[k], a, b);
                if (MODE == 1) r[k] = fma(r[(k + 1) & 15], r[(k + 2) & 15], r[(k + 3) & 15]);
                if (MODE == 2) r[k] = fma(r[(k + 1) & 15], a, r[(k + 3) & 15]);
                if (MODE == 3) r[k] = r[(k + 1) & 15] * r[(k + 2) & 15];
                if (MODE == 4) r[k] = r[(k + 1) & 15] + r[(k + 2) & 15];
            }
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double *out)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000;
    printf("%-46s", name);
    for (int warps : {4, 8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, 10);
        cudaEventRecord(a); k<MODE><<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("  %2dw: %5.2f", warps, (double)iters * 64 * warps / (ms * 1e-3 * 1.965e9));
    }
    printf("   warp-instructions per clock per SM\n");
}
int main()
{
    double *out; cudaMalloc(&out, 148 * 1024 * 8);
    run<0>("DFMA c = fma(c, a, b)            (reuse)", out);
    run<1>("DFMA r[k] = fma(r[k+1], r[k+2], r[k+3])", out);
    run<2>("DFMA r[k] = fma(r[k+1], a, r[k+3])", out);
    run<3>("DMUL r[k] = r[k+1] * r[k+2]", out);
    run<4>("DADD r[k] = r[k+1] + r[k+2]", out);
    return 0;
}

// dmma_bench.cu -- fp64 tensor (mma.sync m8n8k4) vs DFMA issue rate on B200 (development tool)
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CHAINS> __global__ void k_dmma(double *out, int iters)
{
    double c[CHAINS][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) { c[k][0] = k; c[k][1] = -k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) dmma(c[k][0], c[k][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS> __global__ void k_dfma(double *out, int iters)
{
    double c[CHAINS];
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-6;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) c[k] = k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) c[k] = fma(c[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; cudaMalloc(&out, 148 * 1024 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        float ms;
        k_dmma<8><<<148, warps * 32>>>(out, 10);
        cudaEventRecord(a); k_dmma<8><<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, a, b);
        double n = 148.0 * warps * iters * 8;
        printf("DMMA m8n8k4: %2d warps/SM: %.2f ns per warp-instr per SM  -> %.1f cycles@1.965GHz per DMMA per SM, %.2f TFLOP/s\n", warps, ms * 1e6 / (n / 148), ms * 1e6 / (n / 148) * 1.965, n * 512 / (ms * 1e-3) / 1e12);
        k_dfma<8><<<148, warps * 32>>>(out, 10);
        cudaEventRecord(a); k_dfma<8><<<148, warps * 32>>>(out, iters); cudaEventRecord(b); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, a, b);
        printf("DFMA       : %2d warps/SM: %.2f ns per warp-instr per SM  -> %.1f cycles per DFMA per SM, %.2f TFLOP/s\n", warps, ms * 1e6 / (n / 148), ms * 1e6 / (n / 148) * 1.965, n * 64 / (ms * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

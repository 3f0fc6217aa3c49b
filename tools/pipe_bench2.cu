// pipe_bench2.cu -- development tool: SM-wide cost of load flavours on B200 (L1 / shared resident), inline PTX so nothing is hoisted
#include <cuda_runtime.h>
#include <cstdio>
enum { LDG64 = 0, LDG128, LDG64U, LDG128U, LDG64NC, LDG128NC_U, LDS64, LDS128, LDS64U, LDS128U, STS64, STS128, LDG64_2ADDR, LDG128_2ADDR };
template <int K>
__global__ void k_ld(double *out, const double *gbuf, int iters)
{
    __shared__ __align__(16) double sm[4096];
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < 4096; k += blockDim.x) sm[k] = k;
    __syncthreads();
    double acc = 0.0;
    unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    for (int i = 0; i < iters; i++) {
        const int r = (i & 7) * 64;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            double x = 0, y = 0;
            if (K == LDG64) asm volatile("ld.global.f64 %0, [%1];" : "=d"(x) : "l"(gbuf + r + k * 32 + lane));
            if (K == LDG128) asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(gbuf + r + k * 64 + 2 * lane));
            if (K == LDG64U) asm volatile("ld.global.f64 %0, [%1];" : "=d"(x) : "l"(gbuf + r + k));
            if (K == LDG128U) asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(gbuf + r + 2 * k));
            if (K == LDG64NC) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(x) : "l"(gbuf + r + k * 32 + lane));
            if (K == LDG128NC_U) asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(gbuf + r + 2 * k));
            if (K == LDG64_2ADDR) asm volatile("ld.global.f64 %0, [%1];" : "=d"(x) : "l"(gbuf + r + k + (lane >> 4) * 9));       // two cells per warp
            if (K == LDG128_2ADDR) asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(gbuf + r + 2 * k + (lane >> 4) * 18));
            if (K == LDS64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(sbase + 8 * (r + k * 32 + lane)));
            if (K == LDS128) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(sbase + 8 * (r + k * 64 + 2 * lane)));
            if (K == LDS64U) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(sbase + 8 * (r + k)));
            if (K == LDS128U) asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(sbase + 8 * (r + 2 * k)));
            if (K == STS64) asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbase + 8 * (r + k * 32 + lane)), "d"(acc));
            if (K == STS128) asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(sbase + 8 * (r + k * 64 + 2 * lane)), "d"(acc), "d"(acc));
            acc += x + y;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int K> void run(const char *name, double *out, double *gbuf)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000;
    printf("%-44s", name);
    for (int warps : {4, 8, 16, 32}) {
        k_ld<K><<<148, warps * 32>>>(out, gbuf, 10);
        cudaEventRecord(a); k_ld<K><<<148, warps * 32>>>(out, gbuf, iters); cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("  %2dw: %5.2f", warps, ms * 1e-3 * 1.965e9 / ((double)iters * warps * 16));
    }
    printf("   SM-cycles per warp-instruction (%s)\n", cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    double *out, *gbuf; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&gbuf, 8192 * 8); cudaMemset(gbuf, 0, 8192 * 8);
    run<LDG64>("LDG.64  coalesced (256 B / warp)", out, gbuf);
    run<LDG128>("LDG.128 coalesced (512 B / warp)", out, gbuf);
    run<LDG64U>("LDG.64  one address per warp", out, gbuf);
    run<LDG128U>("LDG.128 one address per warp", out, gbuf);
    run<LDG64_2ADDR>("LDG.64  two addresses per warp", out, gbuf);
    run<LDG128_2ADDR>("LDG.128 two addresses per warp", out, gbuf);
    run<LDG64NC>("LDG.64.nc coalesced", out, gbuf);
    run<LDG128NC_U>("LDG.128.nc one address per warp", out, gbuf);
    run<LDS64>("LDS.64  conflict-free (256 B / warp)", out, gbuf);
    run<LDS128>("LDS.128 conflict-free (512 B / warp)", out, gbuf);
    run<LDS64U>("LDS.64  broadcast (one address)", out, gbuf);
    run<LDS128U>("LDS.128 broadcast (one address)", out, gbuf);
    run<STS64>("STS.64  conflict-free", out, gbuf);
    run<STS128>("STS.128 conflict-free", out, gbuf);
    return 0;
}

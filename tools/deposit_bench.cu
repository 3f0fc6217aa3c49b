// deposit_bench.cu -- development tool: throughput of the warp-level deposit reductions of particles.cu in isolation
// (synthetic alpha / beta from registers), per SM, against warps per SM and cells per warp; plus DMMA chain micro-cases.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/deposit_bench.cu -o tools/deposit_bench -lcuda
#include "../qpad_b200/csrc/lib.cu"

template <int M, int MODE>
__global__ void k_dep(double *acc8, double *acc1, int iters, int cells, double *sink)
{
    extern __shared__ double tiles[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *tile = tiles + warp * DepTile<M>::doubles;
    constexpr int P = 2 * M + 1;
    double alpha[2 * P], beta[8];
    for (int k = 0; k < 2 * P; k++) alpha[k] = 1e-3 * (lane + k);
    for (int k = 0; k < 8; k++) beta[k] = 1e-2 * (lane - k);
    const int base = 1 + (blockIdx.x * 37 + warp * 5) % 900;
    const int key = base + (lane * cells) / 32;       // `cells` distinct cells per warp, contiguous lanes
    double s = 0.0;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) warp_deposit_mma<M>(alpha, beta, key, acc8, tile, lane);
        else if (MODE == 1) warp_deposit_q_mma<M>(alpha, key, acc1, tile, lane);
        else if (MODE == 2) {   // 8 dependent DMMAs on register operands
            double c0 = 0, c1 = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) dmma884(c0, c1, alpha[q % (2 * P)], beta[q]);
            s += c0 + c1;
        } else if (MODE == 3) { // 2 x 4 dependent DMMAs
            double c0 = 0, c1 = 0, d0 = 0, d1 = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { dmma884(c0, c1, alpha[q], beta[q]); dmma884(d0, d1, alpha[q + 1], beta[q + 4]); }
            s += c0 + c1 + d0 + d1;
        } else if (MODE == 4) { // 228 DFMAs in 4 chains (the Boris part's fp64 load), no DMMA
            double a0 = alpha[0], a1 = alpha[1], a2 = alpha[2], a3 = alpha[3];
#pragma unroll
            for (int q = 0; q < 57; q++) { a0 = fma(a0, beta[0], beta[1]); a1 = fma(a1, beta[2], beta[3]); a2 = fma(a2, beta[4], beta[5]); a3 = fma(a3, beta[6], beta[7]); }
            s += a0 + a1 + a2 + a3;
        } else if (MODE == 5) { // both: 228 DFMA + 8 dependent DMMA
            double a0 = alpha[0], a1 = alpha[1], a2 = alpha[2], a3 = alpha[3];
#pragma unroll
            for (int q = 0; q < 57; q++) { a0 = fma(a0, beta[0], beta[1]); a1 = fma(a1, beta[2], beta[3]); a2 = fma(a2, beta[4], beta[5]); a3 = fma(a3, beta[6], beta[7]); }
            double c0 = 0, c1 = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) dmma884(c0, c1, a0 + q, a1);
            s += a2 + a3 + c0 + c1;
        }
        alpha[0] += 1e-9;
    }
    if (s == 1.234e-300) sink[0] = s;
}

template <int MODE> static void run(const char *name, double *acc8, double *acc1, double *sink, int cells)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 2000;
    for (int warps : {4, 8, 16, 32}) {
        const size_t smem = sizeof(double) * DepTile<1>::doubles * warps;
        cudaFuncSetAttribute(k_dep<1, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_dep<1, MODE><<<148, warps * 32, smem>>>(acc8, acc1, 10, cells, sink);
        cudaEventRecord(a);
        k_dep<1, MODE><<<148, warps * 32, smem>>>(acc8, acc1, iters, cells, sink);
        cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double per_sm_tile_ns = ms * 1e6 / ((double)iters * warps);
        printf("%-34s cells/warp %d warps/SM %2d: %7.1f ns per tile per SM = %6.0f SM-cycles/tile (%5.0f SMSP-cycles per tile)\n", name, cells, warps, per_sm_tile_ns,
               per_sm_tile_ns * 1.965, per_sm_tile_ns * 1.965 * 4);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    double *acc8, *acc1, *sink;
    cudaMalloc(&acc8, sizeof(double) * 1100 * 3 * 8); cudaMalloc(&acc1, sizeof(double) * 1100 * 3); cudaMalloc(&sink, 64);
    cudaMemset(acc8, 0, sizeof(double) * 1100 * 3 * 8); cudaMemset(acc1, 0, sizeof(double) * 1100 * 3);
    for (int cells : {1, 2, 4}) run<0>("warp_deposit_mma<1> (amj)", acc8, acc1, sink, cells);
    for (int cells : {1, 2, 4}) run<1>("warp_deposit_q_mma<1> (qdep)", acc8, acc1, sink, cells);
    run<2>("8 dependent DMMA", acc8, acc1, sink, 1);
    run<3>("2 chains x 4 DMMA", acc8, acc1, sink, 1);
    run<4>("228 DFMA in 4 chains", acc8, acc1, sink, 1);
    run<5>("228 DFMA + 8 dependent DMMA", acc8, acc1, sink, 1);
    return 0;
}

// barrier_bench.cu -- cost of a hand-rolled grid barrier on B200 under different fence flavours (development tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/barrier_bench tools/barrier_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned ldv(const unsigned *p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned lda(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ldr(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_rel(unsigned *p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void red_rlx(unsigned *p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }

template <int MODE> __device__ __forceinline__ void gbar(unsigned *bar, unsigned &ep, unsigned n)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        ep += n;
        if (MODE == 0) { __threadfence(); atomicAdd(bar, 1u); while ((int)(ldv(bar) - ep) < 0) {} __threadfence(); }
        if (MODE == 1) { red_rel(bar); while ((int)(lda(bar) - ep) < 0) {} }
        if (MODE == 2) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); red_rlx(bar); while ((int)(ldr(bar) - ep) < 0) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
    }
    __syncthreads();
}
// WORK: 0 = nothing between barriers; 1 = every thread does a few fp64 REDs + plain stores (like a deposit phase)
template <int MODE, int WORK> __global__ void __launch_bounds__(512, 1) k(unsigned *bar, int iters, double *acc, double *buf, long long *out)
{
    unsigned ep = 0;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (WORK) {
            for (int k2 = 0; k2 < 8; k2++) atomicAdd(acc + ((gtid >> 5) * 8 + k2) % 3072, 1.0);
            buf[gtid] = (double)i;
        }
        gbar<MODE>(bar, ep, gridDim.x);
    }
    if (gtid == 0) out[0] = clock64() - t0;
}
template <int MODE, int WORK> static void run(const char *name, int grid, unsigned *bar, double *acc, double *buf, long long *out)
{
    int iters = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(bar, 0, 4);
        void *args[] = {&bar, &iters, &acc, &buf, &out};
        cudaEventRecord(a);
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)k<MODE, WORK>, dim3(grid), dim3(512), args, 0, 0);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep) printf("%-40s grid %3d : %.3f us per barrier (%s)\n", name, grid, ms * 1e3 / iters, cudaGetErrorString(e));
    }
}
int main()
{
    unsigned *bar; double *acc, *buf; long long *out;
    cudaMalloc(&bar, 512); cudaMalloc(&acc, 3072 * 8); cudaMalloc(&buf, 148 * 512 * 8); cudaMalloc(&out, 64);
    cudaMemset(acc, 0, 3072 * 8);
    for (int grid : {8, 148}) {
        run<0, 0>("threadfence + atomicAdd + volatile", grid, bar, acc, buf, out);
        run<1, 0>("red.release + ld.acquire", grid, bar, acc, buf, out);
        run<2, 0>("fence.acq_rel + relaxed", grid, bar, acc, buf, out);
        run<0, 1>("threadfence ... with RED/store work", grid, bar, acc, buf, out);
        run<1, 1>("red.release ... with RED/store work", grid, bar, acc, buf, out);
        run<2, 1>("fence.acq_rel ... with RED/store work", grid, bar, acc, buf, out);
    }
    return 0;
}

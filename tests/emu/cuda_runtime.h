// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.  A host stand-in for the CUDA runtime + the CUDA C++ language extensions,
// just large enough to compile the *simple* kernels of qpad_b200/csrc (no warp shuffles, no inline PTX, no cooperative launch)
// with g++ and run them on the CPU: the threads of a CTA are ucontext fibers on ONE OS thread (a __syncthreads() is a yield to the
// round-robin scheduler, so barriers are exact and runs are deterministic), CTAs run one after the other, "device memory" is host
// memory and streams are no-ops.  It exists so that device code written when no GPU time is left (neutral.cu, subcyc.cu, vpot.cu,
// diag.cu) can be checked against the oracle before its first run on a B200.  It is found before the real <cuda_runtime.h>
// because tests/emu is the first -I directory of the emulation build (tests/emu/build.py); nothing in the product uses it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>
#include <vector>

#define QPG_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct emu_stream_s *cudaStream_t;
typedef struct emu_event_s *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

using std::max;
using std::min;
// IEEE round-to-nearest intrinsics: the emulation build is compiled with -ffp-contract=off, so plain operators are exact
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline long long __double_as_longlong(double a) { long long r; memcpy(&r, &a, 8); return r; }
inline double __longlong_as_double(long long a) { double r; memcpy(&r, &a, 8); return r; }
// one OS thread runs every fiber: "atomics" are plain read-modify-writes
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }

namespace emu {
struct Fiber { ucontext_t ctx; bool done; };
inline ucontext_t g_sched;
inline Fiber *g_cur = nullptr;
inline std::function<void()> *g_body = nullptr;
inline long g_launches = 0, g_barriers = 0;
inline void tramp() { (*g_body)(); g_cur->done = true; swapcontext(&g_cur->ctx, &g_sched); }
inline void sync_threads() { g_barriers++; swapcontext(&g_cur->ctx, &g_sched); }
template <class F> void launch(dim3 g, dim3 b, F f)
{
    std::function<void()> body = f;
    g_body = &body; g_launches++;
    gridDim = g; blockDim = b;
    const unsigned nt = b.x * b.y * b.z;
    const size_t stack = 256 * 1024;
    static std::vector<char> stacks;
    if (stacks.size() < stack * nt) stacks.resize(stack * nt);
    std::vector<Fiber> fib(nt);
    for (unsigned bz = 0; bz < g.z; bz++) for (unsigned by = 0; by < g.y; by++) for (unsigned bx = 0; bx < g.x; bx++) {
        blockIdx = uint3{bx, by, bz};
        for (unsigned t = 0; t < nt; t++) {
            getcontext(&fib[t].ctx);
            fib[t].ctx.uc_stack.ss_sp = stacks.data() + stack * t; fib[t].ctx.uc_stack.ss_size = stack; fib[t].ctx.uc_link = &g_sched;
            fib[t].done = false;
            makecontext(&fib[t].ctx, (void (*)())tramp, 0);
        }
        unsigned live = nt;
        while (live) {                                    // one round = one barrier phase of the CTA
            unsigned finished = 0;
            for (unsigned t = 0; t < nt; t++) {
                if (fib[t].done) continue;
                threadIdx = uint3{t % b.x, (t / b.x) % b.y, t / (b.x * b.y)};
                g_cur = &fib[t];
                swapcontext(&g_sched, &fib[t].ctx);
                if (fib[t].done) finished++;
            }
            // CUDA requires every thread of a CTA to reach the same barriers: a round in which some threads ended and others
            // are still waiting at a barrier is legal only if those others end without another barrier -- not checked here
            live -= finished;
        }
    }
}
}  // namespace emu
#define __syncthreads() emu::sync_threads()

// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.  A host stand-in for the CUDA runtime + the CUDA C++ language extensions,
// large enough to compile the per-routine kernels of qpad_b200/csrc (no cooperative launch, no CUDA graphs, no peer memory; the few
// inline-PTX statements have a QPG_EMU branch) with g++ and run them on the CPU: the threads of a CTA are fibers on ONE OS thread;
// __syncthreads() and the warp collectives (__shfl_*_sync, __ballot_sync, __match_any_sync, __syncwarp, the emulated DMMA) are
// yields to a scheduler that releases a barrier / a warp collective only when every participant has arrived, so synchronisation is
// exact, a missing participant is reported as a deadlock, and runs are deterministic.  CTAs run one after the other, "device
// memory" is host memory and streams are no-ops.  It exists so that device code written when no GPU time is left (neutral.cu, subcyc.cu, vpot.cu,
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
#include <cstdio>
#include <map>
#include <numeric>
#include <sys/mman.h>
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
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }   // not a clock: a non-zero constant
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount, cudaDevAttrMaxSharedMemoryPerBlockOptin, cudaDevAttrCooperativeLaunch };
// the emulated device: 12 SMs by default (a cooperative launch keeps 512 fibers per CTA alive), 227 KB of shared memory, cooperative launch
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int) { *v = a == cudaDevAttrMultiProcessorCount ? (getenv("QPAD_EMU_SMS") ? atoi(getenv("QPAD_EMU_SMS")) : 12) : (a == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 232448 : 1); return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// ---- compile-only stubs: CUDA graphs, cooperative launch, occupancy (sim.cu links; these paths fail loudly in emulation) ----------
enum { cudaErrorNotSupported = 801 };
typedef struct emu_graph_s *cudaGraph_t;
typedef struct emu_graphexec_s *cudaGraphExec_t;
typedef struct emu_graphnode_s *cudaGraphNode_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone };
enum cudaGraphNodeType { cudaGraphNodeTypeConditional };
enum cudaGraphConditionalNodeType { cudaGraphCondTypeWhile };
enum { cudaGraphCondAssignDefault = 1, cudaStreamSetCaptureDependencies = 1 };
typedef unsigned long long cudaGraphConditionalHandle;
struct cudaConditionalNodeParams { cudaGraphConditionalHandle handle; cudaGraphConditionalNodeType type; unsigned size; cudaGraph_t *phGraph_out; };
struct cudaGraphNodeParams { cudaGraphNodeType type; cudaConditionalNodeParams conditional; };
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamGetCaptureInfo(cudaStream_t, cudaStreamCaptureStatus *, unsigned long long *, cudaGraph_t *, const cudaGraphNode_t **, size_t *) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphConditionalHandleCreate(cudaGraphConditionalHandle *, cudaGraph_t, unsigned, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphAddNode(cudaGraphNode_t *, cudaGraph_t, const cudaGraphNode_t *, size_t, cudaGraphNodeParams *) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamUpdateCaptureDependencies(cudaStream_t, cudaGraphNode_t *, size_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamBeginCaptureToGraph(cudaStream_t, cudaGraph_t, const cudaGraphNode_t *, const void *, size_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaLaunchCooperativeKernel(const void *, dim3, dim3, void **, size_t, cudaStream_t) { return cudaErrorNotSupported; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }
inline void cudaGraphSetConditional(cudaGraphConditionalHandle, unsigned) {}
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline void __stcg(T *p, T v) { *p = v; }
#define __cluster_dims__(...)
struct uint4 { unsigned x, y, z, w; };
inline long long g_emu_clock = 0;
inline long long clock64() { return ++g_emu_clock; }          // a counter, not a clock: monotonic, one tick per reading
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) {}
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
inline void emu_unsupported_ptx() { fprintf(stderr, "emu: this inline-PTX statement is compile-only in the emulation\n"); abort(); }
template <class T> inline T __ldg(const T *p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
// position of the offset-th set bit of mask counting upwards from bit `base` (offset >= 1), 0xffffffff if there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset)
{
    if (offset == 0) return ((mask >> base) & 1u) ? base : 0xffffffffu;
    if (offset > 0) { for (unsigned k = base; k < 32; k++) if (((mask >> k) & 1u) && --offset == 0) return k; return 0xffffffffu; }
    for (int k = (int)base; k >= 0; k--) if (((mask >> k) & 1u) && ++offset == 0) return (unsigned)k;
    return 0xffffffffu;
}
inline int __double2hiint(double a) { unsigned long long r; memcpy(&r, &a, 8); return (int)(r >> 32); }
inline int __double2loint(double a) { unsigned long long r; memcpy(&r, &a, 8); return (int)(r & 0xffffffffu); }
inline double __hiloint2double(int hi, int lo) { unsigned long long r = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double a; memcpy(&a, &r, 8); return a; }
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
// ---- fibers: a minimal x86-64 context switch (no signal-mask system call, unlike swapcontext) ----------------------------------
#if !defined(__x86_64__)
#error "tests/emu needs x86-64 (the fiber switch is written in assembly)"
#endif
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
enum { ST_RUN = 0, ST_BARRIER, ST_WARP, ST_DONE };
struct Fiber { void *sp; int state; unsigned mask; unsigned seq; };
inline void *g_sched_sp = nullptr;
inline Fiber *g_cur = nullptr;
inline std::function<void()> *g_body = nullptr;
inline long g_launches = 0, g_barriers = 0, g_collectives = 0, g_coop_launches = 0, g_polls = 0;
inline unsigned g_tid = 0;                                   // flattened thread index of the running fiber
inline unsigned long long (*g_slots)[2][32] = nullptr;       // [warp][parity][lane] exchange words of the warp collectives
inline unsigned char *dyn_smem = nullptr;                    // `extern __shared__` storage of the running CTA
inline void yield_to_scheduler() { emu_switch(&g_cur->sp, g_sched_sp); }
inline void tramp()
{
    (*g_body)();
    g_cur->state = ST_DONE;
    yield_to_scheduler();
    abort();
}
inline void sync_threads() { g_barriers++; g_cur->state = ST_BARRIER; yield_to_scheduler(); }
// every lane named in `mask` publishes one 64-bit word; returns once all of them have, with the 32 words of the warp in out[]
inline void warp_exchange(unsigned mask, unsigned long long v, unsigned long long out[32])
{
    Fiber *f = g_cur;
    const unsigned tid = g_tid, w = tid >> 5, lane = tid & 31, par = f->seq & 1;
    g_collectives++;
    g_slots[w][par][lane] = v;
    f->mask = mask; f->state = ST_WARP;
    yield_to_scheduler();
    memcpy(out, g_slots[w][par], sizeof(unsigned long long) * 32);
    f->seq++;
}
// per-CTA storage of a static __shared__ variable of a kernel whose CTAs run concurrently (launch_coop); `id` names the variable
inline std::vector<std::map<int, std::vector<unsigned char>>> *g_statics = nullptr;
inline unsigned g_cta = 0;
inline int g_order = getenv("QPAD_EMU_ORDER") ? atoi(getenv("QPAD_EMU_ORDER")) : 0;
inline unsigned g_ord_rot = 0, g_ord_stride = 1;
inline unsigned long long g_ord_state = 0x9E3779B97F4A7C15ull + (unsigned long long)g_order;
inline void *cta_static(size_t bytes, int id)
{
    auto &v = (*g_statics)[g_cta][id];
    if (v.size() < bytes) v.assign(bytes, 0);
    return v.data();
}
// a polling loop (a load that waits for another CTA) hands the processor on but stays runnable
inline void poll_yield() { g_polls++; g_cur->state = ST_RUN; yield_to_scheduler(); }
inline char *fiber_stacks(size_t bytes)
{
    static char *base = nullptr;
    static size_t cap = 0;
    if (bytes > cap) {
        if (base) munmap(base, cap);
        base = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (base == MAP_FAILED) { perror("emu: mmap of the fiber stacks"); abort(); }
        cap = bytes;
    }
    return base;
}
// concurrent = false: CTAs run one after the other (an ordinary launch).  concurrent = true: the fibers of ALL CTAs are alive at
// once and scheduled round-robin -- a cooperative launch whose CTAs synchronise through global memory (their polling loads call
// poll_yield()).
template <class F> void launch_impl(dim3 g, dim3 b, size_t smem, F f, bool concurrent)
{
    std::function<void()> body = f;
    g_body = &body; g_launches++;
    gridDim = g; blockDim = b;
    const unsigned nt = b.x * b.y * b.z, nw = (nt + 31) / 32, ncta = g.x * g.y * g.z, group = concurrent ? ncta : 1;
    const size_t stack = concurrent ? 96 * 1024 : 256 * 1024;
    char *stacks = fiber_stacks(stack * nt * group + 64);
    std::vector<Fiber> fib((size_t)nt * group);
    std::vector<unsigned long long> slots((size_t)nw * group * 64);
    g_slots = (unsigned long long(*)[2][32])slots.data();
    const size_t dyn_each = (smem + 127) & ~(size_t)63;
    std::vector<unsigned char> dyn(dyn_each * group + 64);
    unsigned char *dyn0 = (unsigned char *)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    std::vector<std::map<int, std::vector<unsigned char>>> statics(group);
    g_statics = &statics;
    for (unsigned first = 0; first < ncta; first += group) {
        for (auto &m : statics) m.clear();
        for (size_t t = 0; t < fib.size(); t++) {
            uintptr_t top = ((uintptr_t)stacks + stack * (t + 1)) & ~(uintptr_t)15;
            void **sp = (void **)(top - 8);          // after the `ret` into tramp: rsp % 16 == 8, as at any function entry
            *--sp = (void *)tramp;
            for (int r = 0; r < 6; r++) *--sp = nullptr;
            fib[t] = Fiber{(void *)sp, ST_RUN, 0u, 0u};
        }
        std::vector<unsigned> live_cta(group, nt);
        size_t live = fib.size();
        while (live) {
            bool ran = false, released = false;
            if (g_order >= 2) {                       // a stride coprime to nt and a rotation, new every round
                g_ord_state = g_ord_state * 6364136223846793005ull + 1442695040888963407ull;
                g_ord_rot = (unsigned)((g_ord_state >> 33) % nt);
                g_ord_stride = 1 + 2 * (unsigned)((g_ord_state >> 13) % 64);
                while (std::gcd(g_ord_stride, nt) != 1) g_ord_stride += 2;
            }
            for (unsigned c = 0; c < group; c++) {
                if (!live_cta[c]) continue;
                const unsigned lin = first + c;
                Fiber *fc = &fib[(size_t)c * nt];
                for (unsigned tt = 0; tt < nt; tt++) {
                    // QPAD_EMU_ORDER: the order in which the runnable threads of a CTA get the processor between two synchronisation
                    // points -- 0 ascending (default), 1 descending, >= 2 a pseudo-random rotation + stride per round (seed).  Correctly
                    // synchronised code gives the same results under every order: a cheap racecheck for missing barriers.
                    unsigned t = tt;
                    if (g_order == 1) t = nt - 1 - tt;
                    else if (g_order >= 2) t = (unsigned)(((unsigned long long)tt * g_ord_stride + g_ord_rot) % nt);
                    if (fc[t].state != ST_RUN) continue;
                    blockIdx = uint3{lin % g.x, (lin / g.x) % g.y, lin / (g.x * g.y)};
                    threadIdx = uint3{t % b.x, (t / b.x) % b.y, t / (b.x * b.y)};
                    g_tid = t; g_cta = c; g_cur = &fc[t]; dyn_smem = dyn0 + dyn_each * c;
                    g_slots = (unsigned long long(*)[2][32])slots.data() + (size_t)c * nw;
                    emu_switch(&g_sched_sp, fc[t].sp);
                    ran = true;
                    if (fc[t].state == ST_DONE) { live--; live_cta[c]--; }
                }
                bool rel_c = false;
                for (unsigned w = 0; w < nw; w++) {      // warp collectives: all lanes named in a waiting lane's mask must wait with that mask
                    const unsigned base = w * 32, nl = std::min(32u, nt - base);
                    for (unsigned l = 0; l < nl; l++) {
                        if (fc[base + l].state != ST_WARP) continue;
                        const unsigned m = fc[base + l].mask & (nl == 32 ? 0xffffffffu : ((1u << nl) - 1u));
                        bool all = true;
                        for (unsigned k = 0; k < nl && all; k++) if ((m >> k) & 1u) all = fc[base + k].state == ST_WARP && fc[base + k].mask == fc[base + l].mask;
                        if (all) { for (unsigned k = 0; k < nl; k++) if ((m >> k) & 1u) fc[base + k].state = ST_RUN; rel_c = true; }
                    }
                }
                if (!rel_c && live_cta[c]) {              // CTA barrier: every live thread of the CTA waits at it
                    unsigned at = 0;
                    for (unsigned t = 0; t < nt; t++) at += fc[t].state == ST_BARRIER;
                    if (at == live_cta[c]) { for (unsigned t = 0; t < nt; t++) if (fc[t].state == ST_BARRIER) fc[t].state = ST_RUN; rel_c = true; }
                }
                released |= rel_c;
            }
            if (live && !ran && !released) {
                size_t nb = 0, nwp = 0;
                for (auto &x : fib) { nb += x.state == ST_BARRIER; nwp += x.state == ST_WARP; }
                fprintf(stderr, "emu: DEADLOCK (CTAs %u..%u of %u): %zu threads live, %zu at __syncthreads, %zu in a warp collective\n", first, first + group - 1, ncta, live, nb, nwp);
                abort();
            }
        }
    }
    g_statics = nullptr;
}
template <class F> void launch(dim3 g, dim3 b, size_t smem, F f) { launch_impl(g, b, smem, f, false); }
template <class F> void launch_coop(dim3 g, dim3 b, size_t smem, F f) { g_coop_launches++; launch_impl(g, b, smem, f, true); }
template <class F> void launch(dim3 g, dim3 b, F f) { launch(g, b, 0, f); }
}  // namespace emu
#define __syncthreads() emu::sync_threads()

// ---- warp collectives on top of emu::warp_exchange ------------------------------------------------------------------------------
namespace emu {
template <class T> inline unsigned long long to_bits(T v) { unsigned long long r = 0; static_assert(sizeof(T) <= 8, "warp word"); memcpy(&r, &v, sizeof(T)); return r; }
template <class T> inline T from_bits(unsigned long long r) { T v; memcpy(&v, &r, sizeof(T)); return v; }
template <class T> inline T shfl_from(unsigned mask, T v, int src, bool valid)
{
    unsigned long long all[32];
    warp_exchange(mask, to_bits(v), all);
    return (valid && ((mask >> src) & 1u)) ? from_bits<T>(all[src]) : v;
}
}  // namespace emu
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = emu::g_tid & 31, seg = lane & ~(width - 1);
    return emu::shfl_from(mask, v, seg | (src & (width - 1)), true);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = emu::g_tid & 31, seg = lane & ~(width - 1), src = lane - (int)delta;
    return emu::shfl_from(mask, v, src < seg ? lane : src, src >= seg);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = emu::g_tid & 31, seg = lane & ~(width - 1), src = lane + (int)delta;
    return emu::shfl_from(mask, v, src >= seg + width ? lane : src, src < seg + width);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32)
{
    const int lane = emu::g_tid & 31, src = lane ^ lm;
    return emu::shfl_from(mask, v, src, (src & ~(width - 1)) == (lane & ~(width - 1)));
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    unsigned long long all[32];
    emu::warp_exchange(mask, pred ? 1ull : 0ull, all);
    unsigned r = 0;
    for (int k = 0; k < 32; k++) if (((mask >> k) & 1u) && all[k]) r |= 1u << k;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v)
{
    unsigned long long all[32];
    const unsigned long long mine = emu::to_bits(v);
    emu::warp_exchange(mask, mine, all);
    unsigned r = 0;
    for (int k = 0; k < 32; k++) if (((mask >> k) & 1u) && all[k] == mine) r |= 1u << k;
    return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned long long all[32]; emu::warp_exchange(mask, 0ull, all); }
// mma.sync.aligned.m8n8k4.row.col.f64: D[8x8] += A[8x4] B[4x8]; lane holds a = A[lane / 4][lane % 4], b = B[lane % 4][lane / 4],
// c0, c1 = C[lane / 4][2 (lane % 4) + {0, 1}]
inline void emu_dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    unsigned long long A[32], B[32];
    emu::warp_exchange(0xffffffffu, emu::to_bits(a), A);
    emu::warp_exchange(0xffffffffu, emu::to_bits(b), B);
    const int lane = emu::g_tid & 31, row = lane >> 2, col = (lane & 3) * 2;
    for (int k = 0; k < 4; k++) {
        const double av = emu::from_bits<double>(A[row * 4 + k]);
        c0 = std::fma(av, emu::from_bits<double>(B[col * 4 + k]), c0);
        c1 = std::fma(av, emu::from_bits<double>(B[(col + 1) * 4 + k]), c1);
    }
}

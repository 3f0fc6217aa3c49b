// p2p.cu -- peer-memory hand-offs of the xi-pipeline between GPUs of one box (one process per GPU).
//
// The reference hands the plasma slice state between pipeline stages with MPI isend/recv pairs
// (species/part2d_class.f03:2355-2450 pipesend/piperecv, fields/field_class.f03:560-700 pipe_send/pipe_recv,
// beam/part3d_comm.f03:278-314).  Here a hand-off is a pack kernel of the producer stage that WRITES THE WIRE RECORD
// STRAIGHT INTO THE CONSUMER GPU'S MEMORY over NVLink (the consumer's buffer is mapped into the producer process through a
// CUDA IPC handle), followed by a one-word flag write; the consumer's stream waits for the flag with a stream memory
// operation (cuStreamWaitValue32: no SM, no spinning kernel, no host involvement).  Nothing else is needed because the
// persistent sweep kernels occupy every SM of a stage's partition -- a library send/recv kernel would have to wait for
// free SMs or take them from the sweeps.
//
//   producer, message n:  wait(own ack  >= n-1) -> pack kernels(dst = peer buffer) -> signal(peer ready = n)
//   consumer, message n:  wait(own ready >= n)  -> unpack kernels                  -> signal(peer ack   = n)
//
// Every flag has ONE writer and counts messages monotonically, so there is nothing to reset and no ABA case.
#include "common.cuh"
#include <cuda.h>

// flag write after everything enqueued on the stream so far: the kernel boundary orders the pack kernels' peer writes
// before it, the system-scope fence + release store order them before the flag for the observer on the other GPU
__global__ void k_p2p_signal(unsigned *flag, unsigned value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// fallback wait when the driver entry point for stream memory operations is not available: one polling thread
__global__ void k_p2p_spin(const unsigned *flag, unsigned value)
{
    unsigned v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - value) >= 0) break;
        __nanosleep(200);
    } while (true);
}

// conditional wait: the stream blocks on the flag only if *count != 0 when the kernel runs.  Used for hand-offs whose payload matters
// only to a stage that holds beam particles (the backward e / b guard slice feeds the beam push alone): a stage without beam
// particles does not wait for its downstream neighbour, so the start-up skew of the pipeline (each stage waits for the first slice
// of the next one) accumulates over the beam-carrying stages only.  One polling thread on an SM the stage's own sweep kernel has
// just left; a producer that never signals trips the trap after ~20 s instead of hanging the GPU.
__global__ void k_p2p_spin_unless_empty(const int *count, const unsigned *flag, unsigned value)
{
    if (*count == 0) return;
    const long long t0 = clock64();
    unsigned v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - value) >= 0) break;
        __nanosleep(200);
        if (clock64() - t0 > 40000000000LL) __trap();
    } while (true);
}

typedef CUresult (*fn_wait32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static fn_wait32_t p2p_wait_entry()
{
    static fn_wait32_t fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        if (!getenv("QPG_P2P_SPIN_WAIT")) {
            void *p = nullptr;
            cudaDriverEntryPointQueryResult st;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess) fn = (fn_wait32_t)p;
            else cudaGetLastError();
        }
    }
    return fn;
}

extern "C" int qpg_wire_alloc(void **dev_ptr, long bytes)
{
    ARG_TRY(dev_ptr && bytes > 0, "bad arg");
    CUDA_TRY(cudaMalloc(dev_ptr, (size_t)bytes));     // plain cudaMalloc: the only kind of allocation cudaIpcGetMemHandle accepts
    CUDA_TRY(cudaMemset(*dev_ptr, 0, (size_t)bytes));
    return 0;
}
extern "C" int qpg_wire_free(void *dev_ptr)
{
    if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
    return 0;
}
extern "C" int qpg_wire_export(void *dev_ptr, unsigned char *handle64)
{
    ARG_TRY(dev_ptr && handle64, "null arg");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int qpg_wire_import(const unsigned char *handle64, void **dev_ptr)
{
    ARG_TRY(dev_ptr && handle64, "null arg");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int qpg_wire_unmap(void *dev_ptr)
{
    if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
extern "C" int qpg_stream_signal(void *cuda_stream, unsigned *flag, unsigned value)
{
    ARG_TRY(flag, "null flag");
    k_p2p_signal<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(flag, value);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_stream_wait(void *cuda_stream, unsigned *flag, unsigned value)
{
    ARG_TRY(flag, "null flag");
    if (fn_wait32_t fn = p2p_wait_entry()) {
        CUresult r = fn((CUstream)cuda_stream, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WAIT_VALUE_GEQ);
        if (r == CUDA_SUCCESS) return 0;
        qpg_set_error("cuStreamWaitValue32 failed (%d)", (int)r);
        return QPG_ERR_CUDA;
    }
    k_p2p_spin<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(flag, value);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_stream_wait_unless_empty(void *cuda_stream, const int *dev_count, unsigned *flag, unsigned value)
{
    ARG_TRY(flag && dev_count, "null arg");
    k_p2p_spin_unless_empty<<<1, 1, 0, (cudaStream_t)cuda_stream>>>(dev_count, flag, value);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
extern "C" int qpg_stream_wait_is_memop(void) { return p2p_wait_entry() != nullptr; }

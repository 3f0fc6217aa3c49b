// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the translation unit of the host-emulation library = the
// unity build of qpad_b200/csrc/lib.cu without p2p.cu (peer memory).  fused.cu (thread-block clusters) is compiled so that sim.cu links
// but cannot run here.  Of the slab drivers of qpg_sim the emulation runs the plain per-slice launches (the default here; a request
// for CUDA-graph replay is ignored) and, after qpg_sim_set_sweep(s, 1), the persistent cooperative sweep kernel of sweep.cu: all
// its CTAs are alive at once (emu::launch_coop) and meet at its hand-rolled grid / team barriers and flagged exchange words
// through "global memory", their polling loads yielding to the scheduler.  The included files are the launch-rewritten copies that tests/emu/build.py writes to _build/; the extern "C"
// entry points are therefore the very code that runs on the GPU, with "device memory" on the host heap.
#include <cuda_runtime.h>
#include "fields.cu.cpp"
#include "particles.cu.cpp"
#include "beam.cu.cpp"
#include "laser.cu.cpp"
#include "fused.cu.cpp"
#include "sweep.cu.cpp"
#include "sim.cu.cpp"
#include "neutral.cu.cpp"
#include "subcyc.cu.cpp"
#include "vpot.cu.cpp"
#include "diag.cu.cpp"

extern "C" {
// p2p.cu (CUDA IPC, stream memory operations) is not part of the emulation; the wire buffers and flags of ONE process are plain
// host memory here, and because every enqueued operation has already run when the call returns, a stream-ordered wait for a flag
// either finds it raised or would wait forever (reported as an error)
int qpg_wire_alloc(void **dev_ptr, long bytes) { *dev_ptr = calloc((size_t)(bytes > 0 ? bytes : 1), 1); return *dev_ptr ? 0 : QPG_ERR_ALLOC; }
int qpg_wire_free(void *dev_ptr) { free(dev_ptr); return 0; }
int qpg_wire_export(void *, unsigned char *) { qpg_set_error("qpg_wire_export: CUDA IPC is not emulated"); return QPG_ERR_UNSUPPORTED; }
int qpg_wire_import(const unsigned char *, void **) { qpg_set_error("qpg_wire_import: CUDA IPC is not emulated"); return QPG_ERR_UNSUPPORTED; }
int qpg_wire_unmap(void *) { return 0; }
int qpg_stream_wait_is_memop(void) { return 0; }
int qpg_stream_signal(void *, unsigned *flag, unsigned value) { *flag = value; return 0; }
int qpg_stream_wait(void *, unsigned *flag, unsigned value)
{
    if ((int)(*flag - value) >= 0) return 0;
    qpg_set_error("qpg_stream_wait: flag %u < %u and its producer has not been enqueued -- on a GPU this stream would wait forever", *flag, value);
    return QPG_ERR_STATE;
}
int qpg_stream_wait_unless_empty(void *st, const int *dev_count, unsigned *flag, unsigned value)
{
    if (*dev_count == 0) return 0;
    return qpg_stream_wait(st, flag, value);
}
long emu_launches(void) { return emu::g_launches; }
long emu_barriers(void) { return emu::g_barriers; }
long emu_collectives(void) { return emu::g_collectives; }
long emu_coop_launches(void) { return emu::g_coop_launches; }
long emu_polls(void) { return emu::g_polls; }
}

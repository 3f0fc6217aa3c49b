// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the translation unit of the host-emulation library = the
// unity build of qpad_b200/csrc/lib.cu without sweep.cu, sim.cu, fused.cu and p2p.cu (cooperative launch, CUDA graphs,
// peer memory).  The included files are the launch-rewritten copies that tests/emu/build.py writes to _build/; the extern "C"
// entry points are therefore the very code that runs on the GPU, with "device memory" on the host heap.
#include <cuda_runtime.h>
#include "fields.cu.cpp"
#include "particles.cu.cpp"
#include "beam.cu.cpp"
#include "laser.cu.cpp"
#include "neutral.cu.cpp"
#include "subcyc.cu.cpp"
#include "vpot.cu.cpp"
#include "diag.cu.cpp"

extern "C" {
long emu_launches(void) { return emu::g_launches; }
long emu_barriers(void) { return emu::g_barriers; }
long emu_collectives(void) { return emu::g_collectives; }
}

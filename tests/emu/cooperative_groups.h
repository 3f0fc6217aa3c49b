// TEST INFRASTRUCTURE ONLY: compile-only stand-in (tests/emu).  The cluster kernels of fused.cu are compiled so that sim.cu links,
// but the emulation runs CTAs one after the other, so a kernel that needs a live cluster cannot run: this_cluster() aborts.
#pragma once
#include <cstdio>
#include <cstdlib>
namespace cooperative_groups {
struct cluster_group {
    void sync() {}
    unsigned block_rank() const { return 0; }
    template <class T> T *map_shared_rank(T *p, unsigned) const { return p; }
};
inline cluster_group this_cluster() { fprintf(stderr, "emu: thread-block clusters are not emulated (use qpg_sim_set_fused(s, 0))\n"); abort(); }
}  // namespace cooperative_groups

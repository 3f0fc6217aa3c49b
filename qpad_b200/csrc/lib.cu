// lib.cu -- unity translation unit of libqpadb200.so (kernels in one module; no relocatable device code needed)
#include "fields.cu"
#include "particles.cu"
#include "beam.cu"
#include "laser.cu"
#include "fused.cu"
#include "sweep.cu"
#include "sim.cu"
#include "p2p.cu"

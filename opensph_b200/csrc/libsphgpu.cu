// Single translation unit of libsphgpu.so: the kernels share __constant__ run/material parameters, so the sources
// are compiled together (no relocatable device code needed).
#include "api.cu"
#include "grid.cu"
#include "pair.cu"
#include "pair_tiled.cu"
#include "pair_symmetric.cu"
#include "stepping.cu"
#include "transfer.cu"
#include "halo.cu"
#include "gravity.cu"
#include "lattice.cu"
#include "components.cu"

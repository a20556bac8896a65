// cuda_shim.h -- TEST INFRASTRUCTURE.  Just enough of the CUDA language for g++ to compile a thread-per-element
// kernel that uses no shared memory, no warp intrinsics and no textures (phasta_b200/csrc/boundary.cuh): the
// qualifiers vanish, the built-in index variables become globals that the harness sets before each "thread",
// __ldg is a load and atomicAdd an add (one thread runs at a time).
#pragma once
#include <cmath>
#include <cstdlib>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __constant__ static
#define __launch_bounds__(...)
struct shim_dim3 {
  unsigned x, y, z;
};
static shim_dim3 blockIdx, threadIdx, blockDim, gridDim;
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline double atomicAdd(double *p, double v) {
  const double o = *p;
  *p += v;
  return o;
}

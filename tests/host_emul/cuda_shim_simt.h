// cuda_shim_simt.h -- TEST INFRASTRUCTURE.  A SIMT emulation just big enough to run the product's cooperative
// kernels (shared memory, __syncthreads, __syncwarp, __shfl_sync, atomicAdd) on the host: every CUDA thread of a
// block is a pthread, the barriers are pthread barriers, the built-in index variables are thread-local, dynamic
// shared memory is one static buffer (blocks run one after the other).  Slow and only meant for the few-element
// meshes of the reference-Fortran fixtures.
#pragma once
#include <pthread.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __constant__ static
#define __shared__
#define __align__(n)
#define __launch_bounds__(...)
struct shim_dim3 {
  unsigned x, y, z;
};
struct double2 {
  double x, y;
};
static thread_local shim_dim3 threadIdx, blockIdx;
static shim_dim3 blockDim, gridDim;
unsigned char smem_raw[1 << 18] __attribute__((aligned(16)));   // `extern __shared__ unsigned char smem_raw[]`
static pthread_barrier_t shim_block_bar, shim_warp_bar[32];
static int shim_warp_buf[32][32];

template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() { pthread_barrier_wait(&shim_block_bar); }
static inline void __syncwarp() { pthread_barrier_wait(&shim_warp_bar[threadIdx.x >> 5]); }
static inline int __shfl_sync(unsigned, int v, int src) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  shim_warp_buf[w][l] = v;
  pthread_barrier_wait(&shim_warp_bar[w]);
  const int r = shim_warp_buf[w][src];
  pthread_barrier_wait(&shim_warp_bar[w]);
  return r;
}
static inline double atomicAdd(double *p, double v) {
  uint64_t *q = reinterpret_cast<uint64_t *>(p);
  uint64_t old = __atomic_load_n(q, __ATOMIC_RELAXED), nw;
  double o;
  do {
    memcpy(&o, &old, 8);
    const double n = o + v;
    memcpy(&nw, &n, 8);
  } while (!__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return o;
}

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __sync_synchronize(); }
template <class T>
static inline T __ldcv(const T *p) { return *reinterpret_cast<const volatile T *>(p); }

// run `kernel()` for a grid of `grid` blocks of `block` threads (block a multiple of 32), block after block
struct ShimArg {
  const std::function<void()> *fn;
  unsigned t, b;
};
static void *shim_thread(void *a) {
  ShimArg *s = static_cast<ShimArg *>(a);
  threadIdx = {s->t, 0, 0};
  blockIdx = {s->b, 0, 0};
  (*s->fn)();
  return nullptr;
}
static void shim_launch(unsigned grid, unsigned block, const std::function<void()> &kernel) {
  blockDim = {block, 1, 1};
  gridDim = {grid, 1, 1};
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, 1 << 20);
  for (unsigned b = 0; b < grid; b++) {
    pthread_barrier_init(&shim_block_bar, nullptr, block);
    for (unsigned w = 0; w < block / 32; w++) pthread_barrier_init(&shim_warp_bar[w], nullptr, 32);
    std::vector<pthread_t> th(block);
    std::vector<ShimArg> args(block);
    for (unsigned t = 0; t < block; t++) {
      args[t] = {&kernel, t, b};
      pthread_create(&th[t], &at, shim_thread, &args[t]);
    }
    for (unsigned t = 0; t < block; t++) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&shim_block_bar);
    for (unsigned w = 0; w < block / 32; w++) pthread_barrier_destroy(&shim_warp_bar[w]);
  }
  pthread_attr_destroy(&at);
}

// cuda_runtime.h -- TEST INFRASTRUCTURE (tests/host_emul/fullhost): a stand-in for the CUDA runtime and the SIMT
// execution model that lets g++ build the WHOLE product library (phasta_b200/csrc/*.cu, passed through cu2cpp.py
// for the <<<...>>> launch syntax) for the host, so that the `-m gpu` test suite can be exercised -- host glue and
// kernels -- where there is no GPU.  "Device" memory is host memory, streams are synchronous, and every CUDA thread
// of a block is a fiber (ucontext) on the calling OS thread: __syncthreads / __syncwarp / warp shuffles yield to a
// round-robin scheduler that releases a barrier once every live fiber of the block (warp) has arrived.
// It is never shipped and nothing in phasta_b200/ refers to it.
#pragma once
#include <ucontext.h>
#include <sys/mman.h>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __constant__ static
#define __align__(n)
#define __launch_bounds__(...)

// ---------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef int cudaStream_t;
struct shim_event { std::chrono::steady_clock::time_point t; };
typedef shim_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double2 { double x, y; };

static inline const char *cudaGetErrorString(cudaError_t) { return "host emulation"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
// "device" allocations carry a 128-byte guard zone on either side, checked when they are freed: a kernel (or a
// cudaMemcpy) that writes past its buffer aborts the test run instead of corrupting a neighbour silently
static inline void *shim_dev_alloc(size_t n) {
  const size_t G = 128;
  unsigned char *raw = (unsigned char *)malloc(n + 2 * G + 16);
  if (!raw) return nullptr;
  memcpy(raw, &n, sizeof n);
  memset(raw + 16, 0xA5, G - 16);
  memset(raw + G + n, 0x5A, G);
  return raw + G;
}
static inline void shim_dev_free(void *p) {
  if (!p) return;
  const size_t G = 128;
  unsigned char *raw = (unsigned char *)p - G;
  size_t n;
  memcpy(&n, raw, sizeof n);
  for (size_t i = 16; i < G; i++)
    if (raw[i] != 0xA5) { fprintf(stderr, "host emulation: write BEFORE a device buffer of %zu bytes\n", n); abort(); }
  for (size_t i = 0; i < G; i++)
    if (raw[G + n + i] != 0x5A) { fprintf(stderr, "host emulation: write PAST a device buffer of %zu bytes (+%zu)\n", n, i); abort(); }
  free(raw);
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)shim_dev_alloc(n ? n : 1); return *p ? cudaSuccess : 2; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return cudaSuccess; }
static inline cudaError_t cudaFree(void *p) { shim_dev_free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *s, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) { memcpy((char *)&sym + off, s, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(T &sym, const void *s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = 0) { memcpy((char *)&sym + off, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new shim_event; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new shim_event; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return 1; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return 1; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

// ---------------------------------------------------------------- SIMT on fibers
struct shim_uint3 { unsigned x, y, z; };
extern shim_uint3 threadIdx, blockIdx, blockDim, gridDim;
extern unsigned char smem_raw[];
void shim_launch(dim3 grid, unsigned block, const std::function<void()> &kernel);
void shim_barrier(int warp_level);          // yields until the block (warp) has arrived
extern uint64_t shim_warp_buf[64][32];

#define SHIM_LAUNCH(GRID, BLOCK, ...) shim_launch(dim3(GRID), (unsigned)(BLOCK), [=]() { __VA_ARGS__; })

static inline void __syncthreads() { shim_barrier(0); }
static inline void __syncwarp(unsigned = 0xffffffffu) { shim_barrier(1); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldcv(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <class T> static inline T shim_shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  shim_warp_buf[w][l] = raw;
  shim_barrier(1);
  T r = v;
  if (src >= 0 && src < 32 && (unsigned)((w << 5) + src) < blockDim.x) memcpy(&r, &shim_warp_buf[w][src], sizeof(T));
  shim_barrier(1);
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return shim_shfl(v, src); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return shim_shfl(v, l + d < 32 ? l + d : l); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return shim_shfl(v, (int)((threadIdx.x & 31) ^ m)); }
static inline double atomicAdd(double *p, double v) { const double o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline int atomicExch(int *p, int v) { const int o = *p; *p = v; return o; }

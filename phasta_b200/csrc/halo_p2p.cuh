// halo_p2p.cuh -- the two kernels of the peer-store halo transport (included by comm.cu; tests/host_emul/halo_host.cpp
// compiles them for the host, one process per rank over a shared-memory arena).
#pragma once
// ---- halo exchange by direct peer stores over NVLink (no NCCL call between pack and unpack) ----------------
// Sender: pack the task's values straight into the RECEIVER's arena (slot = message number & 1), fence at system
// scope, and let the block that finishes last raise the receiver's flag to the message number.  Before
// overwriting a slot the sender waits for the receiver's acknowledgement of the message that used it two messages
// ago (the receiver's unpack kernel writes it into the SENDER's arena), so no ordering assumption about the
// sequence of 'in'/'out' exchanges is needed.  All spins are bounded (error flag instead of a hang).
#ifndef PHB_SPIN_MAX
#define PHB_SPIN_MAX (1ll << 27)
#endif
#ifndef PHB_SPIN_PAUSE
#define PHB_SPIN_PAUSE()  // the host emulation yields the CPU here
#endif
__global__ void k_halo_send(int count, const int *__restrict__ nodes, int nshg, int n, const double *__restrict__ g,
                            double *dst, volatile unsigned long long *peer_flag, volatile unsigned long long *my_ack,
                            unsigned long long msg, unsigned int *ticket, int *err) {
  __shared__ int last;
  if (threadIdx.x == 0 && msg > 2) {
    long long spins = 0;
    const bool dead = *reinterpret_cast<volatile int *>(err) != 0;
    while (!dead && *my_ack + 2 < msg) {
      if (++spins > PHB_SPIN_MAX) { atomicExch(err, 100); break; }
      PHB_SPIN_PAUSE();
    }
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count * n) {
    const int k = t / count, i = t % count;
    dst[t] = g[(size_t)nshg * k + nodes[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence_system();
    *peer_flag = msg;
  }
}
__global__ void k_halo_recv(int count, const int *__restrict__ nodes, int nshg, int n, double *__restrict__ g,
                            const double *src, volatile unsigned long long *my_flag,
                            volatile unsigned long long *peer_ack, unsigned long long msg, int add,
                            unsigned int *ticket, int *err) {
  __shared__ int last;
  if (threadIdx.x == 0) {
    long long spins = 0;
    const bool dead = *reinterpret_cast<volatile int *>(err) != 0;
    while (!dead && *my_flag != msg) {
      if (++spins > PHB_SPIN_MAX) { atomicExch(err, 200); break; }
      PHB_SPIN_PAUSE();
    }
    __threadfence_system();
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count * n) {
    const int k = t / count, i = t % count;
    const double v = __ldcv(src + t);  // written by the peer: never from a stale L1 line
    double *p = g + (size_t)nshg * k + nodes[i];
    *p = add ? (*p + v) : v;
  }
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence_system();
    *peer_ack = msg;
  }
}


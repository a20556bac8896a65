// genadj.cu -- genadj / Asadj (phSolver/common/genadj.f:1-82, asadj.f:1-59) on the device.
//
// The reference grows, element by element, the list of distinct neighbours of every node (asadj.f:24-50, an
// O(degree^2) search per insertion), then selection-sorts each list and compacts them into colm / rowp
// (genadj.f:50-77).  The result is fully determined by the mesh: row i holds the ascending distinct node ids that
// share an element with i (i itself included).  Here: every element emits its nshl^2 ordered node pairs as keys
// row * 2^b + col, one radix sort over the 2b significant bits, one unique pass, a per-row count and an exclusive
// scan.  Integer work, bit-exact against the executed genadj.f (tests/test_golden_f77.py, tests/test_gpu_sparse.py).
// CUB (part of the CUDA toolkit) provides sort / unique / scan; this is one-time set-up, not the timed path.
#include "ctx.h"
#include <cub/cub.cuh>

__global__ void k_adj_pairs(int nshl, int numel, size_t numel_pad, const int *__restrict__ ien, int bits,
                            unsigned long long *__restrict__ keys) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)numel * nshl) return;
  const size_t e = t % numel;
  const int a = (int)(t / numel);
  const unsigned long long row = (unsigned long long)ien[(size_t)a * numel_pad + e];
  for (int b = 0; b < nshl; b++)
    keys[(size_t)(a * nshl + b) * numel + e] = (row << bits) | (unsigned long long)ien[(size_t)b * numel_pad + e];
}
__global__ void k_adj_split(size_t n, const unsigned long long *__restrict__ keys, int bits, int *__restrict__ rowp0,
                            int *__restrict__ rowofblk, int *__restrict__ count) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned long long key = keys[k];
  const int row = (int)(key >> bits), col = (int)(key & ((1ull << bits) - 1ull));
  rowp0[k] = col;
  rowofblk[k] = row;
  atomicAdd(count + row, 1);
}

#define CUB_TRY(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "phb200: genadj: %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      rc = 1;                                                                              \
      goto done;                                                                           \
    }                                                                                      \
  } while (0)

// Builds the 0-based CSR structure of the part on the device: *d_colm0 (nshg+1), *d_rowp0 (nnz_tot), *d_rob
// (row of every entry); the caller owns the three arrays (cudaFree).
int phb_genadj_dev(phb200_ctx *ctx, int **d_colm0, int **d_rowp0, int **d_rob, long long *nnz_tot) {
  const int nshg = ctx->c.nshg;
  cudaStream_t s = ctx->stream;
  int rc = 0;
  size_t npairs = (size_t)ctx->numel_tet * 16;
  for (const ElemGroup &g : ctx->gen) npairs += (size_t)g.numel * g.nshl * g.nshl;
  int bits = 1;
  while ((1ll << bits) < (long long)nshg) bits++;
  unsigned long long *d_keys = nullptr, *d_alt = nullptr, *d_uniq = nullptr;
  void *d_tmp = nullptr;
  long long *d_num = nullptr;
  int *d_cnt = nullptr;
  size_t tmp_bytes = 0, need = 0;
  long long n_unique = 0;
  *d_colm0 = *d_rowp0 = *d_rob = nullptr;
  CUB_TRY(cudaMalloc(&d_keys, sizeof(unsigned long long) * npairs));
  CUB_TRY(cudaMalloc(&d_alt, sizeof(unsigned long long) * npairs));
  CUB_TRY(cudaMalloc(&d_num, sizeof(long long)));
  {
    size_t off = 0;
    if (ctx->numel_tet > 0) {
      const size_t nt = (size_t)ctx->numel_tet * 4;
      k_adj_pairs<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(4, ctx->numel_tet, ctx->numel_pad, ctx->d_ien, bits, d_keys);
      off += (size_t)ctx->numel_tet * 16;
      ctx->launches++;
    }
    for (const ElemGroup &g : ctx->gen) {
      const size_t nt = (size_t)g.numel * g.nshl;
      k_adj_pairs<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(g.nshl, g.numel, g.numel_pad, g.d_ien, bits, d_keys + off);
      off += (size_t)g.numel * g.nshl * g.nshl;
      ctx->launches++;
    }
    CUB_TRY(cudaGetLastError());
  }
  {
    cub::DoubleBuffer<unsigned long long> buf(d_keys, d_alt);
    CUB_TRY(cub::DeviceRadixSort::SortKeys(nullptr, need, buf, npairs, 0, 2 * bits, s));
    tmp_bytes = need;
    CUB_TRY(cub::DeviceSelect::Unique(nullptr, need, d_keys, d_alt, d_num, npairs, s));
    if (need > tmp_bytes) tmp_bytes = need;
    CUB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, (int *)nullptr, (int *)nullptr, nshg + 1, s));
    if (need > tmp_bytes) tmp_bytes = need;
    CUB_TRY(cudaMalloc(&d_tmp, tmp_bytes));
    need = tmp_bytes;
    CUB_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, need, buf, npairs, 0, 2 * bits, s));
    unsigned long long *sorted = buf.Current();
    d_uniq = (sorted == d_keys) ? d_alt : d_keys;
    need = tmp_bytes;
    CUB_TRY(cub::DeviceSelect::Unique(d_tmp, need, sorted, d_uniq, d_num, npairs, s));
    ctx->launches += 8;
  }
  CUB_TRY(cudaMemcpyAsync(&n_unique, d_num, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CUB_TRY(cudaStreamSynchronize(s));
  if (n_unique > 2147483647ll) {
    fprintf(stderr, "phb200: genadj: %lld entries exceed the reference's 32-bit nnz_tot\n", n_unique);
    rc = 1;
    goto done;
  }
  CUB_TRY(cudaMalloc(d_rowp0, sizeof(int) * (size_t)(n_unique + 8)));
  CUB_TRY(cudaMalloc(d_rob, sizeof(int) * (size_t)(n_unique + 8)));
  CUB_TRY(cudaMalloc(d_colm0, sizeof(int) * ((size_t)nshg + 1 + 8)));
  CUB_TRY(cudaMalloc(&d_cnt, sizeof(int) * ((size_t)nshg + 1)));
  CUB_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(int) * ((size_t)nshg + 1), s));
  CUB_TRY(cudaMemsetAsync(*d_colm0, 0, sizeof(int) * ((size_t)nshg + 1 + 8), s));
  CUB_TRY(cudaMemsetAsync(*d_rowp0, 0, sizeof(int) * (size_t)(n_unique + 8), s));
  k_adj_split<<<(unsigned)((n_unique + 255) / 256), 256, 0, s>>>((size_t)n_unique, d_uniq, bits, *d_rowp0, *d_rob, d_cnt);
  ctx->launches++;
  CUB_TRY(cudaGetLastError());
  need = tmp_bytes;
  CUB_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_cnt, *d_colm0, nshg + 1, s));
  CUB_TRY(cudaStreamSynchronize(s));
  *nnz_tot = n_unique;
done:
  if (d_keys) cudaFree(d_keys);
  if (d_alt) cudaFree(d_alt);
  if (d_tmp) cudaFree(d_tmp);
  if (d_num) cudaFree(d_num);
  if (d_cnt) cudaFree(d_cnt);
  if (rc) {
    if (*d_colm0) cudaFree(*d_colm0);
    if (*d_rowp0) cudaFree(*d_rowp0);
    if (*d_rob) cudaFree(*d_rob);
    *d_colm0 = *d_rowp0 = *d_rob = nullptr;
  }
  return rc;
}

// ---------------------------------------------------------------------------
// node -> incident (tet, local node) lists, ascending in the element id (the order in which local.f:67-74 adds
// the element contributions of a block sequence): the deterministic assembly option gathers along them
// ---------------------------------------------------------------------------
__global__ void k_inc_keys(int numel, size_t numel_pad, const int *__restrict__ ien, unsigned long long *__restrict__ keys) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)numel * 4) return;
  const size_t e = t % numel;
  const int a = (int)(t / numel);
  keys[t] = ((unsigned long long)ien[(size_t)a * numel_pad + e] << 34) | ((unsigned long long)e << 2) | (unsigned long long)a;
}
__global__ void k_inc_split(size_t n, const unsigned long long *__restrict__ keys, int *__restrict__ inc, int *__restrict__ count) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned long long key = keys[k];
  inc[k] = (int)(key & ((1ull << 34) - 1ull));
  atomicAdd(count + (int)(key >> 34), 1);
}
int phb_build_incidence(phb200_ctx *ctx) {
  if (ctx->d_inc) return 0;
  const int nshg = ctx->c.nshg, numel = ctx->numel_tet;
  cudaStream_t s = ctx->stream;
  int rc = 0;
  if ((size_t)numel * 4 > 2147483647ull || numel >= (1 << 30)) {
    fprintf(stderr, "phb200: deterministic: more than 2^29 tets per part\n");
    return 1;
  }
  const size_t n = (size_t)numel * 4;
  int bits = 1;
  while ((1ll << bits) < (long long)nshg) bits++;
  unsigned long long *d_keys = nullptr, *d_alt = nullptr;
  void *d_tmp = nullptr;
  int *d_cnt = nullptr;
  size_t tmp_bytes = 0, need = 0;
  CUB_TRY(cudaMalloc(&d_keys, sizeof(unsigned long long) * (n + 1)));
  CUB_TRY(cudaMalloc(&d_alt, sizeof(unsigned long long) * (n + 1)));
  CUB_TRY(cudaMalloc(&ctx->d_inc, sizeof(int) * (n + 1)));
  CUB_TRY(cudaMalloc(&ctx->d_inc_ptr, sizeof(int) * ((size_t)nshg + 1)));
  CUB_TRY(cudaMalloc(&d_cnt, sizeof(int) * ((size_t)nshg + 1)));
  CUB_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(int) * ((size_t)nshg + 1), s));
  if (n > 0) k_inc_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(numel, ctx->numel_pad, ctx->d_ien, d_keys);
  {
    cub::DoubleBuffer<unsigned long long> buf(d_keys, d_alt);
    CUB_TRY(cub::DeviceRadixSort::SortKeys(nullptr, need, buf, n, 0, 34 + bits, s));
    tmp_bytes = need;
    CUB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, (int *)nullptr, (int *)nullptr, nshg + 1, s));
    if (need > tmp_bytes) tmp_bytes = need;
    CUB_TRY(cudaMalloc(&d_tmp, tmp_bytes));
    need = tmp_bytes;
    CUB_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, need, buf, n, 0, 34 + bits, s));
    if (n > 0) k_inc_split<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, buf.Current(), ctx->d_inc, d_cnt);
    need = tmp_bytes;
    CUB_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_cnt, ctx->d_inc_ptr, nshg + 1, s));
    ctx->launches += 6;
  }
  CUB_TRY(cudaStreamSynchronize(s));
done:
  if (d_keys) cudaFree(d_keys);
  if (d_alt) cudaFree(d_alt);
  if (d_tmp) cudaFree(d_tmp);
  if (d_cnt) cudaFree(d_cnt);
  if (rc) {
    if (ctx->d_inc) cudaFree(ctx->d_inc);
    if (ctx->d_inc_ptr) cudaFree(ctx->d_inc_ptr);
    ctx->d_inc = ctx->d_inc_ptr = nullptr;
  }
  return rc;
}

// sparse.cu -- the block-CSR flavour of the path (SolGMRs): CSR structure
// (genadj), element -> CSR block map (sparseloc once, not per assembly),
// Spsi3pre and SparseAp.
//
// Reference: phSolver/common/genadj.f:1-82, asadj.f:1-59,
// common/fillsparse.f:66-126,236-271, compressible/spsi3pre.f:41-221,
// compressible/sparseap.f:26-135.
#include "ctx.h"
#include "mbar.cuh"
#include <algorithm>
#include <cstring>

// k_sparseap_tma: blocks / rows per staged chunk, stages in flight, consumer warps per CTA
#define AP_CB 240
#define AP_CR 64
#define AP_STAGES 2
#define AP_WARPS 15
#define AP_CTAS 2

#if defined(PHB_HOST_FULL) || defined(PHB_HOST_EMUL)
// ---- host emulation only (tests/host_emul): CUB does not build there, so the emulated "device" gets its CSR
//      structure from this host routine; the product path is genadj.cu ----
// ---------------------------------------------------------------------------
// genadj: colm(nshg+1) 1-based row pointers, rowp ascending unique neighbour
// ids incl. self.  The reference grows per-node lists with an O(deg^2) search
// and selection-sorts them (asadj.f:24-50, genadj.f:50-62); the result is the
// sorted unique adjacency, built here from a node->element map in O(N deg log deg).
// ---------------------------------------------------------------------------
static int phb_genadj_host(int nshg, int nelblk, const int *lcblk, const int *const *mien, int nnz, int *colm, int *rowp,
                    int *nnz_tot) {
  std::vector<int> cnt((size_t)nshg + 1, 0);
  for (int b = 0; b < nelblk; b++) {
    const int *lc = lcblk + 10 * b;
    int npro = lc[10] - lc[0], nshl = lc[9];
    for (int a = 0; a < nshl; a++)
      for (int e = 0; e < npro; e++) {
        int v = std::abs(mien[b][e + (size_t)npro * a]);
        if (v < 1 || v > nshg) return 1;
        cnt[v]++;
      }
  }
  std::vector<size_t> start((size_t)nshg + 2, 0);
  for (int i = 1; i <= nshg; i++) start[i + 1] = start[i] + cnt[i];
  // node -> (block, element) incidence
  std::vector<int> inc_b(start[nshg + 1]), inc_e(start[nshg + 1]);
  std::vector<size_t> fill(start.begin(), start.end());
  for (int b = 0; b < nelblk; b++) {
    const int *lc = lcblk + 10 * b;
    int npro = lc[10] - lc[0], nshl = lc[9];
    for (int a = 0; a < nshl; a++)
      for (int e = 0; e < npro; e++) {
        int v = std::abs(mien[b][e + (size_t)npro * a]);
        inc_b[fill[v]] = b;
        inc_e[fill[v]] = e;
        fill[v]++;
      }
  }
  colm[0] = 1;
  size_t icnt = 0;
  std::vector<int> tmp;
  for (int i = 1; i <= nshg; i++) {
    tmp.clear();
    for (size_t t = start[i]; t < start[i + 1]; t++) {
      int b = inc_b[t], e = inc_e[t];
      const int *lc = lcblk + 10 * b;
      int npro = lc[10] - lc[0], nshl = lc[9];
      for (int a = 0; a < nshl; a++) tmp.push_back(std::abs(mien[b][e + (size_t)npro * a]));
    }
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    if (icnt + tmp.size() > (size_t)nnz * nshg) {
      fprintf(stderr, "phb200: genadj: increase nnz (needs more than %d per node)\n", nnz);
      return 1;
    }
    for (int v : tmp) rowp[icnt++] = v;
    colm[i] = (int)icnt + 1;
  }
  *nnz_tot = (int)icnt;
  return 0;
}

int phb_genadj_dev(phb200_ctx *ctx, int **d_colm0, int **d_rowp0, int **d_rob, long long *nnz_tot) {
  const int nshg = ctx->c.nshg, cap = 64;
  std::vector<int> colm((size_t)nshg + 1), rowp((size_t)cap * nshg);
  int ntot = 0;
  if (phb_genadj_host(nshg, ctx->c.nelblk, ctx->h_lcblk.data(), ctx->h_mien.data(), cap, colm.data(), rowp.data(), &ntot))
    return 1;
  PHB_CHECK(cudaMalloc(d_colm0, sizeof(int) * ((size_t)nshg + 1 + 8)));
  PHB_CHECK(cudaMalloc(d_rowp0, sizeof(int) * ((size_t)ntot + 8)));
  PHB_CHECK(cudaMalloc(d_rob, sizeof(int) * ((size_t)ntot + 8)));
  memset(*d_colm0, 0, sizeof(int) * ((size_t)nshg + 1 + 8));
  memset(*d_rowp0, 0, sizeof(int) * ((size_t)ntot + 8));
  for (int i = 0; i <= nshg; i++) (*d_colm0)[i] = colm[i] - 1;
  for (int i = 0; i < nshg; i++)
    for (int k = colm[i] - 1; k < colm[i + 1] - 1; k++) {
      (*d_rowp0)[k] = rowp[k] - 1;
      (*d_rob)[k] = i;
    }
  *nnz_tot = ntot;
  return 0;
}
int phb_build_incidence(phb200_ctx *) {
  fprintf(stderr, "phb200: deterministic: not available under host emulation\n");
  return 1;
}
#endif

// sparseloc (fillsparse.f:236-271) for every (element, a, b): CSR block index
__global__ void k_eloc(int nshl, int numel, size_t numel_pad, const int *__restrict__ ien,
                       const int *__restrict__ colm, const int *__restrict__ rowp, int *__restrict__ eloc) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= numel) return;
  int nd[8];
  for (int a = 0; a < nshl; a++) nd[a] = ien[(size_t)a * numel_pad + e];
  for (int a = 0; a < nshl; a++) {
    const int c0 = colm[nd[a]], n = colm[nd[a] + 1] - c0;
    for (int b = 0; b < nshl; b++) {
      const int target = nd[b];
      int lo = 0, hi = n;  // rowp[c0+lo] <= target < rowp[c0+hi]
      while (hi - lo > 1) {
        int mid = (hi + lo) >> 1;
        if (rowp[c0 + mid] > target) hi = mid; else lo = mid;
      }
      eloc[(size_t)(nshl * a + b) * numel_pad + e] = c0 + lo;
    }
  }
}

// 1-based host arrays -> 0-based device arrays, row of every entry, range check (one thread per row)
__global__ void k_csr_from_ref(int nshg, int nnz_tot, int *__restrict__ colm, int *__restrict__ rowp,
                               int *__restrict__ rowofblk, int *__restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nshg) return;
  const int k0 = colm[i] - 1;  // (colm itself turns 0-based in a second pass: other rows still read it here)
  if (i < nshg) {
    const int k1 = colm[i + 1] - 1;
    if (k0 < 0 || k1 < k0 || k1 > nnz_tot) {
      atomicExch(bad, 1);
    } else {
      for (int k = k0; k < k1; k++) {
        const int j = rowp[k] - 1;
        if (j < 0 || j >= nshg) atomicExch(bad, 2);
        rowp[k] = j;
        rowofblk[k] = i;
      }
    }
  }
}
__global__ void k_sub1(int n, int *v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] -= 1;
}

// Takes ownership of a 0-based CSR structure already on the device (colm0 / rowp0 / rob padded by 8 entries) and
// allocates what hangs off it: lhsK, the element -> block map (sparseloc, once) and SparseAp's row chunks.
// h_colm0: host copy of the 0-based row pointers.
int phb_install_sparse(phb200_ctx *ctx, int *d_colm0, int *d_rowp0, int *d_rob, int nnz_tot, const int *h_colm0) {
  const int nshg = ctx->c.nshg;
  auto F = [](void *p) { if (p) cudaFree(p); };
  F(ctx->d_colm); F(ctx->d_rowp); F(ctx->d_rowofblk); F(ctx->d_lhsK); F(ctx->d_eloc); F(ctx->d_apchunk);
  ctx->d_colm = d_colm0;
  ctx->d_rowp = d_rowp0;
  ctx->d_rowofblk = d_rob;
  ctx->d_lhsK = nullptr; ctx->d_eloc = nullptr; ctx->d_apchunk = nullptr;
  ctx->nnz_tot = nnz_tot;
  // (+8: the bulk copies of k_sparseap_tma start and end on multiples of 4 entries)
  PHB_CHECK(cudaMalloc(&ctx->d_lhsK, sizeof(double) * 25 * ((size_t)nnz_tot + 8)));
  PHB_CHECK(cudaMemsetAsync(ctx->d_lhsK, 0, sizeof(double) * 25 * ((size_t)nnz_tot + 8), ctx->stream));
  PHB_CHECK(cudaMalloc(&ctx->d_eloc, sizeof(int) * 16 * ctx->numel_pad));
  {
    // row chunks of SparseAp: consecutive rows, at most AP_CB blocks and AP_CR rows each (a longer row is a chunk
    // of its own and takes the direct-load path)
    std::vector<int> ch;
    int r = 0;
    while (r < nshg) {
      ch.push_back(r);
      int e = r + 1;
      while (e < nshg && e - r < AP_CR && h_colm0[e + 1] - h_colm0[r] <= AP_CB) e++;
      r = e;
    }
    ch.push_back(nshg);
    ctx->n_apchunk = (int)ch.size() - 1;
    PHB_CHECK(cudaMalloc(&ctx->d_apchunk, sizeof(int) * ch.size()));
    PHB_CHECK(cudaMemcpy(ctx->d_apchunk, ch.data(), sizeof(int) * ch.size(), cudaMemcpyHostToDevice));
  }
  PHB_CHECK(cudaMemsetAsync(ctx->d_eloc, 0, sizeof(int) * 16 * ctx->numel_pad, ctx->stream));
  if (ctx->numel_tet > 0) {
    k_eloc<<<(ctx->numel_tet + 127) / 128, 128, 0, ctx->stream>>>(4, ctx->numel_tet, ctx->numel_pad, ctx->d_ien,
                                                                  ctx->d_colm, ctx->d_rowp, ctx->d_eloc);
    ctx->launches++;
    PHB_CHECK(cudaGetLastError());
  }
  for (ElemGroup &g : ctx->gen) {
    if (g.d_eloc) cudaFree(g.d_eloc);
    const size_t n = (size_t)g.nshl * g.nshl * g.numel_pad;
    PHB_CHECK(cudaMalloc(&g.d_eloc, sizeof(int) * n));
    PHB_CHECK(cudaMemsetAsync(g.d_eloc, 0, sizeof(int) * n, ctx->stream));
    k_eloc<<<(g.numel + 127) / 128, 128, 0, ctx->stream>>>(g.nshl, g.numel, g.numel_pad, g.d_ien, ctx->d_colm,
                                                           ctx->d_rowp, g.d_eloc);
    ctx->launches++;
    PHB_CHECK(cudaGetLastError());
  }
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->have_lhs_sparse = false;
  return 0;
}

// the CSR structure itrdrv owns (genadj's colm / rowp, 1-based) -> device
int phb_set_sparse(phb200_ctx *ctx, const int *colm, const int *rowp, int nnz_tot) {
  const int nshg = ctx->c.nshg;
  if (colm[0] != 1 || colm[nshg] - 1 != nnz_tot) {
    fprintf(stderr, "phb200: set_sparse: colm is not a 1-based row pointer array ending at nnz_tot+1\n");
    return 1;
  }
  std::vector<int> c0((size_t)nshg + 1);
  for (int i = 0; i <= nshg; i++) c0[i] = colm[i] - 1;
  int *d_c = nullptr, *d_r = nullptr, *d_rob = nullptr, *d_bad = nullptr;
  PHB_CHECK(cudaMalloc(&d_c, sizeof(int) * ((size_t)nshg + 1 + 8)));
  PHB_CHECK(cudaMalloc(&d_r, sizeof(int) * ((size_t)nnz_tot + 8)));
  PHB_CHECK(cudaMalloc(&d_rob, sizeof(int) * ((size_t)nnz_tot + 8)));
  PHB_CHECK(cudaMalloc(&d_bad, sizeof(int)));
  PHB_CHECK(cudaMemsetAsync(d_c, 0, sizeof(int) * ((size_t)nshg + 1 + 8), ctx->stream));
  PHB_CHECK(cudaMemsetAsync(d_r, 0, sizeof(int) * ((size_t)nnz_tot + 8), ctx->stream));
  PHB_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  PHB_CHECK(cudaMemcpyAsync(d_c, colm, sizeof(int) * ((size_t)nshg + 1), cudaMemcpyHostToDevice, ctx->stream));
  PHB_CHECK(cudaMemcpyAsync(d_r, rowp, sizeof(int) * (size_t)nnz_tot, cudaMemcpyHostToDevice, ctx->stream));
  k_csr_from_ref<<<(nshg + 1 + 127) / 128, 128, 0, ctx->stream>>>(nshg, nnz_tot, d_c, d_r, d_rob, d_bad);
  k_sub1<<<(nshg + 1 + 255) / 256, 256, 0, ctx->stream>>>(nshg + 1, d_c);
  ctx->launches += 2;
  int bad = 0;
  PHB_CHECK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_bad);
  if (bad) {
    cudaFree(d_c); cudaFree(d_r); cudaFree(d_rob);
    fprintf(stderr, bad == 1 ? "phb200: set_sparse: colm is not monotone within 1..nnz_tot+1\n"
                             : "phb200: set_sparse: rowp entry out of range\n");
    return 1;
  }
  return phb_install_sparse(ctx, d_c, d_r, d_rob, nnz_tot, c0.data());
}

// genadj on the device (genadj.cu), installed as the part's CSR structure; colm / rowp (1-based, the reference's
// arrays) are filled when non-null.  nnz is the reference's per-node capacity of rowp (genadj.f:30,66).
int phb_genadj(phb200_ctx *ctx, int nnz, int *colm, int *rowp, int *nnz_tot) {
  const int nshg = ctx->c.nshg;
  int *d_c = nullptr, *d_r = nullptr, *d_rob = nullptr;
  long long ntot = 0;
  PHB_TRY(phb_genadj_dev(ctx, &d_c, &d_r, &d_rob, &ntot));
  if (ntot > (long long)nnz * nshg) {
    cudaFree(d_c); cudaFree(d_r); cudaFree(d_rob);
    fprintf(stderr, "phb200: genadj: increase nnz (needs more than %d per node)\n", nnz);
    return 1;
  }
  std::vector<int> c0((size_t)nshg + 1);
  PHB_CHECK(cudaMemcpy(c0.data(), d_c, sizeof(int) * ((size_t)nshg + 1), cudaMemcpyDeviceToHost));
  if (colm)
    for (int i = 0; i <= nshg; i++) colm[i] = c0[i] + 1;
  if (rowp) {
    PHB_CHECK(cudaMemcpy(rowp, d_r, sizeof(int) * (size_t)ntot, cudaMemcpyDeviceToHost));
    for (long long k = 0; k < ntot; k++) rowp[k] += 1;
  }
  *nnz_tot = (int)ntot;
  return phb_install_sparse(ctx, d_c, d_r, d_rob, (int)ntot, c0.data());
}

// ---------------------------------------------------------------------------
// Spsi3pre (spsi3pre.f:41-221): thread per CSR block, lhsK(25,k) column-major
// block (entry (f,g) at f + 5 g), L from the row node, U from the column node
// ---------------------------------------------------------------------------
// The 128 blocks of a CTA are one contiguous 25.6 KB piece of lhsK: it is moved through shared memory with
// coalesced loads / stores, and each thread works on its block there (stride 25 doubles: conflict-free for 64-bit
// accesses).  Thread-strided global accesses made this kernel 2.4x slower than its 400 B/block of traffic.
__global__ void __launch_bounds__(128) k_spsi3pre(int nnz_tot, int nshg, const int *__restrict__ rowofblk,
                                                   const int *__restrict__ rowp, const double *__restrict__ BD,
                                                   double *lhsK) {
  __shared__ double sh[128 * 25];
  const int kb = blockIdx.x * 128;
  const int nb = (nnz_tot - kb < 128) ? nnz_tot - kb : 128;
  double *gsrc = lhsK + (size_t)25 * kb;
  for (int t = threadIdx.x; t < nb * 25; t += 128) sh[t] = gsrc[t];
  __syncthreads();
  const int k = kb + threadIdx.x;
  if (k < nnz_tot) {
  const int i = rowofblk[k], j = rowp[k];
  double *blk = sh + 25 * threadIdx.x;
  double B[5][5];  // B[f][g]
#pragma unroll
  for (int g = 0; g < 5; g++)
#pragma unroll
    for (int f = 0; f < 5; f++) B[f][g] = blk[f + 5 * g];
#define LI(r, c) __ldg(BD + (size_t)nshg * (((r)-1) + 5 * ((c)-1)) + i)
  {
    const double l21 = LI(2, 1), l31 = LI(3, 1), l32 = LI(3, 2), l41 = LI(4, 1), l42 = LI(4, 2), l43 = LI(4, 3),
                 l51 = LI(5, 1), l52 = LI(5, 2), l53 = LI(5, 3), l54 = LI(5, 4);
#pragma unroll
    for (int g = 0; g < 5; g++) {
      B[1][g] = B[1][g] - l21 * B[0][g];
      B[2][g] = B[2][g] - l31 * B[0][g] - l32 * B[1][g];
      B[3][g] = B[3][g] - l41 * B[0][g] - l42 * B[1][g] - l43 * B[2][g];
      B[4][g] = B[4][g] - l51 * B[0][g] - l52 * B[1][g] - l53 * B[2][g] - l54 * B[3][g];
    }
  }
#undef LI
#define UJ(r, c) __ldg(BD + (size_t)nshg * (((r)-1) + 5 * ((c)-1)) + j)
  {
    const double u11 = UJ(1, 1), u22 = UJ(2, 2), u33 = UJ(3, 3), u44 = UJ(4, 4), u55 = UJ(5, 5);
    const double u12 = UJ(1, 2), u13 = UJ(1, 3), u14 = UJ(1, 4), u15 = UJ(1, 5), u23 = UJ(2, 3), u24 = UJ(2, 4),
                 u25 = UJ(2, 5), u34 = UJ(3, 4), u35 = UJ(3, 5), u45 = UJ(4, 5);
#pragma unroll
    for (int f = 0; f < 5; f++) {
      B[f][0] = u11 * B[f][0];
      B[f][1] = u22 * (B[f][1] - u12 * B[f][0]);
      B[f][2] = u33 * (B[f][2] - u13 * B[f][0] - u23 * B[f][1]);
      B[f][3] = u44 * (B[f][3] - u14 * B[f][0] - u24 * B[f][1] - u34 * B[f][2]);
      B[f][4] = u55 * (B[f][4] - u15 * B[f][0] - u25 * B[f][1] - u35 * B[f][2] - u45 * B[f][3]);
    }
  }
#undef UJ
#pragma unroll
  for (int g = 0; g < 5; g++)
#pragma unroll
    for (int f = 0; f < 5; f++) blk[f + 5 * g] = B[f][g];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nb * 25; t += 128) gsrc[t] = sh[t];
}

int phb_spsi3pre(phb200_ctx *ctx) {
  if (ctx->nnz_tot <= 0) return 0;
  KScope ks(ctx, KC_I3PRE);
  k_spsi3pre<<<(ctx->nnz_tot + 127) / 128, 128, 0, ctx->stream>>>(ctx->nnz_tot, ctx->c.nshg, ctx->d_rowofblk,
                                                                   ctx->d_rowp, ctx->d_BDiag, ctx->d_lhsK);
  PHB_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// SparseAp (sparseap.f:45-101): one warp per row.  Lane l < 25 owns entry
// l = f + 5 g of every 5x5 block of the row (lhsK(25,k) is entry-fastest), so
// a block is one coalesced 200-byte load, there is no index arithmetic in the
// loop, and the column id is one broadcast load per block; eight blocks are in
// flight per warp.  q(i,f) = sum_g of lanes f+5g at the end (3 shuffles).
// HBM-bound: 204 B per block.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sparseap(int nshg, const int *__restrict__ colm,
                                                   const int *__restrict__ rowp, const double *__restrict__ lhsK,
                                                   const double *__restrict__ p, double *__restrict__ q,
                                                   const int *__restrict__ skip) {
  const int lane = threadIdx.x & 31;
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (row >= nshg || (skip && *skip)) return;
  const bool act = lane < 25;
  const int l = act ? lane : 0;
  const int g = l / 5;
  const double *__restrict__ pg = p + (size_t)nshg * g;
  const int k0 = colm[row], k1 = colm[row + 1];
  const double *__restrict__ a = lhsK + (size_t)25 * k0 + l;
  double acc0 = 0.0, acc1 = 0.0;
  // eight blocks (1.6 KB) in flight per warp; the tail is predicated instead of looped so a typical
  // 15-block row takes two round trips to memory
  for (int k = k0; k < k1; k += 8, a += 200) {
    int j[8];
    double av[8], pv[8];
    // (ptxas keeps this at 32 registers by interleaving the FMAs with the loads; forcing all 24 loads ahead
    // of the FMAs costs 52 registers and half the occupancy, and measured slower: 0.57 vs 0.47 ms)
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const bool ok = k + i < k1;
      j[i] = ok ? __ldg(rowp + k + i) : row;
      av[i] = ok ? __ldcs(a + 25 * i) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) pv[i] = __ldg(pg + j[i]);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      acc0 += av[i] * pv[i];
      acc1 += av[i + 1] * pv[i + 1];
    }
  }
  double acc = act ? acc0 + acc1 : 0.0;
  const double t20 = __shfl_down_sync(0xffffffffu, acc, 20);
  acc += __shfl_down_sync(0xffffffffu, acc, 10);
  acc += __shfl_down_sync(0xffffffffu, acc, 5);
  acc += t20;
  if (lane < 5) q[(size_t)nshg * lane + row] = acc;
}

#if !defined(PHB_HOST_EMUL) && !defined(PHB_HOST_FULL)
// ---------------------------------------------------------------------------
// SparseAp with the matrix streamed through shared memory by the bulk-copy engine (cp.async.bulk + mbarrier).
// The blocks of a run of consecutive rows are ONE contiguous piece of lhsK (and of rowp, colm), so a producer
// warp keeps AP_STAGES such pieces (<= 51 KB each) in flight per SM with no registers tied up, and the 16 consumer
// warps only ever wait on L2-resident gathers of p.  Rows are handed out dynamically inside a stage and a warp that
// finds its stage empty moves on to the next one, so there is no tail at chunk boundaries.
// Copies start / end on multiples of 4 CSR entries (16-byte rule of the bulk copy; the arrays are padded).
// ---------------------------------------------------------------------------
// STAGEA = true: blocks, column ids and row pointers are staged; false: only the indices (4 bytes per block) are, and
// the consumers stream the blocks themselves -- their addresses come from shared memory, so the loads of a whole
// batch of blocks and of the matching entries of p leave together, and the L1 sees every matrix byte once, not twice.
template <bool STAGEA>
struct ApStage {
  double a[STAGEA ? 25 * (AP_CB + 16) : 2];   // + alignment slack (3 + 3) and the over-read of the last batch
  int col[AP_CB + 16];
  int rowptr[AP_CR + 8];
};
template <bool STAGEA>
struct ApSmem {
  static constexpr int NST = STAGEA ? AP_STAGES : 6;
  ApStage<STAGEA> st[NST];
  unsigned long long full[NST], empty[NST];
  int next_row[NST];
  int r0[NST], r1[NST], direct[NST];  // the chunk a stage holds (written by the producer)
};

// p <- node-major copy p5[node][5] of the Krylov vector p[5][nshg]: the 25 lanes of a block then gather their five
// entries of p from ONE 40-byte piece (1-2 sectors) instead of five sectors nshg apart -- ncu showed the L1 91 % busy
__global__ void k_node_major(int nshg, const double *__restrict__ p, double *__restrict__ p5) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nshg * 5) return;
  const int i = (int)(t / 5), g = (int)(t % 5);
  p5[t] = p[(size_t)nshg * g + i];
}

template <bool STAGEA, bool PNM = false>
__global__ void __launch_bounds__(32 * (AP_WARPS + 1), AP_CTAS)
    k_sparseap_tma(int nshg, int nchunk, const int *__restrict__ chunk, const int *__restrict__ colm,
                   const int *__restrict__ rowp, const double *__restrict__ lhsK, const double *__restrict__ p,
                   double *__restrict__ q, const int *__restrict__ skip) {
  extern __shared__ __align__(128) unsigned char ap_smem_raw[];
  ApSmem<STAGEA> &S = *reinterpret_cast<ApSmem<STAGEA> *>(ap_smem_raw);
  constexpr int AP_STAGES_ = ApSmem<STAGEA>::NST;
  if (skip && *skip) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < AP_STAGES_; i++) {
      mbar_init(&S.full[i], 1);
      mbar_init(&S.empty[i], AP_WARPS);
      S.next_row[i] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == AP_WARPS) {
    // ------------------------------------------------------------ producer (one lane)
    if (lane == 0) {
      int it = 0;
      for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
        const int stg = it % AP_STAGES_;
        if (it >= AP_STAGES_) mbar_wait(&S.empty[stg], ((it / AP_STAGES_) - 1) & 1);
        const int r0 = chunk[c], r1 = chunk[c + 1];
        const int k0 = colm[r0], k1 = colm[r1];
        S.next_row[stg] = 0;
        S.r0[stg] = r0;
        S.r1[stg] = r1;
        S.direct[stg] = (k1 - k0 > AP_CB);
        if (k1 - k0 > AP_CB) {  // a single very long row: consumers load it directly
          mbar_arrive(&S.full[stg]);
          continue;
        }
        const int ka = k0 & ~3, kb = (k1 + 3) & ~3, ra = r0 & ~3, rb = (r1 + 1 + 3) & ~3;
        const unsigned ba = (unsigned)(kb - ka) * 200u, bc = (unsigned)(kb - ka) * 4u, br = (unsigned)(rb - ra) * 4u;
        mbar_expect_tx(&S.full[stg], (STAGEA ? ba : 0u) + bc + br);
        if (STAGEA) bulk_g2s(S.st[stg].a, lhsK + (size_t)25 * ka, ba, &S.full[stg]);
        bulk_g2s(S.st[stg].col, rowp + ka, bc, &S.full[stg]);
        bulk_g2s(S.st[stg].rowptr, colm + ra, br, &S.full[stg]);
      }
    }
    return;
  }
  // -------------------------------------------------------------- consumers
  const bool act = lane < 25;
  const int l = act ? lane : 0;
  // PNM: p is the node-major copy, entry g of node j at p[5 j + g]
  const double *__restrict__ pg = PNM ? p + (l / 5) : p + (size_t)nshg * (l / 5);
  constexpr int PS = PNM ? 5 : 1;
  int it = 0;
  for (int c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
    const int stg = it % AP_STAGES_;
    const ApStage<STAGEA> &T = S.st[stg];
    mbar_wait(&S.full[stg], (it / AP_STAGES_) & 1);
    const int r0 = S.r0[stg], r1 = S.r1[stg];
    const int ra = r0 & ~3;
    const bool direct = S.direct[stg] != 0;
    const int ka = T.rowptr[r0 - ra] & ~3;
    for (;;) {
      int r = 0;
      if (lane == 0) r = atomicAdd(&S.next_row[stg], 1);
      r = __shfl_sync(0xffffffffu, r, 0) + r0;
      if (r >= r1) break;
      double acc0 = 0.0, acc1 = 0.0;
      if (direct) {
        // a row that does not fit a stage
        const int k0 = colm[r], k1 = colm[r + 1];
        for (int k = k0; k < k1; k++) acc0 += __ldcs(lhsK + (size_t)25 * k + l) * __ldg(pg + (size_t)PS * __ldg(rowp + k));
      } else {
        const int k0 = T.rowptr[r - ra], k1 = T.rowptr[r - ra + 1];
        const int *cj = T.col + (k0 - ka);
        if (STAGEA) {
          const double *a = T.a + 25 * (k0 - ka) + l;
          // sixteen gathers of p in flight per lane: a typical row (15 blocks) is one round trip to L2
          for (int k = k0; k < k1; k += 16, a += 400, cj += 16) {
            double pv[16];
#pragma unroll
            for (int i = 0; i < 16; i++) pv[i] = __ldg(pg + (size_t)PS * ((k + i < k1) ? cj[i] : r));
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              acc0 += ((k + i < k1) ? a[25 * i] : 0.0) * pv[i];
              acc1 += ((k + i + 1 < k1) ? a[25 * (i + 1)] : 0.0) * pv[i + 1];
            }
          }
        } else {
          const double *a = lhsK + (size_t)25 * k0 + l;
          // eight blocks and their entries of p in flight per lane, no load waits for another one
          for (int k = k0; k < k1; k += 8, a += 200, cj += 8) {
            double av[8], pv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const bool ok = k + i < k1;
              av[i] = ok ? __ldcs(a + 25 * i) : 0.0;
              pv[i] = __ldg(pg + (size_t)PS * (ok ? cj[i] : r));
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              acc0 += av[i] * pv[i];
              acc1 += av[i + 1] * pv[i + 1];
            }
          }
        }
      }
      double acc = act ? acc0 + acc1 : 0.0;
      const double t20 = __shfl_down_sync(0xffffffffu, acc, 20);
      acc += __shfl_down_sync(0xffffffffu, acc, 10);
      acc += __shfl_down_sync(0xffffffffu, acc, 5);
      acc += t20;
      if (lane < 5) q[(size_t)nshg * lane + r] = acc;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[stg]);
  }
}
#endif

__global__ void k_iper_copy5(int n, const int *__restrict__ slaves, const int *__restrict__ iper, int nshg,
                             double *u) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 5) return;
  int j = slaves[t % n], k = t / n;
  u[(size_t)nshg * k + j] = u[(size_t)nshg * k + iper[j]];
}

// out <- A p (sparseap.f:26-135); d_p is the caller's scratch copy (halo 'out' + periodic copy fill its slaves)
int phb_sparseap2(phb200_ctx *ctx, double *d_p, double *d_out, const int *d_skip) {
  const int nshg = ctx->c.nshg;
  cudaStream_t s = ctx->stream;
  if (!ctx->have_lhs_sparse) {
    fprintf(stderr, "phb200: sparseap: no sparse LHS has been assembled (elmgmrs with lhs=1 first)\n");
    return 1;
  }
  PHB_TRY(phb_commu(ctx, d_p, 5, 1));
  if (ctx->n_perslave) {
    KScope ks(ctx, KC_NODE);
    int tot = ctx->n_perslave * 5;
    k_iper_copy5<<<(tot + 255) / 256, 256, 0, s>>>(ctx->n_perslave, ctx->d_perslave, ctx->d_iper, nshg, d_p);
    PHB_CHECK(cudaGetLastError());
  }
  {
    KScope ks(ctx, KC_AP);
#if !defined(PHB_HOST_EMUL) && !defined(PHB_HOST_FULL)
    static bool attr_set = false;
    // PHB200_AP_STAGEA=1: also the blocks go through shared memory (A/B runs; measured equal within 2 %)
    static const bool stage_a = getenv("PHB200_AP_STAGEA") && atoi(getenv("PHB200_AP_STAGEA")) == 1;
    if (!attr_set) {
      PHB_CHECK(cudaFuncSetAttribute(k_sparseap_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ApSmem<true>)));
      PHB_CHECK(cudaFuncSetAttribute(k_sparseap_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ApSmem<false>)));
      PHB_CHECK(cudaFuncSetAttribute(k_sparseap_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(ApSmem<false>)));
      attr_set = true;
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    const int grid = std::min(ctx->n_apchunk, nsm * AP_CTAS);
    // PHB200_AP_NODEMAJOR=0: gather p from the [5][nshg] vector itself (A/B runs)
    static const bool node_major = !(getenv("PHB200_AP_NODEMAJOR") && atoi(getenv("PHB200_AP_NODEMAJOR")) == 0);
    if (stage_a) {
      k_sparseap_tma<true><<<grid, 32 * (AP_WARPS + 1), sizeof(ApSmem<true>), s>>>(
          nshg, ctx->n_apchunk, ctx->d_apchunk, ctx->d_colm, ctx->d_rowp, ctx->d_lhsK, d_p, d_out, d_skip);
    } else if (node_major) {
      const size_t n5 = (size_t)nshg * 5;
      k_node_major<<<(unsigned)((n5 + 255) / 256), 256, 0, s>>>(nshg, d_p, ctx->d_p5);
      ctx->launches++;
      k_sparseap_tma<false, true><<<grid, 32 * (AP_WARPS + 1), sizeof(ApSmem<false>), s>>>(
          nshg, ctx->n_apchunk, ctx->d_apchunk, ctx->d_colm, ctx->d_rowp, ctx->d_lhsK, ctx->d_p5, d_out, d_skip);
    } else {
      k_sparseap_tma<false><<<grid, 32 * (AP_WARPS + 1), sizeof(ApSmem<false>), s>>>(
          nshg, ctx->n_apchunk, ctx->d_apchunk, ctx->d_colm, ctx->d_rowp, ctx->d_lhsK, d_p, d_out, d_skip);
    }
#else
    size_t threads = (size_t)nshg * 32;
    k_sparseap<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(nshg, ctx->d_colm, ctx->d_rowp, ctx->d_lhsK, d_p,
                                                                 d_out, d_skip);
#endif
    PHB_CHECK(cudaGetLastError());
  }
  PHB_TRY(phb_commu(ctx, d_out, 5, 0));
  PHB_TRY(phb_zero_slaves(ctx, d_out, 5, 0));
  return 0;
}

// p <- A p in place (uses d_temp as q)
int phb_sparseap(phb200_ctx *ctx, double *d_u) {
  PHB_TRY(phb_sparseap2(ctx, d_u, ctx->d_temp, nullptr));
  PHB_CHECK(cudaMemcpyAsync(d_u, ctx->d_temp, sizeof(double) * 5 * (size_t)ctx->c.nshg, cudaMemcpyDeviceToDevice,
                            ctx->stream));
  return 0;
}

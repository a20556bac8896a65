// timestep.cu -- the Newton / time-step shell around SolGMR* on the device, so that y, ac, yold, acold and
// Dy never leave HBM between the solves of a step.
//
// Reference: phSolver/compressible/itrPC.f:54-119 (itrPredict), :127-150 (itrCorrect), :205-210 (itrUpdate),
// compressible/itrbc.f:60-199 (itrBC), compressible/rstat.f:94-112 (residual norms), and the flow part of the
// step loop compressible/itrdrv.f:393-457,511-524,590-594.
#include "ctx.h"
#include <cmath>

// itrPredict: the y statement of each ipred branch
__global__ void k_itr_predict_y(size_t n, int ipred, double *__restrict__ y, const double *__restrict__ yold,
                                const double *__restrict__ acold, double alfi, double gami, double Dtgl) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (ipred == 1) y[i] = yold[i];
  else if (ipred == 2) y[i] = yold[i] + alfi / Dtgl * acold[i] * (1.0 - gami);
  else if (ipred == 3) y[i] = yold[i] + alfi / Dtgl * acold[i];
  else {
    const double fct1 = alfi / (1.0 - alfi);
    y[i] = yold[i] + fct1 * (yold[i] - y[i]);
  }
}
// ... and its ac statement
__global__ void k_itr_predict_ac(size_t n, int ipred, const double *__restrict__ y, double *__restrict__ ac,
                                 const double *__restrict__ yold, const double *__restrict__ acold, double almi,
                                 double alfi, double gami, double Dtgl) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (ipred == 1) ac[i] = acold[i] * (1.0 - almi / gami);
  else if (ipred == 2) ac[i] = acold[i] * (1.0 - almi);
  else if (ipred == 3) ac[i] = acold[i];
  else {
    const double fct2 = 1.0 - almi / gami, fct3 = almi / gami / alfi * Dtgl;
    ac[i] = acold[i] * fct2 + (y[i] - yold[i]) * fct3;
  }
}

// itrBC node-wise part (itrbc.f:60-177).  The density branch keeps the reference's target column: the
// pressure computed from (rho_BC, T) (getthm.f:75) lands in y(:,1) (itrbc.f:160-162).
__global__ void k_itr_bc(int nshg, const int *__restrict__ iBC, const double *__restrict__ BC, double Rgas,
                         double *__restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nshg) return;
  const int ib = iBC[i];
  if ((ib & 0x3f) == 0) return;
#define YY(j) y[(size_t)nshg * ((j)-1) + i]
#define BCv(j) BC[(size_t)nshg * ((j)-1) + i]
  if (ib & 2) YY(5) = BCv(2);
  switch ((ib >> 3) & 7) {
    case 1: YY(1) = BCv(3) - BCv(4) * YY(2) - BCv(5) * YY(3); break;
    case 2: YY(2) = BCv(3) - BCv(4) * YY(1) - BCv(5) * YY(3); break;
    case 3:
      YY(1) = BCv(3) - BCv(4) * YY(3);
      YY(2) = BCv(5) - BCv(6) * YY(3);
      break;
    case 4: YY(3) = BCv(3) - BCv(4) * YY(1) - BCv(5) * YY(2); break;
    case 5:
      YY(1) = BCv(3) - BCv(4) * YY(2);
      YY(3) = BCv(5) - BCv(6) * YY(2);
      break;
    case 6:
      YY(2) = BCv(3) - BCv(4) * YY(1);
      YY(3) = BCv(5) - BCv(6) * YY(1);
      break;
    case 7:
      YY(1) = BCv(3);
      YY(2) = BCv(4);
      YY(3) = BCv(5);
      break;
    default: break;
  }
  if (ib & 1) YY(1) = Rgas * BCv(1) * YY(5);
  if (ib & 4) YY(4) = BCv(1);
#undef YY
#undef BCv
}

// y(:,i) = y(iper(:),i), ac likewise (itrbc.f:181-184); only periodic slaves differ from their master
__global__ void k_itr_per(int n, const int *__restrict__ slaves, const int *__restrict__ iper, int nshg,
                          double *__restrict__ y, double *__restrict__ ac) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 5) return;
  const int j = slaves[t % n], k = t / n;
  const size_t o = (size_t)nshg * k;
  y[o + j] = y[o + iper[j]];
  if (ac) ac[o + j] = ac[o + iper[j]];
}

// itrCorrect (itrPC.f:140-147): y -= Dy (with the {p,u,T} -> {u,p,T} shuffle), ac from the updated y
__global__ void k_itr_correct(int nshg, double *__restrict__ y, double *__restrict__ ac,
                              const double *__restrict__ yold, const double *__restrict__ acold,
                              const double *__restrict__ Dy, double fct1, double fct2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nshg) return;
  const int src[5] = {1, 2, 3, 0, 4};
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const size_t o = (size_t)nshg * k + i;
    const double yn = y[o] - Dy[(size_t)nshg * src[k] + i];
    y[o] = yn;
    ac[o] = acold[o] * fct1 + (yn - yold[o]) * fct2;
  }
}

// itrUpdate (itrPC.f:205-210)
__global__ void k_itr_update(size_t n, double *__restrict__ yold, double *__restrict__ acold,
                             const double *__restrict__ y, const double *__restrict__ ac, double fct2, double fct3) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  acold[i] = acold[i] + (ac[i] - acold[i]) * fct2;
  yold[i] = yold[i] + (y[i] - yold[i]) * fct3;
}

// sum of squares of two vectors in one pass (rstat.f:94-101)
__global__ void __launch_bounds__(256) k_sumsq2(size_t n, const double *__restrict__ a, const double *__restrict__ b,
                                                double *out) {
  double sa = 0.0, sb = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    sa += a[i] * a[i];
    sb += b[i] * b[i];
  }
  __shared__ double sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_down_sync(0xffffffffu, sa, o);
    sb += __shfl_down_sync(0xffffffffu, sb, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = sa; sh[1][w] = sb; }
  __syncthreads();
  if (threadIdx.x < 8) {
    sa = sh[0][threadIdx.x];
    sb = sh[1][threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      sa += __shfl_down_sync(0xffu, sa, o);
      sb += __shfl_down_sync(0xffu, sb, o);
    }
    if (threadIdx.x == 0) {
      atomicAdd(out, sa);
      atomicAdd(out + 1, sb);
    }
  }
}

static inline unsigned nblk(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

int phb_itrpredict(phb200_ctx *ctx, const phb200_step *st, int ipred) {
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  if (ipred < 1 || ipred > 4) {
    fprintf(stderr, "phb200: itrpredict: ipred=%d (1..4)\n", ipred);
    return 1;
  }
  {
    KScope ks(ctx, KC_NODE);
    k_itr_predict_y<<<nblk(n5, 256), 256, 0, ctx->stream>>>(n5, ipred, ctx->d_y, ctx->d_yold, ctx->d_acold, st->alfi,
                                                            st->gami, st->Dtgl);
    PHB_CHECK(cudaGetLastError());
  }
  if (ipred != 1) PHB_TRY(phb_itrbc(ctx, 1));  // itrPC.f:80,93,109
  {
    KScope ks(ctx, KC_NODE);
    k_itr_predict_ac<<<nblk(n5, 256), 256, 0, ctx->stream>>>(n5, ipred, ctx->d_y, ctx->d_ac, ctx->d_yold,
                                                             ctx->d_acold, st->almi, st->alfi, st->gami, st->Dtgl);
    PHB_CHECK(cudaGetLastError());
  }
  return 0;
}

// itrBC on any (y, ac) pair in the global {u,v,w,p,T} layout; Au1MFG applies it to the perturbed state with
// ires=2, where ac is not touched (itrbc.f:177-191)
int phb_itrbc_vec(phb200_ctx *ctx, double *d_y, double *d_ac, int ires) {
  const int nshg = ctx->c.nshg;
  {
    KScope ks(ctx, KC_NODE);
    k_itr_bc<<<nblk(nshg, 256), 256, 0, ctx->stream>>>(nshg, ctx->d_iBC, ctx->d_BC, ctx->c.Rgas, d_y);
    if (ctx->n_perslave) {
      k_itr_per<<<nblk((size_t)ctx->n_perslave * 5, 256), 256, 0, ctx->stream>>>(
          ctx->n_perslave, ctx->d_perslave, ctx->d_iper, nshg, d_y, ires != 2 ? d_ac : nullptr);
      ctx->launches++;
    }
    PHB_CHECK(cudaGetLastError());
  }
  if (ctx->c.numpe > 1) {
    PHB_TRY(phb_commu(ctx, d_y, 5, 1));
    if (ires != 2) PHB_TRY(phb_commu(ctx, d_ac, 5, 1));
  }
  return 0;
}

int phb_itrbc(phb200_ctx *ctx, int ires) { return phb_itrbc_vec(ctx, ctx->d_y, ctx->d_ac, ires); }

int phb_itrcorrect(phb200_ctx *ctx, const phb200_step *st) {
  KScope ks(ctx, KC_NODE);
  const double fct1 = 1.0 - st->almi / st->gami, fct2 = st->almi * st->Dtgl / st->gami / st->alfi;
  k_itr_correct<<<nblk(ctx->c.nshg, 256), 256, 0, ctx->stream>>>(ctx->c.nshg, ctx->d_y, ctx->d_ac, ctx->d_yold,
                                                                 ctx->d_acold, ctx->d_Dy, fct1, fct2);
  PHB_CHECK(cudaGetLastError());
  return 0;
}

int phb_itrupdate(phb200_ctx *ctx, const phb200_step *st) {
  KScope ks(ctx, KC_NODE);
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  k_itr_update<<<nblk(n5, 256), 256, 0, ctx->stream>>>(n5, ctx->d_yold, ctx->d_acold, ctx->d_y, ctx->d_ac,
                                                       1.0 / st->almi, 1.0 / st->alfi);
  PHB_CHECK(cudaGetLastError());
  return 0;
}

// rstat (rstat.f:94-112): totres(1:2) = sqrt(allreduce(sum res^2, sum b^2)) / nshgt
int phb_rstat(phb200_ctx *ctx, long long nshgt, double *totres) {
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_CHECK(cudaMemsetAsync(ctx->d_dots, 0, sizeof(double) * 2, ctx->stream));
  {
    KScope ks(ctx, KC_BLAS);
    unsigned g = nblk(n5, 256 * 8);
    if (g > 1184) g = 1184;
    if (g < 1) g = 1;
    k_sumsq2<<<g, 256, 0, ctx->stream>>>(n5, ctx->d_res, ctx->d_rmes, ctx->d_dots);
    PHB_CHECK(cudaGetLastError());
  }
  PHB_TRY(phb_allreduce_sum(ctx, ctx->d_dots, 2));
  PHB_CHECK(cudaMemcpyAsync(ctx->h_dots, ctx->d_dots, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  totres[0] = sqrt(ctx->h_dots[0]) / (double)nshgt;
  totres[1] = sqrt(ctx->h_dots[1]) / (double)nshgt;
  return 0;
}

// One time step of itrdrv.f's flow sequence on the resident state: predictor, nitr x (SolGMRe|s, rstat,
// itrCorrect, itrBC), itrUpdate.  stats[6*it + ..] = totres(1), totres(2), iKs, lGMRES, lhs, 0.
int phb_timestep(phb200_ctx *ctx, const phb200_step *st0, int ipred, int nitr, int sparse, int LHSupd,
                 long long nshgt, int *ntotGM, double *stats) {
  if (nitr < 1 || LHSupd < 1) {
    fprintf(stderr, "phb200: timestep: nitr and LHSupd must be >= 1\n");
    return 1;
  }
  if (sparse != 0 && sparse != 1) {
    // 2 would pair a CSR assembly with the matrix-free solve, which needs phb_elmmfg's rmes: not a step flavour
    fprintf(stderr, "phb200: timestep: sparse must be 0 (EBE) or 1 (CSR), got %d\n", sparse);
    return 1;
  }
  phb200_step st = *st0;
  st.nitr = nitr;
  PHB_TRY(phb_itrpredict(ctx, &st, ipred));
  PHB_TRY(phb_itrbc(ctx, 1));  // itrdrv.f:394
  for (int it = 1; it <= nitr; it++) {
    ctx->ifuncs++;
    st.iter = it;
    st.lhs = 1 - (((ctx->ifuncs - 1) % LHSupd) > 0 ? 1 : 0);  // itrdrv.f:456,511
    st.iprec = st.lhs;
    PHB_CHECK(cudaMemsetAsync(ctx->d_aerfrc, 0, sizeof(double) * 4, ctx->stream));  // itrdrv.f:437-442
    int iKs = 0, lG = 0;
    PHB_TRY(phb_elmgmre(ctx, &st, sparse));
    PHB_TRY(phb_solve(ctx, &st, sparse, &iKs, &lG, ntotGM));
    double tot[2];
    PHB_TRY(phb_rstat(ctx, nshgt, tot));
    if (stats) {
      double *s = stats + 6 * (it - 1);
      s[0] = tot[0]; s[1] = tot[1]; s[2] = iKs; s[3] = lG; s[4] = st.lhs; s[5] = 0.0;
    }
    PHB_TRY(phb_itrcorrect(ctx, &st));
    PHB_TRY(phb_itrbc(ctx, 1));
  }
  PHB_TRY(phb_itrupdate(ctx, &st));
  PHB_TRY(phb_itrbc_vec(ctx, ctx->d_yold, ctx->d_acold, 1));  // itrdrv.f:652
  return 0;
}

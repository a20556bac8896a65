// mfg.cu -- the matrix-free flavour of the Newton linear solve on the device:
// ElmMFG (compressible/elmmfg.f:1-256), ItrRes (itrres.f:1-171), Au1MFG / Au2MFG (au1mfg.f:1-98,
// au2mfg.f:1-120), itrFDI (itrfdi.f:1-139), yshuffle (shuffle.f:1-27).  The Krylov loop is phb_solve
// (solver.cu) with flavour 2; the element kernels (k_asires, the e3bdg mode of k_asigmr_*) are in assembly.cu.
//
// Ap = one residual-class element sweep (no EGmass: ~150 B and ~6 kflop per tet per Ap instead of 3 200 B),
// wrapped in node-wise kernels (perturb + i3LU backward + yshuffle, itrBC, i3LU forward + difference).
#include "ctx.h"
#include <cmath>

static inline unsigned nblk(size_t n, int b) { return (unsigned)((n + b - 1) / b); }

// v <- U^-1 (ypre + eps * dir) reordered {p,u,T} -> {u,p,T}: the head of Au1MFG (au1mfg.f:58-68),
// i3LU 'backward' (i3lu.f:118-141) and yshuffle 'old2new' (shuffle.f:8-13) in one pass.  dir may be null.
__global__ void k_mfg_perturb(int nshg, const double *__restrict__ ypre, double eps, const double *__restrict__ dir,
                              const double *__restrict__ D, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nshg) return;
  double r[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const double y0 = ypre[(size_t)nshg * k + i];
    r[k] = dir ? (y0 + eps * dir[(size_t)nshg * k + i]) : y0;
  }
#define DG(a, b) D[(size_t)nshg * (((a)-1) + 5 * ((b)-1)) + i]
  r[4] = DG(5, 5) * r[4];
  r[3] = DG(4, 4) * (r[3] - r[4] * DG(4, 5));
  r[2] = DG(3, 3) * (r[2] - r[4] * DG(3, 5) - r[3] * DG(3, 4));
  r[1] = DG(2, 2) * (r[1] - r[4] * DG(2, 5) - r[3] * DG(2, 4) - r[2] * DG(2, 3));
  r[0] = DG(1, 1) * (r[0] - r[4] * DG(1, 5) - r[3] * DG(1, 4) - r[2] * DG(1, 3) - r[1] * DG(1, 2));
#undef DG
  out[i] = r[1];
  out[(size_t)nshg + i] = r[2];
  out[(size_t)nshg * 2 + i] = r[3];
  out[(size_t)nshg * 3 + i] = r[0];
  out[(size_t)nshg * 4 + i] = r[4];
}

// ypre = U . new2old(y) (solmfg.f:126-129: yshuffle 'new2old', i3LU 'product' i3lu.f:152-165)
__global__ void k_mfg_ypre(int nshg, const double *__restrict__ y, const double *__restrict__ D,
                           double *__restrict__ ypre) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nshg) return;
  double r[5] = {y[(size_t)nshg * 3 + i], y[i], y[(size_t)nshg + i], y[(size_t)nshg * 2 + i], y[(size_t)nshg * 4 + i]};
#define DG(a, b) D[(size_t)nshg * (((a)-1) + 5 * ((b)-1)) + i]
  r[0] = r[0] / DG(1, 1) + r[1] * DG(1, 2) + r[2] * DG(1, 3) + r[3] * DG(1, 4) + r[4] * DG(1, 5);
  r[1] = r[1] / DG(2, 2) + r[2] * DG(2, 3) + r[3] * DG(2, 4) + r[4] * DG(2, 5);
  r[2] = r[2] / DG(3, 3) + r[3] * DG(3, 4) + r[4] * DG(3, 5);
  r[3] = r[3] / DG(4, 4) + r[4] * DG(4, 5);
  r[4] = r[4] / DG(5, 5);
#undef DG
#pragma unroll
  for (int k = 0; k < 5; k++) ypre[(size_t)nshg * k + i] = r[k];
}

// out = (a - b) / e                       (au1mfg.f:84)
__global__ void k_mfg_diff(size_t n, double *__restrict__ out, const double *__restrict__ a,
                           const double *__restrict__ b, double e) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (a[i] - b[i]) / e;
}
// out = res - (a - b) / (2 eps)           (au2mfg.f:105)
__global__ void k_mfg_diff2(size_t n, double *__restrict__ out, const double *__restrict__ res,
                            const double *__restrict__ a, const double *__restrict__ b, double eps) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = res[i] - (a[i] - b[i]) / (2.0 * eps);
}
// v = ((v - 2 rmes) / epsM)^2  or  v = v^2   (itrfdi.f:84,125)
__global__ void k_mfg_sq(size_t n, double *__restrict__ v, const double *__restrict__ rmes, double epsM) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double t = rmes ? (v[i] - 2.0 * rmes[i]) / epsM : v[i];
    v[i] = t * t;
  }
}

static int mfg_alloc(phb200_ctx *ctx) {
  if (!ctx->d_mfg) PHB_CHECK(cudaMalloc(&ctx->d_mfg, sizeof(double) * 4 * 5 * (size_t)ctx->c.nshg));
  return 0;
}
static inline double *mfg_ypre(phb200_ctx *ctx) { return ctx->d_mfg; }
static inline double *mfg_work(phb200_ctx *ctx, int k) { return ctx->d_mfg + (size_t)(1 + k) * 5 * ctx->c.nshg; }

// ItrRes (itrres.f:58-165): d_rmes += modified residual of d_yp, halo sum, bc3Res.  Jactyp = 0 (itrPC.f:29):
// no boundary-element part.  d_rmes is not zeroed here (itrFDI accumulates two calls).
int phb_itrres(phb200_ctx *ctx, const double *d_yp, double *d_rmes, int iabres, int ires) {
  PHB_TRY(phb_asires(ctx, d_yp, d_rmes, iabres, ires));
  PHB_TRY(phb_commu(ctx, d_rmes, 5, 0));
  PHB_TRY(phb_bc3res_vec(ctx, d_rmes));
  return 0;
}

// ElmMFG (elmmfg.f:60-250): res (ires=3 == the ElmGMRe residual), e3bdg block diagonal when iprec/=0, and the
// modified residual of the base state
int phb_elmmfg(phb200_ctx *ctx, const phb200_step *st) {
  phb200_step s2 = *st;
  s2.lhs = 0;  // itrdrv.f:496
  PHB_TRY(phb_elmgmre(ctx, &s2, 0));
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_CHECK(cudaMemsetAsync(ctx->d_rmes, 0, sizeof(double) * n5, ctx->stream));
  // ElmMFG runs e3 with ires=3: with discontinuity capturing its modified residual is not the one ItrRes (ires=2)
  // computes for the same state (k_asires DCM 3 vs 2)
  PHB_TRY(phb_itrres(ctx, ctx->d_y, ctx->d_rmes, 0, 3));
  return 0;
}

// ypre = U new2old(y) (solmfg.f:126-129)
int phb_mfg_begin(phb200_ctx *ctx) {
  PHB_TRY(mfg_alloc(ctx));
  KScope ks(ctx, KC_NODE);
  k_mfg_ypre<<<nblk(ctx->c.nshg, 128), 128, 0, ctx->stream>>>(ctx->c.nshg, ctx->d_y, ctx->d_BDiag, mfg_ypre(ctx));
  PHB_CHECK(cudaGetLastError());
  return 0;
}

// out <- [L^-1] Rm( itrBC( U^-1 (ypre + eps dir) ) ); v is scratch for the perturbed state
static int perturbed_res(phb200_ctx *ctx, double *v, double eps, const double *dir, double *out, bool zero_out,
                         bool with_itrbc, int iabres, bool forward) {
  const int nshg = ctx->c.nshg;
  const size_t n5 = (size_t)5 * nshg;
  {
    KScope ks(ctx, KC_NODE);
    k_mfg_perturb<<<nblk(nshg, 128), 128, 0, ctx->stream>>>(nshg, mfg_ypre(ctx), eps, dir, ctx->d_BDiag, v);
    PHB_CHECK(cudaGetLastError());
  }
  if (with_itrbc) PHB_TRY(phb_itrbc_vec(ctx, v, nullptr, 2));
  if (zero_out) PHB_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * n5, ctx->stream));
  PHB_TRY(phb_itrres(ctx, v, out, iabres));
  if (forward) PHB_TRY(phb_i3lu(ctx, ctx->d_BDiag, out, 1));
  return 0;
}

// Au1MFG (au1mfg.f:52-90): u <- ( L^-1 Rm(U^-1(ypre + e u)) - rmes ) / e, in place
int phb_au1mfg(phb200_ctx *ctx, double *d_u) {
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  const double e = ctx->eGMRES;
  if (!(e > 0.0)) {
    fprintf(stderr, "phb200: au1mfg: eGMRES = %g (itrFDI has not run: first call needs iter=1, mod(istep,20)=0)\n", e);
    return 1;
  }
  double *v = mfg_work(ctx, 0), *w = mfg_work(ctx, 1);
  PHB_TRY(perturbed_res(ctx, v, e, d_u, w, true, true, 0, true));
  KScope ks(ctx, KC_BLAS);
  k_mfg_diff<<<nblk(n5, 256) > 1184 ? 1184 : nblk(n5, 256), 256, 0, ctx->stream>>>(n5, d_u, w, ctx->d_rmes, e);
  PHB_CHECK(cudaGetLastError());
  return 0;
}

// Au2MFG (au2mfg.f:52-112): out <- res - (Rm(+eps Dy) - Rm(-eps Dy)) / (2 eps), eps = epsM^(2/3) / |Dy|
int phb_au2mfg(phb200_ctx *ctx, double *d_out) {
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *v = mfg_work(ctx, 0), *w1 = mfg_work(ctx, 1), *w2 = mfg_work(ctx, 2);
  PHB_CHECK(cudaMemcpyAsync(v, ctx->d_Dy, sizeof(double) * n5, cudaMemcpyDeviceToDevice, ctx->stream));
  {
    KScope ks(ctx, KC_BLAS);
    k_mfg_sq<<<1184, 256, 0, ctx->stream>>>(n5, v, nullptr, 0.0);
    PHB_CHECK(cudaGetLastError());
  }
  double summed = 0.0;
  PHB_TRY(phb_sumgat_dev(ctx, v, n5, &summed));
  const double eps = pow(ctx->c.epsM, 0.6666666666666666666666666666667) / sqrt(summed);
  PHB_TRY(perturbed_res(ctx, v, eps, ctx->d_Dy, w1, true, true, 0, true));
  PHB_TRY(perturbed_res(ctx, v, -eps, ctx->d_Dy, w2, true, true, 0, true));
  KScope ks(ctx, KC_BLAS);
  k_mfg_diff2<<<1184, 256, 0, ctx->stream>>>(n5, d_out, ctx->d_res, w1, w2, eps);
  PHB_CHECK(cudaGetLastError());
  return 0;
}

// itrFDI (itrfdi.f:52-132) with uBrg = res (solmfg.f:150-157): eGMRES = 2 sqrt(epsA / SDnrm)
int phb_itrfdi(phb200_ctx *ctx) {
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  const double epsM = ctx->c.epsM;
  double *v = mfg_work(ctx, 0), *w = mfg_work(ctx, 1);
  PHB_TRY(perturbed_res(ctx, v, 0.0, nullptr, w, true, false, 1, true));
  {
    KScope ks(ctx, KC_BLAS);
    k_mfg_sq<<<1184, 256, 0, ctx->stream>>>(n5, w, nullptr, 0.0);
    PHB_CHECK(cudaGetLastError());
  }
  double summed = 0.0;
  PHB_TRY(phb_sumgat_dev(ctx, w, n5, &summed));
  const double epsA = (epsM * epsM) * sqrt(summed);
  const double epsSD = sqrt(epsM);
  PHB_TRY(perturbed_res(ctx, v, epsSD, ctx->d_res, w, true, false, 0, false));
  PHB_TRY(perturbed_res(ctx, v, -epsSD, ctx->d_res, w, false, false, 0, true));
  {
    KScope ks(ctx, KC_BLAS);
    k_mfg_sq<<<1184, 256, 0, ctx->stream>>>(n5, w, ctx->d_rmes, epsM);
    PHB_CHECK(cudaGetLastError());
  }
  PHB_TRY(phb_sumgat_dev(ctx, w, n5, &summed));
  ctx->eGMRES = 2.0 * sqrt(epsA / sqrt(summed));
  return 0;
}

// incomp.cu -- the incompressible flavour of the path (BASELINE.json configs[3]): ElmGMR into block-CSR and the
// lesSparse matrix-vector products.  Reference (paths relative to /root/reference/phSolver):
//   incompressible/elmgmr.f:1-330 (ElmGMR), asiq.f + e3q.f + e3qvar.f (diffusive-flux projection), asigmr.f + e3.f,
//   e3ivar.f + e3res.f:300-400 (e3resStrongPDE), e3stab.f (itau=0, e3gijd), e3res.f:1-200 (e3Res), e3lhs.f:1-230
//   (e3LHS), bc3lhs.f, common/fillsparse.f:1-65 (fillsparseI), bc3res.f + bc3per.f, lesSparse.f:204-492.
//
// Kernels (sm_100a, FP64):
//   k_inc_asiq<NSHL,NQ>      thread = element: q = 2 mu sym(grad u) projected on the nodes (9 + 1 atomics per node)
//   k_inc_asigmr<NSHL,NQ,LHS> thread = element.  Pass 1 walks the quadrature points once: metric, interpolation,
//                            strong residual, Shakib tau, weak residual rl (scattered with red.f64); it keeps 16
//                            scalars per point.  Pass 2 builds the element tangent one block row at a time: the 13
//                            entries of block (a,b) (3x3 K + G1..3 + C) are summed over the points in registers
//                            (same order as the reference's point loop), bc3LHS acts on the block in registers
//                            (row operations by node a's code, column operations by node b's, lower local node
//                            first, as the reference's node loop does), and the block is added to lhsK/lhsP at the
//                            CSR slot found once by sparseloc (phb_set_sparse).  HBM-bound: 13 doubles out per
//                            (a,b) block, ~150 B in per node; the flops (about 6 kflop per tet) hide under it.
//   k_les_apfull / apkg / apngt / apg   warp = CSR row, lane = entry: K p, G p, C p gathered; the reference's
//                            transposed scatter  q(j,1:3) -= pLhs(1:3,k) p(i,4)  becomes a gather through `tpos`
//                            (the slot of the transposed entry; the adjacency pattern of genadj is symmetric), so
//                            no atomics and a deterministic sum.
#include "ctx.h"
#include <algorithm>
#include <cstring>
#include <vector>

struct IncTab {
  int nq, nshl, lcsyst, pad;
  double N[8][8];      // N[q][a]      shp(lcsyst,a,q)
  double dN[8][8][3];  // dN[q][a][i]  shgl(lcsyst,i,a,q)
  double Qwt[8];
};
__constant__ IncTab c_it[3];  // 0 tets, 1 hexes, 2 wedges

struct IncPhys {
  double rho, rmu, bf[3];
  double tmps;    // 1 - flmpr                      (e3res.f:36)
  double lhsFct;  // alfi * gami * Delt(itseq)      (e3lhs.f:27)
  double lhmFct;  // almi * (1 - flmpl)             (e3lhs.f:28)
  double dts;     // Dtgl * dtsfct                  (e3stab.f:40)
  double ff;      // taucfct / dtsfct               (e3stab.f:63)
  int iconvflow, idiff, matflg5, lhs;
};
__constant__ IncPhys c_ip;
#include "bnd_pack.h"
struct IncBndPhys {
  double rho, rmu;                  // getDiff (incompressible/getdiff.f:24-27), iLSet = 0, DNS
  int iviscflux, iconvflow, itwmod;
};
__constant__ IncBndPhys c_ibp;
__constant__ BndTables c_ibnd[4];   // face tables of lcsyst 1..4 (index lcsyst-1)
#include "inc_boundary.cuh"


static int inc_tab_index(int lcsyst) { return lcsyst == 1 ? 0 : (lcsyst == 2 ? 1 : 2); }

static int upload_inc_tables(phb200_ctx *ctx) {
  if (ctx->have_inc_tabs) return 0;
  const phb200_common &c = ctx->c;
  const double *shp = ctx->h_shp.data(), *shgl = ctx->h_shgl.data();
  auto fill = [&](int lcsyst, int nshl) -> int {
    IncTab t;
    memset(&t, 0, sizeof t);
    const int top = lcsyst - 1;
    t.nq = c.nint[top];
    t.nshl = nshl;
    t.lcsyst = lcsyst;
    if (t.nq < 1 || t.nq > 8) {
      fprintf(stderr, "phb200: inc_elmgmr: quadrature rule with %d points not supported\n", t.nq);
      return 1;
    }
    for (int q = 0; q < t.nq; q++) {
      t.Qwt[q] = c.Qwt[top + PHB200_MAXTOP * q];
      for (int a = 0; a < nshl; a++) {
        t.N[q][a] = shp[top + PHB200_MAXTOP * (a + PHB200_MAXSH * q)];
        for (int i = 0; i < 3; i++) t.dN[q][a][i] = shgl[top + PHB200_MAXTOP * (i + 3 * (a + PHB200_MAXSH * q))];
      }
    }
    PHB_CHECK(cudaMemcpyToSymbol(c_it, &t, sizeof t, sizeof(IncTab) * inc_tab_index(lcsyst)));
    return 0;
  };
  if (ctx->numel_tet > 0) PHB_TRY(fill(1, 4));
  for (const ElemGroup &g : ctx->gen) PHB_TRY(fill(g.lcsyst, g.nshl));
  ctx->have_inc_tabs = true;
  return 0;
}

// ---------------------------------------------------------------------------
// element geometry at one quadrature point (common/e3metric.f:22-77; the metric block of e3qvar.f:17-72 is the
// same arithmetic)
template <int NSHL>
__device__ __forceinline__ void inc_metric(const double xl[NSHL][3], const double (*dN)[3], double Qw,
                                           double shg[NSHL][3], double x[3][3], double &W) {
  double d[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
#pragma unroll
      for (int n = 0; n < NSHL; n++) s += xl[n][i] * dN[n][j];
      d[i][j] = s;
    }
  x[0][0] = d[1][1] * d[2][2] - d[2][1] * d[1][2];
  x[0][1] = d[2][1] * d[0][2] - d[0][1] * d[2][2];
  x[0][2] = d[0][1] * d[1][2] - d[0][2] * d[1][1];
  double tmp = 1.0 / (x[0][0] * d[0][0] + x[0][1] * d[1][0] + x[0][2] * d[2][0]);
  x[0][0] *= tmp; x[0][1] *= tmp; x[0][2] *= tmp;
  x[1][0] = (d[1][2] * d[2][0] - d[1][0] * d[2][2]) * tmp;
  x[1][1] = (d[0][0] * d[2][2] - d[2][0] * d[0][2]) * tmp;
  x[1][2] = (d[1][0] * d[0][2] - d[0][0] * d[1][2]) * tmp;
  x[2][0] = (d[1][0] * d[2][1] - d[1][1] * d[2][0]) * tmp;
  x[2][1] = (d[2][0] * d[0][1] - d[0][0] * d[2][1]) * tmp;
  x[2][2] = (d[0][0] * d[1][1] - d[0][1] * d[1][0]) * tmp;
  W = Qw / tmp;
#pragma unroll
  for (int n = 0; n < NSHL; n++)
#pragma unroll
    for (int i = 0; i < 3; i++) shg[n][i] = dN[n][0] * x[0][i] + dN[n][1] * x[1][i] + dN[n][2] * x[2][i];
}

// e3gijd (incompressible/e3stab.f:330-420): g = {11,22,33,12,23,13}
template <bool TET>
__device__ __forceinline__ void inc_gijd(const double d[3][3], double g[6]) {
  if (!TET) {
    g[0] = d[0][0] * d[0][0] + d[1][0] * d[1][0] + d[2][0] * d[2][0];
    g[3] = d[0][0] * d[0][1] + d[1][0] * d[1][1] + d[2][0] * d[2][1];
    g[1] = d[0][1] * d[0][1] + d[1][1] * d[1][1] + d[2][1] * d[2][1];
    g[4] = d[0][1] * d[0][2] + d[1][1] * d[1][2] + d[2][1] * d[2][2];
    g[5] = d[0][0] * d[0][2] + d[1][0] * d[1][2] + d[2][0] * d[2][2];
    g[2] = d[0][2] * d[0][2] + d[1][2] * d[1][2] + d[2][2] * d[2][2];
  } else {
    const double c1 = 1.259921049894873e+00, c2 = 6.299605249474365e-01;
    double t1, t2, t3;
    t1 = c1 * d[0][0] + c2 * (d[1][0] + d[2][0]);
    t2 = c1 * d[1][0] + c2 * (d[0][0] + d[2][0]);
    t3 = c1 * d[2][0] + c2 * (d[0][0] + d[1][0]);
    g[0] = d[0][0] * t1 + d[1][0] * t2 + d[2][0] * t3;
    t1 = c1 * d[0][1] + c2 * (d[1][1] + d[2][1]);
    t2 = c1 * d[1][1] + c2 * (d[0][1] + d[2][1]);
    t3 = c1 * d[2][1] + c2 * (d[0][1] + d[1][1]);
    g[1] = d[0][1] * t1 + d[1][1] * t2 + d[2][1] * t3;
    g[3] = d[0][0] * t1 + d[1][0] * t2 + d[2][0] * t3;
    t1 = c1 * d[0][2] + c2 * (d[1][2] + d[2][2]);
    t2 = c1 * d[1][2] + c2 * (d[0][2] + d[2][2]);
    t3 = c1 * d[2][2] + c2 * (d[0][2] + d[1][2]);
    g[2] = d[0][2] * t1 + d[1][2] * t2 + d[2][2] * t3;
    g[4] = d[0][1] * t1 + d[1][1] * t2 + d[2][1] * t3;
    g[5] = d[0][0] * t1 + d[1][0] * t2 + d[2][0] * t3;
  }
}

// ---------------------------------------------------------------------------
// AsIq + e3q + e3qvar (asiq.f:1-68, e3q.f:1-100): qres(nshg,9), rmass(nshg) += element contributions
template <int NSHL, int NQ>
__global__ void __launch_bounds__(128) k_inc_asiq(int tab, int numel, size_t numel_pad, int nshg, int numnp,
                                                   const int *__restrict__ ien, const double *__restrict__ x,
                                                   const double *__restrict__ y, double *__restrict__ qres,
                                                   double *__restrict__ rmass) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= numel) return;
  const IncTab &T = c_it[tab];
  int nd[NSHL];
  double xl[NSHL][3], ul[NSHL][3];
#pragma unroll
  for (int a = 0; a < NSHL; a++) {
    nd[a] = ien[(size_t)a * numel_pad + e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      xl[a][i] = __ldg(x + (size_t)numnp * i + nd[a]);
      ul[a][i] = __ldg(y + (size_t)nshg * i + nd[a]);
    }
  }
  double ql[NSHL][9], rm[NSHL];
#pragma unroll
  for (int a = 0; a < NSHL; a++) {
    rm[a] = 0.0;
#pragma unroll
    for (int k = 0; k < 9; k++) ql[a][k] = 0.0;
  }
  const double rmu = c_ip.rmu;
#pragma unroll 1
  for (int q = 0; q < NQ; q++) {
    double shg[NSHL][3], dxidx[3][3], W;
    inc_metric<NSHL>(xl, T.dN[q], T.Qwt[q], shg, dxidx, W);
    double g[3][3];  // g[i][m] = d u_m / d x_i
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int m = 0; m < 3; m++) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < NSHL; n++) s += shg[n][i] * ul[n][m];
        g[i][m] = s;
      }
    double qd[9];  // qdi(1..9), e3q.f:41-49
    qd[0] = 2.0 * rmu * g[0][0];
    qd[3] = rmu * (g[0][1] + g[1][0]);
    qd[6] = rmu * (g[0][2] + g[2][0]);
    qd[1] = rmu * (g[0][1] + g[1][0]);
    qd[4] = 2.0 * rmu * g[1][1];
    qd[7] = rmu * (g[1][2] + g[2][1]);
    qd[2] = rmu * (g[0][2] + g[2][0]);
    qd[5] = rmu * (g[1][2] + g[2][1]);
    qd[8] = 2.0 * rmu * g[2][2];
#pragma unroll
    for (int a = 0; a < NSHL; a++) {
      const double nw = T.N[q][a] * W;
#pragma unroll
      for (int k = 0; k < 9; k++) ql[a][k] += nw * qd[k];
      rm[a] += nw;
    }
  }
#pragma unroll
  for (int a = 0; a < NSHL; a++) {
#pragma unroll
    for (int k = 0; k < 9; k++) atomicAdd(qres + (size_t)nshg * k + nd[a], ql[a][k]);
    atomicAdd(rmass + nd[a], rm[a]);
  }
}

// ---------------------------------------------------------------------------
// what pass 2 of the assembly needs from one quadrature point (e3lhs.f:27-43)
struct IncQP {
  double tlW, tsFct, tauM, tauC, rmu, tauBar, tauMr;  // already scaled by lhsFct * WdetJ as e3LHS does
  double u[3], r[3], uB[3];
};

// bc3LHS (incompressible/bc3lhs.f) on one 3x3 block K[3*r+c]: column operation of the node the block's columns
// belong to, row operation of the node its rows belong to; velocity codes 0 and 7 do nothing (:13-14)
__device__ __noinline__ void inc_bc_col(double *K, int code, double b4, double b5, double b6) {
  if (code == 1 || code == 2 || code == 4) {
    const int pv = code == 1 ? 0 : (code == 2 ? 1 : 2);
    const int o1 = pv == 0 ? 1 : 0, o2 = pv == 2 ? 1 : 2;
    for (int r = 0; r < 3; r++) {
      K[3 * r + o1] = K[3 * r + o1] - b4 * K[3 * r + pv];
      K[3 * r + o2] = K[3 * r + o2] - b5 * K[3 * r + pv];
      K[3 * r + pv] = 0.0;
    }
  } else {
    int p1, p2, fr;
    if (code == 3) { p1 = 0; p2 = 1; fr = 2; }
    else if (code == 5) { p1 = 0; p2 = 2; fr = 1; }
    else { p1 = 1; p2 = 2; fr = 0; }
    for (int r = 0; r < 3; r++) {
      K[3 * r + fr] = K[3 * r + fr] - b4 * K[3 * r + p1] - b6 * K[3 * r + p2];
      K[3 * r + p1] = 0.0;
      K[3 * r + p2] = 0.0;
    }
  }
}
__device__ __noinline__ void inc_bc_row(double *K, int code, double b4, double b5, double b6) {
  if (code == 1 || code == 2 || code == 4) {
    const int pv = code == 1 ? 0 : (code == 2 ? 1 : 2);
    const int o1 = pv == 0 ? 1 : 0, o2 = pv == 2 ? 1 : 2;
    for (int c = 0; c < 3; c++) {
      K[3 * o1 + c] = K[3 * o1 + c] - b4 * K[3 * pv + c];
      K[3 * o2 + c] = K[3 * o2 + c] - b5 * K[3 * pv + c];
      K[3 * pv + c] = 0.0;
    }
  } else {
    int p1, p2, fr;
    if (code == 3) { p1 = 0; p2 = 1; fr = 2; }
    else if (code == 5) { p1 = 0; p2 = 2; fr = 1; }
    else { p1 = 1; p2 = 2; fr = 0; }
    for (int c = 0; c < 3; c++) {
      double v = K[3 * fr + c] - b4 * K[3 * p1 + c];
      // bc3lhs.f:355-361: code 6 drops the BC(:,6) term on the first two of its three row statements
      if (!(code == 6 && c < 2)) v = v - b6 * K[3 * p2 + c];
      K[3 * fr + c] = v;
      K[3 * p1 + c] = 0.0;
      K[3 * p2 + c] = 0.0;
    }
  }
}
__device__ __forceinline__ void inc_bc_diag(double *K, int code) {
  if (code == 1) K[0] = 1.0;
  else if (code == 2) K[4] = 1.0;
  else if (code == 4) K[8] = 1.0;
  else if (code == 3) { K[0] = 1.0; K[4] = 1.0; }
  else if (code == 5) { K[0] = 1.0; K[8] = 1.0; }
  else if (code == 6) { K[4] = 1.0; K[8] = 1.0; }
}

// AsIGMR + e3 (+ bc3LHS + fillsparseI when LHS): one thread per element
template <int NSHL, int NQ, bool LHS>
__global__ void __launch_bounds__(128) k_inc_asigmr(int tab, int numel, size_t numel_pad, int nshg, int numnp,
                                                     const int *__restrict__ ien, const double *__restrict__ x,
                                                     const double *__restrict__ y, const double *__restrict__ ac,
                                                     const double *__restrict__ qres, const int *__restrict__ iBC,
                                                     const double *__restrict__ BC, const int *__restrict__ eloc,
                                                     double *__restrict__ res, double *__restrict__ lhsK,
                                                     double *__restrict__ lhsP) {
  constexpr bool TET = (NSHL == 4);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= numel) return;
  const IncTab &T = c_it[tab];
  const double rho = c_ip.rho, rmu = c_ip.rmu;
  const int iconv = c_ip.iconvflow;
  int nd[NSHL];
  double xl[NSHL][3], yl[NSHL][4], al[NSHL][3];
#pragma unroll
  for (int a = 0; a < NSHL; a++) {
    const int A = ien[(size_t)a * numel_pad + e];
    nd[a] = A;
#pragma unroll
    for (int i = 0; i < 3; i++) xl[a][i] = __ldg(x + (size_t)numnp * i + A);
    yl[a][0] = __ldg(y + (size_t)nshg * 3 + A);  // localy.f:47-72: {u,v,w,p} -> {p,u,v,w}
#pragma unroll
    for (int i = 0; i < 3; i++) {
      yl[a][1 + i] = __ldg(y + (size_t)nshg * i + A);
      al[a][i] = __ldg(ac + (size_t)nshg * i + A);
    }
  }
  double rl[NSHL][4];
#pragma unroll
  for (int a = 0; a < NSHL; a++)
#pragma unroll
    for (int m = 0; m < 4; m++) rl[a][m] = 0.0;
  IncQP qp[LHS ? NQ : 1];
  double shg[NSHL][3], dxidx[3][3], W;
  // ------------------------------------------------------------------ pass 1: residual
#pragma unroll 1
  for (int q = 0; q < NQ; q++) {
    inc_metric<NSHL>(xl, T.dN[q], T.Qwt[q], shg, dxidx, W);
    // e3ivar.f:34-40,88-118
    double pres = 0.0, u[3] = {0.0, 0.0, 0.0}, aci[3] = {0.0, 0.0, 0.0};
    double g[3][4];  // g[i][m] = d Y_m / d x_i, Y = {p,u1,u2,u3}
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int m = 0; m < 4; m++) g[i][m] = 0.0;
#pragma unroll
    for (int n = 0; n < NSHL; n++) {
      const double Nn = T.N[q][n];
      pres += Nn * yl[n][0];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        u[i] += Nn * yl[n][1 + i];
        aci[i] += Nn * al[n][i];
      }
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int m = 0; m < 4; m++) g[i][m] += shg[n][i] * yl[n][m];
    }
    double divq[3] = {0.0, 0.0, 0.0};
    if (c_ip.idiff >= 1) {  // e3ivar.f:122-135: div of the projected diffusive flux
#pragma unroll
      for (int n = 0; n < NSHL; n++) {
#pragma unroll
        for (int m = 0; m < 3; m++)
          divq[m] = divq[m] + shg[n][0] * __ldg(qres + (size_t)nshg * m + nd[n]) +
                    shg[n][1] * __ldg(qres + (size_t)nshg * (3 + m) + nd[n]) +
                    shg[n][2] * __ldg(qres + (size_t)nshg * (6 + m) + nd[n]);
      }
    }
    // e3resStrongPDE (e3res.f:300-400)
    double src[3] = {0.0, 0.0, 0.0};
    if (c_ip.matflg5 == 1) { src[0] = c_ip.bf[0]; src[1] = c_ip.bf[1]; src[2] = c_ip.bf[2]; }
    double r[3];
#pragma unroll
    for (int m = 0; m < 3; m++)
      r[m] = (aci[m] + u[0] * g[0][1 + m] + u[1] * g[1][1 + m] + u[2] * g[2][1 + m] - src[m]) * rho + g[m][0] - divq[m];
    if (iconv == 1) {
      const double divu = (g[0][1] + g[1][2] + g[2][3]) * rho;
#pragma unroll
      for (int m = 0; m < 3; m++) r[m] = r[m] + u[m] * divu;
    }
    // e3stab, itau = 0 (e3stab.f:38-66,205-225)
    double gd[6];
    inc_gijd<TET>(dxidx, gd);
    const double rhoinv = 1.0 / rho, rnu = rmu * rhoinv, dts = c_ip.dts;
    double tauM = ((2.0 * dts) * (2.0 * dts) +
                   (u[0] * (gd[0] * u[0] + gd[3] * u[1] + gd[5] * u[2]) + u[1] * (gd[3] * u[0] + gd[1] * u[1] + gd[4] * u[2]) +
                    u[2] * (gd[5] * u[0] + gd[4] * u[1] + gd[2] * u[2]))) +
                  36.0 * (rnu * rnu) *
                      (gd[0] * gd[0] + gd[1] * gd[1] + gd[2] * gd[2] + 2.0 * (gd[3] * gd[3] + gd[4] * gd[4] + gd[5] * gd[5]));
    const double fact = sqrt(tauM);
    const double tauC = rho * 0.125 * fact / (gd[0] + gd[1] + gd[2]) * c_ip.ff;
    tauM = 1.0 / fact;
    double tauBar = r[0] * (gd[0] * r[0] + gd[3] * r[1] + gd[5] * r[2]) + r[1] * (gd[3] * r[0] + gd[1] * r[1] + gd[4] * r[2]) +
                    r[2] * (gd[5] * r[0] + gd[4] * r[1] + gd[2] * r[2]);
    if (tauBar != 0.0) tauBar = tauM / sqrt(tauBar);
    double uBar[3];
#pragma unroll
    for (int m = 0; m < 3; m++) uBar[m] = u[m] - tauM * r[m] * rhoinv;
    // e3Res (e3res.f:33-200)
    double rNa[3], rG[3][3];
#pragma unroll
    for (int m = 0; m < 3; m++) rNa[m] = aci[m] * c_ip.tmps - src[m];
    const double tmp = -pres + tauC * (g[0][1] + g[1][2] + g[2][3]);
    const double s12 = rmu * (g[1][1] + g[0][2]);
    const double s23 = rmu * (g[2][2] + g[1][3]);
    const double s13 = rmu * (g[0][3] + g[2][1]);
    rG[0][0] = 2.0 * rmu * g[0][1] + tmp;
    rG[0][1] = s12;
    rG[0][2] = s13;
    rG[1][0] = s12;
    rG[1][1] = 2.0 * rmu * g[1][2] + tmp;
    rG[1][2] = s23;
    rG[2][0] = s13;
    rG[2][1] = s23;
    rG[2][2] = 2.0 * rmu * g[2][3] + tmp;
    if (iconv == 2) {
#pragma unroll
      for (int m = 0; m < 3; m++)
        rNa[m] = rNa[m] + uBar[0] * g[0][1 + m] + uBar[1] * g[1][1 + m] + uBar[2] * g[2][1 + m];
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] - u[i] * u[j] * rho;
    }
    {
      double t[3] = {tauM * r[0], tauM * r[1], tauM * r[2]};
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] + t[i] * u[j];
      if (iconv == 1) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] + t[j] * u[i];
      }
      if (iconv == 2) {
#pragma unroll
        for (int m = 0; m < 3; m++) t[m] = tauBar * (r[0] * g[0][1 + m] + r[1] * g[1][1 + m] + r[2] * g[2][1 + m]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] + t[i] * r[j];
      }
    }
#pragma unroll
    for (int m = 0; m < 3; m++) rNa[m] = rNa[m] * rho;
#pragma unroll
    for (int a = 0; a < NSHL; a++) {
      const double Na = T.N[q][a];
      rl[a][3] = rl[a][3] + W * (shg[a][0] * uBar[0] + shg[a][1] * uBar[1] + shg[a][2] * uBar[2]);
#pragma unroll
      for (int m = 0; m < 3; m++)
        rl[a][m] = rl[a][m] - W * (Na * rNa[m] + shg[a][0] * rG[m][0] + shg[a][1] * rG[m][1] + shg[a][2] * rG[m][2]);
    }
    if (LHS) {  // e3lhs.f:27-43
      IncQP &Q = qp[q];
      const double tlW = c_ip.lhsFct * W;
      const double t1 = tlW * rho;
      Q.tlW = tlW;
      Q.tsFct = c_ip.lhmFct * W * rho;
      Q.tauM = tlW * tauM;
      Q.tauC = tlW * tauC;
      Q.rmu = tlW * rmu;
      Q.tauMr = Q.tauM / rho;
      if (iconv == 2) {
        Q.tauBar = c_ip.lhsFct * W * tauBar;
#pragma unroll
        for (int m = 0; m < 3; m++) Q.uB[m] = t1 * uBar[m];
      } else {
        Q.tauBar = 0.0;
#pragma unroll
        for (int m = 0; m < 3; m++) Q.uB[m] = t1 * u[m];
      }
#pragma unroll
      for (int m = 0; m < 3; m++) { Q.u[m] = u[m]; Q.r[m] = r[m]; }
    }
  }
  // local(res, rl, 'scatter') (common/local.f:67-74)
#pragma unroll
  for (int a = 0; a < NSHL; a++)
#pragma unroll
    for (int m = 0; m < 4; m++) atomicAdd(res + (size_t)nshg * m + nd[a], rl[a][m]);
  if (!LHS) return;
  // ------------------------------------------------------------------ pass 2: tangent, one block row at a time
  int code[NSHL];
#pragma unroll
  for (int a = 0; a < NSHL; a++) code[a] = (__ldg(iBC + nd[a]) >> 3) & 7;
#pragma unroll 1
  for (int a = 0; a < NSHL; a++) {
#pragma unroll 1
    for (int b = 0; b < NSHL; b++) {
      double K[9], G[4];
#pragma unroll
      for (int k = 0; k < 9; k++) K[k] = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) G[k] = 0.0;
      const int lo = a < b ? a : b, hi = a < b ? b : a;
#pragma unroll 1
      for (int q = 0; q < NQ; q++) {
        if (!TET || q == 0) inc_metric<NSHL>(xl, T.dN[q], T.Qwt[q], shg, dxidx, W);  // linear tets: constant
        const IncQP &Q = qp[q];
        const double Na = T.N[q][a], Nb = T.N[q][b];
        const double *gb = shg[b], *gl = shg[lo], *gh = shg[hi];
        // mass + advection (e3lhs.f:49-60)
        const double t1u = Q.uB[0] * gb[0] + Q.uB[1] * gb[1] + Q.uB[2] * gb[2];
        const double x2 = Q.tsFct * Na * Nb + t1u * Na;
        K[0] += x2;
        K[4] += x2;
        K[8] += x2;
        // diffusion + SUPG + continuity stabilisation (e3lhs.f:64-171): computed by the reference for b <= aa
        // from the lower node's t1, t2, t3 and mirrored into (b,aa)
        double t1[3], t2[3], t3[3];
        const double y1 = Q.tauM * (Q.u[0] * gl[0] + Q.u[1] * gl[1] + Q.u[2] * gl[2]) * rho;
        const double y2 = Q.tauBar * (Q.r[0] * gl[0] + Q.r[1] * gl[1] + Q.r[2] * gl[2]);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          t1[i] = Q.tauC * gl[i];
          t2[i] = Q.rmu * gl[i];
          t3[i] = t2[i] + y1 * Q.u[i] + y2 * Q.r[i];
        }
        const double tmp = t3[0] * gh[0] + t3[1] * gh[1] + t3[2] * gh[2];
        if (a == b) {
          K[0] = K[0] + tmp + t1[0] * gh[0] + t2[0] * gh[0];
          K[4] = K[4] + tmp + t1[1] * gh[1] + t2[1] * gh[1];
          K[8] = K[8] + tmp + t1[2] * gh[2] + t2[2] * gh[2];
          double z = t1[0] * gh[1] + t2[1] * gh[0];
          K[1] += z;
          K[3] += z;
          z = t1[0] * gh[2] + t2[2] * gh[0];
          K[2] += z;
          K[6] += z;
          z = t1[1] * gh[2] + t2[2] * gh[1];
          K[5] += z;
          K[7] += z;
        } else {
          K[0] += tmp + t1[0] * gh[0] + t2[0] * gh[0];
          K[4] += tmp + t1[1] * gh[1] + t2[1] * gh[1];
          K[8] += tmp + t1[2] * gh[2] + t2[2] * gh[2];
          // M[i][j] = t1[i] gh[j] + t2[j] gh[i]; lower blocks (a > b) take M, mirrored ones its transpose
          const double m01 = t1[0] * gh[1] + t2[1] * gh[0], m02 = t1[0] * gh[2] + t2[2] * gh[0];
          const double m10 = t1[1] * gh[0] + t2[0] * gh[1], m12 = t1[1] * gh[2] + t2[2] * gh[1];
          const double m20 = t1[2] * gh[0] + t2[0] * gh[2], m21 = t1[2] * gh[1] + t2[1] * gh[2];
          if (a > b) {
            K[1] += m01; K[2] += m02; K[3] += m10; K[5] += m12; K[6] += m20; K[7] += m21;
          } else {
            K[3] += m01; K[6] += m02; K[1] += m10; K[7] += m12; K[2] += m20; K[5] += m21;
          }
        }
        // G (e3lhs.f:176-186) and C (:191-203, mirrored in e3.f:108-113)
#pragma unroll
        for (int i = 0; i < 3; i++) G[i] += (Q.tlW * gb[i]) * Na;
        G[3] = G[3] + (Q.tauMr * gl[0]) * gh[0] + (Q.tauMr * gl[1]) * gh[1] + (Q.tauMr * gl[2]) * gh[2];
      }
      // bc3LHS: the lower local node's operation first (bc3lhs.f loops inod = 1..nshl, columns then rows)
      const int ca = code[a], cb = code[b];
      const bool ra = (ca != 0 && ca != 7), rb = (cb != 0 && cb != 7);
      if (ra || rb) {
        const double a4 = __ldg(BC + (size_t)nshg * 3 + nd[a]), a5 = __ldg(BC + (size_t)nshg * 4 + nd[a]),
                     a6 = __ldg(BC + (size_t)nshg * 5 + nd[a]);
        const double b4 = __ldg(BC + (size_t)nshg * 3 + nd[b]), b5 = __ldg(BC + (size_t)nshg * 4 + nd[b]),
                     b6 = __ldg(BC + (size_t)nshg * 5 + nd[b]);
        if (a < b) {
          if (ra) inc_bc_row(K, ca, a4, a5, a6);
          if (rb) inc_bc_col(K, cb, b4, b5, b6);
        } else {
          if (rb) inc_bc_col(K, cb, b4, b5, b6);
          if (ra) inc_bc_row(K, ca, a4, a5, a6);
          if (a == b) inc_bc_diag(K, ca);
        }
      }
      // fillsparseI (common/fillsparse.f:20-46)
      const size_t k = (size_t)eloc[(size_t)(NSHL * a + b) * numel_pad + e];
#pragma unroll
      for (int m = 0; m < 9; m++) atomicAdd(lhsK + 9 * k + m, K[m]);
#pragma unroll
      for (int m = 0; m < 4; m++) atomicAdd(lhsP + 4 * k + m, G[m]);
    }
  }
}

// ---------------------------------------------------------------------------
// Linear tets: N_a,i, the metric tensor, grad Y and div q are constant over the element, so the point loop only
// interpolates u, Y,t, p and evaluates tau; the tangent needs no point loop at all once a handful of point sums
// are kept (e3lhs.f with constant N_a,i):
//   K_ij(a,b) = d_ij [ M_ab + UB_a . g_b + mu~ g_a.g_b + g_a^T UU g_b + g_a^T RR g_b ] + tauC~ g_b,i g_a,j + mu~ g_b,j g_a,i
//   G_i(a,b)  = TL_a g_b,i        C(a,b) = tauMr~ g_a.g_b
// with M_ab = sum_q tsFct N_a N_b, UB_a = sum_q rho tlW ubar N_a, TL_a = sum_q tlW N_a, UU = sum_q tlW tauM rho u u^T,
// RR = sum_q tlW tauBar r r^T, x~ = sum_q tlW x.  Node data comes from the 208-byte node records (13 16-byte loads
// per node); the 13 entries of every block go through a per-warp shared tile so that one reduction instruction
// covers whole CSR blocks (4 L2 sectors each) instead of one double of 32 different blocks.
#define INC_NREC 26
#ifndef INC_TET_MINB
#define INC_TET_MINB 3
#endif
template <int NQ, bool LHS>
__global__ void __launch_bounds__(128, INC_TET_MINB) k_inc_asigmr_tet(int numel, size_t numel_pad, int nshg,
                                                         const int *__restrict__ ien, const double *__restrict__ aos,
                                                         const int *__restrict__ iBC, const double *__restrict__ BC,
                                                         const int *__restrict__ eloc, double *__restrict__ res,
                                                         double *__restrict__ lhsK, double *__restrict__ lhsP) {
  __shared__ double stage_all[LHS ? 4 * 32 * 13 : 1];
  const int e0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e0 < numel;
  const int e = valid ? e0 : numel - 1;
  const int lane = threadIdx.x & 31;
  const IncTab &T = c_it[0];
  const double rho = c_ip.rho, rmu = c_ip.rmu;
  const int iconv = c_ip.iconvflow;
  int nd[4];
  double xl[4][3], yl[4][4], al[4][3], ql[4][9];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int A = ien[(size_t)a * numel_pad + e];
    nd[a] = A;
    const double2 *rec = reinterpret_cast<const double2 *>(aos + (size_t)A * INC_NREC);
    double v[INC_NREC];
#pragma unroll
    for (int k = 0; k < 11; k++) {  // x(3) Y(5) Y,t(5) q(9) = 22 doubles
      const double2 t = __ldg(rec + k);
      v[2 * k] = t.x;
      v[2 * k + 1] = t.y;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) xl[a][i] = v[i];
#pragma unroll
    for (int m = 0; m < 4; m++) yl[a][m] = v[3 + m];
#pragma unroll
    for (int i = 0; i < 3; i++) al[a][i] = v[9 + i];
#pragma unroll
    for (int k = 0; k < 9; k++) ql[a][k] = v[13 + k];
  }
  double shg[4][3], dxidx[3][3], W1;
  inc_metric<4>(xl, T.dN[0], 1.0, shg, dxidx, W1);
  double g[3][4];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int m = 0; m < 4; m++) g[i][m] = shg[0][i] * yl[0][m] + shg[1][i] * yl[1][m] + shg[2][i] * yl[2][m] + shg[3][i] * yl[3][m];
  double divq[3] = {0.0, 0.0, 0.0};
  if (c_ip.idiff >= 1) {
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
      for (int m = 0; m < 3; m++) divq[m] = divq[m] + shg[n][0] * ql[n][m] + shg[n][1] * ql[n][3 + m] + shg[n][2] * ql[n][6 + m];
  }
  double gd[6];
  inc_gijd<true>(dxidx, gd);
  const double trg = gd[0] + gd[1] + gd[2];
  const double rhoinv = 1.0 / rho, rnu = rmu * rhoinv, dts = c_ip.dts;
  const double visc2 = 36.0 * (rnu * rnu) *
                       (gd[0] * gd[0] + gd[1] * gd[1] + gd[2] * gd[2] + 2.0 * (gd[3] * gd[3] + gd[4] * gd[4] + gd[5] * gd[5]));
  double src[3] = {0.0, 0.0, 0.0};
  if (c_ip.matflg5 == 1) { src[0] = c_ip.bf[0]; src[1] = c_ip.bf[1]; src[2] = c_ip.bf[2]; }
  const double divu = g[0][1] + g[1][2] + g[2][3];
  const double s12 = rmu * (g[1][1] + g[0][2]), s23 = rmu * (g[2][2] + g[1][3]), s13 = rmu * (g[0][3] + g[2][1]);
  double rl[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int m = 0; m < 4; m++) rl[a][m] = 0.0;
  // point sums for the tangent
  double UB[4][3], UU[6], RR[6], sC = 0.0, sMr = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) { UU[k] = 0.0; RR[k] = 0.0; }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 3; i++) UB[a][i] = 0.0;
#pragma unroll 1
  for (int q = 0; q < NQ; q++) {
    const double W = T.Qwt[q] * W1;
    double pres = 0.0, u[3] = {0.0, 0.0, 0.0}, aci[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int n = 0; n < 4; n++) {
      const double Nn = T.N[q][n];
      pres += Nn * yl[n][0];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        u[i] += Nn * yl[n][1 + i];
        aci[i] += Nn * al[n][i];
      }
    }
    double r[3];
#pragma unroll
    for (int m = 0; m < 3; m++)
      r[m] = (aci[m] + u[0] * g[0][1 + m] + u[1] * g[1][1 + m] + u[2] * g[2][1 + m] - src[m]) * rho + g[m][0] - divq[m];
    if (iconv == 1) {
#pragma unroll
      for (int m = 0; m < 3; m++) r[m] = r[m] + u[m] * (divu * rho);
    }
    double tauM = ((2.0 * dts) * (2.0 * dts) +
                   (u[0] * (gd[0] * u[0] + gd[3] * u[1] + gd[5] * u[2]) + u[1] * (gd[3] * u[0] + gd[1] * u[1] + gd[4] * u[2]) +
                    u[2] * (gd[5] * u[0] + gd[4] * u[1] + gd[2] * u[2]))) + visc2;
    const double fact = sqrt(tauM);
    const double tauC = rho * 0.125 * fact / trg * c_ip.ff;
    tauM = 1.0 / fact;
    double tauBar = r[0] * (gd[0] * r[0] + gd[3] * r[1] + gd[5] * r[2]) + r[1] * (gd[3] * r[0] + gd[1] * r[1] + gd[4] * r[2]) +
                    r[2] * (gd[5] * r[0] + gd[4] * r[1] + gd[2] * r[2]);
    if (tauBar != 0.0) tauBar = tauM * rsqrt(tauBar);
    double uBar[3];
#pragma unroll
    for (int m = 0; m < 3; m++) uBar[m] = u[m] - tauM * r[m] * rhoinv;
    double rNa[3], rG[3][3];
#pragma unroll
    for (int m = 0; m < 3; m++) rNa[m] = aci[m] * c_ip.tmps - src[m];
    const double tmp = -pres + tauC * divu;
    rG[0][0] = 2.0 * rmu * g[0][1] + tmp; rG[0][1] = s12; rG[0][2] = s13;
    rG[1][0] = s12; rG[1][1] = 2.0 * rmu * g[1][2] + tmp; rG[1][2] = s23;
    rG[2][0] = s13; rG[2][1] = s23; rG[2][2] = 2.0 * rmu * g[2][3] + tmp;
    if (iconv == 2) {
#pragma unroll
      for (int m = 0; m < 3; m++) rNa[m] = rNa[m] + uBar[0] * g[0][1 + m] + uBar[1] * g[1][1 + m] + uBar[2] * g[2][1 + m];
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] - u[i] * u[j] * rho;
    }
    {
      double t[3] = {tauM * r[0], tauM * r[1], tauM * r[2]};
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] + t[i] * u[j];
      if (iconv == 1) {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] + t[j] * u[i];
      } else {
#pragma unroll
        for (int m = 0; m < 3; m++) t[m] = tauBar * (r[0] * g[0][1 + m] + r[1] * g[1][1 + m] + r[2] * g[2][1 + m]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) rG[i][j] = rG[i][j] + t[i] * r[j];
      }
    }
#pragma unroll
    for (int m = 0; m < 3; m++) rNa[m] = rNa[m] * rho;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double Na = T.N[q][a];
      rl[a][3] = rl[a][3] + W * (shg[a][0] * uBar[0] + shg[a][1] * uBar[1] + shg[a][2] * uBar[2]);
#pragma unroll
      for (int m = 0; m < 3; m++)
        rl[a][m] = rl[a][m] - W * (Na * rNa[m] + shg[a][0] * rG[m][0] + shg[a][1] * rG[m][1] + shg[a][2] * rG[m][2]);
    }
    if (LHS) {
      const double tlW = c_ip.lhsFct * W;
      const double tM = tlW * tauM;
      sC += tlW * tauC;
      sMr += tM / rho;
      const double tMr = tM * rho;
      UU[0] += tMr * u[0] * u[0]; UU[1] += tMr * u[1] * u[1]; UU[2] += tMr * u[2] * u[2];
      UU[3] += tMr * u[0] * u[1]; UU[4] += tMr * u[1] * u[2]; UU[5] += tMr * u[0] * u[2];
      if (iconv == 2) {
        const double tB = tlW * tauBar;
        RR[0] += tB * r[0] * r[0]; RR[1] += tB * r[1] * r[1]; RR[2] += tB * r[2] * r[2];
        RR[3] += tB * r[0] * r[1]; RR[4] += tB * r[1] * r[2]; RR[5] += tB * r[0] * r[2];
      }
      const double t1 = tlW * rho;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const double w = t1 * T.N[q][a];
#pragma unroll
        for (int i = 0; i < 3; i++) UB[a][i] += w * (iconv == 2 ? uBar[i] : u[i]);
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int m = 0; m < 4; m++) atomicAdd(res + (size_t)nshg * m + nd[a], rl[a][m]);
  }
  if (!LHS) return;
  // ------------------------------------------------------------------ tangent blocks
  double *stage = stage_all + (threadIdx.x >> 5) * (32 * 13);
  if (iconv == 2) {
#pragma unroll
    for (int k = 0; k < 6; k++) UU[k] += RR[k];
  }
  double sW = 0.0;
#pragma unroll
  for (int q = 0; q < NQ; q++) sW += T.Qwt[q];
  const double sMu = c_ip.lhsFct * W1 * sW * rmu;      // sum_q tlW rmu
  const double cM = c_ip.lhmFct * W1 * rho, cTL = c_ip.lhsFct * W1;
  int code[4];
#pragma unroll
  for (int a = 0; a < 4; a++) code[a] = (__ldg(iBC + nd[a]) >> 3) & 7;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    double TLa = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; q++) TLa += T.Qwt[q] * T.N[q][a];
    TLa *= cTL;
    // h_a = (UU + RR) g_a
    const double *ga = shg[a];
    const double ha[3] = {UU[0] * ga[0] + UU[3] * ga[1] + UU[5] * ga[2], UU[3] * ga[0] + UU[1] * ga[1] + UU[4] * ga[2],
                          UU[5] * ga[0] + UU[4] * ga[1] + UU[2] * ga[2]};
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const double *gb = shg[b];
      double Mab = 0.0;
#pragma unroll
      for (int q = 0; q < NQ; q++) Mab += T.Qwt[q] * T.N[q][a] * T.N[q][b];
      const double gg = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
      const double dcom = cM * Mab + (UB[a][0] * gb[0] + UB[a][1] * gb[1] + UB[a][2] * gb[2]) + sMu * gg +
                          (ha[0] * gb[0] + ha[1] * gb[1] + ha[2] * gb[2]);
      double K[9], G[4];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) K[3 * i + j] = sC * gb[i] * ga[j] + sMu * gb[j] * ga[i] + (i == j ? dcom : 0.0);
#pragma unroll
      for (int i = 0; i < 3; i++) G[i] = TLa * gb[i];
      G[3] = sMr * gg;
      const int ca = code[a], cb = code[b];
      const bool ra = (ca != 0 && ca != 7), rb = (cb != 0 && cb != 7);
      if (ra || rb) {
        double Kt[9];
#pragma unroll
        for (int k = 0; k < 9; k++) Kt[k] = K[k];
        const double a4 = __ldg(BC + (size_t)nshg * 3 + nd[a]), a5 = __ldg(BC + (size_t)nshg * 4 + nd[a]),
                     a6 = __ldg(BC + (size_t)nshg * 5 + nd[a]);
        const double b4 = __ldg(BC + (size_t)nshg * 3 + nd[b]), b5 = __ldg(BC + (size_t)nshg * 4 + nd[b]),
                     b6 = __ldg(BC + (size_t)nshg * 5 + nd[b]);
        if (a < b) {
          if (ra) inc_bc_row(Kt, ca, a4, a5, a6);
          if (rb) inc_bc_col(Kt, cb, b4, b5, b6);
        } else {
          if (rb) inc_bc_col(Kt, cb, b4, b5, b6);
          if (ra) inc_bc_row(Kt, ca, a4, a5, a6);
          if (a == b) inc_bc_diag(Kt, ca);
        }
#pragma unroll
        for (int k = 0; k < 9; k++) K[k] = Kt[k];
      }
      // fillsparseI through the per-warp tile: lanes 0..12 / 16..28 add the 13 entries of two blocks per step.
      // (The reduction rate is set by the SM's REDG issue, ~1.3 cycles per lane; grouping equal slots of the warp
      // with match.any before the reduction was measured and costs more than the 2x fewer reductions save.)
      const int kslot = valid ? eloc[(size_t)(4 * a + b) * numel_pad + e] : -1;
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 9; m++) stage[lane * 13 + m] = K[m];
#pragma unroll
      for (int m = 0; m < 4; m++) stage[lane * 13 + 9 + m] = G[m];
      __syncwarp();
      const int m = lane & 15, half = lane >> 4;
#pragma unroll
      for (int it = 0; it < 16; it++) {
        const int el = 2 * it + half;
        const int kk = __shfl_sync(0xffffffffu, kslot, el);
        if (m < 13 && kk >= 0) {
          const double v = stage[el * 13 + m];
          if (m < 9) atomicAdd(lhsK + (size_t)9 * kk + m, v);
          else atomicAdd(lhsP + (size_t)4 * kk + (m - 9), v);
        }
      }
    }
  }
}

// bc3Res (incompressible/bc3res.f:1-90, intpres=0) after bc3per: node-wise
__global__ void k_inc_bc3res(int nshg, const int *__restrict__ iBC, const double *__restrict__ BC, double *res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nshg) return;
  const int ib = iBC[i], code = (ib >> 3) & 7;
  if (!(ib & 4) && code == 0 && !(ib & (1 << 11))) return;
  double r1 = res[i], r2 = res[(size_t)nshg + i], r3 = res[(size_t)2 * nshg + i];
  const double b4 = BC[(size_t)nshg * 3 + i], b5 = BC[(size_t)nshg * 4 + i], b6 = BC[(size_t)nshg * 5 + i];
  if (ib & 4) res[(size_t)3 * nshg + i] = 0.0;
  switch (code) {
    case 1: r2 = r2 - b4 * r1; r3 = r3 - b5 * r1; r1 = 0.0; break;
    case 2: r1 = r1 - b4 * r2; r3 = r3 - b5 * r2; r2 = 0.0; break;
    case 3: r3 = r3 - b4 * r1 - b6 * r2; r1 = 0.0; r2 = 0.0; break;
    case 4: r1 = r1 - b4 * r3; r2 = r2 - b5 * r3; r3 = 0.0; break;
    case 5: r2 = r2 - b4 * r1 - b6 * r3; r1 = 0.0; r3 = 0.0; break;
    case 6: r1 = r1 - b4 * r2 - b6 * r3; r2 = 0.0; r3 = 0.0; break;
    case 7: r1 = r2 = r3 = 0.0; break;
    default: break;
  }
  if (ib & (1 << 11)) r1 = r2 = r3 = 0.0;
  res[i] = r1;
  res[(size_t)nshg + i] = r2;
  res[(size_t)2 * nshg + i] = r3;
}

template <int NSHL, int NQ>
static int launch_group(phb200_ctx *ctx, int tab, int numel, size_t numel_pad, const int *d_ien, const int *d_eloc,
                        bool lhs, bool asiq) {
  const phb200_common &c = ctx->c;
  const int nb = (numel + 127) / 128;
  cudaStream_t s = ctx->stream;
  if (asiq) {
    KScope ks(ctx, KC_ASIQ);
    k_inc_asiq<NSHL, NQ><<<nb, 128, 0, s>>>(tab, numel, numel_pad, c.nshg, c.numnp, d_ien, ctx->d_x, ctx->d_y,
                                            ctx->d_qres, ctx->d_rmass);
  } else {
    KScope ks(ctx, KC_ASM);
    if (lhs)
      k_inc_asigmr<NSHL, NQ, true><<<nb, 128, 0, s>>>(tab, numel, numel_pad, c.nshg, c.numnp, d_ien, ctx->d_x, ctx->d_y,
                                                      ctx->d_ac, ctx->d_qres, ctx->d_iBC, ctx->d_BC, d_eloc,
                                                      ctx->d_res4, ctx->d_lhsK9, ctx->d_lhsP4);
    else
      k_inc_asigmr<NSHL, NQ, false><<<nb, 128, 0, s>>>(tab, numel, numel_pad, c.nshg, c.numnp, d_ien, ctx->d_x, ctx->d_y,
                                                       ctx->d_ac, ctx->d_qres, ctx->d_iBC, ctx->d_BC, d_eloc,
                                                       ctx->d_res4, ctx->d_lhsK9, ctx->d_lhsP4);
  }
  PHB_CHECK(cudaGetLastError());
  return 0;
}

template <int NQ>
static int launch_tet(phb200_ctx *ctx, bool lhs) {
  KScope ks(ctx, KC_ASM);
  const int nb = (ctx->numel_tet + 127) / 128;
  if (lhs)
    k_inc_asigmr_tet<NQ, true><<<nb, 128, 0, ctx->stream>>>(ctx->numel_tet, ctx->numel_pad, ctx->c.nshg, ctx->d_ien,
                                                            ctx->d_nodeaos, ctx->d_iBC, ctx->d_BC, ctx->d_eloc,
                                                            ctx->d_res4, ctx->d_lhsK9, ctx->d_lhsP4);
  else
    k_inc_asigmr_tet<NQ, false><<<nb, 128, 0, ctx->stream>>>(ctx->numel_tet, ctx->numel_pad, ctx->c.nshg, ctx->d_ien,
                                                             ctx->d_nodeaos, ctx->d_iBC, ctx->d_BC, ctx->d_eloc,
                                                             ctx->d_res4, ctx->d_lhsK9, ctx->d_lhsP4);
  PHB_CHECK(cudaGetLastError());
  return 0;
}

static int launch_all(phb200_ctx *ctx, bool lhs, bool asiq) {
  static const bool generic_tets = getenv("PHB200_INC_GENERIC") && atoi(getenv("PHB200_INC_GENERIC")) != 0;
  if (ctx->numel_tet > 0 && !asiq && !generic_tets && ctx->tet_uniform_rule) {
    const int nq = ctx->c.nint[0];
    PHB_TRY(phb_pack_nodes(ctx, ctx->inc_idiff >= 1));
    if (nq == 4) PHB_TRY(launch_tet<4>(ctx, lhs));
    else if (nq == 1) PHB_TRY(launch_tet<1>(ctx, lhs));
    else { fprintf(stderr, "phb200: inc_elmgmr: tet rule with %d points not supported\n", nq); return 1; }
  } else if (ctx->numel_tet > 0) {
    const int nq = ctx->c.nint[0];
    if (nq == 4) PHB_TRY((launch_group<4, 4>(ctx, 0, ctx->numel_tet, ctx->numel_pad, ctx->d_ien, ctx->d_eloc, lhs, asiq)));
    else if (nq == 1) PHB_TRY((launch_group<4, 1>(ctx, 0, ctx->numel_tet, ctx->numel_pad, ctx->d_ien, ctx->d_eloc, lhs, asiq)));
    else { fprintf(stderr, "phb200: inc_elmgmr: tet rule with %d points not supported\n", nq); return 1; }
  }
  for (const ElemGroup &g : ctx->gen) {
    if (g.nshl == 8 && g.nq == 8)
      PHB_TRY((launch_group<8, 8>(ctx, 1, g.numel, g.numel_pad, g.d_ien, g.d_eloc, lhs, asiq)));
    else if (g.nshl == 6 && g.nq == 6)
      PHB_TRY((launch_group<6, 6>(ctx, 2, g.numel, g.numel_pad, g.d_ien, g.d_eloc, lhs, asiq)));
    else { fprintf(stderr, "phb200: inc_elmgmr: topology nshl=%d with %d points not supported\n", g.nshl, g.nq); return 1; }
  }
  return 0;
}

// transposed-entry map: tpos[k] = slot of (j,i) for the entry k = (i,j); genadj's pattern is symmetric
__global__ void k_tpos(int nnz_tot, const int *__restrict__ rowofblk, const int *__restrict__ colm,
                       const int *__restrict__ rowp, int *__restrict__ tpos, int *bad) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz_tot) return;
  const int i = rowofblk[k], j = rowp[k];
  int lo = colm[j], hi = colm[j + 1] - 1, found = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1, v = rowp[mid];
    if (v == i) { found = mid; break; }
    if (v < i) lo = mid + 1; else hi = mid - 1;
  }
  if (found < 0) { atomicAdd(bad, 1); found = k; }
  tpos[k] = found;
}

static int inc_alloc(phb200_ctx *ctx) {
  const size_t nshg = ctx->c.nshg, nnz = (size_t)std::max(ctx->nnz_tot, 1);
  if (!ctx->d_res4) {
    PHB_CHECK(cudaMalloc(&ctx->d_res4, sizeof(double) * 4 * nshg));
    PHB_CHECK(cudaMalloc(&ctx->d_lesp, sizeof(double) * 4 * nshg));
    PHB_CHECK(cudaMalloc(&ctx->d_lesq, sizeof(double) * 4 * nshg));
    PHB_CHECK(cudaMalloc(&ctx->d_lesp4, sizeof(double) * 4 * nshg));
    PHB_CHECK(cudaMemsetAsync(ctx->d_lesp, 0, sizeof(double) * 4 * nshg, ctx->stream));
  }
  if (!ctx->d_lhsK9 && ctx->nnz_tot > 0) {
    PHB_CHECK(cudaMalloc(&ctx->d_lhsK9, sizeof(double) * 9 * nnz));
    PHB_CHECK(cudaMalloc(&ctx->d_lhsP4, sizeof(double) * 4 * nnz));
    PHB_CHECK(cudaMalloc(&ctx->d_tpos, sizeof(int) * nnz));
    int *d_bad;
    PHB_CHECK(cudaMalloc(&d_bad, sizeof(int)));
    PHB_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    k_tpos<<<(ctx->nnz_tot + 255) / 256, 256, 0, ctx->stream>>>(ctx->nnz_tot, ctx->d_rowofblk, ctx->d_colm, ctx->d_rowp,
                                                                ctx->d_tpos, d_bad);
    ctx->launches++;
    int bad = 0;
    PHB_CHECK(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PHB_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bad);
    if (bad) {
      fprintf(stderr, "phb200: inc_elmgmr: colm/rowp is not structurally symmetric (%d entries)\n", bad);
      return 1;
    }
  }
  return 0;
}

void phb_inc_free(phb200_ctx *ctx) {
  void *p[] = {ctx->d_res4, ctx->d_lhsK9, ctx->d_lhsP4, ctx->d_lesp, ctx->d_lesq, ctx->d_lesp4, ctx->d_tpos};
  for (void *q : p)
    if (q) cudaFree(q);
  ctx->d_res4 = ctx->d_lhsK9 = ctx->d_lhsP4 = ctx->d_lesp = ctx->d_lesq = ctx->d_lesp4 = nullptr;
  ctx->d_tpos = nullptr;
  if (ctx->d_nsrflist) cudaFree(ctx->d_nsrflist);
  ctx->d_nsrflist = nullptr;
  ctx->have_inc_btabs = false;
}

// the boundary blocks of ElmGMR (incompressible/elmgmr.f:246-320): flxID = 0 (elmgmr.f:130), AsBMFG per block
static int inc_boundary(phb200_ctx *ctx, const phb200_incomp *ip) {
  const phb200_common &c = ctx->c;
  cudaStream_t s = ctx->stream;
  const int nshg = c.nshg;
  if (!ctx->have_inc_btabs) {
    const double *shpb = ctx->h_shpb.data(), *shglb = ctx->h_shglb.data();
    auto fill = [&](int lcsyst, int nshl) -> int {
      BndTables b;
      if (phb_bnd_fill_tables(&b, lcsyst, nshl, c.nintb, c.Qwtb, shpb, shglb)) return 1;
      PHB_CHECK(cudaMemcpyToSymbol(c_ibnd, &b, sizeof b, sizeof(BndTables) * (lcsyst - 1)));
      return 0;
    };
    if (ctx->numelb > 0) PHB_TRY(fill(1, 4));
    for (const BndGroup &g : ctx->bgen) PHB_TRY(fill(g.lcsyst, g.nshl));
    PHB_CHECK(cudaMalloc(&ctx->d_nsrflist, sizeof(int) * 1001));
    ctx->have_inc_btabs = true;
  }
  IncBndPhys bp;
  bp.rho = ip->rho; bp.rmu = ip->rmu;
  bp.iviscflux = ip->iviscflux; bp.iconvflow = ip->iconvflow; bp.itwmod = ip->itwmod;
  PHB_CHECK(cudaMemcpyToSymbolAsync(c_ibp, &bp, sizeof bp, 0, cudaMemcpyHostToDevice, s));
  if (ip->nsrflist) {
    // pageable source: the copy is staged before the call returns, the caller's array may go away afterwards
    PHB_CHECK(cudaMemcpyAsync(ctx->d_nsrflist, ip->nsrflist, sizeof(int) * 1001, cudaMemcpyHostToDevice, s));
  } else {
    PHB_CHECK(cudaMemsetAsync(ctx->d_nsrflist, 0, sizeof(int) * 1001, s));
  }
  PHB_CHECK(cudaMemsetAsync(ctx->d_aerfrc + 4, 0, sizeof(double) * 10 * 1001, s));
  KScope ks(ctx, KC_ASM);
  long long nl = 0;
#define PHB_IBND_LAUNCH(NSHL, NSHLB, LCS, N, IEN, IB, BCBP)                                                       \
  k_inc_asbmfg<NSHL, NSHLB, LCS><<<((N) + 127) / 128, 128, 0, s>>>((N), nshg, c.numnp, (IEN), (IB), (BCBP), ctx->d_x, \
                                                                  ctx->d_y, ctx->d_nsrflist, ctx->d_res4, ctx->d_aerfrc)
  if (ctx->numelb > 0) {
    PHB_IBND_LAUNCH(4, 3, 1, ctx->numelb, ctx->d_ienb, ctx->d_iBCB, ctx->d_BCB);
    nl++;
  }
  for (const BndGroup &g : ctx->bgen) {
    if (g.lcsyst == 2) PHB_IBND_LAUNCH(8, 4, 2, g.n, g.d_ien, g.d_iBCB, g.d_BCB);
    else if (g.lcsyst == 3) PHB_IBND_LAUNCH(6, 3, 3, g.n, g.d_ien, g.d_iBCB, g.d_BCB);
    else PHB_IBND_LAUNCH(6, 4, 4, g.n, g.d_ien, g.d_iBCB, g.d_BCB);
    nl++;
  }
#undef PHB_IBND_LAUNCH
  ctx->launches += nl - 1;  // KScope counted one
  PHB_CHECK(cudaGetLastError());
  return 0;
}

// ElmGMR (incompressible/elmgmr.f:1-330) on the device-resident y / ac
int phb_inc_elmgmr(phb200_ctx *ctx, const phb200_incomp *ip) {
  const phb200_common &c = ctx->c;
  const size_t nshg = c.nshg;
  cudaStream_t s = ctx->stream;
  if (c.nelblb > 0 && ctx->bnd_deformable) {
    fprintf(stderr, "phb200: inc_elmgmr: deformable-wall boundary elements (iBCB bit 4, ideformwall) are not built\n");
    return 1;
  }
  if (ip->itau != 0 || ip->ipord != 1 || (ip->idiff != 0 && ip->idiff != 1) || (ip->iconvflow != 1 && ip->iconvflow != 2) ||
      (ip->matflg5 != 0 && ip->matflg5 != 1)) {
    fprintf(stderr, "phb200: inc_elmgmr: supported: itau=0, ipord=1, idiff 0|1, iconvflow 1|2, matflg(5,1) 0|1\n");
    return 1;
  }
  if (ip->lhs == 1 && (ctx->nnz_tot <= 0 || (!ctx->d_eloc && ctx->numel_tet > 0))) {
    fprintf(stderr, "phb200: inc_elmgmr: no CSR structure (call phb200_set_sparse first)\n");
    return 1;
  }
  PHB_TRY(upload_inc_tables(ctx));
  PHB_TRY(inc_alloc(ctx));
  IncPhys p;
  p.rho = ip->rho; p.rmu = ip->rmu;
  for (int i = 0; i < 3; i++) p.bf[i] = ip->bf[i];
  p.tmps = 1.0 - ip->flmpr;
  p.lhsFct = ip->alfi * ip->gami * ip->Delt;
  p.lhmFct = ip->almi * (1.0 - ip->flmpl);
  p.dts = ip->Dtgl * ip->dtsfct;
  p.ff = ip->taucfct / ip->dtsfct;
  p.iconvflow = ip->iconvflow; p.idiff = ip->idiff; p.matflg5 = ip->matflg5; p.lhs = ip->lhs;
  PHB_CHECK(cudaMemcpyToSymbolAsync(c_ip, &p, sizeof p, 0, cudaMemcpyHostToDevice, s));
  ctx->inc_idiff = ip->idiff;
  if (ip->idiff == 1) {  // elmgmr.f:44-86
    PHB_CHECK(cudaMemsetAsync(ctx->d_qres, 0, sizeof(double) * 12 * nshg, s));
    PHB_CHECK(cudaMemsetAsync(ctx->d_rmass, 0, sizeof(double) * nshg, s));
    PHB_TRY(launch_all(ctx, false, true));
    PHB_TRY(phb_qpbc(ctx));
  }
  PHB_CHECK(cudaMemsetAsync(ctx->d_res4, 0, sizeof(double) * 4 * nshg, s));
  if (ip->lhs == 1) {
    PHB_CHECK(cudaMemsetAsync(ctx->d_lhsK9, 0, sizeof(double) * 9 * (size_t)ctx->nnz_tot, s));
    PHB_CHECK(cudaMemsetAsync(ctx->d_lhsP4, 0, sizeof(double) * 4 * (size_t)ctx->nnz_tot, s));
  }
  PHB_TRY(launch_all(ctx, ip->lhs == 1, false));
  if (c.nelblb > 0) PHB_TRY(inc_boundary(ctx, ip));  // elmgmr.f:246-320
  PHB_TRY(phb_commu(ctx, ctx->d_res4, 4, 0));
  // bc3Res: bc3per (periodic sum, rows owned by another part zeroed) then the essential-BC projections
  PHB_TRY(phb_bc3per(ctx, ctx->d_res4, 4));
  PHB_TRY(phb_zero_slaves(ctx, ctx->d_res4, 4, 0));
  {
    KScope ks(ctx, KC_NODE);
    k_inc_bc3res<<<(unsigned)((nshg + 255) / 256), 256, 0, s>>>((int)nshg, ctx->d_iBC, ctx->d_BC, ctx->d_res4);
    PHB_CHECK(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------
// lesSparse.f products.  MODE bits: 1 = K p(:,1:3) into q(:,1:3); 2 = -G^T p(:,4) into q(:,1:3) (through tpos);
// 4 = G p(:,1:3) into the scalar row; 8 = + C p(:,4).
// Half a warp per CSR row (about 15 entries per row on tet meshes), lane = entry: each lane streams its entry's
// kLhs (72 B) and pLhs (32 B, two 16-byte loads), fetches the transposed entry's pLhs through tpos and gathers p from
// 32-byte node records (p is packed first: one L2 sector per gathered column instead of four); the four row sums
// are reduced with shuffles inside the half warp.  (A 16-lanes-per-entry variant, every load a contiguous run,
// was measured 3x slower: too few bytes in flight per warp.)
__global__ void k_les_pack(int nshg, int ncol, int scalar_only, const double *__restrict__ p, double *__restrict__ p4) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * nshg) return;
  const int i = t >> 2, c = t & 3;
  double v = 0.0;
  if (scalar_only) { if (c == 3) v = p[i]; }
  else if (c < ncol) v = p[(size_t)nshg * c + i];
  p4[t] = v;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_les_ap(int nshg, const int *__restrict__ colm, const int *__restrict__ rowp,
                                                 const int *__restrict__ tpos, const double *__restrict__ lhsK,
                                                 const double *__restrict__ lhsP, const double *__restrict__ p4,
                                                 double *__restrict__ q, int qcol4) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, l16 = threadIdx.x & 15;
  const bool live = row < nshg;
  int k0 = 0, k1 = 0;
  if (live) { k0 = colm[row]; k1 = colm[row + 1]; }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int k = k0 + l16; k < k1; k += 16) {
    const int j = __ldg(rowp + k);
    const double2 pa = __ldg(reinterpret_cast<const double2 *>(p4 + (size_t)4 * j));
    const double2 pb = __ldg(reinterpret_cast<const double2 *>(p4 + (size_t)4 * j) + 1);
    const double p1 = pa.x, p2 = pa.y, p3 = pb.x, pp = pb.y;
    if (MODE & 1) {
      const double *K = lhsK + (size_t)9 * k;  // lesSparse.f:283-294: the three rows use kLhs entries 1,4,7 / 2,5,8 / 3,6,9
      s0 = s0 + __ldcs(K + 0) * p1 + __ldcs(K + 3) * p2 + __ldcs(K + 6) * p3;
      s1 = s1 + __ldcs(K + 1) * p1 + __ldcs(K + 4) * p2 + __ldcs(K + 7) * p3;
      s2 = s2 + __ldcs(K + 2) * p1 + __ldcs(K + 5) * p2 + __ldcs(K + 8) * p3;
    }
    if (MODE & 2) {
      const double2 *Pt = reinterpret_cast<const double2 *>(lhsP + (size_t)4 * __ldg(tpos + k));
      const double2 ta = __ldg(Pt), tb = __ldg(Pt + 1);
      s0 -= ta.x * pp;
      s1 -= ta.y * pp;
      s2 -= tb.x * pp;
    }
    if (MODE & (4 | 8)) {
      const double2 *P = reinterpret_cast<const double2 *>(lhsP + (size_t)4 * k);
      const double2 ga = __ldg(P), gb = __ldg(P + 1);
      if (MODE & 4) s3 = s3 + ga.x * p1 + ga.y * p2 + gb.x * p3;
      if (MODE & 8) s3 += gb.y * pp;
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  if (live && l16 == 0) {
    if (MODE & (1 | 2)) {
      q[row] = s0;
      q[(size_t)nshg + row] = s1;
      q[(size_t)2 * nshg + row] = s2;
    }
    if (MODE & (4 | 8)) q[(size_t)qcol4 * nshg + row] = s3;
  }
}

// kind 0 ApG, 1 ApKG, 2 ApNGt, 3 ApNGtC, 4 ApFull on device vectors (column-major (nshg,ncol) like the reference)
int phb_les_ap(phb200_ctx *ctx, int kind, const double *d_p, double *d_q) {
  if (!ctx->d_lhsK9) {
    fprintf(stderr, "phb200: les_ap: no incompressible LHS (call phb200_inc_elmgmr with lhs=1 first)\n");
    return 1;
  }
  const int nshg = ctx->c.nshg;
  cudaStream_t s = ctx->stream;
  static const int ncol[5] = {1, 4, 3, 4, 4};
  if (kind < 0 || kind > 4) { fprintf(stderr, "phb200: les_ap: kind %d\n", kind); return 1; }
  {
    KScope ks(ctx, KC_BLAS);
    k_les_pack<<<(4 * nshg + 255) / 256, 256, 0, s>>>(nshg, ncol[kind], kind == 0, d_p, ctx->d_lesp4);
    PHB_CHECK(cudaGetLastError());
  }
  const unsigned nb = (unsigned)(((size_t)nshg * 16 + 127) / 128);
  KScope ks(ctx, KC_AP);
#define LES(MODE, qc)                                                                                           \
  k_les_ap<MODE><<<nb, 128, 0, s>>>(nshg, ctx->d_colm, ctx->d_rowp, ctx->d_tpos, ctx->d_lhsK9, ctx->d_lhsP4, \
                                    ctx->d_lesp4, d_q, qc)
  switch (kind) {
    case 0: LES(2, 0); break;
    case 1: LES(1 | 2, 0); break;
    case 2: LES(4, 0); break;
    case 3: LES(4 | 8, 0); break;
    default: LES(1 | 2 | 4 | 8, 3); break;
  }
#undef LES
  PHB_CHECK(cudaGetLastError());
  return 0;
}

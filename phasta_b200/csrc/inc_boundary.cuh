// inc_boundary.cuh -- the boundary integral of the incompressible code on the device: AsBMFG + e3b + e3bvar
// (incompressible/asbmfg.f:1-68, e3b.f:1-262, e3bvar.f:1-230) with rigid walls (ideformwall = 0: vdot, rlKwall and
// the boundary xKebe are zero, so the bc3lhs / fillsparseI of elmgmr.f:303-313 add nothing), for linear tets (LCSYST
// 1), hexes (2), wedges with a triangular (3) or quadrilateral (4) boundary face.  Thread = boundary element; no
// shared memory or warp intrinsics, so tests/host_emul/ compiles this file for the host as well.
// Included by incomp.cu after it has defined c_ibp (IncBndPhys) and c_ibnd (BndTables[4], index lcsyst-1).
#pragma once
#include "bnd_pack.h"

// struct IncBndPhys { double rho, rmu; int iviscflux, iconvflow, itwmod; } and the two constants are defined by
// the including translation unit (incomp.cu; tests/host_emul/inc_bnd_host.cpp)

// aer[0..2] Force, aer[4 + 10*surf + k] flxID(k+1,surf); nsrflist(0:MAXSURF) (common.h:98-108)
template <int NSHL, int NSHLB, int LCSYST>
__global__ void __launch_bounds__(128) k_inc_asbmfg(int nb, int nshg, int numnp, const int *__restrict__ ienb,
                                                     const int *__restrict__ iBCB, const double *__restrict__ BCB,
                                                     const double *__restrict__ x, const double *__restrict__ y,
                                                     const int *__restrict__ nsrflist, double *__restrict__ res,
                                                     double *__restrict__ aer) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb) return;
  const BndTables &tb = c_ibnd[LCSYST - 1];
  // lnode (getbnodes, hierarchic.f:90-168) and the two edge ends of the normal: "curl into element for tets,
  // all others out" (e3bvar.f:100-119), 0-based
  int ln[NSHLB];
#pragma unroll
  for (int k = 0; k < NSHLB; k++) ln[k] = k;
  if (LCSYST == 4) { ln[1] = 3; ln[2] = 4; ln[3] = 1; }
  const int ipt2 = (LCSYST == 1) ? 1 : (LCSYST == 2) ? 3 : (LCSYST == 3) ? 2 : 1;
  const int ipt3 = (LCSYST == 1) ? 2 : (LCSYST == 2) ? 1 : (LCSYST == 3) ? 1 : 3;
  int nd[NSHL];
  double xl[NSHL][3], yl[NSHL][4];  // localy: {p, u1, u2, u3}
#pragma unroll
  for (int a = 0; a < NSHL; a++) {
    nd[a] = ienb[(size_t)a * nb + e];
#pragma unroll
    for (int i = 0; i < 3; i++) xl[a][i] = __ldg(x + (size_t)numnp * i + nd[a]);
    yl[a][0] = __ldg(y + (size_t)nshg * 3 + nd[a]);
    yl[a][1] = __ldg(y + nd[a]);
    yl[a][2] = __ldg(y + (size_t)nshg * 1 + nd[a]);
    yl[a][3] = __ldg(y + (size_t)nshg * 2 + nd[a]);
  }
  const int ibcb = iBCB[e], surf = abs(iBCB[nb + e]);
  const int listed = (surf <= 1000) ? nsrflist[surf] : 0;
  double v1[3], v2[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    v1[i] = xl[ipt2][i] - xl[0][i];
    v2[i] = xl[ipt3][i] - xl[0][i];
  }
  const double t1 = v1[1] * v2[2] - v2[1] * v1[2];
  const double t2 = v2[0] * v1[2] - v1[0] * v2[2];
  const double t3 = v1[0] * v2[1] - v2[0] * v1[1];
  const double tinv = 1.0 / sqrt(t1 * t1 + t2 * t2 + t3 * t3);
  const double bn[3] = {t1 * tinv, t2 * tinv, t3 * tinv};
  const double rmu = c_ibp.rmu, rho = c_ibp.rho;
  double rl[NSHLB][4];
#pragma unroll
  for (int n = 0; n < NSHLB; n++)
#pragma unroll
    for (int m = 0; m < 4; m++) rl[n][m] = 0.0;
  double frc[3] = {0, 0, 0}, flx[5] = {0, 0, 0, 0, 0};
  const int nq = tb.nq;
  for (int q = 0; q < nq; q++) {
    // e3bvar.f:129-142
    const double WdetJb = (LCSYST == 3) ? tb.Qwt[q] / (2.0 * tinv) : tb.Qwt[q] / (4.0 * tinv);
    double J[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) J[i][j] = 0.0;
#pragma unroll
    for (int a = 0; a < NSHL; a++)
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) J[i][j] += xl[a][i] * tb.dN[q][a][j];
    double d[3][3];
    d[0][0] = J[1][1] * J[2][2] - J[2][1] * J[1][2];
    d[0][1] = J[2][1] * J[0][2] - J[0][1] * J[2][2];
    d[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    const double dinv = 1.0 / (d[0][0] * J[0][0] + d[0][1] * J[1][0] + d[0][2] * J[2][0]);
    d[0][0] *= dinv;
    d[0][1] *= dinv;
    d[0][2] *= dinv;
    d[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * dinv;
    d[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * dinv;
    d[1][2] = (J[1][0] * J[0][2] - J[0][0] * J[1][2]) * dinv;
    d[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * dinv;
    d[2][1] = (J[2][0] * J[0][1] - J[0][0] * J[2][1]) * dinv;
    d[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * dinv;
    double Y[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < NSHLB; k++)
#pragma unroll
      for (int m = 0; m < 4; m++) Y[m] += tb.N[q][ln[k]] * yl[ln[k]][m];
    double gl[3][4];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int m = 0; m < 4; m++) gl[i][m] = 0.0;
#pragma unroll
    for (int a = 0; a < NSHL; a++)
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int m = 0; m < 4; m++) gl[i][m] += tb.dN[q][a][i] * yl[a][m];
    double gr[3][4];  // gr[j][m] = dY_m/dx_j (m = 1..3 velocity)
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int m = 1; m < 4; m++) gr[j][m] = d[0][j] * gl[0][m] + d[1][j] * gl[1][m] + d[2][j] * gl[2][m];
    double pres = Y[0];
    const double u1 = Y[1], u2 = Y[2], u3 = Y[3];
    const double *g1 = gr[0], *g2 = gr[1], *g3 = gr[2];
    double unm = bn[0] * u1 + bn[1] * u2 + bn[2] * u3;
    double tau1n = bn[0] * 2.0 * rmu * g1[1] + bn[1] * (rmu * (g2[1] + g1[2])) + bn[2] * (rmu * (g3[1] + g1[3]));
    double tau2n = bn[0] * (rmu * (g2[1] + g1[2])) + bn[1] * 2.0 * rmu * g2[2] + bn[2] * (rmu * (g3[2] + g2[3]));
    double tau3n = bn[0] * (rmu * (g3[1] + g1[3])) + bn[1] * (rmu * (g3[2] + g2[3])) + bn[2] * 2.0 * rmu * g3[3];
    const double tn = bn[0] * tau1n + bn[1] * tau2n + bn[2] * tau3n;  // e3bvar.f: the normal part goes to pres
    pres = pres - tn;
    tau1n = (tau1n - bn[0] * tn) * c_ibp.iviscflux;
    tau2n = (tau2n - bn[1] * tn) * c_ibp.iviscflux;
    tau3n = (tau3n - bn[2] * tn) * c_ibp.iviscflux;
    if (listed != 0) {  // flxID before the natural BCs replace the computed values (e3b.f:64-78)
      flx[0] += WdetJb;
      flx[1] -= WdetJb * unm;
      flx[2] -= (tau1n - bn[0] * pres) * WdetJb;
      flx[3] -= (tau2n - bn[1] * pres) * WdetJb;
      flx[4] -= (tau3n - bn[2] * pres) * WdetJb;
    }
    // natural BCs (e3b.f:80-112): shape(lnode(n)) * BCB(:,n,k)
    if (ibcb & 7) {
      double bv[5] = {0, 0, 0, 0, 0};
#pragma unroll
      for (int n = 0; n < NSHLB; n++)
#pragma unroll
        for (int k = 0; k < 5; k++) bv[k] += tb.N[q][ln[n]] * __ldg(BCB + (size_t)(k * NSHLB + n) * nb + e);
      if (ibcb & 1) unm = bv[0];
      if (ibcb & 2) pres = bv[1];
      if (ibcb & 4) { tau1n = bv[2]; tau2n = bv[3]; tau3n = bv[4]; }
    }
    double rNa[4];
    rNa[0] = -WdetJb * (tau1n - bn[0] * pres);
    rNa[1] = -WdetJb * (tau2n - bn[1] * pres);
    rNa[2] = -WdetJb * (tau3n - bn[2] * pres);
    rNa[3] = WdetJb * unm;
    if (c_ibp.iconvflow == 1) {  // conservative form: convective boundary integral (e3b.f:176-183)
      const double rou = rho * unm;
      rNa[0] += WdetJb * rou * u1;
      rNa[1] += WdetJb * rou * u2;
      rNa[2] += WdetJb * rou * u3;
    }
#pragma unroll
    for (int n = 0; n < NSHLB; n++)
#pragma unroll
      for (int m = 0; m < 4; m++) rl[n][m] -= tb.N[q][ln[n]] * rNa[m];
    if (listed == 1) {  // e3b.f:224-240
      frc[0] += (tau1n - bn[0] * pres) * WdetJb;
      frc[1] += (tau2n - bn[1] * pres) * WdetJb;
      frc[2] += (tau3n - bn[2] * pres) * WdetJb;
    }
  }
  // local (res, rl, ienb, nflow, 'scatter'): res(:,1:3) momentum, res(:,4) continuity
#pragma unroll
  for (int n = 0; n < NSHLB; n++)
#pragma unroll
    for (int m = 0; m < 4; m++) atomicAdd(res + (size_t)nshg * m + nd[ln[n]], rl[n][m]);
  if (listed != 0)
    for (int k = 0; k < 5; k++) atomicAdd(aer + 4 + 10 * surf + k, flx[k]);
  if (listed == 1 && (c_ibp.itwmod == 1 || c_ibp.itwmod == -1))
    for (int k = 0; k < 3; k++) atomicAdd(aer + k, -frc[k]);
}

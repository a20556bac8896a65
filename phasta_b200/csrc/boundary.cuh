// boundary.cuh -- device code shared by assembly.cu and its host emulation (tests/host_emul/bnd_host.cpp, which
// compiles this file with g++ behind a small shim: thread-per-element kernels without shared memory or warp
// intrinsics run unchanged as plain loops): the physical parameters in constant memory, the localy gather, the
// diffusivities, and the boundary-flux kernel for hexes and wedges.
#pragma once
#include "bnd_pack.h"

// c_ph (PhysParams) and c_bnd (BndTables[3]) are defined by the including translation unit before this point

// getDiff (compressible/getdiff.f:127-171), DNS
__device__ __forceinline__ void diffusivities(double T, double cp, double &mu, double &lam, double &con) {
  const double pt66 = 0.6666666666666666666666666666667;
  if (c_ph.matflg2 == 0)
    mu = c_ph.mu0;
  else
    mu = c_ph.mu0 * (T / c_ph.Tref) * sqrt(T / c_ph.Tref) * (c_ph.Tref + c_ph.Ssuth) / (T + c_ph.Ssuth);
  lam = (c_ph.matflg3 == 0) ? (-pt66 * mu) : ((c_ph.dat131 - pt66) * mu);
  con = mu * cp / c_ph.pr;
}

// localy (common/localy.f:47-72): global {u,v,w,p,T} -> local {p,u,v,w,T}
__device__ __forceinline__ void gather_y(const double *__restrict__ y, int nshg, int node, double yl[5]) {
  yl[0] = __ldg(y + (size_t)nshg * 3 + node);
  yl[1] = __ldg(y + node);
  yl[2] = __ldg(y + (size_t)nshg * 1 + node);
  yl[3] = __ldg(y + (size_t)nshg * 2 + node);
  yl[4] = __ldg(y + (size_t)nshg * 4 + node);
}

// ---------------------------------------------------------------------------
// The same for hexes (LCSYST 2, quadrilateral face), wedges with a triangular (3) or quadrilateral (4) boundary
// face: lnode of getbnodes (hierarchic.f:119-168), the per-topology normals and WdetJb of e3bvar.f:139-176,
// grad Y through the volume metric at the face points (e3bvar.f:182-262).  Thread = boundary element.
// ---------------------------------------------------------------------------
template <int NSHL, int NSHLB, int LCSYST>
__global__ void __launch_bounds__(128) k_asbmfg_gen(int nb, int nshg, int numnp, const int *__restrict__ ienb,
                                                     const int *__restrict__ iBCB, const double *__restrict__ BCB,
                                                     const double *__restrict__ x, const double *__restrict__ y,
                                                     double *__restrict__ res, double *__restrict__ aer,
                                                     int do_force) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb) return;
  const BndTables &tb = c_bnd[LCSYST - 2];
  // lnode (0-based): the element nodes that lie on the boundary face
  int ln[NSHLB];
#pragma unroll
  for (int k = 0; k < NSHLB; k++) ln[k] = k;
  if (LCSYST == 4) { ln[1] = 3; ln[2] = 4; ln[3] = 1; }
  int nd[NSHL];
  double xl[NSHL][3], yl[NSHL][5];
#pragma unroll
  for (int a = 0; a < NSHL; a++) {
    nd[a] = ienb[(size_t)a * nb + e];
#pragma unroll
    for (int i = 0; i < 3; i++) xl[a][i] = __ldg(x + (size_t)numnp * i + nd[a]);
    gather_y(y, nshg, nd[a], yl[a]);
  }
  const int ibcb = iBCB[e], surf = abs(iBCB[nb + e]);
  double v1[3], v2[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    v1[i] = xl[1][i] - xl[0][i];
    v2[i] = xl[2][i] - xl[0][i];
  }
  double rl[NSHLB][5];
#pragma unroll
  for (int n = 0; n < NSHLB; n++)
#pragma unroll
    for (int m = 0; m < 5; m++) rl[n][m] = 0.0;
  double frc[4] = {0, 0, 0, 0}, flx[5] = {0, 0, 0, 0, 0};
  const int nq = tb.nq;
  for (int q = 0; q < nq; q++) {
    // deformation gradient dxdxib(i,j) = sum_n xlb(n,i) shglb(j,n) (e3bvar.f:126-136)
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
    double t1, t2, t3;
    if (LCSYST == 4) {  // e3bvar.f:139-146
      t1 = J[1][0] * J[2][2] - J[1][2] * J[2][0];
      t2 = J[2][0] * J[0][2] - J[2][2] * J[0][0];
      t3 = J[0][0] * J[1][2] - J[0][2] * J[1][0];
    } else {            // e3bvar.f:152-155
      t1 = -v1[1] * v2[2] + v2[1] * v1[2];
      t2 = -v2[0] * v1[2] + v1[0] * v2[2];
      t3 = -v1[0] * v2[1] + v2[0] * v1[1];
    }
    const double tinv = 1.0 / sqrt(t1 * t1 + t2 * t2 + t3 * t3);
    const double bn[3] = {t1 * tinv, t2 * tinv, t3 * tinv};
    double WdetJb;      // e3bvar.f:163-176
    if (LCSYST == 3) WdetJb = (1.0 - tb.Qwt[q]) / (4.0 * tinv);
    else if (LCSYST == 4) WdetJb = tb.Qwt[q] / tinv;
    else WdetJb = tb.Qwt[q] / (4.0 * tinv);
    // inverse of the deformation gradient, d[i][j] = dxidxb(i+1,j+1) (e3bvar.f:186-212)
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
    // state on the face nodes (e3bvar.f:94-103), local and global grad Y (e3bvar.f:216-262)
    double Y[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < NSHLB; k++)
#pragma unroll
      for (int m = 0; m < 5; m++) Y[m] += tb.N[q][ln[k]] * yl[ln[k]][m];
    double gl[3][5];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int m = 0; m < 5; m++) gl[i][m] = 0.0;
#pragma unroll
    for (int a = 0; a < NSHL; a++)
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int m = 0; m < 5; m++) gl[i][m] += tb.dN[q][a][i] * yl[a][m];
    double gr[3][5];  // gr[j][m] = dY_m/dx_j
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int m = 0; m < 5; m++) gr[j][m] = d[0][j] * gl[0][m] + d[1][j] * gl[1][m] + d[2][j] * gl[2][m];
    const double pres = Y[0], u1 = Y[1], u2 = Y[2], u3 = Y[3], T = Y[4];
    const double rk = 0.5 * (u1 * u1 + u2 * u2 + u3 * u3);
    const double rho = pres / (c_ph.Rgas * T);
    const double ei = T * (c_ph.Rgas / c_ph.gamma1);
    const double cp = c_ph.Rgas * c_ph.gamma / c_ph.gamma1;
    // natural BC values interpolated on the face (e3bvar.f:330-356): shpb(lnode(n)) * BCB(:,n,k)
    double bv[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int n = 0; n < NSHLB; n++)
#pragma unroll
      for (int k = 0; k < 6; k++) bv[k] += tb.N[q][ln[n]] * __ldg(BCB + (size_t)(k * NSHLB + n) * nb + e);
    double rou, un, pb = bv[1];
    if (!(ibcb & 1)) {
      un = bn[0] * u1 + bn[1] * u2 + bn[2] * u3;
      rou = rho * un;
    } else {
      rou = bv[0];
      un = rou / rho;
    }
    if (!(ibcb & 2)) pb = pres;
    double F[5];
    F[0] = rou;
    F[1] = rou * u1 + bn[0] * pb;
    F[2] = rou * u2 + bn[1] * pb;
    F[3] = rou * u3 + bn[2] * pb;
    F[4] = rou * (ei + rk) + un * pb;
    double mu, lam, con;
    diffusivities(T, cp, mu, lam, con);
    const double l2m = lam + 2.0 * mu;
    const double *g1 = gr[0], *g2 = gr[1], *g3 = gr[2];
    const double tau1n = bn[0] * (l2m * g1[1] + lam * g2[2] + lam * g3[3]) + bn[1] * (mu * (g2[1] + g1[2])) +
                         bn[2] * (mu * (g3[1] + g1[3]));
    const double tau2n = bn[0] * (mu * (g2[1] + g1[2])) + bn[1] * (lam * g1[1] + l2m * g2[2] + lam * g3[3]) +
                         bn[2] * (mu * (g3[2] + g2[3]));
    const double tau3n = bn[0] * (mu * (g3[1] + g1[3])) + bn[1] * (mu * (g3[2] + g2[3])) +
                         bn[2] * (lam * g1[1] + lam * g2[2] + l2m * g3[3]);
    double Fv2 = bv[2], Fv3 = bv[3], Fv4 = bv[4], Fh5 = bv[5];
    if (!(ibcb & 4)) { Fv2 = tau1n; Fv3 = tau2n; Fv4 = tau3n; }
    const double Fv5 = u1 * Fv2 + u2 * Fv3 + u3 * Fv4;
    const double heat = -con * (bn[0] * g1[4] + bn[1] * g2[4] + bn[2] * g3[4]);
    if (!(ibcb & 8)) Fh5 = heat;
    F[1] -= Fv2; F[2] -= Fv3; F[3] -= Fv4;
    F[4] = F[4] - Fv5 + Fh5;
#pragma unroll
    for (int n = 0; n < NSHLB; n++) {
      const double wn = WdetJb * tb.N[q][ln[n]];
#pragma unroll
      for (int m = 0; m < 5; m++) rl[n][m] += wn * F[m];
    }
    // flxID (e3b.f:305-321) and aerodynamic forces (e3b.f:325-345)
    flx[0] += WdetJb;
    flx[1] -= WdetJb * rou;
    flx[2] -= (tau1n - bn[0] * pres) * WdetJb;
    flx[3] -= (tau2n - bn[1] * pres) * WdetJb;
    flx[4] -= (tau3n - bn[2] * pres) * WdetJb;
    if (!(ibcb & 1)) {
      frc[0] += (pres * bn[0] - tau1n) * WdetJb;
      frc[1] += (pres * bn[1] - tau2n) * WdetJb;
      frc[2] += (pres * bn[2] - tau3n) * WdetJb;
      frc[3] += -heat * WdetJb;
    }
  }
#pragma unroll
  for (int n = 0; n < NSHLB; n++)
#pragma unroll
    for (int m = 0; m < 5; m++) atomicAdd(res + (size_t)nshg * m + nd[ln[n]], rl[n][m]);
  if (surf != 0 && surf <= 1000)
    for (int k = 0; k < 5; k++) atomicAdd(aer + 4 + 10 * surf + k, flx[k]);
  if (do_force)
    for (int k = 0; k < 4; k++) atomicAdd(aer + k, frc[k]);
}


/*
 * oracle_elem.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 *
 * Restatement of the element-level routines of PHASTA's compressible
 * interior assembly: AsIGMR -> e3 -> e3ivar/getthm/getDiff/e3metric/e3mtrx/
 * e3conv/e3visc/e3LS/e3tau/e3massr/e3juel/e3massl/e3wmlt, and AsIq -> e3q.
 * Restricted to the settings of BASELINE.json's configs: ipress=0 (ideal
 * gas), itau in {0}, iDC=0, Navier=1, DNS, no level set, no body force,
 * ipord=1.  Index conventions inside this file are the Fortran 1-based ones
 * (arrays are over-allocated by one) so formulas read like the reference.
 */
#include "phasta_oracle.h"
#include "oracle_internal.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- per-quadrature-point work state of one element (the automatic arrays
 * of e3, compressible/e3.f:61-95, for a single iel) ---- */
typedef struct qpstate {
  int nshl, nenl, lcsyst, intp, ngauss;
  double shape[ORC_MAXSH + 1], shdrv[4][ORC_MAXSH + 1];
  double dui[6], aci[6], g1yi[6], g2yi[6], g3yi[6];
  double shg[ORC_MAXSH + 1][4], dxidx[4][4], WdetJ;
  double rho, pres, T, ei, h, alfap, betaT, cp, cv, rk, u1, u2, u3;
  double divqi[6], rmu, rlm, rlm2mu, con;
  double A0[6][6], A1[6][6], A2[6][6], A3[6][6];
  double rLyi[6], rLymi[6], tau[6];
  double ri[21], rmi[21], stiff[16][16];
  /* discontinuity capturing (iDC /= 0): e3mtrx.f:232-298, e3tau.f:177-247, e3dc.f */
  double A0DC[5], A0inv[16], dVdY[16], giju[7], rTLS, raLS, DC;
} qpstate;

/* getthm, ipress=0 (compressible/getthm.f:111,148) ithm=6 */
static void getthm6(const orc_common *c, double pres, double T, double *rho,
                    double *ei) {
  *rho = pres / (c->Rgas * T);
  *ei = T * (c->Rgas / c->gamma1);
}
/* getthm ithm=7 (compressible/getthm.f:111,148,163-169) */
static void getthm7(const orc_common *c, double pres, double T, double *rho,
                    double *ei, double *h, double *cv, double *cp,
                    double *alfap, double *betaT) {
  getthm6(c, pres, T, rho, ei);
  *h = T * (c->Rgas * c->gamma / c->gamma1);
  *cv = c->Rgas / c->gamma1;
  *cp = c->Rgas * c->gamma / c->gamma1;
  *alfap = 1.0 / T;
  *betaT = 1.0 / pres;
}

/* getDiff (compressible/getdiff.f:54-269), DNS branch: constant viscosity
 * (matflg(2,1)=0, :127) or Sutherland (:156-157); rlm :161-165; con :171;
 * xmut = 0 (:180); final combination :263-266 */
static void getdiff(const orc_common *c, double T, double cp, double *rmu,
                    double *rlm, double *rlm2mu, double *con) {
  const double pt66 = 0.6666666666666666666666666666667;
  double mu;
  if (c->matflg2 == 0) {
    mu = c->datmat121;
  } else {
    mu = c->datmat121 * (T / c->datmat221) * sqrt(T / c->datmat221) *
         (c->datmat221 + c->datmat321) / (T + c->datmat321);
  }
  double lm;
  if (c->matflg3 == 0)
    lm = -pt66 * mu;
  else
    lm = (c->datmat131 - pt66) * mu;
  double k = mu * cp / c->pr;
  double xmut = 0.0;
  lm = lm - pt66 * xmut;
  mu = mu + xmut;
  *rmu = mu;
  *rlm = lm;
  *rlm2mu = lm + 2.0 * mu;
  *con = k + xmut * cp / c->pr;
}

/* e3metric (common/e3metric.f:8-80); xl(nenl,3) of this element */
static void e3metric(const orc_common *c, qpstate *s, double xl[][4]) {
  double dxdxi[4][4];
  memset(dxdxi, 0, sizeof dxdxi);
  for (int n = 1; n <= s->nenl; n++)
    for (int i = 1; i <= 3; i++)
      for (int j = 1; j <= 3; j++) dxdxi[i][j] += xl[n][i] * s->shdrv[j][n];
  double(*d)[4] = s->dxidx;
  d[1][1] = dxdxi[2][2] * dxdxi[3][3] - dxdxi[3][2] * dxdxi[2][3];
  d[1][2] = dxdxi[3][2] * dxdxi[1][3] - dxdxi[1][2] * dxdxi[3][3];
  d[1][3] = dxdxi[1][2] * dxdxi[2][3] - dxdxi[1][3] * dxdxi[2][2];
  double tmp = 1.0 / (d[1][1] * dxdxi[1][1] + d[1][2] * dxdxi[2][1] +
                      d[1][3] * dxdxi[3][1]);
  d[1][1] *= tmp;
  d[1][2] *= tmp;
  d[1][3] *= tmp;
  d[2][1] = (dxdxi[2][3] * dxdxi[3][1] - dxdxi[2][1] * dxdxi[3][3]) * tmp;
  d[2][2] = (dxdxi[1][1] * dxdxi[3][3] - dxdxi[3][1] * dxdxi[1][3]) * tmp;
  d[2][3] = (dxdxi[2][1] * dxdxi[1][3] - dxdxi[1][1] * dxdxi[2][3]) * tmp;
  d[3][1] = (dxdxi[2][1] * dxdxi[3][2] - dxdxi[2][2] * dxdxi[3][1]) * tmp;
  d[3][2] = (dxdxi[3][1] * dxdxi[1][2] - dxdxi[1][1] * dxdxi[3][2]) * tmp;
  d[3][3] = (dxdxi[1][1] * dxdxi[2][2] - dxdxi[1][2] * dxdxi[2][1]) * tmp;
  s->WdetJ = QWT(c, s->lcsyst, s->intp) / tmp;
  for (int n = 1; n <= s->nshl; n++)
    for (int i = 1; i <= 3; i++)
      s->shg[n][i] = s->shdrv[1][n] * d[1][i] + s->shdrv[2][n] * d[2][i] +
                     s->shdrv[3][n] * d[3][i];
}

/* e3ivar (compressible/e3ivar.f:1-489).  ycl/acl(nshl,5) element-local in
 * {p,u1,u2,u3,T} order, ql(nshl,idflx). */
static void e3ivar(const orc_common *c, qpstate *s, double yl[][6], double ycl[][6],
                   double acl[][6], double xl[][4], double ql[][13]) {
  int nshl = s->nshl;
  /* dui and the gradients come from yl, the point state from ycl; the two
   * alias in AsIGMR / AsIMFG and differ in AsIRes (perturbed vs base state) */
  for (int m = 1; m <= 5; m++) s->dui[m] = 0.0;
  for (int n = 1; n <= nshl; n++)
    for (int m = 1; m <= 5; m++) s->dui[m] += s->shape[n] * yl[n][m];
  /* conservative variables (:164-185) */
  s->rk = 0.5 * (s->dui[2] * s->dui[2] + s->dui[3] * s->dui[3] +
                 s->dui[4] * s->dui[4]);
  getthm6(c, s->dui[1], s->dui[5], &s->rho, &s->ei);
  s->dui[1] = s->rho;
  s->dui[2] = s->rho * s->dui[2];
  s->dui[3] = s->rho * s->dui[3];
  s->dui[4] = s->rho * s->dui[4];
  s->dui[5] = s->rho * (s->ei + s->rk);
  /* primitive variables at the point (:186-203) */
  s->pres = s->u1 = s->u2 = s->u3 = s->T = 0.0;
  for (int n = 1; n <= nshl; n++) {
    s->pres += s->shape[n] * ycl[n][1];
    s->u1 += s->shape[n] * ycl[n][2];
    s->u2 += s->shape[n] * ycl[n][3];
    s->u3 += s->shape[n] * ycl[n][4];
    s->T += s->shape[n] * ycl[n][5];
  }
  /* acceleration (:222-231) */
  for (int m = 1; m <= 5; m++) s->aci[m] = 0.0;
  for (int n = 1; n <= nshl; n++)
    for (int m = 1; m <= 5; m++) s->aci[m] += s->shape[n] * acl[n][m];
  /* thermodynamics (:236-252) */
  s->rk = 0.5 * (s->u1 * s->u1 + s->u2 * s->u2 + s->u3 * s->u3);
  getthm7(c, s->pres, s->T, &s->rho, &s->ei, &s->h, &s->cv, &s->cp, &s->alfap,
          &s->betaT);
  getdiff(c, s->T, s->cp, &s->rmu, &s->rlm, &s->rlm2mu, &s->con);
  e3metric(c, s, xl);
  /* global gradients (:259-356) */
  for (int m = 1; m <= 5; m++) s->g1yi[m] = s->g2yi[m] = s->g3yi[m] = 0.0;
  for (int n = 1; n <= nshl; n++)
    for (int m = 1; m <= 5; m++) {
      s->g1yi[m] += s->shg[n][1] * yl[n][m];
      s->g2yi[m] += s->shg[n][2] * yl[n][m];
      s->g3yi[m] += s->shg[n][3] * yl[n][m];
    }
  /* div q (:358-395) */
  for (int m = 1; m <= 5; m++) s->divqi[m] = 0.0;
  if (c->idiff >= 1 && (c->ires == 3 || c->ires == 1)) {
    for (int n = 1; n <= nshl; n++) {
      s->divqi[1] += s->shg[n][1] * ql[n][1] + s->shg[n][2] * ql[n][5] +
                     s->shg[n][3] * ql[n][9];
      s->divqi[2] += s->shg[n][1] * ql[n][2] + s->shg[n][2] * ql[n][6] +
                     s->shg[n][3] * ql[n][10];
      s->divqi[3] += s->shg[n][1] * ql[n][3] + s->shg[n][2] * ql[n][7] +
                     s->shg[n][3] * ql[n][11];
      s->divqi[4] += s->shg[n][1] * ql[n][4] + s->shg[n][2] * ql[n][8] +
                     s->shg[n][3] * ql[n][12];
    }
  }
}

/* the iDC /= 0 tail of e3mtrx (compressible/e3mtrx.f:232-298): A0DC, A0^-1 (15 symmetric entries), dV/dY */
static void e3mtrx_dc(qpstate *s) {
  double rho = s->rho, T = s->T, u1 = s->u1, u2 = s->u2, u3 = s->u3, rk = s->rk, h = s->h, cp = s->cp;
  double alfap = s->alfap, betaT = s->betaT;
  double s1 = 1.0 / (rho * rho * betaT * T);
  double cv = cp - (alfap * alfap * T / rho / betaT);
  s->A0DC[1] = (rho * betaT) * (rho * betaT) * s1;
  s->A0DC[2] = -rho * alfap * rho * betaT * s1;
  s->A0DC[3] = rho / T;
  s->A0DC[4] = (-rho * alfap) * (-rho * alfap) * s1 + (rho * cv / (T * T));
  double fact1 = 1.0 / (rho * cv * (T * T));
  double d = alfap * T / rho / betaT;
  double e1bar = h - rk, e2bar = e1bar - d, e3bar = e2bar - cv * T;
  double e5bar = e1bar * e1bar - 2 * e1bar * d + 2 * rk * cv * T + cp * T / rho / betaT;
  double c1bar = u1 * u1 + cv * T, c2bar = u2 * u2 + cv * T, c3bar = u3 * u3 + cv * T;
  double u12 = u1 * u2, u31 = u3 * u1, u23 = u2 * u3;
  double *Ai = s->A0inv;
  Ai[1] = e5bar * fact1;
  Ai[2] = c1bar * fact1;
  Ai[3] = c2bar * fact1;
  Ai[4] = c3bar * fact1;
  Ai[5] = 1 * fact1;
  Ai[6] = u1 * e3bar * fact1;
  Ai[7] = u2 * e3bar * fact1;
  Ai[8] = u3 * e3bar * fact1;
  Ai[9] = -e2bar * fact1;
  Ai[10] = u12 * fact1;
  Ai[11] = u31 * fact1;
  Ai[12] = -u1 * fact1;
  Ai[13] = u23 * fact1;
  Ai[14] = -u2 * fact1;
  Ai[15] = -u3 * fact1;
  fact1 = 1 / T;
  double fact2 = fact1 / T;
  double *V = s->dVdY;
  V[1] = fact1 / rho;
  V[2] = -fact1 * u1;
  V[3] = fact1;
  V[4] = -fact1 * u2;
  V[5] = 0.0;
  V[6] = fact1;
  V[7] = -fact1 * u3;
  V[8] = 0.0;
  V[9] = 0.0;
  V[10] = fact1;
  V[11] = -(h - rk) * fact2;
  V[12] = -fact2 * u1;
  V[13] = -fact2 * u2;
  V[14] = -fact2 * u3;
  V[15] = fact2;
}

/* e3mtrx (compressible/e3mtrx.f:75-227) */
static void e3mtrx(qpstate *s) {
  double rho = s->rho, u1 = s->u1, u2 = s->u2, u3 = s->u3;
  memset(s->A0, 0, sizeof s->A0);
  memset(s->A1, 0, sizeof s->A1);
  memset(s->A2, 0, sizeof s->A2);
  memset(s->A3, 0, sizeof s->A3);
  double drdp = rho * s->betaT;
  double drdT = -rho * s->alfap;
  double(*A0)[6] = s->A0, (*A1)[6] = s->A1, (*A2)[6] = s->A2,
  (*A3)[6] = s->A3;
  A0[5][1] = drdp * (s->h + s->rk) - s->alfap * s->T;
  double e2p = A0[5][1] + 1.0;
  double e3p = rho * (s->h + s->rk);
  double e4p = drdT * (s->h + s->rk) + rho * s->cp;
  A0[1][1] = drdp;
  A0[1][5] = drdT;
  A0[2][1] = drdp * u1;
  A0[2][2] = rho;
  A0[2][5] = drdT * u1;
  A0[3][1] = drdp * u2;
  A0[3][3] = rho;
  A0[3][5] = drdT * u2;
  A0[4][1] = drdp * u3;
  A0[4][4] = rho;
  A0[4][5] = drdT * u3;
  A0[5][2] = rho * u1;
  A0[5][3] = rho * u2;
  A0[5][4] = rho * u3;
  A0[5][5] = e4p;

  A1[1][1] = drdp * u1;
  A1[1][2] = rho;
  A1[1][5] = drdT * u1;
  A1[2][1] = drdp * u1 * u1 + 1;
  A1[2][2] = 2.0 * rho * u1;
  A1[2][5] = drdT * u1 * u1;
  A1[3][1] = drdp * u1 * u2;
  A1[3][2] = rho * u2;
  A1[3][3] = rho * u1;
  A1[3][5] = drdT * u1 * u2;
  A1[4][1] = drdp * u1 * u3;
  A1[4][2] = rho * u3;
  A1[4][4] = rho * u1;
  A1[4][5] = drdT * u1 * u3;
  A1[5][1] = u1 * e2p;
  A1[5][2] = e3p + rho * u1 * u1;
  A1[5][3] = rho * u1 * u2;
  A1[5][4] = rho * u1 * u3;
  A1[5][5] = u1 * e4p;

  A2[1][1] = drdp * u2;
  A2[1][3] = rho;
  A2[1][5] = drdT * u2;
  A2[2][1] = drdp * u1 * u2;
  A2[2][2] = rho * u2;
  A2[2][3] = rho * u1;
  A2[2][5] = drdT * u1 * u2;
  A2[3][1] = drdp * u2 * u2 + 1;
  A2[3][3] = 2.0 * rho * u2;
  A2[3][5] = drdT * u2 * u2;
  A2[4][1] = drdp * u2 * u3;
  A2[4][3] = rho * u3;
  A2[4][4] = rho * u2;
  A2[4][5] = drdT * u2 * u3;
  A2[5][1] = u2 * e2p;
  A2[5][2] = rho * u1 * u2;
  A2[5][3] = e3p + rho * u2 * u2;
  A2[5][4] = rho * u2 * u3;
  A2[5][5] = u2 * e4p;

  A3[1][1] = drdp * u3;
  A3[1][4] = rho;
  A3[1][5] = drdT * u3;
  A3[2][1] = drdp * u1 * u3;
  A3[2][2] = rho * u3;
  A3[2][4] = rho * u1;
  A3[2][5] = drdT * u1 * u3;
  A3[3][1] = drdp * u3 * u2;
  A3[3][3] = rho * u3;
  A3[3][4] = rho * u2;
  A3[3][5] = drdT * u3 * u2;
  A3[4][1] = drdp * u3 * u3 + 1;
  A3[4][4] = 2.0 * rho * u3;
  A3[4][5] = drdT * u3 * u3;
  A3[5][1] = u3 * e2p;
  A3[5][2] = rho * u1 * u3;
  A3[5][3] = rho * u2 * u3;
  A3[5][4] = e3p + rho * u3 * u3;
  A3[5][5] = u3 * e4p;
}

/* EGe(r,c): element matrix of one element inside the (numel,nedof,nedof)
 * array: r,c 1-based */
#define EGE(r, c) EG[(size_t)eg_stride * (((r)-1) + (size_t)nedof * ((c)-1))]

/* e3conv (compressible/e3conv.f:68-352) */
static void e3conv(const orc_common *c, qpstate *s, double *EG,
                   size_t eg_stride, int nedof) {
  double rho = s->rho, u1 = s->u1, u2 = s->u2, u3 = s->u3, pres = s->pres;
  double *ri = s->ri;
  if (c->ires == 1 || c->ires == 3) {
    ri[1] = (-u1) * rho;
    ri[2] = (-u1) * rho * u1 - pres;
    ri[3] = (-u1) * rho * u2;
    ri[4] = (-u1) * rho * u3;
    ri[5] = (-u1) * rho * (s->ei + s->rk) - u1 * pres;
    ri[6] = (-u2) * rho;
    ri[7] = (-u2) * rho * u1;
    ri[8] = (-u2) * rho * u2 - pres;
    ri[9] = (-u2) * rho * u3;
    ri[10] = (-u2) * rho * (s->ei + s->rk) - u2 * pres;
    ri[11] = (-u3) * rho;
    ri[12] = (-u3) * rho * u1;
    ri[13] = (-u3) * rho * u2;
    ri[14] = (-u3) * rho * u3 - pres;
    ri[15] = (-u3) * rho * (s->ei + s->rk) - u3 * pres;
  }
  /* rLyi = A_i Y,i (:100-179); structural zeros contribute exactly 0 */
  for (int m = 1; m <= 5; m++) {
    double acc = 0.0;
    for (int n = 1; n <= 5; n++) acc += s->A1[m][n] * s->g1yi[n];
    for (int n = 1; n <= 5; n++) acc += s->A2[m][n] * s->g2yi[n];
    for (int n = 1; n <= 5; n++) acc += s->A3[m][n] * s->g3yi[n];
    s->rLyi[m] = acc;
  }
  if (c->ires == 2 || c->ires == 3)
    for (int m = 1; m <= 5; m++) s->rmi[15 + m] = s->rLyi[m];
  if (c->lhs == 1) {
    double AiNbi[6][6];
    for (int j = 1; j <= s->nshl; j++) {
      int j0 = 5 * (j - 1);
      double fact1 = s->WdetJ * s->shg[j][1];
      double fact2 = s->WdetJ * s->shg[j][2];
      double fact3 = s->WdetJ * s->shg[j][3];
      for (int m = 1; m <= 5; m++)
        for (int n = 1; n <= 5; n++)
          AiNbi[m][n] =
              fact1 * s->A1[m][n] + fact2 * s->A2[m][n] + fact3 * s->A3[m][n];
      for (int i = 1; i <= s->nshl; i++) {
        int i0 = 5 * (i - 1);
        for (int jdof = 1; jdof <= 5; jdof++)
          for (int m = 1; m <= 5; m++)
            EGE(i0 + m, j0 + jdof) += s->shape[i] * AiNbi[m][jdof];
      }
    }
  }
}

/* e3visc (compressible/e3visc.f:57-357), itau<10, rlsli=0 */
/* K_ij of e3visc.f:69-139 into a 15x15 array of 5x5 blocks (sparse pattern) */
static void visc_stiff(const qpstate *s, double (*st)[16]) {
  double rmu = s->rmu, rlm = s->rlm, rlm2mu = s->rlm2mu, con = s->con;
  double u1 = s->u1, u2 = s->u2, u3 = s->u3;
  st[2][2] = rlm2mu;
  st[3][3] = rmu;
  st[4][4] = rmu;
  st[5][2] = rlm2mu * u1;
  st[5][3] = rmu * u2;
  st[5][4] = rmu * u3;
  st[5][5] = con;
  st[2][8] = rlm;
  st[3][7] = rmu;
  st[5][7] = rmu * u2;
  st[5][8] = rlm * u1;
  st[2][14] = rlm;
  st[4][12] = rmu;
  st[5][12] = rmu * u3;
  st[5][14] = rlm * u1;
  st[7][3] = rmu;
  st[8][2] = rlm;
  st[10][2] = rlm * u2;
  st[10][3] = rmu * u1;
  st[7][7] = rmu;
  st[8][8] = rlm2mu;
  st[9][9] = rmu;
  st[10][7] = rmu * u1;
  st[10][8] = rlm2mu * u2;
  st[10][9] = rmu * u3;
  st[10][10] = con;
  st[8][14] = rlm;
  st[9][13] = rmu;
  st[10][13] = rmu * u3;
  st[10][14] = rlm * u2;
  st[12][4] = rmu;
  st[14][2] = rlm;
  st[15][2] = rlm * u3;
  st[15][4] = rmu * u1;
  st[13][9] = rmu;
  st[14][8] = rlm;
  st[15][8] = rlm * u3;
  st[15][9] = rmu * u2;
  st[12][12] = rmu;
  st[13][13] = rmu;
  st[14][14] = rlm2mu;
  st[15][12] = rmu * u1;
  st[15][13] = rmu * u2;
  st[15][14] = rlm2mu * u3;
  st[15][15] = con;
}

static void e3visc(const orc_common *c, qpstate *s) {
  double rmu = s->rmu, rlm = s->rlm, rlm2mu = s->rlm2mu, con = s->con;
  double u1 = s->u1, u2 = s->u2, u3 = s->u3;
  if (c->lhs == 1) visc_stiff(s, s->stiff);
  double *g1 = s->g1yi, *g2 = s->g2yi, *g3 = s->g3yi, *rmi = s->rmi,
         *ri = s->ri;
  /* x1 (:278-292) */
  rmi[2] = rlm2mu * g1[2] + rlm * g2[3] + rlm * g3[4];
  rmi[3] = rmu * g1[3] + rmu * g2[2];
  rmi[4] = rmu * g1[4] + rmu * g3[2];
  rmi[5] = rlm2mu * u1 * g1[2] + rmu * u2 * g1[3] + rmu * u3 * g1[4] +
           rmu * u2 * g2[2] + rlm * u1 * g2[3] + rmu * u3 * g3[2] +
           rlm * u1 * g3[4] + con * g1[5];
  for (int m = 2; m <= 5; m++) ri[m] += rmi[m];
  /* x2 (:300-317) */
  rmi[7] = rmu * g1[3] + rmu * g2[2];
  rmi[8] = rlm * g1[2] + rlm2mu * g2[3] + rlm * g3[4];
  rmi[9] = rmu * g2[4] + rmu * g3[3];
  rmi[10] = rlm * u2 * g1[2] + rmu * u1 * g1[3] + rmu * u1 * g2[2] +
            rlm2mu * u2 * g2[3] + rmu * u3 * g2[4] + rmu * u3 * g3[3] +
            rlm * u2 * g3[4] + con * g2[5];
  for (int m = 7; m <= 10; m++) ri[m] += rmi[m];
  /* x3 (:325-343) */
  rmi[12] = rmu * g1[4] + rmu * g3[2];
  rmi[13] = rmu * g2[4] + rmu * g3[3];
  rmi[14] = rlm * g1[2] + rlm * g2[3] + rlm2mu * g3[4];
  rmi[15] = rlm * u3 * g1[2] + rmu * u1 * g1[4] + rlm * u3 * g2[3] +
            rmu * u2 * g2[4] + rmu * u1 * g3[2] + rmu * u2 * g3[3] +
            rlm2mu * u3 * g3[4] + con * g3[5];
  for (int m = 12; m <= 15; m++) ri[m] += rmi[m];
}

/* e3gijd (compressible/e3tau.f:1397-1507) */
static void e3gijd(const qpstate *s, double gijd[7]) {
  const double(*d)[4] = s->dxidx;
  if (s->lcsyst >= 2) {
    gijd[1] = d[1][1] * d[1][1] + d[2][1] * d[2][1] + d[3][1] * d[3][1];
    gijd[2] = d[1][1] * d[1][2] + d[2][1] * d[2][2] + d[3][1] * d[3][2];
    gijd[3] = d[1][2] * d[1][2] + d[2][2] * d[2][2] + d[3][2] * d[3][2];
    gijd[4] = d[1][1] * d[1][3] + d[2][1] * d[2][3] + d[3][1] * d[3][3];
    gijd[5] = d[1][2] * d[1][3] + d[2][2] * d[2][3] + d[3][2] * d[3][3];
    gijd[6] = d[1][3] * d[1][3] + d[2][3] * d[2][3] + d[3][3] * d[3][3];
  } else {
    const double c1 = 1.259921049894873e+00, c2 = 6.299605249474365e-01;
    double t1, t2, t3;
    t1 = c1 * d[1][1] + c2 * (d[2][1] + d[3][1]);
    t2 = c1 * d[2][1] + c2 * (d[1][1] + d[3][1]);
    t3 = c1 * d[3][1] + c2 * (d[1][1] + d[2][1]);
    gijd[1] = d[1][1] * t1 + d[2][1] * t2 + d[3][1] * t3;
    t1 = c1 * d[1][2] + c2 * (d[2][2] + d[3][2]);
    t2 = c1 * d[2][2] + c2 * (d[1][2] + d[3][2]);
    t3 = c1 * d[3][2] + c2 * (d[1][2] + d[2][2]);
    gijd[2] = d[1][1] * t1 + d[2][1] * t2 + d[3][1] * t3;
    gijd[3] = d[1][2] * t1 + d[2][2] * t2 + d[3][2] * t3;
    t1 = c1 * d[1][3] + c2 * (d[2][3] + d[3][3]);
    t2 = c1 * d[2][3] + c2 * (d[1][3] + d[3][3]);
    t3 = c1 * d[3][3] + c2 * (d[1][3] + d[2][3]);
    gijd[4] = d[1][1] * t1 + d[2][1] * t2 + d[3][1] * t3;
    gijd[5] = d[1][2] * t1 + d[2][2] * t2 + d[3][2] * t3;
    gijd[6] = d[1][3] * t1 + d[2][3] * t2 + d[3][3] * t3;
  }
}

/* e3tau, itau=0 branch (compressible/e3tau.f:53-58,140-183,273-279) */
static void e3tau(const orc_common *c, qpstate *s) {
  double gijd[7];
  e3gijd(s, gijd);
  if (c->itau != 0) {
    fprintf(stderr, "oracle e3tau: only itau=0 restated\n");
    abort();
  }
  double fff = 36.0;
  if (c->ipord == 2) fff = 60.0;
  if (c->ipord == 3) fff = 128.0;
  double dts = (c->iremoveStabTimeTerm == 1) ? 0.0 : c->dtsfct * c->Dtgl;
  double rho = s->rho, u1 = s->u1, u2 = s->u2, u3 = s->u3, rmu = s->rmu;
  s->tau[2] =
      rho * rho *
          ((2.0 * dts) * (2.0 * dts) +
           (u1 * (u1 * gijd[1] + 2.0 * (u2 * gijd[2] + u3 * gijd[4])) +
            u2 * (u2 * gijd[3] + 2.0 * u3 * gijd[5]) + u3 * u3 * gijd[6])) +
      fff * rmu * rmu *
          (gijd[1] * gijd[1] + gijd[3] * gijd[3] + gijd[6] * gijd[6] +
           2.0 * (gijd[2] * gijd[2] + gijd[4] * gijd[4] + gijd[5] * gijd[5]));
  double fact = sqrt(s->tau[2]);
  s->tau[1] =
      0.125 * fact / (rho * (gijd[1] + gijd[3] + gijd[6])) * c->taucfct;
  s->tau[2] = 1.0 / fact;
  s->tau[3] = s->tau[2] / s->cv * c->temper;
  double rt[6];
  if (c->iDC != 0)
    for (int m = 1; m <= 5; m++) rt[m] = s->rLyi[m]; /* rLyitemp, e3tau.f:177 */
  if (c->ires == 3 || c->ires == 1) {
    s->rLyi[1] *= s->tau[1];
    s->rLyi[2] *= s->tau[2];
    s->rLyi[3] *= s->tau[2];
    s->rLyi[4] *= s->tau[2];
    s->rLyi[5] *= s->tau[3];
  }
  if (c->iDC != 0) { /* e3tau.f:186-247 */
    const double *V = s->dVdY, *Ai = s->A0inv, *r = s->rLyi;
    s->rTLS = rt[1] * (r[1] * V[1] + V[2] * r[2] + V[4] * r[3] + r[4] * V[7] + V[11] * r[5]) +
              rt[2] * (r[2] * V[3] + r[3] * V[5] + V[8] * r[4] + r[5] * V[12]) +
              rt[3] * (r[3] * V[6] + V[9] * r[4] + V[13] * r[5]) + rt[4] * (r[4] * V[10] + V[14] * r[5]) +
              rt[5] * (V[15] * r[5]);
    s->raLS = 2.0 * rt[4] * rt[5] * Ai[15] + 2.0 * rt[3] * rt[5] * Ai[14] + 2.0 * rt[1] * rt[2] * Ai[6] +
              2.0 * rt[2] * rt[3] * Ai[10] + 2.0 * rt[2] * rt[4] * Ai[11] + 2.0 * rt[1] * rt[3] * Ai[7] +
              2.0 * rt[3] * rt[4] * Ai[13] + 2.0 * rt[2] * rt[5] * Ai[12] + 2.0 * rt[1] * rt[4] * Ai[8] +
              2.0 * rt[1] * rt[5] * Ai[9] + rt[1] * rt[1] * Ai[1] + rt[2] * rt[2] * Ai[2] + rt[3] * rt[3] * Ai[3] +
              rt[4] * rt[4] * Ai[4] + rt[5] * rt[5] * Ai[5];
    double gu[7];
    gu[1] = gijd[1];
    gu[2] = gijd[3];
    gu[3] = gijd[6];
    gu[4] = gijd[2];
    gu[5] = gijd[4];
    gu[6] = gijd[5];
    double detI = 1.0 / (gu[1] * gu[2] * gu[3] - gu[1] * gu[6] * gu[6] - gu[4] * gu[4] * gu[3] +
                         gu[4] * gu[5] * gu[6] * 2.0 - gu[5] * gu[5] * gu[2]);
    s->giju[1] = detI * (gu[2] * gu[3] - gu[6] * gu[6]);
    s->giju[2] = detI * (gu[1] * gu[3] - gu[5] * gu[5]);
    s->giju[3] = detI * (gu[1] * gu[2] - gu[4] * gu[4]);
    s->giju[4] = detI * (gu[5] * gu[6] - gu[4] * gu[3]);
    s->giju[5] = detI * (gu[4] * gu[6] - gu[5] * gu[2]);
    s->giju[6] = detI * (gu[4] * gu[5] - gu[1] * gu[6]);
  }
  if (c->ires != 1) {
    s->rLymi[1] *= s->tau[1];
    s->rLymi[2] *= s->tau[2];
    s->rLymi[3] *= s->tau[2];
    s->rLymi[4] *= s->tau[2];
    s->rLymi[5] *= s->tau[3];
  }
}

/* e3LS (compressible/e3ls.f:87-771) */
static void e3ls(const orc_common *c, qpstate *s, double *EG, size_t eg_stride,
                 int nedof) {
  double fct1 = c->almi / c->gami / c->alfi * c->Dtgl;
  if (c->ires != 1)
    for (int m = 1; m <= 5; m++) s->rLymi[m] = s->rLyi[m] + fct1 * s->dui[m];
  if (c->ires == 1 || c->ires == 3) {
    for (int m = 1; m <= 5; m++) {
      double acc = s->rLyi[m];
      for (int n = 1; n <= 5; n++) acc += s->A0[m][n] * s->aci[n];
      s->rLyi[m] = acc;
    }
  }
  if (c->idiff >= 1 && (c->ires == 3 || c->ires == 1)) {
    s->rLyi[2] -= s->divqi[1];
    s->rLyi[3] -= s->divqi[2];
    s->rLyi[4] -= s->divqi[3];
    s->rLyi[5] -= s->divqi[4];
  }
  e3tau(c, s);
  double(*A[4])[6] = {NULL, s->A1, s->A2, s->A3};
  if (c->ires != 1) {
    for (int i = 1; i <= 3; i++)
      for (int m = 1; m <= 5; m++) {
        double acc = 0.0;
        for (int n = 1; n <= 5; n++) acc += A[i][m][n] * s->rLymi[n];
        s->rmi[5 * (i - 1) + m] = acc + s->rmi[5 * (i - 1) + m];
      }
  }
  if (c->ires == 3 || c->ires == 1) {
    for (int i = 1; i <= 3; i++)
      for (int m = 1; m <= 5; m++) {
        double acc = 0.0;
        for (int n = 1; n <= 5; n++) acc += A[i][m][n] * s->rLyi[n];
        s->ri[5 * (i - 1) + m] = acc + s->ri[5 * (i - 1) + m];
      }
  }
  if (c->lhs == 1) {
    double Atau[6][6], AtauA0[4][6][6];
    for (int ii = 1; ii <= 3; ii++) {
      for (int i = 1; i <= 5; i++) {
        Atau[i][1] = A[ii][i][1] * s->tau[1];
        Atau[i][2] = A[ii][i][2] * s->tau[2];
        Atau[i][3] = A[ii][i][3] * s->tau[2];
        Atau[i][4] = A[ii][i][4] * s->tau[2];
        Atau[i][5] = A[ii][i][5] * s->tau[3];
      }
      for (int j = 1; j <= 5; j++)
        for (int i = 1; i <= 5; i++)
          AtauA0[ii][i][j] = Atau[i][1] * s->A0[1][j] + Atau[i][2] * s->A0[2][j] +
                             Atau[i][3] * s->A0[3][j] + Atau[i][4] * s->A0[4][j] +
                             Atau[i][5] * s->A0[5][j];
      for (int jj = 1; jj <= 3; jj++)
        for (int j = 1; j <= 5; j++)
          for (int i = 1; i <= 5; i++)
            s->stiff[i + 5 * (ii - 1)][j + 5 * (jj - 1)] +=
                (Atau[i][1] * A[jj][1][j] + Atau[i][2] * A[jj][2][j] +
                 Atau[i][3] * A[jj][3][j] + Atau[i][4] * A[jj][4][j] +
                 Atau[i][5] * A[jj][5][j]);
    }
    /* LS time term (:713-760) */
    for (int i = 1; i <= s->nshl; i++) {
      int i0 = 5 * (i - 1);
      for (int idof = 1; idof <= 5; idof++)
        for (int jdof = 1; jdof <= 5; jdof++)
          Atau[idof][jdof] = s->shg[i][1] * AtauA0[1][idof][jdof] +
                             s->shg[i][2] * AtauA0[2][idof][jdof] +
                             s->shg[i][3] * AtauA0[3][idof][jdof];
      for (int j = 1; j <= s->nshl; j++) {
        int j0 = 5 * (j - 1);
        double fact =
            s->shape[j] * s->WdetJ * c->almi / c->gami / c->alfi * c->Dtgl;
        for (int idof = 1; idof <= 5; idof++)
          for (int jdof = 1; jdof <= 5; jdof++)
            EGE(i0 + idof, j0 + jdof) += fact * Atau[idof][jdof];
      }
    }
  }
}

/* e3DC (compressible/e3dc.f:1-330): discontinuity-capturing viscosity DC (iDC = 1, 2, 3), its flux
 * DC g^ij A0 Y,j into ri/rmi and its tangent DC g^ij A0 into stiff.  The statement for rmi(:,11) reads
 * rmi(:,12) and gAgyi(:,12) (e3dc.f:262) -- kept. */
static void e3dc(const orc_common *c, qpstate *s) {
  const double *g[4] = {NULL, s->g1yi, s->g2yi, s->g3yi};
  const double *gj = s->giju, *D = s->A0DC;
  double A0g[16], gA[16], yy[7];
  for (int i = 1; i <= 3; i++)
    for (int m = 1; m <= 5; m++)
      A0g[5 * (i - 1) + m] = s->A0[m][1] * g[i][1] + s->A0[m][2] * g[i][2] + s->A0[m][3] * g[i][3] +
                             s->A0[m][4] * g[i][4] + s->A0[m][5] * g[i][5];
  for (int m = 1; m <= 5; m++) {
    gA[m] = gj[1] * A0g[m] + gj[4] * A0g[5 + m] + gj[5] * A0g[10 + m];
    gA[5 + m] = gj[4] * A0g[m] + gj[2] * A0g[5 + m] + gj[6] * A0g[10 + m];
    gA[10 + m] = gj[5] * A0g[m] + gj[6] * A0g[5 + m] + gj[3] * A0g[10 + m];
  }
  for (int i = 1; i <= 3; i++)
    yy[i] = D[1] * (g[i][1] * g[i][1]) + 2.0 * g[i][1] * D[2] * g[i][5] + D[3] * (g[i][2] * g[i][2]) +
            D[3] * (g[i][3] * g[i][3]) + D[3] * (g[i][4] * g[i][4]) + D[4] * (g[i][5] * g[i][5]);
  const int pi[4] = {0, 1, 1, 2}, pj[4] = {0, 2, 3, 3};
  for (int k = 1; k <= 3; k++) {
    const double *a = g[pi[k]], *b = g[pj[k]];
    yy[3 + k] = a[1] * D[1] * b[1] + a[1] * D[2] * b[5] + a[2] * D[3] * b[2] + a[3] * D[3] * b[3] +
                a[4] * D[3] * b[4] + a[5] * D[2] * b[1] + a[5] * D[4] * b[5];
  }
  double gnorm = 1.0 / (gj[1] * yy[1] + 2.0 * gj[4] * yy[4] + 2.0 * gj[5] * yy[5] + gj[2] * yy[2] +
                        2.0 * gj[6] * yy[6] + gj[3] * yy[3] + c->epsM);
  double DC = 0.0;
  if (c->iDC == 1) {
    double fact = 1.0;
    if (c->ipord == 2) fact = 0.9;
    if (c->ipord == 3) fact = 0.75;
    DC = fmax(0.0, (fact * sqrt(s->raLS * gnorm)) - (s->rTLS * gnorm));
  } else if (c->iDC == 2) {
    DC = 2.0 * s->rTLS * gnorm;
  } else if (c->iDC == 3) {
    double fact = 1.0;
    if (c->ipord == 2) fact = 0.5;
    DC = fmin(fmax(0.0, fact * sqrt(s->raLS * gnorm) - s->rTLS * gnorm), 2.0 * s->rTLS * gnorm);
  }
  s->DC = DC;
  if (c->ires == 1 || c->ires == 3) {
    for (int k = 1; k <= 15; k++) {
      s->ri[k] = s->ri[k] + DC * gA[k];
      if (k != 11) s->rmi[k] = s->rmi[k] + DC * gA[k];
      else s->rmi[11] = s->rmi[12] + DC * gA[12]; /* the reference's statement (:262), before rmi(:,12) is updated */
    }
  }
  if (c->ires == 2)
    for (int k = 1; k <= 15; k++) s->rmi[k] = s->rmi[k] + DC * gA[k];
  if (c->iprec == 1 && c->lhs == 1) {
    for (int j = 1; j <= 5; j++)
      for (int i = 1; i <= 5; i++) {
        double dt = s->A0[i][j] * DC;
        s->stiff[i][j] += dt * gj[1];
        s->stiff[i][j + 5] += dt * gj[4];
        s->stiff[i][j + 10] += dt * gj[5];
        s->stiff[i + 5][j] += dt * gj[4];
        s->stiff[i + 5][j + 5] += dt * gj[2];
        s->stiff[i + 5][j + 10] += dt * gj[6];
        s->stiff[i + 10][j] += dt * gj[5];
        s->stiff[i + 10][j + 5] += dt * gj[6];
        s->stiff[i + 10][j + 10] += dt * gj[3];
      }
  }
}

/* e3massr (compressible/e3massr.f:33-76) */
static void e3massr(const orc_common *c, qpstate *s) {
  if (c->ires == 1 || c->ires == 3)
    for (int m = 1; m <= 5; m++) {
      double acc = s->ri[15 + m];
      for (int n = 1; n <= 5; n++) acc += s->A0[m][n] * s->aci[n];
      s->ri[15 + m] = acc;
    }
  if (c->ires != 1) {
    double fct1 = c->almi / c->gami / c->alfi * c->Dtgl;
    for (int m = 1; m <= 5; m++) s->rmi[15 + m] += fct1 * s->dui[m];
  }
}

/* e3juel (compressible/e3juel.f:50-151), ires in {1,3} part; yl aliases
 * ycl and is overwritten (asigmr.f:75) */
static void e3juel(const orc_common *c, qpstate *s, double yl[][6],
                   double acl[][6], double rl[][6]) {
  double fact = s->WdetJ / (QWT(c, s->lcsyst, s->intp) * 15.0);
  if (c->ires == 1 || c->ires == 3) {
    double ub[6];
    for (int m = 1; m <= 5; m++)
      ub[m] = acl[1][m] + acl[2][m] + acl[3][m] + acl[4][m];
    for (int i = 1; i <= s->nshl; i++)
      for (int m = 1; m <= 5; m++) yl[i][m] = fact * (acl[i][m] + ub[m]);
    for (int i = 1; i <= s->nshl; i++)
      for (int m = 1; m <= 5; m++) {
        double acc = rl[i][m];
        for (int n = 1; n <= 5; n++) acc += s->A0[m][n] * yl[i][n];
        rl[i][m] = acc;
      }
  }
}

/* e3massl (compressible/e3massl.f:29-64), bcool=0 */
static void e3massl(const orc_common *c, qpstate *s, double *EG,
                    size_t eg_stride, int nedof) {
  double temp = s->WdetJ * c->almi / c->gami / c->alfi * c->Dtgl;
  for (int j = 1; j <= s->nshl; j++) {
    int j0 = 5 * (j - 1);
    for (int i = 1; i <= s->nshl; i++) {
      int i0 = 5 * (i - 1);
      double shpij = s->shape[i] * s->shape[j];
      double fact = shpij * temp;
      for (int jdof = 1; jdof <= 5; jdof++)
        for (int m = 1; m <= 5; m++)
          EGE(i0 + m, j0 + jdof) += fact * s->A0[m][jdof];
    }
  }
}

/* e3wmlt (compressible/e3wmlt.f:60-223) */
static void e3wmlt(const orc_common *c, qpstate *s, double rl[][6], double rml[][6], double *EG,
                   size_t eg_stride, int nedof) {
  double W = s->WdetJ;
  if ((c->ires == 2 || c->ires == 3) && rml) /* :95-122 */
    for (int i = 1; i <= s->nshl; i++)
      for (int m = 1; m <= 5; m++)
        rml[i][m] += W * (s->shg[i][1] * s->rmi[m] + s->shg[i][2] * s->rmi[5 + m] +
                          s->shg[i][3] * s->rmi[10 + m] + s->shape[i] * s->rmi[15 + m]);
  if (c->ires == 1 || c->ires == 3)
    for (int i = 1; i <= s->nshl; i++)
      for (int m = 1; m <= 5; m++)
        rl[i][m] += W * (s->shg[i][1] * s->ri[m] + s->shg[i][2] * s->ri[5 + m] +
                         s->shg[i][3] * s->ri[10 + m]);
  if (s->ngauss == 1 && s->nshl == 4) {
    /* mass already exactly integrated; body force absent (:126-134) */
  } else {
    for (int i = 1; i <= s->nshl; i++)
      for (int m = 1; m <= 5; m++) rl[i][m] += s->shape[i] * W * s->ri[15 + m];
  }
  if (c->lhs == 1) {
    double stif1[6][6], stif2[6][6], stif3[6][6];
    for (int j = 1; j <= s->nshl; j++) {
      int j0 = 5 * (j - 1);
      double shg1 = W * s->shg[j][1], shg2 = W * s->shg[j][2],
             shg3 = W * s->shg[j][3];
      for (int jdof = 1; jdof <= 5; jdof++)
        for (int idof = 1; idof <= 5; idof++) {
          stif1[idof][jdof] = shg1 * s->stiff[idof][jdof] +
                              shg2 * s->stiff[idof][jdof + 5] +
                              shg3 * s->stiff[idof][jdof + 10];
          stif2[idof][jdof] = shg1 * s->stiff[idof + 5][jdof] +
                              shg2 * s->stiff[idof + 5][jdof + 5] +
                              shg3 * s->stiff[idof + 5][jdof + 10];
          stif3[idof][jdof] = shg1 * s->stiff[idof + 10][jdof] +
                              shg2 * s->stiff[idof + 10][jdof + 5] +
                              shg3 * s->stiff[idof + 10][jdof + 10];
        }
      for (int i = 1; i <= s->nshl; i++) {
        int i0 = 5 * (i - 1);
        for (int jdof = 1; jdof <= 5; jdof++)
          for (int m = 1; m <= 5; m++)
            EGE(i0 + m, j0 + jdof) += s->shg[i][1] * stif1[m][jdof] +
                                      s->shg[i][2] * stif2[m][jdof] +
                                      s->shg[i][3] * stif3[m][jdof];
      }
    }
  }
}

/* getshp for ipord=1 (common/hierarchic.f:28-55): copy table columns */
static void getshp(const orc_part *p, qpstate *s) {
  for (int n = 1; n <= s->nshl; n++) {
    s->shape[n] = SHP(p, s->lcsyst, n, s->intp);
    for (int i = 1; i <= 3; i++) s->shdrv[i][n] = SHGL(p, s->lcsyst, i, n, s->intp);
  }
}

/* e3bdg (compressible/e3bdg.f:1-2594), itau<10, bcool=0, ivart>=2, ngauss>1:
 * the block-diagonal preconditioner of the matrix-free solver built directly,
 *   BDiagl(a) += N_a W N_a,i A_i                       (:34-189, "Ex-E3conv")
 *              + N_a N_a W fct1 A0                      (:218-246, "Ex-e3mass")
 *              + N_a W fct1 N_a,i A_i tau A0            (:251-343)
 *              + W N_a,i N_a,j (K_ij + A_i tau A_j)     (:344-2590)
 * i.e. the four terms of EGmass(a,a).  The reference writes every entry out
 * with the structural zeros of A_i removed; the sums below are the same terms
 * in loop form (differences are at round-off level). */
static void e3bdg(const orc_common *c, const qpstate *s, double BDl[][6][6]) {
  const double(*A[4])[6] = {NULL, s->A1, s->A2, s->A3};
  double fct1 = c->almi / c->gami / c->alfi * c->Dtgl;
  double W = s->WdetJ;
  double K[16][16];
  memset(K, 0, sizeof K);
  if (c->Navier == 1) visc_stiff(s, K);
  double Atau[4][6][6], AtauA0[4][6][6];
  for (int ii = 1; ii <= 3; ii++) {
    for (int i = 1; i <= 5; i++) {
      Atau[ii][i][1] = A[ii][i][1] * s->tau[1];
      Atau[ii][i][2] = A[ii][i][2] * s->tau[2];
      Atau[ii][i][3] = A[ii][i][3] * s->tau[2];
      Atau[ii][i][4] = A[ii][i][4] * s->tau[2];
      Atau[ii][i][5] = A[ii][i][5] * s->tau[3];
    }
    for (int j = 1; j <= 5; j++)
      for (int i = 1; i <= 5; i++) {
        double acc = 0.0;
        for (int k = 1; k <= 5; k++) acc += Atau[ii][i][k] * s->A0[k][j];
        AtauA0[ii][i][j] = acc;
      }
    for (int jj = 1; jj <= 3; jj++)
      for (int j = 1; j <= 5; j++)
        for (int i = 1; i <= 5; i++) {
          double acc = 0.0;
          for (int k = 1; k <= 5; k++) acc += Atau[ii][i][k] * A[jj][k][j];
          K[i + 5 * (ii - 1)][j + 5 * (jj - 1)] += acc;
        }
  }
  for (int a = 1; a <= s->nshl; a++) {
    double Na = s->shape[a];
    for (int i = 1; i <= 5; i++)
      for (int k = 1; k <= 5; k++) {
        double v = 0.0;
        v += Na * W * (s->shg[a][1] * s->A1[i][k] + s->shg[a][2] * s->A2[i][k] + s->shg[a][3] * s->A3[i][k]);
        v += (Na * Na) * (W * fct1) * s->A0[i][k];
        v += (Na * W * fct1) * (s->shg[a][1] * AtauA0[1][i][k] + s->shg[a][2] * AtauA0[2][i][k] +
                                s->shg[a][3] * AtauA0[3][i][k]);
        for (int ii = 1; ii <= 3; ii++)
          for (int jj = 1; jj <= 3; jj++)
            v += W * s->shg[a][ii] * s->shg[a][jj] * K[i + 5 * (ii - 1)][k + 5 * (jj - 1)];
        BDl[a][i][k] += v;
      }
  }
}

/* e3 (compressible/e3.f:1-306) for element iel of a block; EG points at
 * EGmass(iel_global,1,1), stride numel.  yl: see e3ivar; rml / BDl nullable. */
static void e3_element(const orc_part *p, int lcsyst, int nshl, int nenl,
                       int ngauss, double yl[][6], double ycl[][6], double acl[][6],
                       double xl[][4], double ql[][13], double rl[][6], double rml[][6],
                       double BDl[][6][6], double *EG, size_t eg_stride, int nedof) {
  const orc_common *c = &p->c;
  qpstate s;
  s.nshl = nshl;
  s.nenl = nenl;
  s.lcsyst = lcsyst;
  s.ngauss = ngauss;
  for (int intp = 1; intp <= ngauss; intp++) {
    if (QWT(c, lcsyst, intp) == 0.0) continue;
    s.intp = intp;
    getshp(p, &s);
    memset(s.ri, 0, sizeof s.ri);
    memset(s.rmi, 0, sizeof s.rmi);
    if (c->lhs == 1) memset(s.stiff, 0, sizeof s.stiff);
    e3ivar(c, &s, yl, ycl, acl, xl, ql);
    e3mtrx(&s);
    if (c->iDC != 0) e3mtrx_dc(&s);
    e3conv(c, &s, EG, eg_stride, nedof);
    if (c->Navier == 1) e3visc(c, &s);
    e3ls(c, &s, EG, eg_stride, nedof);
    if (c->iDC != 0) e3dc(c, &s); /* e3.f:217-222 */
    if (ngauss == 1 && nshl == 4)
      e3juel(c, &s, yl, acl, rl);
    else
      e3massr(c, &s);
    if (c->lhs == 1) e3massl(c, &s, EG, eg_stride, nedof);
    if (c->iprec == 1 && c->lhs != 1 && BDl) e3bdg(c, &s, BDl); /* e3.f:258-285 */
    e3wmlt(c, &s, rl, rml, EG, eg_stride, nedof);
  }
}

/* AsIGMR (compressible/asigmr.f:1-119) for one block.  The per-element
 * loop is hoisted outside the routine sequence (every statement of the
 * reference is element-wise), but the scatter keeps local.f's order. */
void orc_asigmr(const orc_part *p, int iblk, const double *qres, double *res,
                double *BDiag, double *EGmass) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nshl = lc[9];
  int npro = lc[10] - iel;
  int ngauss = c->nint[lcsyst - 1];
  const int *ien = p->ien + p->ien_off[iblk];
  int nshg = c->nshg, nedof = c->nedof, idflx = (c->idiff >= 1) ? 12 : 0;
  size_t numel = (size_t)c->numel;
  double(*rl)[ORC_MAXSH + 1][6] = calloc((size_t)npro, sizeof *rl);
  for (int e = 0; e < npro; e++) {
    double ycl[ORC_MAXSH + 1][6], acl[ORC_MAXSH + 1][6], xl[ORC_MAXSH + 1][4],
        ql[ORC_MAXSH + 1][13];
    memset(ql, 0, sizeof ql);
    for (int n = 1; n <= nshl; n++) {
      int A = ien[e + (size_t)npro * (n - 1)] - 1;
      /* localy (common/localy.f:47-72): {u,v,w,p,T} -> {p,u,v,w,T} */
      ycl[n][1] = p->y[A + (size_t)nshg * 3];
      ycl[n][2] = p->y[A + (size_t)nshg * 0];
      ycl[n][3] = p->y[A + (size_t)nshg * 1];
      ycl[n][4] = p->y[A + (size_t)nshg * 2];
      ycl[n][5] = p->y[A + (size_t)nshg * 4];
      acl[n][1] = p->ac[A + (size_t)nshg * 3];
      acl[n][2] = p->ac[A + (size_t)nshg * 0];
      acl[n][3] = p->ac[A + (size_t)nshg * 1];
      acl[n][4] = p->ac[A + (size_t)nshg * 2];
      acl[n][5] = p->ac[A + (size_t)nshg * 4];
      for (int i = 1; i <= 3; i++)
        xl[n][i] = p->x[A + (size_t)c->numnp * (i - 1)];
      for (int k = 1; k <= idflx; k++)
        ql[n][k] = qres[A + (size_t)nshg * (k - 1)];
    }
    /* the reference passes the strided section EGmass(iel:inum,:,:)
     * (elmgmr.f:160); gfortran packs it into a contiguous temporary.  Here the
     * element's nedof x nedof slab is accumulated contiguously and copied out. */
    double EGl[1600]; /* nedof <= 40 (hexes) */
    if (EGmass) {
      if (nedof * nedof > 1600) { fprintf(stderr, "orc_asigmr: nedof>40\n"); abort(); }
      memset(EGl, 0, sizeof(double) * (size_t)nedof * nedof);
    }
    e3_element(p, lcsyst, nshl, nenl, ngauss, ycl, ycl, acl, xl, ql, rl[e], NULL, NULL,
               EGmass ? EGl : NULL, 1, nedof);
    if (EGmass) {
      double *EG = EGmass + (size_t)(iel - 1 + e);
      for (int k = 0; k < nedof * nedof; k++) EG[numel * (size_t)k] += EGl[k];
    }
  }
  /* local(res, rl, 'scatter') (common/local.f:67-74): dof, node, element */
  for (int j = 1; j <= 5; j++)
    for (int i = 1; i <= nshl; i++)
      for (int e = 0; e < npro; e++) {
        int A = ien[e + (size_t)npro * (i - 1)] - 1;
        res[A + (size_t)nshg * (j - 1)] += rl[e][i][j];
      }
  /* BDiag extraction + scatter (asigmr.f:92-102) */
  if (c->iprec != 0 && EGmass) {
    for (int k = 1; k <= 5; k++)
      for (int j = 1; j <= 5; j++)
        for (int i = 1; i <= nshl; i++)
          for (int e = 0; e < npro; e++) {
            int A = ien[e + (size_t)npro * (i - 1)] - 1;
            int i0 = (i - 1) * 5 + j, j0 = (i - 1) * 5 + k;
            BDiag[A + (size_t)nshg * ((j - 1) + 5 * (k - 1))] +=
                EGmass[(size_t)(iel - 1 + e) +
                       numel * ((i0 - 1) + (size_t)nedof * (j0 - 1))];
          }
  }
  free(rl);
}

/* AsIq + e3q + e3qvar (compressible/asiq.f:1-71, e3q.f:1-246, e3qvar.f) */
/* gather of one element's nodal data (localy / localx / local, asigmr.f:53-62) */
static void gather_elem(const orc_part *p, const int *ien, int npro, int e, int nshl, const double *y,
                        const double *ac, const double *qres, double ycl[][6], double acl[][6],
                        double xl[][4], double ql[][13]) {
  const orc_common *c = &p->c;
  int nshg = c->nshg, idflx = (c->idiff >= 1 && qres) ? 12 : 0;
  for (int n = 1; n <= nshl; n++) {
    int A = ien[e + (size_t)npro * (n - 1)] - 1;
    static const int src[6] = {0, 3, 0, 1, 2, 4}; /* localy.f:47-72 */
    for (int m = 1; m <= 5; m++) {
      if (ycl) ycl[n][m] = y[A + (size_t)nshg * src[m]];
      if (acl) acl[n][m] = ac[A + (size_t)nshg * src[m]];
    }
    if (xl)
      for (int i = 1; i <= 3; i++) xl[n][i] = p->x[A + (size_t)c->numnp * (i - 1)];
    if (ql) {
      memset(ql[n], 0, sizeof(double) * 13);
      for (int k = 1; k <= idflx; k++) ql[n][k] = qres[A + (size_t)nshg * (k - 1)];
    }
  }
}

/* AsIMFG (compressible/asimfg.f:1-110) for one block: ires=3 residual rl,
 * modified residual rml and (iprec=1, lhs/=1) the e3bdg block diagonal. */
void orc_asimfg(const orc_part *p, int iblk, const double *qres, double *res, double *rmes,
                double *BDiag) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nshl = lc[9];
  int npro = lc[10] - iel;
  int ngauss = c->nint[lcsyst - 1];
  const int *ien = p->ien + p->ien_off[iblk];
  int nshg = c->nshg;
  if (ngauss == 1 && nshl == 4) {
    fprintf(stderr, "orc_asimfg: the matrix-free path is restated for ngauss>1 only\n");
    abort();
  }
  for (int e = 0; e < npro; e++) {
    double ycl[ORC_MAXSH + 1][6], acl[ORC_MAXSH + 1][6], xl[ORC_MAXSH + 1][4], ql[ORC_MAXSH + 1][13];
    double rl[ORC_MAXSH + 1][6], rml[ORC_MAXSH + 1][6], BDl[ORC_MAXSH + 1][6][6];
    memset(rl, 0, sizeof rl);
    memset(rml, 0, sizeof rml);
    memset(BDl, 0, sizeof BDl);
    gather_elem(p, ien, npro, e, nshl, p->y, p->ac, qres, ycl, acl, xl, ql);
    e3_element(p, lcsyst, nshl, nenl, ngauss, ycl, ycl, acl, xl, ql, rl, rml, BDl, NULL, 1, c->nedof);
    for (int i = 1; i <= nshl; i++) {
      int A = ien[e + (size_t)npro * (i - 1)] - 1;
      for (int j = 1; j <= 5; j++) {
        res[A + (size_t)nshg * (j - 1)] += rl[i][j];
        rmes[A + (size_t)nshg * (j - 1)] += rml[i][j];
        if (c->iprec != 0)
          for (int k = 1; k <= 5; k++) BDiag[A + (size_t)nshg * ((j - 1) + 5 * (k - 1))] += BDl[i][j][k];
      }
    }
  }
}

/* AsIRes (compressible/asires.f:1-94) for one block: ires=2 modified residual
 * with yl = yp (perturbed, (nshg,5) {u,p,T} global order) and ycl = p->y. */
void orc_asires(const orc_part *p, int iblk, const double *yp, double *rmes, int iabres) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nshl = lc[9];
  int npro = lc[10] - iel;
  int ngauss = c->nint[lcsyst - 1];
  const int *ien = p->ien + p->ien_off[iblk];
  int nshg = c->nshg;
  for (int e = 0; e < npro; e++) {
    double yl[ORC_MAXSH + 1][6], ycl[ORC_MAXSH + 1][6], acl[ORC_MAXSH + 1][6], xl[ORC_MAXSH + 1][4],
        ql[ORC_MAXSH + 1][13];
    double rml[ORC_MAXSH + 1][6];
    memset(rml, 0, sizeof rml);
    memset(ql, 0, sizeof ql);
    gather_elem(p, ien, npro, e, nshl, p->y, p->ac, NULL, ycl, acl, xl, NULL);
    gather_elem(p, ien, npro, e, nshl, yp, p->ac, NULL, yl, NULL, NULL, NULL);
    e3_element(p, lcsyst, nshl, nenl, ngauss, yl, ycl, acl, xl, ql, rml, rml, NULL, NULL, 1, c->nedof);
    for (int i = 1; i <= nshl; i++) {
      int A = ien[e + (size_t)npro * (i - 1)] - 1;
      for (int j = 1; j <= 5; j++) {
        double v = rml[i][j];
        if (iabres == 1) v = fabs(v);
        rmes[A + (size_t)nshg * (j - 1)] += v;
      }
    }
  }
}

void orc_asiq(const orc_part *p, int iblk, double *qres, double *rmass) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nshl = lc[9];
  int npro = lc[10] - iel;
  int ngauss = c->nint[lcsyst - 1];
  const int *ien = p->ien + p->ien_off[iblk];
  int nshg = c->nshg;
  double(*qll)[ORC_MAXSH + 1][13] = calloc((size_t)npro, sizeof *qll);
  double(*rml)[ORC_MAXSH + 1] = calloc((size_t)npro, sizeof *rml);
  for (int e = 0; e < npro; e++) {
    double ycl[ORC_MAXSH + 1][6], xl[ORC_MAXSH + 1][4];
    for (int n = 1; n <= nshl; n++) {
      int A = ien[e + (size_t)npro * (n - 1)] - 1;
      ycl[n][1] = p->y[A + (size_t)nshg * 3];
      ycl[n][2] = p->y[A + (size_t)nshg * 0];
      ycl[n][3] = p->y[A + (size_t)nshg * 1];
      ycl[n][4] = p->y[A + (size_t)nshg * 2];
      ycl[n][5] = p->y[A + (size_t)nshg * 4];
      for (int i = 1; i <= 3; i++)
        xl[n][i] = p->x[A + (size_t)c->numnp * (i - 1)];
    }
    qpstate s;
    s.nshl = nshl;
    s.nenl = nenl;
    s.lcsyst = lcsyst;
    s.ngauss = ngauss;
    for (int intp = 1; intp <= ngauss; intp++) {
      if (QWT(c, lcsyst, intp) == 0.0) continue;
      s.intp = intp;
      getshp(p, &s);
      /* e3qvar (e3qvar.f:60-135) */
      s.pres = s.u1 = s.u2 = s.u3 = s.T = 0.0;
      for (int n = 1; n <= nshl; n++) {
        s.pres += s.shape[n] * ycl[n][1];
        s.u1 += s.shape[n] * ycl[n][2];
        s.u2 += s.shape[n] * ycl[n][3];
        s.u3 += s.shape[n] * ycl[n][4];
        s.T += s.shape[n] * ycl[n][5];
      }
      getthm7(c, s.pres, s.T, &s.rho, &s.ei, &s.h, &s.cv, &s.cp, &s.alfap,
              &s.betaT);
      e3metric(c, &s, xl);
      for (int m = 1; m <= 5; m++) s.g1yi[m] = s.g2yi[m] = s.g3yi[m] = 0.0;
      for (int n = 1; n <= nshl; n++)
        for (int m = 1; m <= 5; m++) {
          s.g1yi[m] += s.shg[n][1] * ycl[n][m];
          s.g2yi[m] += s.shg[n][2] * ycl[n][m];
          s.g3yi[m] += s.shg[n][3] * ycl[n][m];
        }
      getdiff(c, s.T, s.cp, &s.rmu, &s.rlm, &s.rlm2mu, &s.con);
      double qdi[13];
      double rmu = s.rmu, rlm = s.rlm, rlm2mu = s.rlm2mu, con = s.con;
      double u1 = s.u1, u2 = s.u2, u3 = s.u3;
      double *g1 = s.g1yi, *g2 = s.g2yi, *g3 = s.g3yi;
      /* e3q.f:103-146 */
      qdi[1] = rlm2mu * g1[2] + rlm * g2[3] + rlm * g3[4];
      qdi[2] = rmu * g1[3] + rmu * g2[2];
      qdi[3] = rmu * g1[4] + rmu * g3[2];
      qdi[4] = rlm2mu * u1 * g1[2] + rmu * u2 * g1[3] + rmu * u3 * g1[4] +
               rmu * u2 * g2[2] + rlm * u1 * g2[3] + rmu * u3 * g3[2] +
               rlm * u1 * g3[4] + con * g1[5];
      qdi[5] = rmu * g1[3] + rmu * g2[2];
      qdi[6] = rlm * g1[2] + rlm2mu * g2[3] + rlm * g3[4];
      qdi[7] = rmu * g2[4] + rmu * g3[3];
      qdi[8] = rlm * u2 * g1[2] + rmu * u1 * g1[3] + rmu * u1 * g2[2] +
               rlm2mu * u2 * g2[3] + rmu * u3 * g2[4] + rmu * u3 * g3[3] +
               rlm * u2 * g3[4] + con * g2[5];
      qdi[9] = rmu * g1[4] + rmu * g3[2];
      qdi[10] = rmu * g2[4] + rmu * g3[3];
      qdi[11] = rlm * g1[2] + rlm * g2[3] + rlm2mu * g3[4];
      qdi[12] = rlm * u3 * g1[2] + rmu * u1 * g1[4] + rlm * u3 * g2[3] +
                rmu * u2 * g2[4] + rmu * u1 * g3[2] + rmu * u2 * g3[3] +
                rlm2mu * u3 * g3[4] + con * g3[5];
      for (int i = 1; i <= nshl; i++) {
        for (int k = 1; k <= 12; k++)
          qll[e][i][k] += s.shape[i] * s.WdetJ * qdi[k];
        if (c->idiff == 1) rml[e][i] += s.shape[i] * s.WdetJ;
      }
    }
  }
  for (int j = 1; j <= 12; j++)
    for (int i = 1; i <= nshl; i++)
      for (int e = 0; e < npro; e++) {
        int A = ien[e + (size_t)npro * (i - 1)] - 1;
        qres[A + (size_t)nshg * (j - 1)] += qll[e][i][j];
      }
  for (int i = 1; i <= nshl; i++)
    for (int e = 0; e < npro; e++) {
      int A = ien[e + (size_t)npro * (i - 1)] - 1;
      rmass[A] += rml[e][i];
    }
  free(qll);
  free(rml);
}

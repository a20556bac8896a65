/*
 * oracle_incomp.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 *
 * CPU restatement of the incompressible element assembly into block-CSR and of
 * the sparse matrix-vector products its Krylov solver calls back
 * (BASELINE.json configs[3]):
 *   ElmGMR            phSolver/incompressible/elmgmr.f:1-330
 *   AsIq/e3q/e3qvar   incompressible/asiq.f:1-68, e3q.f:1-137, e3qvar.f:1-100, common/qpbc.f
 *   AsIGMR/e3         incompressible/asigmr.f:1-110, e3.f:1-120
 *   getDiff           incompressible/getdiff.f:1-40 (iLSet=0, DNS: rho, mu constant)
 *   e3ivar + e3resStrongPDE  incompressible/e3ivar.f:1-250, e3res.f:300-400
 *   e3stab + e3gijd   incompressible/e3stab.f:1-230,330-420 (itau=0)
 *   e3Res             incompressible/e3res.f:1-200
 *   e3LHS             incompressible/e3lhs.f:1-230
 *   bc3LHS            incompressible/bc3lhs.f:1-380
 *   fillsparseI       common/fillsparse.f:1-65
 *   bc3Res/bc3per     incompressible/bc3res.f:1-90, bc3per.f:1-45
 *   fLesSparseApG/ApKG/ApNGt/ApNGtC/ApFull  incompressible/lesSparse.f:204-492
 * Scope: DNS (iLES=iRANS=iLSet=0), ipord=1, itau=0, idiff in {0,1}, iconvflow in
 * {1,2}, constant body force (matflg(5,1) in {0,1}), the boundary integral
 * AsBMFG/e3b/e3bvar with rigid walls (ideformwall=0) on tet, hex and wedge
 * faces, ipvsq=0.  Pinned against the
 * reference's own Fortran executed by f77np (tests/golden/f77_incomp_*.npz).
 */
#include "oracle_internal.h"
#include "oracle_incomp.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int orc_sizeof_incomp(void) { return (int)sizeof(orc_incomp); }

typedef struct iqp {
  int nshl, nenl, lcsyst, intp;
  double shape[ORC_MAXSH + 1], shdrv[4][ORC_MAXSH + 1];
  double shg[ORC_MAXSH + 1][4], dxidx[4][4], WdetJ;
  double pres, u1, u2, u3, aci[4], g1yi[5], g2yi[5], g3yi[5], divqi[4];
  double rho, rmu, rLui[4], src[4], tauC, tauM, tauBar, uBar[4];
} iqp;

static void getshp_i(const orc_part *p, iqp *s) {
  for (int n = 1; n <= s->nshl; n++) {
    s->shape[n] = SHP(p, s->lcsyst, n, s->intp);
    for (int i = 1; i <= 3; i++) s->shdrv[i][n] = SHGL(p, s->lcsyst, i, n, s->intp);
  }
}

/* e3metric (common/e3metric.f:8-80) and, statement for statement the same arithmetic,
 * the metric block of e3qvar (incompressible/e3qvar.f:17-72) */
static void metric_i(const orc_common *c, iqp *s, double xl[][4]) {
  double dxdxi[4][4];
  memset(dxdxi, 0, sizeof dxdxi);
  for (int n = 1; n <= s->nenl; n++)
    for (int i = 1; i <= 3; i++)
      for (int j = 1; j <= 3; j++) dxdxi[i][j] += xl[n][i] * s->shdrv[j][n];
  double(*d)[4] = s->dxidx;
  d[1][1] = dxdxi[2][2] * dxdxi[3][3] - dxdxi[3][2] * dxdxi[2][3];
  d[1][2] = dxdxi[3][2] * dxdxi[1][3] - dxdxi[1][2] * dxdxi[3][3];
  d[1][3] = dxdxi[1][2] * dxdxi[2][3] - dxdxi[1][3] * dxdxi[2][2];
  double tmp = 1.0 / (d[1][1] * dxdxi[1][1] + d[1][2] * dxdxi[2][1] + d[1][3] * dxdxi[3][1]);
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
      s->shg[n][i] = s->shdrv[1][n] * d[1][i] + s->shdrv[2][n] * d[2][i] + s->shdrv[3][n] * d[3][i];
}

/* e3ivar (incompressible/e3ivar.f:34-118,194-199) + e3resStrongPDE (e3res.f:300-400) */
static void e3ivar_i(const orc_common *c, const orc_incomp *ip, iqp *s, double yl[][6], double acl[][6],
                     double xl[][4], double ql[][10]) {
  s->pres = s->u1 = s->u2 = s->u3 = 0.0;
  for (int n = 1; n <= s->nshl; n++) {
    s->pres += s->shape[n] * yl[n][1];
    s->u1 += s->shape[n] * yl[n][2];
    s->u2 += s->shape[n] * yl[n][3];
    s->u3 += s->shape[n] * yl[n][4];
  }
  s->aci[1] = s->aci[2] = s->aci[3] = 0.0;
  for (int n = 1; n <= s->nshl; n++)
    for (int i = 1; i <= 3; i++) s->aci[i] += s->shape[n] * acl[n][i + 1];
  metric_i(c, s, xl);
  for (int m = 1; m <= 4; m++) s->g1yi[m] = s->g2yi[m] = s->g3yi[m] = 0.0;
  for (int n = 1; n <= s->nshl; n++)
    for (int m = 1; m <= 4; m++) {
      s->g1yi[m] += s->shg[n][1] * yl[n][m];
      s->g2yi[m] += s->shg[n][2] * yl[n][m];
      s->g3yi[m] += s->shg[n][3] * yl[n][m];
    }
  s->divqi[1] = s->divqi[2] = s->divqi[3] = 0.0;
  if (ip->idiff >= 1)
    for (int n = 1; n <= s->nshl; n++) {
      s->divqi[1] = s->divqi[1] + s->shg[n][1] * ql[n][1] + s->shg[n][2] * ql[n][4] + s->shg[n][3] * ql[n][7];
      s->divqi[2] = s->divqi[2] + s->shg[n][1] * ql[n][2] + s->shg[n][2] * ql[n][5] + s->shg[n][3] * ql[n][8];
      s->divqi[3] = s->divqi[3] + s->shg[n][1] * ql[n][3] + s->shg[n][2] * ql[n][6] + s->shg[n][3] * ql[n][9];
    }
  /* e3resStrongPDE */
  s->src[1] = s->src[2] = s->src[3] = 0.0;
  if (ip->matflg5 == 1)
    for (int i = 1; i <= 3; i++) s->src[i] = ip->bf[i - 1];
  double rho = s->rho;
  s->rLui[1] = (s->aci[1] + s->u1 * s->g1yi[2] + s->u2 * s->g2yi[2] + s->u3 * s->g3yi[2] - s->src[1]) * rho +
               s->g1yi[1] - s->divqi[1];
  s->rLui[2] = (s->aci[2] + s->u1 * s->g1yi[3] + s->u2 * s->g2yi[3] + s->u3 * s->g3yi[3] - s->src[2]) * rho +
               s->g2yi[1] - s->divqi[2];
  s->rLui[3] = (s->aci[3] + s->u1 * s->g1yi[4] + s->u2 * s->g2yi[4] + s->u3 * s->g3yi[4] - s->src[3]) * rho +
               s->g3yi[1] - s->divqi[3];
  if (ip->iconvflow == 1) {
    double divu = (s->g1yi[2] + s->g2yi[3] + s->g3yi[4]) * rho;
    s->rLui[1] = s->rLui[1] + s->u1 * divu;
    s->rLui[2] = s->rLui[2] + s->u2 * divu;
    s->rLui[3] = s->rLui[3] + s->u3 * divu;
  }
}

/* e3gijd (incompressible/e3stab.f:330-420): g(1..6) = 11,22,33,12,23,13 */
static void e3gijd_i(const iqp *s, double g[7]) {
  const double(*d)[4] = s->dxidx;
  if (s->lcsyst >= 2) {
    g[1] = d[1][1] * d[1][1] + d[2][1] * d[2][1] + d[3][1] * d[3][1];
    g[4] = d[1][1] * d[1][2] + d[2][1] * d[2][2] + d[3][1] * d[3][2];
    g[2] = d[1][2] * d[1][2] + d[2][2] * d[2][2] + d[3][2] * d[3][2];
    g[5] = d[1][2] * d[1][3] + d[2][2] * d[2][3] + d[3][2] * d[3][3];
    g[6] = d[1][1] * d[1][3] + d[2][1] * d[2][3] + d[3][1] * d[3][3];
    g[3] = d[1][3] * d[1][3] + d[2][3] * d[2][3] + d[3][3] * d[3][3];
  } else {
    const double c1 = 1.259921049894873e+00, c2 = 6.299605249474365e-01;
    double t1, t2, t3;
    t1 = c1 * d[1][1] + c2 * (d[2][1] + d[3][1]);
    t2 = c1 * d[2][1] + c2 * (d[1][1] + d[3][1]);
    t3 = c1 * d[3][1] + c2 * (d[1][1] + d[2][1]);
    g[1] = d[1][1] * t1 + d[2][1] * t2 + d[3][1] * t3;
    t1 = c1 * d[1][2] + c2 * (d[2][2] + d[3][2]);
    t2 = c1 * d[2][2] + c2 * (d[1][2] + d[3][2]);
    t3 = c1 * d[3][2] + c2 * (d[1][2] + d[2][2]);
    g[2] = d[1][2] * t1 + d[2][2] * t2 + d[3][2] * t3;
    g[4] = d[1][1] * t1 + d[2][1] * t2 + d[3][1] * t3;
    t1 = c1 * d[1][3] + c2 * (d[2][3] + d[3][3]);
    t2 = c1 * d[2][3] + c2 * (d[1][3] + d[3][3]);
    t3 = c1 * d[3][3] + c2 * (d[1][3] + d[2][3]);
    g[3] = d[1][3] * t1 + d[2][3] * t2 + d[3][3] * t3;
    g[5] = d[1][2] * t1 + d[2][2] * t2 + d[3][2] * t3;
    g[6] = d[1][1] * t1 + d[2][1] * t2 + d[3][1] * t3;
  }
}

/* e3stab, itau=0 (incompressible/e3stab.f:38-66,205-225) */
static void e3stab_i(const orc_incomp *ip, iqp *s) {
  double g[7];
  e3gijd_i(s, g);
  const double fff = 36.0; /* ipord == 1 */
  double u1 = s->u1, u2 = s->u2, u3 = s->u3;
  double rhoinv = 1.0 / s->rho;
  double rnu = s->rmu * rhoinv;
  double dts = ip->Dtgl * ip->dtsfct;
  double tauM = ((2.0 * dts) * (2.0 * dts) +
                 (u1 * (g[1] * u1 + g[4] * u2 + g[6] * u3) + u2 * (g[4] * u1 + g[2] * u2 + g[5] * u3) +
                  u3 * (g[6] * u1 + g[5] * u2 + g[3] * u3))) +
                fff * (rnu * rnu) *
                    (g[1] * g[1] + g[2] * g[2] + g[3] * g[3] + 2.0 * (g[4] * g[4] + g[5] * g[5] + g[6] * g[6])) +
                0.0; /* omegasq (no rotation) */
  double fact = sqrt(tauM);
  double ff = ip->taucfct / ip->dtsfct;
  s->tauC = s->rho * 0.125 * fact / (g[1] + g[2] + g[3]) * ff;
  s->tauM = 1.0 / fact;
  double *r = s->rLui;
  double tb = r[1] * (g[1] * r[1] + g[4] * r[2] + g[6] * r[3]) + r[2] * (g[4] * r[1] + g[2] * r[2] + g[5] * r[3]) +
              r[3] * (g[6] * r[1] + g[5] * r[2] + g[3] * r[3]);
  if (tb != 0.0) tb = s->tauM / sqrt(tb);
  s->tauBar = tb;
  s->uBar[1] = u1 - s->tauM * r[1] * rhoinv;
  s->uBar[2] = u2 - s->tauM * r[2] * rhoinv;
  s->uBar[3] = u3 - s->tauM * r[3] * rhoinv;
}

/* e3Res (incompressible/e3res.f:1-200), iLES=0, no rotation */
static void e3res_i(const orc_incomp *ip, const iqp *s, double rl[][5]) {
  double rNa[4], rGNa[4][4];
  double tmps = 1.0 - ip->flmpr;
  const double *g1 = s->g1yi, *g2 = s->g2yi, *g3 = s->g3yi;
  double u1 = s->u1, u2 = s->u2, u3 = s->u3, rmu = s->rmu, rho = s->rho;
  for (int i = 1; i <= 3; i++) rNa[i] = s->aci[i] * tmps - s->src[i];
  double tmp = -s->pres + s->tauC * (g1[2] + g2[3] + g3[4]);
  double tmp1 = rmu * (g2[2] + g1[3]);
  double tmp2 = rmu * (g3[3] + g2[4]);
  double tmp3 = rmu * (g1[4] + g3[2]);
  if (ip->iconvflow == 2) {
    rNa[1] = rNa[1] + s->uBar[1] * g1[2] + s->uBar[2] * g2[2] + s->uBar[3] * g3[2];
    rNa[2] = rNa[2] + s->uBar[1] * g1[3] + s->uBar[2] * g2[3] + s->uBar[3] * g3[3];
    rNa[3] = rNa[3] + s->uBar[1] * g1[4] + s->uBar[2] * g2[4] + s->uBar[3] * g3[4];
    rGNa[1][1] = 2.0 * rmu * g1[2] + tmp;
    rGNa[1][2] = tmp1;
    rGNa[1][3] = tmp3;
    rGNa[2][1] = tmp1;
    rGNa[2][2] = 2.0 * rmu * g2[3] + tmp;
    rGNa[2][3] = tmp2;
    rGNa[3][1] = tmp3;
    rGNa[3][2] = tmp2;
    rGNa[3][3] = 2.0 * rmu * g3[4] + tmp;
  } else {
    rGNa[1][1] = 2.0 * rmu * g1[2] + tmp - u1 * u1 * rho;
    rGNa[1][2] = tmp1 - u1 * u2 * rho;
    rGNa[1][3] = tmp3 - u1 * u3 * rho;
    rGNa[2][1] = tmp1 - u1 * u2 * rho;
    rGNa[2][2] = 2.0 * rmu * g2[3] + tmp - u2 * u2 * rho;
    rGNa[2][3] = tmp2 - u3 * u2 * rho;
    rGNa[3][1] = tmp3 - u1 * u3 * rho;
    rGNa[3][2] = tmp2 - u3 * u2 * rho;
    rGNa[3][3] = 2.0 * rmu * g3[4] + tmp - u3 * u3 * rho;
  }
  tmp1 = s->tauM * s->rLui[1];
  tmp2 = s->tauM * s->rLui[2];
  tmp3 = s->tauM * s->rLui[3];
  double t[4] = {0, tmp1, tmp2, tmp3}, u[4] = {0, u1, u2, u3};
  for (int i = 1; i <= 3; i++)
    for (int j = 1; j <= 3; j++) rGNa[i][j] = rGNa[i][j] + t[i] * u[j];
  if (ip->iconvflow == 1)
    for (int i = 1; i <= 3; i++)
      for (int j = 1; j <= 3; j++) rGNa[i][j] = rGNa[i][j] + t[j] * u[i];
  if (ip->iconvflow == 2) {
    const double *r = s->rLui;
    t[1] = s->tauBar * (r[1] * g1[2] + r[2] * g2[2] + r[3] * g3[2]);
    t[2] = s->tauBar * (r[1] * g1[3] + r[2] * g2[3] + r[3] * g3[3]);
    t[3] = s->tauBar * (r[1] * g1[4] + r[2] * g2[4] + r[3] * g3[4]);
    for (int i = 1; i <= 3; i++)
      for (int j = 1; j <= 3; j++) rGNa[i][j] = rGNa[i][j] + t[i] * r[j];
  }
  for (int i = 1; i <= 3; i++) rNa[i] = rNa[i] * rho;
  double W = s->WdetJ;
  for (int a = 1; a <= s->nshl; a++) {
    const double *sg = s->shg[a];
    rl[a][4] = rl[a][4] + W * (sg[1] * s->uBar[1] + sg[2] * s->uBar[2] + sg[3] * s->uBar[3]);
    for (int i = 1; i <= 3; i++)
      rl[a][i] = rl[a][i] - W * (s->shape[a] * rNa[i] + sg[1] * rGNa[i][1] + sg[2] * rGNa[i][2] + sg[3] * rGNa[i][3]);
  }
}

/* e3LHS (incompressible/e3lhs.f:1-230).  K[k][a][b] = xKebe(:,k,a,b), k=3(row-1)+col;
 * G[k][a][b] = xGoC(:,k,a,b). */
#define XK(k, a, b) K[((k)-1) * nn + ((a)-1) * nshl + ((b)-1)]
#define XG(k, a, b) G[((k)-1) * nn + ((a)-1) * nshl + ((b)-1)]
static void e3lhs_i(const orc_incomp *ip, iqp *s, double *K, double *G) {
  int nshl = s->nshl, nn = nshl * nshl;
  double lhsFct = ip->alfi * ip->gami * ip->Delt;
  double lhmFct = ip->almi * (1.0 - ip->flmpl);
  double W = s->WdetJ, rho = s->rho;
  double tlW = lhsFct * W;
  double tmp1 = tlW * rho;
  double tauM = tlW * s->tauM, tauC = tlW * s->tauC, rmu = tlW * s->rmu;
  double tsFct = lhmFct * W * rho;
  double tauBar, uB[4];
  double u[4] = {0, s->u1, s->u2, s->u3};
  if (ip->iconvflow == 2) {
    tauBar = lhsFct * W * s->tauBar;
    for (int i = 1; i <= 3; i++) uB[i] = tmp1 * s->uBar[i];
  } else {
    tauBar = 0.0;
    for (int i = 1; i <= 3; i++) uB[i] = tmp1 * u[i];
  }
  for (int b = 1; b <= nshl; b++) {
    double t1 = uB[1] * s->shg[b][1] + uB[2] * s->shg[b][2] + uB[3] * s->shg[b][3];
    for (int a = 1; a <= nshl; a++) {
      double x1 = tsFct * s->shape[a] * s->shape[b];
      double x2 = x1 + t1 * s->shape[a];
      XK(1, a, b) += x2;
      XK(5, a, b) += x2;
      XK(9, a, b) += x2;
    }
  }
  const double *r = s->rLui;
  for (int b = 1; b <= nshl; b++) {
    const double *sb = s->shg[b];
    double t1[4], t2[4], t3[4];
    for (int i = 1; i <= 3; i++) {
      t1[i] = tauC * sb[i];
      t2[i] = rmu * sb[i];
    }
    double y1 = tauM * (u[1] * sb[1] + u[2] * sb[2] + u[3] * sb[3]) * rho;
    double y2 = tauBar * (r[1] * sb[1] + r[2] * sb[2] + r[3] * sb[3]);
    for (int i = 1; i <= 3; i++) t3[i] = t2[i] + y1 * u[i] + y2 * r[i];
    int a = b;
    const double *sa = s->shg[a];
    double tmp = t3[1] * sa[1] + t3[2] * sa[2] + t3[3] * sa[3];
    XK(1, a, b) = XK(1, a, b) + tmp + t1[1] * sa[1] + t2[1] * sa[1];
    XK(5, a, b) = XK(5, a, b) + tmp + t1[2] * sa[2] + t2[2] * sa[2];
    XK(9, a, b) = XK(9, a, b) + tmp + t1[3] * sa[3] + t2[3] * sa[3];
    double z;
    z = t1[1] * sa[2] + t2[2] * sa[1];
    XK(2, a, b) += z;
    XK(4, b, a) += z;
    z = t1[1] * sa[3] + t2[3] * sa[1];
    XK(3, a, b) += z;
    XK(7, b, a) += z;
    z = t1[2] * sa[3] + t2[3] * sa[2];
    XK(6, a, b) += z;
    XK(8, b, a) += z;
    for (a = b + 1; a <= nshl; a++) {
      sa = s->shg[a];
      tmp = t3[1] * sa[1] + t3[2] * sa[2] + t3[3] * sa[3];
      z = tmp + t1[1] * sa[1] + t2[1] * sa[1];
      XK(1, a, b) += z;
      XK(1, b, a) += z;
      z = tmp + t1[2] * sa[2] + t2[2] * sa[2];
      XK(5, a, b) += z;
      XK(5, b, a) += z;
      z = tmp + t1[3] * sa[3] + t2[3] * sa[3];
      XK(9, a, b) += z;
      XK(9, b, a) += z;
      z = t1[1] * sa[2] + t2[2] * sa[1];
      XK(2, a, b) += z;
      XK(4, b, a) += z;
      z = t1[1] * sa[3] + t2[3] * sa[1];
      XK(3, a, b) += z;
      XK(7, b, a) += z;
      z = t1[2] * sa[1] + t2[1] * sa[2];
      XK(4, a, b) += z;
      XK(2, b, a) += z;
      z = t1[2] * sa[3] + t2[3] * sa[2];
      XK(6, a, b) += z;
      XK(8, b, a) += z;
      z = t1[3] * sa[1] + t2[1] * sa[3];
      XK(7, a, b) += z;
      XK(3, b, a) += z;
      z = t1[3] * sa[2] + t2[2] * sa[3];
      XK(8, a, b) += z;
      XK(6, b, a) += z;
    }
  }
  for (int b = 1; b <= nshl; b++) {
    double t1[4];
    for (int i = 1; i <= 3; i++) t1[i] = tlW * s->shg[b][i];
    for (int a = 1; a <= nshl; a++)
      for (int i = 1; i <= 3; i++) XG(i, a, b) += t1[i] * s->shape[a];
  }
  tauM = tauM / rho;
  for (int b = 1; b <= nshl; b++) {
    double t1[4];
    for (int i = 1; i <= 3; i++) t1[i] = tauM * s->shg[b][i];
    for (int a = b; a <= nshl; a++)
      XG(4, a, b) = XG(4, a, b) + t1[1] * s->shg[a][1] + t1[2] * s->shg[a][2] + t1[3] * s->shg[a][3];
  }
}

/* bc3LHS (incompressible/bc3lhs.f:1-380) on one element's xKebe: node-sequential column
 * then row eliminations; velocity codes 0 and 7 leave the block alone (:13-14).  The
 * second row operation of code 6 omits the BC(:,6) term on two of its three statements
 * (:355-361) -- kept. */
static void bc3lhs_i(const orc_part *p, const int *nodes, int nshl, double *K) {
  int nn = nshl * nshl, nshg = p->c.nshg;
  for (int inod = 1; inod <= nshl; inod++) {
    int in = nodes[inod - 1];
    int code = (p->iBC[in] >> 3) & 7;
    if (code == 0 || code == 7) continue;
    double b4 = p->BC[in + (size_t)nshg * 3], b5 = p->BC[in + (size_t)nshg * 4], b6 = p->BC[in + (size_t)nshg * 5];
    if (code == 1 || code == 2 || code == 4) {
      int pv = code == 1 ? 1 : (code == 2 ? 2 : 3);
      int o1 = pv == 1 ? 2 : 1, o2 = pv == 3 ? 2 : 3;
      for (int r = 1; r <= 3; r++)
        for (int i = 1; i <= nshl; i++)
          XK(3 * (r - 1) + o1, i, inod) = XK(3 * (r - 1) + o1, i, inod) - b4 * XK(3 * (r - 1) + pv, i, inod);
      for (int r = 1; r <= 3; r++)
        for (int i = 1; i <= nshl; i++)
          XK(3 * (r - 1) + o2, i, inod) = XK(3 * (r - 1) + o2, i, inod) - b5 * XK(3 * (r - 1) + pv, i, inod);
      for (int r = 1; r <= 3; r++)
        for (int i = 1; i <= nshl; i++) XK(3 * (r - 1) + pv, i, inod) = 0.0;
      for (int c = 1; c <= 3; c++)
        for (int i = 1; i <= nshl; i++)
          XK(3 * (o1 - 1) + c, inod, i) = XK(3 * (o1 - 1) + c, inod, i) - b4 * XK(3 * (pv - 1) + c, inod, i);
      for (int c = 1; c <= 3; c++)
        for (int i = 1; i <= nshl; i++)
          XK(3 * (o2 - 1) + c, inod, i) = XK(3 * (o2 - 1) + c, inod, i) - b5 * XK(3 * (pv - 1) + c, inod, i);
      for (int c = 1; c <= 3; c++)
        for (int i = 1; i <= nshl; i++) XK(3 * (pv - 1) + c, inod, i) = 0.0;
      XK(3 * (pv - 1) + pv, inod, inod) = 1.0;
    } else {
      int p1, p2, fr;
      if (code == 3) { p1 = 1; p2 = 2; fr = 3; }
      else if (code == 5) { p1 = 1; p2 = 3; fr = 2; }
      else { p1 = 2; p2 = 3; fr = 1; }
      for (int r = 1; r <= 3; r++)
        for (int i = 1; i <= nshl; i++)
          XK(3 * (r - 1) + fr, i, inod) = XK(3 * (r - 1) + fr, i, inod) - b4 * XK(3 * (r - 1) + p1, i, inod) -
                                          b6 * XK(3 * (r - 1) + p2, i, inod);
      for (int r = 1; r <= 3; r++)
        for (int i = 1; i <= nshl; i++) XK(3 * (r - 1) + p1, i, inod) = XK(3 * (r - 1) + p2, i, inod) = 0.0;
      for (int c = 1; c <= 3; c++)
        for (int i = 1; i <= nshl; i++) {
          double v = XK(3 * (fr - 1) + c, inod, i) - b4 * XK(3 * (p1 - 1) + c, inod, i);
          if (!(code == 6 && c < 3)) v = v - b6 * XK(3 * (p2 - 1) + c, inod, i);
          XK(3 * (fr - 1) + c, inod, i) = v;
        }
      for (int c = 1; c <= 3; c++)
        for (int i = 1; i <= nshl; i++) XK(3 * (p1 - 1) + c, inod, i) = XK(3 * (p2 - 1) + c, inod, i) = 0.0;
      XK(3 * (p1 - 1) + p1, inod, inod) = 1.0;
      XK(3 * (p2 - 1) + p2, inod, inod) = 1.0;
    }
  }
}

/* sparseloc (common/fillsparse.f:236-271), 1-based result */
static int sparseloc_i(const int *list, int n, int target) {
  int rowvl = 1, rowvh = n + 1;
  while (rowvh - rowvl > 1) {
    int rowv = (rowvh + rowvl) / 2;
    if (list[rowv - 1] > target) rowvh = rowv; else rowvl = rowv;
  }
  return rowvl;
}

static void gather_i(const orc_part *p, const int *ien, int npro, int e, int nshl, double yl[][6], double acl[][6],
                     double xl[][4], double ql[][10], int idflx) {
  const orc_common *c = &p->c;
  int nshg = c->nshg;
  static const int src[6] = {0, 3, 0, 1, 2, 4}; /* localy.f:47-72 */
  for (int n = 1; n <= nshl; n++) {
    int A = abs(ien[e + (size_t)npro * (n - 1)]) - 1;
    for (int m = 1; m <= 5; m++) {
      yl[n][m] = p->y[A + (size_t)nshg * src[m]];
      if (acl) acl[n][m] = p->ac[A + (size_t)nshg * src[m]];
    }
    for (int i = 1; i <= 3; i++) xl[n][i] = p->x[A + (size_t)c->numnp * (i - 1)];
    if (ql) {
      memset(ql[n], 0, sizeof(double) * 10);
      for (int k = 1; k <= idflx; k++) ql[n][k] = p->qres[A + (size_t)nshg * (k - 1)];
    }
  }
}

/* AsIq + e3q + e3qvar for one block (idiff = 1) */
static void asiq_i(const orc_part *p, const orc_incomp *ip, int iblk) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nshl = lc[9], npro = lc[10] - iel;
  int ngauss = c->nint[lcsyst - 1], nshg = c->nshg;
  const int *ien = p->ien + p->ien_off[iblk];
  double(*ql)[ORC_MAXSH + 1][10] = calloc((size_t)npro, sizeof *ql);
  double(*rm)[ORC_MAXSH + 1] = calloc((size_t)npro, sizeof *rm);
  for (int e = 0; e < npro; e++) {
    double yl[ORC_MAXSH + 1][6], xl[ORC_MAXSH + 1][4];
    gather_i(p, ien, npro, e, nshl, yl, NULL, xl, NULL, 0);
    iqp s;
    s.nshl = nshl;
    s.nenl = nenl;
    s.lcsyst = lcsyst;
    for (int intp = 1; intp <= ngauss; intp++) {
      if (QWT(c, lcsyst, intp) == 0.0) continue;
      s.intp = intp;
      getshp_i(p, &s);
      metric_i(c, &s, xl);
      double g1[5] = {0}, g2[5] = {0}, g3[5] = {0}, qdi[10];
      for (int n = 1; n <= nshl; n++)
        for (int m = 2; m <= 4; m++) {
          g1[m] += s.shg[n][1] * yl[n][m];
          g2[m] += s.shg[n][2] * yl[n][m];
          g3[m] += s.shg[n][3] * yl[n][m];
        }
      double rmu = ip->rmu;
      qdi[1] = 2.0 * rmu * g1[2];
      qdi[4] = rmu * (g1[3] + g2[2]);
      qdi[7] = rmu * (g1[4] + g3[2]);
      qdi[2] = rmu * (g1[3] + g2[2]);
      qdi[5] = 2.0 * rmu * g2[3];
      qdi[8] = rmu * (g2[4] + g3[3]);
      qdi[3] = rmu * (g1[4] + g3[2]);
      qdi[6] = rmu * (g2[4] + g3[3]);
      qdi[9] = 2.0 * rmu * g3[4];
      for (int i = 1; i <= nshl; i++) {
        for (int k = 1; k <= 9; k++) ql[e][i][k] = ql[e][i][k] + s.shape[i] * s.WdetJ * qdi[k];
        rm[e][i] = rm[e][i] + s.shape[i] * s.WdetJ;
      }
    }
  }
  for (int k = 1; k <= 9; k++)
    for (int i = 1; i <= nshl; i++)
      for (int e = 0; e < npro; e++) {
        int A = abs(ien[e + (size_t)npro * (i - 1)]) - 1;
        p->qres[A + (size_t)nshg * (k - 1)] += ql[e][i][k];
      }
  for (int i = 1; i <= nshl; i++)
    for (int e = 0; e < npro; e++) {
      int A = abs(ien[e + (size_t)npro * (i - 1)]) - 1;
      p->rmass[A] += rm[e][i];
    }
  free(ql);
  free(rm);
}

/* qpbc (common/qpbc.f:1-90) for idflx = (nflow-1)*nsd = 9 */
static void qpbc_i(int nparts, orc_part *parts) {
  double **q = malloc(sizeof(double *) * nparts), **rm = malloc(sizeof(double *) * nparts);
  for (int m = 0; m < nparts; m++) {
    q[m] = parts[m].qres;
    rm[m] = parts[m].rmass;
  }
  orc_commu(nparts, parts, q, 9, 0);
  orc_commu(nparts, parts, rm, 1, 0);
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    int nshg = p->c.nshg;
    for (int j = 0; j < nshg; j++)
      if (p->iBC[j] & (1 << 10)) {
        int i = p->iper[j] - 1;
        p->rmass[i] += p->rmass[j];
        for (int k = 0; k < 9; k++) p->qres[i + (size_t)nshg * k] += p->qres[j + (size_t)nshg * k];
      }
    for (int j = 0; j < nshg; j++)
      if (p->iBC[j] & (1 << 10)) {
        int i = p->iper[j] - 1;
        p->rmass[j] = p->rmass[i];
        for (int k = 0; k < 9; k++) p->qres[j + (size_t)nshg * k] = p->qres[i + (size_t)nshg * k];
      }
    for (int j = 0; j < nshg; j++) p->rmass[j] = 1.0 / p->rmass[j];
    for (int k = 0; k < 9; k++)
      for (int j = 0; j < nshg; j++) p->qres[j + (size_t)nshg * k] = p->rmass[j] * p->qres[j + (size_t)nshg * k];
  }
  orc_commu(nparts, parts, q, 9, 1);
  free(q);
  free(rm);
}

/* AsIGMR + e3 + bc3LHS + fillsparseI for one block */
static void asigmr_i(const orc_part *p, const orc_incomp *ip, int iblk, double *res, double *lhsK, double *lhsP,
                     double *xKebe_out, double *xGoC_out) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nshl = lc[9], npro = lc[10] - iel;
  int ngauss = c->nint[lcsyst - 1], nshg = c->nshg, nn = nshl * nshl;
  int idflx = ip->idiff >= 1 ? 9 : 0;
  const int *ien = p->ien + p->ien_off[iblk];
  double(*rl)[ORC_MAXSH + 1][5] = calloc((size_t)npro, sizeof *rl);
  double *K = malloc(sizeof(double) * 9 * nn), *G = malloc(sizeof(double) * 4 * nn);
  for (int e = 0; e < npro; e++) {
    double yl[ORC_MAXSH + 1][6], acl[ORC_MAXSH + 1][6], xl[ORC_MAXSH + 1][4], ql[ORC_MAXSH + 1][10];
    gather_i(p, ien, npro, e, nshl, yl, acl, xl, ql, idflx);
    memset(K, 0, sizeof(double) * 9 * nn);
    memset(G, 0, sizeof(double) * 4 * nn);
    iqp s;
    s.nshl = nshl;
    s.nenl = nenl;
    s.lcsyst = lcsyst;
    for (int intp = 1; intp <= ngauss; intp++) {
      if (QWT(c, lcsyst, intp) == 0.0) continue;
      s.intp = intp;
      getshp_i(p, &s);
      s.rho = ip->rho;
      s.rmu = ip->rmu;
      e3ivar_i(c, ip, &s, yl, acl, xl, ql);
      e3stab_i(ip, &s);
      e3res_i(ip, &s, rl[e]);
      if (ip->lhs == 1) e3lhs_i(ip, &s, K, G);
    }
    if (ip->lhs == 1) {
      for (int ib = 1; ib <= nshl; ib++)
        for (int ia = 1; ia <= ib - 1; ia++) XG(4, ia, ib) = XG(4, ib, ia); /* e3.f:108-113 */
      int nodes[ORC_MAXSH];
      for (int n = 0; n < nshl; n++) nodes[n] = abs(ien[e + (size_t)npro * n]) - 1;
      if (ip->ipord == 1) bc3lhs_i(p, nodes, nshl, K);
      if (xKebe_out) {
        /* xKebe(numel,9,nshl,nshl), xGoC(numel,4,nshl,nshl) over the whole part, for the parity tests */
        size_t numel = (size_t)c->numel, ge = (size_t)(iel - 1 + e);
        int nsm = c->nshape;
        for (int k = 1; k <= 9; k++)
          for (int a = 1; a <= nshl; a++)
            for (int b = 1; b <= nshl; b++)
              xKebe_out[ge + numel * ((k - 1) + 9 * ((a - 1) + (size_t)nsm * (b - 1)))] = XK(k, a, b);
        for (int k = 1; k <= 4; k++)
          for (int a = 1; a <= nshl; a++)
            for (int b = 1; b <= nshl; b++)
              xGoC_out[ge + numel * ((k - 1) + 4 * ((a - 1) + (size_t)nsm * (b - 1)))] = XG(k, a, b);
      }
      /* fillsparseI (common/fillsparse.f:1-65) */
      for (int a = 1; a <= nshl; a++) {
        int i = nodes[a - 1];
        int cs = p->colm[i], n = p->colm[i + 1] - cs;
        for (int b = 1; b <= nshl; b++) {
          int k = sparseloc_i(p->rowp + (cs - 1), n, nodes[b - 1] + 1) + cs - 1;
          for (int m = 1; m <= 9; m++) lhsK[(m - 1) + 9 * (size_t)(k - 1)] += XK(m, a, b);
          for (int m = 1; m <= 4; m++) lhsP[(m - 1) + 4 * (size_t)(k - 1)] += XG(m, a, b);
        }
      }
    }
  }
  /* local(res, rl, ien, nflow, 'scatter') (common/local.f:67-74) */
  for (int j = 1; j <= 4; j++)
    for (int i = 1; i <= nshl; i++)
      for (int e = 0; e < npro; e++) {
        int A = abs(ien[e + (size_t)npro * (i - 1)]) - 1;
        res[A + (size_t)nshg * (j - 1)] += rl[e][i][j];
      }
  free(rl);
  free(K);
  free(G);
}

/* bc3per (incompressible/bc3per.f:1-45): periodic sum, then zero the rows another part owns */
static void bc3per_i(const orc_part *p, double *r, int nQs) {
  int nshg = p->c.nshg;
  for (int j = 0; j < nshg; j++)
    if ((p->iBC[j] & (1 << 10)) || (p->iBC[j] & (1 << 12))) {
      int i = p->iper[j] - 1;
      for (int k = 0; k < nQs; k++) {
        r[i + (size_t)nshg * k] += r[j + (size_t)nshg * k];
        r[j + (size_t)nshg * k] = 0.0;
      }
    }
  if (p->c.numpe > 1) {
    const int *il = p->ilwork;
    int numtask = il[0], itkbeg = 1;
    for (int it = 0; it < numtask; it++) {
      int iacc = il[itkbeg + 1], numseg = il[itkbeg + 3];
      if (iacc == 0)
        for (int is = 1; is <= numseg; is++) {
          int isgbeg = il[itkbeg + 2 + 2 * is], lenseg = il[itkbeg + 3 + 2 * is];
          for (int k = 0; k < nQs; k++)
            for (int a = isgbeg - 1; a < isgbeg - 1 + lenseg; a++) r[a + (size_t)nshg * k] = 0.0;
        }
      itkbeg += 4 + 2 * numseg;
    }
  }
}

/* bc3Res (incompressible/bc3res.f:1-90), intpres = 0 */
static void bc3res_i(const orc_part *p, double *res) {
  int nshg = p->c.nshg;
  bc3per_i(p, res, 4);
#define RS(i, k) res[(i) + (size_t)nshg * ((k)-1)]
#define BCv(i, k) p->BC[(i) + (size_t)nshg * ((k)-1)]
  for (int i = 0; i < nshg; i++) {
    int ib = p->iBC[i], code = (ib >> 3) & 7;
    if (ib & 4) RS(i, 4) = 0.0;
    switch (code) {
      case 1:
        RS(i, 2) = RS(i, 2) - BCv(i, 4) * RS(i, 1);
        RS(i, 3) = RS(i, 3) - BCv(i, 5) * RS(i, 1);
        RS(i, 1) = 0.0;
        break;
      case 2:
        RS(i, 1) = RS(i, 1) - BCv(i, 4) * RS(i, 2);
        RS(i, 3) = RS(i, 3) - BCv(i, 5) * RS(i, 2);
        RS(i, 2) = 0.0;
        break;
      case 3:
        RS(i, 3) = RS(i, 3) - BCv(i, 4) * RS(i, 1) - BCv(i, 6) * RS(i, 2);
        RS(i, 1) = RS(i, 2) = 0.0;
        break;
      case 4:
        RS(i, 1) = RS(i, 1) - BCv(i, 4) * RS(i, 3);
        RS(i, 2) = RS(i, 2) - BCv(i, 5) * RS(i, 3);
        RS(i, 3) = 0.0;
        break;
      case 5:
        RS(i, 2) = RS(i, 2) - BCv(i, 4) * RS(i, 1) - BCv(i, 6) * RS(i, 3);
        RS(i, 1) = RS(i, 3) = 0.0;
        break;
      case 6:
        RS(i, 1) = RS(i, 1) - BCv(i, 4) * RS(i, 2) - BCv(i, 6) * RS(i, 3);
        RS(i, 2) = RS(i, 3) = 0.0;
        break;
      case 7:
        RS(i, 1) = RS(i, 2) = RS(i, 3) = 0.0;
        break;
      default:
        break;
    }
    if (ib & (1 << 11)) RS(i, 1) = RS(i, 2) = RS(i, 3) = 0.0;
  }
#undef RS
#undef BCv
}

/* AsBMFG + e3b + e3bvar of the incompressible code (incompressible/asbmfg.f:1-68, e3b.f:1-262, e3bvar.f:1-230) for one
 * boundary block, rigid walls (ideformwall = 0: vdot = rlKwall = 0, xKebe = 0 so the bc3lhs / fillsparseI that follow
 * in elmgmr.f:303-313 add nothing).  getbnodes lnode (hierarchic.f:90-190); normals "curl into element for tets, all
 * others out" (e3bvar.f:100-127), WdetJb per topology (e3bvar.f:129-142).  flxID / Force go to p->aerfrc as in the
 * compressible path: Force(1:3) at [0..2], flxID(k,iface) at [4 + 10 iface + k-1]. */
static void asbmfg_i(const orc_part *p, const orc_incomp *ip, int iblk, double *res) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblkb + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nenbl = lc[5], nshl = lc[8], nshlb = lc[9];
  int npro = lc[10] - iel, nshg = c->nshg;
  if (lcsyst == 3) lcsyst = nenbl; /* elmgmr.f:267 */
  int lnode[5] = {0, 1, 2, 3, 4}, ipt2, ipt3;
  if (lcsyst == 4) {
    lnode[2] = 4;
    lnode[3] = 5;
    lnode[4] = 2;
  }
  if (lcsyst == 1) {
    ipt2 = 2;
    ipt3 = 3;
  } else if (lcsyst == 2) {
    ipt2 = 4;
    ipt3 = 2;
  } else if (lcsyst == 3) {
    ipt2 = 3;
    ipt3 = 2;
  } else if (lcsyst == 4) {
    ipt2 = 2;
    ipt3 = 4;
  } else {
    fprintf(stderr, "orc_inc_elmgmr: boundary block lcsyst %d not restated\n", lcsyst);
    abort();
  }
  int ngaussb = c->nintb[lcsyst - 1];
  const int *ien = p->ienb + p->ienb_off[iblk];
  const int *iBCB = p->iBCB + p->iBCB_off[iblk];
  const double *BCB = p->BCB + p->BCB_off[iblk];
  double(*rl)[9][5] = calloc((size_t)npro, sizeof *rl);
  for (int intp = 1; intp <= ngaussb; intp++) {
    double sum1 = 0, sum2 = 0, sum3 = 0; /* Force uses sum() over the block (e3b.f:236-238) */
    for (int e = 0; e < npro; e++) {
      double yl[9][5], xlb[9][4];
      for (int n = 1; n <= nshl; n++) {
        int A = abs(ien[e + (size_t)npro * (n - 1)]) - 1;
        yl[n][1] = p->y[A + (size_t)nshg * 3];
        yl[n][2] = p->y[A + (size_t)nshg * 0];
        yl[n][3] = p->y[A + (size_t)nshg * 1];
        yl[n][4] = p->y[A + (size_t)nshg * 2];
        for (int i = 1; i <= 3; i++) xlb[n][i] = p->x[A + (size_t)c->numnp * (i - 1)];
      }
      double shape[9], shdrv[4][9];
      for (int n = 1; n <= nshl; n++) {
        shape[n] = SHPB(p, lcsyst, n, intp);
        for (int i = 1; i <= 3; i++) shdrv[i][n] = SHGLB(p, lcsyst, i, n, intp);
      }
      double rmu = ip->rmu, rho = ip->rho;
      double pres = 0, u1 = 0, u2 = 0, u3 = 0;
      for (int k = 1; k <= nshlb; k++) {
        int n = lnode[k];
        pres += shape[n] * yl[n][1];
        u1 += shape[n] * yl[n][2];
        u2 += shape[n] * yl[n][3];
        u3 += shape[n] * yl[n][4];
      }
      double dxdxib[4][4];
      memset(dxdxib, 0, sizeof dxdxib);
      for (int n = 1; n <= nenl; n++)
        for (int i = 1; i <= 3; i++)
          for (int j = 1; j <= 3; j++) dxdxib[i][j] += xlb[n][i] * shdrv[j][n];
      double v1[4], v2[4];
      for (int i = 1; i <= 3; i++) {
        v1[i] = xlb[ipt2][i] - xlb[1][i];
        v2[i] = xlb[ipt3][i] - xlb[1][i];
      }
      double t1 = v1[2] * v2[3] - v2[2] * v1[3];
      double t2 = v2[1] * v1[3] - v1[1] * v2[3];
      double t3 = v1[1] * v2[2] - v2[1] * v1[2];
      double temp = 1.0 / sqrt(t1 * t1 + t2 * t2 + t3 * t3);
      double bn[4] = {0, t1 * temp, t2 * temp, t3 * temp};
      double WdetJb = (lcsyst == 3) ? QWTB(c, lcsyst, intp) / (2.0 * temp) : QWTB(c, lcsyst, intp) / (4.0 * temp);
      double d[4][4];
      d[1][1] = dxdxib[2][2] * dxdxib[3][3] - dxdxib[3][2] * dxdxib[2][3];
      d[1][2] = dxdxib[3][2] * dxdxib[1][3] - dxdxib[1][2] * dxdxib[3][3];
      d[1][3] = dxdxib[1][2] * dxdxib[2][3] - dxdxib[1][3] * dxdxib[2][2];
      temp = 1.0 / (d[1][1] * dxdxib[1][1] + d[1][2] * dxdxib[2][1] + d[1][3] * dxdxib[3][1]);
      d[1][1] *= temp;
      d[1][2] *= temp;
      d[1][3] *= temp;
      d[2][1] = (dxdxib[2][3] * dxdxib[3][1] - dxdxib[2][1] * dxdxib[3][3]) * temp;
      d[2][2] = (dxdxib[1][1] * dxdxib[3][3] - dxdxib[3][1] * dxdxib[1][3]) * temp;
      d[2][3] = (dxdxib[2][1] * dxdxib[1][3] - dxdxib[1][1] * dxdxib[2][3]) * temp;
      d[3][1] = (dxdxib[2][1] * dxdxib[3][2] - dxdxib[2][2] * dxdxib[3][1]) * temp;
      d[3][2] = (dxdxib[3][1] * dxdxib[1][2] - dxdxib[1][1] * dxdxib[3][2]) * temp;
      d[3][3] = (dxdxib[1][1] * dxdxib[2][2] - dxdxib[1][2] * dxdxib[2][1]) * temp;
      double gl[4][5];
      memset(gl, 0, sizeof gl);
      for (int n = 1; n <= nshl; n++)
        for (int i = 1; i <= 3; i++)
          for (int m = 1; m <= 4; m++) gl[i][m] += shdrv[i][n] * yl[n][m];
      double g1[5], g2[5], g3[5];
      for (int m = 2; m <= 4; m++) {
        g1[m] = d[1][1] * gl[1][m] + d[2][1] * gl[2][m] + d[3][1] * gl[3][m];
        g2[m] = d[1][2] * gl[1][m] + d[2][2] * gl[2][m] + d[3][2] * gl[3][m];
        g3[m] = d[1][3] * gl[1][m] + d[2][3] * gl[2][m] + d[3][3] * gl[3][m];
      }
      double unm = bn[1] * u1 + bn[2] * u2 + bn[3] * u3;
      double tau1n = bn[1] * 2.0 * rmu * g1[2] + bn[2] * (rmu * (g2[2] + g1[3])) + bn[3] * (rmu * (g3[2] + g1[4]));
      double tau2n = bn[1] * (rmu * (g2[2] + g1[3])) + bn[2] * 2.0 * rmu * g2[3] + bn[3] * (rmu * (g3[3] + g2[4]));
      double tau3n = bn[1] * (rmu * (g3[2] + g1[4])) + bn[2] * (rmu * (g3[3] + g2[4])) + bn[3] * 2.0 * rmu * g3[4];
      double tn = bn[1] * tau1n + bn[2] * tau2n + bn[3] * tau3n;
      pres = pres - tn;
      tau1n = tau1n - bn[1] * tn;
      tau2n = tau2n - bn[2] * tn;
      tau3n = tau3n - bn[3] * tn;
      tau1n = tau1n * ip->iviscflux;
      tau2n = tau2n * ip->iviscflux;
      tau3n = tau3n * ip->iviscflux;
      /* ---- e3b.f:53-118 ---- */
      int ibcb = iBCB[e], iface = abs(iBCB[e + (size_t)npro]);
      int listed = ip->nsrflist && iface <= 1000 && ip->nsrflist[iface] != 0;
      if (listed && p->aerfrc) {
        double *fl = p->aerfrc + 4 + 10 * iface;
        fl[0] = fl[0] + WdetJb;
        fl[1] = fl[1] - WdetJb * unm;
        fl[2] = fl[2] - (tau1n - bn[1] * pres) * WdetJb;
        fl[3] = fl[3] - (tau2n - bn[2] * pres) * WdetJb;
        fl[4] = fl[4] - (tau3n - bn[3] * pres) * WdetJb;
      }
#define BCBv(n, k) BCB[e + (size_t)npro * (((n)-1) + (size_t)nshlb * ((k)-1))]
      if (ibcb & 1) {
        unm = 0;
        for (int n = 1; n <= nshlb; n++) unm = unm + shape[lnode[n]] * BCBv(n, 1);
      }
      if (ibcb & 2) {
        pres = 0;
        for (int n = 1; n <= nshlb; n++) pres = pres + shape[lnode[n]] * BCBv(n, 2);
      }
      if (ibcb & 4) {
        tau1n = tau2n = tau3n = 0;
        for (int n = 1; n <= nshlb; n++) {
          tau1n = tau1n + shape[lnode[n]] * BCBv(n, 3);
          tau2n = tau2n + shape[lnode[n]] * BCBv(n, 4);
          tau3n = tau3n + shape[lnode[n]] * BCBv(n, 5);
        }
      }
#undef BCBv
      if (ibcb & 16) {
        fprintf(stderr, "orc_inc_elmgmr: deformable-wall boundary elements (iBCB bit 4) are not restated\n");
        abort();
      }
      double rNa[5];
      rNa[1] = -WdetJb * (tau1n - bn[1] * pres - 0.0);
      rNa[2] = -WdetJb * (tau2n - bn[2] * pres - 0.0);
      rNa[3] = -WdetJb * (tau3n - bn[3] * pres - 0.0);
      rNa[4] = WdetJb * unm;
      if (ip->iconvflow == 1) {
        double rou = rho * unm;
        rNa[1] = rNa[1] + WdetJb * rou * u1;
        rNa[2] = rNa[2] + WdetJb * rou * u2;
        rNa[3] = rNa[3] + WdetJb * rou * u3;
      }
      for (int k = 1; k <= nshlb; k++) {
        int n = lnode[k];
        for (int m = 1; m <= 4; m++) rl[e][n][m] = rl[e][n][m] - shape[n] * rNa[m];
      }
      if (abs(ip->itwmod) == 1 && listed) { /* e3b.f:224-240, ires != 2, iter == nitr; nsrflist == 1 */
        if (ip->nsrflist[iface] == 1) {
          sum1 += (tau1n - bn[1] * pres) * WdetJb;
          sum2 += (tau2n - bn[2] * pres) * WdetJb;
          sum3 += (tau3n - bn[3] * pres) * WdetJb;
        }
      }
    }
    if (abs(ip->itwmod) == 1 && p->aerfrc) {
      p->aerfrc[0] = p->aerfrc[0] - sum1;
      p->aerfrc[1] = p->aerfrc[1] - sum2;
      p->aerfrc[2] = p->aerfrc[2] - sum3;
    }
  }
  for (int j = 1; j <= 4; j++)
    for (int i = 1; i <= nshl; i++)
      for (int e = 0; e < npro; e++) {
        int A = abs(ien[e + (size_t)npro * (i - 1)]) - 1;
        res[A + (size_t)nshg * (j - 1)] += rl[e][i][j];
      }
  free(rl);
}

/* ElmGMR (incompressible/elmgmr.f:1-330).  res[m] (nshg,4),
 * lhsK[m] (9,nnz_tot), lhsP[m] (4,nnz_tot) per part; qres/rmass of the parts are work
 * space; xKebe/xGoC (numel,9|4,nshape,nshape) per part optional (post-bc3LHS). */
void orc_inc_elmgmr(int nparts, orc_part *parts, const orc_incomp *ip, double **res, double **lhsK, double **lhsP,
                    double **xKebe, double **xGoC) {
  if (ip->itau != 0 || ip->ipord != 1) {
    fprintf(stderr, "orc_inc_elmgmr: only itau=0, ipord=1 are restated\n");
    abort();
  }
  if (ip->idiff == 1) {
    for (int m = 0; m < nparts; m++) {
      orc_part *p = &parts[m];
      memset(p->qres, 0, sizeof(double) * (size_t)p->c.nshg * 9);
      memset(p->rmass, 0, sizeof(double) * (size_t)p->c.nshg);
      for (int iblk = 0; iblk < p->c.nelblk; iblk++) asiq_i(p, ip, iblk);
    }
    qpbc_i(nparts, parts);
  }
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    size_t nshg = (size_t)p->c.nshg;
    memset(res[m], 0, sizeof(double) * nshg * 4);
    if (ip->lhs == 1) {
      size_t nnz_tot = (size_t)(p->colm[nshg] - 1);
      memset(lhsK[m], 0, sizeof(double) * 9 * nnz_tot);
      memset(lhsP[m], 0, sizeof(double) * 4 * nnz_tot);
    }
    for (int iblk = 0; iblk < p->c.nelblk; iblk++)
      asigmr_i(p, ip, iblk, res[m], lhsK[m], lhsP[m], xKebe ? xKebe[m] : NULL, xGoC ? xGoC[m] : NULL);
    if (p->aerfrc) memset(p->aerfrc + 4, 0, sizeof(double) * 10 * 1001); /* flxID = zero (elmgmr.f:130) */
    for (int iblk = 0; iblk < p->c.nelblb; iblk++) asbmfg_i(p, ip, iblk, res[m]);
  }
  if (nparts > 1) orc_commu(nparts, parts, res, 4, 0);
  for (int m = 0; m < nparts; m++) bc3res_i(&parts[m], res[m]);
}

/* bc3Res alone on one part's res(nshg,4) (incompressible/bc3res.f): a linear map, which lets a test isolate what the
 * boundary blocks contributed to the final residual */
void orc_inc_bc3res(orc_part *p, double *res) { bc3res_i(p, res); }

/* ---- lesSparse.f: the matrix-vector products of the coupled momentum/continuity system ---- */
#define KL(m, k) kLhs[((m)-1) + 9 * (size_t)((k)-1)]
#define PL(m, k) pLhs[((m)-1) + 4 * (size_t)((k)-1)]
#define PV(j, m) pv[((j)-1) + (size_t)n * ((m)-1)]
#define QV(j, m) q[((j)-1) + (size_t)n * ((m)-1)]

/* fLesSparseApG (lesSparse.f:204-245): q(n,3) = -G^T-scatter of the scalar p(n) */
void orc_les_apg(int n, const int *col, const int *row, const double *pLhs, const double *pv, double *q) {
  for (int i = 1; i <= n; i++) QV(i, 1) = QV(i, 2) = QV(i, 3) = 0.0;
  for (int i = 1; i <= n; i++) {
    double pisave = pv[i - 1];
    for (int k = col[i - 1]; k <= col[i] - 1; k++) {
      int j = row[k - 1];
      QV(j, 1) = QV(j, 1) - PL(1, k) * pisave;
      QV(j, 2) = QV(j, 2) - PL(2, k) * pisave;
      QV(j, 3) = QV(j, 3) - PL(3, k) * pisave;
    }
  }
}

/* fLesSparseApKG (lesSparse.f:252-320), ipvsq=0: q(n,3) = K p(:,1:3) - G^T-scatter p(:,4) */
void orc_les_apkg(int n, const int *col, const int *row, const double *kLhs, const double *pLhs, const double *pv,
                  double *q) {
  for (int i = 1; i <= n; i++) QV(i, 1) = QV(i, 2) = QV(i, 3) = 0.0;
  for (int i = 1; i <= n; i++) {
    double t1 = 0, t2 = 0, t3 = 0, pisave = PV(i, 4);
    for (int k = col[i - 1]; k <= col[i] - 1; k++) {
      int j = row[k - 1];
      t1 = t1 + KL(1, k) * PV(j, 1) + KL(4, k) * PV(j, 2) + KL(7, k) * PV(j, 3);
      t2 = t2 + KL(2, k) * PV(j, 1) + KL(5, k) * PV(j, 2) + KL(8, k) * PV(j, 3);
      t3 = t3 + KL(3, k) * PV(j, 1) + KL(6, k) * PV(j, 2) + KL(9, k) * PV(j, 3);
      QV(j, 1) = QV(j, 1) - PL(1, k) * pisave;
      QV(j, 2) = QV(j, 2) - PL(2, k) * pisave;
      QV(j, 3) = QV(j, 3) - PL(3, k) * pisave;
    }
    QV(i, 1) = QV(i, 1) + t1;
    QV(i, 2) = QV(i, 2) + t2;
    QV(i, 3) = QV(i, 3) + t3;
  }
}

/* fLesSparseApNGt (lesSparse.f:327-360) and ApNGtC (:368-403): q(n) = G p(:,1:3) [+ C p(:,4)] */
void orc_les_apngt(int n, const int *col, const int *row, const double *pLhs, const double *pv, double *q,
                   int withC) {
  for (int i = n; i >= 1; i--) {
    double t = 0;
    for (int k = col[i - 1]; k <= col[i] - 1; k++) {
      int j = row[k - 1];
      if (withC)
        t = t + PL(1, k) * PV(j, 1) + PL(2, k) * PV(j, 2) + PL(3, k) * PV(j, 3) + PL(4, k) * PV(j, 4);
      else
        t = t + PL(1, k) * PV(j, 1) + PL(2, k) * PV(j, 2) + PL(3, k) * PV(j, 3);
    }
    q[i - 1] = t;
  }
}

/* fLesSparseApFull (lesSparse.f:410-486), ipvsq=0: q(n,4) = [K -G^T; G C] p(n,4) */
void orc_les_apfull(int n, const int *col, const int *row, const double *kLhs, const double *pLhs, const double *pv,
                    double *q) {
  for (int i = 1; i <= n; i++) QV(i, 1) = QV(i, 2) = QV(i, 3) = 0.0;
  for (int i = 1; i <= n; i++) {
    double t1 = 0, t2 = 0, t3 = 0, t4 = 0, pisave = PV(i, 4);
    for (int k = col[i - 1]; k <= col[i] - 1; k++) {
      int j = row[k - 1];
      t1 = t1 + KL(1, k) * PV(j, 1) + KL(4, k) * PV(j, 2) + KL(7, k) * PV(j, 3);
      t2 = t2 + KL(2, k) * PV(j, 1) + KL(5, k) * PV(j, 2) + KL(8, k) * PV(j, 3);
      t3 = t3 + KL(3, k) * PV(j, 1) + KL(6, k) * PV(j, 2) + KL(9, k) * PV(j, 3);
      t4 = t4 + PL(1, k) * PV(j, 1) + PL(2, k) * PV(j, 2) + PL(3, k) * PV(j, 3) + PL(4, k) * PV(j, 4);
      QV(j, 1) = QV(j, 1) - PL(1, k) * pisave;
      QV(j, 2) = QV(j, 2) - PL(2, k) * pisave;
      QV(j, 3) = QV(j, 3) - PL(3, k) * pisave;
    }
    QV(i, 1) = QV(i, 1) + t1;
    QV(i, 2) = QV(i, 2) + t2;
    QV(i, 3) = QV(i, 3) + t3;
    QV(i, 4) = t4;
  }
}

/*
 * oracle_boundary.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 * Boundary-element flux AsBMFG -> e3b -> e3bvar (compressible/asbmfg.f:1-66,
 * e3b.f:1-386, e3bvar.f:1-374), ires=1, Navier=1, iLHScond=0, for linear
 * tets (triangular face, lcsyst 1), hexes (quadrilateral face, lcsyst 2) and
 * wedges with a triangular (lcsyst 3) or quadrilateral (lcsyst 4) boundary
 * face: lnode from getbnodes (common/hierarchic.f:90-190), the per-topology
 * normals and WdetJb of e3bvar.f:139-165.
 */
#include "oracle_internal.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* genint.f:30-75 (symtri rules, Qwtb*2) + genshpb.f:22-30 (shpTet at
 * Qptb(1,1:3,i), shglb/2).  Pinned by tests/golden (tri points/weights). */
void orc_tri_tables(int rule, int *nintb, double *Qwtb, double *shpb,
                    double *shglb) {
  double pts[3][3], w[3];
  int n;
  if (rule == 1) {
    n = 1;
    pts[0][0] = pts[0][1] = pts[0][2] = 0.333333333333333;
    w[0] = 1.0;
  } else if (rule == 2) {
    n = 3;
    const double a = 0.666666666666667, b = 0.166666666666667;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) pts[i][j] = (i == j) ? a : b;
      w[i] = 0.333333333333333;
    }
  } else {
    fprintf(stderr, "orc_tri_tables: rule %d not restated\n", rule);
    abort();
  }
  nintb[0] = n;
  for (int i = 0; i < n; i++) {
    Qwtb[0 + ORC_MAXTOP * i] = 2.0 * w[i];
    double L[4] = {pts[i][0], pts[i][1], pts[i][2],
                   1.0 - pts[i][0] - pts[i][1] - pts[i][2]};
    for (int a = 0; a < 4; a++) {
      shpb[0 + ORC_MAXTOP * (a + ORC_MAXSH * i)] = L[a];
      for (int j = 0; j < 3; j++) {
        double dN = (a == 3) ? -1.0 : ((a == j) ? 1.0 : 0.0);
        shglb[0 + ORC_MAXTOP * (j + 3 * (a + ORC_MAXSH * i))] = dN / 2.0;
      }
    }
  }
}

static void getdiff_b(const orc_common *c, double T, double cp, double *rmu,
                      double *rlm, double *rlm2mu, double *con) {
  /* same as getDiff in oracle_elem.c (compressible/getdiff.f), DNS */
  const double pt66 = 0.6666666666666666666666666666667;
  double mu = (c->matflg2 == 0)
                  ? c->datmat121
                  : c->datmat121 * (T / c->datmat221) * sqrt(T / c->datmat221) *
                        (c->datmat221 + c->datmat321) / (T + c->datmat321);
  double lm = (c->matflg3 == 0) ? -pt66 * mu : (c->datmat131 - pt66) * mu;
  *rmu = mu;
  *rlm = lm;
  *rlm2mu = lm + 2.0 * mu;
  *con = mu * cp / c->pr;
}

/* AsBMFG + e3b + e3bvar for one boundary block */
void orc_asbmfg(const orc_part *p, int iblk, double *res) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblkb + 10 * iblk;
  int iel = lc[0], lcsyst = lc[2], nenl = lc[4], nenbl = lc[5], nshl = lc[8],
      nshlb = lc[9];
  int npro = lc[10] - iel;
  if (lcsyst == 3) lcsyst = nenbl; /* elmgmr.f:191 */
  /* getbnodes (hierarchic.f:102-168), ipord = 1 */
  int lnode[5] = {0, 1, 2, 3, 4};
  int ok = (lcsyst == 1 && nshl == 4 && nshlb == 3 && nenbl == 3) ||
           (lcsyst == 2 && nshl == 8 && nshlb == 4 && nenbl == 4) ||
           (lcsyst == 3 && nshl == 6 && nshlb == 3 && nenbl == 3) ||
           (lcsyst == 4 && nshl == 6 && nshlb == 4 && nenbl == 4);
  if (!ok) {
    fprintf(stderr, "orc_asbmfg: boundary block lcsyst %d nshl %d nshlb %d not restated\n", lcsyst, nshl, nshlb);
    abort();
  }
  if (lcsyst == 4) {
    lnode[2] = 4;
    lnode[3] = 5;
    lnode[4] = 2;
  }
  int ngaussb = c->nintb[lcsyst - 1];
  const int *ien = p->ienb + p->ienb_off[iblk];
  const int *iBCB = p->iBCB + p->iBCB_off[iblk];     /* (npro,2)         */
  const double *BCB = p->BCB + p->BCB_off[iblk];     /* (npro,nshlb,6)   */
  int nshg = c->nshg;
  double(*rl)[9][6] = calloc((size_t)npro, sizeof *rl);
  for (int e = 0; e < npro; e++) {
    double yl[9][6], xlb[9][4];
    for (int n = 1; n <= nshl; n++) {
      int A = ien[e + (size_t)npro * (n - 1)] - 1;
      yl[n][1] = p->y[A + (size_t)nshg * 3];
      yl[n][2] = p->y[A + (size_t)nshg * 0];
      yl[n][3] = p->y[A + (size_t)nshg * 1];
      yl[n][4] = p->y[A + (size_t)nshg * 2];
      yl[n][5] = p->y[A + (size_t)nshg * 4];
      for (int i = 1; i <= 3; i++)
        xlb[n][i] = p->x[A + (size_t)c->numnp * (i - 1)];
    }
    int ibcb = iBCB[e];
    for (int intp = 1; intp <= ngaussb; intp++) {
      if (QWTB(c, lcsyst, intp) == 0.0) continue;
      double shape[9], shdrv[4][9];
      for (int n = 1; n <= nshl; n++) {
        shape[n] = SHPB(p, lcsyst, n, intp);
        for (int i = 1; i <= 3; i++) shdrv[i][n] = SHGLB(p, lcsyst, i, n, intp);
      }
      /* ---- e3bvar (e3bvar.f:78-356) ---- */
      double pres = 0, u1 = 0, u2 = 0, u3 = 0, T = 0;
      for (int k = 1; k <= nshlb; k++) {
        int n = lnode[k];
        pres += shape[n] * yl[n][1];
        u1 += shape[n] * yl[n][2];
        u2 += shape[n] * yl[n][3];
        u3 += shape[n] * yl[n][4];
        T += shape[n] * yl[n][5];
      }
      double rk = 0.5 * (u1 * u1 + u2 * u2 + u3 * u3);
      double rho = pres / (c->Rgas * T);
      double ei = T * (c->Rgas / c->gamma1);
      double cp = c->Rgas * c->gamma / c->gamma1;
      double dxdxib[4][4];
      memset(dxdxib, 0, sizeof dxdxib);
      for (int n = 1; n <= nenl; n++)
        for (int i = 1; i <= 3; i++)
          for (int j = 1; j <= 3; j++) dxdxib[i][j] += xlb[n][i] * shdrv[j][n];
      double v1[4], v2[4];
      for (int i = 1; i <= 3; i++) {
        v1[i] = xlb[2][i] - xlb[1][i];
        v2[i] = xlb[3][i] - xlb[1][i];
      }
      double t1, t2, t3;
      if (lcsyst == 4) { /* e3bvar.f:139-146 */
        t1 = dxdxib[2][1] * dxdxib[3][3] - dxdxib[2][3] * dxdxib[3][1];
        t2 = dxdxib[3][1] * dxdxib[1][3] - dxdxib[3][3] * dxdxib[1][1];
        t3 = dxdxib[1][1] * dxdxib[2][3] - dxdxib[1][3] * dxdxib[2][1];
      } else if (lcsyst == 1) {
        t1 = v1[2] * v2[3] - v2[2] * v1[3];
        t2 = v2[1] * v1[3] - v1[1] * v2[3];
        t3 = v1[1] * v2[2] - v2[1] * v1[2];
      } else { /* e3bvar.f:152-155 */
        t1 = -v1[2] * v2[3] + v2[2] * v1[3];
        t2 = -v2[1] * v1[3] + v1[1] * v2[3];
        t3 = -v1[1] * v2[2] + v2[1] * v1[2];
      }
      double temp = 1.0 / sqrt(t1 * t1 + t2 * t2 + t3 * t3);
      double bn[4] = {0, t1 * temp, t2 * temp, t3 * temp};
      double WdetJb; /* e3bvar.f:163-176 */
      if (lcsyst == 3)
        WdetJb = (1 - QWTB(c, lcsyst, intp)) / (4.0 * temp);
      else if (lcsyst == 4)
        WdetJb = QWTB(c, lcsyst, intp) / temp;
      else
        WdetJb = QWTB(c, lcsyst, intp) / (4.0 * temp);
      double d[4][4];
      d[1][1] = dxdxib[2][2] * dxdxib[3][3] - dxdxib[3][2] * dxdxib[2][3];
      d[1][2] = dxdxib[3][2] * dxdxib[1][3] - dxdxib[1][2] * dxdxib[3][3];
      d[1][3] = dxdxib[1][2] * dxdxib[2][3] - dxdxib[1][3] * dxdxib[2][2];
      temp = 1.0 / (d[1][1] * dxdxib[1][1] + d[1][2] * dxdxib[2][1] +
                    d[1][3] * dxdxib[3][1]);
      d[1][1] *= temp;
      d[1][2] *= temp;
      d[1][3] *= temp;
      d[2][1] = (dxdxib[2][3] * dxdxib[3][1] - dxdxib[2][1] * dxdxib[3][3]) * temp;
      d[2][2] = (dxdxib[1][1] * dxdxib[3][3] - dxdxib[3][1] * dxdxib[1][3]) * temp;
      d[2][3] = (dxdxib[2][1] * dxdxib[1][3] - dxdxib[1][1] * dxdxib[2][3]) * temp;
      d[3][1] = (dxdxib[2][1] * dxdxib[3][2] - dxdxib[2][2] * dxdxib[3][1]) * temp;
      d[3][2] = (dxdxib[3][1] * dxdxib[1][2] - dxdxib[1][1] * dxdxib[3][2]) * temp;
      d[3][3] = (dxdxib[1][1] * dxdxib[2][2] - dxdxib[1][2] * dxdxib[2][1]) * temp;
      double gl[4][6];
      memset(gl, 0, sizeof gl);
      for (int n = 1; n <= nshl; n++)
        for (int i = 1; i <= 3; i++)
          for (int m = 1; m <= 5; m++) gl[i][m] += shdrv[i][n] * yl[n][m];
      double g1[6], g2[6], g3[6];
      for (int m = 2; m <= 5; m++) {
        g1[m] = d[1][1] * gl[1][m] + d[2][1] * gl[2][m] + d[3][1] * gl[3][m];
        g2[m] = d[1][2] * gl[1][m] + d[2][2] * gl[2][m] + d[3][2] * gl[3][m];
        g3[m] = d[1][3] * gl[1][m] + d[2][3] * gl[2][m] + d[3][3] * gl[3][m];
      }
      double rou = 0, pb = 0, Fv2 = 0, Fv3 = 0, Fv4 = 0, Fh5 = 0;
      for (int n = 1; n <= nshlb; n++) {
#define BCBv(n, k) BCB[e + (size_t)npro * (((n)-1) + (size_t)nshlb * ((k)-1))]
        double sh = shape[lnode[n]];
        rou += sh * BCBv(n, 1);
        pb += sh * BCBv(n, 2);
        Fv2 += sh * BCBv(n, 3);
        Fv3 += sh * BCBv(n, 4);
        Fv4 += sh * BCBv(n, 5);
        Fh5 += sh * BCBv(n, 6);
#undef BCBv
      }
      /* ---- e3b (e3b.f:123-283) ---- */
      double un;
      if (!(ibcb & 1)) {
        un = bn[1] * u1 + bn[2] * u2 + bn[3] * u3;
        rou = rho * un;
      } else {
        un = rou / rho;
      }
      if (!(ibcb & 2)) pb = pres;
      double F1 = rou;
      double F2 = rou * u1 + bn[1] * pb;
      double F3 = rou * u2 + bn[2] * pb;
      double F4 = rou * u3 + bn[3] * pb;
      double F5 = rou * (ei + rk) + un * pb;
      double rmu, rlm, rlm2mu, con;
      getdiff_b(c, T, cp, &rmu, &rlm, &rlm2mu, &con);
      double tau1n = bn[1] * (rlm2mu * g1[2] + rlm * g2[3] + rlm * g3[4]) +
                     bn[2] * (rmu * (g2[2] + g1[3])) +
                     bn[3] * (rmu * (g3[2] + g1[4]));
      double tau2n = bn[1] * (rmu * (g2[2] + g1[3])) +
                     bn[2] * (rlm * g1[2] + rlm2mu * g2[3] + rlm * g3[4]) +
                     bn[3] * (rmu * (g3[3] + g2[4]));
      double tau3n = bn[1] * (rmu * (g3[2] + g1[4])) +
                     bn[2] * (rmu * (g3[3] + g2[4])) +
                     bn[3] * (rlm * g1[2] + rlm * g2[3] + rlm2mu * g3[4]);
      if (!(ibcb & 4)) {
        Fv2 = tau1n;
        Fv3 = tau2n;
        Fv4 = tau3n;
      }
      double Fv5 = u1 * Fv2 + u2 * Fv3 + u3 * Fv4;
      double heat = -con * (bn[1] * g1[5] + bn[2] * g2[5] + bn[3] * g3[5]);
      if (!(ibcb & 8)) Fh5 = heat;
      F2 = F2 - Fv2;
      F3 = F3 - Fv3;
      F4 = F4 - Fv4;
      F5 = F5 - Fv5 + Fh5;
      if (p->aerfrc) { /* flxID (e3b.f:305-321), Force/HFlux (e3b.f:325-345, iter==nitr) */
        int iface = abs(iBCB[e + (size_t)npro]);
        double *fl = p->aerfrc + 4 + 10 * iface;
        if (iface != 0) {
          fl[0] += WdetJb;
          fl[1] -= WdetJb * rou;
          fl[2] -= (tau1n - bn[1] * pres) * WdetJb;
          fl[3] -= (tau2n - bn[2] * pres) * WdetJb;
          fl[4] -= (tau3n - bn[3] * pres) * WdetJb;
        }
        if (!(ibcb & 1)) {
          p->aerfrc[0] += (pres * bn[1] - tau1n) * WdetJb;
          p->aerfrc[1] += (pres * bn[2] - tau2n) * WdetJb;
          p->aerfrc[2] += (pres * bn[3] - tau3n) * WdetJb;
          p->aerfrc[3] += -heat * WdetJb;
        }
      }
      for (int k = 1; k <= nshlb; k++) {
        int n = lnode[k];
        rl[e][n][1] += WdetJb * shape[n] * F1;
        rl[e][n][2] += WdetJb * shape[n] * F2;
        rl[e][n][3] += WdetJb * shape[n] * F3;
        rl[e][n][4] += WdetJb * shape[n] * F4;
        rl[e][n][5] += WdetJb * shape[n] * F5;
      }
    }
  }
  for (int j = 1; j <= 5; j++)
    for (int i = 1; i <= nshl; i++)
      for (int e = 0; e < npro; e++) {
        int A = ien[e + (size_t)npro * (i - 1)] - 1;
        res[A + (size_t)nshg * (j - 1)] += rl[e][i][j];
      }
  free(rl);
}

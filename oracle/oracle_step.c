/* oracle_step.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 *
 * CPU restatement of the Newton / time-step shell around SolGMR*: the
 * predictor-multicorrector routines of phSolver/compressible/itrPC.f, itrBC
 * (compressible/itrbc.f:1-199), the residual statistics of rstat
 * (compressible/rstat.f:94-112) and the flow part of the step loop in
 * compressible/itrdrv.f:393-457,511-524,590-594 (stepseq "0 1 0 1 ...":
 * solve, update, solve, update, then itrUpdate).
 */
#include "oracle_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_elmgmrs(int nparts, orc_part *parts);

/* itrPredict (itrPC.f:54-119) without the itrBC calls of ipred 2..4, which the
 * caller issues (they sit between the y and the ac statement). */
static void predict_y(const orc_part *p, int ipred, double *y, const double *yold,
                      const double *acold) {
  const orc_common *c = &p->c;
  size_t n = (size_t)c->nshg * c->ndof;
  double almi = c->almi, alfi = c->alfi, gami = c->gami, Dtgl = c->Dtgl;
  (void)almi;
  for (size_t i = 0; i < n; i++) {
    if (ipred == 1) y[i] = yold[i];                                      /* :71 */
    else if (ipred == 2) y[i] = yold[i] + alfi / Dtgl * acold[i] * (1.0 - gami); /* :79 */
    else if (ipred == 3) y[i] = yold[i] + alfi / Dtgl * acold[i];        /* :92 */
    else {                                                               /* :105-108 */
      double fct1 = alfi / (1.0 - alfi);
      y[i] = yold[i] + fct1 * (yold[i] - y[i]);
    }
  }
}
static void predict_ac(const orc_part *p, int ipred, const double *y, double *ac,
                       const double *yold, const double *acold) {
  const orc_common *c = &p->c;
  size_t n = (size_t)c->nshg * c->ndof;
  double almi = c->almi, alfi = c->alfi, gami = c->gami, Dtgl = c->Dtgl;
  for (size_t i = 0; i < n; i++) {
    if (ipred == 1) ac[i] = acold[i] * (1.0 - almi / gami);              /* :72 */
    else if (ipred == 2) ac[i] = acold[i] * (1.0 - almi);                /* :85 */
    else if (ipred == 3) ac[i] = acold[i];                               /* :98 */
    else {                                                               /* :106-107,114 */
      double fct2 = 1.0 - almi / gami, fct3 = almi / gami / alfi * Dtgl;
      ac[i] = acold[i] * fct2 + (y[i] - yold[i]) * fct3;
    }
  }
}

/* itrBC (itrbc.f:60-199), ylimit off, iabc=0.  y, ac (nshg,5) {u1,u2,u3,p,T}.
 * NOTE (kept on purpose): the density branch writes the pressure computed from
 * (rho_BC, T) into y(:,1) -- the x1-velocity slot of the global ordering
 * (itrbc.f:143-163). */
void orc_itrbc(int nparts, orc_part *parts, double **y, double **ac, int ires) {
  for (int m = 0; m < nparts; m++) {
    const orc_part *p = &parts[m];
    const orc_common *c = &p->c;
    int nshg = c->nshg;
    double *Y = y[m], *A = ac[m];
#define YY(i, j) Y[(i) + (size_t)nshg * ((j)-1)]
#define BCv(i, j) p->BC[(i) + (size_t)nshg * ((j)-1)]
    int anyrho = 0;
    for (int i = 0; i < nshg; i++) {
      int ib = p->iBC[i];
      if (ib & 2) YY(i, 5) = BCv(i, 2);                                   /* :60-62 */
      switch ((ib >> 3) & 7) {                                            /* :69-128 */
        case 1: YY(i, 1) = BCv(i, 3) - BCv(i, 4) * YY(i, 2) - BCv(i, 5) * YY(i, 3); break;
        case 2: YY(i, 2) = BCv(i, 3) - BCv(i, 4) * YY(i, 1) - BCv(i, 5) * YY(i, 3); break;
        case 3:
          YY(i, 1) = BCv(i, 3) - BCv(i, 4) * YY(i, 3);
          YY(i, 2) = BCv(i, 5) - BCv(i, 6) * YY(i, 3);
          break;
        case 4: YY(i, 3) = BCv(i, 3) - BCv(i, 4) * YY(i, 1) - BCv(i, 5) * YY(i, 2); break;
        case 5:
          YY(i, 1) = BCv(i, 3) - BCv(i, 4) * YY(i, 2);
          YY(i, 3) = BCv(i, 5) - BCv(i, 6) * YY(i, 2);
          break;
        case 6:
          YY(i, 2) = BCv(i, 3) - BCv(i, 4) * YY(i, 1);
          YY(i, 3) = BCv(i, 5) - BCv(i, 6) * YY(i, 1);
          break;
        case 7:
          YY(i, 1) = BCv(i, 3);
          YY(i, 2) = BCv(i, 4);
          YY(i, 3) = BCv(i, 5);
          break;
        default: break;
      }
      if (ib & 1) anyrho = 1;
    }
    if (anyrho)                                                           /* :137-166 */
      for (int i = 0; i < nshg; i++)
        if (p->iBC[i] & 1) YY(i, 1) = c->Rgas * BCv(i, 1) * YY(i, 5);     /* getthm.f:75 */
    for (int i = 0; i < nshg; i++)                                        /* :170-177 */
      if (p->iBC[i] & 4) YY(i, 4) = BCv(i, 1);
    for (int j = 1; j <= 5; j++)                                          /* :181-184 */
      for (int i = 0; i < nshg; i++) {
        int mst = p->iper[i] - 1;
        YY(i, j) = YY(mst, j);
        if (ires != 2) A[i + (size_t)nshg * (j - 1)] = A[mst + (size_t)nshg * (j - 1)];
      }
#undef YY
#undef BCv
  }
  if (parts[0].c.numpe > 1) {                                             /* :188-191 */
    orc_commu(nparts, parts, y, 5, 1);
    if (ires != 2) orc_commu(nparts, parts, ac, 5, 1);
  }
}

/* itrCorrect (itrPC.f:127-150) */
void orc_itrcorrect(const orc_part *p, double *y, double *ac, const double *yold,
                    const double *acold, const double *Dy) {
  const orc_common *c = &p->c;
  int nshg = c->nshg;
  for (int i = 0; i < nshg; i++) {
    for (int k = 0; k < 3; k++) y[i + (size_t)nshg * k] -= Dy[i + (size_t)nshg * (k + 1)];
    y[i + (size_t)nshg * 3] -= Dy[i];
    y[i + (size_t)nshg * 4] -= Dy[i + (size_t)nshg * 4];
  }
  double fct1 = 1.0 - c->almi / c->gami;
  double fct2 = c->almi * c->Dtgl / c->gami / c->alfi;
  size_t n = (size_t)nshg * 5;
  for (size_t i = 0; i < n; i++) ac[i] = acold[i] * fct1 + (y[i] - yold[i]) * fct2;
}

/* itrUpdate (itrPC.f:205-210) */
void orc_itrupdate(const orc_part *p, double *yold, double *acold, const double *y,
                   const double *ac) {
  const orc_common *c = &p->c;
  double fct2 = 1.0 / c->almi, fct3 = 1.0 / c->alfi;
  size_t n = (size_t)c->nshg * 5;
  for (size_t i = 0; i < n; i++) {
    acold[i] = acold[i] + (ac[i] - acold[i]) * fct2;
    yold[i] = yold[i] + (y[i] - yold[i]) * fct3;
  }
}

/* rstat (rstat.f:94-112): totres(1:2) = sqrt(sum res^2, sum b^2)/nshgt */
void orc_rstat(int nparts, orc_part *parts, int nshgt, double *totres) {
  double s[2] = {0.0, 0.0};
  for (int m = 0; m < nparts; m++) {
    const orc_part *p = &parts[m];
    size_t n = (size_t)p->c.nshg * 5;
    double a = 0.0, b = 0.0;
    for (int j = 0; j < 5; j++)
      for (int i = 0; i < p->c.nshg; i++) {
        double r = p->res[i + (size_t)p->c.nshg * j], q = p->rmes[i + (size_t)p->c.nshg * j];
        a += r * r;
        b += q * q;
      }
    (void)n;
    s[0] += a;
    s[1] += b;
  }
  totres[0] = sqrt(s[0]) / (double)nshgt;
  totres[1] = sqrt(s[1]) / (double)nshgt;
}

/* One time step of itrdrv.f (flow solves only): predictor, nitr x (solve,
 * itrCorrect, itrBC), itrUpdate.  y/ac/yold/acold per part are caller-owned
 * and updated in place; parts[m].y / .ac must point at y[m] / ac[m].
 * sparse: 0 SolGMRe, 1 SolGMRs.  LHSupd: itrdrv.f:456,511 (lhs = 1 -
 * min(1, mod(ifuncs-1, LHSupd))); *ifuncs persists across steps.
 * stats[6*it + {0,1}] = totres(1:2), [2] = iKs, [3] = lGMRES, [4] = lhs. */
void orc_timestep(int nparts, orc_part *parts, double **y, double **ac, double **yold,
                  double **acold, int ipred, int nitr, int sparse, int LHSupd, int nshgt,
                  int *ifuncs, int *ntotGM, double *stats) {
  int K = parts[0].c.Kspace;
  double *HBrg = calloc((size_t)(K + 1) * K, sizeof(double));
  double *eBrg = calloc((size_t)K + 1, sizeof(double)), *yBrg = calloc((size_t)K + 1, sizeof(double));
  double *Rcos = calloc((size_t)K + 1, sizeof(double)), *Rsin = calloc((size_t)K + 1, sizeof(double));
  for (int m = 0; m < nparts; m++) predict_y(&parts[m], ipred, y[m], yold[m], acold[m]);
  if (ipred != 1) orc_itrbc(nparts, parts, y, ac, 1);
  for (int m = 0; m < nparts; m++) predict_ac(&parts[m], ipred, y[m], ac[m], yold[m], acold[m]);
  orc_itrbc(nparts, parts, y, ac, 1);                                     /* itrdrv.f:394 */
  for (int it = 1; it <= nitr; it++) {
    *ifuncs += 1;
    int lhs = 1 - ((((*ifuncs - 1) % LHSupd) > 0) ? 1 : 0);
    for (int m = 0; m < nparts; m++) {
      parts[m].c.lhs = lhs;
      parts[m].c.iprec = lhs;
    }
    for (int m = 0; m < nparts; m++)                                      /* itrdrv.f:437-442 */
      if (parts[m].aerfrc) memset(parts[m].aerfrc, 0, sizeof(double) * 4);
    int iKs = 0, lG = 0;
    if (sparse)
      orc_solgmrs(nparts, parts, HBrg, eBrg, yBrg, Rcos, Rsin, &iKs, &lG, ntotGM);
    else
      orc_solgmre(nparts, parts, HBrg, eBrg, yBrg, Rcos, Rsin, &iKs, &lG, ntotGM);
    orc_rstat(nparts, parts, nshgt, stats + 6 * (it - 1));
    stats[6 * (it - 1) + 2] = iKs;
    stats[6 * (it - 1) + 3] = lG;
    stats[6 * (it - 1) + 4] = lhs;
    for (int m = 0; m < nparts; m++)
      orc_itrcorrect(&parts[m], y[m], ac[m], yold[m], acold[m], parts[m].Dy);
    orc_itrbc(nparts, parts, y, ac, 1);
  }
  for (int m = 0; m < nparts; m++) orc_itrupdate(&parts[m], yold[m], acold[m], y[m], ac[m]);
  orc_itrbc(nparts, parts, yold, acold, 1);                               /* itrdrv.f:652 */
  free(HBrg); free(eBrg); free(yBrg); free(Rcos); free(Rsin);
}

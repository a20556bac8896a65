/*
 * oracle_mfg.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 *
 * The matrix-free flavour of the Newton linear solve: SolMFG
 * (compressible/solmfg.f:1-381), ElmMFG (elmmfg.f:1-256), ItrRes
 * (itrres.f:1-171), Au1MFG / Au2MFG (au1mfg.f:1-98, au2mfg.f:1-120), itrFDI
 * (itrfdi.f:1-139) and yshuffle (shuffle.f:1-27).  The element level
 * (AsIMFG / AsIRes / e3 with ires=2|3 / e3bdg) is in oracle_elem.c.
 */
#include "oracle_internal.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_asimfg(const orc_part *p, int iblk, const double *qres, double *res, double *rmes,
                double *BDiag);
void orc_asires(const orc_part *p, int iblk, const double *yp, double *rmes, int iabres);
void orc_gmres_core_mfg(int nparts, orc_part *parts, void (*Ap)(int, orc_part *, double **),
                        void (*restart)(int, orc_part *, double **), int minIters, double *HBrg,
                        double *eBrg, double *yBrg, double *Rcos, double *Rsin, int *iKs,
                        int *lGMRES, int *ntotGM);

/* ElmMFG (elmmfg.f:60-250): ires=3 residual, modified residual and (iprec/=0)
 * the e3bdg block diagonal.  Jactyp=0 (itrPC.f:29): boundary elements add to
 * res only (asbmfg.f:58-59, e3b.f:296-297). */
void orc_elmmfg(int nparts, orc_part *parts) {
  for (int m = 0; m < nparts; m++) parts[m].c.ires = 3;
  if (parts[0].c.idiff == 1 || parts[0].c.idiff == 3) {
    for (int m = 0; m < nparts; m++) {
      orc_part *p = &parts[m];
      size_t nshg = (size_t)p->c.nshg;
      memset(p->qres, 0, sizeof(double) * nshg * 12);
      memset(p->rmass, 0, sizeof(double) * nshg);
      for (int iblk = 0; iblk < p->c.nelblk; iblk++) orc_asiq(p, iblk, p->qres, p->rmass);
    }
    orc_qpbc(nparts, parts);
  }
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    const orc_common *c = &p->c;
    size_t nshg = (size_t)c->nshg;
    memset(p->res, 0, sizeof(double) * nshg * 5);
    memset(p->rmes, 0, sizeof(double) * nshg * 5);
    if (c->iprec != 0) memset(p->BDiag, 0, sizeof(double) * nshg * 25);
    for (int iblk = 0; iblk < c->nelblk; iblk++)
      orc_asimfg(p, iblk, p->qres, p->res, p->rmes, p->BDiag);
    if (p->aerfrc) memset(p->aerfrc + 4, 0, sizeof(double) * 10 * 1001);
    for (int iblk = 0; iblk < c->nelblb; iblk++) orc_asbmfg(p, iblk, p->res);
  }
  if (nparts > 1) { /* :226-232 */
    double **g = malloc(sizeof(double *) * nparts);
    for (int m = 0; m < nparts; m++) g[m] = parts[m].res;
    orc_commu(nparts, parts, g, 5, 0);
    for (int m = 0; m < nparts; m++) g[m] = parts[m].rmes;
    orc_commu(nparts, parts, g, 5, 0);
    if (parts[0].c.iprec != 0) {
      for (int m = 0; m < nparts; m++) g[m] = parts[m].BDiag;
      orc_commu(nparts, parts, g, 25, 0);
    }
    free(g);
  }
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    orc_bc3res(p, p->res);
    orc_bc3res(p, p->rmes);
    if (p->c.iprec != 0) orc_bc3bdg(p, p->BDiag);
  }
}

/* ItrRes (itrres.f:58-165): rmes += modified residual of the perturbed state
 * yp (global {u,p,T} order) with coefficients frozen at parts[m].y; sets
 * ires=2 and iprec=0 in COMMON as the reference does.  rmes is NOT zeroed
 * (the callers do; itrFDI accumulates two calls on purpose). */
void orc_itrres(int nparts, orc_part *parts, double **yp, double **rmes, int iabres) {
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    p->c.ires = 2;
    p->c.iprec = 0;
    for (int iblk = 0; iblk < p->c.nelblk; iblk++) orc_asires(p, iblk, yp[m], rmes[m], iabres);
  }
  if (nparts > 1) orc_commu(nparts, parts, rmes, 5, 0);
  for (int m = 0; m < nparts; m++) orc_bc3res(&parts[m], rmes[m]);
}

/* yshuffle (shuffle.f:8-24) */
static void yshuffle(const orc_part *p, double *r, int old2new) {
  int nshg = p->c.nshg;
  for (int i = 0; i < nshg; i++) {
    double v[5];
    for (int k = 0; k < 5; k++) v[k] = r[i + (size_t)nshg * k];
    if (old2new) { /* {p,u1,u2,u3,T} -> {u1,u2,u3,p,T} */
      r[i] = v[1];
      r[i + (size_t)nshg] = v[2];
      r[i + (size_t)nshg * 2] = v[3];
      r[i + (size_t)nshg * 3] = v[0];
    } else {
      r[i] = v[3];
      r[i + (size_t)nshg] = v[0];
      r[i + (size_t)nshg * 2] = v[1];
      r[i + (size_t)nshg * 3] = v[2];
    }
  }
}

/* state shared by the Ap callbacks of one SolMFG call */
static struct {
  double **ypre;
  double eGMRES;
  double **work; /* uBtmp */
  double **work2;
} M;

/* R(ypre + eps*dir) preconditioned: the body shared by Au1MFG (au1mfg.f:58-84),
 * Au2MFG (au2mfg.f:61-105) and itrFDI (itrfdi.f:58-121): v <- ypre + eps*dir,
 * i3LU backward, old2new, [itrBC], ItrRes into out (out zeroed iff zero_out),
 * [i3LU forward].  v is overwritten with the perturbed state. */
static void perturbed_res(int nparts, orc_part *parts, double **v, double eps, double **dir,
                          double **out, int zero_out, int with_itrbc, int iabres, int forward) {
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    size_t n = (size_t)p->c.nshg * 5;
    if (dir)
      for (size_t i = 0; i < n; i++) v[m][i] = M.ypre[m][i] + eps * dir[m][i];
    else
      memcpy(v[m], M.ypre[m], sizeof(double) * n);
    orc_i3lu(&p->c, p->BDiag, v[m], 2);
    yshuffle(p, v[m], 1);
    if (zero_out) memset(out[m], 0, sizeof(double) * n);
  }
  if (with_itrbc) orc_itrbc(nparts, parts, v, v, 2);
  orc_itrres(nparts, parts, v, out, iabres);
  if (forward)
    for (int m = 0; m < nparts; m++) orc_i3lu(&parts[m].c, parts[m].BDiag, out[m], 1);
}

/* Au1MFG (au1mfg.f:52-90): u <- ( L^-1 Rm(U^-1(ypre + e u)) - rmes ) / e */
void orc_au1mfg(int nparts, orc_part *parts, double **u) {
  double e = M.eGMRES;
  perturbed_res(nparts, parts, u, e, u, M.work, 1, 1, 0, 1);
  for (int m = 0; m < nparts; m++) {
    size_t n = (size_t)parts[m].c.nshg * 5;
    for (size_t i = 0; i < n; i++) u[m][i] = (M.work[m][i] - parts[m].rmes[i]) / e;
  }
}

/* Au2MFG (au2mfg.f:52-112): temp <- res - (Rm(+eps Dy) - Rm(-eps Dy)) / (2 eps) */
static void au2mfg(int nparts, orc_part *parts, double **t) {
  double **dy = malloc(sizeof(double *) * nparts);
  for (int m = 0; m < nparts; m++) {
    size_t n = (size_t)parts[m].c.nshg * 5;
    for (size_t i = 0; i < n; i++) t[m][i] = parts[m].Dy[i] * parts[m].Dy[i];
    dy[m] = parts[m].Dy;
  }
  double summed = orc_sumgat(nparts, parts, t, 5);
  double eps = pow(parts[0].c.epsM, 0.6666666666666666666666666666667) / sqrt(summed); /* epsM**pt66 */
  perturbed_res(nparts, parts, t, eps, dy, M.work, 1, 1, 0, 1);
  perturbed_res(nparts, parts, t, -eps, dy, M.work2, 1, 1, 0, 1);
  for (int m = 0; m < nparts; m++) {
    size_t n = (size_t)parts[m].c.nshg * 5;
    for (size_t i = 0; i < n; i++)
      t[m][i] = parts[m].res[i] - (M.work[m][i] - M.work2[m][i]) / (2.0 * eps);
  }
  free(dy);
}

/* itrFDI (itrfdi.f:52-132): the finite-difference interval eGMRES */
static double itrfdi(int nparts, orc_part *parts, double **dirres) {
  double epsM = parts[0].c.epsM;
  double **t = malloc(sizeof(double *) * nparts);
  for (int m = 0; m < nparts; m++) t[m] = parts[m].temp;
  /* |Rm| with absolute element contributions (iabres=1), no itrBC here */
  perturbed_res(nparts, parts, t, 0.0, NULL, M.work, 1, 0, 1, 1);
  for (int m = 0; m < nparts; m++) {
    size_t n = (size_t)parts[m].c.nshg * 5;
    for (size_t i = 0; i < n; i++) M.work[m][i] = M.work[m][i] * M.work[m][i];
  }
  double epsA = (epsM * epsM) * sqrt(orc_sumgat(nparts, parts, M.work, 5));
  double epsSD = sqrt(epsM);
  /* rtmp = Rm(+epsSD res) + Rm(-epsSD res): one forward reduction of the sum */
  perturbed_res(nparts, parts, t, epsSD, dirres, M.work, 1, 0, 0, 0);
  perturbed_res(nparts, parts, t, -epsSD, dirres, M.work, 0, 0, 0, 1);
  for (int m = 0; m < nparts; m++) {
    size_t n = (size_t)parts[m].c.nshg * 5;
    for (size_t i = 0; i < n; i++) {
      double v = (M.work[m][i] - 2.0 * parts[m].rmes[i]) / epsM;
      M.work[m][i] = v * v;
    }
  }
  double SDnrm = sqrt(orc_sumgat(nparts, parts, M.work, 5));
  free(t);
  return 2.0 * sqrt(epsA / SDnrm);
}

/* for the tests: one Au1MFG application with a given interval */
void orc_mfg_begin(int nparts, orc_part *parts, double eGMRES) {
  M.ypre = malloc(sizeof(double *) * nparts);
  M.work = malloc(sizeof(double *) * nparts);
  M.work2 = malloc(sizeof(double *) * nparts);
  M.eGMRES = eGMRES;
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    size_t n = (size_t)p->c.nshg * 5;
    M.ypre[m] = malloc(sizeof(double) * n);
    M.work[m] = malloc(sizeof(double) * n);
    M.work2[m] = malloc(sizeof(double) * n);
    memcpy(M.ypre[m], p->y, sizeof(double) * n); /* ypre = y(:,1:nflow) (solmfg.f:126) */
    yshuffle(p, M.ypre[m], 0);
    orc_i3lu(&p->c, p->BDiag, M.ypre[m], 3);
  }
}
void orc_mfg_end(int nparts) {
  for (int m = 0; m < nparts; m++) {
    free(M.ypre[m]);
    free(M.work[m]);
    free(M.work2[m]);
  }
  free(M.ypre);
  free(M.work);
  free(M.work2);
}

/* SolMFG (solmfg.f:86-375).  eGMRES lives in COMMON /itrpar/ (common.h:217):
 * in/out; recomputed by itrFDI when iter==1 and mod(istep,20)==0. */
void orc_solmfg(int nparts, orc_part *parts, double *HBrg, double *eBrg, double *yBrg,
                double *Rcos, double *Rsin, int *iKs, int *lGMRES, int *ntotGM, double *eGMRES,
                int iter, int istep) {
  orc_elmmfg(nparts, parts);
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    size_t n = (size_t)p->c.nshg * 5;
    if (p->c.iprec != 0) orc_i3lu(&p->c, p->BDiag, p->res, 0);
    orc_i3lu(&p->c, p->BDiag, p->res, 1);
    orc_i3lu(&p->c, p->BDiag, p->rmes, 1);
    memset(p->Dy, 0, sizeof(double) * n);
  }
  orc_mfg_begin(nparts, parts, *eGMRES);
  double **t = malloc(sizeof(double *) * nparts);
  for (int m = 0; m < nparts; m++) {
    size_t n = (size_t)parts[m].c.nshg * 5;
    for (size_t i = 0; i < n; i++) parts[m].temp[i] = parts[m].res[i] * parts[m].res[i];
    t[m] = parts[m].temp;
  }
  double unorm = sqrt(orc_sumgat(nparts, parts, t, 5));
  double epsM = parts[0].c.epsM;
  if (!(unorm < 100.0 * epsM * epsM) && iter == 1 && (istep % 20) == 0) { /* :150-157 */
    for (int m = 0; m < nparts; m++) t[m] = parts[m].res;
    M.eGMRES = *eGMRES = itrfdi(nparts, parts, t);
  }
  for (int m = 0; m < nparts; m++) parts[m].c.ires = 2; /* :161 */
  orc_gmres_core_mfg(nparts, parts, orc_au1mfg, au2mfg, parts[0].c.minIters, HBrg, eBrg, yBrg, Rcos,
                     Rsin, iKs, lGMRES, ntotGM);
  for (int m = 0; m < nparts; m++) {
    orc_i3lu(&parts[m].c, parts[m].BDiag, parts[m].Dy, 2); /* :357 */
    parts[m].c.ires = 3;
  }
  orc_mfg_end(nparts);
  free(t);
}

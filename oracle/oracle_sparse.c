/*
 * oracle_sparse.c -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h).
 * The block-CSR flavour of the path: genadj/Asadj, ElmGMRs + fillsparseC,
 * Spsi3pre, SparseAp, SolGMRs (common/genadj.f, common/asadj.f,
 * common/fillsparse.f:66-126,236-271, compressible/elmgmr.f:280-612,
 * spsi3pre.f, sparseap.f, solgmr.f:368-744).
 */
#include "oracle_internal.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void orc_gmres_core(int nparts, orc_part *parts,
                    void (*Ap)(int, orc_part *, double **), int minIters,
                    int restart_recompute, double *HBrg, double *eBrg,
                    double *yBrg, double *Rcos, double *Rsin, int *iKs,
                    int *lGMRES, int *ntotGM);

/* genadj + Asadj (genadj.f:1-82, asadj.f:1-59).  row_fill_list keeps the
 * reference's first-seen order; the selection sort of genadj.f:50-62 then
 * emits each row ascending.  Returns nnz_tot = icnt. */
int orc_genadj(const orc_part *p, int nnz, int *colm, int *rowp) {
  const orc_common *c = &p->c;
  int nshg = c->nshg, cap = 15 * nnz;
  int *adjcnt = calloc((size_t)nshg, sizeof(int));
  int *fill = malloc(sizeof(int) * (size_t)nshg * cap);
  for (int iblk = 0; iblk < c->nelblk; iblk++) {
    const int *lc = p->lcblk + 10 * iblk;
    int iel = lc[0], nshl = lc[9], npro = lc[10] - iel;
    const int *ien = p->ien + p->ien_off[iblk];
    for (int i = 0; i < npro; i++) {
      int ndlist[ORC_MAXSH];
      for (int j = 0; j < nshl; j++) ndlist[j] = abs(ien[i + (size_t)npro * j]);
      for (int j = 0; j < nshl; j++) {
        int jnd = ndlist[j] - 1;
        int jl = adjcnt[jnd];
        for (int k = 0; k < nshl; k++) {
          int knd = ndlist[k], broke = 0;
          for (int l = 0; l < jl; l++)
            if (fill[(size_t)jnd * cap + l] == knd) { broke = 1; break; }
          if (!broke) {
            if (jl + 1 > cap) { fprintf(stderr, "increase overflow factor in genadj\n"); abort(); }
            fill[(size_t)jnd * cap + jl++] = knd;
          }
        }
        adjcnt[jnd] = jl;
      }
    }
  }
  colm[0] = 1;
  for (int i = 0; i < nshg; i++) colm[i + 1] = colm[i] + adjcnt[i];
  int icnt = 0, ibig = 10 * nshg;
  for (int i = 0; i < nshg; i++) {
    int ncol = adjcnt[i];
    int *t = fill + (size_t)i * cap;
    for (int j = 0; j < ncol; j++) {
      int imin = t[0], mloc = 0;
      for (int l = 1; l < ncol; l++)
        if (t[l] < imin) { imin = t[l]; mloc = l; }
      if (icnt >= nnz * nshg) { fprintf(stderr, "increase nnz in genmat\n"); abort(); }
      rowp[icnt++] = imin;
      t[mloc] = ibig;
    }
  }
  free(adjcnt);
  free(fill);
  return icnt;
}

/* sparseloc (fillsparse.f:236-271), 1-based result */
static int sparseloc(const int *list, int n, int target) {
  int rowvl = 1, rowvh = n + 1;
  while (rowvh - rowvl > 1) {
    int rowv = (rowvh + rowvl) / 2;
    if (list[rowv - 1] > target) rowvh = rowv; else rowvl = rowv;
  }
  return rowvl;
}

/* fillsparseC (fillsparse.f:66-126) for one block; EGb is the block's packed
 * EGmass(npro,nedof,nedof) */
static void fillsparsec(const orc_part *p, int iblk, const double *EGb) {
  const orc_common *c = &p->c;
  const int *lc = p->lcblk + 10 * iblk;
  int iel = lc[0], nshl = lc[9], npro = lc[10] - iel, nedof = c->nedof;
  const int *ien = p->ien + p->ien_off[iblk];
  for (int e = 0; e < npro; e++)
    for (int aa = 1; aa <= nshl; aa++) {
      int i = abs(ien[e + (size_t)npro * (aa - 1)]);
      int cc = p->colm[i - 1];
      int n = p->colm[i] - cc;
      int r = (aa - 1) * 5;
      for (int b = 1; b <= nshl; b++) {
        int s = (b - 1) * 5;
        int k = sparseloc(p->rowp + (cc - 1), n, abs(ien[e + (size_t)npro * (b - 1)])) + cc - 1;
        for (int g = 1; g <= 5; g++) {
          int t = (g - 1) * 5;
          for (int f = 1; f <= 5; f++)
            p->lhsK[(t + f - 1) + (size_t)25 * (k - 1)] +=
                EGb[e + (size_t)npro * ((r + f - 1) + (size_t)nedof * (s + g - 1))];
        }
      }
    }
}

/* ElmGMRs (elmgmr.f:280-612): same as ElmGMRe but each block's EGmass is a
 * temporary that is bc3LHS'd and scattered into lhsK */
void orc_elmgmrs(int nparts, orc_part *parts) {
  for (int m = 0; m < nparts; m++) parts[m].c.ires = 1;
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
    orc_common *c = &p->c;
    size_t nshg = (size_t)c->nshg;
    int nnz_tot = p->colm[c->nshg] - 1;
    memset(p->res, 0, sizeof(double) * nshg * 5);
    if (c->lhs == 1) memset(p->lhsK, 0, sizeof(double) * 25 * (size_t)nnz_tot);
    if (c->iprec != 0) memset(p->BDiag, 0, sizeof(double) * nshg * 25);
    /* run AsIGMR/bc3LHS block by block on a one-block EGmass: emulate with
     * a private part view whose numel is the block's npro */
    for (int iblk = 0; iblk < c->nelblk; iblk++) {
      const int *lc = p->lcblk + 10 * iblk;
      int iel = lc[0], npro = lc[10] - iel;
      double *EGb = NULL;
      if (c->lhs == 1) EGb = calloc((size_t)npro * c->nedof * c->nedof, sizeof(double));
      /* shift so that EGmass(iel_global) addresses the temporary: the
       * routines index EGmass[(iel-1+e) + numel*...] */
      orc_part q = *p;
      int lcb[20];
      memcpy(lcb, lc, sizeof(int) * 20);
      lcb[10] = lcb[10] - lcb[0] + 1; /* next block start */
      lcb[0] = 1;
      q.lcblk = lcb;
      int64_t off0 = p->ien_off[iblk];
      q.ien_off = &off0;
      q.c.numel = npro;
      q.c.nelblk = 1;
      orc_asigmr(&q, 0, p->qres, p->res, p->BDiag, EGb);
      if (c->lhs == 1) {
        orc_bc3lhs_block(&q, 0, EGb);
        fillsparsec(p, iblk, EGb);
        free(EGb);
      }
    }
    if (p->aerfrc) memset(p->aerfrc + 4, 0, sizeof(double) * 10 * 1001);
    for (int iblk = 0; iblk < c->nelblb; iblk++) orc_asbmfg(p, iblk, p->res);
  }
  if (nparts > 1) {
    double **g = malloc(sizeof(double *) * nparts);
    for (int m = 0; m < nparts; m++) g[m] = parts[m].res;
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
    if (p->c.iprec != 0) orc_bc3bdg(p, p->BDiag);
  }
}

/* Spsi3pre (spsi3pre.f:41-221) */
void orc_spsi3pre(int nparts, orc_part *parts) {
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    int nshg = p->c.nshg;
#define K(f, k) p->lhsK[((f)-1) + (size_t)25 * ((k)-1)]
#define BD(i, a, b) p->BDiag[((i)-1) + (size_t)nshg * (((a)-1) + 5 * ((b)-1))]
    for (int i = 1; i <= nshg; i++)
      for (int k = p->colm[i - 1]; k <= p->colm[i] - 1; k++)
        for (int g = 0; g < 5; g++) { /* column g of the block: entries 5g+1..5g+5 */
          int o = 5 * g;
          K(o + 2, k) = K(o + 2, k) - BD(i, 2, 1) * K(o + 1, k);
          K(o + 3, k) = K(o + 3, k) - BD(i, 3, 1) * K(o + 1, k) - BD(i, 3, 2) * K(o + 2, k);
          K(o + 4, k) = K(o + 4, k) - BD(i, 4, 1) * K(o + 1, k) - BD(i, 4, 2) * K(o + 2, k) -
                        BD(i, 4, 3) * K(o + 3, k);
          K(o + 5, k) = K(o + 5, k) - BD(i, 5, 1) * K(o + 1, k) - BD(i, 5, 2) * K(o + 2, k) -
                        BD(i, 5, 3) * K(o + 3, k) - BD(i, 5, 4) * K(o + 4, k);
        }
    for (int i = 1; i <= nshg; i++)
      for (int k = p->colm[i - 1]; k <= p->colm[i] - 1; k++) {
        int j = p->rowp[k - 1];
        for (int f = 1; f <= 5; f++) { /* row f of the block: entries f, f+5, ... */
          K(f, k) = BD(j, 1, 1) * K(f, k);
          K(f + 5, k) = BD(j, 2, 2) * (K(f + 5, k) - BD(j, 1, 2) * K(f, k));
          K(f + 10, k) = BD(j, 3, 3) * (K(f + 10, k) - BD(j, 1, 3) * K(f, k) - BD(j, 2, 3) * K(f + 5, k));
          K(f + 15, k) = BD(j, 4, 4) * (K(f + 15, k) - BD(j, 1, 4) * K(f, k) - BD(j, 2, 4) * K(f + 5, k) -
                                        BD(j, 3, 4) * K(f + 10, k));
          K(f + 20, k) = BD(j, 5, 5) * (K(f + 20, k) - BD(j, 1, 5) * K(f, k) - BD(j, 2, 5) * K(f + 5, k) -
                                        BD(j, 3, 5) * K(f + 10, k) - BD(j, 4, 5) * K(f + 15, k));
        }
      }
#undef K
#undef BD
  }
}

/* SparseAp (sparseap.f:26-135) */
void orc_sparseap(int nparts, orc_part *parts, double **u) {
  orc_commu(nparts, parts, u, 5, 1);
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    int nshg = p->c.nshg;
    double *pp = u[m];
    for (int j = 0; j < 5; j++)
      for (int i = 0; i < nshg; i++) pp[i + (size_t)nshg * j] = pp[(p->iper[i] - 1) + (size_t)nshg * j];
    double *q = calloc((size_t)nshg * 5, sizeof(double));
    for (int i = 1; i <= nshg; i++) {
      double tmp[5] = {0, 0, 0, 0, 0};
      for (int k = p->colm[i - 1]; k <= p->colm[i] - 1; k++) {
        int j = p->rowp[k - 1] - 1;
        const double *Kk = p->lhsK + (size_t)25 * (k - 1);
        for (int f = 0; f < 5; f++)
          tmp[f] = tmp[f] + Kk[f] * pp[j] + Kk[f + 5] * pp[j + (size_t)nshg] +
                   Kk[f + 10] * pp[j + (size_t)nshg * 2] + Kk[f + 15] * pp[j + (size_t)nshg * 3] +
                   Kk[f + 20] * pp[j + (size_t)nshg * 4];
      }
      for (int f = 0; f < 5; f++) q[(i - 1) + (size_t)nshg * f] += tmp[f];
    }
    memcpy(pp, q, sizeof(double) * (size_t)nshg * 5);
    free(q);
  }
  orc_commu(nparts, parts, u, 5, 0);
  for (int m = 0; m < nparts; m++) {
    /* zero the rows owned elsewhere (sparseap.f:113-133) */
    const orc_part *p = &parts[m];
    if (p->c.numpe <= 1) continue;
    const int *il = p->ilwork;
    int numtask = il[0], itk = 1, nshg = p->c.nshg;
    for (int t = 0; t < numtask; t++) {
      int iacc = il[itk + 1], numseg = il[itk + 3];
      if (iacc == 0)
        for (int is = 0; is < numseg; is++) {
          int b = il[itk + 4 + 2 * is], ln = il[itk + 5 + 2 * is];
          for (int k = 0; k < ln; k++)
            for (int j = 0; j < 5; j++) u[m][(b + k - 1) + (size_t)nshg * j] = 0.0;
        }
      itk += 4 + 2 * numseg;
    }
  }
}

/* SolGMRs (solgmr.f:440-744) */
void orc_solgmrs(int nparts, orc_part *parts, double *HBrg, double *eBrg,
                 double *yBrg, double *Rcos, double *Rsin, int *iKs,
                 int *lGMRES, int *ntotGM) {
  orc_elmgmrs(nparts, parts);
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    size_t n = (size_t)p->c.nshg * 5;
    if (p->rmes) memcpy(p->rmes, p->res, sizeof(double) * n);
    if (p->c.iprec != 0) orc_i3lu(&p->c, p->BDiag, p->res, 0);
  }
  if (parts[0].c.iprec != 0 && nparts > 1) { /* commu(BDiag,'out') (:473-475) */
    double **g = malloc(sizeof(double *) * nparts);
    for (int m = 0; m < nparts; m++) g[m] = parts[m].BDiag;
    orc_commu(nparts, parts, g, 25, 1);
    free(g);
  }
  for (int m = 0; m < nparts; m++) {
    orc_part *p = &parts[m];
    orc_i3lu(&p->c, p->BDiag, p->res, 1);
    memset(p->Dy, 0, sizeof(double) * (size_t)p->c.nshg * 5);
  }
  if (parts[0].c.lhs == 1) orc_spsi3pre(nparts, parts); /* guarded here (:495), SURVEY B4 */
  /* restart recomputation is keyed on the EBE counter lGMRES of COMMON
   * (solgmr.f:526), which is 0 when SolGMRs runs: never taken */
  orc_gmres_core(nparts, parts, orc_sparseap, parts[0].c.minIters, 0, HBrg, eBrg, yBrg, Rcos, Rsin, iKs,
                 lGMRES, ntotGM);
  for (int m = 0; m < nparts; m++) orc_i3lu(&parts[m].c, parts[m].BDiag, parts[m].Dy, 2);
}

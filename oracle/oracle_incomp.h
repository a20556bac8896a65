/* oracle_incomp.h -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h): the incompressible
 * element assembly into block-CSR and the lesSparse matrix-vector products. */
#ifndef ORACLE_INCOMP_H
#define ORACLE_INCOMP_H
#include "phasta_oracle.h"

/* the COMMON scalars the incompressible path reads beyond orc_common:
 * /solpar/ iconvflow, /genpar/ itau idiff ipord lhs, /matdat/ datmat(1,1,1) rho,
 * datmat(1,2,1) mu, matflg(5,1) + datmat(1:3,5,1) body force, /timdat/ flmpl flmpr
 * Delt(itseq) Dtgl almi alfi gami, /genpar/ dtsfct taucfct (common.h:184-255) */
typedef struct orc_incomp {
  int iconvflow, itau, idiff, ipord, lhs, matflg5;
  double rho, rmu, bf[3];
  double flmpl, flmpr, Delt, Dtgl, almi, alfi, gami, dtsfct, taucfct;
  /* boundary integral (incompressible/e3b.f, e3bvar.f): /nomodule/ iviscflux, /turbvari/ itwmod,
   * /aerfrc/ nsrflist(0:MAXSURF) (NULL = no surface in the flux list); ideformwall = 0 */
  int iviscflux, itwmod;
  const int *nsrflist;
} orc_incomp;

#ifdef __cplusplus
extern "C" {
#endif
int orc_sizeof_incomp(void);
void orc_inc_elmgmr(int nparts, orc_part *parts, const orc_incomp *ip, double **res, double **lhsK, double **lhsP,
                    double **xKebe, double **xGoC);
void orc_inc_bc3res(orc_part *p, double *res);
void orc_les_apg(int n, const int *col, const int *row, const double *pLhs, const double *p, double *q);
void orc_les_apkg(int n, const int *col, const int *row, const double *kLhs, const double *pLhs, const double *p,
                  double *q);
void orc_les_apngt(int n, const int *col, const int *row, const double *pLhs, const double *p, double *q, int withC);
void orc_les_apfull(int n, const int *col, const int *row, const double *kLhs, const double *pLhs, const double *p,
                    double *q);
#ifdef __cplusplus
}
#endif
#endif

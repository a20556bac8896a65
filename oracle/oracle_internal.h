/* oracle_internal.h -- TEST INFRASTRUCTURE ONLY (see phasta_oracle.h). */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H
#include "phasta_oracle.h"
#include <stddef.h>

/* Fortran-layout accessors, all indices 1-based */
#define QWT(c, lcsyst, intp) \
  ((c)->Qwt[((lcsyst)-1) + ORC_MAXTOP * ((intp)-1)])
#define QWTB(c, lcsyst, intp) \
  ((c)->Qwtb[((lcsyst)-1) + ORC_MAXTOP * ((intp)-1)])
#define SHP(p, lcsyst, n, intp) \
  ((p)->shp[((lcsyst)-1) + ORC_MAXTOP * (((n)-1) + ORC_MAXSH * ((intp)-1))])
#define SHGL(p, lcsyst, i, n, intp)                      \
  ((p)->shgl[((lcsyst)-1) +                              \
             ORC_MAXTOP * (((i)-1) + 3 * (((n)-1) + ORC_MAXSH * ((intp)-1)))])
#define SHPB(p, lcsyst, n, intp) \
  ((p)->shpb[((lcsyst)-1) + ORC_MAXTOP * (((n)-1) + ORC_MAXSH * ((intp)-1))])
#define SHGLB(p, lcsyst, i, n, intp)                      \
  ((p)->shglb[((lcsyst)-1) +                              \
              ORC_MAXTOP * (((i)-1) + 3 * (((n)-1) + ORC_MAXSH * ((intp)-1)))])

void orc_asigmr(const orc_part *p, int iblk, const double *qres, double *res,
                double *BDiag, double *EGmass);
void orc_asiq(const orc_part *p, int iblk, double *qres, double *rmass);
void orc_asbmfg(const orc_part *p, int iblk, double *res);
void orc_bc3lhs_block(const orc_part *p, int iblk, double *EGmass);
void orc_bc3res(const orc_part *p, double *res);
void orc_bc3bdg(const orc_part *p, double *BDiag);
void orc_qpbc(int nparts, orc_part *parts);

#endif

/* Test stand-in for the Fortran executable (tests/test_fortran_abi.py): DEFINES the COMMON blocks that
 * libphb200_f.so reads (phasta_b200/csrc/fortran_commons.h) and fills them from the two structs the rest of the
 * suite already uses, i.e. the inverse of fortran_abi.c's fill_common / fill_step.  Not product code. */
#define PHB_COMMON_EXTERN
#include <string.h>
#include "../../phasta_b200/csrc/fortran_commons.h"
#include "../../include/phb200.h"

void drv_fill_commons(const phb200_common *c, const phb200_step *st, const int *lcblk, const int *lcblkb, int nnz_tot) {
  conpar_.nshg = c->nshg; conpar_.numnp = c->numnp; conpar_.numel = c->numel; conpar_.numelb = c->numelb;
  conpar_.nflow = c->nflow; conpar_.ndof = c->ndof; conpar_.navier = c->Navier; conpar_.nnz_tot = nnz_tot;
  conpar_.nedof = c->nedof;
  genpar_.ndofBC = c->ndofBC; shpdat_.nshape = c->nshape;
  elmpar_.nelblk = c->nelblk; elmpar_.nelblb = c->nelblb; fronts_.nlwork = c->nlwork;
  workfc_.numpe = c->numpe; workfc_.myrank = c->myrank; workfc_.master = 0;
  genpar_.ipord = c->ipord; genpar_.idiff = c->idiff; genpar_.itau = c->itau;
  genpar_.iremoveStabTimeTerm = c->iremoveStabTimeTerm; genpar_.EntropyPressure = c->EntropyPressure;
  solpar_.iDC = c->iDC; solpar_.Kspace = c->Kspace; solpar_.nGMRES = c->nGMRES; incomp_.minIters = c->minIters;
  matdat_.matflg[0][1] = c->matflg2; matdat_.matflg[0][2] = c->matflg3;
  mmatpar_.Rgas = c->Rgas; mmatpar_.gamma = c->gamma; mmatpar_.gamma1 = c->gamma1; mmatpar_.pr = c->pr;
  matdat_.datmat[0][1][0] = c->datmat121; matdat_.datmat[0][1][1] = c->datmat221;
  matdat_.datmat[0][1][2] = c->datmat321; matdat_.datmat[0][2][0] = c->datmat131;
  precis_.epsM = c->epsM; genpar_.dtsfct = c->dtsfct; genpar_.taucfct = c->taucfct; outpar_.temper = c->temper;
  for (int i = 0; i < MAXTOP; i++) {
    intpt_.nint[i] = c->nint[i];
    intpt_.nintb[i] = c->nintb[i];
  }
  memcpy(intpt_.Qwt, c->Qwt, sizeof(c->Qwt));
  memcpy(intpt_.Qwtb, c->Qwtb, sizeof(c->Qwtb));
  memcpy(blkdat_.lcblk, lcblk, sizeof(int) * 10 * (c->nelblk + 1));
  if (lcblkb) memcpy(blkdat_.lcblkb, lcblkb, sizeof(int) * 10 * (c->nelblb + 1));
  genpar_.lhs = st->lhs; genpar_.iprec = st->iprec;
  timdat_.iter = st->iter; timdat_.nitr = st->nitr; timdat_.lstep = st->lstep; timdat_.istep = st->istep;
  timdat_.Dtgl = st->Dtgl; timdat_.almi = st->almi; timdat_.alfi = st->alfi; timdat_.gami = st->gami;
  timdat_.etol = st->etol;
  memset(&itrpar_, 0, sizeof(itrpar_));
}
/* iKs, lGMRES, ntotGM, iKss, lGMRESs, ntotGMs */
void drv_get_itrpar(int *out, double *eGMRES) {
  out[0] = itrpar_.iKs; out[1] = itrpar_.lGMRES; out[2] = itrpar_.ntotGM;
  out[3] = itrpar_.iKss; out[4] = itrpar_.lGMRESs; out[5] = itrpar_.ntotGMs;
  *eGMRES = itrpar_.eGMRES;
}

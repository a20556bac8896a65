/* CPU stand-in for libphb200.so (tests/test_fortran_abi.py, no GPU): records what fortran_abi.c hands to the C-ABI so
 * that the COMMON-block -> phb200_common / phb200_step mapping, the block registration and the /itrpar/ write-back
 * can be checked where there is no device.  Not product code; never linked into the product. */
#include <string.h>
#include "../../include/phb200.h"

struct phb200_ctx { int dummy; };
static struct phb200_ctx g_one;
phb200_common stub_common;
phb200_step stub_step;
int stub_calls[8];            /* init, solgmre, solgmrs, solmfg, set_sparse, finalize, comm_init */
const void *stub_ptrs[16];    /* lcblk, mien[0], mien[last], x, iBC, BC, iper, ilwork, shp, shgl, shpb, shglb, colm, rowp */
int stub_nnz_tot, stub_device;

int phb200_init(phb200_ctx **ctx, const phb200_common *c, const int *lcblk, const int *const *mien, const int *lcblkb,
                const int *const *mienb, const int *const *miBCB, const double *const *mBCB, const double *x,
                const int *iBC, const double *BC, const int *iper, const int *ilwork, const double *shp,
                const double *shgl, const double *shpb, const double *shglb, int device) {
  (void)lcblkb; (void)mienb; (void)miBCB; (void)mBCB;
  stub_common = *c;
  stub_calls[0]++;
  stub_ptrs[0] = lcblk; stub_ptrs[1] = mien[0]; stub_ptrs[2] = mien[c->nelblk - 1];
  stub_ptrs[3] = x; stub_ptrs[4] = iBC; stub_ptrs[5] = BC; stub_ptrs[6] = iper; stub_ptrs[7] = ilwork;
  stub_ptrs[8] = shp; stub_ptrs[9] = shgl; stub_ptrs[10] = shpb; stub_ptrs[11] = shglb;
  stub_device = device;
  *ctx = &g_one;
  return 0;
}
void phb200_finalize(phb200_ctx *ctx) { (void)ctx; stub_calls[5]++; }
int phb200_nccl_unique_id(void *id128) { memset(id128, 7, 128); return 0; }
int phb200_comm_init(phb200_ctx *ctx, const void *id128) { (void)ctx; stub_calls[6] += ((const char *)id128)[5] == 7; return 0; }
int phb200_solgmre(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res, double *rmes,
                   double *BDiag, double *Dy, double *HBrg, double *eBrg, double *yBrg, double *Rcos, double *Rsin,
                   int *iKs, int *lGMRES, int *ntotGM) {
  (void)ctx; (void)y; (void)ac; (void)rmes; (void)BDiag; (void)HBrg; (void)eBrg; (void)yBrg; (void)Rcos; (void)Rsin;
  stub_step = *st;
  stub_calls[1]++;
  res[0] = 11.0; Dy[0] = 12.0;
  *iKs = 17; *lGMRES = 0; *ntotGM += 17;
  return 0;
}
int phb200_set_sparse(phb200_ctx *ctx, const int *colm, const int *rowp, int nnz_tot) {
  (void)ctx;
  stub_calls[4]++;
  stub_ptrs[12] = colm; stub_ptrs[13] = rowp; stub_nnz_tot = nnz_tot;
  return 0;
}
int phb200_solgmrs(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res, double *rmes,
                   double *BDiag, double *Dy, double *HBrg, double *eBrg, double *yBrg, double *Rcos, double *Rsin,
                   int *iKs, int *lGMRESs, int *ntotGM) {
  (void)ctx; (void)y; (void)ac; (void)rmes; (void)BDiag; (void)HBrg; (void)eBrg; (void)yBrg; (void)Rcos; (void)Rsin;
  stub_step = *st;
  stub_calls[2]++;
  res[0] = 21.0; Dy[0] = 22.0;
  *iKs = 9; *lGMRESs = 1; *ntotGM += 9;
  return 0;
}
int phb200_solmfg(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res, double *BDiag,
                  double *Dy, double *HBrg, int *iKs, int *lGMRES, int *ntotGM, double *eGMRES) {
  (void)ctx; (void)y; (void)ac; (void)BDiag; (void)HBrg;
  stub_step = *st;
  stub_calls[3]++;
  res[0] = 31.0; Dy[0] = 32.0;
  *iKs = 5; *lGMRES = 0; *ntotGM += 5; *eGMRES = 2.0 * *eGMRES;
  return 0;
}

/*
 * phb200_fortran.h -- the entry points of libphb200_f.so (phasta_b200/csrc/fortran_abi.c): the reference's OWN solver
 * routines by name and argument list, gfortran external-procedure convention (lower case + underscore, every
 * argument by reference), on top of the C-ABI of phb200.h.  A Fortran caller needs no declaration at all -- these
 * prototypes document the boundary for C readers and let tests/test_fortran_abi.py check that the library exports
 * every one of them.  Hidden inputs: the COMMON blocks of phSolver/common/common.h (mirrors:
 * phasta_b200/csrc/fortran_commons.h); hidden outputs: COMMON /itrpar/.  See INTEGRATION.md section 2b.
 */
#ifndef PHB200_FORTRAN_H
#define PHB200_FORTRAN_H
#ifdef __cplusplus
extern "C" {
#endif

/* subroutine SolGMRe (y, ac, yold, acold, x, iBC, BC, EGmass, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin, iper, ilwork,
 *                     shp, shgl, shpb, shglb, Dy, rerr)     phSolver/compressible/solgmr.f:1-5; call site itrdrv.f:515-524
 * /itrpar/ iKs, lGMRES, ntotGM are written.  EGmass is not touched (the element matrices stay in HBM). */
void solgmre_(double *y, double *ac, double *yold, double *acold, double *x, int *iBC, double *BC, double *EGmass,
              double *res, double *BDiag, double *HBrg, double *eBrg, double *yBrg, double *Rcos, double *Rsin,
              int *iper, int *ilwork, double *shp, double *shgl, double *shpb, double *shglb, double *Dy,
              double *rerr);

/* subroutine SolGMRs (y, ac, yold, acold, x, iBC, BC, col, row, lhsk, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin, iper,
 *                     ilwork, shp, shgl, shpb, shglb, Dy, rerr)   solgmr.f:368-373; call site itrdrv.f:477-487
 * col / row = genadj's colm / rowp (itrdrv.f:163-169); /itrpar/ iKss, lGMRESs, ntotGMs are written; lhsk is not touched. */
void solgmrs_(double *y, double *ac, double *yold, double *acold, double *x, int *iBC, double *BC, int *col, int *row,
              double *lhsk, double *res, double *BDiag, double *HBrg, double *eBrg, double *yBrg, double *Rcos,
              double *Rsin, int *iper, int *ilwork, double *shp, double *shgl, double *shpb, double *shglb,
              double *Dy, double *rerr);

/* subroutine SolMFG (y, ac, yold, acold, x, iBC, BC, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin, iper, ilwork, shp, shgl,
 *                    shpb, shglb, Dy, rerr)                 phSolver/compressible/solmfg.f:1-5; call site itrdrv.f:496-505
 * /itrpar/ iKs, lGMRES, ntotGM and eGMRES (in and out, itrfdi.f:139) are written. */
void solmfg_(double *y, double *ac, double *yold, double *acold, double *x, int *iBC, double *BC, double *res,
             double *BDiag, double *HBrg, double *eBrg, double *yBrg, double *Rcos, double *Rsin, int *iper,
             int *ilwork, double *shp, double *shgl, double *shpb, double *shglb, double *Dy, double *rerr);

/* The one thing C cannot reach: mien(iblk)%p of module pointer_data (common/pointer.f:42-47).  Called once per block
 * after genblk / genbkb by the five-line Fortran routine of INTEGRATION.md 2b.  iblk is 1-based. */
void phb200_register_block_(const int *iblk, const int *ien);
void phb200_register_blockb_(const int *iblk, const int *ienb, const int *iBCB, const double *BCB);

/* Multi-rank runs (one MPI rank per GPU): the master asks for the 128-byte NCCL id, the Fortran side broadcasts it
 * with MPI_BCAST and every rank hands it back before the first solve (replaces MPI_COMM_WORLD of commu.f /
 * mpitools.f for the halo exchange and the dot products). */
void phb200_fortran_unique_id_(unsigned char *id128);
void phb200_fortran_comm_id_(const unsigned char *id128);

/* before MPI_Finalize: releases the device context */
void phb200_fortran_finalize_(void);

#ifdef __cplusplus
}
#endif
#endif

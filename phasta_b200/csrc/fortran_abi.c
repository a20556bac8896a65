/*
 * fortran_abi.c -- the reference's OWN entry points, by name and argument list, on top of libphb200.so:
 *
 *     solgmre_   phSolver/compressible/solgmr.f:1-5      (call site itrdrv.f:515-524)
 *     solgmrs_   phSolver/compressible/solgmr.f:368-373  (call site itrdrv.f:477-487)
 *     solmfg_    phSolver/compressible/solmfg.f:1-5      (call site itrdrv.f:496-505)
 *
 * with gfortran's external-procedure convention (lower case + underscore, every argument by reference), exactly as
 * SolGMRp does for the PETSc flavour (phSolver/compressible/solgmrpetsc.c:59-65).  A PHASTA build that drops
 * solgmr.f / solmfg.f from its source list and links libphb200_f.so + libphb200.so runs itrdrv.f unchanged.
 *
 * The hidden inputs are read where the Fortran routines read them: the COMMON blocks of common.h (their C mirrors
 * follow common_c.h:86-665; /blkdat/, /intpt/ and /shpdat/ have no mirror there and follow common.h:92-96,111,172).
 * The one input C cannot reach is the module array mien(iblk)%p of `use pointer_data` (pointer.f:42-47): the
 * build adds the five-line Fortran routine of INTEGRATION.md section 2b that hands every block pointer to
 * phb200_register_block_ once, after genblk.
 *
 * The device context is created on the first solve (the mesh, BC and table arguments arrive with every call,
 * solgmr.f:1-8, and are fixed for a run) and lives until phb200_fortran_finalize_.  Plain C: this file is the host
 * side of the boundary and needs no CUDA header.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/phb200.h"

#include "fortran_commons.h"

/* ---- block pointers handed over by the Fortran side (pointer_data is not reachable from C) ---- */
static const int **g_mien, **g_mienb, **g_miBCB;
static const double **g_mBCB;
static int g_cap;
static phb200_ctx *g_ctx;
static int g_sparse_set;
static unsigned char g_nccl_id[128];
static int g_have_id;

static void die(const char *where, const char *what) {
  /* the reference's error() (common/error.f) prints and stops the run; so does the drop-in */
  fprintf(stderr, "phb200: %s: %s\n", where, what);
  fflush(stderr);
  abort();
}
static void grow(int n) {
  if (n <= g_cap) return;
  int cap = g_cap ? g_cap : 1024;
  while (cap < n) cap *= 2;
  g_mien = (const int **)realloc(g_mien, sizeof(*g_mien) * cap);
  g_mienb = (const int **)realloc(g_mienb, sizeof(*g_mienb) * cap);
  g_miBCB = (const int **)realloc(g_miBCB, sizeof(*g_miBCB) * cap);
  g_mBCB = (const double **)realloc(g_mBCB, sizeof(*g_mBCB) * cap);
  if (!g_mien || !g_mienb || !g_miBCB || !g_mBCB) die("register_block", "out of memory");
  for (int i = g_cap; i < cap; i++) {
    g_mien[i] = g_mienb[i] = g_miBCB[i] = NULL;
    g_mBCB[i] = NULL;
  }
  g_cap = cap;
}
/* call phb200_register_block(iblk, mien(iblk)%p)           -- interior block iblk (1-based) */
void phb200_register_block_(const int *iblk, const int *ien) {
  if (*iblk < 1 || *iblk > MAXBLK) die("register_block", "iblk out of range");
  grow(*iblk);
  g_mien[*iblk - 1] = ien;
}
/* call phb200_register_blockb(iblk, mienb(iblk)%p, miBCB(iblk)%p, mBCB(iblk)%p) -- boundary block */
void phb200_register_blockb_(const int *iblk, const int *ienb, const int *iBCB, const double *BCB) {
  if (*iblk < 1 || *iblk > MAXBLK) die("register_blockb", "iblk out of range");
  grow(*iblk);
  g_mienb[*iblk - 1] = ienb;
  g_miBCB[*iblk - 1] = iBCB;
  g_mBCB[*iblk - 1] = BCB;
}

/* Multi-rank runs (one MPI rank per GPU): rank 0 asks for the NCCL id, the Fortran side broadcasts the 128 bytes
 * with MPI_Bcast and every rank hands them back before the first solve:
 *     if (myrank.eq.master) call phb200_fortran_unique_id(id)
 *     call MPI_BCAST(id, 128, MPI_BYTE, master, MPI_COMM_WORLD, ierr)
 *     call phb200_fortran_comm_id(id)                                                                          */
void phb200_fortran_unique_id_(unsigned char *id128) {
  if (phb200_nccl_unique_id(id128)) die("unique_id", "NCCL is not available");
}
void phb200_fortran_comm_id_(const unsigned char *id128) {
  memcpy(g_nccl_id, id128, 128);
  g_have_id = 1;
}

static void fill_common(phb200_common *c) {
  memset(c, 0, sizeof(*c));
  c->nshg = conpar_.nshg; c->numnp = conpar_.numnp; c->numel = conpar_.numel; c->numelb = conpar_.numelb;
  c->nflow = conpar_.nflow; c->ndof = conpar_.ndof; c->ndofBC = genpar_.ndofBC;
  c->nshape = shpdat_.nshape; c->nedof = conpar_.nflow * shpdat_.nshape;   /* itrdrv.f:513 */
  c->nelblk = elmpar_.nelblk; c->nelblb = elmpar_.nelblb; c->nlwork = fronts_.nlwork;
  c->numpe = workfc_.numpe; c->myrank = workfc_.myrank;
  c->ipord = genpar_.ipord; c->idiff = genpar_.idiff; c->itau = genpar_.itau;
  c->iremoveStabTimeTerm = genpar_.iremoveStabTimeTerm; c->EntropyPressure = genpar_.EntropyPressure;
  c->iDC = solpar_.iDC; c->Navier = conpar_.navier; c->Kspace = solpar_.Kspace; c->nGMRES = solpar_.nGMRES;
  c->minIters = incomp_.minIters;
  c->matflg2 = matdat_.matflg[0][1]; c->matflg3 = matdat_.matflg[0][2];      /* matflg(2,1), matflg(3,1) */
  c->Rgas = mmatpar_.Rgas; c->gamma = mmatpar_.gamma; c->gamma1 = mmatpar_.gamma1; c->pr = mmatpar_.pr;
  c->datmat121 = matdat_.datmat[0][1][0]; c->datmat221 = matdat_.datmat[0][1][1];   /* datmat(1:3,2,1) */
  c->datmat321 = matdat_.datmat[0][1][2]; c->datmat131 = matdat_.datmat[0][2][0];   /* datmat(1,3,1)   */
  c->epsM = precis_.epsM; c->dtsfct = genpar_.dtsfct; c->taucfct = genpar_.taucfct; c->temper = outpar_.temper;
  for (int i = 0; i < MAXTOP; i++) {
    c->nint[i] = intpt_.nint[i];
    c->nintb[i] = intpt_.nintb[i];
  }
  memcpy(c->Qwt, intpt_.Qwt, sizeof(c->Qwt));      /* Qwt(MAXTOP,MAXQPT): same column-major image */
  memcpy(c->Qwtb, intpt_.Qwtb, sizeof(c->Qwtb));
}
static void fill_step(phb200_step *st) {
  st->lhs = genpar_.lhs; st->iprec = genpar_.iprec; st->iter = timdat_.iter; st->nitr = timdat_.nitr;
  st->lstep = timdat_.lstep; st->istep = timdat_.istep;
  st->Dtgl = timdat_.Dtgl; st->almi = timdat_.almi; st->alfi = timdat_.alfi; st->gami = timdat_.gami;
  st->etol = timdat_.etol;
}

static phb200_ctx *context(const double *x, const int *iBC, const double *BC, const int *iper, const int *ilwork,
                           const double *shp, const double *shgl, const double *shpb, const double *shglb) {
  if (g_ctx) return g_ctx;
  phb200_common c;
  fill_common(&c);
  if (g_cap < c.nelblk) die("solgmr", "interior blocks were not registered (phb200_register_block, INTEGRATION.md 2b)");
  for (int b = 0; b < c.nelblk; b++)
    if (!g_mien[b]) die("solgmr", "an interior block was not registered");
  for (int b = 0; b < c.nelblb; b++)
    if (g_cap <= b || !g_mienb[b]) die("solgmr", "a boundary block was not registered");
  const char *dev = getenv("PHB200_DEVICE"); /* one process per GPU: default = local rank modulo visible devices */
  int device = dev ? atoi(dev) : -1;
  if (device < 0) {
    const char *lr = getenv("OMPI_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = getenv("MV2_COMM_WORLD_LOCAL_RANK");
    if (!lr) lr = getenv("SLURM_LOCALID");
    device = lr ? atoi(lr) : 0;
  }
  if (phb200_init(&g_ctx, &c, &blkdat_.lcblk[0][0], g_mien, c.nelblb ? &blkdat_.lcblkb[0][0] : NULL, g_mienb, g_miBCB,
                  g_mBCB, x, iBC, BC, iper, ilwork, shp, shgl, shpb, shglb, device))
    die("solgmr", "phb200_init failed");
  if (c.numpe > 1) {
    if (!g_have_id) die("solgmr", "numpe > 1 but no NCCL id was handed over (phb200_fortran_comm_id)");
    if (phb200_comm_init(g_ctx, g_nccl_id)) die("solgmr", "phb200_comm_init failed");
  }
  return g_ctx;
}

/* subroutine SolGMRe (y, ac, yold, acold, x, iBC, BC, EGmass, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin, iper,
 *                     ilwork, shp, shgl, shpb, shglb, Dy, rerr)                               solgmr.f:1-5
 * EGmass(numel,nedof,nedof) stays in HBM (12.9 GB per 4 M tets); the caller's array is not written.  yold, acold
 * and rerr are not read by the reference's routine either (solgmr.f:95-347). */
void solgmre_(double *y, double *ac, double *yold, double *acold, double *x, int *iBC, double *BC, double *EGmass,
              double *res, double *BDiag, double *HBrg, double *eBrg, double *yBrg, double *Rcos, double *Rsin,
              int *iper, int *ilwork, double *shp, double *shgl, double *shpb, double *shglb, double *Dy,
              double *rerr) {
  (void)yold; (void)acold; (void)EGmass; (void)rerr;
  phb200_ctx *ctx = context(x, iBC, BC, iper, ilwork, shp, shgl, shpb, shglb);
  phb200_step st;
  fill_step(&st);
  if (phb200_solgmre(ctx, y, ac, &st, res, NULL, BDiag, Dy, HBrg, eBrg, yBrg, Rcos, Rsin, &itrpar_.iKs,
                     &itrpar_.lGMRES, &itrpar_.ntotGM))
    die("solgmre", "solve failed");
}

/* subroutine SolGMRs (y, ac, yold, acold, x, iBC, BC, col, row, lhsk, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin,
 *                     iper, ilwork, shp, shgl, shpb, shglb, Dy, rerr)                         solgmr.f:368-373
 * col / row are genadj's colm(nshg+1) / rowp(nnz*nshg) (itrdrv.f:163-169); lhsk(nflow*nflow,nnz_tot) stays in HBM. */
void solgmrs_(double *y, double *ac, double *yold, double *acold, double *x, int *iBC, double *BC, int *col, int *row,
              double *lhsk, double *res, double *BDiag, double *HBrg, double *eBrg, double *yBrg, double *Rcos,
              double *Rsin, int *iper, int *ilwork, double *shp, double *shgl, double *shpb, double *shglb,
              double *Dy, double *rerr) {
  (void)yold; (void)acold; (void)lhsk; (void)rerr;
  phb200_ctx *ctx = context(x, iBC, BC, iper, ilwork, shp, shgl, shpb, shglb);
  if (!g_sparse_set) {
    if (phb200_set_sparse(ctx, col, row, conpar_.nnz_tot)) die("solgmrs", "set_sparse failed");
    g_sparse_set = 1;
  }
  phb200_step st;
  fill_step(&st);
  if (phb200_solgmrs(ctx, y, ac, &st, res, NULL, BDiag, Dy, HBrg, eBrg, yBrg, Rcos, Rsin, &itrpar_.iKss,
                     &itrpar_.lGMRESs, &itrpar_.ntotGMs))
    die("solgmrs", "solve failed");
}

/* subroutine SolMFG (y, ac, yold, acold, x, iBC, BC, res, BDiag, HBrg, eBrg, yBrg, Rcos, Rsin, iper, ilwork,
 *                    shp, shgl, shpb, shglb, Dy, rerr)                                         solmfg.f:1-5 */
void solmfg_(double *y, double *ac, double *yold, double *acold, double *x, int *iBC, double *BC, double *res,
             double *BDiag, double *HBrg, double *eBrg, double *yBrg, double *Rcos, double *Rsin, int *iper,
             int *ilwork, double *shp, double *shgl, double *shpb, double *shglb, double *Dy, double *rerr) {
  (void)yold; (void)acold; (void)rerr;
  phb200_ctx *ctx = context(x, iBC, BC, iper, ilwork, shp, shgl, shpb, shglb);
  phb200_step st;
  fill_step(&st);
  (void)eBrg; (void)yBrg; (void)Rcos; (void)Rsin;   /* scratch of the Fortran routine; nothing reads them after it */
  if (phb200_solmfg(ctx, y, ac, &st, res, BDiag, Dy, HBrg, &itrpar_.iKs, &itrpar_.lGMRES, &itrpar_.ntotGM,
                    &itrpar_.eGMRES))                /* COMMON /itrpar/ eGMRES in and out (itrfdi.f:139) */
    die("solmfg", "solve failed");
}

/* call phb200_fortran_finalize() -- before MPI_Finalize */
void phb200_fortran_finalize_(void) {
  if (g_ctx) phb200_finalize(g_ctx);
  g_ctx = NULL;
  g_sparse_set = 0;
}

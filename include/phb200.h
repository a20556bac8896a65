/*
 * phb200.h -- C ABI of libphb200.so: the B200-native drop-in for PHASTA's
 * compressible implicit-step hot path (element assembly + EBE GMRES).
 *
 * The boundary is the SolGMRe call in phSolver/compressible/itrdrv.f:515-524;
 * the precedent for a C function behind that call site is SolGMRp
 * (phSolver/compressible/solgmrpetsc.c:59-65).  All arguments are plain
 * pointers to caller-owned arrays in the reference's in-memory (Fortran,
 * column-major, 1-based ids) layouts; the hidden COMMON-block inputs
 * (phSolver/common/common.h:35-268) travel in phb200_common / phb200_step.
 * INTEGRATION.md shows the ISO_C_BINDING shim that binds these entry points.
 *
 * Every function returns 0 on success; on failure it prints
 * "phb200: <routine>: <what>" to stderr (the reference's error() convention,
 * phSolver/common/error.f) and returns non-zero so the Fortran shim can call
 * error()/MPI_ABORT.  There is no CPU fallback: without a CUDA device
 * phb200_init fails.
 */
#ifndef PHB200_H
#define PHB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PHB200_MAXTOP 6   /* common.h:17-20 MAXTOP */
#define PHB200_MAXSH 32   /* MAXSH  */
#define PHB200_MAXQPT 125 /* MAXQPT */

/* Scalars fixed for a run: /conpar/ /blkdat/ /fronts/ /workfc/ /genpar/
 * /solpar/ /incomp/ /mmatpar/ /matdat/ /precis/ /outpar/ /intpt/
 * (common.h:92-96,111,121-125,184-189,220-231,247). */
typedef struct phb200_common {
  int nshg, numnp, numel, numelb, nflow, ndof, ndofBC, nshape, nedof;
  int nelblk, nelblb, nlwork, numpe, myrank;
  int ipord, idiff, itau, iremoveStabTimeTerm, EntropyPressure;
  int iDC, Navier, Kspace, nGMRES, minIters;
  int matflg2, matflg3; /* matflg(2,1) viscosity model, matflg(3,1) */
  double Rgas, gamma, gamma1, pr;
  double datmat121, datmat221, datmat321, datmat131;
  double epsM, dtsfct, taucfct, temper;
  int nint[PHB200_MAXTOP], nintb[PHB200_MAXTOP];
  double Qwt[PHB200_MAXTOP * PHB200_MAXQPT];  /* Qwt(MAXTOP,MAXQPT)  */
  double Qwtb[PHB200_MAXTOP * PHB200_MAXQPT]; /* Qwtb(MAXTOP,MAXQPT) */
} phb200_common;

/* Scalars that change per call: /timdat/ (common.h:252-255; itrPC.f:18-47)
 * and the per-solve switches itrdrv sets (itrdrv.f:456-457,511-512). */
typedef struct phb200_step {
  int lhs, iprec, iter, nitr, lstep;
  int istep; /* steps taken in this run (COMMON /timdat/): SolMFG recomputes eGMRES when mod(istep,20)==0 */
  double Dtgl, almi, alfi, gami, etol;
} phb200_step;

/* Incompressible flavour (PHASTA_INCOMPRESSIBLE; BASELINE.json configs[3]): the COMMON scalars its element
 * routines read beyond phb200_common -- /solpar/ iconvflow (input_fform.cc:650-653), /genpar/ itau idiff ipord
 * lhs dtsfct taucfct, /matdat/ datmat(1,1,1) rho, datmat(1,2,1) mu, matflg(5,1) + datmat(1:3,5,1) constant body
 * force, /timdat/ flmpl flmpr Delt(itseq) Dtgl almi alfi gami (common.h:184-255; incompressible/itrPC.f:18-26). */
typedef struct phb200_incomp {
  int iconvflow, itau, idiff, ipord, lhs, matflg5;
  double rho, rmu, bf[3];
  double flmpl, flmpr, Delt, Dtgl, almi, alfi, gami, dtsfct, taucfct;
  /* the boundary integral (incompressible/asbmfg.f, e3b.f, e3bvar.f): /nomodule/ iviscflux, /turbvari/ itwmod
   * (|itwmod| = 1 integrates Force), /aerfrc/ nsrflist(0:MAXSURF) -- the 1001 switches that put a surface ID
   * in the flux / force list (common.h:98-108); NULL = no surface listed.  Rigid walls only (ideformwall = 0). */
  int iviscflux, itwmod;
  const int *nsrflist;
} phb200_incomp;

typedef struct phb200_ctx phb200_ctx;

/* One-time setup, called after genadj in itrdrv (itrdrv.f:169).  Replaces the
 * mesh/BC/table arguments SolGMRe receives on every call (solgmr.f:1-8):
 * x, iBC, BC, iper, ilwork, shp, shgl, shpb, shglb, plus lcblk/lcblkb
 * (common.h:111) and the pointer_data block arrays mien/mienb/miBCB/mBCB
 * (pointer.f:42-47) -- mien[iblk] points at mien(iblk)%p (npro,nshl).
 * ilwork is the in-memory array after ctypes (iother 0-based, ctypes.f:47).
 * device: CUDA ordinal (one process per GPU). */
int phb200_init(phb200_ctx **ctx, const phb200_common *c, const int *lcblk,
                const int *const *mien, const int *lcblkb,
                const int *const *mienb, const int *const *miBCB,
                const double *const *mBCB, const double *x, const int *iBC,
                const double *BC, const int *iper, const int *ilwork,
                const double *shp, const double *shgl, const double *shpb,
                const double *shglb, int device);
void phb200_finalize(phb200_ctx *ctx);

/* Multi-GPU plumbing: replaces MPI_COMM_WORLD of commu.f / mpitools.f with an
 * NCCL communicator over NVLink.  Rank 0 creates the 128-byte id, the host
 * broadcasts it (MPI_Bcast in the Fortran shim, torch.distributed here). */
int phb200_nccl_unique_id(void *id128);
int phb200_comm_init(phb200_ctx *ctx, const void *id128);

/* SolGMRe (solgmr.f:1-362).  y, ac: (nshg,ndof) {u1,u2,u3,p,T}.  Outputs:
 * res (nshg,nflow) block-diagonal-preconditioned residual, rmes un-
 * preconditioned b (solgmr.f:83) (nullable), Dy (nshg,nflow) {p,u,v,w,T}
 * (= solinc), HBrg(Kspace+1,Kspace)/eBrg/yBrg/Rcos/Rsin (nullable), and the
 * /itrpar/ counters iKs, lGMRES, ntotGM (+= iterations).  EGmass and BDiag
 * stay device-resident (nothing outside SolGMRe reads them);
 * BDiag (LU factors) is copied back only if the pointer is non-null. */
int phb200_solgmre(phb200_ctx *ctx, const double *y, const double *ac,
                   const phb200_step *step, double *res, double *rmes,
                   double *BDiag, double *Dy, double *HBrg, double *eBrg,
                   double *yBrg, double *Rcos, double *Rsin, int *iKs,
                   int *lGMRES, int *ntotGM);

/* ElmGMRe (elmgmr.f:1-274): res, BDiag(nshg,5,5), EGmass(numel,nedof,nedof)
 * (any output pointer may be null).  qres(nshg,12) nullable (debug seam). */
int phb200_elmgmre(phb200_ctx *ctx, const double *y, const double *ac,
                   const phb200_step *step, double *res, double *BDiag,
                   double *EGmass, double *qres);

/* ---- block-CSR flavour: SolGMRs (solgmr.f:368-744, itrdrv.f:477-487) ----
 * genadj (common/genadj.f:1-82): colm(nshg+1) 1-based row pointers, rowp
 * ascending 1-based column ids incl. the diagonal, nnz_tot; rowp has capacity
 * nnz*nshg like the reference (input.f:151).  (The reference calls these
 * `colm`/`rowp` at the call site and `col`/`row` inside sparseap.f:18-20.)
 * Built ON THE DEVICE (csrc/genadj.cu: radix sort of the node pairs, unique, scan; 0.16 s for 32 M tets) and
 * INSTALLED as the part's CSR structure (no phb200_set_sparse needed afterwards); colm / rowp may be NULL when the
 * caller does not want the host copies. */
int phb200_genadj(phb200_ctx *ctx, int nnz, int *colm, int *rowp, int *nnz_tot);
/* hand the CSR structure itrdrv owns (itrdrv.f:167-169) to the device once */
int phb200_set_sparse(phb200_ctx *ctx, const int *colm, const int *rowp,
                      int nnz_tot);
/* ElmGMRs + fillsparseC (elmgmr.f:280-612, fillsparse.f:66-126); lhsK
 * (25,nnz_tot) out, nullable */
int phb200_elmgmrs(phb200_ctx *ctx, const double *y, const double *ac,
                   const phb200_step *step, double *res, double *BDiag,
                   double *lhsK);
/* Spsi3pre (spsi3pre.f:41-221), SparseAp (sparseap.f:26-135) */
int phb200_spsi3pre(phb200_ctx *ctx, double *lhsK);
int phb200_sparseap(phb200_ctx *ctx, double *p);
/* SolGMRs: same outputs as phb200_solgmre; lhsK stays device-resident */
int phb200_solgmrs(phb200_ctx *ctx, const double *y, const double *ac,
                   const phb200_step *step, double *res, double *rmes,
                   double *BDiag, double *Dy, double *HBrg, double *eBrg,
                   double *yBrg, double *Rcos, double *Rsin, int *iKs,
                   int *lGMRESs, int *ntotGM);
int phb200_dev_elmgmrs(phb200_ctx *ctx, const phb200_step *step);
int phb200_dev_solve_sparse(phb200_ctx *ctx, const phb200_step *step, int *iKs,
                            int *lGMRESs, int *ntotGM);
int phb200_dev_sparseap(phb200_ctx *ctx, int slot);

/* ---- incompressible flavour: ElmGMR into block-CSR + the lesSparse products ----
 * ElmGMR (incompressible/elmgmr.f:1-330 called from SolFlow, incompressible/solfar.f): AsIq/e3q + qpbc (idiff=1),
 * AsIGMR/e3 (e3ivar, e3stab itau=0, e3Res, e3LHS), bc3LHS, fillsparseI (common/fillsparse.f:1-65), halo 'in',
 * bc3Res.  y, ac (nshg,ndof) as for SolGMRe; res (nshg,4) {mom1,mom2,mom3,continuity}; lhsK (9,nnz_tot) with the
 * 3x3 block entry (r,c) at 3(r-1)+c, lhsP (4,nnz_tot) = {G1,G2,G3,C}; the CSR structure is the one of
 * phb200_set_sparse.  Output pointers may be null (the matrices stay device-resident for phb200_les_ap).
 * Boundary-element blocks (AsBMFG/e3b/e3bvar of the incompressible code, elmgmr.f:246-320) add their flux to res
 * and integrate flxID / Force (read back with phb200_get_aerfrc); deformable-wall elements (iBCB bit 4) are refused. */
int phb200_inc_elmgmr(phb200_ctx *ctx, const double *y, const double *ac, const phb200_incomp *ip, double *res,
                      double *lhsK, double *lhsP);
int phb200_inc_dev_elmgmr(phb200_ctx *ctx, const phb200_incomp *ip);
/* fLesSparseAp{G,KG,NGt,NGtC,Full} (incompressible/lesSparse.f:204-492), the matrix-vector products the
 * reference's Krylov solver calls back, on the device-resident lhsK/lhsP of the last phb200_inc_elmgmr:
 * kind 0 ApG  p(n)   -> q(n,3);  1 ApKG  p(n,4) -> q(n,3);  2 ApNGt p(n,3) -> q(n);
 *      3 ApNGtC p(n,4) -> q(n);  4 ApFull p(n,4) -> q(n,4). */
int phb200_les_ap(phb200_ctx *ctx, int kind, const double *p, double *q);
/* ApFull on device work vectors (Ap/s timing) */
int phb200_inc_dev_apfull(phb200_ctx *ctx);

/* Finer seams, all on host arrays: */
/* i3LU (i3lu.f:1-181) code 0 LU_Fact / 1 forward / 2 backward / 3 product */
int phb200_i3lu(phb200_ctx *ctx, double *Diag, double *r, int code);
/* i3pre (i3pre.f:1-146) on the device-resident EGmass/BDiag of the last
 * elmgmre; EGmass out nullable */
int phb200_i3pre(phb200_ctx *ctx, double *EGmass);
/* Au1GMR (au1gmr.f:1-106) with the resident EGmass: u(nshg,5) in/out */
int phb200_au1gmr(phb200_ctx *ctx, double *u);
/* bc3per (bc3per.f:1-46) */
int phb200_bc3per(phb200_ctx *ctx, double *r);
/* commu (commu.f:1-297): code 0 'in ', 1 'out'; global(nshg,n) */
int phb200_commu(phb200_ctx *ctx, double *global, int n, int code);
/* sumgat (mpitools.f:98-137) */
int phb200_sumgat(phb200_ctx *ctx, const double *u, int n, double *summed);

/* COMMON /aerfrc/ (common.h:106) integrated by the boundary elements
 * (e3b.f:305-345): Force(3) and HFlux accumulate over calls made with
 * step.iter==step.nitr (zero=1 resets them, as itrdrv.f:437-442 does each
 * step); flxID(10,0:MAXSURF) holds the last ElmGMRe's per-surface fluxes.
 * Any pointer may be null. */
int phb200_get_aerfrc(phb200_ctx *ctx, double *Force, double *HFlux,
                      double *flxID, int zero);

/* ---- Newton / time-step shell on the device ----
 * The predictor-multicorrector routines itrdrv calls around SolGMR*
 * (itrdrv.f:393-394,590-594): itrPredict (itrPC.f:54-119, ipred 1..4),
 * itrBC (itrbc.f:60-199; ylimit off, iabc=0), itrCorrect (itrPC.f:127-150),
 * itrUpdate (itrPC.f:205-210) act on the device-resident y/ac (phb200_set_state)
 * and yold/acold (phb200_set_old_state), so a step's vectors never cross PCIe.
 * phb200_rstat: totres(1:2) of rstat.f:94-112 from the last solve's res / rmes.
 * phb200_timestep: one whole step of itrdrv.f's flow sequence "0 1 0 1 ...":
 * predictor, nitr x (SolGMRe | SolGMRs, rstat, itrCorrect, itrBC), itrUpdate;
 * lhs = 1 - min(1, mod(ifuncs-1, LHSupd)) (itrdrv.f:456,511).  stats (nitr,6
 * row-major, nullable): totres(1), totres(2), iKs, lGMRES, lhs, 0. */
int phb200_set_old_state(phb200_ctx *ctx, const double *yold, const double *acold);
int phb200_get_state(phb200_ctx *ctx, double *y, double *ac, double *yold, double *acold);
int phb200_itrpredict(phb200_ctx *ctx, const phb200_step *step, int ipred);
int phb200_itrbc(phb200_ctx *ctx, int ires);
int phb200_itrcorrect(phb200_ctx *ctx, const phb200_step *step);
int phb200_itrupdate(phb200_ctx *ctx, const phb200_step *step);
int phb200_rstat(phb200_ctx *ctx, long long nshgt, double *totres);
int phb200_timestep(phb200_ctx *ctx, const phb200_step *step, int ipred, int nitr, int sparse, int LHSupd,
                    long long nshgt, int *ntotGM, double *stats);

/* HBM-resident path (what bench.py's `value` times: inputs already on the
 * device).  set_state uploads y/ac once; the dev_* calls run on them. */
int phb200_set_state(phb200_ctx *ctx, const double *y, const double *ac);
int phb200_dev_elmgmre(phb200_ctx *ctx, const phb200_step *step);
/* i3LU + i3pre + GMRES + back-substitution on the resident system */
int phb200_dev_solve(phb200_ctx *ctx, const phb200_step *step, int *iKs,
                     int *lGMRES, int *ntotGM);
/* one Au1GMR + bc3per on Krylov slot `slot` -> slot+1 (for Ap/s timing) */
int phb200_dev_ap(phb200_ctx *ctx, int slot);
int phb200_get_res(phb200_ctx *ctx, double *res);
int phb200_get_dy(phb200_ctx *ctx, double *Dy);
int phb200_get_bdiag(phb200_ctx *ctx, double *BDiag);
int phb200_get_egmass(phb200_ctx *ctx, double *EGmass);
/* spot reads for meshes whose whole LHS does not fit on the host: EGmass(e0+1:e0+n,:,:) of the reference's element
 * order as out(n,nedof,nedof), and lhsK(:,k0+1:k0+n) as out(25,n) */
int phb200_get_egmass_range(phb200_ctx *ctx, long long e0, int n, double *EGmass);
int phb200_get_lhsk_range(phb200_ctx *ctx, long long k0, long long n, double *lhsK);

/* Timing on the library's own stream (bench.py): record event `slot`
 * (0..15), elapsed ms between two recorded events, full sync, and the number
 * of kernel launches issued so far. */
int phb200_event_record(phb200_ctx *ctx, int slot);
int phb200_event_elapsed_ms(phb200_ctx *ctx, int a, int b, float *ms);
int phb200_sync(phb200_ctx *ctx);
long long phb200_launch_count(phb200_ctx *ctx);
/* accumulated device time (ms) and launch count of kernel class k since the
 * last reset: 0 assembly(AsIGMR), 1 AsIq, 2 Ap(EBE), 3 i3pre, 4 Krylov BLAS-1,
 * 5 node-wise BC/LU, 6 halo pack/unpack.  Timed with CUDA events only when
 * profiling is switched on (phb200_profile(ctx,1)). */
int phb200_profile(phb200_ctx *ctx, int on);
int phb200_profile_get(phb200_ctx *ctx, int k, float *ms, long long *launches);
int phb200_profile_reset(phb200_ctx *ctx);
/* FP64 FMA peak microbenchmark (MEASURED_PEAKS.json has no FP64 number):
 * returns achieved TFLOP/s of a register-resident DFMA chain kernel. */
int phb200_fp64_peak(phb200_ctx *ctx, double *tflops);
/* Deterministic assembly option (north_star: 'colored or warp-aggregated scatter-add'; the reference adds element
 * contributions in element order, local.f:67-74).  on != 0: AsIq and the lhs=1 tet assembly store their per-element
 * contributions and a node-wise gather sums them in ascending element order, so qres, res and BDiag (and EGmass,
 * which never used atomics) are bit-for-bit reproducible from run to run; costs one extra pass over 960 B/element.
 * Default off: FP64 atomics, results equal to round-off.  Parts of linear tets without boundary-element blocks. */
int phb200_set_deterministic(phb200_ctx *ctx, int on);
/* FP64 tensor-core peak (mma.sync.m8n8k4.f64 chains): the evidence for keeping 5x5 blocks on the FMA pipe */
int phb200_dmma_peak(phb200_ctx *ctx, double *tflops);
/* FP64 scatter-add microbenchmark: warp-wide red.global.add.f64 on the 25 contiguous doubles of pseudo-random
 * 200-byte blocks out of nblk (the address pattern of fillsparseC into lhsK): G doubles added per second. */
int phb200_red_peak(phb200_ctx *ctx, long long nblk, double *gadds_per_s);
/* flush L2 by writing a >126 MB scratch buffer (bench hygiene) */
int phb200_flush_l2(phb200_ctx *ctx);

const char *phb200_version(void);
/* struct sizes, so a binding can verify its mirror of the two structs */
int phb200_sizeof_common(void);
int phb200_sizeof_incomp(void);
int phb200_sizeof_step(void);
/* test transport: several parts on ONE GPU, one host thread per part (see
 * csrc/comm.cu); stands in for phb200_comm_init on a single-GPU box */
int phb200_local_group_join(phb200_ctx *ctx, int nranks);

/* ---- matrix-free flavour: SolMFG (compressible/solmfg.f:1-381), called from the same ladder in itrdrv
 * (itrdrv.f:494-505, lhs=0, iprec per LHSupd).  No EGmass / lhsK is ever stored: the Ap is a finite
 * difference of the modified residual (Au1MFG, au1mfg.f:52-90; ItrRes/AsIRes, itrres.f, asires.f) and the
 * block-diagonal preconditioner is built directly (e3bdg.f).  eGMRES is COMMON /itrpar/'s interval
 * (common.h:217): in/out, recomputed by itrFDI (itrfdi.f) when st->iter==1 and mod(st->istep,20)==0.
 * y must have been through itrBC (itrdrv.f:394), as Au1MFG applies itrBC to the perturbed state. */
int phb200_solmfg(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                  double *BDiag, double *Dy, double *HBrg, int *iKs, int *lGMRES, int *ntotGM, double *eGMRES);
/* ElmMFG (elmmfg.f:1-256): res, modified residual rmes, e3bdg block diagonal (iprec/=0); any may be NULL */
int phb200_elmmfg(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                  double *rmes, double *BDiag);
/* ItrRes (itrres.f:1-171) of yp(nshg,5) against the state of the last phb200_elmmfg / phb200_solmfg call */
int phb200_itrres(phb200_ctx *ctx, const double *yp, double *rmes, int iabres);
/* Au1MFG (au1mfg.f:1-98) in place on u(nshg,5); setup!=0 first performs solmfg.f:97-135 (LU_Fact, forward
 * reduction of res / rmes, ypre) on the outputs of phb200_elmmfg */
int phb200_au1mfg(phb200_ctx *ctx, double *u, double eGMRES, int setup);
/* COMMON /itrpar/ eGMRES (common.h:217) as kept by the context: set!=0 stores *e, else reads it */
int phb200_egmres(phb200_ctx *ctx, double *e, int set);
/* HBM-resident variants (state from phb200_set_state) */
int phb200_dev_elmmfg(phb200_ctx *ctx, const phb200_step *st);
int phb200_dev_solve_mfg(phb200_ctx *ctx, const phb200_step *st, int *iKs, int *lGMRES, int *ntotGM);
int phb200_dev_au1mfg(phb200_ctx *ctx, int slot);

#ifdef __cplusplus
}
#endif
#endif

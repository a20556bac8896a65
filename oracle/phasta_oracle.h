/*
 * phasta_oracle.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * CPU restatement (plain C, double precision) of the PHASTA compressible
 * implicit-step hot path, used as the parity checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * PINNED: the reference ships no golden vectors for this path
 * and no Fortran compiler exists here, so the reference's own unmodified
 * Fortran sources are executed by the f77np interpreter (tests/golden/f77np.py,
 * make_golden_f77.py) and this oracle reproduces their outputs (res bit for
 * bit, EGmass/BDiag/Dy to round-off, colm/rowp exactly; tests/test_golden_f77.py);
 * the halo exchange against the executed ctypes.f + commu.f (tests/test_reference_commu.py), whole time steps
 * against itrdrv.f's sequence (tests/test_timestep.py), the incompressible flavour (tests/test_incomp.py).
 * Quadrature/shape tables are pinned against the reference's own C generators
 * (oracle/_ref, tests/golden/tables_ref.npz).
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * it restates.  Array layouts are the Fortran ones (column-major, 1-based
 * node ids inside ien / iper / ilwork).
 */
#ifndef PHASTA_ORACLE_H
#define PHASTA_ORACLE_H
#include <stdint.h>

#define ORC_MAXTOP 6   /* phSolver/common/common.h:17-20 */
#define ORC_MAXSH 32
#define ORC_MAXQPT 125

/* snapshot of the COMMON-block scalars the path reads (common.h:35-268) */
typedef struct orc_common {
  /* /conpar/ */
  int nshg, numnp, numel, numelb, nflow, ndof, ndofBC, nshape, nedof;
  /* /blkdat/, /fronts/, /workfc/ */
  int nelblk, nelblb, nlwork, numpe, myrank;
  /* /genpar/ */
  int ipord, idiff, itau, iprec, lhs, ires, iremoveStabTimeTerm, EntropyPressure;
  /* /solpar/ /incomp/ */
  int iDC, Navier, Kspace, nGMRES, minIters;
  /* material: matflg(2,1) viscosity model, matflg(3,1) bulk visc flag */
  int matflg2, matflg3;
  int pad0;
  /* /mmatpar/ /matdat/ /precis/ /outpar/ */
  double Rgas, gamma, gamma1, pr, datmat121, datmat221, datmat321, datmat131;
  double epsM, dtsfct, taucfct, temper;
  /* /timdat/ */
  double Dtgl, almi, alfi, gami, etol;
  /* /intpt/ */
  int nint[ORC_MAXTOP], nintb[ORC_MAXTOP];
  double Qwt[ORC_MAXTOP * ORC_MAXQPT];  /* Qwt(MAXTOP,MAXQPT)  */
  double Qwtb[ORC_MAXTOP * ORC_MAXQPT]; /* Qwtb(MAXTOP,MAXQPT) */
} orc_common;

/* one mesh part (= one MPI rank of the reference) */
typedef struct orc_part {
  orc_common c;
  const int *lcblk;       /* lcblk(10,nelblk+1)                                */
  const int *ien;         /* all mien(iblk)%p concatenated, each (npro,nshl)   */
  const int64_t *ien_off; /* offset of block iblk inside ien                   */
  const int *lcblkb;      /* lcblkb(10,nelblb+1)                               */
  const int *ienb;        /* mienb concatenated, each (npro,nshl)              */
  const int64_t *ienb_off;
  const int *iBCB;        /* miBCB concatenated, each (npro,2)                 */
  const int64_t *iBCB_off;
  const double *BCB;      /* mBCB concatenated, each (npro,nshlb,ndBCB)        */
  const int64_t *BCB_off;
  const double *x;        /* x(numnp,3)                                        */
  const int *iBC;         /* iBC(nshg)                                         */
  const double *BC;       /* BC(nshg,ndofBC)                                   */
  const int *iper;        /* iper(nshg) 1-based                                */
  const int *ilwork;      /* ilwork(nlwork), iother 0-based (after ctypes.f)   */
  const double *shp;      /* shp(MAXTOP,MAXSH,MAXQPT)                          */
  const double *shgl;     /* shgl(MAXTOP,3,MAXSH,MAXQPT)                       */
  const double *shpb;
  const double *shglb;
  /* state in, results out; all caller-allocated */
  const double *y;  /* y(nshg,ndof)  {u1,u2,u3,p,T}                            */
  const double *ac; /* ac(nshg,ndof)                                           */
  double *res;      /* res(nshg,nflow) {p,u1,u2,u3,T} order                    */
  double *rmes;     /* rmes(nshg,nflow)                                        */
  double *BDiag;    /* BDiag(nshg,nflow,nflow)                                 */
  double *EGmass;   /* EGmass(numel,nedof,nedof)                               */
  double *qres;     /* qres(nshg,idflx)                                        */
  double *rmass;    /* rmass(nshg)                                             */
  double *Dy;       /* Dy(nshg,nflow)                                          */
  double *uBrg;     /* uBrg(nshg,nflow,Kspace+1)                               */
  double *temp;     /* temp(nshg,nflow)                                        */
  double *lhsK;     /* lhsK(nflow*nflow,nnz_tot) (sparse path)                 */
  const int *colm;  /* colm(nshg+1)                                            */
  const int *rowp;  /* rowp(nnz_tot)                                           */
  double *aerfrc;   /* Force(3), HFlux, flxID(10,0:1000) (common.h:106); nullable  */
} orc_part;

#ifdef __cplusplus
extern "C" {
#endif

/* tables: genint.f:30-75 + genshp.f:34-37 for linear tets */
void orc_tet_tables(int rule, int *nint, double *Qwt, double *shp, double *shgl);
void orc_tri_tables(int rule, int *nintb, double *Qwtb, double *shpb, double *shglb);

/* ElmGMRe (elmgmr.f:1-274); all parts at once so commu can run in-process */
void orc_elmgmre(int nparts, orc_part *parts);
/* i3LU (i3lu.f:1-181): code 0 LU_Fact, 1 forward, 2 backward, 3 product */
void orc_i3lu(const orc_common *c, double *Diag, double *r, int code);
/* i3pre (i3pre.f:1-146) */
void orc_i3pre(int nparts, orc_part *parts);
/* Au1GMR (au1gmr.f:1-106): u <- A u in place, u(nshg,nflow) per part */
void orc_au1gmr(int nparts, orc_part *parts, double **u);
/* bc3per (bc3per.f:1-46) */
void orc_bc3per(const orc_part *p, double *r, int nQs);
/* commu (commu.f:1-297) code 0 'in', 1 'out' on global(nshg,n) per part */
void orc_commu(int nparts, orc_part *parts, double **global, int n, int code);
/* sumgat (mpitools.f:98-137) */
double orc_sumgat(int nparts, orc_part *parts, double **u, int n);
/* SolGMRe (solgmr.f:1-362).  HBrg(Kspace+1,Kspace), eBrg, yBrg, Rcos, Rsin
 * caller-allocated; returns iKs, lGMRES; ntotGM incremented. */
void orc_solgmre(int nparts, orc_part *parts, double *HBrg, double *eBrg,
                 double *yBrg, double *Rcos, double *Rsin, int *iKs,
                 int *lGMRES, int *ntotGM);
/* genadj/Asadj (genadj.f:1-82, asadj.f:1-59): returns nnz_tot, fills colm
 * (nshg+1) and rowp (capacity nnz*nshg, compacted) */
int orc_genadj(const orc_part *p, int nnz, int *colm, int *rowp);
/* ElmGMRs / SparseAp / Spsi3pre / SolGMRs (elmgmr.f:280-612, sparseap.f,
 * spsi3pre.f, solgmr.f:368-744) */
void orc_elmgmrs(int nparts, orc_part *parts);
void orc_sparseap(int nparts, orc_part *parts, double **u);
void orc_spsi3pre(int nparts, orc_part *parts);
void orc_solgmrs(int nparts, orc_part *parts, double *HBrg, double *eBrg,
                 double *yBrg, double *Rcos, double *Rsin, int *iKs,
                 int *lGMRES, int *ntotGM);
/* Newton / time-step shell (oracle_step.c): itrBC (itrbc.f:1-199), itrCorrect /
 * itrUpdate (itrPC.f:127-150,205-210), rstat norms (rstat.f:94-112) and one
 * whole step of itrdrv.f's flow sequence (predictor, nitr x (solve, correct,
 * itrBC), update). */
void orc_itrbc(int nparts, orc_part *parts, double **y, double **ac, int ires);
void orc_itrcorrect(const orc_part *p, double *y, double *ac, const double *yold,
                    const double *acold, const double *Dy);
void orc_itrupdate(const orc_part *p, double *yold, double *acold, const double *y,
                   const double *ac);
void orc_rstat(int nparts, orc_part *parts, int nshgt, double *totres);
void orc_timestep(int nparts, orc_part *parts, double **y, double **ac, double **yold,
                  double **acold, int ipred, int nitr, int sparse, int LHSupd, int nshgt,
                  int *ifuncs, int *ntotGM, double *stats);
int orc_sizeof_part(void);
int orc_sizeof_common(void);

#ifdef __cplusplus
}
#endif
#endif

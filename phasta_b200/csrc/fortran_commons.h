/*
 * fortran_commons.h -- C mirrors of the COMMON blocks the reference's solver entry points read
 * (phSolver/common/common.h:53-54,92-96,111,121-125,172-189,217-262; C precedent: common_c.h:86-665; /blkdat/,
 * /intpt/ and /shpdat/ have no mirror there).  gfortran names a COMMON block <name>_ .
 * PHB_COMMON_EXTERN is `extern` for the drop-in (the Fortran executable owns the storage) and empty for the test
 * stand-in that defines them (tests/fortran_abi/commons.c).
 */
#ifndef PHB_FORTRAN_COMMONS_H
#define PHB_FORTRAN_COMMONS_H
#ifndef PHB_COMMON_EXTERN
#define PHB_COMMON_EXTERN extern
#endif
#define MAXBLK 50000 /* common.h:17-24 */
#define MAXTS 100
#define MAXTOP 6
#define MAXQPT 125
#define MAXSH 32

PHB_COMMON_EXTERN struct { int master, numpe, myrank; } workfc_;
PHB_COMMON_EXTERN struct { int maxfront, nlwork; } fronts_;
PHB_COMMON_EXTERN struct { long long nshgt, minowned, maxowned; int numper, nshg0; } newdim_;
PHB_COMMON_EXTERN struct {
  int numnp, numel, numelb, numpbc, nen, nfaces, numflx, ndof, iALE, icoord, navier, irs, iexec, necho, ichem, iRK,
      nedof, nshg, nnz, istop, nflow, nnz_tot, idtn, ncorpsize, iownnodes, usingpetsc, numerr;
} conpar_;
PHB_COMMON_EXTERN struct { int lcblk[MAXBLK + 1][10], lcblkb[MAXBLK + 1][10]; } blkdat_;
PHB_COMMON_EXTERN struct { int nshape, nshapeb, maxshb, nshl, nshlb, nfath, ntopsh, nsonmax; } shpdat_;
PHB_COMMON_EXTERN struct {
  int lelCat, lcsyst, iorder, nenb, nelblk, nelblb, ndofl, nsymdl, nenl, nfacel, nenbl, intind, mattyp;
} elmpar_;
PHB_COMMON_EXTERN struct {
  double E3nsd;
  int I3nsd, nsymdf, ndofBC, ndiBCB, ndBCB, Jactyp, jump, ires, iprec, iprev, ibound, idiff, lhs, itau, ipord, ipred,
      lstres, iepstm;
  double dtsfct, taucfct;
  int ibksiz, iabc, isurf, idflx;
  double Bo;
  int EntropyPressure, irampViscOutlet, istretchOutlet, iremoveStabTimeTerm, iLHScond;
} genpar_;
PHB_COMMON_EXTERN struct {
  double Qpt[MAXQPT][4][MAXTOP], Qwt[MAXQPT][MAXTOP], Qptb[MAXQPT][4][MAXTOP], Qwtb[MAXQPT][MAXTOP];
  int nint[MAXTOP], nintb[MAXTOP], ngauss, ngaussb, intp, maxnint;
} intpt_;
PHB_COMMON_EXTERN struct { double eGMRES; int lGMRES, lGMRESs, iKs, iKss, ntotGM, ntotGMs; } itrpar_;
PHB_COMMON_EXTERN struct { double pr, Planck, Stefan, Nh, Rh, Rgas, gamma, gamma1, s0; } mmatpar_;
PHB_COMMON_EXTERN struct { double datmat[MAXTS][7][3]; int matflg[MAXTS][6]; int nummat, mexist; } matdat_;
PHB_COMMON_EXTERN struct { double ro, vel, temper, press, entrop; int ntout; } outpar_;
PHB_COMMON_EXTERN struct { double epsM; int iabres; } precis_;
PHB_COMMON_EXTERN struct { int imap, ivart, iDC, iPcond, Kspace, nGMRES, iconvflow, iconvsclr, idcsclr[2]; } solpar_;
PHB_COMMON_EXTERN struct {
  double time, CFLfld, CFLsld, Dtgl, Dtmax, alpha, etol;
  int lstep, ifunc, itseq, istep, iter, nitr;
  double almi, alfi, gami, flmpl, flmpr, dtol[2];
  int iCFLworst, lskeep;
} timdat_;
PHB_COMMON_EXTERN struct { int numeqns[100], minIters, maxIters; } incomp_;

#endif

// asm_host.cpp -- TEST INFRASTRUCTURE.  Compiles the device code of phasta_b200/csrc/assembly.cu for the host
// (PHB_HOST_EMUL) behind the SIMT shim and runs the hex / wedge assembly kernel k_asigmr_gen -- node records from
// k_pack_nodes, AsIGMR + e3 (+ e3dc) + BDiag + bc3LHS, EBE tiles or residual only -- on a fixture-sized mesh.
// Not a fallback: nothing in phasta_b200/ loads it.
#include "cuda_shim_simt.h"
#define PHB_HOST_EMUL 1
#define EG_TILE 32
#include "../../phasta_b200/csrc/assembly.cu"

// phys: Rgas gamma gamma1 pr mu0 Tref Ssuth dat131 dtsfct taucfct temper Dtgl fct1 epsM
// iphys: matflg2 matflg3 idiff iremove ipord lhs iprec iDC
extern "C" int asm_host_gen(int nshl, int lhs_mode, int numel, int nshg, int numnp, const int *ien /* [nshl][numel_pad] 0-based */,
                            const double *x, const double *y, const double *ac, const double *q, const int *iBC,
                            const double *BC, const int *nint, const double *Qwt, const double *shp, const double *shgl,
                            const double *phys, const int *iphys, double *res, double *BDiag, double *EG) {
  const int lcsyst = (nshl == 8) ? 2 : 3, tab = (nshl == 8) ? 0 : 1, top = lcsyst - 1, nq = nint[top];
  if ((nshl == 8 && nq != 8) || (nshl == 6 && nq != 6)) return -1;
  PhysParams p;
  memset(&p, 0, sizeof p);
  p.Rgas = phys[0]; p.gamma = phys[1]; p.gamma1 = phys[2]; p.pr = phys[3]; p.mu0 = phys[4]; p.Tref = phys[5];
  p.Ssuth = phys[6]; p.dat131 = phys[7]; p.dtsfct = phys[8]; p.taucfct = phys[9]; p.temper = phys[10];
  p.Dtgl = phys[11]; p.fct1 = phys[12]; p.epsM = phys[13];
  p.matflg2 = iphys[0]; p.matflg3 = iphys[1]; p.idiff = iphys[2]; p.iremove = iphys[3]; p.ipord = iphys[4];
  p.lhs = iphys[5]; p.iprec = iphys[6]; p.iDC = iphys[7];
  c_ph = p;
  GenTables gt;
  memset(&gt, 0, sizeof gt);
  gt.nq = nq; gt.nshl = nshl;
  for (int qq = 0; qq < nq; qq++) {
    gt.Qwt[qq] = Qwt[top + PHB200_MAXTOP * qq];
    for (int a = 0; a < nshl; a++) {
      gt.N[qq][a] = shp[top + PHB200_MAXTOP * (a + PHB200_MAXSH * qq)];
      for (int i = 0; i < 3; i++) gt.dN[qq][a][i] = shgl[top + PHB200_MAXTOP * (i + 3 * (a + PHB200_MAXSH * qq))];
    }
  }
  c_gen[tab] = gt;
  std::vector<double> aos((size_t)nshg * NREC);
  {
    const int tot = nshg * NREC, nb = (tot + 255) / 256;
    double *pa = aos.data();
    shim_launch(nb, 256, [=]() { k_pack_nodes(nshg, numnp, x, y, ac, q, p.idiff >= 1, pa); });
  }
  const size_t numel_pad = (size_t)((numel + 31) / 32) * 32;
  const int ntiles = (numel + 31) / 32;
  const bool dc = p.iDC != 0;
  double *pa = aos.data();
#define RUN(NSHL, NQ, LHS, DC)                                                                              \
  shim_launch(ntiles, 32 * NQ, [=]() {                                                                      \
    k_asigmr_gen<NSHL, NQ, LHS, DC>(tab, numel, numel_pad, nshg, ntiles, ien, pa, iBC, BC, res, BDiag, EG,  \
                                    nullptr, nullptr);                                                      \
  })
  if (nshl == 8) {
    if (lhs_mode == 1) { if (dc) RUN(8, 8, 1, true); else RUN(8, 8, 1, false); }
    else { if (dc) RUN(8, 8, 0, true); else RUN(8, 8, 0, false); }
  } else {
    if (lhs_mode == 1) { if (dc) RUN(6, 6, 1, true); else RUN(6, 6, 1, false); }
    else { if (dc) RUN(6, 6, 0, true); else RUN(6, 6, 0, false); }
  }
#undef RUN
  return 0;
}

// ---- shared set-up for the matrix-free entry points below ---------------------------------------------------
static void set_phys(const double *phys, const int *iphys) {
  PhysParams p;
  memset(&p, 0, sizeof p);
  p.Rgas = phys[0]; p.gamma = phys[1]; p.gamma1 = phys[2]; p.pr = phys[3]; p.mu0 = phys[4]; p.Tref = phys[5];
  p.Ssuth = phys[6]; p.dat131 = phys[7]; p.dtsfct = phys[8]; p.taucfct = phys[9]; p.temper = phys[10];
  p.Dtgl = phys[11]; p.fct1 = phys[12]; p.epsM = phys[13];
  p.matflg2 = iphys[0]; p.matflg3 = iphys[1]; p.idiff = iphys[2]; p.iremove = iphys[3]; p.ipord = iphys[4];
  p.lhs = iphys[5]; p.iprec = iphys[6]; p.iDC = iphys[7];
  c_ph = p;
}
static int set_tables(int nshl, const int *nint, const double *Qwt, const double *shp, const double *shgl) {
  if (nshl == 4) {
    TetTables t;
    memset(&t, 0, sizeof t);
    t.nq = nint[0];
    if (t.nq != 4) return -1;
    for (int q = 0; q < t.nq; q++) {
      t.Qwt[q] = Qwt[0 + PHB200_MAXTOP * q];
      for (int a = 0; a < 4; a++) {
        t.N[q][a] = shp[0 + PHB200_MAXTOP * (a + PHB200_MAXSH * q)];
        for (int i = 0; i < 3; i++) t.dN[q][a][i] = shgl[0 + PHB200_MAXTOP * (i + 3 * (a + PHB200_MAXSH * q))];
      }
    }
    c_tet = t;
    return 0;
  }
  const int top = (nshl == 8) ? 1 : 2, tab = (nshl == 8) ? 0 : 1, nq = nint[top];
  if (nq != nshl) return -1;
  GenTables gt;
  memset(&gt, 0, sizeof gt);
  gt.nq = nq; gt.nshl = nshl;
  for (int qq = 0; qq < nq; qq++) {
    gt.Qwt[qq] = Qwt[top + PHB200_MAXTOP * qq];
    for (int a = 0; a < nshl; a++) {
      gt.N[qq][a] = shp[top + PHB200_MAXTOP * (a + PHB200_MAXSH * qq)];
      for (int i = 0; i < 3; i++) gt.dN[qq][a][i] = shgl[top + PHB200_MAXTOP * (i + 3 * (a + PHB200_MAXSH * qq))];
    }
  }
  c_gen[tab] = gt;
  return 0;
}
static std::vector<double> pack(int nshg, int numnp, const double *x, const double *y, const double *ac,
                                const double *q) {
  std::vector<double> aos((size_t)nshg * NREC);
  double *pa = aos.data();
  const int with_q = c_ph.idiff >= 1;
  shim_launch((nshg * NREC + 255) / 256, 256, [=]() { k_pack_nodes(nshg, numnp, x, y, ac, q, with_q, pa); });
  return aos;
}

// ElmMFG's element pass (e3 with lhs=0, iprec=1: residual + e3bdg block diagonal), tets / hexes / wedges, with or
// without discontinuity capturing: k_asigmr_tet<32,4,3,DCON> / k_asigmr_gen<NSHL,NQ,3,DCON>
extern "C" int asm_host_bdg(int nshl, int numel, int nshg, int numnp, const int *ien, const double *x, const double *y,
                            const double *ac, const double *q, const int *iBC, const double *BC, const int *nint,
                            const double *Qwt, const double *shp, const double *shgl, const double *phys,
                            const int *iphys, double *res, double *BDiag) {
  set_phys(phys, iphys);
  if (set_tables(nshl, nint, Qwt, shp, shgl)) return -1;
  std::vector<double> aos = pack(nshg, numnp, x, y, ac, q);
  double *pa = aos.data();
  const size_t numel_pad = (size_t)((numel + 31) / 32) * 32;
  const int ntiles = (numel + 31) / 32, tab = (nshl == 8) ? 0 : 1;
  const bool dc = c_ph.iDC != 0;
#define RUNG(NSHL, NQ, DC)                                                                                    \
  shim_launch(ntiles, 32 * NQ, [=]() {                                                                        \
    k_asigmr_gen<NSHL, NQ, 3, DC>(tab, numel, numel_pad, nshg, ntiles, ien, pa, iBC, BC, res, BDiag, nullptr, \
                                  nullptr, nullptr);                                                          \
  })
#define RUNT(DC)                                                                                              \
  shim_launch(ntiles, 128, [=]() {                                                                            \
    k_asigmr_tet<32, 4, 3, DC>(numel, numel_pad, nshg, numnp, ntiles, ien, pa, iBC, BC, res, BDiag, nullptr,  \
                               nullptr, nullptr);                                                             \
  })
  if (nshl == 4) { if (dc) RUNT(true); else RUNT(false); }
  else if (nshl == 8) { if (dc) RUNG(8, 8, true); else RUNG(8, 8, false); }
  else { if (dc) RUNG(6, 6, true); else RUNG(6, 6, false); }
#undef RUNG
#undef RUNT
  return 0;
}

// AsIRes (the modified residual of the state yp around the base state y): k_asires<NSHL,NQ,DCM>, DCM from iDC and
// ires (2 ItrRes / Au1MFG, 3 ElmMFG) as phb_asires chooses it
extern "C" int asm_host_asires(int nshl, int ires, int iabres, int numel, int nshg, int numnp, const int *ien,
                               const double *x, const double *y, const double *ac, const double *q, const double *yp,
                               const int *nint, const double *Qwt, const double *shp, const double *shgl,
                               const double *phys, const int *iphys, double *rmes) {
  set_phys(phys, iphys);
  if (set_tables(nshl, nint, Qwt, shp, shgl)) return -1;
  std::vector<double> aos = pack(nshg, numnp, x, y, ac, q);
  double *pa = aos.data();
  const size_t numel_pad = (size_t)((numel + 31) / 32) * 32;
  const int tab = (nshl == 8) ? 0 : 1;
  const int dcm = (c_ph.iDC != 0) ? (ires == 3 ? 3 : 2) : 0;
  const int nb = (numel + 127) / 128;
#define RUNR(NSHL, NQ, DCM) \
  shim_launch(nb, 128, [=]() { k_asires<NSHL, NQ, DCM>(tab, numel, numel_pad, nshg, ien, pa, yp, rmes, iabres); })
#define RUNR3(NSHL, NQ) \
  do { if (dcm == 3) RUNR(NSHL, NQ, 3); else if (dcm == 2) RUNR(NSHL, NQ, 2); else RUNR(NSHL, NQ, 0); } while (0)
  if (nshl == 4) RUNR3(4, 4);
  else if (nshl == 8) RUNR3(8, 8);
  else RUNR3(6, 6);
#undef RUNR3
#undef RUNR
  return 0;
}


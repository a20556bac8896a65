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

// bnd_pack.h -- host-side preparation of the boundary-element groups that are not tets (hexes with a
// quadrilateral face, wedges with a triangular or quadrilateral face): kind of a lcblkb column, the device layout
// of ienb / iBCB / BCB, and the face tables.  Plain C++ (no CUDA calls) so that api.cu, assembly.cu and the host
// emulation of k_asbmfg_gen (tests/host_emul/bnd_host.cpp) share one copy.
// Reference: common/genbkbPosix.f:103-123 (lcblkb rows), compressible/elmgmr.f:180-222 (boundary block loop),
// common/genshpb.f, common/genint.f (face rules).
#pragma once
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/phb200.h"

// physical parameters of one call (constant memory on the device)
struct PhysParams {
  double Rgas, gamma, gamma1, pr, mu0, Tref, Ssuth, dat131;
  double dtsfct, taucfct, temper, Dtgl, fct1;  // fct1 = almi/gami/alfi*Dtgl
  int matflg2, matflg3, idiff, iremove, ipord, lhs, iprec, iDC;
  double epsM;
};

// volume shape functions at the face points of lcsyst 2 (hex), 3 (wedge / triangular face), 4 (wedge /
// quadrilateral face): N[q][a] = shpb(lcsyst,a,q), dN[q][a][i] = shglb(lcsyst,i,a,q), Qwt[q] = Qwtb(lcsyst,q)
struct BndTables {
  int nq;
  double N[4][8];
  double dN[4][8][3];
  double Qwt[4];
};

// kinds: 0 tets (k_asbmfg_tet), 1 hexes, 2 wedges / triangular face, 3 wedges / quadrilateral face; -1 unsupported
static const int PHB_BND_NSHL[4] = {4, 8, 6, 6}, PHB_BND_NSHLB[4] = {3, 4, 3, 4}, PHB_BND_LCSYST[4] = {1, 2, 3, 4};
static inline int phb_bnd_kind(const int *lc) {
  int lcsyst = lc[2];
  const int nenl = lc[4], nenbl = lc[5], nshl = lc[8], nshlb = lc[9];
  if (lcsyst == 3) lcsyst = nenbl;  // elmgmr.f:191
  if (lcsyst == 1 && nenl == 4 && nshl == 4 && nshlb == 3 && nenbl == 3) return 0;
  if (lcsyst == 2 && nenl == 8 && nshl == 8 && nshlb == 4 && nenbl == 4) return 1;
  if (lcsyst == 3 && nenl == 6 && nshl == 6 && nshlb == 3 && nenbl == 3) return 2;
  if (lcsyst == 4 && nenl == 6 && nshl == 6 && nshlb == 4 && nenbl == 4) return 3;
  return -1;
}

// Concatenate the blocks of one kind: ienb -> [nshl][nb] 0-based, iBCB -> [2][nb], BCB(e,n,j) -> [(j*nshlb+n)*nb+e].
// Returns the number of elements, -1 on a connectivity entry out of range.
static inline int phb_bnd_pack(int kind, int nelblb, const int *lcblkb, const int *const *mienb,
                               const int *const *miBCB, const double *const *mBCB, int nshg,
                               std::vector<int> &ienb, std::vector<int> &ib, std::vector<double> &bcb) {
  const int nshl = PHB_BND_NSHL[kind], nshlb = PHB_BND_NSHLB[kind];
  int nb = 0;
  for (int b = 0; b < nelblb; b++)
    if (phb_bnd_kind(lcblkb + 10 * b) == kind) nb += lcblkb[10 * b + 10] - lcblkb[10 * b];
  ienb.assign((size_t)nshl * nb, 0);
  ib.assign((size_t)2 * nb, 0);
  bcb.assign((size_t)6 * nshlb * nb, 0.0);
  size_t e0 = 0;
  for (int b = 0; b < nelblb; b++) {
    const int *lc = lcblkb + 10 * b;
    if (phb_bnd_kind(lc) != kind) continue;
    const int npro = lc[10] - lc[0];
    for (int e = 0; e < npro; e++) {
      for (int a = 0; a < nshl; a++) {
        int v = mienb[b][e + (size_t)npro * a];
        if (v < 0) v = -v;
        if (v < 1 || v > nshg) return -1;
        ienb[(size_t)a * nb + e0 + e] = v - 1;
      }
      ib[e0 + e] = miBCB[b][e];
      ib[(size_t)nb + e0 + e] = miBCB[b][e + (size_t)npro];
      for (int j = 0; j < 6; j++)
        for (int n = 0; n < nshlb; n++)
          bcb[(size_t)(j * nshlb + n) * nb + e0 + e] = mBCB[b][e + (size_t)npro * (n + nshlb * j)];
    }
    e0 += npro;
  }
  return nb;
}

static inline int phb_bnd_fill_tables(BndTables *b, int lcsyst, int nshl, const int *nintb, const double *Qwtb,
                                      const double *shpb, const double *shglb) {
  memset(b, 0, sizeof *b);
  const int top = lcsyst - 1;
  b->nq = nintb[top];
  if (b->nq < 1 || b->nq > 4) {
    fprintf(stderr, "phb200: init: boundary rule with %d points for lcsyst %d not supported\n", b->nq, lcsyst);
    return 1;
  }
  for (int q = 0; q < b->nq; q++) {
    b->Qwt[q] = Qwtb[top + PHB200_MAXTOP * q];
    for (int a = 0; a < nshl; a++) {
      b->N[q][a] = shpb[top + PHB200_MAXTOP * (a + PHB200_MAXSH * q)];
      for (int i = 0; i < 3; i++) b->dN[q][a][i] = shglb[top + PHB200_MAXTOP * (i + 3 * (a + PHB200_MAXSH * q))];
    }
  }
  return 0;
}

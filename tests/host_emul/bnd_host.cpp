// bnd_host.cpp -- TEST INFRASTRUCTURE.  Runs the product's boundary-flux kernel for hexes and wedges
// (phasta_b200/csrc/boundary.cuh: k_asbmfg_gen) on the host, one "thread" after the other, with the product's
// own group packing and table filling (bnd_pack.h), so that its arithmetic and data layout can be checked against
// the reference-Fortran fixtures where there is no GPU.  It is not a fallback: nothing in phasta_b200/ loads it.
#include "cuda_shim.h"
#include "../../phasta_b200/csrc/bnd_pack.h"
static PhysParams c_ph;
static BndTables c_bnd[3];
#include "../../phasta_b200/csrc/boundary.cuh"

template <int NSHL, int NSHLB, int LCS>
static void run(int nb, int nshg, int numnp, const int *ien, const int *ib, const double *bcb, const double *x,
                const double *y, double *res, double *aer, int do_force) {
  blockDim = {128, 1, 1};
  gridDim = {(unsigned)((nb + 127) / 128), 1, 1};
  for (unsigned b = 0; b < gridDim.x; b++)
    for (unsigned t = 0; t < 128; t++) {
      blockIdx = {b, 0, 0};
      threadIdx = {t, 0, 0};
      k_asbmfg_gen<NSHL, NSHLB, LCS>(nb, nshg, numnp, ien, ib, bcb, x, y, res, aer, do_force);
    }
}

// phys: Rgas, gamma, gamma1, pr, mu0, Tref, Ssuth, dat131; iphys: matflg2, matflg3
extern "C" int bnd_host_asbmfg(int nelblb, const int *lcblkb, const int *const *mienb, const int *const *miBCB,
                               const double *const *mBCB, int nshg, int numnp, const double *x, const double *y,
                               const int *nintb, const double *Qwtb, const double *shpb, const double *shglb,
                               const double *phys, const int *iphys, double *res, double *aer, int do_force) {
  memset(&c_ph, 0, sizeof c_ph);
  c_ph.Rgas = phys[0]; c_ph.gamma = phys[1]; c_ph.gamma1 = phys[2]; c_ph.pr = phys[3];
  c_ph.mu0 = phys[4]; c_ph.Tref = phys[5]; c_ph.Ssuth = phys[6]; c_ph.dat131 = phys[7];
  c_ph.matflg2 = iphys[0]; c_ph.matflg3 = iphys[1];
  int done = 0;
  for (int k = 1; k < 4; k++) {
    std::vector<int> ien, ib;
    std::vector<double> bcb;
    const int nb = phb_bnd_pack(k, nelblb, lcblkb, mienb, miBCB, mBCB, nshg, ien, ib, bcb);
    if (nb < 0) return -1;
    if (nb == 0) continue;
    const int lcs = PHB_BND_LCSYST[k];
    if (phb_bnd_fill_tables(&c_bnd[lcs - 2], lcs, PHB_BND_NSHL[k], nintb, Qwtb, shpb, shglb)) return -2;
    if (lcs == 2) run<8, 4, 2>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, res, aer, do_force);
    else if (lcs == 3) run<6, 3, 3>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, res, aer, do_force);
    else run<6, 4, 4>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, res, aer, do_force);
    done += nb;
  }
  return done;
}

// inc_bnd_host.cpp -- TEST INFRASTRUCTURE.  Runs the product's incompressible boundary-integral kernel
// (phasta_b200/csrc/inc_boundary.cuh: k_inc_asbmfg) on the host, one "thread" after the other, with the product's
// own group packing and table filling (bnd_pack.h).  Not a fallback: nothing in phasta_b200/ loads it.
#include "cuda_shim.h"
#include "../../phasta_b200/csrc/bnd_pack.h"
struct IncBndPhys {
  double rho, rmu;
  int iviscflux, iconvflow, itwmod;
};
static IncBndPhys c_ibp;
static BndTables c_ibnd[4];
#include "../../phasta_b200/csrc/inc_boundary.cuh"

template <int NSHL, int NSHLB, int LCS>
static void run(int nb, int nshg, int numnp, const int *ien, const int *ib, const double *bcb, const double *x,
                const double *y, const int *nsrflist, double *res, double *aer) {
  blockDim = {128, 1, 1};
  gridDim = {(unsigned)((nb + 127) / 128), 1, 1};
  for (unsigned b = 0; b < gridDim.x; b++)
    for (unsigned t = 0; t < 128; t++) {
      blockIdx = {b, 0, 0};
      threadIdx = {t, 0, 0};
      k_inc_asbmfg<NSHL, NSHLB, LCS>(nb, nshg, numnp, ien, ib, bcb, x, y, nsrflist, res, aer);
    }
}

extern "C" int inc_bnd_host_asbmfg(int nelblb, const int *lcblkb, const int *const *mienb, const int *const *miBCB,
                                   const double *const *mBCB, int nshg, int numnp, const double *x, const double *y,
                                   const int *nintb, const double *Qwtb, const double *shpb, const double *shglb,
                                   double rho, double rmu, int iviscflux, int iconvflow, int itwmod,
                                   const int *nsrflist, double *res, double *aer) {
  c_ibp.rho = rho; c_ibp.rmu = rmu;
  c_ibp.iviscflux = iviscflux; c_ibp.iconvflow = iconvflow; c_ibp.itwmod = itwmod;
  int done = 0;
  for (int k = 0; k < 4; k++) {
    std::vector<int> ien, ib;
    std::vector<double> bcb;
    const int nb = phb_bnd_pack(k, nelblb, lcblkb, mienb, miBCB, mBCB, nshg, ien, ib, bcb);
    if (nb < 0) return -1;
    if (nb == 0) continue;
    const int lcs = PHB_BND_LCSYST[k];
    if (phb_bnd_fill_tables(&c_ibnd[lcs - 1], lcs, PHB_BND_NSHL[k], nintb, Qwtb, shpb, shglb)) return -2;
    if (lcs == 1) run<4, 3, 1>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, nsrflist, res, aer);
    else if (lcs == 2) run<8, 4, 2>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, nsrflist, res, aer);
    else if (lcs == 3) run<6, 3, 3>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, nsrflist, res, aer);
    else run<6, 4, 4>(nb, nshg, numnp, ien.data(), ib.data(), bcb.data(), x, y, nsrflist, res, aer);
    done += nb;
  }
  return done;
}

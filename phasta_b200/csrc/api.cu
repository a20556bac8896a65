// api.cu -- the extern "C" boundary declared in include/phb200.h.
#include "ctx.h"
#include "bnd_pack.h"
#include <cstring>
#include <new>

static int fail(const char *routine, const char *what) {
  fprintf(stderr, "phb200: %s: %s\n", routine, what);
  return 1;
}

template <class T>
static int dev_alloc(T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  PHB_CHECK(cudaMalloc((void **)p, sizeof(T) * n));
  return 0;
}

// Concatenate the blocks of one topology: ien -> [nshl][numel_pad] 0-based, refel -> position in file order.
static int upload_group(phb200_ctx *ctx, const int *lcblk, const int *const *mien, int lcsyst, int nshl, int *numel_out,
                        size_t *pad_out, int **d_ien, int **d_refel) {
  const phb200_common &c = ctx->c;
  int numel = 0;
  for (int b = 0; b < c.nelblk; b++) {
    const int *lc = lcblk + 10 * b;
    if (lc[2] == lcsyst && lc[9] == nshl) numel += lc[10] - lc[0];
  }
  size_t pad = ((size_t)numel + 63) / 64 * 64;
  if (pad == 0) pad = 64;
  std::vector<int> ien((size_t)nshl * pad, 0), refel((size_t)(numel ? numel : 1), 0);
  size_t e0 = 0;
  for (int b = 0; b < c.nelblk; b++) {
    const int *lc = lcblk + 10 * b;
    if (lc[2] != lcsyst || lc[9] != nshl) continue;
    const int npro = lc[10] - lc[0];
    const int *ib = mien[b];
    for (int a = 0; a < nshl; a++)
      for (int e = 0; e < npro; e++) {
        int v = ib[e + (size_t)npro * a];
        if (v < 0) v = -v;
        if (v < 1 || v > c.nshg) return fail("init", "ien entry out of range");
        ien[(size_t)a * pad + e0 + e] = v - 1;
      }
    for (int e = 0; e < npro; e++) refel[e0 + e] = lc[0] - 1 + e;
    e0 += npro;
  }
  PHB_TRY(dev_alloc(d_ien, ien.size()));
  PHB_CHECK(cudaMemcpy(*d_ien, ien.data(), sizeof(int) * ien.size(), cudaMemcpyHostToDevice));
  PHB_TRY(dev_alloc(d_refel, refel.size()));
  PHB_CHECK(cudaMemcpy(*d_refel, refel.data(), sizeof(int) * refel.size(), cudaMemcpyHostToDevice));
  *numel_out = numel;
  *pad_out = pad;
  return 0;
}

extern "C" const char *phb200_version(void) { return "phb200 0.1 (sm_100a)"; }
extern "C" int phb200_sizeof_common(void) { return (int)sizeof(phb200_common); }
extern "C" int phb200_sizeof_incomp(void) { return (int)sizeof(phb200_incomp); }
extern "C" int phb200_sizeof_step(void) { return (int)sizeof(phb200_step); }

extern "C" int phb200_init(phb200_ctx **out, const phb200_common *c, const int *lcblk, const int *const *mien,
                           const int *lcblkb, const int *const *mienb, const int *const *miBCB,
                           const double *const *mBCB, const double *x, const int *iBC, const double *BC,
                           const int *iper, const int *ilwork, const double *shp, const double *shgl,
                           const double *shpb, const double *shglb, int device) {
  if (!out || !c) return fail("init", "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("init", "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("init", "bad device ordinal");
  if (c->nflow != 5 || c->ndof != 5) return fail("init", "only nflow=ndof=5 (no scalars) supported");
  if (c->ipord != 1) return fail("init", "only ipord=1 supported");
  if (c->itau != 0) return fail("init", "only itau=0 (Shakib diagonal tau) supported");
  if (c->iDC < 0 || c->iDC > 3) return fail("init", "iDC must be 0..3 (e3dc.f)");
  if (c->Navier != 1) return fail("init", "Navier must be 1");
  if (c->EntropyPressure != 0) return fail("init", "EntropyPressure=1 not supported");
  PHB_CHECK(cudaSetDevice(device));
  phb200_ctx *ctx = new (std::nothrow) phb200_ctx();
  if (!ctx) return fail("init", "out of host memory");
  ctx->c = *c;
  ctx->device = device;
  ctx->nccl = nullptr;
  ctx->local_group = false;
  ctx->launches = 0;
  ctx->profiling = false;
  ctx->have_lhs = false;
  memset(ctx->kc_ms, 0, sizeof ctx->kc_ms);
  memset(ctx->kc_n, 0, sizeof ctx->kc_n);
  PHB_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  PHB_CHECK(cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking));
  PHB_CHECK(cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming));
  PHB_CHECK(cudaEventCreateWithFlags(&ctx->ev_ac, cudaEventDisableTiming));
  ctx->ac_pending = false;
  for (int i = 0; i < 16; i++) PHB_CHECK(cudaEventCreate(&ctx->ev[i]));
  PHB_CHECK(cudaEventCreate(&ctx->pev0));
  PHB_CHECK(cudaEventCreate(&ctx->pev1));
  const int nshg = c->nshg, numnp = c->numnp;
  // ---- connectivity: blocks grouped by topology (genblkPosix.f:52-96), file order kept inside a group
  {
    int total = 0;
    for (int b = 0; b < c->nelblk; b++) total += lcblk[10 * b + 10] - lcblk[10 * b];
    if (total != c->numel) return fail("init", "lcblk does not add up to numel");
    int nt = 0;
    PHB_TRY(upload_group(ctx, lcblk, mien, 1, 4, &nt, &ctx->numel_pad, &ctx->d_ien, &ctx->d_refel_tet));
    ctx->numel_tet = nt;
    int covered = nt;
    const int topo[2][2] = {{2, 8}, {3, 6}};  // hexes, wedges: (lcsyst, nshl)
    for (int t = 0; t < 2; t++) {
      ElemGroup g;
      memset(&g, 0, sizeof g);
      g.lcsyst = topo[t][0];
      g.nshl = topo[t][1];
      g.nq = c->nint[g.lcsyst - 1];
      g.tab = t;
      PHB_TRY(upload_group(ctx, lcblk, mien, g.lcsyst, g.nshl, &g.numel, &g.numel_pad, &g.d_ien, &g.d_refel));
      if (g.numel == 0) {
        cudaFree(g.d_ien);
        cudaFree(g.d_refel);
        continue;
      }
      if (g.nq != g.nshl) return fail("init", "hex/wedge blocks need quadrature rule 2 (8-pt hexes, 6-pt wedges)");
      covered += g.numel;
      ctx->gen.push_back(g);
    }
    if (covered != c->numel)
      return fail("init", "only linear tet (lcsyst=1,nshl=4), hex (2,8) and wedge (3,6) blocks are supported");
    if (c->nedof < 5 * 4 || (!ctx->gen.empty() && c->nedof < 5 * 6))
      return fail("init", "nedof smaller than 5*nshl of a block");
  }
  // ---- nodal data
  PHB_TRY(dev_alloc(&ctx->d_x, (size_t)3 * numnp));
  PHB_CHECK(cudaMemcpy(ctx->d_x, x, sizeof(double) * 3 * (size_t)numnp, cudaMemcpyHostToDevice));
  PHB_TRY(dev_alloc(&ctx->d_iBC, (size_t)nshg));
  PHB_CHECK(cudaMemcpy(ctx->d_iBC, iBC, sizeof(int) * (size_t)nshg, cudaMemcpyHostToDevice));
  PHB_TRY(dev_alloc(&ctx->d_BC, (size_t)c->ndofBC * nshg));
  PHB_CHECK(cudaMemcpy(ctx->d_BC, BC, sizeof(double) * (size_t)c->ndofBC * nshg, cudaMemcpyHostToDevice));
  {
    std::vector<int> ip(nshg), sl;
    for (int i = 0; i < nshg; i++) {
      ip[i] = iper[i] - 1;
      if (ip[i] < 0 || ip[i] >= nshg) return fail("init", "iper entry out of range");
      if (iBC[i] & (1 << 10)) sl.push_back(i);
      if (iBC[i] & (1 << 11)) return fail("init", "SPEBC (iBC bit 11) not supported");
    }
    for (int j : sl)
      if (iBC[ip[j]] & (1 << 10)) return fail("init", "chained periodic masters not supported");
    PHB_TRY(dev_alloc(&ctx->d_iper, (size_t)nshg));
    PHB_CHECK(cudaMemcpy(ctx->d_iper, ip.data(), sizeof(int) * (size_t)nshg, cudaMemcpyHostToDevice));
    ctx->n_perslave = (int)sl.size();
    PHB_TRY(dev_alloc(&ctx->d_perslave, sl.size()));
    if (!sl.empty())
      PHB_CHECK(cudaMemcpy(ctx->d_perslave, sl.data(), sizeof(int) * sl.size(), cudaMemcpyHostToDevice));
  }
  // ---- boundary elements (genbkbPosix.f:113-123 lcblkb rows; asbmfg.f)
  ctx->numelb = 0;
  ctx->d_ienb = nullptr; ctx->d_iBCB = nullptr; ctx->d_BCB = nullptr;
  ctx->have_inc_btabs = false; ctx->d_nsrflist = nullptr; ctx->bnd_deformable = false;
  if (c->nelblb > 0) {
    if (!lcblkb || !mienb || !miBCB || !mBCB || !shpb || !shglb) return fail("init", "null boundary arrays");
    for (int b = 0; b < c->nelblb; b++) {
      if (phb_bnd_kind(lcblkb + 10 * b) < 0)
        return fail("init", "boundary block is not a linear tet, hex or wedge (lcsyst 1..4) with its face on lnode");
      const int npro = lcblkb[10 * b + 10] - lcblkb[10 * b];
      for (int e = 0; e < npro; e++)
        if (miBCB[b][e] & 16) ctx->bnd_deformable = true;   // e3b.f (incompressible): vessel-wall elements
    }
    for (int k = 0; k < 4; k++) {
      std::vector<int> ienb, ib;
      std::vector<double> bcb;
      const int nb = phb_bnd_pack(k, c->nelblb, lcblkb, mienb, miBCB, mBCB, nshg, ienb, ib, bcb);
      if (nb < 0) return fail("init", "ienb entry out of range");
      if (nb == 0) continue;
      int *d_i = nullptr, *d_b = nullptr;
      double *d_v = nullptr;
      PHB_TRY(dev_alloc(&d_i, ienb.size()));
      PHB_CHECK(cudaMemcpy(d_i, ienb.data(), sizeof(int) * ienb.size(), cudaMemcpyHostToDevice));
      PHB_TRY(dev_alloc(&d_b, ib.size()));
      PHB_CHECK(cudaMemcpy(d_b, ib.data(), sizeof(int) * ib.size(), cudaMemcpyHostToDevice));
      PHB_TRY(dev_alloc(&d_v, bcb.size()));
      PHB_CHECK(cudaMemcpy(d_v, bcb.data(), sizeof(double) * bcb.size(), cudaMemcpyHostToDevice));
      if (k == 0) {
        ctx->numelb = nb;
        ctx->d_ienb = d_i; ctx->d_iBCB = d_b; ctx->d_BCB = d_v;
      } else {
        BndGroup g;
        g.lcsyst = PHB_BND_LCSYST[k]; g.nshl = PHB_BND_NSHL[k]; g.nshlb = PHB_BND_NSHLB[k]; g.n = nb;
        g.d_ien = d_i; g.d_iBCB = d_b; g.d_BCB = d_v;
        ctx->bgen.push_back(g);
      }
    }
  }
  PHB_TRY(dev_alloc(&ctx->d_aerfrc, (size_t)4 + 10 * 1001));
  PHB_CHECK(cudaMemset(ctx->d_aerfrc, 0, sizeof(double) * (4 + 10 * 1001)));
  PHB_TRY(phb_halo_setup(ctx, ilwork));
  PHB_TRY(phb_upload_tables(ctx, shp, shgl, shpb, shglb));
  // ---- state and work arrays
  const size_t n5 = (size_t)5 * nshg;
  PHB_TRY(dev_alloc(&ctx->d_y, n5));
  PHB_TRY(dev_alloc(&ctx->d_ac, n5));
  PHB_TRY(dev_alloc(&ctx->d_qres, (size_t)12 * nshg));
  PHB_TRY(dev_alloc(&ctx->d_rmass, (size_t)nshg));
  PHB_TRY(dev_alloc(&ctx->d_nodeaos, (size_t)26 * nshg));
  PHB_TRY(dev_alloc(&ctx->d_res, n5));
  PHB_TRY(dev_alloc(&ctx->d_rmes, n5));
  PHB_TRY(dev_alloc(&ctx->d_Dy, n5));
  PHB_TRY(dev_alloc(&ctx->d_temp, n5));
  PHB_TRY(dev_alloc(&ctx->d_BDiag, (size_t)25 * nshg));
  ctx->d_BDtmp = nullptr;
  if (c->numpe > 1) PHB_TRY(dev_alloc(&ctx->d_BDtmp, (size_t)25 * nshg));
  ctx->d_EG = nullptr;  // allocated by the first EBE lhs=1 assembly
  ctx->d_yold = ctx->d_acold = nullptr;
  ctx->d_mfg = nullptr;
  ctx->d_mail = nullptr; ctx->d_peer_mail = nullptr; ctx->d_ticket = nullptr; ctx->d_p2p_err = nullptr;
  ctx->p2p = false; ctx->p2p_seq = 0;
  for (int r = 0; r < 64; r++) ctx->peer_mapped[r] = nullptr;
  ctx->d_res4 = ctx->d_lhsK9 = ctx->d_lhsP4 = ctx->d_lesp = ctx->d_lesq = ctx->d_lesp4 = nullptr;
  ctx->d_tpos = nullptr;
  ctx->have_inc_tabs = false;
  if (shpb && shglb) {
    ctx->h_shpb.assign(shpb, shpb + (size_t)PHB200_MAXTOP * PHB200_MAXSH * PHB200_MAXQPT);
    ctx->h_shglb.assign(shglb, shglb + (size_t)PHB200_MAXTOP * 3 * PHB200_MAXSH * PHB200_MAXQPT);
  }
  ctx->h_shp.assign(shp, shp + (size_t)PHB200_MAXTOP * PHB200_MAXSH * PHB200_MAXQPT);
  ctx->h_shgl.assign(shgl, shgl + (size_t)PHB200_MAXTOP * 3 * PHB200_MAXSH * PHB200_MAXQPT);
  ctx->eGMRES = 0.0;
  ctx->ifuncs = 0;
  ctx->nnz_tot = 0;
  ctx->d_colm = ctx->d_rowp = ctx->d_rowofblk = ctx->d_eloc = nullptr;
  ctx->d_lhsK = nullptr;
  ctx->have_lhs_sparse = false;
  // keep the host block structure for genadj
  ctx->h_lcblk.assign(lcblk, lcblk + 10 * (c->nelblk + 1));
  ctx->h_mien.assign(mien, mien + c->nelblk);
  PHB_TRY(dev_alloc(&ctx->d_uBrg, n5 * (size_t)(c->Kspace + 1)));
  PHB_TRY(dev_alloc(&ctx->d_dots, (size_t)c->Kspace + 8));
  PHB_CHECK(cudaMallocHost(&ctx->h_dots, sizeof(double) * ((size_t)c->Kspace + 8)));
  PHB_TRY(dev_alloc(&ctx->d_ptmp, n5));
  PHB_TRY(dev_alloc(&ctx->d_p5, n5));
  {
    const size_t nk = phb_kry_doubles(c->Kspace), nf = (size_t)c->Kspace + 4;
    PHB_TRY(dev_alloc(&ctx->d_kry, nk));
    PHB_CHECK(cudaMallocHost(&ctx->h_kry, sizeof(double) * nk));
    PHB_TRY(dev_alloc(&ctx->d_kflag, nf));
    PHB_CHECK(cudaMallocHost(&ctx->h_kflag, sizeof(int) * nf));
    for (int i = 0; i < 4; i++) PHB_CHECK(cudaEventCreateWithFlags(&ctx->kev[i], cudaEventDisableTiming));
  }
  ctx->scratch_bytes = (size_t)256 << 20;
  PHB_TRY(dev_alloc((char **)&ctx->d_scratch, ctx->scratch_bytes));
  PHB_CHECK(cudaMemset(ctx->d_qres, 0, sizeof(double) * 12 * (size_t)nshg));
  const int K = c->Kspace;
  ctx->HBrg.assign((size_t)(K + 1) * K, 0.0);
  ctx->eBrg.assign(K + 1, 0.0);
  ctx->yBrg.assign(K + 1, 0.0);
  ctx->Rcos.assign(K + 1, 0.0);
  ctx->Rsin.assign(K + 1, 0.0);
  PHB_CHECK(cudaDeviceSynchronize());
  *out = ctx;
  return 0;
}

extern "C" void phb200_finalize(phb200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  phb_comm_free(ctx);
  phb_inc_free(ctx);
  void *ptrs[] = {ctx->d_ien, ctx->d_iBC, ctx->d_BC, ctx->d_iper, ctx->d_x, ctx->d_perslave, ctx->d_halo_nodes,
                  ctx->d_slave_nodes, ctx->d_sendbuf, ctx->d_recvbuf, ctx->d_y, ctx->d_ac, ctx->d_qres,
                  ctx->d_rmass, ctx->d_res, ctx->d_rmes, ctx->d_Dy, ctx->d_temp, ctx->d_BDiag, ctx->d_BDtmp,
                  ctx->d_EG, ctx->d_uBrg, ctx->d_dots, ctx->d_scratch, ctx->d_ienb, ctx->d_iBCB,
                  ctx->d_BCB, ctx->d_aerfrc, ctx->d_colm, ctx->d_rowp, ctx->d_rowofblk, ctx->d_eloc, ctx->d_lhsK,
                  ctx->d_nodeaos, ctx->d_yold, ctx->d_acold, ctx->d_mfg, ctx->d_ptmp, ctx->d_p5, ctx->d_kry, ctx->d_kflag, ctx->d_apchunk, ctx->d_elc, ctx->d_inc, ctx->d_inc_ptr};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (ctx->d_refel_tet) cudaFree(ctx->d_refel_tet);
  for (ElemGroup &g : ctx->gen) {
    void *gp[] = {g.d_ien, g.d_refel, g.d_EG, g.d_eloc};
    for (void *p : gp)
      if (p) cudaFree(p);
  }
  for (BndGroup &g : ctx->bgen) {
    void *gp[] = {g.d_ien, g.d_iBCB, g.d_BCB};
    for (void *p : gp)
      if (p) cudaFree(p);
  }
  if (ctx->h_dots) cudaFreeHost(ctx->h_dots);
  if (ctx->h_kry) cudaFreeHost(ctx->h_kry);
  if (ctx->h_kflag) cudaFreeHost(ctx->h_kflag);
  for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->kev[i]);
  for (int i = 0; i < 16; i++) cudaEventDestroy(ctx->ev[i]);
  cudaEventDestroy(ctx->pev0);
  cudaEventDestroy(ctx->pev1);
  cudaEventDestroy(ctx->ev_main);
  cudaEventDestroy(ctx->ev_ac);
  cudaStreamDestroy(ctx->cstream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int phb200_nccl_unique_id(void *id128) { return phb_comm_unique_id(id128); }
extern "C" int phb200_comm_init(phb200_ctx *ctx, const void *id128) { return phb_comm_init(ctx, id128); }

#define ENTER(ctx)                                  \
  if (!(ctx)) return fail(__func__, "null context"); \
  PHB_CHECK(cudaSetDevice((ctx)->device));

static int h2d(phb200_ctx *ctx, double *d, const double *h, size_t n) {
  PHB_CHECK(cudaMemcpyAsync(d, h, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
static int d2h(phb200_ctx *ctx, double *h, const double *d, size_t n) {
  PHB_CHECK(cudaMemcpyAsync(h, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

// device tiles of one topology group -> EGmass(numel,nedof,nedof); rows/columns beyond 5*nshl stay zero
// (on mixed meshes nedof = 5*max nshl)
__global__ void k_eg_to_ref(int ngrp, int nd, int numel, int nedof, const int *__restrict__ refel,
                            const double *__restrict__ EG, double *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t tot = (size_t)ngrp * nd * nd;
  if (t >= tot) return;
  size_t e = t % ngrp;
  int k = (int)(t / ngrp), r = k % nd, c = k / nd;
  out[refel[e] + (size_t)numel * (r + (size_t)nedof * c)] =
      EG[((e / EG_TILE) * (size_t)(nd * nd) + (size_t)(r + nd * c)) * EG_TILE + (e % EG_TILE)];
}

extern "C" int phb200_set_state(phb200_ctx *ctx, const double *y, const double *ac) {
  ENTER(ctx);
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_TRY(h2d(ctx, ctx->d_y, y, n5));
  PHB_TRY(h2d(ctx, ctx->d_ac, ac, n5));
  return 0;
}
// y on the compute stream, ac on the copy stream (after everything already queued on the compute stream, which
// may still read the old d_ac); phb_elmgmre orders its first reader of d_ac behind ev_ac
static int set_state_split(phb200_ctx *ctx, const double *y, const double *ac) {
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_TRY(h2d(ctx, ctx->d_y, y, n5));
  // Y,t follows Y on the link (so Y is not slowed down) but on the copy stream, i.e. concurrently with AsIq
  PHB_CHECK(cudaEventRecord(ctx->ev_main, ctx->stream));
  PHB_CHECK(cudaStreamWaitEvent(ctx->cstream, ctx->ev_main, 0));
  PHB_CHECK(cudaMemcpyAsync(ctx->d_ac, ac, sizeof(double) * n5, cudaMemcpyHostToDevice, ctx->cstream));
  PHB_CHECK(cudaEventRecord(ctx->ev_ac, ctx->cstream));
  ctx->ac_pending = true;
  return 0;
}
extern "C" int phb200_dev_elmgmre(phb200_ctx *ctx, const phb200_step *st) {
  ENTER(ctx);
  return phb_elmgmre(ctx, st);
}
extern "C" int phb200_dev_solve(phb200_ctx *ctx, const phb200_step *st, int *iKs, int *lGMRES, int *ntotGM) {
  ENTER(ctx);
  return phb_solve(ctx, st, 0, iKs, lGMRES, ntotGM);
}
extern "C" int phb200_dev_ap(phb200_ctx *ctx, int slot) {
  ENTER(ctx);
  if (slot < 0 || slot >= ctx->c.Kspace) return fail("dev_ap", "slot out of range");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *src = ctx->d_uBrg + (size_t)slot * n5, *dst = src + n5;
  // as in the Krylov loop: a scratch copy gets the halo / periodic fill, the product lands in the next slot
  PHB_CHECK(cudaMemcpyAsync(ctx->d_ptmp, src, sizeof(double) * n5, cudaMemcpyDeviceToDevice, ctx->stream));
  PHB_TRY(phb_au1gmr2(ctx, ctx->d_ptmp, dst, nullptr));
  return phb_bc3per(ctx, dst, 5);
}
extern "C" int phb200_get_res(phb200_ctx *ctx, double *res) {
  ENTER(ctx);
  PHB_TRY(d2h(ctx, res, ctx->d_res, (size_t)5 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_get_dy(phb200_ctx *ctx, double *Dy) {
  ENTER(ctx);
  PHB_TRY(d2h(ctx, Dy, ctx->d_Dy, (size_t)5 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_get_bdiag(phb200_ctx *ctx, double *BD) {
  ENTER(ctx);
  PHB_TRY(d2h(ctx, BD, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_get_egmass(phb200_ctx *ctx, double *EGmass) {
  ENTER(ctx);
  const int numel = ctx->c.numel, nedof = ctx->c.nedof;
  if (numel == 0) return 0;
  if (!ctx->have_lhs) return fail("get_egmass", "no EBE LHS has been assembled");
  double *d_out = nullptr;
  size_t tot = (size_t)numel * nedof * nedof;
  PHB_CHECK(cudaMalloc(&d_out, sizeof(double) * tot));
  PHB_CHECK(cudaMemsetAsync(d_out, 0, sizeof(double) * tot, ctx->stream));
  if (ctx->numel_tet > 0) {
    size_t nthr = (size_t)ctx->numel_tet * 400;
    k_eg_to_ref<<<(unsigned)((nthr + 255) / 256), 256, 0, ctx->stream>>>(ctx->numel_tet, 20, numel, nedof,
                                                                         ctx->d_refel_tet, ctx->d_EG, d_out);
    ctx->launches++;
  }
  for (const ElemGroup &g : ctx->gen) {
    const int nd = 5 * g.nshl;
    size_t nthr = (size_t)g.numel * nd * nd;
    k_eg_to_ref<<<(unsigned)((nthr + 255) / 256), 256, 0, ctx->stream>>>(g.numel, nd, numel, nedof, g.d_refel, g.d_EG,
                                                                         d_out);
    ctx->launches++;
  }
  PHB_CHECK(cudaGetLastError());
  PHB_TRY(d2h(ctx, EGmass, d_out, tot));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_out);
  return 0;
}

// EGmass(e0+1 : e0+n, :, :) of the reference's element order, as out(n,nedof,nedof): spot checks on meshes whose
// whole EGmass does not fit on the host (bench.py's parity leg, tests/test_gpu_at_size.py)
__global__ void k_eg_to_ref_range(int ngrp, int nd, int e0, int n, int nedof, const int *__restrict__ refel,
                                  const double *__restrict__ EG, double *__restrict__ out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)ngrp * nd) return;
  const size_t e = t % ngrp;
  const int c = (int)(t / ngrp), pos = refel[e] - e0;
  if (pos < 0 || pos >= n) return;
  for (int r = 0; r < nd; r++)
    out[pos + (size_t)n * (r + (size_t)nedof * c)] =
        EG[((e / EG_TILE) * (size_t)(nd * nd) + (size_t)(r + nd * c)) * EG_TILE + (e % EG_TILE)];
}
extern "C" int phb200_get_egmass_range(phb200_ctx *ctx, long long e0, int n, double *EGmass) {
  ENTER(ctx);
  const int numel = ctx->c.numel, nedof = ctx->c.nedof;
  if (!EGmass || e0 < 0 || n < 0 || e0 + n > numel) return fail("get_egmass_range", "range outside 0..numel");
  if (n == 0) return 0;
  if (!ctx->have_lhs) return fail("get_egmass_range", "no EBE LHS has been assembled");
  double *d_out = nullptr;
  const size_t tot = (size_t)n * nedof * nedof;
  PHB_CHECK(cudaMalloc(&d_out, sizeof(double) * tot));
  PHB_CHECK(cudaMemsetAsync(d_out, 0, sizeof(double) * tot, ctx->stream));
  if (ctx->numel_tet > 0) {
    const size_t nthr = (size_t)ctx->numel_tet * 20;
    k_eg_to_ref_range<<<(unsigned)((nthr + 255) / 256), 256, 0, ctx->stream>>>(ctx->numel_tet, 20, (int)e0, n, nedof,
                                                                               ctx->d_refel_tet, ctx->d_EG, d_out);
    ctx->launches++;
  }
  for (const ElemGroup &g : ctx->gen) {
    const int nd = 5 * g.nshl;
    const size_t nthr = (size_t)g.numel * nd;
    k_eg_to_ref_range<<<(unsigned)((nthr + 255) / 256), 256, 0, ctx->stream>>>(g.numel, nd, (int)e0, n, nedof, g.d_refel,
                                                                               g.d_EG, d_out);
    ctx->launches++;
  }
  PHB_CHECK(cudaGetLastError());
  PHB_TRY(d2h(ctx, EGmass, d_out, tot));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_out);
  return 0;
}

extern "C" int phb200_elmgmre(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                              double *BDiag, double *EGmass, double *qres) {
  ENTER(ctx);
  if (!y || !ac || !st) return fail("elmgmre", "null argument");
  PHB_TRY(set_state_split(ctx, y, ac));
  PHB_TRY(phb_elmgmre(ctx, st));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res, (size_t)5 * ctx->c.nshg));
  if (BDiag && st->iprec) PHB_TRY(d2h(ctx, BDiag, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  if (qres) PHB_TRY(d2h(ctx, qres, ctx->d_qres, (size_t)12 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  if (EGmass && st->lhs == 1) PHB_TRY(phb200_get_egmass(ctx, EGmass));
  return 0;
}

extern "C" int phb200_solgmre(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                              double *rmes, double *BDiag, double *Dy, double *HBrg, double *eBrg, double *yBrg,
                              double *Rcos, double *Rsin, int *iKs, int *lGMRES, int *ntotGM) {
  ENTER(ctx);
  if (!y || !ac || !st || !Dy || !iKs || !lGMRES || !ntotGM) return fail("solgmre", "null argument");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_TRY(set_state_split(ctx, y, ac));
  PHB_TRY(phb_elmgmre(ctx, st));
  PHB_TRY(phb_solve(ctx, st, 0, iKs, lGMRES, ntotGM));
  PHB_TRY(d2h(ctx, Dy, ctx->d_Dy, n5));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res, n5));
  if (rmes) PHB_TRY(d2h(ctx, rmes, ctx->d_rmes, n5));
  if (BDiag) PHB_TRY(d2h(ctx, BDiag, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  const int K = ctx->c.Kspace;
  if (HBrg) memcpy(HBrg, ctx->HBrg.data(), sizeof(double) * (size_t)(K + 1) * K);
  if (eBrg) memcpy(eBrg, ctx->eBrg.data(), sizeof(double) * (K + 1));
  if (yBrg) memcpy(yBrg, ctx->yBrg.data(), sizeof(double) * (K + 1));
  if (Rcos) memcpy(Rcos, ctx->Rcos.data(), sizeof(double) * (K + 1));
  if (Rsin) memcpy(Rsin, ctx->Rsin.data(), sizeof(double) * (K + 1));
  return 0;
}


// ---- block-CSR flavour -------------------------------------------------------
extern "C" int phb200_genadj(phb200_ctx *ctx, int nnz, int *colm, int *rowp, int *nnz_tot) {
  ENTER(ctx);
  if (!nnz_tot) return fail("genadj", "null argument");
  return phb_genadj(ctx, nnz, colm, rowp, nnz_tot);
}
extern "C" int phb200_set_sparse(phb200_ctx *ctx, const int *colm, const int *rowp, int nnz_tot) {
  ENTER(ctx);
  if (!colm || !rowp) return fail("set_sparse", "null argument");
  return phb_set_sparse(ctx, colm, rowp, nnz_tot);
}
static int get_lhsk(phb200_ctx *ctx, double *lhsK) {
  PHB_CHECK(cudaMemcpyAsync(lhsK, ctx->d_lhsK, sizeof(double) * 25 * (size_t)ctx->nnz_tot, cudaMemcpyDeviceToHost,
                            ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// lhsK(:, k0+1 : k0+n) (blocks of the CSR structure given to set_sparse), as out(25,n)
extern "C" int phb200_get_lhsk_range(phb200_ctx *ctx, long long k0, long long n, double *lhsK) {
  ENTER(ctx);
  if (!lhsK || k0 < 0 || n < 0 || k0 + n > ctx->nnz_tot) return fail("get_lhsk_range", "range outside 0..nnz_tot");
  if (!ctx->have_lhs_sparse) return fail("get_lhsk_range", "no sparse LHS has been assembled");
  PHB_CHECK(cudaMemcpyAsync(lhsK, ctx->d_lhsK + 25 * (size_t)k0, sizeof(double) * 25 * (size_t)n, cudaMemcpyDeviceToHost,
                            ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_elmgmrs(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                              double *BDiag, double *lhsK) {
  ENTER(ctx);
  if (!y || !ac || !st) return fail("elmgmrs", "null argument");
  PHB_TRY(phb200_set_state(ctx, y, ac));
  PHB_TRY(phb_elmgmre(ctx, st, 1));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res, (size_t)5 * ctx->c.nshg));
  if (BDiag && st->iprec) PHB_TRY(d2h(ctx, BDiag, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  if (lhsK && st->lhs == 1) PHB_TRY(get_lhsk(ctx, lhsK));
  return 0;
}
extern "C" int phb200_spsi3pre(phb200_ctx *ctx, double *lhsK) {
  ENTER(ctx);
  if (!ctx->have_lhs_sparse) return fail("spsi3pre", "no sparse LHS has been assembled");
  PHB_TRY(phb_spsi3pre(ctx));
  if (lhsK) PHB_TRY(get_lhsk(ctx, lhsK));
  return 0;
}
extern "C" int phb200_sparseap(phb200_ctx *ctx, double *p) {
  ENTER(ctx);
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *d_u = ctx->d_uBrg;
  PHB_TRY(h2d(ctx, d_u, p, n5));
  PHB_TRY(phb_sparseap(ctx, d_u));
  PHB_TRY(d2h(ctx, p, d_u, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_solgmrs(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                              double *rmes, double *BDiag, double *Dy, double *HBrg, double *eBrg, double *yBrg,
                              double *Rcos, double *Rsin, int *iKs, int *lGMRESs, int *ntotGM) {
  ENTER(ctx);
  if (!y || !ac || !st || !Dy || !iKs || !lGMRESs || !ntotGM) return fail("solgmrs", "null argument");
  if (!ctx->d_lhsK) return fail("solgmrs", "no CSR structure (call phb200_set_sparse with colm/rowp first)");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_TRY(set_state_split(ctx, y, ac));
  PHB_TRY(phb_elmgmre(ctx, st, 1));
  PHB_TRY(phb_solve(ctx, st, 1, iKs, lGMRESs, ntotGM));
  PHB_TRY(d2h(ctx, Dy, ctx->d_Dy, n5));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res, n5));
  if (rmes) PHB_TRY(d2h(ctx, rmes, ctx->d_rmes, n5));
  if (BDiag) PHB_TRY(d2h(ctx, BDiag, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  const int K = ctx->c.Kspace;
  if (HBrg) memcpy(HBrg, ctx->HBrg.data(), sizeof(double) * (size_t)(K + 1) * K);
  if (eBrg) memcpy(eBrg, ctx->eBrg.data(), sizeof(double) * (K + 1));
  if (yBrg) memcpy(yBrg, ctx->yBrg.data(), sizeof(double) * (K + 1));
  if (Rcos) memcpy(Rcos, ctx->Rcos.data(), sizeof(double) * (K + 1));
  if (Rsin) memcpy(Rsin, ctx->Rsin.data(), sizeof(double) * (K + 1));
  return 0;
}
// ---------------------------------------------------------------------------
// matrix-free flavour (solmfg.f, elmmfg.f, itrres.f, au1mfg.f)
// ---------------------------------------------------------------------------
extern "C" int phb200_elmmfg(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                             double *rmes, double *BDiag) {
  ENTER(ctx);
  if (!y || !ac || !st) return fail("elmmfg", "null argument");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  PHB_TRY(phb200_set_state(ctx, y, ac));
  PHB_TRY(phb_elmmfg(ctx, st));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res, n5));
  if (rmes) PHB_TRY(d2h(ctx, rmes, ctx->d_rmes, n5));
  if (BDiag && st->iprec != 0) PHB_TRY(d2h(ctx, BDiag, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// ItrRes (itrres.f): rmes(nshg,5) = modified residual of yp(nshg,5) {u,v,w,p,T}, coefficients frozen at the
// state of the last phb200_elmmfg / phb200_solmfg call
extern "C" int phb200_itrres(phb200_ctx *ctx, const double *yp, double *rmes, int iabres) {
  ENTER(ctx);
  if (!yp || !rmes) return fail("itrres", "null argument");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *d_yp = ctx->d_uBrg, *d_out = ctx->d_uBrg + n5;
  PHB_TRY(h2d(ctx, d_yp, yp, n5));
  PHB_CHECK(cudaMemsetAsync(d_out, 0, sizeof(double) * n5, ctx->stream));
  PHB_TRY(phb_itrres(ctx, d_yp, d_out, iabres));
  PHB_TRY(d2h(ctx, rmes, d_out, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// the set-up of solmfg.f:97-135 on the outputs of phb200_elmmfg (LU_Fact, forward reduction of res and rmes,
// ypre) followed by ONE Au1MFG of u(nshg,5) with the given interval (a parity seam)
extern "C" int phb200_au1mfg(phb200_ctx *ctx, double *u, double eGMRES, int setup) {
  ENTER(ctx);
  if (!u) return fail("au1mfg", "null argument");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  if (setup) {
    PHB_TRY(phb_i3lu(ctx, ctx->d_BDiag, nullptr, 0));
    PHB_TRY(phb_i3lu(ctx, ctx->d_BDiag, ctx->d_res, 1));
    PHB_TRY(phb_i3lu(ctx, ctx->d_BDiag, ctx->d_rmes, 1));
    PHB_TRY(phb_mfg_begin(ctx));
  }
  ctx->eGMRES = eGMRES;
  double *d_u = ctx->d_uBrg;
  PHB_TRY(h2d(ctx, d_u, u, n5));
  PHB_TRY(phb_au1mfg(ctx, d_u));
  PHB_TRY(d2h(ctx, u, d_u, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// SolMFG (solmfg.f:1-381).  eGMRES is COMMON /itrpar/'s finite-difference interval: in/out, recomputed by
// itrFDI when st->iter==1 and mod(st->istep,20)==0.
extern "C" int phb200_solmfg(phb200_ctx *ctx, const double *y, const double *ac, const phb200_step *st, double *res,
                             double *BDiag, double *Dy, double *HBrg, int *iKs, int *lGMRES, int *ntotGM,
                             double *eGMRES) {
  ENTER(ctx);
  if (!y || !ac || !st || !Dy || !iKs || !lGMRES || !ntotGM || !eGMRES) return fail("solmfg", "null argument");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  ctx->eGMRES = *eGMRES;
  PHB_TRY(phb200_set_state(ctx, y, ac));
  PHB_TRY(phb_elmmfg(ctx, st));
  PHB_TRY(phb_solve(ctx, st, 2, iKs, lGMRES, ntotGM));
  PHB_TRY(d2h(ctx, Dy, ctx->d_Dy, n5));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res, n5));
  if (BDiag) PHB_TRY(d2h(ctx, BDiag, ctx->d_BDiag, (size_t)25 * ctx->c.nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  const int K = ctx->c.Kspace;
  if (HBrg) memcpy(HBrg, ctx->HBrg.data(), sizeof(double) * (size_t)(K + 1) * K);
  *eGMRES = ctx->eGMRES;
  return 0;
}
// COMMON /itrpar/ eGMRES (common.h:217): set != 0 stores *e, else reads it back
extern "C" int phb200_egmres(phb200_ctx *ctx, double *e, int set) {
  if (!ctx || !e) return fail("egmres", "null argument");
  if (set) ctx->eGMRES = *e; else *e = ctx->eGMRES;
  return 0;
}
// HBM-resident variants for bench.py: state set by phb200_set_state
extern "C" int phb200_dev_elmmfg(phb200_ctx *ctx, const phb200_step *st) {
  ENTER(ctx);
  return phb_elmmfg(ctx, st);
}
extern "C" int phb200_dev_solve_mfg(phb200_ctx *ctx, const phb200_step *st, int *iKs, int *lGMRES, int *ntotGM) {
  ENTER(ctx);
  return phb_solve(ctx, st, 2, iKs, lGMRES, ntotGM);
}
extern "C" int phb200_dev_au1mfg(phb200_ctx *ctx, int slot) {
  ENTER(ctx);
  if (slot < 0 || slot >= ctx->c.Kspace) return fail("dev_au1mfg", "slot out of range");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *src = ctx->d_uBrg + (size_t)slot * n5, *dst = src + n5;
  PHB_CHECK(cudaMemcpyAsync(dst, src, sizeof(double) * n5, cudaMemcpyDeviceToDevice, ctx->stream));
  return phb_au1mfg(ctx, dst);
}

extern "C" int phb200_dev_elmgmrs(phb200_ctx *ctx, const phb200_step *st) {
  ENTER(ctx);
  return phb_elmgmre(ctx, st, 1);
}
extern "C" int phb200_dev_solve_sparse(phb200_ctx *ctx, const phb200_step *st, int *iKs, int *lGMRESs, int *ntotGM) {
  ENTER(ctx);
  return phb_solve(ctx, st, 1, iKs, lGMRESs, ntotGM);
}
extern "C" int phb200_dev_sparseap(phb200_ctx *ctx, int slot) {
  ENTER(ctx);
  if (slot < 0 || slot >= ctx->c.Kspace) return fail("dev_sparseap", "slot out of range");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *src = ctx->d_uBrg + (size_t)slot * n5, *dst = src + n5;
  PHB_CHECK(cudaMemcpyAsync(ctx->d_ptmp, src, sizeof(double) * n5, cudaMemcpyDeviceToDevice, ctx->stream));
  PHB_TRY(phb_sparseap2(ctx, ctx->d_ptmp, dst, nullptr));
  return phb_bc3per(ctx, dst, 5);
}

// ---- incompressible flavour (incomp.cu) ------------------------------------
extern "C" int phb200_inc_elmgmr(phb200_ctx *ctx, const double *y, const double *ac, const phb200_incomp *ip,
                                 double *res, double *lhsK, double *lhsP) {
  ENTER(ctx);
  if (!y || !ac || !ip) return fail("inc_elmgmr", "null argument");
  PHB_TRY(phb200_set_state(ctx, y, ac));
  PHB_TRY(phb_inc_elmgmr(ctx, ip));
  if (res) PHB_TRY(d2h(ctx, res, ctx->d_res4, (size_t)4 * ctx->c.nshg));
  if (lhsK && ip->lhs == 1) PHB_TRY(d2h(ctx, lhsK, ctx->d_lhsK9, (size_t)9 * ctx->nnz_tot));
  if (lhsP && ip->lhs == 1) PHB_TRY(d2h(ctx, lhsP, ctx->d_lhsP4, (size_t)4 * ctx->nnz_tot));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_inc_dev_elmgmr(phb200_ctx *ctx, const phb200_incomp *ip) {
  ENTER(ctx);
  if (!ip) return fail("inc_dev_elmgmr", "null argument");
  return phb_inc_elmgmr(ctx, ip);
}
extern "C" int phb200_les_ap(phb200_ctx *ctx, int kind, const double *p, double *q) {
  ENTER(ctx);
  if (!p || !q) return fail("les_ap", "null argument");
  if (kind < 0 || kind > 4) return fail("les_ap", "kind must be 0..4");
  if (!ctx->d_lesp) return fail("les_ap", "no incompressible LHS (call phb200_inc_elmgmr first)");
  static const int ncp[5] = {1, 4, 3, 4, 4}, ncq[5] = {3, 3, 1, 1, 4};
  const size_t nshg = ctx->c.nshg;
  PHB_TRY(h2d(ctx, ctx->d_lesp, p, ncp[kind] * nshg));
  PHB_TRY(phb_les_ap(ctx, kind, ctx->d_lesp, ctx->d_lesq));
  PHB_TRY(d2h(ctx, q, ctx->d_lesq, ncq[kind] * nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_inc_dev_apfull(phb200_ctx *ctx) {
  ENTER(ctx);
  if (!ctx->d_lesp) return fail("inc_dev_apfull", "no incompressible LHS (call phb200_inc_elmgmr first)");
  return phb_les_ap(ctx, 4, ctx->d_lesp, ctx->d_lesq);
}

// ---- finer seams on host arrays: stage through d_temp / d_BDiag ------------
extern "C" int phb200_i3lu(phb200_ctx *ctx, double *Diag, double *r, int code) {
  ENTER(ctx);
  const size_t nshg = ctx->c.nshg;
  if (Diag) PHB_TRY(h2d(ctx, ctx->d_BDiag, Diag, 25 * nshg));
  double *d_r = ctx->d_uBrg;  // slot 1 as staging
  if (code != 0) {
    if (!r) return fail("i3lu", "null r");
    PHB_TRY(h2d(ctx, d_r, r, 5 * nshg));
  }
  PHB_TRY(phb_i3lu(ctx, ctx->d_BDiag, d_r, code));
  if (code == 0 && Diag) PHB_TRY(d2h(ctx, Diag, ctx->d_BDiag, 25 * nshg));
  if (code != 0) PHB_TRY(d2h(ctx, r, d_r, 5 * nshg));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_i3pre(phb200_ctx *ctx, double *EGmass) {
  ENTER(ctx);
  PHB_TRY(phb_i3pre(ctx));
  if (EGmass) PHB_TRY(phb200_get_egmass(ctx, EGmass));
  return 0;
}
extern "C" int phb200_au1gmr(phb200_ctx *ctx, double *u) {
  ENTER(ctx);
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *d_u = ctx->d_uBrg;
  PHB_TRY(h2d(ctx, d_u, u, n5));
  PHB_TRY(phb_au1gmr(ctx, d_u));
  PHB_TRY(d2h(ctx, u, d_u, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_bc3per(phb200_ctx *ctx, double *r) {
  ENTER(ctx);
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  double *d_u = ctx->d_uBrg;
  PHB_TRY(h2d(ctx, d_u, r, n5));
  PHB_TRY(phb_bc3per(ctx, d_u, 5));
  PHB_TRY(d2h(ctx, r, d_u, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_commu(phb200_ctx *ctx, double *global, int n, int code) {
  ENTER(ctx);
  if (n < 1 || n > 25) return fail("commu", "n must be 1..25");
  const size_t len = (size_t)n * ctx->c.nshg;
  double *d = (n <= 5) ? ctx->d_uBrg : ctx->d_scratch;
  if (n > 5 && len * sizeof(double) > ctx->scratch_bytes) return fail("commu", "vector too long for the staging buffer");
  PHB_TRY(h2d(ctx, d, global, len));
  PHB_TRY(phb_commu(ctx, d, n, code));
  PHB_TRY(d2h(ctx, global, d, len));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_sumgat(phb200_ctx *ctx, const double *u, int n, double *summed) {
  ENTER(ctx);
  const size_t len = (size_t)n * ctx->c.nshg;
  if (len * sizeof(double) > ctx->scratch_bytes) return fail("sumgat", "vector too long");
  PHB_TRY(h2d(ctx, ctx->d_scratch, u, len));
  return phb_sumgat_dev(ctx, ctx->d_scratch, len, summed);
}

// ---- Newton / time-step shell (timestep.cu) --------------------------------------------------------------
extern "C" int phb200_set_old_state(phb200_ctx *ctx, const double *yold, const double *acold) {
  ENTER(ctx);
  if (!yold || !acold) return fail("set_old_state", "null argument");
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  if (!ctx->d_yold) {
    PHB_TRY(dev_alloc(&ctx->d_yold, n5));
    PHB_TRY(dev_alloc(&ctx->d_acold, n5));
  }
  PHB_TRY(h2d(ctx, ctx->d_yold, yold, n5));
  PHB_TRY(h2d(ctx, ctx->d_acold, acold, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
extern "C" int phb200_get_state(phb200_ctx *ctx, double *y, double *ac, double *yold, double *acold) {
  ENTER(ctx);
  const size_t n5 = (size_t)5 * ctx->c.nshg;
  if ((yold || acold) && !ctx->d_yold) return fail("get_state", "no old state on the device");
  if (y) PHB_TRY(d2h(ctx, y, ctx->d_y, n5));
  if (ac) PHB_TRY(d2h(ctx, ac, ctx->d_ac, n5));
  if (yold) PHB_TRY(d2h(ctx, yold, ctx->d_yold, n5));
  if (acold) PHB_TRY(d2h(ctx, acold, ctx->d_acold, n5));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
#define NEED_OLD(ctx, name) \
  if (!(ctx)->d_yold) return fail(name, "no old state (call phb200_set_old_state first)");
extern "C" int phb200_itrpredict(phb200_ctx *ctx, const phb200_step *st, int ipred) {
  ENTER(ctx);
  NEED_OLD(ctx, "itrpredict");
  return phb_itrpredict(ctx, st, ipred);
}
extern "C" int phb200_itrbc(phb200_ctx *ctx, int ires) {
  ENTER(ctx);
  return phb_itrbc(ctx, ires);
}
extern "C" int phb200_itrcorrect(phb200_ctx *ctx, const phb200_step *st) {
  ENTER(ctx);
  NEED_OLD(ctx, "itrcorrect");
  return phb_itrcorrect(ctx, st);
}
extern "C" int phb200_itrupdate(phb200_ctx *ctx, const phb200_step *st) {
  ENTER(ctx);
  NEED_OLD(ctx, "itrupdate");
  return phb_itrupdate(ctx, st);
}
extern "C" int phb200_rstat(phb200_ctx *ctx, long long nshgt, double *totres) {
  ENTER(ctx);
  if (!totres || nshgt < 1) return fail("rstat", "bad argument");
  return phb_rstat(ctx, nshgt, totres);
}
extern "C" int phb200_timestep(phb200_ctx *ctx, const phb200_step *st, int ipred, int nitr, int sparse, int LHSupd,
                               long long nshgt, int *ntotGM, double *stats) {
  ENTER(ctx);
  NEED_OLD(ctx, "timestep");
  if (!st || !ntotGM || nshgt < 1) return fail("timestep", "bad argument");
  if (sparse && !ctx->d_lhsK) return fail("timestep", "no CSR structure (call phb200_set_sparse first)");
  return phb_timestep(ctx, st, ipred, nitr, sparse, LHSupd, nshgt, ntotGM, stats);
}

// /aerfrc/ (common.h:106): Force(3), HFlux accumulate over calls with iter==nitr
// (e3b.f:325-345; itrdrv zeroes them per step, itrdrv.f:437-442 -> zero=1);
// flxID(10,0:MAXSURF) is per ElmGMRe call (elmgmr.f:122).
extern "C" int phb200_get_aerfrc(phb200_ctx *ctx, double *Force, double *HFlux, double *flxID, int zero) {
  ENTER(ctx);
  std::vector<double> h((size_t)4 + 10 * 1001);
  PHB_CHECK(cudaMemcpyAsync(h.data(), ctx->d_aerfrc, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  if (Force) memcpy(Force, h.data(), sizeof(double) * 3);
  if (HFlux) *HFlux = h[3];
  if (flxID) memcpy(flxID, h.data() + 4, sizeof(double) * 10 * 1001);
  if (zero) PHB_CHECK(cudaMemsetAsync(ctx->d_aerfrc, 0, sizeof(double) * 4, ctx->stream));
  return 0;
}

// ---- instrumentation ---------------------------------------------------------
extern "C" int phb200_event_record(phb200_ctx *ctx, int slot) {
  ENTER(ctx);
  if (slot < 0 || slot >= 16) return fail("event_record", "slot 0..15");
  PHB_CHECK(cudaEventRecord(ctx->ev[slot], ctx->stream));
  return 0;
}
extern "C" int phb200_event_elapsed_ms(phb200_ctx *ctx, int a, int b, float *ms) {
  ENTER(ctx);
  PHB_CHECK(cudaEventSynchronize(ctx->ev[b]));
  PHB_CHECK(cudaEventElapsedTime(ms, ctx->ev[a], ctx->ev[b]));
  return 0;
}
extern "C" int phb200_sync(phb200_ctx *ctx) {
  ENTER(ctx);
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  return phb_p2p_check(ctx);  // a peer wait that timed out shows up here, not as a hang
}
extern "C" long long phb200_launch_count(phb200_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int phb200_profile(phb200_ctx *ctx, int on) {
  ENTER(ctx);
  ctx->profiling = on != 0;
  return 0;
}
extern "C" int phb200_profile_get(phb200_ctx *ctx, int k, float *ms, long long *launches) {
  if (!ctx || k < 0 || k >= KC_N) return 1;
  if (ms) *ms = ctx->kc_ms[k];
  if (launches) *launches = ctx->kc_n[k];
  return 0;
}
extern "C" int phb200_profile_reset(phb200_ctx *ctx) {
  if (!ctx) return 1;
  memset(ctx->kc_ms, 0, sizeof ctx->kc_ms);
  memset(ctx->kc_n, 0, sizeof ctx->kc_n);
  return 0;
}
extern "C" int phb200_fp64_peak(phb200_ctx *ctx, double *tflops) {
  ENTER(ctx);
  return phb_fp64_peak(ctx, tflops);
}
extern "C" int phb200_set_deterministic(phb200_ctx *ctx, int on) {
  ENTER(ctx);
  return phb_set_deterministic(ctx, on);
}
extern "C" int phb200_dmma_peak(phb200_ctx *ctx, double *tflops) {
  ENTER(ctx);
  return phb_dmma_peak(ctx, tflops);
}
extern "C" int phb200_red_peak(phb200_ctx *ctx, long long nblk, double *gadds_per_s) {
  ENTER(ctx);
  if (!gadds_per_s) return fail("red_peak", "null argument");
  if (nblk < 0) return phb_bulkred_peak(ctx, -nblk, gadds_per_s);  // negative: the bulk-copy-engine variant
  return phb_red_peak(ctx, nblk, gadds_per_s);
}
extern "C" int phb200_flush_l2(phb200_ctx *ctx) {
  ENTER(ctx);
  PHB_CHECK(cudaMemsetAsync(ctx->d_scratch, 0, ctx->scratch_bytes, ctx->stream));
  return 0;
}

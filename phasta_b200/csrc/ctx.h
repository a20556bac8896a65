// ctx.h -- device-resident state of one mesh part (one GPU) and shared helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/phb200.h"
#include "halo_task.h"

#define PHB_CHECK(call)                                                        \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "phb200: %s:%d: %s\n", __FILE__, __LINE__,               \
              cudaGetErrorString(e_));                                         \
      return 1;                                                                \
    }                                                                          \
  } while (0)

#define PHB_TRY(expr)            \
  do {                           \
    int r_ = (expr);             \
    if (r_) return r_;           \
  } while (0)

// kernel classes for the built-in profile (phb200_profile_get)
enum { KC_ASM = 0, KC_ASIQ = 1, KC_AP = 2, KC_I3PRE = 3, KC_BLAS = 4, KC_NODE = 5, KC_HALO = 6, KC_N = 7 };

// EGmass device layout: 32-element tiles, each tile is the reference's block
// layout EGmass(npro=32,nedof,nedof) (asigmr.f:36): EG[tile][c][r][lane].
#define EG_TILE 32
__host__ __device__ inline size_t eg_index(size_t e, int r, int c, int nedof) {
  return ((e / EG_TILE) * (size_t)(nedof * nedof) + (size_t)(r + nedof * c)) * EG_TILE + (e % EG_TILE);
}

// interior elements of one non-tet topology (hexes lcsyst=2, wedges lcsyst=3),
// handled by the generic-topology kernels; same tile layout as the tets with
// nd = 5*nshl rows/columns per element matrix
struct ElemGroup {
  int lcsyst, nshl, nq, tab;  // tab: index into the generic shape tables (0 hex, 1 wedge)
  int numel;
  size_t numel_pad;
  int *d_ien;     // [nshl][numel_pad] 0-based
  int *d_refel;   // [numel] 0-based position of the element in the reference's (file) order
  double *d_EG;   // [tile][c][r][32]
  int *d_eloc;    // [nshl*nshl][numel_pad] CSR block of element block (a,b)
};

// boundary elements that are not tets: one group per kind -- hexes (quadrilateral face, lcsyst 2), wedges with a
// triangular (3) or quadrilateral (4) boundary face (elmgmr.f:191 turns the wedge's lcsyst into nenbl)
struct BndGroup {
  int lcsyst, nshl, nshlb, n;
  int *d_ien;     // [nshl][n] 0-based
  int *d_iBCB;    // [2][n]  iBCB(:,1) flux codes, iBCB(:,2) surfID
  double *d_BCB;  // [6][nshlb][n]  BCB(e,k,j) -> [(j*nshlb+k)*n + e]
};


// Krylov scalars resident on the device (solver.cu): offsets, in doubles, into d_kry / h_kry.
// HBrg(Kspace+1,Kspace), eBrg, yBrg, Rcos, Rsin as in solgmr.f:46-49; hcol = the Gram-Schmidt coefficients of the
// column being built; scale = H(iKs+1,iKs) of the newest basis vector (it is normalised when the next iteration
// reads it); red = one slot of PHB_MAILW reduction results per Gram-Schmidt pass of the current iteration.
struct KryLayout {
  int K, npass;
  size_t H, e, y, rc, rs, hcol, scale, epsnrm, red, total;
  __host__ __device__ explicit KryLayout(int K_) : K(K_) {
    npass = K / 4 + 3;
    H = 0;
    e = H + (size_t)(K + 1) * K;
    y = e + K + 1;
    rc = y + K + 1;
    rs = rc + K + 1;
    hcol = rs + K + 1;
    scale = hcol + K + 2;
    epsnrm = scale + 1;
    red = epsnrm + 1;
    total = red + (size_t)npass * PHB_MAILW;
  }
};
static inline size_t phb_kry_doubles(int K) { return KryLayout(K).total; }

struct phb200_ctx {
  phb200_common c;
  int device;
  cudaStream_t stream;
  // second stream for the host->device copy of Y,t in the host-array entry points: it runs under AsIq/qpbc,
  // which only need Y (api.cu set_state_split; phb_elmgmre waits on ev_ac before the first reader of d_ac)
  cudaStream_t cstream;
  cudaEvent_t ev_main, ev_ac;
  bool ac_pending;
  // ---- mesh: lcsyst==1 blocks concatenated in file order (specialised tet kernels);
  //      other topologies live in `gen` (generic kernels)
  std::vector<ElemGroup> gen;
  int *d_refel_tet;    // [numel_tet] position of each tet in the reference's element order
  int numel_tet;       // elements in tet blocks
  size_t numel_pad;    // padded to EG_TILE
  int *d_ien;          // [4][numel_pad] 0-based, padding repeats node 0 (masked)
  int *d_iBC;          // [nshg]
  double *d_BC;        // [ndofBC][nshg]
  int *d_iper;         // [nshg] 0-based
  double *d_x;         // [3][numnp]
  int n_perslave;      // periodic slave nodes (iBC bit 10)
  int *d_perslave;     // their ids
  // ---- boundary elements (tet volume element, tri face = local nodes 1..3)
  int numelb;          // boundary TETS in all blocks (the other kinds: bgen)
  std::vector<BndGroup> bgen;
  int *d_ienb;         // [4][numelb] 0-based
  int *d_iBCB;         // [2][numelb]  iBCB(:,1) flux codes, iBCB(:,2) surfID
  double *d_BCB;       // [6][3][numelb]  BCB(e,n,k) -> [(k*3+n)*numelb + e]
  double *d_aerfrc;    // Force(3), HFlux, then flxID(10,0:MAXSURF)
  // ---- halo (ilwork)
  std::vector<HaloTask> tasks;
  int *d_halo_nodes;   // concatenated node lists of all tasks
  int n_slave_nodes;   // nodes owned by another part (iacc==0 tasks)
  int *d_slave_nodes;
  double *d_sendbuf, *d_recvbuf;
  size_t halo_cap;     // doubles per buffer
  void *nccl;          // ncclComm_t
  // ---- NVLink peer-memory mailboxes (comm.cu): every rank maps every other rank's mailbox through CUDA IPC; small
  //      all-reduces (the Krylov dot products) are done by the reducing kernel itself with peer stores + flags
  double *d_mail;             // this rank's mailbox: vals[2][PHB_MAXR][PHB_MAILW] doubles, then seq[2][PHB_MAXR] u64
  double **d_peer_mail;       // device array [numpe] of mailbox pointers (own pointer at myrank)
  void *peer_mapped[64];      // host copies of the mapped peer pointers (for cudaIpcCloseMemHandle)
  unsigned long long p2p_seq; // all-reduces issued so far (identical on all ranks)
  unsigned int *d_ticket;     // last-block detection for the fused reduction kernels
  int *d_p2p_err;             // set by a kernel whose wait on a peer flag timed out
  bool p2p;
  bool local_group;    // in-process multi-part transport (tests)
  // ---- state / results
  double *d_y, *d_ac;            // [5][nshg] {u,v,w,p,T}
  double *d_yold, *d_acold;      // state at the beginning of the step (timestep.cu), allocated on first use
  int ifuncs;                    // flow solves so far (itrdrv.f:432, drives LHSupd)
  double *d_qres, *d_rmass;      // [12][nshg], [nshg]
  double *d_nodeaos;             // [nshg][26] node records for the element gathers (assembly.cu)
  double *d_res, *d_rmes, *d_Dy, *d_temp;  // [5][nshg]
  double *d_BDiag;               // [25][nshg]  BDiag(nshg,5,5)
  double *d_BDtmp;               // scratch copy for i3pre's commu 'out'
  double *d_EG;                  // tiles, see eg_index (allocated on first EBE lhs=1 call)
  // ---- block-CSR flavour (SolGMRs)
  int nnz_tot;
  int *d_colm, *d_rowp;          // 0-based CSR: colm[nshg+1] row pointers, rowp[nnz_tot] column ids
  int *d_rowofblk;               // row of every block
  int *d_eloc;                   // [16][numel_pad] CSR block of element block (a,b)
  double *d_lhsK;                // [nnz_tot][25]  lhsK(25,nnz_tot)
  int *d_apchunk, n_apchunk;     // row chunks of SparseAp (first row of each, + nshg)
  bool have_lhs_sparse;
  double *d_uBrg;                // [Kspace+1][5][nshg]
  double *d_dots;                // device scalars for fused MGS
  double *h_dots;                // pinned
  // Krylov loop without the host in it (solver.cu): the Ap input with slaves filled, the Hessenberg / Givens state
  // and the per-iteration status words on the device, pinned mirrors, one event per in-flight iteration
  double *d_ptmp;                // [5][nshg]
  double *d_p5;                  // [nshg][5] node-major copy of the vector SparseAp gathers from
  double *d_kry, *h_kry;         // layout: KryLayout (solver.cu)
  int *d_kflag, *h_kflag;        // [0] done, [1] iKs at convergence, [2 + iK] status of iteration iK (1 run, 2 converged)
  cudaEvent_t kev[4];
  double *d_scratch;             // L2 flush / fp64 peak
  size_t scratch_bytes;
  bool have_lhs;                 // EGmass/BDiag hold a (preconditioned) system
  // ---- incompressible flavour (incomp.cu): res(nshg,4), lhsK(9,nnz_tot), lhsP(4,nnz_tot), position of the
  //      transposed CSR entry of every entry, two work vectors [4][nshg] for the lesSparse products
  double *d_res4, *d_lhsK9, *d_lhsP4, *d_lesp, *d_lesq, *d_lesp4;
  int *d_tpos;
  bool have_inc_tabs;
  int inc_idiff;
  std::vector<double> h_shpb, h_shglb; // the same for the boundary-face tables
  bool have_inc_btabs;                 // c_ibnd uploaded, d_nsrflist allocated
  int *d_nsrflist;                     // nsrflist(0:MAXSURF) of the last incompressible call
  bool bnd_deformable;                 // some boundary element carries iBCB bit 4 (deformable wall)
  std::vector<double> h_shp, h_shgl;   // host copies of shp / shgl (the incompressible kernels build their tables lazily)
  // ---- matrix-free flavour (SolMFG): ypre, two work vectors [3][5][nshg]; eGMRES of COMMON /itrpar/
  double *d_mfg;
  double eGMRES;
  bool tet_uniform_rule;         // same N_a,xi and weight at every tet quadrature point
  // host copies of the block structure (pointers stay caller-owned, as mien(iblk)%p does)
  std::vector<int> h_lcblk;
  std::vector<const int *> h_mien;
  // host-side Hessenberg work (solgmr.f:46-49)
  std::vector<double> HBrg, eBrg, yBrg, Rcos, Rsin;
  // ---- deterministic assembly option (assembly.cu: per-element contributions + ordered node gather)
  bool deterministic;
  double *d_elc;                 // [120][numel_pad]
  int *d_inc_ptr, *d_inc;        // node -> (element*4 + local node), ascending in the element id
  // ---- instrumentation
  long long launches;
  cudaEvent_t ev[16];
  bool profiling;
  cudaEvent_t pev0, pev1;
  float kc_ms[KC_N];
  long long kc_n[KC_N];
};

// launch bookkeeping: counts launches, optionally brackets with events
struct KScope {
  phb200_ctx *c;
  int k;
  KScope(phb200_ctx *c_, int k_) : c(c_), k(k_) {
    c->launches++;
    c->kc_n[k]++;
    if (c->profiling) cudaEventRecord(c->pev0, c->stream);
  }
  ~KScope() {
    if (c->profiling) {
      cudaEventRecord(c->pev1, c->stream);
      cudaEventSynchronize(c->pev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, c->pev0, c->pev1);
      c->kc_ms[k] += ms;
    }
  }
};

// assembly.cu
int phb_upload_tables(phb200_ctx *ctx, const double *shp, const double *shgl, const double *shpb,
                      const double *shglb);
int phb_elmgmre(phb200_ctx *ctx, const phb200_step *st, int sparse = 0);
int phb_alloc_eg(phb200_ctx *ctx);
int phb_asires(phb200_ctx *ctx, const double *d_yp, double *d_rmes, int iabres, int ires = 2);
int phb_bc3res_vec(phb200_ctx *ctx, double *d_r);
int phb_qpbc(phb200_ctx *ctx);
int phb_pack_nodes(phb200_ctx *ctx, int with_q);
// solver.cu
int phb_i3lu(phb200_ctx *ctx, double *d_Diag, double *d_r, int code);
int phb_i3pre(phb200_ctx *ctx);
int phb_au1gmr(phb200_ctx *ctx, double *d_u);
int phb_au1gmr2(phb200_ctx *ctx, double *d_p, double *d_out, const int *d_skip);  // d_p is modified (halo, periodic slaves)
int phb_bc3per(phb200_ctx *ctx, double *d_r, int n);
int phb_zero_slaves(phb200_ctx *ctx, double *d_r, int n, int identity);
int phb_sumgat_dev(phb200_ctx *ctx, const double *d_u, size_t len, double *out);
int phb_solve(phb200_ctx *ctx, const phb200_step *st, int sparse, int *iKs, int *lGMRES, int *ntotGM);
int phb_fp64_peak(phb200_ctx *ctx, double *tflops);
int phb_dmma_peak(phb200_ctx *ctx, double *tflops);
int phb_bulkred_peak(phb200_ctx *ctx, long long nblk, double *gadds_per_s);
int phb_red_peak(phb200_ctx *ctx, long long nblk, double *gadds_per_s);
// timestep.cu
int phb_itrpredict(phb200_ctx *ctx, const phb200_step *st, int ipred);
int phb_itrbc(phb200_ctx *ctx, int ires);
int phb_itrbc_vec(phb200_ctx *ctx, double *d_y, double *d_ac, int ires);
// mfg.cu (matrix-free flavour)
int phb_elmmfg(phb200_ctx *ctx, const phb200_step *st);
int phb_itrres(phb200_ctx *ctx, const double *d_yp, double *d_rmes, int iabres, int ires = 2);
int phb_mfg_begin(phb200_ctx *ctx);
int phb_au1mfg(phb200_ctx *ctx, double *d_u);
int phb_au2mfg(phb200_ctx *ctx, double *d_out);
int phb_itrfdi(phb200_ctx *ctx);
int phb_itrcorrect(phb200_ctx *ctx, const phb200_step *st);
int phb_itrupdate(phb200_ctx *ctx, const phb200_step *st);
int phb_rstat(phb200_ctx *ctx, long long nshgt, double *totres);
int phb_timestep(phb200_ctx *ctx, const phb200_step *st, int ipred, int nitr, int sparse, int LHSupd,
                 long long nshgt, int *ntotGM, double *stats);
// incomp.cu (incompressible ElmGMR into block-CSR, lesSparse products)
int phb_inc_elmgmr(phb200_ctx *ctx, const phb200_incomp *ip);
int phb_les_ap(phb200_ctx *ctx, int kind, const double *d_p, double *d_q);
void phb_inc_free(phb200_ctx *ctx);
// sparse.cu
int phb_set_sparse(phb200_ctx *ctx, const int *colm, const int *rowp, int nnz_tot);
int phb_genadj(phb200_ctx *ctx, int nnz, int *colm, int *rowp, int *nnz_tot);
int phb_build_incidence(phb200_ctx *ctx);  // genadj.cu
int phb_set_deterministic(phb200_ctx *ctx, int on);  // assembly.cu
int phb_genadj_dev(phb200_ctx *ctx, int **d_colm0, int **d_rowp0, int **d_rob, long long *nnz_tot);  // genadj.cu
int phb_spsi3pre(phb200_ctx *ctx);
int phb_sparseap(phb200_ctx *ctx, double *d_u);
int phb_sparseap2(phb200_ctx *ctx, double *d_p, double *d_out, const int *d_skip);
// comm.cu
int phb_halo_setup(phb200_ctx *ctx, const int *ilwork);
int phb_commu(phb200_ctx *ctx, double *d_global, int n, int code);
int phb_allreduce_sum(phb200_ctx *ctx, double *d_vals, int n);
int phb_p2p_check(phb200_ctx *ctx);
// PHB_MAXR ranks x PHB_MAILW doubles per mailbox: halo_task.h
// All-reduce of up to PHB_MAILW doubles over NVLink peer memory, done by the kernel that produced them (called by
// ONE warp of one block).  Every rank owns a mailbox  vals[2][MAXR][MAILW] | seq[2][MAXR]  that all peers have
// mapped through CUDA IPC (comm.cu p2p_setup).  For all-reduce number `seq` (the same on all ranks: the solver is
// SPMD) lane r of the calling warp
//   1. stores this rank's partial values into slot [seq&1][me] of rank r's mailbox, fences at system scope, then
//      stores `seq` into the matching flag;
//   2. spins on flag [seq&1][r] of its OWN mailbox until rank r's contribution has arrived;
//   3. lane 0 adds the `world` contributions in rank order -- every rank computes the same bits.
// Two slots (parity of seq) suffice: a rank can only start all-reduce seq+2 after seq+1 completed, and seq+1
// completes only after every rank has contributed to it, i.e. after every rank has finished reading seq.
// The spin is bounded; a time-out raises *err (checked on the host at the end of the solve) instead of hanging.
struct PhbP2P {
  double *const *peer;  // [world] mailbox pointers
  int me, world;
  unsigned long long seq;
  int *err;
};
#ifdef __CUDACC__
static __device__ __forceinline__ void phb_p2p_allreduce_warp(const PhbP2P &p, volatile double *vals, int n) {
  const int lane = threadIdx.x & 31;
  const int par = (int)(p.seq & 1ull);
  const size_t voff = (size_t)par * PHB_MAXR * PHB_MAILW, foff = (size_t)2 * PHB_MAXR * PHB_MAILW;
  if (lane < p.world) {
    volatile double *mb = p.peer[lane];
    for (int k = 0; k < n; k++) mb[voff + (size_t)p.me * PHB_MAILW + k] = vals[k];
    __threadfence_system();
    volatile unsigned long long *fl = reinterpret_cast<volatile unsigned long long *>(p.peer[lane] + foff);
    fl[par * PHB_MAXR + p.me] = p.seq;
  }
  double *mine = p.peer[p.me];
  if (lane < p.world) {
    volatile unsigned long long *fl = reinterpret_cast<volatile unsigned long long *>(mine + foff);
    long long spins = 0;
    const bool dead = *reinterpret_cast<volatile int *>(p.err) != 0;  // an earlier wait timed out: do not wait again
    while (!dead && fl[par * PHB_MAXR + lane] != p.seq) {
      if (++spins > (1ll << 27)) { atomicExch(p.err, 1 + lane); break; }
    }
    __threadfence_system();
  }
  __syncwarp();
  if (lane == 0) {
    volatile double *v = mine + voff;
    for (int k = 0; k < n; k++) {
      double s = 0.0;
      for (int r = 0; r < p.world; r++) s += v[(size_t)r * PHB_MAILW + k];
      vals[k] = s;
    }
  }
  __syncwarp();
}
#endif
PhbP2P phb_p2p_next(phb200_ctx *ctx);
int phb_comm_init(phb200_ctx *ctx, const void *id128);
int phb_comm_unique_id(void *id128);
void phb_comm_free(phb200_ctx *ctx);

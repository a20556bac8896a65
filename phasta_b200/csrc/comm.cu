// comm.cu -- commu (phSolver/common/commu.f:1-297) and the sumgat allreduce
// (common/mpitools.f:107-137) on the device.
//
// ilwork keeps its reference meaning (commu.f:131-143): per task
// {itag, iacc (0 = slave/send on 'in', 1 = master/recv on 'in'), iother,
// numseg, (isgbeg,lenseg)*}.  Each task's segments are flattened once into a
// node list; 'in' = slaves pack -> send, masters recv -> add (dof-outer,
// node-inner, tasks in ilwork order: commu.f:268-294); 'out' = masters pack
// -> send, slaves recv -> overwrite.  Transport is NCCL point-to-point over
// NVLink (one process per GPU), resolved with dlopen so a single-GPU run
// needs no NCCL at all.  A second transport, the in-process "local group"
// (several parts on ONE GPU, one host thread each, a barrier with a time-out), exists
// so the partitioned path can be parity-tested on a single-GPU box.
#include "ctx.h"
#include <dlfcn.h>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>

// ---- minimal NCCL surface (nccl.h 2.27), resolved at run time -------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };
static struct {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
} N;

static int nccl_load() {
  if (N.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !N.h; i++) N.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!N.h) {
    fprintf(stderr, "phb200: comm_init: cannot dlopen libnccl.so.2: %s\n", dlerror());
    return 1;
  }
#define SYM(f)                                                    \
  *(void **)(&N.f) = dlsym(N.h, "nccl" #f);                       \
  if (!N.f) {                                                     \
    fprintf(stderr, "phb200: comm_init: missing nccl" #f "\n");   \
    return 1;                                                     \
  }
  SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(AllGather) SYM(GroupStart)
  SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
  return 0;
}
#define NCCL_CHECK(call)                                                                   \
  do {                                                                                     \
    ncclResult_t r_ = (call);                                                              \
    if (r_ != 0) {                                                                         \
      fprintf(stderr, "phb200: nccl %s:%d: %s\n", __FILE__, __LINE__, N.GetErrorString(r_)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

// ---- in-process local group ------------------------------------------------
// A barrier that gives up: a member that never arrives (its thread died on an exception, or left through an error
// return) turns into an error on every waiting member after PHB200_LOCAL_TIMEOUT_S seconds (default 60) instead of
// a process that can never exit.  `failed` is sticky for the group: once one member has reported an error or timed
// out, every later wait returns at once.
struct TimedBarrier {
  std::mutex mu;
  std::condition_variable cv;
  int n = 0, waiting = 0;
  unsigned long gen = 0;
  bool failed = false;
  void reset(int nranks) {
    std::lock_guard<std::mutex> lk(mu);
    n = nranks;
    waiting = 0;
    gen++;
    failed = false;
  }
  // err != 0: this member arrives carrying an error; everybody (including it) gets 1 back
  int wait(int err = 0) {
    static const double tmo = [] {
      const char *e = getenv("PHB200_LOCAL_TIMEOUT_S");
      double t = e ? atof(e) : 60.0;
      return t > 0 ? t : 60.0;
    }();
    std::unique_lock<std::mutex> lk(mu);
    if (err) failed = true;
    if (failed) {
      cv.notify_all();
      return 1;
    }
    const unsigned long my = gen;
    if (++waiting == n) {
      waiting = 0;
      gen++;
      cv.notify_all();
      return 0;
    }
    const bool ok = cv.wait_for(lk, std::chrono::duration<double>(tmo), [&] { return gen != my || failed; });
    if (!ok) {
      failed = true;
      cv.notify_all();
      fprintf(stderr, "phb200: local group: barrier timed out after %.0f s (%d of %d members arrived)\n", tmo, waiting, n);
      return 1;
    }
    return failed ? 1 : 0;
  }
};

struct LocalGroup {
  int n = 0;
  TimedBarrier bar;
  phb200_ctx *members[64] = {};
  double red[64][PHB_MAILW] = {};
};
static LocalGroup g_local;
static std::mutex g_local_mu;

extern "C" int phb200_local_group_join(phb200_ctx *ctx, int nranks) {
  std::lock_guard<std::mutex> lk(g_local_mu);
  if (nranks > 64 || nranks < 1 || ctx->c.myrank < 0 || ctx->c.myrank >= nranks) return 1;
  // a new group (different size, or the slot is taken by another live context) starts with a clean barrier
  if (g_local.n != nranks || g_local.bar.failed || (g_local.members[ctx->c.myrank] && g_local.members[ctx->c.myrank] != ctx)) {
    g_local.bar.reset(nranks);
    g_local.n = nranks;
    for (auto &m : g_local.members) m = nullptr;
  }
  g_local.members[ctx->c.myrank] = ctx;
  ctx->nccl = nullptr;
  ctx->local_group = true;
  return 0;
}

// ---------------------------------------------------------------------------
int phb_halo_setup(phb200_ctx *ctx, const int *il) {
  ctx->tasks.clear();
  ctx->d_halo_nodes = nullptr;
  ctx->d_slave_nodes = nullptr;
  ctx->n_slave_nodes = 0;
  ctx->d_sendbuf = ctx->d_recvbuf = nullptr;
  ctx->halo_cap = 0;
  if (ctx->c.numpe <= 1 || !il || ctx->c.nlwork < 1) return 0;
  std::vector<int> nodes, slaves;
  phb_parse_ilwork(il, ctx->tasks, nodes, slaves);
  if (nodes.empty()) return 0;
  PHB_CHECK(cudaMalloc(&ctx->d_halo_nodes, sizeof(int) * nodes.size()));
  PHB_CHECK(cudaMemcpy(ctx->d_halo_nodes, nodes.data(), sizeof(int) * nodes.size(), cudaMemcpyHostToDevice));
  if (!slaves.empty()) {
    PHB_CHECK(cudaMalloc(&ctx->d_slave_nodes, sizeof(int) * slaves.size()));
    PHB_CHECK(cudaMemcpy(ctx->d_slave_nodes, slaves.data(), sizeof(int) * slaves.size(), cudaMemcpyHostToDevice));
    ctx->n_slave_nodes = (int)slaves.size();
  }
  ctx->halo_cap = nodes.size() * 25;
  PHB_CHECK(cudaMalloc(&ctx->d_sendbuf, sizeof(double) * ctx->halo_cap));
  PHB_CHECK(cudaMalloc(&ctx->d_recvbuf, sizeof(double) * ctx->halo_cap));
  return 0;
}

// buf[(k*count + t)] <-> global[node_t + nshg*k]
__global__ void k_halo_pack(int count, const int *__restrict__ nodes, int nshg, int n, const double *__restrict__ g,
                            double *__restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * n) return;
  int k = t / count, i = t % count;
  buf[t] = g[(size_t)nshg * k + nodes[i]];
}
__global__ void k_halo_unpack(int count, const int *__restrict__ nodes, int nshg, int n, double *__restrict__ g,
                              const double *__restrict__ buf, int add) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * n) return;
  int k = t / count, i = t % count;
  double *p = g + (size_t)nshg * k + nodes[i];
  *p = add ? (*p + buf[t]) : buf[t];
}

int phb_commu(phb200_ctx *ctx, double *g, int n, int code) {
  if (ctx->c.numpe <= 1 || ctx->tasks.empty()) return 0;
  if (!ctx->nccl && !ctx->local_group) {
    fprintf(stderr, "phb200: commu: numpe=%d but no communicator (call phb200_comm_init)\n", ctx->c.numpe);
    return 1;
  }
  if (n > 25) {
    fprintf(stderr, "phb200: commu: n=%d > 25 unsupported\n", n);
    return 1;
  }
  cudaStream_t s = ctx->stream;
  const int nshg = ctx->c.nshg;
  // sender role: iacc==0 on 'in', iacc==1 on 'out'
  const int send_role = (code == 0) ? 0 : 1;
  {
    KScope ks(ctx, KC_HALO);
    for (auto &h : ctx->tasks)
      if (h.iacc == send_role && h.count > 0) {
        int tot = h.count * n;
        k_halo_pack<<<(tot + 255) / 256, 256, 0, s>>>(h.count, ctx->d_halo_nodes + h.offset, nshg, n, g,
                                                      ctx->d_sendbuf + (size_t)h.offset * 25);
      }
    PHB_CHECK(cudaGetLastError());
  }
  if (ctx->nccl) {
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    NCCL_CHECK(N.GroupStart());
    for (auto &h : ctx->tasks) {
      size_t cnt = (size_t)h.count * n;
      if (h.iacc == send_role)
        NCCL_CHECK(N.Send(ctx->d_sendbuf + (size_t)h.offset * 25, cnt, ncclFloat64, h.peer, comm, s));
      else
        NCCL_CHECK(N.Recv(ctx->d_recvbuf + (size_t)h.offset * 25, cnt, ncclFloat64, h.peer, comm, s));
    }
    NCCL_CHECK(N.GroupEnd());
  } else {
    // local group: every rank's packed data must be complete before peers read it
    // (errors travel through the barrier so that no member is left waiting for one that has returned)
    int err = cudaStreamSynchronize(s) != cudaSuccess;
    if (g_local.bar.wait(err)) return 1;
    for (auto &h : ctx->tasks)
      if (h.iacc != send_role && !err) {
        phb200_ctx *peer = g_local.members[h.peer];
        const HaloTask *ph = nullptr;
        if (peer)
          for (auto &t : peer->tasks)
            if (t.tag == h.tag && t.peer == ctx->c.myrank && t.iacc == send_role) ph = &t;
        if (!ph || ph->count != h.count) {
          fprintf(stderr, "phb200: commu: unmatched task tag %d\n", h.tag);
          err = 1;
          break;
        }
        err = cudaMemcpyAsync(ctx->d_recvbuf + (size_t)h.offset * 25, peer->d_sendbuf + (size_t)ph->offset * 25,
                              sizeof(double) * h.count * n, cudaMemcpyDeviceToDevice, s) != cudaSuccess;
      }
    err |= cudaStreamSynchronize(s) != cudaSuccess;
    if (g_local.bar.wait(err)) return 1;
  }
  {
    KScope ks(ctx, KC_HALO);
    for (auto &h : ctx->tasks)
      if (h.iacc != send_role && h.count > 0) {
        int tot = h.count * n;
        k_halo_unpack<<<(tot + 255) / 256, 256, 0, s>>>(h.count, ctx->d_halo_nodes + h.offset, nshg, n, g,
                                                        ctx->d_recvbuf + (size_t)h.offset * 25, code == 0);
      }
    PHB_CHECK(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------
// small all-reduces over NVLink peer memory: protocol and device routine in ctx.h (phb_p2p_allreduce_warp)
__global__ void k_p2p_allreduce(PhbP2P p, double *vals, int n) {
  phb_p2p_allreduce_warp(p, vals, n);
}

PhbP2P phb_p2p_next(phb200_ctx *ctx) {
  PhbP2P p;
  p.peer = ctx->d_peer_mail;
  p.me = ctx->c.myrank;
  p.world = ctx->c.numpe;
  p.seq = ++ctx->p2p_seq;
  p.err = ctx->d_p2p_err;
  return p;
}

int phb_p2p_check(phb200_ctx *ctx) {
  if (!ctx->p2p) return 0;
  int e = 0;
  PHB_CHECK(cudaMemcpyAsync(&e, ctx->d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  if (e) {
    fprintf(stderr, "phb200: peer all-reduce timed out waiting for rank %d\n", e - 1);
    return 1;
  }
  return 0;
}

int phb_allreduce_sum(phb200_ctx *ctx, double *d_vals, int n) {
  if (ctx->c.numpe <= 1) return 0;
  if (ctx->p2p && n <= PHB_MAILW) {
    KScope ks(ctx, KC_HALO);
    k_p2p_allreduce<<<1, 32, 0, ctx->stream>>>(phb_p2p_next(ctx), d_vals, n);
    PHB_CHECK(cudaGetLastError());
    return 0;
  }
  if (ctx->nccl) {
    NCCL_CHECK(N.AllReduce(d_vals, d_vals, (size_t)n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    return 0;
  }
  if (ctx->local_group) {
    if (n < 1 || n > PHB_MAILW) return 1;
    double v[PHB_MAILW];
    int err = cudaMemcpyAsync(v, d_vals, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess;
    err |= cudaStreamSynchronize(ctx->stream) != cudaSuccess;
    for (int k = 0; k < n; k++) g_local.red[ctx->c.myrank][k] = v[k];
    if (g_local.bar.wait(err)) return 1;
    double s[PHB_MAILW] = {};
    for (int r = 0; r < g_local.n; r++)
      for (int k = 0; k < n; k++) s[k] += g_local.red[r][k];
    if (g_local.bar.wait(0)) return 1;
    PHB_CHECK(cudaMemcpyAsync(d_vals, s, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    PHB_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
  }
  fprintf(stderr, "phb200: allreduce: numpe=%d but no communicator\n", ctx->c.numpe);
  return 1;
}

// map every rank's mailbox into this process (handles travel through one ncclAllGather).
// Every rank takes part in the same collectives whatever happens locally: a failure (no IPC handle, a mapping that
// does not open) only lowers this rank's vote, and the unanimous all-reduce at the end decides for everybody, so no
// rank is left waiting in a collective another one has skipped.  On "no" everything allocated here is released and
// the small all-reduces stay on NCCL.
static int p2p_setup(phb200_ctx *ctx) {
  const int world = ctx->c.numpe, me = ctx->c.myrank;
  ctx->p2p = false;
  const char *env = getenv("PHB200_P2P");
  if (env && atoi(env) == 0) return 0;
  if (world > PHB_MAXR) return 0;
  const size_t mail_dbl = phb_mailbox_doubles();
  int ok = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  unsigned char *d_h = nullptr;
  double *d_ok = nullptr;
  ok &= cudaMalloc(&ctx->d_mail, sizeof(double) * mail_dbl) == cudaSuccess;
  ok &= cudaMalloc(&ctx->d_ticket, sizeof(unsigned int)) == cudaSuccess;
  ok &= cudaMalloc(&ctx->d_p2p_err, sizeof(int)) == cudaSuccess;
  ok &= cudaMalloc(&d_h, (size_t)(world + 1) * sizeof(mine)) == cudaSuccess;
  ok &= cudaMalloc(&d_ok, sizeof(double)) == cudaSuccess;
  if (!ok) {  // out of device memory: nothing collective can be attempted at all
    cudaGetLastError();
    fprintf(stderr, "phb200: comm_init: cannot allocate the peer mailbox\n");
    return 1;
  }
  cudaMemset(ctx->d_mail, 0, sizeof(double) * mail_dbl);
  cudaMemset(ctx->d_ticket, 0, sizeof(unsigned int));
  cudaMemset(ctx->d_p2p_err, 0, sizeof(int));
  if (cudaIpcGetMemHandle(&mine, ctx->d_mail) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  PHB_CHECK(cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  NCCL_CHECK(N.AllGather(d_h, d_h + sizeof(mine), sizeof(mine), /*ncclInt8*/ 0, (ncclComm_t)ctx->nccl, ctx->stream));
  std::vector<cudaIpcMemHandle_t> all(world);
  PHB_CHECK(cudaMemcpyAsync(all.data(), d_h + sizeof(mine), (size_t)world * sizeof(mine), cudaMemcpyDeviceToHost,
                            ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_h);
  // a rank that got no handle votes "no" first, so that nobody tries to open its (zeroed) handle
  double okd = ok ? 0.0 : 1.0;
  PHB_CHECK(cudaMemcpy(d_ok, &okd, sizeof(double), cudaMemcpyHostToDevice));
  NCCL_CHECK(N.AllReduce(d_ok, d_ok, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  PHB_CHECK(cudaMemcpyAsync(&okd, d_ok, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  std::vector<double *> ptrs(world, nullptr);
  if (okd == 0.0) {
    for (int r = 0; r < world; r++) {
      ctx->peer_mapped[r] = nullptr;
      if (r == me) { ptrs[r] = ctx->d_mail; continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      ctx->peer_mapped[r] = p;
      ptrs[r] = (double *)p;
    }
  } else {
    ok = 0;
  }
  // every rank must take the same decision: agree through one NCCL all-reduce
  okd = ok ? 0.0 : 1.0;
  PHB_CHECK(cudaMemcpy(d_ok, &okd, sizeof(double), cudaMemcpyHostToDevice));
  NCCL_CHECK(N.AllReduce(d_ok, d_ok, 1, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  PHB_CHECK(cudaMemcpyAsync(&okd, d_ok, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PHB_CHECK(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_ok);
  if (okd != 0.0) {
    if (me == 0) fprintf(stderr, "phb200: comm_init: peer mapping unavailable on %d rank(s); small all-reduces stay on NCCL\n", (int)okd);
    for (int r = 0; r < world && r < 64; r++)
      if (ctx->peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->peer_mapped[r]); ctx->peer_mapped[r] = nullptr; }
    cudaFree(ctx->d_mail); cudaFree(ctx->d_ticket); cudaFree(ctx->d_p2p_err);
    ctx->d_mail = nullptr; ctx->d_ticket = nullptr; ctx->d_p2p_err = nullptr;
    return 0;
  }
  PHB_CHECK(cudaMalloc(&ctx->d_peer_mail, sizeof(double *) * world));
  PHB_CHECK(cudaMemcpy(ctx->d_peer_mail, ptrs.data(), sizeof(double *) * world, cudaMemcpyHostToDevice));
  ctx->p2p_seq = 0;
  ctx->p2p = true;
  return 0;
}

int phb_comm_unique_id(void *id128) {
  PHB_TRY(nccl_load());
  ncclUniqueId id;
  NCCL_CHECK(N.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

int phb_comm_init(phb200_ctx *ctx, const void *id128) {
  PHB_TRY(nccl_load());
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  PHB_CHECK(cudaSetDevice(ctx->device));
  NCCL_CHECK(N.CommInitRank(&comm, ctx->c.numpe, id, ctx->c.myrank));
  ctx->nccl = comm;
  ctx->local_group = false;
  PHB_TRY(p2p_setup(ctx));
  return 0;
}

void phb_comm_free(phb200_ctx *ctx) {
  if (ctx->p2p) {
    // drain this rank.s stream before the peers. mappings of its mailbox go away
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < ctx->c.numpe && r < 64; r++)
      if (ctx->peer_mapped[r]) cudaIpcCloseMemHandle(ctx->peer_mapped[r]);
    ctx->p2p = false;
  }
  if (ctx->d_mail) cudaFree(ctx->d_mail);
  if (ctx->d_peer_mail) cudaFree(ctx->d_peer_mail);
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  if (ctx->d_p2p_err) cudaFree(ctx->d_p2p_err);
  ctx->d_mail = nullptr; ctx->d_peer_mail = nullptr; ctx->d_ticket = nullptr; ctx->d_p2p_err = nullptr;
  if (ctx->nccl && N.CommDestroy) N.CommDestroy((ncclComm_t)ctx->nccl);
  ctx->nccl = nullptr;
  if (ctx->local_group) {
    std::lock_guard<std::mutex> lk(g_local_mu);
    g_local.members[ctx->c.myrank] = nullptr;
  }
}

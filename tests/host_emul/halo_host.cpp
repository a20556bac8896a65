// halo_host.cpp -- TEST INFRASTRUCTURE.  One PROCESS per rank runs the product's peer-store halo transport on the host:
// the task parsing, arena layout, task pairing and message addressing of halo_task.h and the kernels of
// halo_p2p.cuh (under the SIMT shim), over arenas that live in one shared-memory mapping the way the GPUs' arenas
// are mapped into each other through CUDA IPC.  The ranks synchronise through the protocol's own flags and
// acknowledgements only.  Not a fallback: nothing in phasta_b200/ loads it.
#include <sched.h>
#include "cuda_shim_simt.h"
#define PHB_SPIN_MAX (1ll << 34)
#define PHB_SPIN_PAUSE() sched_yield()
#include "../../phasta_b200/csrc/halo_task.h"
#include "../../phasta_b200/csrc/halo_p2p.cuh"

static std::vector<HaloTask> g_tasks;
static std::vector<int> g_nodes, g_slaves;
static std::vector<unsigned> g_tickets;
static size_t g_halo_cap;
static int g_err, g_me, g_world;
static double *g_base;
static size_t g_stride;

extern "C" long halo_host_arena_total(long halo_cap) { return (long)phb_arena_layout((size_t)halo_cap).total; }
extern "C" int halo_host_table_words() { return PHB_P2P_W; }

// parse ilwork, publish this rank's table; returns halo_cap
extern "C" long halo_host_setup(int me, int world, const int *ilwork, double *arenas, long stride, int *tables) {
  g_tasks.clear(); g_nodes.clear(); g_slaves.clear();
  phb_parse_ilwork(ilwork, g_tasks, g_nodes, g_slaves);
  g_halo_cap = g_nodes.size() * 25;
  g_tickets.assign(g_tasks.size() + 1, 0u);
  g_err = 0; g_me = me; g_world = world; g_base = arenas; g_stride = (size_t)stride;
  phb_p2p_mytab(g_tasks, g_halo_cap, tables + (size_t)PHB_P2P_W * me);
  return (long)g_halo_cap;
}
extern "C" int halo_host_pair(const int *tables) { return phb_p2p_pair(g_me, g_world, tables, g_tasks) ? 1 : 0; }

// commu(global(nshg,n), 'in ' | 'out') by peer stores: the loop of comm.cu commu_p2p, launches through the shim
extern "C" int halo_host_commu(double *g, int nshg, int n, int code) {
  const int send_role = (code == 0) ? 0 : 1;
  auto arena = [&](int r) { return g_base + g_stride * (size_t)r; };
  const PhbArena A = phb_arena_layout(g_halo_cap);
  for (size_t ti = 0; ti < g_tasks.size(); ti++) {
    HaloTask &h = g_tasks[ti];
    if (h.iacc != send_role) continue;
    const PhbHaloMsg m = phb_p2p_send_msg(h, ti, n, arena(g_me), arena(h.peer), A);
    const int count = h.count;
    const int *nodes = g_nodes.data() + h.offset;
    unsigned *tk = g_tickets.data() + ti;
    shim_launch((m.tot + 255) / 256, 256,
                [=]() { k_halo_send(count, nodes, nshg, n, g, m.data, m.flag, m.ack, m.msg, tk, &g_err); });
  }
  for (size_t ti = 0; ti < g_tasks.size(); ti++) {
    HaloTask &h = g_tasks[ti];
    if (h.iacc == send_role) continue;
    const PhbHaloMsg m = phb_p2p_recv_msg(h, ti, n, arena(g_me), arena(h.peer), A, g_halo_cap);
    const int count = h.count;
    const int *nodes = g_nodes.data() + h.offset;
    unsigned *tk = g_tickets.data() + ti;
    shim_launch((m.tot + 255) / 256, 256, [=]() {
      k_halo_recv(count, nodes, nshg, n, g, m.data, m.flag, m.ack, m.msg, code == 0, tk, &g_err);
    });
  }
  return g_err;
}

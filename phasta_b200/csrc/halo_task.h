// halo_task.h -- the host-side bookkeeping of the partition-boundary exchange, plain C++ (no CUDA calls) so that
// comm.cu and the multi-process host emulation (tests/host_emul/halo_host.cpp) share one copy: the ilwork tasks
// (common/commu.f:131-143, ctypes.f:40-120), the layout of the peer-visible arena, the pairing of a rank's tasks with
// its peers' and the addresses a send / receive touches.
#pragma once
#include <cstddef>
#include <vector>

#define PHB_MAXR 16   // ranks a mailbox has room for
#define PHB_MAILW 12  // doubles per all-reduce (one blocked Gram-Schmidt pass: 4 dots + 6 Gram entries + 1 norm)

struct HaloTask {
  int peer, iacc, tag, count;  // count = number of nodes (all segments)
  int offset;                  // into d_halo_nodes
  // NVLink peer-memory transport (comm.cu): the matching task on the peer, where its data lands in the peer's
  // arena, and how many messages this rank has sent / received on this task
  int peer_task, peer_offset;
  size_t peer_cap;             // the peer's halo_cap = stride between its two data slots
  unsigned long long sendn, recvn;
};

// ilwork (after ctypes.f:47, iother 0-based): numtask, then per task tag, iacc, iother, numseg, (isgbeg, lenseg)*
static inline void phb_parse_ilwork(const int *il, std::vector<HaloTask> &tasks, std::vector<int> &nodes,
                                    std::vector<int> &slaves) {
  int numtask = il[0], itk = 1;
  for (int t = 0; t < numtask; t++) {
    HaloTask h;
    h.tag = il[itk];
    h.iacc = il[itk + 1];
    h.peer = il[itk + 2];
    int numseg = il[itk + 3];
    h.offset = (int)nodes.size();
    for (int s = 0; s < numseg; s++) {
      int beg = il[itk + 4 + 2 * s], len = il[itk + 5 + 2 * s];
      for (int k = 0; k < len; k++) {
        nodes.push_back(beg + k - 1);
        if (h.iacc == 0) slaves.push_back(beg + k - 1);
      }
    }
    h.count = (int)nodes.size() - h.offset;
    h.peer_task = -1;
    h.peer_offset = 0;
    h.peer_cap = 0;
    h.sendn = h.recvn = 0;
    tasks.push_back(h);
    itk += 4 + 2 * numseg;
  }
}

// One peer-visible arena per rank: all-reduce mailbox | halo flags [64][2] | halo acks [64] | halo data [2][halo_cap].
// The offsets (in doubles from the arena base) are the same on every rank -- a rank addresses its PEERS' arenas
// with them; only the data slot stride (the peer's halo_cap) differs per rank and travels in the task table.
#define PHB_P2P_MAXT 64
struct PhbArena {
  size_t flag_off, ack_off, data_off, total;
};
static inline PhbArena phb_arena_layout(size_t halo_cap) {
  PhbArena A;
  const size_t mail_dbl = (size_t)2 * PHB_MAXR * PHB_MAILW + (size_t)2 * PHB_MAXR;
  A.flag_off = mail_dbl;
  A.ack_off = A.flag_off + 2 * PHB_P2P_MAXT;
  A.data_off = A.ack_off + PHB_P2P_MAXT;
  A.total = A.data_off + 2 * halo_cap;
  return A;
}

// the table a rank publishes: [ntask or -1, halo_cap, (tag, iacc, peer, offset, count) * PHB_P2P_MAXT]
#define PHB_P2P_REC 5
#define PHB_P2P_W (2 + PHB_P2P_MAXT * PHB_P2P_REC)
static inline void phb_p2p_mytab(const std::vector<HaloTask> &tasks, size_t halo_cap, int *mytab) {
  for (int i = 0; i < PHB_P2P_W; i++) mytab[i] = 0;
  const size_t ntask = tasks.size();
  const int fits = ((int)ntask <= PHB_P2P_MAXT && halo_cap < ((size_t)1 << 31)) ? 1 : 0;
  mytab[0] = fits ? (int)ntask : -1;
  mytab[1] = (int)halo_cap;
  for (size_t t = 0; t < ntask && fits; t++) {
    const HaloTask &h = tasks[t];
    int *r = &mytab[2 + PHB_P2P_REC * t];
    r[0] = h.tag; r[1] = h.iacc; r[2] = h.peer; r[3] = h.offset; r[4] = h.count;
  }
}
// alltab: the `world` tables one after the other.  Fills peer_task / peer_offset / peer_cap of every task of rank
// `me`; false when a table did not fit or a task has no partner (same tag, opposite role, same length, pointing back)
static inline bool phb_p2p_pair(int me, int world, const int *alltab, std::vector<HaloTask> &tasks) {
  bool good = true;
  for (int r = 0; r < world; r++)
    if (alltab[(size_t)PHB_P2P_W * r] < 0) good = false;
  for (size_t t = 0; t < tasks.size() && good; t++) {
    HaloTask &h = tasks[t];
    const int *pt = &alltab[(size_t)PHB_P2P_W * h.peer];
    h.peer_task = -1;
    for (int k = 0; k < pt[0]; k++) {
      const int *r = pt + 2 + PHB_P2P_REC * k;
      if (r[0] == h.tag && r[2] == me && r[1] != h.iacc && r[4] == h.count) {
        h.peer_task = k;
        h.peer_offset = r[3];
        h.peer_cap = (size_t)pt[1];
      }
    }
    if (h.peer_task < 0) good = false;
  }
  return good;
}

// what one message touches.  Message m of a task uses data slot m & 1; the sender stores into the RECEIVER's arena
// (slot stride = the receiver's halo_cap, offset = the receiver's task offset) and raises the receiver's flag
// [peer_task][slot]; the receiver reads its own arena and acknowledges into the SENDER's arena at acks[peer_task].
struct PhbHaloMsg {
  double *data;                          // send: destination in the peer's arena; receive: source in my arena
  volatile unsigned long long *flag;     // send: the peer's flag; receive: my flag
  volatile unsigned long long *ack;      // send: my ack (written by the peer); receive: the peer's ack
  unsigned long long msg;
  int tot;
};
static inline PhbHaloMsg phb_p2p_send_msg(HaloTask &h, size_t ti, int n, double *my_arena, double *peer_arena,
                                          const PhbArena &A) {
  PhbHaloMsg m;
  m.msg = ++h.sendn;
  const int slot = (int)(m.msg & 1ull);
  m.tot = h.count * n;
  m.data = peer_arena + A.data_off + (size_t)slot * h.peer_cap + (size_t)h.peer_offset * 25;
  m.flag = reinterpret_cast<volatile unsigned long long *>(peer_arena + A.flag_off) + (2 * h.peer_task + slot);
  m.ack = reinterpret_cast<volatile unsigned long long *>(my_arena + A.ack_off) + ti;
  return m;
}
static inline PhbHaloMsg phb_p2p_recv_msg(HaloTask &h, size_t ti, int n, double *my_arena, double *peer_arena,
                                          const PhbArena &A, size_t my_halo_cap) {
  PhbHaloMsg m;
  m.msg = ++h.recvn;
  const int slot = (int)(m.msg & 1ull);
  m.tot = h.count * n;
  m.data = my_arena + A.data_off + (size_t)slot * my_halo_cap + (size_t)h.offset * 25;
  m.flag = reinterpret_cast<volatile unsigned long long *>(my_arena + A.flag_off) + (2 * ti + slot);
  m.ack = reinterpret_cast<volatile unsigned long long *>(peer_arena + A.ack_off) + h.peer_task;
  return m;
}

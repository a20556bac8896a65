// halo_task.h -- the host-side bookkeeping of the partition-boundary exchange, plain C++ (no CUDA calls): the ilwork
// tasks (common/commu.f:131-143, ctypes.f:40-120) and the size of the peer-visible all-reduce mailbox.
#pragma once
#include <cstddef>
#include <vector>

#define PHB_MAXR 16   // ranks a mailbox has room for
#define PHB_MAILW 12  // doubles per all-reduce (one blocked Gram-Schmidt pass: 4 dots + 6 Gram entries + 1 norm)

struct HaloTask {
  int peer, iacc, tag, count;  // count = number of nodes (all segments)
  int offset;                  // into d_halo_nodes
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
    tasks.push_back(h);
    itk += 4 + 2 * numseg;
  }
}

// the mailbox every rank exposes to its peers: vals[2][PHB_MAXR][PHB_MAILW] doubles, then seq[2][PHB_MAXR] u64 (ctx.h)
static inline size_t phb_mailbox_doubles() { return (size_t)2 * PHB_MAXR * PHB_MAILW + (size_t)2 * PHB_MAXR; }

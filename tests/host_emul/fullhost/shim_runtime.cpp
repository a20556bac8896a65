// shim_runtime.cpp -- the fiber scheduler behind tests/host_emul/fullhost/cuda_runtime.h (TEST INFRASTRUCTURE)
#include "cuda_runtime.h"

shim_uint3 threadIdx, blockIdx, blockDim, gridDim;
unsigned char smem_raw[1 << 18] __attribute__((aligned(16)));
uint64_t shim_warp_buf[64][32];

namespace {
enum { RUN = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };
constexpr size_t STACK = 256 * 1024;
struct Fiber {
  ucontext_t ctx;
  int state;
};
std::recursive_mutex g_mu;
std::vector<Fiber> g_fib;
char *g_stacks = nullptr;
size_t g_nstacks = 0;
ucontext_t g_sched;
int g_cur = -1;
const std::function<void()> *g_fn = nullptr;

void trampoline() {
  (*g_fn)();
  g_fib[g_cur].state = DONE;
  swapcontext(&g_fib[g_cur].ctx, &g_sched);
}
}  // namespace

void shim_barrier(int warp_level) {
  Fiber &f = g_fib[g_cur];
  f.state = warp_level ? AT_WARP : AT_BLOCK;
  swapcontext(&f.ctx, &g_sched);
}

void shim_launch(dim3 grid3, unsigned block, const std::function<void()> &kernel) {
  const unsigned grid = grid3.x * grid3.y;
  std::lock_guard<std::recursive_mutex> lock(g_mu);   // launches from several host threads (in-process parts) serialise
  if (block == 0 || grid == 0) return;
  if (block > 2048) { fprintf(stderr, "shim_launch: block of %u threads\n", block); abort(); }
  if (g_nstacks < block) {
    if (g_stacks) munmap(g_stacks, g_nstacks * STACK);
    g_stacks = (char *)mmap(nullptr, (size_t)block * STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    g_nstacks = block;
  }
  g_fib.resize(block);
  blockDim = {block, 1, 1};
  gridDim = {grid3.x, grid3.y, 1};
  g_fn = &kernel;
  for (unsigned b = 0; b < grid; b++) {
    blockIdx = {b % grid3.x, b / grid3.x, 0};
    for (unsigned t = 0; t < block; t++) {
      getcontext(&g_fib[t].ctx);
      g_fib[t].ctx.uc_stack.ss_sp = g_stacks + (size_t)t * STACK;
      g_fib[t].ctx.uc_stack.ss_size = STACK;
      g_fib[t].ctx.uc_link = &g_sched;
      makecontext(&g_fib[t].ctx, trampoline, 0);
      g_fib[t].state = RUN;
    }
    unsigned done = 0;
    while (done < block) {
      bool progress = false;
      for (unsigned t = 0; t < block; t++) {
        if (g_fib[t].state != RUN) continue;
        g_cur = (int)t;
        threadIdx = {t, 0, 0};
        swapcontext(&g_sched, &g_fib[t].ctx);
        progress = true;
        if (g_fib[t].state == DONE) done++;
      }
      // release the warp barriers every live fiber of the warp has reached, then the block barrier
      for (unsigned w = 0; w * 32 < block; w++) {
        bool all = true, any = false;
        for (unsigned t = w * 32; t < block && t < w * 32 + 32; t++) {
          if (g_fib[t].state == DONE) continue;
          if (g_fib[t].state == AT_WARP) any = true; else all = false;
        }
        if (all && any) {
          for (unsigned t = w * 32; t < block && t < w * 32 + 32; t++)
            if (g_fib[t].state == AT_WARP) g_fib[t].state = RUN;
          progress = true;
        }
      }
      bool all = true, any = false;
      for (unsigned t = 0; t < block; t++) {
        if (g_fib[t].state == DONE) continue;
        if (g_fib[t].state == AT_BLOCK) any = true; else all = false;
      }
      if (all && any) {
        for (unsigned t = 0; t < block; t++)
          if (g_fib[t].state == AT_BLOCK) g_fib[t].state = RUN;
        progress = true;
      }
      if (!progress) { fprintf(stderr, "shim_launch: barrier deadlock in block %u\n", b); abort(); }
    }
  }
  g_fn = nullptr;
}

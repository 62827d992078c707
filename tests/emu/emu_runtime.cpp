// TEST INFRASTRUCTURE: fiber scheduler of the CPU emulator (see emu_shim.h).
#include "emu_shim.h"
#include <dlfcn.h>
#include <map>
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace emu {
Idx thread_idx{ 0, 0, 0 }, block_idx{ 0, 0, 0 }, block_dim{ 1, 1, 1 }, grid_dim{ 1, 1, 1 };
alignas(16) static double dyn_smem_buf[232448 / 8 + 16];
double* const dyn_smem = dyn_smem_buf;

namespace {
constexpr size_t kStack = 512 * 1024;
struct Fiber
{
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = true;
  void* site = nullptr; // return address of the barrier / collective the fiber waits in
  int kind = 0;         // 1 block barrier, 2 warp barrier, 3 yield
};
std::vector<Fiber> fibers;
ucontext_t sched;
int cur = -1, nthreads = 0;
const std::function<void()>* body = nullptr;
struct Bar
{
  int count = 0;
  unsigned gen = 0;
};
Bar cta_bar;
Bar warp_bar[64];
unsigned long long progress = 0; // bumped whenever a barrier completes
WarpBuf warp_bufs[64];
unsigned long long ticks = 0;
int or_acc = 0;

void
trampoline()
{
  (*body)();
  fibers[(size_t)cur].done = true;
  swapcontext(&fibers[(size_t)cur].ctx, &sched);
}
void
wait_on(Bar& b, int n)
{
  const unsigned g = b.gen;
  if (++b.count == n) {
    b.count = 0;
    ++b.gen;
    ++progress;
  } else {
    while (b.gen == g) yield();
  }
}
} // namespace

unsigned long long
globaltimer()
{
  return ticks += 1000;
}
void
yield()
{
  swapcontext(&fibers[(size_t)cur].ctx, &sched);
}
void
block_barrier()
{
  fibers[(size_t)cur].site = __builtin_return_address(0);
  fibers[(size_t)cur].kind = 1;
  wait_on(cta_bar, nthreads);
}
void
warp_barrier()
{
  const int w = (int)(thread_idx.x >> 5);
  const int lanes = std::min(32, nthreads - 32 * w);
  fibers[(size_t)cur].site = __builtin_return_address(0);
  fibers[(size_t)cur].kind = 2;
  wait_on(warp_bar[w], lanes);
}
WarpBuf&
warp_buf()
{
  return warp_bufs[thread_idx.x >> 5];
}

void
run_grid(int grid, int block, const std::function<void()>& fn)
{
  if (block > 2048 || block <= 0) std::abort();
  if ((int)fibers.size() < block) fibers.resize((size_t)block);
  body = &fn;
  nthreads = block;
  block_dim = Idx{ (unsigned)block, 1, 1 };
  grid_dim = Idx{ (unsigned)grid, 1, 1 };
  for (int b = 0; b < grid; ++b) {
    block_idx = Idx{ (unsigned)b, 0, 0 };
    cta_bar = Bar();
    for (auto& wb : warp_bar) wb = Bar();
    or_acc = 0;
    for (int t = 0; t < block; ++t) {
      Fiber& f = fibers[(size_t)t];
      if (!f.stack) f.stack = static_cast<char*>(std::malloc(kStack));
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = kStack;
      f.ctx.uc_link = &sched;
      f.done = false;
      makecontext(&f.ctx, trampoline, 0);
    }
    int live = block;
    unsigned long long rounds = 0, last_progress = progress, idle_rounds = 0;
    auto stall_report = [&](const char* why) {
      std::fprintf(stderr, "emu: CTA %d stalled (%s); where its threads wait (library offset: threads):\n", b, why);
      std::map<std::pair<int, void*>, int> hist;
      for (int t = 0; t < block; ++t) {
        if (!fibers[(size_t)t].done) hist[{ fibers[(size_t)t].kind, fibers[(size_t)t].site }]++;
      }
      for (auto& kv : hist) {
        Dl_info di{};
        dladdr(kv.first.second, &di);
        std::fprintf(stderr, "  %s at +0x%zx : %d\n", kv.first.first == 1 ? "block barrier" : kv.first.first == 2 ? "warp collective" : "yield",
                     (size_t)((char*)kv.first.second - (char*)di.dli_fbase), kv.second);
      }
      std::fprintf(stderr, "  (%d thread(s) returned)  resolve with: addr2line -f -C -e tests/emu/libpqp_emu.so 0x<offset>\n", block - live);
      std::abort();
    };
    while (live > 0) {
      live = 0;
      for (int t = 0; t < block; ++t) {
        Fiber& f = fibers[(size_t)t];
        if (f.done) continue;
        cur = t;
        thread_idx = Idx{ (unsigned)t, 0, 0 };
        swapcontext(&sched, &f.ctx);
        if (!f.done) ++live;
      }
      if (live > 0 && cta_bar.count > 0 && cta_bar.count + (block - live) == block && live < block)
        stall_report("threads returned while others wait at a block barrier");
      // every fiber runs to its next barrier in one round: a round without a completed barrier means every
      // live thread is waiting -> the barriers can never complete (unpaired / divergent barriers)
      if (progress == last_progress) {
        if (++idle_rounds > 64) stall_report("no barrier completes: divergent or unpaired barriers");
      } else {
        idle_rounds = 0;
        last_progress = progress;
      }
      ++rounds;
    }
  }
  body = nullptr;
}
} // namespace emu

int
__syncthreads_or(int p)
{
  // all threads contribute, then all read, then the accumulator is cleared
  static int acc = 0, result = 0;
  if (p) acc = 1;
  emu::block_barrier();
  result = acc;
  emu::block_barrier();
  const int r = result;
  acc = 0;
  emu::block_barrier();
  return r;
}

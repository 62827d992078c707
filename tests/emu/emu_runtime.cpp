// TEST INFRASTRUCTURE: fiber scheduler of the CPU emulator (see emu_shim.h).
#include "emu_shim.h"
#include <dlfcn.h>
#include <map>
// EXPERIMENTAL race detection (`make race`): built with -fsanitize=thread every emulated CUDA thread is a TSan
// fiber, switches carry NO synchronisation, and the only happens-before edges are the block / warp barriers (release
// on arrival, acquire on departure, one sync object per barrier generation in flight). Conflicting accesses of two
// CUDA threads that no barrier separates are then reported with both source lines. It catches such conflicts in
// small kernels (see the self-test in emu_race_main.cpp) but is NOT reliable on the full solver: TSan keeps only four
// accesses per 8-byte word, and the round-1 `active_set_change` race, re-introduced on purpose, went unreported.
// A clean run therefore proves nothing; a report is worth reading.
#if defined(__SANITIZE_THREAD__)
#include <sanitizer/tsan_interface.h>
#define EMU_TSAN 1
#define EMU_NO_TSAN __attribute__((no_sanitize("thread")))
#else
#define EMU_NO_TSAN
#endif
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
  void* tsan = nullptr; // TSan fiber context (race-detection build)
  bool started = false;
};
void* sched_tsan = nullptr;
std::vector<Fiber> fibers;
ucontext_t sched;
int cur = -1, nthreads = 0;
const std::function<void()>* body = nullptr;
struct Bar
{
  int count = 0;
  unsigned gen = 0;
  char token[4] = {}; // TSan sync objects, one per barrier generation in flight (a fast thread arrives at barrier
                      // g + 1 before a slow one has left g: a single object would leak the later release into it)
};
Bar cta_bar;
Bar warp_bar[64];
unsigned long long progress = 0; // bumped whenever a barrier completes
WarpBuf warp_bufs[64];
unsigned long long ticks = 0;
unsigned long long barriers0 = 0; // block barriers passed by thread 0 (EMU_CLOCK=barriers)
int or_acc = 0;

EMU_NO_TSAN void
to_scheduler()
{
#ifdef EMU_TSAN
  __tsan_switch_to_fiber(sched_tsan, __tsan_switch_to_fiber_no_sync);
#endif
  swapcontext(&fibers[(size_t)cur].ctx, &sched);
}
void
trampoline()
{
  (*body)();
  fibers[(size_t)cur].done = true;
#ifdef EMU_TSAN
  __tsan_switch_to_fiber(sched_tsan, 0); // kernel end: everything the thread did happens before the host continues
  swapcontext(&fibers[(size_t)cur].ctx, &sched);
#else
  to_scheduler();
#endif
}
EMU_NO_TSAN void
wait_on(Bar& b, int n)
{
  const unsigned g = b.gen;
#ifdef EMU_TSAN
  __tsan_release(&b.token[g & 3]); // everything this thread did before the barrier ...
#endif
  if (++b.count == n) {
    b.count = 0;
    ++b.gen;
    ++progress;
  } else {
    while (b.gen == g) to_scheduler();
  }
#ifdef EMU_TSAN
  __tsan_acquire(&b.token[g & 3]); // ... happens before everything any thread does after it
#endif
}
} // namespace

EMU_NO_TSAN unsigned long long
globaltimer()
{
  // EMU_CLOCK=barriers: the clock counts the block barriers thread 0 has passed, so PQP_PROFILE=1 reports the number of
  // barrier intervals per phase (the kernels are bound by dependent latency per interval, not by work)
  static const bool count_barriers = [] { const char* e = std::getenv("EMU_CLOCK"); return e && !std::strcmp(e, "barriers"); }();
  if (count_barriers) return barriers0;
  return ticks += 1000;
}
EMU_NO_TSAN void
yield()
{
  to_scheduler();
}
EMU_NO_TSAN void
block_barrier()
{
  if (cur == 0) ++barriers0;
  fibers[(size_t)cur].site = __builtin_return_address(0);
  fibers[(size_t)cur].kind = 1;
  wait_on(cta_bar, nthreads);
}
EMU_NO_TSAN void
warp_barrier()
{
  const int w = (int)(thread_idx.x >> 5);
  const int lanes = std::min(32, nthreads - 32 * w);
  fibers[(size_t)cur].site = __builtin_return_address(0);
  fibers[(size_t)cur].kind = 2;
  wait_on(warp_bar[w], lanes);
}
EMU_NO_TSAN WarpBuf&
warp_buf()
{
  return warp_bufs[thread_idx.x >> 5];
}

EMU_NO_TSAN void
run_grid(int grid, int block, const std::function<void()>& fn)
{
#ifdef EMU_TSAN
  sched_tsan = __tsan_get_current_fiber();
#endif
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
      f.started = false;
      makecontext(&f.ctx, trampoline, 0);
#ifdef EMU_TSAN
      if (f.tsan) __tsan_destroy_fiber(f.tsan);
      f.tsan = __tsan_create_fiber(0);
#endif
    }
    int live = block;
    static const int order_mode = [] { const char* e = std::getenv("EMU_ORDER"); return !e ? 0 : !std::strcmp(e, "reverse") ? 1 : !std::strcmp(e, "stride") ? 2 : 0; }();
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
      for (int tt = 0; tt < block; ++tt) {
        // EMU_ORDER=reverse|stride runs the threads of a CTA in another order between barriers: a result that changes
        // with the order is a data race in the kernel (or a barrier the kernel forgot)
        const int t = order_mode == 1 ? block - 1 - tt : order_mode == 2 ? (int)(((long long)tt * 37 + 11) % block) : tt;
        Fiber& f = fibers[(size_t)t];
        if (f.done) continue;
        cur = t;
        thread_idx = Idx{ (unsigned)t, 0, 0 };
#ifdef EMU_TSAN
        // the launch orders the host's writes before the thread's first instruction; later switches carry no
        // synchronisation (only barriers order CUDA threads)
        __tsan_switch_to_fiber(f.tsan, f.started ? __tsan_switch_to_fiber_no_sync : 0);
        f.started = true;
#endif
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

EMU_NO_TSAN int
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

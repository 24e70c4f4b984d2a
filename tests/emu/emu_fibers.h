// TEST INFRASTRUCTURE ONLY -- the execution model shared by the two host emulations of the CUDA
// sources (tests/emu/harness_prelude.h: single kernels cut out of a .cu file; tests/emu/cpu_build/:
// the whole library).  Nothing here is shipped or used by the product.
//
// One OS thread per WARP of a thread block; the (up to) 32 lanes of a warp are cooperative fibers of
// that thread.  A lane runs until it reaches a barrier, then hands the thread to the next lane:
//   * warp barrier (__syncwarp, and the two barriers inside every emulated shuffle / mma.sync): the
//     lanes take turns until all 32 have arrived -- a round of user-space context switches;
//   * block barrier (__syncthreads): the lanes of each warp gather the same way, the last one meets
//     the other warps at a pthread barrier.
// Blocks of a grid run one after the other.  The first version of these harnesses ran one std::thread
// per CUDA thread with pthread barriers everywhere and spent most of its time in futex calls (the CPU
// suite took 10 minutes; 20x less inside the kernels now).
// Context switch: six callee-saved registers + the stack pointer on x86-64 (no system call), ucontext
// elsewhere.  As on the GPU, every lane of a warp must reach the same warp barriers and every thread of
// a block the same block barriers; a lane left alone at a barrier aborts with a message instead of hanging.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <functional>
#include <thread>
#include <vector>

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };

// ---------------------------------------------------------------- contexts
#if defined(__x86_64__)
struct EmuCtx { void* sp = nullptr; };
__attribute__((naked, noinline, used)) static void emu_switch_sp(void** /*save_sp: rdi*/, void* /*load_sp: rsi*/) {
  asm volatile(
      "pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
      "movq %rsp, (%rdi)\n\t"
      "movq %rsi, %rsp\n\t"
      "popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\t"
      "ret\n\t");
}
inline void emu_ctx_make(EmuCtx& c, char* stack, size_t size, void (*fn)()) {
  void** sp = (void**)(stack + size);       // 16-byte aligned top
  *--sp = nullptr;                          // keeps rsp = 8 (mod 16) at the entry of fn, as after a call
  *--sp = (void*)fn;                        // target of the first `ret`
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  c.sp = sp;
}
inline void emu_ctx_switch(EmuCtx& from, EmuCtx& to) { emu_switch_sp(&from.sp, to.sp); }
#else
struct EmuCtx { ucontext_t uc; };
inline void emu_ctx_make(EmuCtx& c, char* stack, size_t size, void (*fn)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = size;
  c.uc.uc_link = nullptr;
  makecontext(&c.uc, fn, 0);
}
inline void emu_ctx_switch(EmuCtx& from, EmuCtx& to) { swapcontext(&from.uc, &to.uc); }
#endif

// ---------------------------------------------------------------- warps and lanes
struct EmuFiber {
  EmuCtx ctx;
  EmuIdx tid;
  unsigned ltid = 0;                        // linear thread id inside the block
  bool done = false;
};
struct EmuWarp {                            // one OS thread
  EmuFiber fib[32];
  int n = 0, cur = 0, live = 0;
  EmuCtx main;
  unsigned arrived = 0, gen = 0;            // warp barrier
  unsigned cta_arrived = 0, cta_gen = 0;    // warp-local part of the block barrier
};
inline thread_local EmuWarp* emu_w = nullptr;
inline thread_local EmuIdx blockIdx;
#define threadIdx (emu_w->fib[emu_w->cur].tid)
#define emu_ltid (emu_w->fib[emu_w->cur].ltid)
inline unsigned emu_nthreads = 0, emu_nwarps = 0;
inline pthread_barrier_t emu_bar;           // across the warps (OS threads) of a block
inline const std::function<void()>* emu_body = nullptr;
constexpr size_t EMU_STACK = 256 * 1024;    // per lane; pages are touched lazily
inline char* emu_stacks = nullptr;          // [1024][EMU_STACK]

// hand the OS thread to the next unfinished lane of this warp
inline void emu_yield() {
  EmuWarp* w = emu_w;
  const int from = w->cur;
  int to = from;
  do { to = (to + 1) % w->n; } while (w->fib[to].done && to != from);
  if (to == from) {
    fprintf(stderr, "emulation: thread %u waits at a barrier that the rest of its warp never reaches\n", w->fib[from].ltid);
    abort();
  }
  w->cur = to;
  emu_ctx_switch(w->fib[from].ctx, w->fib[to].ctx);
}
// the same for a lane that polls something another warp will change (an mbarrier): false if no other lane
// of this warp can run, and the caller should give up the OS thread instead
inline bool emu_yield_if_possible() {
  EmuWarp* w = emu_w;
  for (int l = 0; l < w->n; ++l)
    if (l != w->cur && !w->fib[l].done) {
      emu_yield();
      return true;
    }
  return false;
}
inline void emu_warp_barrier() {
  EmuWarp* w = emu_w;
  const unsigned g = w->gen;
  if (++w->arrived == (unsigned)w->n) {
    w->arrived = 0;
    ++w->gen;
    return;
  }
  while (w->gen == g) emu_yield();
}
inline void emu_barrier() {
  EmuWarp* w = emu_w;
  const unsigned g = w->cta_gen;
  if (++w->cta_arrived == (unsigned)w->n) {   // last lane of this warp: meet the other warps
    w->cta_arrived = 0;
    if (emu_nwarps > 1) pthread_barrier_wait(&emu_bar);
    ++w->cta_gen;
    return;
  }
  while (w->cta_gen == g) emu_yield();
}
inline void emu_fiber_main() {
  (*emu_body)();
  EmuWarp* w = emu_w;
  EmuFiber& me = w->fib[w->cur];
  me.done = true;
  if (--w->live == 0) emu_ctx_switch(me.ctx, w->main);
  int to = w->cur;
  do { to = (to + 1) % w->n; } while (w->fib[to].done);
  w->cur = to;
  emu_ctx_switch(me.ctx, w->fib[to].ctx);
  abort();                                  // a finished lane is never resumed
}

// all blocks of one launch, as seen by warp `wid`
inline void emu_run_warp(unsigned wid, EmuIdx grid, EmuIdx block, std::vector<double>* smem) {
  EmuWarp* w = new EmuWarp;
  emu_w = w;
  w->n = (int)(emu_nthreads - 32 * wid < 32u ? emu_nthreads - 32 * wid : 32u);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        if (wid == 0)
          for (auto& v : *smem) v = NAN;    // fresh block: stale shared memory cannot help
        if (emu_nwarps > 1) pthread_barrier_wait(&emu_bar);
        w->arrived = w->cta_arrived = 0;
        w->live = w->n;
        for (int l = 0; l < w->n; ++l) {
          EmuFiber& f = w->fib[l];
          const unsigned t = 32 * wid + (unsigned)l;
          f.ltid = t;
          f.tid.x = t % block.x;
          f.tid.y = (t / block.x) % block.y;
          f.tid.z = t / (block.x * block.y);
          f.done = false;
          emu_ctx_make(f.ctx, emu_stacks + (size_t)t * EMU_STACK, EMU_STACK, emu_fiber_main);
        }
        w->cur = 0;
        emu_ctx_switch(w->main, w->fib[0].ctx);   // back here when the last lane has finished
        if (emu_nwarps > 1) pthread_barrier_wait(&emu_bar);
      }
  emu_w = nullptr;
  delete w;
}

// every block of a grid (extents in grid.x/y/z, each >= 1) of block.x * block.y * block.z threads;
// `smem` is the dynamic shared memory of the running block (poisoned with NaN at every block start)
inline void emu_run_grid(EmuIdx grid, EmuIdx block, std::vector<double>* smem, const std::function<void()>& fn) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads == 0 || nthreads > 1024) { fprintf(stderr, "emulation: %u threads per block\n", nthreads); abort(); }
  if (!emu_stacks) {
    emu_stacks = (char*)mmap(nullptr, 1024 * EMU_STACK, PROT_READ | PROT_WRITE,
                             MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (emu_stacks == (char*)MAP_FAILED) { perror("emulation: mmap of the lane stacks"); abort(); }
    for (size_t t = 0; t < 1024; ++t)       // guard page below every stack: an overflow faults instead of
      mprotect(emu_stacks + t * EMU_STACK, 4096, PROT_NONE);   // scribbling over the neighbouring lane
  }
  emu_nthreads = nthreads;
  emu_nwarps = (nthreads + 31) / 32;
  emu_body = &fn;
  if (emu_nwarps == 1) {
    emu_run_warp(0, grid, block, smem);
  } else {
    pthread_barrier_init(&emu_bar, nullptr, emu_nwarps);
    std::vector<std::thread> pool;
    pool.reserve(emu_nwarps);
    for (unsigned wid = 0; wid < emu_nwarps; ++wid)
      pool.emplace_back([=]() { emu_run_warp(wid, grid, block, smem); });
    for (auto& th : pool) th.join();
    pthread_barrier_destroy(&emu_bar);
  }
  emu_body = nullptr;
}

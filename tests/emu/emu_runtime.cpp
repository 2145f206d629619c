// Test-only SIMT emulator runtime (see emu_runtime.h).
#include "emu_runtime.h"

#include <sys/mman.h>

namespace emu {

Globals g;
Fiber* cur = nullptr;

static const size_t kStack = 256 * 1024;
static std::vector<char*> stack_pool;
static int order_mode = -1;  // 0 fwd, 1 reverse, 2 random
static uint32_t lcg = 12345u;

static char* get_stack(size_t i) {
  while (stack_pool.size() <= i) {
    void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("emu mmap"); abort(); }
    stack_pool.push_back((char*)p);
  }
  return stack_pool[i];
}

#if EMU_FAST_SWITCH
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
static inline void to_sched() { emu_switch(&cur->sp, g.sched_sp); }
static inline void to_fiber(Fiber& f) { emu_switch(&g.sched_sp, f.sp); }
#else
static inline void to_sched() { swapcontext(&cur->ctx, &g.sched); }
static inline void to_fiber(Fiber& f) { swapcontext(&g.sched, &f.ctx); }
#endif

static void trampoline() {
  (*g.body)();
  cur->done = true;
  g.alive--;
  g.progress++;
  // a thread that exits no longer takes part in block barriers (CUDA semantics for exited threads)
  if (g.alive > 0 && g.arrived == g.alive) { g.arrived = 0; g.gen++; }
  to_sched();
  abort();      // a finished fiber is never resumed
}

void yield() { to_sched(); }

void syncthreads() {
  int gen = g.gen;
  if (++g.arrived == g.alive) {
    g.arrived = 0;
    g.gen++;
    g.progress++;
  } else {
    while (g.gen == gen) yield();
  }
}

void warp_barrier() {
  Warp& w = g.warps[cur->warp];
  int gen = w.gen;
  if (++w.arrived == w.nlanes) {
    w.arrived = 0;
    w.gen++;
    g.progress++;
  } else {
    while (w.gen == gen) yield();
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  if (order_mode < 0) {
    const char* e = getenv("EMU_ORDER");
    order_mode = (e && !strcmp(e, "reverse")) ? 1 : (e && !strcmp(e, "random")) ? 2 : 0;
  }
  int nt = (int)(block.x * block.y * block.z);
  if (nt <= 0 || nt > 1024) { fprintf(stderr, "emu: bad block size %d\n", nt); abort(); }
  g.bdim = block;
  g.gdim = grid;
  g.nthreads = nt;
  g.body = &body;
  g.dynsmem.assign(smem + 64, 0);
  g.fibers.resize(nt);
  int nw = (nt + 31) / 32;
  g.warps.resize(nw);
  std::vector<int> order(nt);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g.block = uint3{bx, by, bz};
        g.alive = nt;
        g.arrived = 0;
        g.gen = 0;
        for (int w = 0; w < nw; ++w) {
          g.warps[w].arrived = 0;
          g.warps[w].gen = 0;
          g.warps[w].nlanes = std::min(32, nt - 32 * w);
        }
        for (int i = 0; i < nt; ++i) {
          Fiber& f = g.fibers[i];
          f.lin = i;
          f.tid = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
          f.lane = i % 32;
          f.warp = i / 32;
          f.done = false;
          f.stack = get_stack(i);
#if EMU_FAST_SWITCH
          {   // initial frame: six callee-saved registers (don't care) + return address = trampoline; at its entry rsp % 16 == 8
            uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + kStack) & ~uintptr_t(15);
            void** sp = reinterpret_cast<void**>(top - 8);        // slot that a caller's `call` would have filled (alignment)
            *--sp = reinterpret_cast<void*>(trampoline);
            for (int k = 0; k < 6; ++k) *--sp = nullptr;
            f.sp = sp;
          }
#else
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = &g.sched;
          makecontext(&f.ctx, (void (*)())trampoline, 0);
#endif
          order[i] = (order_mode == 1) ? nt - 1 - i : i;
        }
        long idle_passes = 0;
        while (g.alive > 0) {
          if (order_mode == 2) {
            for (int i = nt - 1; i > 0; --i) {
              lcg = lcg * 1664525u + 1013904223u;
              std::swap(order[i], order[(lcg >> 8) % (i + 1)]);
            }
          }
          long before = g.progress;
          for (int k = 0; k < nt; ++k) {
            Fiber& f = g.fibers[order[k]];
            if (f.done) continue;
            cur = &f;
            to_fiber(f);
          }
          if (g.progress == before) {
            if (++idle_passes > 4) {
              fprintf(stderr, "emu: DEADLOCK in block (%u,%u,%u): %d threads alive, %d at barrier\n", bx, by, bz, g.alive, g.arrived);
              abort();
            }
          } else {
            idle_passes = 0;
          }
        }
      }
  cur = nullptr;
}

}  // namespace emu

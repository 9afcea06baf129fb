// tests/emul/simt/simt_emu.cpp — TEST HARNESS ONLY.  Fiber scheduler of the SIMT emulator (see cuda_runtime.h).
//
// One CUDA thread = one fiber with its own stack; a block's fibers are multiplexed on one OS thread by a
// run queue.  A fiber leaves the CPU only at a rendezvous (warp collective, __syncthreads, grid.sync) or in an
// explicit emu::yield() of a spin loop, so the interleaving is deterministic for a given launch.  A normal launch
// runs its blocks one after the other on the calling thread; a cooperative launch gives every block its own OS
// thread and grid.sync() is a pthread barrier entered by the last fiber of each block.  With YT_EMU_SCHED_SEED=n the
// next fiber is drawn at random from the run queue instead of in FIFO order.
#include <pthread.h>
#include <stdio.h>
#include <sys/mman.h>

#include <deque>
#include <thread>
#include <vector>

#include "cuda_runtime.h"

extern "C" void emu_switch(void **save_sp, void *next_sp);
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

namespace emu {

thread_local Coords *t_coords = nullptr;

namespace {
constexpr size_t STACK_BYTES = 96 * 1024;
enum State { RUNNABLE, WAITING, DONE };

struct Block;
struct Fiber {
  void *sp = nullptr;
  char *stack = nullptr;
  Coords co;
  State state = DONE;
  unsigned warp = 0;
};
struct Warp {
  uint64_t vals[32];
  uint64_t snap[2][32];
  uint32_t snap_live[2] = {0, 0};
  uint32_t arrived = 0, live = 0;
  unsigned gen = 0;
  std::vector<Fiber *> waiters;
};
struct Grid {
  pthread_barrier_t bar;
};
struct Block {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::deque<Fiber *> runq;
  std::vector<Fiber *> sync_waiters;
  unsigned n_live = 0, sync_arrived = 0;
  bool sync_is_grid = false;
  Grid *grid = nullptr;
  void *main_sp = nullptr;
  Fiber *current = nullptr;
  const std::function<void()> *body = nullptr;
  void *dyn = nullptr;
};
thread_local Block *t_blk = nullptr;

struct StackPool {  // stacks are recycled between launches of the same OS thread
  std::vector<char *> free_list;
  char *get() {
    if (!free_list.empty()) { char *s = free_list.back(); free_list.pop_back(); return s; }
    void *p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("simt_emu: mmap"); abort(); }
    return (char *)p;
  }
  void put(char *s) { free_list.push_back(s); }
  ~StackPool() { for (char *s : free_list) munmap(s, STACK_BYTES); }
};
thread_local StackPool t_stacks;

void to_scheduler(Fiber *f) { emu_switch(&f->sp, t_blk->main_sp); }
void block_here(Fiber *f) { f->state = WAITING; to_scheduler(f); }
void wake(Block *b, Fiber *f) { f->state = RUNNABLE; b->runq.push_back(f); }

void complete_warp(Block *b, Warp &w) {  // every live lane has contributed
  const unsigned g = w.gen & 1u;
  memcpy(w.snap[g], w.vals, sizeof(w.vals));
  w.snap_live[g] = w.live;
  w.gen++;
  w.arrived = 0;
  for (Fiber *f : w.waiters) wake(b, f);
  w.waiters.clear();
}
void complete_sync(Block *b) {
  if (b->sync_is_grid && b->grid) pthread_barrier_wait(&b->grid->bar);  // the other fibers of this block all wait
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
  b->sync_arrived = 0;
  b->sync_is_grid = false;
  for (Fiber *f : b->sync_waiters) wake(b, f);
  b->sync_waiters.clear();
}
void rendezvous(bool grid) {
  Block *b = t_blk;
  Fiber *f = b->current;
  if (grid) b->sync_is_grid = true;
  if (++b->sync_arrived == b->n_live) { complete_sync(b); return; }
  b->sync_waiters.push_back(f);
  block_here(f);
}

void fiber_main() {
  Block *b = t_blk;
  Fiber *f = b->current;
  (*b->body)();
  // exit: the lane / thread no longer takes part in collectives
  f->state = DONE;
  Warp &w = b->warps[f->warp];
  w.live &= ~(1u << f->co.lane);
  b->n_live--;
  if (w.live && w.arrived == w.live && !w.waiters.empty()) complete_warp(b, w);
  if (b->n_live && b->sync_arrived == b->n_live && !b->sync_waiters.empty()) complete_sync(b);
  to_scheduler(f);
  abort();  // a finished fiber is never resumed
}

void run_block(unsigned bid, unsigned grid, unsigned threads, size_t smem, const std::function<void()> &body, Grid *g) {
  Block blk;
  blk.grid = g;
  blk.body = &body;
  blk.n_live = threads;
  blk.fibers.resize(threads);
  blk.warps.resize((threads + 31) / 32);
  if (smem) { blk.dyn = aligned_alloc(128, (smem + 127) / 128 * 128); memset(blk.dyn, 0xcd, smem); }
  for (unsigned t = 0; t < threads; t++) {
    Fiber &f = blk.fibers[t];
    f.co.tid = dim3(t);
    f.co.bid = dim3(bid);
    f.co.bdim = dim3(threads);
    f.co.gdim = dim3(grid);
    f.co.lane = t & 31u;
    f.warp = t >> 5;
    f.state = RUNNABLE;
    f.stack = t_stacks.get();
    blk.warps[f.warp].live |= 1u << f.co.lane;
    // initial frame: six callee-saved registers, then the "return address" emu_switch jumps to
    void **top = (void **)(f.stack + STACK_BYTES);
    void **sp = top - 8;
    for (int i = 0; i < 6; i++) sp[i] = nullptr;
    sp[6] = (void *)&fiber_main;
    sp[7] = nullptr;
    f.sp = sp;
    blk.runq.push_back(&f);
  }
  const char *ss = getenv("YT_EMU_SCHED_SEED");
  const uint64_t sched_seed = ss ? strtoull(ss, nullptr, 10) : 0;
  uint64_t rng = (sched_seed * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)bid + 1) * 0xD1B54A32D192ED03ull;
  if (!rng) rng = 1;
  Block *outer_blk = t_blk;
  Coords *outer_co = t_coords;
  t_blk = &blk;
  unsigned done = 0;
  while (done < threads) {
    if (blk.runq.empty()) {
      fprintf(stderr, "simt_emu: DEADLOCK in block %u: %u of %u threads finished, %u at a block/grid barrier\n", bid, done,
              threads, blk.sync_arrived);
      abort();
    }
    if (sched_seed) {  // YT_EMU_SCHED_SEED: any runnable fiber may go next (order-dependence / missing-barrier hunt)
      rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
      std::swap(blk.runq.front(), blk.runq[rng % blk.runq.size()]);
    }
    Fiber *f = blk.runq.front();
    blk.runq.pop_front();
    blk.current = f;
    t_coords = &f->co;
    emu_switch(&blk.main_sp, f->sp);
    if (f->state == DONE && f->stack) { t_stacks.put(f->stack); f->stack = nullptr; done++; }
  }
  free(blk.dyn);
  t_blk = outer_blk;
  t_coords = outer_co;
}
}  // namespace

void yield() {
  Block *b = t_blk;
  Fiber *f = b->current;
  b->runq.push_back(f);
  to_scheduler(f);
}

const uint64_t *warp_gather(uint64_t v, uint32_t *live) {
  Block *b = t_blk;
  Fiber *f = b->current;
  Warp &w = b->warps[f->warp];
  const unsigned g = w.gen & 1u;
  w.vals[f->co.lane] = v;
  w.arrived |= 1u << f->co.lane;
  if (w.arrived == w.live) complete_warp(b, w);
  else { w.waiters.push_back(f); block_here(f); }
  *live = w.snap_live[g];
  return w.snap[g];
}
void block_sync() { rendezvous(false); }
void grid_sync() { rendezvous(true); }
void *dyn_smem() { return t_blk->dyn; }

unsigned n_sm() {
  const char *e = getenv("YT_EMU_SMS");
  const int n = e ? atoi(e) : 2;
  return (unsigned)(n < 1 ? 1 : n);
}

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body) {
  for (unsigned b = 0; b < grid; b++) run_block(b, grid, block, smem, body, nullptr);
}

void launch_cooperative(unsigned grid, unsigned block, size_t smem, const std::function<void()> &body) {
  Grid g;
  pthread_barrier_init(&g.bar, nullptr, grid);
  std::vector<std::thread> th;
  for (unsigned b = 0; b < grid; b++) th.emplace_back([&, b] { run_block(b, grid, block, smem, body, &g); });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&g.bar);
}

}  // namespace emu

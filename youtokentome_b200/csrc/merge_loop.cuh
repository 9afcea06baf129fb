// merge_loop.cuh — phase 4 of training: the persistent merge loop (included by train.cu).
//
// Replaces the reference's main loop + worker_doing_merge + PriorityQueue
// (bpe.cpp:1121-1282, 601-811, 149-314).  One cooperative kernel runs every merge iteration on
// the device; per iteration:
//   1. exact arg-max over the pair table under MergeCandidate::operator< (bpe.cpp:110-126)
//   2. grid barrier; every block reduces the per-block winners redundantly (no second barrier)
//   3. apply x y -> z: the packed words are organised in TILES (windows of `tile_tok` token
//      slots, cut at word boundaries).  RESIDENT mode: block b keeps tile b in shared memory for
//      the whole launch (the working set of deduplicated words fits the 148 x 227 KB of SMEM:
//      no HBM/L2 token traffic inside the loop at all).  STREAMING mode (token buffer larger than
//      the SMEM of the chip): tiles are staged into shared memory each iteration, coalesced,
//      and modified words are written through to HBM.
//      Inside a tile a lane tests one word (has_pair); words that hold the pair are rewritten by
//      the whole warp in parallel (ballot/popc compaction, run-parity rule for x == y) and their
//      pair multiset is re-counted (-old, +new) with one table update per lane.
//   4. grid barrier.
#pragma once

constexpr int SWEEP_UNROLL = 8;  // arg-max sweep: table counts in flight per thread

struct LoopArgs {
  uint32_t *tok;
  const uint32_t *off;
  const uint64_t *freq;
  uint64_t n_words;
  const uint2 *tile_desc;      // n_tiles + 1 entries (first word, its token offset); last = (n_words, n_slots)
  uint32_t stream_tok_cap;     // STREAMING: token / word capacity of ONE of the two pipeline stages
  uint32_t stream_word_cap;
  uint32_t n_stage;            // STREAMING: pipeline depth
  uint32_t dbg;                // diagnostics (env YTTM_DBG): 1 = consumers skip the scan, 2 = scalar scan, 8 = per-block apply timers
  uint4 *defer;                // STREAMING: per-block lists of words to rewrite after the tile scan
  uint32_t defer_cap;          // entries per block
  uint32_t n_tiles;
  uint32_t resident;           // 1: block b owns tile b and keeps it in shared memory
  uint32_t smem_tok_cap;       // token capacity of the shared tile buffer
  uint32_t smem_word_cap;      // word capacity of the shared tile buffer
  PairTab tab;
  YtLoopCtl *ctl;
  unsigned long long *blockbest;  // 4 per block: count, prio, slot, (pad)
  uint32_t *rules;                // 3 per merge
  unsigned long long *rfreq;
  uint32_t first_new_id;          // id of merge number 0
  uint32_t max_total;             // stop when ctl->n_done reaches this
  uint32_t max_iters;             // iterations allowed in this launch
  uint32_t key_limit;             // leave for a rebuild above this table occupancy
};

struct Best { unsigned long long c, prio, slot; };
__device__ __forceinline__ bool better(const Best &a, const Best &b) {  // a beats b
  return a.c > b.c || (a.c == b.c && a.prio > b.prio);
}
#ifndef YT_SIMT_EMU
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#else   // tests/emul/simt: the CPU clock
inline unsigned long long gtimer() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
#endif
__device__ __forceinline__ Best warp_best(Best v) {
  for (int o = 16; o > 0; o >>= 1) {
    Best w;
    w.c = __shfl_xor_sync(0xffffffffu, v.c, o);
    w.prio = __shfl_xor_sync(0xffffffffu, v.prio, o);
    w.slot = __shfl_xor_sync(0xffffffffu, v.slot, o);
    if (better(w, v)) v = w;
  }
  return v;
}

struct MergeOp { uint32_t x, y, z; unsigned long long key; };

// Per-warp queue of pending table updates in shared memory: the updates of several rewritten
// words are issued together, one per lane, so their L2 round trips overlap.
constexpr int UQ_CAP = 64;
constexpr int MAX_STAGES = 16;
constexpr int CLAIM_WORDS = 1024;  // claim bitmap: up to 32768 words per shared-memory tile
struct WarpQueue {
  unsigned long long *key;  // UQ_CAP entries of this warp
  long long *delta;
  uint32_t n;               // warp-uniform
};
template <class Tab>
__device__ __forceinline__ void uq_drain(WarpQueue &q, unsigned lane, const Tab &tab) {
  __syncwarp();
  for (uint32_t i = lane; i < q.n; i += 32) {
    const long long d = q.delta[i];
    if (d > 0) pair_add<true>(tab, q.key[i], d);  // additions are mostly pairs with the new token: CAS first
    else pair_add<false>(tab, q.key[i], d);
  }
  __syncwarp();
  q.n = 0;
}
// every lane may contribute one update (has == true); direct == true bypasses the queue
template <class Tab>
__device__ __forceinline__ void uq_push(WarpQueue &q, bool direct, bool has, unsigned long long key, long long delta,
                                        unsigned lane, const Tab &tab) {
  if (direct) { if (has) pair_add(tab, key, delta); return; }
  const unsigned m = __ballot_sync(0xffffffffu, has);
  if (has) {
    const uint32_t i = q.n + __popc(m & ((1u << lane) - 1u));
    q.key[i] = key;
    q.delta[i] = delta;
  }
  q.n += __popc(m);
}

// Run structure of <= 32 live tokens held one per lane (t == DEAD beyond the n live ones): every
// run start owns floor(L/2) self pairs and one cross pair to the next run (for_each_pair()).
struct RunInfo { bool start; uint32_t nxt, L, b; unsigned starts; };
__device__ __forceinline__ RunInfo warp_runs(uint32_t t, uint32_t n, unsigned lane) {
  RunInfo r;
  const bool valid = lane < n;
  uint32_t prev = __shfl_up_sync(0xffffffffu, t, 1);
  r.start = valid && (lane == 0 || prev != t);
  r.starts = __ballot_sync(0xffffffffu, r.start);
  const unsigned above = lane == 31 ? 0u : r.starts & ~((2u << lane) - 1u);
  r.nxt = above ? (uint32_t)__ffs(above) - 1u : n;
  r.b = __shfl_sync(0xffffffffu, t, r.nxt & 31);
  r.L = r.nxt - lane;
  return r;
}

// One warp rewrites one word that holds (x,y): st = its `cap` token slots (shared or global
// memory), gt = optional write-through copy in global memory.  Table updates: only the pairs
// whose run is touched by a merge are sent (-old, +new); untouched runs cancel exactly.
// Returns the number of merges.
template <class Tab>
__device__ __forceinline__ uint32_t warp_apply_word(uint32_t *st, uint32_t cap, uint32_t *gt, long long f,
                                                    const MergeOp &op, unsigned lane, const Tab &tab,
                                                    WarpQueue &q) {
  if (cap > 32) {  // long word: scalar path on lane 0 (exact, slow)
    uint32_t merges = 0;
    if (lane == 0) {
      for_each_pair(st, cap, [&](uint64_t key, uint64_t mult) {
        if (key != op.key) pair_add(tab, key, -(long long)mult * f);
      });
      merges = rewrite_word(st, cap, op.x, op.y, op.z);
      for_each_pair(st, cap, [&](uint64_t key, uint64_t mult) { pair_add(tab, key, (long long)mult * f); });
    }
    __syncwarp();
    if (gt)
      for (uint32_t i = lane; i < cap; i += 32) gt[i] = st[i];
    return __shfl_sync(0xffffffffu, merges, 0);
  }
  const uint32_t t = lane < cap ? st[lane] : DEAD;
  const bool valid = t != DEAD;
  const uint32_t n = __popc(__ballot_sync(0xffffffffu, valid));  // live tokens form a prefix
  const RunInfo ro = warp_runs(t, n, lane);
  // greedy left-to-right matches
  const uint32_t nx = __shfl_down_sync(0xffffffffu, t, 1);
  bool m = valid && lane + 1 < n && t == op.x && nx == op.y;
  if (op.x == op.y) {  // inside a run x^L only every second position starts a pair (bpe.cpp:654-690)
    const unsigned upto = ro.starts & (lane == 31 ? 0xffffffffu : ((2u << lane) - 1u));
    const uint32_t run_start = upto ? 31u - (uint32_t)__clz(upto) : 0u;
    m = m && (((lane - run_start) & 1u) == 0u);
  }
  const unsigned mm = __ballot_sync(0xffffffffu, m);
  const unsigned touched = mm | (mm << 1);                       // first and second token of every merged pair
  const bool keep = valid && !((mm << 1) >> lane & 1u);
  const unsigned km = __ballot_sync(0xffffffffu, keep);
  const uint32_t newpos = __popc(km & ((1u << lane) - 1u));
  const uint32_t n2 = __popc(km);
  const uint32_t nv = m ? op.z : t;
  // an old run [lane, nxt) (plus the first token of the next run) without a touched token is unchanged
  const unsigned span_mask = (ro.nxt >= 31 ? 0xffffffffu : ((2u << ro.nxt) - 1u)) & ~((1u << lane) - 1u);
  const bool old_changed = ro.start && (touched & span_mask) != 0u;
  const unsigned unchanged_starts = __ballot_sync(0xffffffffu, ro.start && !old_changed);
  __syncwarp();  // every lane holds its old token in a register before any slot is overwritten
  if (keep) { st[newpos] = nv; if (gt) gt[newpos] = nv; }
  if (valid && lane >= n2) { st[lane] = DEAD; if (gt) gt[lane] = DEAD; }
  // new lane p came from old lane src = (p+1)-th kept lane: fetch its token with a shuffle (st may be
  // global memory in the deferred path: no re-read through L1); its run is unchanged iff that old run was
  const uint32_t src = lane < n2 ? __fns(km, 0, lane + 1) : 0u;
  const uint32_t moved = __shfl_sync(0xffffffffu, nv, src & 31);
  const uint32_t t2 = lane < n2 ? moved : DEAD;
  const RunInfo rn = warp_runs(t2, n2, lane);
  const bool new_changed = rn.start && !((unchanged_starts >> (src & 31)) & 1u);
  // the four possible updates of this lane
  const bool h0 = old_changed && ro.L >= 2 && pair_key(t, t) != op.key;
  const bool h1 = old_changed && ro.nxt < n && pair_key(t, ro.b) != op.key;
  const bool h2 = new_changed && rn.L >= 2;
  const bool h3 = new_changed && rn.nxt < n2;
  const uint32_t total = __popc(__ballot_sync(0xffffffffu, h0)) + __popc(__ballot_sync(0xffffffffu, h1)) +
                         __popc(__ballot_sync(0xffffffffu, h2)) + __popc(__ballot_sync(0xffffffffu, h3));
  const bool direct = total > UQ_CAP;
  if (!direct && q.n + total > UQ_CAP) uq_drain(q, lane, tab);
  uq_push(q, direct, h0, pair_key(t, t), -f * (long long)(ro.L >> 1), lane, tab);
  uq_push(q, direct, h1, pair_key(t, ro.b), -f, lane, tab);
  uq_push(q, direct, h2, pair_key(t2, t2), f * (long long)(rn.L >> 1), lane, tab);
  uq_push(q, direct, h3, pair_key(t2, rn.b), f, lane, tab);
  return n - n2;
}

// One RESIDENT tile: its token slots sit in shared memory (`span` slots at stok, word w at
// [soff[w], soff[w+1])).  The scan is TOKEN-parallel: a lane compares four consecutive slots and
// the one after them with (x, y) — consecutive lanes read consecutive 16-byte pieces of shared
// memory, every lane does the same work whatever the word lengths.  A hit cannot straddle two words: the first token
// of every word is (derived from) the "▁" token, which only ever occurs at position 0, so y —
// the second element of an in-word pair — is never a word-initial token; tail padding (DEAD)
// matches nothing.  Each hit is mapped to its word (binary search in the offsets), the word is
// claimed once through a bitmap, and claimed words are rewritten by the whole warp.
template <class Tab>
__device__ __forceinline__ unsigned long long process_tile(uint32_t *stok, const uint32_t *soff, uint32_t nw,
                                                           uint32_t span, uint32_t *claim, const uint64_t *gfreq,
                                                           const MergeOp &op, const Tab &tab, WarpQueue &q) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  unsigned long long dead = 0;
  // a lane compares four tokens from one 16-byte shared load (stok is 16-byte aligned, its capacity a
  // multiple of 4); the token after them comes from the next lane
  for (uint32_t base = wid * 128; base < span; base += nwarp * 128) {
    const uint32_t p = base + lane * 4;
    uint4 v = make_uint4(DEAD, DEAD, DEAD, DEAD);
    if (p < span) v = *reinterpret_cast<const uint4 *>(stok + p);
    uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (lane == 31) nxt = p + 4 < span ? stok[p + 4] : DEAD;
    uint32_t m = (v.x == op.x && v.y == op.y ? 1u : 0u) | (v.y == op.x && v.z == op.y ? 2u : 0u) |
                 (v.z == op.x && v.w == op.y ? 4u : 0u) | (v.w == op.x && nxt == op.y ? 8u : 0u);
    if (p + 4 >= span) m &= span > p + 1 ? (1u << (span - p - 1)) - 1u : 0u;  // occurrence i needs i + 1 < span
    while (__ballot_sync(0xffffffffu, m != 0)) {  // one round per hit of the busiest lane (almost always one)
      uint32_t w = 0, o = 0, cap = 0;
      bool own = false;
      if (m) {
        const uint32_t i = p + (uint32_t)__ffs(m) - 1u;
        m &= m - 1u;
        // re-read: an earlier round (or another warp) may already have rewritten this word
        if (stok[i] == op.x && stok[i + 1] == op.y) {
          uint32_t lo = 0, hi = nw;  // largest w with soff[w] <= i
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (soff[mid] <= i) lo = mid; else hi = mid;
          }
          w = lo;
          o = soff[w];
          cap = soff[w + 1] - o;
          own = !((atomicOr(&claim[w >> 5], 1u << (w & 31)) >> (w & 31)) & 1u);
        }
      }
      const long long fw = own ? (long long)gfreq[w] : 0;  // owners fetch their word's frequency together
      unsigned mask = __ballot_sync(0xffffffffu, own);
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t oj = __shfl_sync(0xffffffffu, o, j), cj = __shfl_sync(0xffffffffu, cap, j);
        const uint32_t wj = __shfl_sync(0xffffffffu, w, j);
        const long long f = __shfl_sync(0xffffffffu, fw, j);
        dead += warp_apply_word(stok + oj, cj, nullptr, f, op, lane, tab, q);
        if (lane == 0) atomicAnd(&claim[wj >> 5], ~(1u << (wj & 31)));  // bitmap all zero between tiles
      }
    }
  }
  return lane == 0 ? dead : 0ull;  // pending table updates stay queued (drained by the caller)
}

// Oversized tile (a word longer than the shared buffer): thread per word straight on global memory.
template <class Tab>
__device__ __forceinline__ unsigned long long process_tile_direct(uint32_t *tok, const uint32_t *off,
                                                                  const uint64_t *freq, uint32_t w0, uint32_t w1,
                                                                  const MergeOp &op, const Tab &tab) {
  unsigned long long dead = 0;
  for (uint32_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
    uint32_t o = off[w], wcap = off[w + 1] - o;
    uint32_t *t = tok + o;
    if (!has_pair(t, wcap, op.x, op.y)) continue;
    long long f = (long long)freq[w];
    for_each_pair(t, wcap, [&](uint64_t key, uint64_t mult) {
      if (key != op.key) pair_add(tab, key, -(long long)mult * f);
    });
    dead += rewrite_word(t, wcap, op.x, op.y, op.z);
    for_each_pair(t, wcap, [&](uint64_t key, uint64_t mult) { pair_add(tab, key, (long long)mult * f); });
  }
  return dead;
}

// ---- TMA (bulk async copy) staging of a tile: global -> shared, completion on an mbarrier ----------
#ifndef YT_SIMT_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 28)) asm volatile("trap;");  // a lost transaction must not hang the box
  }
}
#else
// tests/emul/simt (CPU emulation of the kernels, test harness only): the mbarrier word is modelled as
// { pending arrivals : 16, arrival count : 16, outstanding transaction bytes : 31, phase : 1 }; a bulk copy is an
// immediate memcpy that completes its bytes on the barrier; a waiter yields to the other fibers of the block.
struct EmuMbar { uint64_t pending : 16, count : 16, tx : 31, phase : 1; };
inline EmuMbar *emu_bar(unsigned long long *bar) { return reinterpret_cast<EmuMbar *>(bar); }
inline void emu_bar_check(EmuMbar *m) {
  if (m->pending == 0 && m->tx == 0) { m->phase ^= 1u; m->pending = m->count; }
}
inline void mbar_init(unsigned long long *bar, uint32_t count) {
  EmuMbar *m = emu_bar(bar);
  m->pending = count; m->count = count; m->tx = 0; m->phase = 0;
}
inline void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {  // one arrival + expected bytes
  EmuMbar *m = emu_bar(bar);
  m->tx += bytes;
  m->pending -= 1;
  emu_bar_check(m);
}
inline void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  if (((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) { fprintf(stderr, "emu: misaligned bulk copy\n"); abort(); }
  memcpy(dst, src, bytes);
  EmuMbar *m = emu_bar(bar);
  m->tx -= bytes;
  emu_bar_check(m);
}
inline void mbar_arrive(unsigned long long *bar) {
  EmuMbar *m = emu_bar(bar);
  m->pending -= 1;
  emu_bar_check(m);
}
inline void mbar_wait(unsigned long long *bar, uint32_t parity) {  // returns once the phase of that parity is over
  for (uint64_t spin = 0; emu_bar(bar)->phase == (parity & 1u); spin++) {
    if (spin > (1u << 24)) { fprintf(stderr, "emu: mbarrier wait never completes\n"); abort(); }
    emu::yield();
  }
}
#endif
// bytes of a 16-byte aligned window [lo & ~3, roundup(hi, 4)) over uint32 elements
__device__ __forceinline__ uint32_t win_lo(uint32_t lo) { return lo & ~3u; }
__device__ __forceinline__ uint32_t win_bytes(uint32_t lo, uint32_t hi) { return (((hi + 3u) & ~3u) - (lo & ~3u)) * 4u; }

__device__ __forceinline__ void load_tile(const LoopArgs &a, uint32_t w0, uint32_t w1, uint32_t *stok, uint32_t *soff) {
  const uint32_t o0 = a.off[w0], span = a.off[w1] - o0, nw = w1 - w0;  // resident mode: plain coalesced copy, once
  for (uint32_t i = threadIdx.x; i <= nw; i += blockDim.x) soff[i] = a.off[w0 + i] - o0;
  const uint32_t *src = a.tok + o0;
  for (uint32_t i = threadIdx.x; i < span; i += blockDim.x) stok[i] = __ldcg(src + i);
}

#ifndef YT_SIMT_EMU
extern __shared__ __align__(16) uint32_t yt_dyn_smem[];
#else
#define yt_dyn_smem (reinterpret_cast<uint32_t *>(emu::dyn_smem()))
#endif

template <bool WIDE>
__device__ __forceinline__ void merge_loop_body(const LoopArgs &a) {
  const auto &tab = TabView<WIDE>::of(a.tab);  // WIDE: the table probed four slots per round trip (experimental)
  cg::grid_group grid = cg::this_grid();
  __shared__ Best s_warp[32];
  __shared__ Best s_best;
  __shared__ unsigned long long s_dead;
  __shared__ uint32_t s_defer_n, s_direct;
  __shared__ unsigned long long s_tpre;
  const bool dbgt = (a.dbg & 8u) != 0;  // per-block apply-phase timing (diagnostic)
  // dynamic shared memory: [update queues: 32 warps x UQ_CAP x 16 B][tile tokens][tile offsets]
  unsigned long long *uq_keys = reinterpret_cast<unsigned long long *>(yt_dyn_smem);
  long long *uq_deltas = reinterpret_cast<long long *>(uq_keys + 32 * UQ_CAP);
  uint32_t *s_claim = reinterpret_cast<uint32_t *>(uq_deltas + 32 * UQ_CAP);  // 2 bitmaps of CLAIM_WORDS x 32 flags
  uint32_t *stok = s_claim + 2 * CLAIM_WORDS;
  uint32_t *soff = stok + a.smem_tok_cap;
  // RESIDENT only: word frequencies behind the offsets (8-byte aligned: both caps are even)
  unsigned long long *sfreq = reinterpret_cast<unsigned long long *>(soff + a.smem_word_cap + 2);
  for (uint32_t i = threadIdx.x; i < 2 * CLAIM_WORDS; i += blockDim.x) s_claim[i] = 0;
  // STREAMING carve of the same region: NSTAGE stages of (tokens, offsets), each 16-byte aligned;
  // full[s]: TMA bytes landed (tx count), empty[s]: all consumer warps are done with stage s
  __shared__ __align__(8) unsigned long long s_full[MAX_STAGES], s_empty[MAX_STAGES];
  const uint32_t n_stage = a.n_stage;
  const uint32_t stage_words = a.stream_tok_cap + a.stream_word_cap;  // uint32 per stage
  if (!a.resident && threadIdx.x == 0) {
    for (uint32_t st = 0; st < n_stage; st++) { mbar_init(&s_full[st], 1); mbar_init(&s_empty[st], (blockDim.x >> 5) - 1);  /* consumer warps */ }
#ifndef YT_SIMT_EMU
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }
  // STREAMING pipeline state of this thread (producer lane: empty-phase bits, consumers: full-phase bits)
  uint32_t pipe_used = 0, pipe_phase = 0, pipe_stage = 0;
  WarpQueue uq;
  uq.key = uq_keys + (threadIdx.x >> 5) * UQ_CAP;
  uq.delta = uq_deltas + (threadIdx.x >> 5) * UQ_CAP;
  uq.n = 0;
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t gstride = (uint64_t)gridDim.x * blockDim.x;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const uint64_t cap = a.tab.mask + 1;
  const uint32_t n_done0 = a.ctl->n_done;

  // resident mode: this block's tile moves into shared memory once
  uint32_t rw0 = 0, rw1 = 0;
  if (a.resident && blockIdx.x < a.n_tiles) {
    rw0 = a.tile_desc[blockIdx.x].x;
    rw1 = a.tile_desc[blockIdx.x + 1].x;
    if (rw1 > rw0) {
      load_tile(a, rw0, rw1, stok, soff);
      for (uint32_t i = threadIdx.x; i < rw1 - rw0; i += blockDim.x) sfreq[i] = a.freq[rw0 + i];
    }
  }
  __syncthreads();

  for (uint32_t it = 0; it < a.max_iters; ++it) {
    const uint32_t n_done = n_done0 + it;
    if (n_done >= a.max_total) break;
    // ---------------- arg-max over the table under MergeCandidate::operator< (bpe.cpp:110-126)
    unsigned long long tq0 = gtid == 0 ? gtimer() : 0, tq1 = 0, tq2 = 0, tq3 = 0;
    Best b{0, 0, 0};
    // the counts of SWEEP_UNROLL slots are requested together: one L2 round trip per batch instead of
    // one per slot (a thread owns cap / (148 * 1024) slots, 7 at the 1 M-slot table of the 100 MB corpus)
    for (uint64_t i0 = gtid; i0 < cap; i0 += gstride * SWEEP_UNROLL) {
      unsigned long long c[SWEEP_UNROLL];
#pragma unroll
      for (int u = 0; u < SWEEP_UNROLL; u++) {
        const uint64_t i = i0 + (uint64_t)u * gstride;
        c[u] = i < cap ? __ldcg(a.tab.cnts + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < SWEEP_UNROLL; u++) {
        if (c[u] != 0 && c[u] >= b.c) {
          const uint64_t i = i0 + (uint64_t)u * gstride;
          const unsigned long long k = __ldcg(a.tab.keys + i);
          Best cand{c[u], pair_prio((uint32_t)(k >> 32), (uint32_t)k), i};
          if (better(cand, b)) b = cand;
        }
      }
    }
    b = warp_best(b);
    if (lane == 0) s_warp[wid] = b;
    __syncthreads();
    if (wid == 0) {
      Best v = lane < nwarp ? s_warp[lane] : Best{0, 0, 0};
      v = warp_best(v);
      if (lane == 0) {
        a.blockbest[4 * blockIdx.x + 0] = v.c;
        a.blockbest[4 * blockIdx.x + 1] = v.prio;
        a.blockbest[4 * blockIdx.x + 2] = v.slot;
      }
    }
    if (gtid == 0) tq1 = gtimer();
    grid.sync();
    if (gtid == 0) tq2 = gtimer();
    const unsigned long long tw0 = dbgt && lane == 0 ? gtimer() : 0;
    {  // every block reduces the per-block winners redundantly: one entry per thread, one round trip
      Best v{0, 0, 0};
      for (unsigned j = threadIdx.x; j < gridDim.x; j += blockDim.x) {
        Best w{__ldcg(a.blockbest + 4 * j), __ldcg(a.blockbest + 4 * j + 1), __ldcg(a.blockbest + 4 * j + 2)};
        if (better(w, v)) v = w;
      }
      const unsigned used_warps = (min(gridDim.x, blockDim.x) + 31) >> 5;
      if (wid < used_warps) {
        v = warp_best(v);
        if (lane == 0) s_warp[wid] = v;
      }
      __syncthreads();
      if (wid == 0) {
        Best u = lane < used_warps ? s_warp[lane] : Best{0, 0, 0};
        u = warp_best(u);
        if (lane == 0) { s_best = u; s_dead = 0; s_tpre = 0; }
      }
    }
    __syncthreads();
    const Best win = s_best;
    if (win.c == 0) {  // no pair left: "merged only" (bpe.cpp:1137-1145)
      if (gtid == 0) a.ctl->stop = 1;
      break;
    }
    MergeOp op;
    {  // (x, y) is recoverable from the priority word (pair_prio), no dependent table load needed
      const uint32_t mx = 0xffffffffu - (uint32_t)(win.prio >> 32);
      const uint32_t mn = 0x7fffffffu - (uint32_t)((win.prio & 0xffffffffull) >> 1);
      op.x = (win.prio & 1ull) ? mx : mn;
      op.y = (win.prio & 1ull) ? mn : mx;
      op.key = pair_key(op.x, op.y);
    }
    op.z = a.first_new_id + n_done;
    if (gtid == 0) {
      a.rules[3 * n_done + 0] = op.x; a.rules[3 * n_done + 1] = op.y; a.rules[3 * n_done + 2] = op.z;
      a.rfreq[n_done] = win.c;
      a.tab.cnts[win.slot] = 0;  // every occurrence of (x,y) is merged below; no deltas are sent for it
      a.ctl->n_done = n_done + 1;
    }
    // ---------------- apply x y -> z
    unsigned long long dead = 0;
    if (a.resident) {
      if (rw1 > rw0) dead = process_tile(stok, soff, rw1 - rw0, soff[rw1 - rw0], s_claim,
                                           reinterpret_cast<const uint64_t *>(sfreq), op, tab, uq);
    } else {
      // STREAMING: this block owns a contiguous chunk of tiles that flows through an n_stage ring of
      // shared-memory stages.  Lane 0 of warp 0 is the PRODUCER: it keeps up to n_stage tiles in
      // flight with TMA bulk copies (tokens + offsets, 16-byte aligned windows, completion on
      // full[s]) and refills a stage as soon as all consumer warps have released it (empty[s]).
      // Warps 1.. are CONSUMERS: they scan a tile out of shared memory token-parallel and append
      // the words that hold (x,y) to the block's deferred list; no block barrier in the tile loop.
      // The deferred words are rewritten afterwards straight in HBM, one warp per word, so their
      // latencies overlap each other instead of stalling the copy pipeline.
      const uint32_t per_block = (a.n_tiles + gridDim.x - 1) / gridDim.x;
      const uint32_t k_first = min(a.n_tiles, blockIdx.x * per_block);
      const uint32_t my_tiles = min(a.n_tiles, k_first + per_block) - k_first;
      uint4 *defer = a.defer + (size_t)blockIdx.x * a.defer_cap;
      if (threadIdx.x == 0) { s_defer_n = 0; s_direct = 0; }
      __syncthreads();
      auto staged = [&](uint2 d0, uint2 d1) {  // does this tile go through shared memory?
        return d1.x > d0.x && win_bytes(d0.y, d1.y) <= a.stream_tok_cap * 4u &&
               win_bytes(d0.x, d1.x + 1) <= a.stream_word_cap * 4u;
      };
      // tile descriptors travel in lane-distributed batches: lane j of a warp holds tile_desc[base + j],
      // tiles base .. base+30 take (d0, d1) from lanes (j, j+1) by shuffle, so the L2 latency of the
      // descriptor loads is paid once per 31 tiles instead of once per tile
      auto load_batch = [&](uint32_t base) {
        const uint32_t idx = min(k_first + base + lane, k_first + my_tiles);
        return a.tile_desc[idx];
      };
      if (wid == 0) {
        uint32_t used = pipe_used, ephase = pipe_phase, st = pipe_stage;  // barrier phases persist across merges
        uint2 batch = make_uint2(0, 0);
        for (uint32_t t = 0; t < my_tiles; t++) {
          const uint32_t j = t % 31;
          if (j == 0) batch = load_batch(t);
          uint2 d0, d1;
          d0.x = __shfl_sync(0xffffffffu, batch.x, j);     d0.y = __shfl_sync(0xffffffffu, batch.y, j);
          d1.x = __shfl_sync(0xffffffffu, batch.x, j + 1); d1.y = __shfl_sync(0xffffffffu, batch.y, j + 1);
          if (!staged(d0, d1)) continue;
          if (lane == 0) {
            if ((used >> st) & 1u) mbar_wait(&s_empty[st], (ephase >> st) & 1u);
            uint32_t *dst = stok + st * stage_words;
            const uint32_t bt = win_bytes(d0.y, d1.y), bo = win_bytes(d0.x, d1.x + 1);
            mbar_expect_tx(&s_full[st], bt + bo);
            tma_bulk_g2s(dst, a.tok + win_lo(d0.y), bt, &s_full[st]);
            tma_bulk_g2s(dst + a.stream_tok_cap, a.off + win_lo(d0.x), bo, &s_full[st]);
          }
          if ((used >> st) & 1u) ephase ^= 1u << st;
          used |= 1u << st;
          st = st + 1 == n_stage ? 0 : st + 1;
          __syncwarp();
        }
        pipe_used = used; pipe_phase = ephase; pipe_stage = st;
      } else {
        uint32_t fphase = pipe_phase, st = pipe_stage;
        const unsigned cw = wid - 1, ncw = nwarp - 1;  // consumer warp index / count
        uint2 batch = make_uint2(0, 0);
        for (uint32_t t = 0; t < my_tiles; t++) {
          const uint32_t j = t % 31;
          if (j == 0) batch = load_batch(t);
          uint2 d0, d1;
          d0.x = __shfl_sync(0xffffffffu, batch.x, j);     d0.y = __shfl_sync(0xffffffffu, batch.y, j);
          d1.x = __shfl_sync(0xffffffffu, batch.x, j + 1); d1.y = __shfl_sync(0xffffffffu, batch.y, j + 1);
          if (d1.x <= d0.x) continue;
          if (!staged(d0, d1)) { if (cw == 0 && lane == 0) s_direct = 1; continue; }  // oversized: direct pass below
          mbar_wait(&s_full[st], (fphase >> st) & 1u);
          fphase ^= 1u << st;
          const uint32_t *stg = stok + st * stage_words;     // 16-byte aligned window of the stage
          const uint32_t head = d0.y - win_lo(d0.y);         // 0..3 tokens of the previous tile in front
          const uint32_t *tk = stg + head;
          const uint32_t *of = stg + a.stream_tok_cap + (d0.x - win_lo(d0.x));
          const uint32_t span = d1.y - d0.y, nw = d1.x - d0.x, obase = d0.y;
          auto report = [&](uint32_t i) {  // token i of the tile starts an (x,y) occurrence
            uint32_t lo = 0, hi = nw;      // largest w with of[w] - obase <= i
            while (hi - lo > 1) {
              const uint32_t mid = (lo + hi) >> 1;
              if (of[mid] - obase <= i) lo = mid; else hi = mid;
            }
            const uint32_t o = of[lo] - obase, wcap = of[lo + 1] - obase - o;
            for (uint32_t q2 = o; q2 < i; q2++)  // only the first hit of a word reports it
              if (tk[q2] == op.x && tk[q2 + 1] == op.y) return;
            const uint32_t slot = atomicAdd(&s_defer_n, 1u);
            if (slot < a.defer_cap) defer[slot] = make_uint4(d0.x + lo, obase + o, wcap, 0u);
            else s_direct = 1;  // list full: the direct pass below picks the rest up
          };
          if (a.dbg & 1u) {
          } else if (a.dbg & 2u) {
            for (uint32_t base = cw * 32; base < span; base += ncw * 32) {
              const uint32_t i = base + lane;
              const bool hit = i + 1 < span && tk[i] == op.x && tk[i + 1] == op.y;
              if (!__ballot_sync(0xffffffffu, hit)) continue;
              if (hit) report(i);
            }
          } else {
            // eight tokens per lane from two 16-byte shared loads; the token after them comes from the
            // next lane.  p = position in the aligned window; occurrence p is inside the tile iff
            // head <= p and p + 1 < total (a word-initial token is never y, so none straddles tiles).
            // Measured on B200, 2 stages: scalar 45 %, 4 per lane 74-87 %, 8 per lane 75-92 % of HBM peak.
            const uint32_t total = head + span;
            for (uint32_t base = cw * 256; base < total; base += ncw * 256) {
              const uint32_t p = base + lane * 8;
              uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u), u = v;
              if (p < total) v = *reinterpret_cast<const uint4 *>(stg + p);
              if (p + 4 < total) u = *reinterpret_cast<const uint4 *>(stg + p + 4);
              uint32_t nxt = __shfl_down_sync(0xffffffffu, v.x, 1);
              if (lane == 31) nxt = p + 8 < total ? stg[p + 8] : ~0u;
              uint32_t m = (v.x == op.x && v.y == op.y ? 1u : 0u) | (v.y == op.x && v.z == op.y ? 2u : 0u) |
                           (v.z == op.x && v.w == op.y ? 4u : 0u) | (v.w == op.x && u.x == op.y ? 8u : 0u) |
                           (u.x == op.x && u.y == op.y ? 16u : 0u) | (u.y == op.x && u.z == op.y ? 32u : 0u) |
                           (u.z == op.x && u.w == op.y ? 64u : 0u) | (u.w == op.x && nxt == op.y ? 128u : 0u);
              if (!__ballot_sync(0xffffffffu, m)) continue;
              while (m) {
                const uint32_t pk = p + (uint32_t)__ffs(m) - 1u;
                m &= m - 1u;
                if (pk >= head && pk + 1 < total) report(pk - head);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[st]);
          st = st + 1 == n_stage ? 0 : st + 1;
        }
        pipe_phase = fphase; pipe_stage = st;
      }
      __syncthreads();
      {
        const uint32_t n_def = min(s_defer_n, a.defer_cap);
        for (uint32_t j = wid; j < n_def; j += nwarp) {
          const uint4 e = defer[j];
          const long long f = (long long)a.freq[e.x];
          const uint32_t merges = warp_apply_word(a.tok + e.y, e.z, nullptr, f, op, lane, tab, uq);
          if (lane == 0) dead += merges;
        }
        if (s_direct && my_tiles) {  // oversized tiles / overflowed list: exact thread-per-word pass on global memory
          __syncthreads();
          const uint32_t w_lo = a.tile_desc[k_first].x, w_hi = a.tile_desc[k_first + my_tiles].x;
          dead += process_tile_direct(a.tok, a.off, a.freq, w_lo, w_hi, op, tab);
        }
      }
      // write-through stores (generic proxy) must be ordered before the next iteration's bulk loads
#ifndef YT_SIMT_EMU
      asm volatile("fence.proxy.async;" ::: "memory");
#endif
    }
    if (dbgt && lane == 0) atomicMax(&s_tpre, gtimer() - tw0);
    if (uq.n) uq_drain(uq, lane, tab);  // one batch of table updates per warp and iteration
    for (int o = 16; o > 0; o >>= 1) dead += __shfl_xor_sync(0xffffffffu, dead, o);
    if (lane == 0 && dead) atomicAdd(&s_dead, dead);
    __syncthreads();
    if (threadIdx.x == 0 && s_dead) atomicAdd(&a.ctl->dead, s_dead);
    if (dbgt && threadIdx.x == 0) {
      const unsigned long long t = gtimer() - tw0;
      atomicMax(&a.ctl->blk[it & 1][0], t);
      atomicAdd(&a.ctl->blk[it & 1][1], t);
      atomicMax(&a.ctl->blk[it & 1][2], s_tpre);
    }
    if (gtid == 0) tq3 = gtimer();
    grid.sync();
    if (gtid == 0) {
      unsigned long long tq4 = gtimer();
      if (dbgt) {
        a.ctl->t_phase[4] += __ldcg(&a.ctl->blk[it & 1][0]);
        a.ctl->t_phase[5] += __ldcg(&a.ctl->blk[it & 1][1]) / gridDim.x;
        a.ctl->t_phase[6] += __ldcg(&a.ctl->blk[it & 1][2]);
        a.ctl->blk[it & 1][0] = 0; a.ctl->blk[it & 1][1] = 0; a.ctl->blk[it & 1][2] = 0;
      }
      a.ctl->t_phase[0] += tq1 - tq0; a.ctl->t_phase[1] += tq2 - tq1;
      a.ctl->t_phase[2] += tq3 - tq2; a.ctl->t_phase[3] += tq4 - tq3;
      a.ctl->iters += 1;
    }
    // ---------------- uniform exit checks (every block reads the same values)
    uint32_t nk = __ldcg(&a.ctl->n_keys), ov = __ldcg(&a.ctl->overflow);
    unsigned long long dd = __ldcg(&a.ctl->dead), sl = __ldcg(&a.ctl->slots);
    if (ov || nk > a.key_limit) { if (gtid == 0) a.ctl->stop = 2; break; }
    if (dd * 4 > sl && sl > 65536) { if (gtid == 0) a.ctl->stop = 3; break; }
  }
  // resident tiles go back to HBM on every exit path
  __syncthreads();
  if (a.resident && rw1 > rw0) {
    const uint32_t o0 = a.off[rw0], span = a.off[rw1] - o0;
    for (uint32_t i = threadIdx.x; i < span; i += blockDim.x) a.tok[o0 + i] = stok[i];
  }
}
__global__ void __launch_bounds__(1024, 1) merge_loop_kernel(LoopArgs a) { merge_loop_body<false>(a); }
__global__ void __launch_bounds__(1024, 1) merge_loop_wide_kernel(LoopArgs a) { merge_loop_body<true>(a); }  // YTTM_LOOP_WIDEPROBE

// ---- tile planning -----------------------------------------------------------------------------
// tile k = words whose first token slot lies in [k*q, (k+1)*q); tile_desc[k] = (its first word,
// that word's token offset); tile_desc[n_tiles] = (n_words, n_slots).
__global__ void tile_desc_kernel(const uint32_t *__restrict__ off, uint64_t n_words, uint32_t q, uint32_t n_tiles,
                                 uint2 *__restrict__ desc) {
  uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const uint32_t ow = off[w];
  const uint32_t k1 = ow / q;
  const int64_t k0 = w == 0 ? -1 : (int64_t)(off[w - 1] / q);
  for (int64_t k = k0 + 1; k <= (int64_t)k1; k++) desc[k] = make_uint2((uint32_t)w, ow);
  if (w + 1 == n_words)
    for (uint32_t k = k1 + 1; k <= n_tiles; k++) desc[k] = make_uint2((uint32_t)n_words, off[n_words]);
}
__global__ void tile_stats_kernel(const uint2 *__restrict__ desc, uint32_t n_tiles,
                                  uint32_t *out /* [0] max span, [1] max words */) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_tiles) return;
  uint2 d0 = desc[k], d1 = desc[k + 1];
  if (d1.x <= d0.x) return;
  atomicMax(out, d1.y - d0.y);
  atomicMax(out + 1, d1.x - d0.x);
}

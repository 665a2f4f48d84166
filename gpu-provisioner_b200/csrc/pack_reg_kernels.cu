// pack_reg_kernels.cu — the register-resident form of the pack (Scheduler.Solve's sequential loop,
// K/controllers/provisioning/scheduling/scheduler.go:394-421, :488-675) for batches in which no pod defines a
// requirement key and there are no existing nodes: the common provisioning batch, and BASELINE config[1].
//
// pack_kernels.cu walks s.newNodeClaims in order, lane group per claim, claim state in global memory.  A step there
// is a chain of dependent loads (order -> claim state -> tables -> type mask) and several CTA barriers.  On this
// machine a warp's dependent instruction stream runs at 6-10 cycles per instruction whatever it does, so a step costs
// what its longest role ISSUES (and, with sixteen warps on one SM, what all of them issue together: the kernel fills
// 42 % of its SM's issue slots); this form is built to issue little:
//
//   * every in-flight NodeClaim is owned by one thread and lives in its registers: the accumulated requests as
//     32-bit multiples of each resource's granule (the gcd of every quantity of that resource in the problem — exact;
//     a problem whose scaled values do not fit falls back to pack_kernel);
//   * the claim ORDER is never read by the scan.  Every claim thread knows its claims' POSITIONS in s.newNodeClaims,
//     so "first feasible claim in order" (scheduler.go:566-575) == "feasible claim of least position": one warp
//     min-reduce + one shared-memory atomicMin + one barrier per pod.  Positions are kept up to date locally: between
//     two of the reference's sort.Slice calls (scheduler.go:499) one claim changed, and unless sort.Slice takes the
//     real pdqsort path the result is that claim moving to the end of the run of claims with as many pods (a new claim:
//     behind the last claim with one pod) — a shift by one of a range of positions, which the winner derives from a
//     count table (claims with more than k pods) and publishes in one word for the next step;
//   * the minimum of a step is taken BEFORE the previous step's shift is known (that costs no second barrier): over
//     priorities 2 * pos + 1 in which the claim that changed took a value between its new neighbours' old ones — every
//     other claim keeps its place among the others, so the order of the priorities is the order of the true positions;
//   * the ORDER WARP does nothing in an ordinary step.  Only when sort.Slice could take the real pdqsort path — fewer
//     than 50 claims and an unsorted order, the changed position next to one of choosePivot's nine samples (the keys
//     decide, and nobody holds them), or a pod count beyond the count table — do the claim threads lay the order out
//     in shared memory as it is, the order warp runs the exact emulation (b200_core.cuh GoSort), positions are re-read
//     and the minimum is taken again.  That is lazy in the reference's own way: its sort runs at the start of the
//     NEXT add(), so the last step's change is simply never applied and Results.NewNodeClaims needs no undo;
//   * a claim of selector-less pods keeps exactly those instance types of its template's pre-filtered set that fit
//     its requests (nodeclaim.go:144-146, :392-423; the rows "alloc[r] >= v" are nested), so CanAdd's filter result
//     is non-empty iff the FEASIBILITY TABLE cell of the new total's per-resource ranks is set: a bit table over
//     rank space, built once per Solve by feas_table_kernel from the bit rows.  The rank of a total is one byte load
//     from a direct value -> rank table (resources with a small scaled range), one 8-byte bucket entry for the one
//     resource whose range is too large (memory), or — when the catalog's values do not bucket — a 16-byte window of the
//     next distinct allocatable values and four compares per resource.  No type mask is read or written in the loop; the
//     final InstanceTypeOptions are the first pod's stage-A row inside the rows of the final totals;
//   * the queue warp prepares pod rows 32 at a time, a batch ahead (Queue.Pop, queue.go:46-60: until the first
//     pass over the queue is over the pop order is the FFD order), and appends pushed-back pods (queue.go:63-66) to
//     a side list that the second phase then cycles through with the lastLen rule.
//
// The host twin (tests/native/emul.cpp emul_solve_reg) runs the same algorithm serially — positions, shift words, the
// minimum over priorities that lack the last shift, the sort on demand — and is checked against the oracle on the CPU
// (fuzz, and config[1] with the kernel's own counters: 318 meetings, 284 full pdqsorts); the kernel is checked against
// the oracle in tests/test_gpu.py.  A problem that needs more claims than the threads hold leaves status 0 and the next
// form, then pack_kernel (launched right after, skipping finished problems), redoes it.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <numeric>
#include <type_traits>

#include "device_types.cuh"
#include "kernels.h"

namespace b200 {

namespace {

#ifndef B200S_REG_THREADS
#define B200S_REG_THREADS 512
#endif
constexpr int REG_THREADS = B200S_REG_THREADS;
#ifndef B200S_REG_PREFETCH
#define B200S_REG_PREFETCH 0
#endif
#ifndef B200S_REG_DEFER
#define B200S_REG_DEFER 1
#endif
constexpr bool REG_PREFETCH = B200S_REG_PREFETCH != 0;  // read the next pod's row before the step's barrier
constexpr bool REG_DEFER = B200S_REG_DEFER != 0;        // the queue warp writes the winners' placements out
constexpr int REG_NW = REG_THREADS / 32;
constexpr int REG_CW = REG_NW - 2;     // claim warps; then the order warp and the queue warp
constexpr int REG_NCT = REG_CW * 32;   // claim threads
constexpr int REG_ORDW = REG_CW, REG_QW = REG_CW + 1;
constexpr int RING = 64;
constexpr int RT = REG_RT;             // resource slots of a claim (catalog resources, padded)
constexpr unsigned PRIO_INF = 0xffffffffu;

struct RingEnt {  // a popped pod's row: updateCachedPodData's products (scheduler.go:467-486) through its class ids
  int32_t q[8];   // request per resource slot, in granules
  uint64_t tol;
  int32_t p, tnew;
};

struct RegShared {
  uint64_t taints[MAXT];
  int32_t daemon[MAXT][8];
  uint64_t all_taints;
  unsigned best[3], best2[3];
  // what step t did to the order, for the claim threads (read after the next step's barrier): positions [lo, hi] move
  // by one (mv_word); and, when sort.Slice might take the real pdqsort path on it, what the order warp needs to run
  // that sort: the changed position, whether the claim was appended, the number of claims
  unsigned mv[2];
  int32_t meet_a, meet_n, done, _pad3;
  // first pass over the queue: what step t did, for the queue warp to write out one step later (the winner's stores
  // to global memory are off the step's critical path): claim id, or -1 nothing to write, -2 a new claim (written by
  // its owner)
  int32_t wlog[RING];
  int32_t head, len, n_pushed, _pad;
  RingEnt ring[RING];
};

__host__ __device__ inline size_t reg_align16(size_t x) { return (x + 15) & ~(size_t)15; }

constexpr int REG_KCAP = 2048;  // pod counts below this have an entry in the count table

struct RegLayout {
  size_t tab, win, bits, ord, key, pos, gc, total;
};

__host__ __device__ inline RegLayout reg_layout(int tab_bytes, int win_entries, int T, int wpt, int cap) {
  RegLayout l;
  size_t o = reg_align16(sizeof(RegShared));
  l.tab = o; o += reg_align16((size_t)tab_bytes);
  l.win = o; o += reg_align16((size_t)win_entries * 16);
  l.bits = o; o += reg_align16((size_t)(T > 0 ? T : 1) * wpt * 4);
  l.ord = o; o += reg_align16((size_t)cap * 4);
  l.key = o; o += reg_align16((size_t)cap * 4);
  l.pos = o; o += reg_align16((size_t)cap * 4);
  l.gc = o; o += reg_align16((size_t)REG_KCAP * 8);   // the claim threads' own two copies (see the loop)
  l.total = o;
  return l;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// The full pdqsort by one warp (b200_core.cuh GoSort), with the order before it saved for the undo and the claim ->
// position map rebuilt after it.  One copy of the code: the step loop's instruction footprint matters.
__device__ __forceinline__ void ord_full_sort(int32_t* ord, int32_t* key, int32_t* pos_of, int32_t* snap, int n) {
  const int lane = threadIdx.x & 31;
  if (snap) {
    for (int i = lane; i < n; i += 32) snap[i] = ord[i];
    __syncwarp();
  }
  GoSort gs{ord, key, true};
  gs.sort(n);
  __syncwarp();
  for (int i = lane; i < n; i += 32) pos_of[ord[i]] = i;
  __syncwarp();
}

// What sort.Slice (scheduler.go:499) does to an order that was sorted before one element changed (see pack_kernels.cu
// warp_normalise for the derivation), by one warp on the explicit arrays; keeps pos_of (claim -> position).
__device__ __forceinline__ void ord_move(int32_t* ord, int32_t* key, int32_t* pos_of, int src, int dst) {
  const int lane = threadIdx.x & 31;
  const int x = ord[src], kx = key[src];
  __syncwarp();
  if (src < dst) {
    for (int cb = src + 1; cb <= dst; cb += 32) {
      const int i = cb + lane;
      int v = 0, kv = 0;
      if (i <= dst) { v = ord[i]; kv = key[i]; }
      __syncwarp();
      if (i <= dst) { ord[i - 1] = v; key[i - 1] = kv; pos_of[v] = i - 1; }
      __syncwarp();
    }
  } else {
    for (int ce = src - 1; ce >= dst; ce -= 32) {
      const int i = ce - lane;
      int v = 0, kv = 0;
      if (i >= dst) { v = ord[i]; kv = key[i]; }
      __syncwarp();
      if (i >= dst) { ord[i + 1] = v; key[i + 1] = kv; pos_of[v] = i + 1; }
      __syncwarp();
    }
  }
  if (lane == 0) { ord[dst] = x; key[dst] = kx; pos_of[x] = dst; }
  __syncwarp();
}

__device__ __forceinline__ bool ord_normalise(int32_t* ord, int32_t* key, int32_t* pos_of, int n, int pos, bool appended) {
  const int lane = threadIdx.x & 31;
  if (n <= 1) return false;
  const bool sorted = appended ? !(key[n - 1] < key[n - 2]) : (pos == n - 1 || !(key[pos + 1] < key[pos]));
  if (sorted) return false;
  bool increasing = false;
  if (n >= 50) {  // choosePivot_func's nine samples: no swap <=> the three triples and their middles are non-decreasing
    const int q = n / 4;
    int v = 0;
    if (lane < 9) v = key[q * (lane / 3 + 1) + (lane % 3) - 1];
    const int up = __shfl_down_sync(0xffffffffu, v, 1);
    const int mid_next = __shfl_down_sync(0xffffffffu, v, 3);
    bool bad = false;
    if (lane < 9 && (lane % 3) != 2) bad = up < v;
    if (lane == 1 || lane == 4) bad = bad || mid_next < v;
    increasing = __ballot_sync(0xffffffffu, bad) == 0;
  }
  if (!increasing) {
    ord_full_sort(ord, key, pos_of, nullptr, n);
    return true;
  }
  // partialInsertionSort re-inserts the displaced element next to its equals
  const int kx = key[pos];
  int dst = pos;
  if (!appended) {
    for (int b0 = pos + 1;; b0 += 32) {
      const int i = b0 + lane;
      const bool stop = i >= n || !(key[i] < kx);
      const unsigned sb = __ballot_sync(0xffffffffu, stop);
      if (sb) { dst = b0 + __ffs((int)sb) - 2; break; }
    }
  } else {
    for (int t0 = pos - 1;; t0 -= 32) {
      const int i = t0 - lane;
      const bool stop = i < 0 || !(kx < key[i]);
      const unsigned sb = __ballot_sync(0xffffffffu, stop);
      if (sb) { dst = t0 - (__ffs((int)sb) - 1) + 1; break; }
    }
  }
  if (dst != pos) ord_move(ord, key, pos_of, pos, dst);
  return false;
}

constexpr int REG_CTHREADS = (REG_NW - 1) * 32;  // the step loop's barrier: every warp but the order warp
__device__ __forceinline__ void step_bar() { named_bar_sync(2, REG_CTHREADS); }
// the order warp and the claim warps (the queue warp has no part in the sort)
__device__ __forceinline__ void meet_bar() { named_bar_sync(3, REG_CTHREADS); }
__device__ __forceinline__ void claim_bar() { named_bar_sync(4, REG_NCT); }
// lane 0: *p = min(*p, v) on shared memory, predicated instead of branched
__device__ __forceinline__ void lane0_min(unsigned* p, unsigned v, int lane) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("{ .reg .pred p0; setp.eq.s32 p0, %2, 0; @p0 red.shared.min.u32 [%0], %1; }" ::"r"(a), "r"(v), "r"(lane) : "memory");
}

// the order change of a step as the claim threads see it: bits [11:0] lo, [23:12] hi, [24] the positions lo..hi move up
// (a new claim was inserted before them) instead of down (a claim that gained a pod passed them), [25] sort.Slice may
// take the real pdqsort path on it: the order warp runs it.  No claim sits at position 0xFFF: that is the empty range.
constexpr unsigned MV_EMPTY = 0xFFFu | (0xFFFu << 12), MV_UP = 1u << 24, MV_MEET = 1u << 25;
__device__ __forceinline__ unsigned mv_word(int lo, int hi, bool up, bool meet) {
  return (lo <= hi ? ((unsigned)lo | ((unsigned)hi << 12)) : MV_EMPTY) | (up ? MV_UP : 0u) | (meet ? MV_MEET : 0u);
}

struct RegParams {
  int32_t R, W, K, NB, T;
  int32_t wpt;                // feasibility-table words per template
  int32_t cap;                // claims the CTA's threads can hold
  int32_t tab_bytes, win_entries;
  // per resource slot (direct-table slots first, window slots last; slots past the catalog's resources are inert)
  int32_t perm[RT];           // catalog resource of the slot, -1 for padding
  int32_t lim[RT];            // direct slot: the value -> rank table has lim + 1 entries (lim = 2 * maxv + 1)
  int32_t toff[RT];           // direct slot: offset into the byte table; window slot: offset into the window table
  int32_t stride[RT];         // of the slot's rank in the feasibility table
  int32_t row0[RT];           // first bit row "alloc >= dv[0]" of the slot's resource
  int32_t bsh;                // bucket slot: log2 of the bucket width
  int64_t scale[RT];          // granule
  const uint8_t* tab;         // direct value -> rank tables (global; staged into shared memory)
  const int4* win;            // windows {dv[i], dv[i+1], dv[i+2], dv[i+3]} (global; staged)
  const uint32_t* bits;       // feasibility table (global; staged)
  const int32_t* daemon;      // [MAXT][8] template daemon overhead in granules, by slot
  const int32_t* rq;          // [RT][NQ] request classes in granules, by slot
  const uint64_t* rows;       // bit rows (global; epilogue only)
  int8_t bslot[MAXK];         // bounded slot of key k, or -1
};

// One in-flight NodeClaim in its owner's registers.
struct Slot {
  int32_t x[RT];       // Spec.Resources.Requests in granules
  uint64_t wrk, nwrk;  // table row (rank) of each window slot's total, 8 bits each; nwrk: after this step's pod
  uint32_t pos;        // position in s.newNodeClaims (sorted by len(Pods)) as of the last order change applied
  uint32_t prio;       // what the step's minimum is taken over: 2 * pos + 1, except right after this claim moved
  uint32_t cur;        // what this slot entered the step's minimum with: prio when the pod fits, else PRIO_INF
  int32_t npods;       // len(Pods)
  int32_t tmpl;
  uint64_t taint;      // the template's taints (bit per distinct taint)
};

__device__ __forceinline__ void slot_clear(Slot& s) {
#pragma unroll
  for (int r = 0; r < RT; r++) s.x[r] = 0;
  s.wrk = 0; s.nwrk = 0; s.pos = PRIO_INF; s.prio = PRIO_INF; s.npods = 0; s.tmpl = 0; s.cur = PRIO_INF; s.taint = 0;
}

// NodeClaim.CanAdd for a pod without requirement keys (nodeclaim.go:114-163): whether the filter result of the new
// total is non-empty — the feasibility-table cell of its ranks — and the taints (nodeclaim.go:116).  Returns the
// claim's priority when the pod fits, PRIO_INF otherwise (a slot without a claim has priority PRIO_INF).  Straight-
// line code.  NWIN == 1: the last slot's rank comes from a bucket table {the one distinct value in the bucket, ranks
// below the bucket}; NWIN == RT: every slot tracks its rank and looks at a window of the next distinct values.
template <int NWIN>
__device__ __forceinline__ unsigned slot_eval(Slot& s, const RegParams& prm, const uint8_t* s_tab, const int4* s_win,
                                              const uint32_t* s_bits, const int32_t (&q)[RT], uint64_t tol) {
  int idx = 0;
  uint64_t nw = 0;
  constexpr int NDIR = NWIN == RT ? 0 : RT - 1;
#pragma unroll
  for (int r = 0; r < NDIR; r++) {
    // the table reaches 2 * maxv + 1 >= any claim total (<= maxv) + any pod request (clamped to maxv + 1)
    idx += (int)s_tab[prm.toff[r] + s.x[r] + q[r]] * prm.stride[r];
  }
  if constexpr (NWIN == 1) {
    constexpr int r = RT - 1;
    const int x = s.x[r] + q[r];
    const int2 b = reinterpret_cast<const int2*>(s_win)[prm.toff[r] + (x >> prm.bsh)];
    idx += (b.y + (b.x < x ? 1 : 0)) * prm.stride[r];
  } else {
#pragma unroll
    for (int r = 0; r < RT; r++) {
      const int sh = 8 * r;
      const int x = s.x[r] + q[r];
      int i = (int)((s.wrk >> sh) & 0xFFull);
      const int4 w = s_win[prm.toff[r] + i];
      int c = (w.x < x) + (w.y < x) + (w.z < x) + (w.w < x);
      if (c == 4) {  // more than three rows ahead (rare)
        int j = i + 4;
        while (s_win[prm.toff[r] + j].x < x) j++;  // the list ends in INT32_MAX
        c = j - i;
      }
      i += c;
      idx += i * prm.stride[r];
      nw |= (uint64_t)i << sh;
    }
  }
  const bool f = ((s_bits[s.tmpl * prm.wpt + (idx >> 5)] >> (idx & 31)) & 1u) != 0 && (s.taint & ~tol) == 0;
  s.nwrk = nw;
  s.cur = f ? s.prio : PRIO_INF;
  return s.cur;
}

// NewNodeClaim + Add (nodeclaim.go:83-109, :168-180): requests = daemon overhead + the pod; the claim goes behind the
// last claim with one pod, position dst — in the not yet shifted numbering of the others: just before their dst
template <int NWIN>
__device__ __forceinline__ void slot_open(Slot& s, const RegParams& prm, const int4* s_win, const int32_t* daemon,
                                          const int32_t (&q)[RT], int tnew, uint64_t taint, unsigned pos, int dst) {
  uint64_t nw = 0;
#pragma unroll
  for (int r = 0; r < RT; r++) {
    s.x[r] = daemon[r] + q[r];
    if (NWIN == RT) {
      int i = 0;
      while (s_win[prm.toff[r] + i].x < s.x[r]) i++;
      nw |= (uint64_t)i << (8 * r);
    }
  }
  s.wrk = nw;
  s.tmpl = tnew;
  s.taint = taint;
  s.npods = 1;
  s.pos = pos;
  s.prio = 2u * (unsigned)dst;
  s.cur = PRIO_INF;
}

// PROF: cycle counters of the warp leaders into PackProblem::prof.  Claim warp 0: [0] step start -> barrier (evaluate),
// [1] wait at the barrier, [7] barrier -> winner known (read the minimum, shift positions), [2] the rest of the step
// (commit), [8] cycles in steps that met the order warp, [9] such steps, [3] steps.  Order warp: [4] waiting for a
// meeting, [5] sorting, [10] full pdqsorts.  Queue warp: [11] its own work before the barrier.
template <int NWIN, int CPT, bool PROF>
__global__ void __launch_bounds__(REG_THREADS, 1) pack_reg_kernel(const __grid_constant__ RegParams prm,
                                                                  const TmplDev* __restrict__ tg, PodsDev pods, PackArrays A,
                                                                  PackProblem* __restrict__ problems) {
  extern __shared__ __align__(16) unsigned char smem[];
  RegShared& S = *reinterpret_cast<RegShared*>(smem);
  const RegLayout L = reg_layout(prm.tab_bytes, prm.win_entries, prm.T, prm.wpt, prm.cap);
  uint8_t* s_tab = reinterpret_cast<uint8_t*>(smem + L.tab);
  int4* s_win = reinterpret_cast<int4*>(smem + L.win);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(smem + L.bits);
  int32_t* s_ord = reinterpret_cast<int32_t*>(smem + L.ord);
  int32_t* s_key = reinterpret_cast<int32_t*>(smem + L.key);
  int32_t* s_pos = reinterpret_cast<int32_t*>(smem + L.pos);
  int32_t* s_Gc = reinterpret_cast<int32_t*>(smem + L.gc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_claim = warp < REG_CW;

  PackProblem& pb = problems[blockIdx.x];
  const int base = pb.pod_begin, n_pods = pb.pod_end - pb.pod_begin;
  if (pb.node_end != pb.node_begin) return;  // existing nodes: pack_kernel's problem (uniform)
  if (pb.status == 1) return;                // finished by the smaller form of this kernel (uniform)
  int32_t* push_list = reinterpret_cast<int32_t*>(A.c_hi + (size_t)base * prm.R);  // scratch: pods pushed back, a ring of n_pods
  int32_t* c_first = A.c_nopt + base;  // scratch: the pod that opened the claim
  const int32_t* queue = A.queue + base;

  // ---- stage tables ----
  for (int i = tid; i < MAXT; i += REG_THREADS) {
    S.taints[i] = i < tg->T ? tg->taints[i] : 0ull;
    for (int r = 0; r < 8; r++) S.daemon[i][r] = prm.daemon[i * 8 + r];
  }
  for (int i = tid; i < prm.tab_bytes; i += REG_THREADS) s_tab[i] = prm.tab[i];
  for (int i = tid; i < prm.win_entries; i += REG_THREADS) s_win[i] = prm.win[i];
  for (int i = tid; i < prm.T * prm.wpt; i += REG_THREADS) s_bits[i] = prm.bits[i];
  for (int i = tid; i < REG_KCAP; i += REG_THREADS) { s_Gc[i] = 0; s_Gc[REG_KCAP + i] = 0; }
  if (tid == 0) {
    uint64_t at = 0;
    for (int t = 0; t < tg->T; t++) at |= tg->taints[t];
    S.all_taints = at;
    for (int i = 0; i < 3; i++) { S.best[i] = PRIO_INF; S.best2[i] = PRIO_INF; }
    S.mv[0] = MV_EMPTY; S.mv[1] = MV_EMPTY;
    S.done = 0;
    for (int i = 0; i < RING; i++) S.wlog[i] = -1;
    S.head = 0; S.len = 0; S.n_pushed = 0;
  }
  if (warp == REG_QW && lane < n_pods) {  // the first batch of pod rows
    RingEnt& e = S.ring[lane];
    const int p = queue[lane];
    const int rq = pods.req_id[p];
#pragma unroll
    for (int r = 0; r < 8; r++) e.q[r] = r < RT ? prm.rq[(size_t)(r < RT ? r : 0) * pods.NQ + rq] : 0;
    e.tol = pods.sig_tol[pods.sig_id[p]];
    e.p = p;
    e.tnew = A.a_tmpl[p];
  }
  __syncthreads();

  // ---- per-thread claim state: claim c = j * REG_NCT + tid lives in slot j ----
  Slot sl0, sl1, sl2, sl3;
  slot_clear(sl0); slot_clear(sl1); slot_clear(sl2); slot_clear(sl3);
// F(j, slot) for every slot this form has
#define FOR_SLOTS(F)                  \
  do {                                \
    F(0, sl0);                         \
    if constexpr (CPT > 1) F(1, sl1);  \
    if constexpr (CPT > 2) F(2, sl2);  \
    if constexpr (CPT > 3) F(3, sl3);  \
  } while (0)
  int n = 0, seq = 0;  // uniform across the CTA
  // queue warp: the batch being prepared (one pod per lane), staged over the steps of a batch
  int pf_p = -1, pf_rq = 0, pf_sig = 0, pf_tnew = -1;
  int32_t pf_q[RT];
  uint64_t pf_tol = 0;
#pragma unroll
  for (int r = 0; r < RT; r++) pf_q[r] = 0;
  bool overflow = false;
  unsigned long long prof[7] = {0, 0, 0, 0, 0, 0, 0};
  long long tick = PROF ? clock64() : 0;
#define RPROF(i)                                    \
  do {                                              \
    if (PROF && (tid & 31) == 0) {                  \
      const long long now_ = clock64();             \
      prof[i] += (unsigned long long)(now_ - tick); \
      tick = now_;                                  \
    }                                               \
  } while (0)

  auto pop_slow = [&](int t) {  // second phase (one warp): Queue.Pop with the lastLen rule over the pushed pods
    int p = -1;
    if (lane == 0 && S.len > 0) {
      const int cand = push_list[S.head];
      if (A.lastlen[cand] != S.len) {
        p = cand;
        S.head = S.head + 1 == n_pods ? 0 : S.head + 1;
        S.len--;
      }
    }
    p = __shfl_sync(0xffffffffu, p, 0);
    RingEnt& e = S.ring[t & (RING - 1)];
    if (p >= 0) {
      const int rq = pods.req_id[p];
      if (lane < 8) e.q[lane] = lane < RT ? prm.rq[(size_t)lane * pods.NQ + rq] : 0;
      if (lane == 0) { e.tol = pods.sig_tol[pods.sig_id[p]]; e.p = p; e.tnew = A.a_tmpl[p]; }
    } else if (lane == 0) {
      e.p = -1;
    }
    __syncwarp();
  };

  // queue warp: writes out what step t did (see RegShared::wlog)
  auto qw_log = [&](int t) {
    const int w = S.wlog[t & (RING - 1)];
    __syncwarp();  // every lane has read the entry before lane 0 clears it
    if (w != -1) {
      if (lane == 0) {
        if (w >= 0) {
          const int p = S.ring[t & (RING - 1)].p;
          A.pod_kind[p] = B200S_PLACE_CLAIM;
          A.pod_index[p] = w;
          A.pod_seq[p] = seq;
        }
        S.wlog[t & (RING - 1)] = -1;
      }
      seq++;
    }
  };
  unsigned t_end = 0;  // steps taken (claim warps)
  if (warp == REG_ORDW) {
    // ================= the order warp: runs sort.Slice when the step loop asks for it =================
    for (;;) {
      meet_bar();  // A: the claim threads have written the order as it was before the change, and the new pod count
      if (S.done) break;
      RPROF(0);
      const int ma = S.meet_a;  // [11:0] changed position, [12] appended, [31:16] claims
      const bool gen = ord_normalise(s_ord, s_key, s_pos, ma >> 16, ma & 0xFFF, (ma & 0x1000) != 0);
      if (PROF && gen) prof[4]++;
      RPROF(1);
      meet_bar();  // B: positions are ready
    }
  } else if (warp == REG_QW) {
    // ================= the queue warp: pod rows of the next batch (the ring's other half), a stage every few steps ====
    for (int t = 0; t < n_pods; t++) {
      const int ph = t & 31, i = (t & ~31) + 32 + lane;
      if (i < n_pods) {
        if (ph == 0) pf_p = queue[i];
        else if (ph == 4) { pf_rq = pods.req_id[pf_p]; pf_sig = pods.sig_id[pf_p]; pf_tnew = A.a_tmpl[pf_p]; }
        else if (ph == 8) {
#pragma unroll
          for (int r = 0; r < RT; r++) pf_q[r] = prm.rq[(size_t)r * pods.NQ + pf_rq];
          pf_tol = pods.sig_tol[pf_sig];
        } else if (ph == 16) {
          RingEnt& d = S.ring[i & (RING - 1)];  // the slot of step i - RING, consumed before this batch began
#pragma unroll
          for (int r = 0; r < RT; r++) d.q[r] = pf_q[r];
#pragma unroll
          for (int r = RT; r < 8; r++) d.q[r] = 0;
          d.tol = pf_tol; d.p = pf_p; d.tnew = pf_tnew;
        }
      }
      if (REG_DEFER && t > 1) qw_log(t - 2);  // this iteration runs beside step t - 1: the log of step t - 2 is complete
      RPROF(0);
      step_bar();
    }
    if (REG_DEFER && n_pods > 1) qw_log(n_pods - 2);  // (the last step's winner writes its placement itself)
  } else {
    // ================= claim warps: the step loop =================
    int s3i = 0;  // t % 3
    int32_t nx_q[RT];
    uint64_t nx_tol = 0;
    int nx_p = -1, nx_tnew = -1;
#pragma unroll
    for (int r = 0; r < RT; r++) nx_q[r] = 0;
    // one pod: returns false when the queue is exhausted (second phase) or the claims do not fit this form
    auto step = [&](const unsigned t, auto phase2_tag) -> bool {
      constexpr bool PHASE2 = decltype(phase2_tag)::value;
      // the pod's row: read during the step before in the first pass (the rows are written a batch ahead)
      int32_t q[RT];
      uint64_t tol;
      int ep, etnew;
      if (PHASE2 || t == 0 || !REG_PREFETCH) {
        const RingEnt& e = S.ring[t & (RING - 1)];
        tol = e.tol; ep = e.p; etnew = e.tnew;
#pragma unroll
        for (int r = 0; r < RT; r++) q[r] = e.q[r];
      } else {
        tol = nx_tol; ep = nx_p; etnew = nx_tnew;
#pragma unroll
        for (int r = 0; r < RT; r++) q[r] = nx_q[r];
      }
      if (PHASE2 && ep < 0) return false;
      const int s3 = s3i, s3n = s3i == 2 ? 0 : s3i + 1;
      s3i = s3n;

      // ---- before the barrier: addToInflightNode — every claim decides for itself, the first feasible in the order wins
      if (tid == 0) { S.best[s3n] = PRIO_INF; S.best2[s3n] = PRIO_INF; }
      // slot 0, or slots 0 and 1 together once a second slot is in use anywhere (straight-line code: the two evaluations
      // interleave; a slot without a claim has priority PRIO_INF), the others only in warps that hold such claims
      unsigned cand = PRIO_INF;
      // Guards are CTA-wide (is the slot in use anywhere), not per warp: with per-warp guards the warps drift apart
      // and the step got slower (6.93 against 6.66 ms in one run)
      const int wbase = 0;
      if (CPT > 1 && REG_NCT + wbase < n) {  // (uniform) some thread holds a second claim
        cand = slot_eval<NWIN>(sl0, prm, s_tab, s_win, s_bits, q, tol);
        if constexpr (CPT > 1) cand = min(cand, slot_eval<NWIN>(sl1, prm, s_tab, s_win, s_bits, q, tol));
      } else if (wbase < n) {  // only first claims
        cand = slot_eval<NWIN>(sl0, prm, s_tab, s_win, s_bits, q, tol);
      }
#define EVAL_SLOT(j, sl)                                                               \
  if ((j) > 1 && (j) * REG_NCT + (tid & ~31) < n) /* this warp holds a claim in slot j */ \
    cand = min(cand, slot_eval<NWIN>(sl, prm, s_tab, s_win, s_bits, q, tol))
      FOR_SLOTS(EVAL_SLOT);
#undef EVAL_SLOT
      const unsigned wmin = __reduce_min_sync(0xffffffffu, cand);
      lane0_min(&S.best[s3], wmin, lane);
      if (!PHASE2 && REG_PREFETCH) {  // the next pod's row
        const RingEnt& e1 = S.ring[(t + 1u) & (RING - 1)];
        nx_tol = e1.tol; nx_p = e1.p; nx_tnew = e1.tnew;
#pragma unroll
        for (int r = 0; r < RT; r++) nx_q[r] = e1.q[r];
      }
      RPROF(0);
      if (PHASE2) claim_bar(); else step_bar();  // B1
      RPROF(1);
      // ---- the previous step's change of the order: the sort at the start of this add() (scheduler.go:499) ----
      // The minimum above was taken over priorities that do not have it yet: every claim but the one that changed keeps
      // its place among the others, and that one took a priority between its new neighbours' old ones.
      const unsigned mv = S.mv[(t + 1u) & 1u];
      unsigned g = S.best[s3];
      if (mv & MV_MEET) {  // uniform
        // sort.Slice may take the real pdqsort path (ties permuted): the claim threads lay out the order as it is —
        // positions before the change, the new pod count —, the order warp sorts it, positions are re-read and the
        // minimum is taken again.
#define LAYOUT_SLOT(j, sl)                                 \
  if ((j) * REG_NCT + tid < n) {                           \
    s_ord[sl.pos] = (j) * REG_NCT + tid;                   \
    s_key[sl.pos] = sl.npods;                              \
    s_pos[(j) * REG_NCT + tid] = (int)sl.pos;              \
  }
        FOR_SLOTS(LAYOUT_SLOT);
#undef LAYOUT_SLOT
        meet_bar();  // A
        meet_bar();  // B
        cand = PRIO_INF;
#define RELOAD_SLOT(j, sl)                               \
  if ((j) * REG_NCT + tid < n) {                         \
    sl.pos = (unsigned)s_pos[(j) * REG_NCT + tid];       \
    sl.prio = 2u * sl.pos + 1u;                          \
    sl.cur = sl.cur != PRIO_INF ? sl.prio : PRIO_INF;    \
    cand = min(cand, sl.cur);                            \
  }
        FOR_SLOTS(RELOAD_SLOT);
#undef RELOAD_SLOT
        const unsigned wmin2 = __reduce_min_sync(0xffffffffu, cand);
        lane0_min(&S.best2[s3], wmin2, lane);
        claim_bar();  // B2
        g = S.best2[s3];
        RPROF(5);
        if (PROF) prof[6]++;
      }
      // the winner is the slot whose priority entered the minimum as g; then every slot takes the previous step's shift
      // (the claim that moved marked its position: bit 31) and its priority for the next minimum
      int mine = -1;
      {
        const bool shift = (mv & MV_MEET) == 0;
        const unsigned lo = mv & 0xFFFu, span = shift ? ((mv >> 12) & 0xFFFu) - lo : 0u;
        const unsigned d = (mv & MV_UP) ? 1u : 0xFFFFFFFFu;
#define SHIFT_SLOT(j, sl)                                     \
  if ((j) * REG_NCT + wbase < n) { /* uniform: the slot is in use somewhere */ \
    if (sl.cur == g) mine = (j);                              \
    if (sl.pos - lo <= span && shift) sl.pos += d;            \
    sl.pos &= 0x7FFFFFFFu;                                    \
    sl.prio = 2u * sl.pos + 1u;                               \
  }
        FOR_SLOTS(SHIFT_SLOT);
#undef SHIFT_SLOT
      }
      RPROF(3);

      if (g != PRIO_INF) {
        // ---- NodeClaim.Add (nodeclaim.go:168-180) by the owner ----
        if (mine >= 0) {
          int a = 0, k = 0;
#define WIN_SLOT(j, sl) \
  if (mine == (j)) { a = (int)sl.pos; k = sl.npods; }
          FOR_SLOTS(WIN_SLOT);
#undef WIN_SLOT
          // The claim now has k + 1 pods and belongs at the end of the run of claims with k pods, position dst (the
          // claims it passes move down by one).  sort.Slice gets there by partialInsertionSort unless there are fewer
          // than 50 claims or choosePivot's nine samples, which include the positions next to a, show an inversion; the
          // keys are not here, so that is left to the order warp at the next step's sort.  Only winners touch the count
          // table Gc[x] = claims with more than x pods, one per step, a barrier apart.
          const bool big = k + 1 >= REG_KCAP;
          const int kk = big ? REG_KCAP - 1 : k;
          const int gk = s_Gc[kk];
          if (!big) s_Gc[kk] = gk + 1;
          int dst = n - gk - 1;
          const int qq = n >> 2;
          const bool near = (unsigned)(a - qq + 1) <= 2u || (unsigned)(a - 2 * qq + 1) <= 2u || (unsigned)(a - 3 * qq + 1) <= 2u;
          const bool meet = big || (dst != a && (n < 50 || near));
          if (big) dst = a;
          S.mv[t & 1u] = mv_word(a + 1, dst, false, meet);
          if (meet) S.meet_a = a | (n << 16);
#define COMMIT_SLOT(j, sl)                                                   \
  if (mine == (j)) {                                                         \
    _Pragma("unroll") for (int r = 0; r < RT; r++) sl.x[r] += q[r];          \
    sl.wrk = sl.nwrk;                                                        \
    sl.npods = k + 1;                                                        \
    sl.prio = 2u * (unsigned)dst + 2u;                                       \
    sl.pos = meet ? (unsigned)a : (0x80000000u | (unsigned)dst);             \
  }
          FOR_SLOTS(COMMIT_SLOT);
#undef COMMIT_SLOT
          if (PHASE2 || !REG_DEFER || (int)t == n_pods - 1) {  // (the queue warp's last look at the log is one step earlier)
            const int p = ep;
            A.pod_kind[p] = B200S_PLACE_CLAIM;
            A.pod_index[p] = mine * REG_NCT + tid;
            A.pod_seq[p] = seq;
          } else {
            S.wlog[t & (RING - 1)] = mine * REG_NCT + tid;  // the queue warp writes it out
          }
        }
        seq++;
      } else {
        const int tnew = etnew, p = ep;
        if (tnew >= 0) {
          // ---- addToNewNodeClaim: the template walk was done for every pod by stage A (scheduler.go:587-675) ----
          if (n >= prm.cap) { overflow = true; return false; }  // uniform: the next form redoes this problem
          const int c = n;
          if (tid == c % REG_NCT) {
            // appended (scheduler.go:672); the next sort puts it behind the last claim with one pod — below 50 claims
            // by the real pdqsort on any unsorted input, from 50 on always by re-insertion (the last position is no
            // sample)
            const int dst = n - s_Gc[1];
            const bool meet = n + 1 < 50 && dst != n;
#define OPEN_SLOT(j, sl) \
  if ((j) == c / REG_NCT) slot_open<NWIN>(sl, prm, s_win, S.daemon[tnew], q, tnew, S.taints[tnew], meet ? (unsigned)n : (0x80000000u | (unsigned)dst), dst)
            FOR_SLOTS(OPEN_SLOT);
#undef OPEN_SLOT
            S.mv[t & 1u] = mv_word(dst, n - 1, true, meet);
            S.meet_a = n | 0x1000 | ((n + 1) << 16);
            if (!PHASE2 && REG_DEFER && (int)t != n_pods - 1) S.wlog[t & (RING - 1)] = -2;
            c_first[c] = p;
            A.pod_kind[p] = B200S_PLACE_CLAIM;
            A.pod_index[p] = c;
            A.pod_seq[p] = seq;
          }
          n++;
          seq++;
        } else if (tid == 0) {
          // ---- Queue.Push (queue.go:63-66): the pod stays in PodErrors ----
          if (!PHASE2) {
            // the reference's queue holds the n_pods - (t + 1) pods not yet popped and the pods pushed so far
            push_list[S.n_pushed] = p;
            A.lastlen[p] = (n_pods - ((int)t + 1)) + S.n_pushed + 1;
            S.n_pushed++;
          } else {
            int tail = S.head + S.len;
            if (tail >= n_pods) tail -= n_pods;
            push_list[tail] = p;
            S.len++;
            A.lastlen[p] = S.len;
          }
          A.pod_kind[p] = B200S_PLACE_UNSCHEDULED;
          A.pod_index[p] = -1;
          S.mv[t & 1u] = MV_EMPTY;
        }
      }
      if (PHASE2) {  // the next pod: Queue.Pop with the lastLen rule, by the warp that pushed
        if (warp == 0) pop_slow((int)t + 1);
        claim_bar();
      }
      RPROF(2);
      if (PROF && tid == 0) prof[4]++;
      return true;
    };

    unsigned t = 0;
    for (; (int)t < n_pods; t++)
      if (!step(t, std::false_type{})) break;
    if (overflow) {
      for (int t2 = (int)t + 1; t2 < n_pods; t2++) step_bar();  // the queue warp's remaining barriers
    } else {
      // the first pass over the queue is over: what is left are the pods that were pushed back
      claim_bar();
      if (S.n_pushed != 0) {
        if (warp == 0) {
          if (lane == 0) { S.head = 0; S.len = S.n_pushed; }
          __syncwarp();
          pop_slow((int)t);
        }
        claim_bar();
        for (;; t++)
          if (!step(t, std::true_type{})) break;
      }
    }
    t_end = t;
    // Results.NewNodeClaims is the order after the last sort the reference ran plus what the last add() did to it: the
    // last step's change is not applied (its claim keeps its place; a new claim is last)
    if (tid == 0) { S.done = overflow ? 2 : 1; S.meet_n = n; }
    meet_bar();
  }

  // ---- epilogue ----
  __syncthreads();
  overflow = S.done == 2;
  n = S.meet_n;
  if (overflow) {  // status stays 0: the next form redoes the problem from a clean queue state
    for (int i = tid; i < n_pods; i += REG_THREADS) A.lastlen[base + i] = 0;
    return;
  }
  {
    const int R = prm.R, W = prm.W, K = prm.K, NB = prm.NB;
    int64_t* c_req = A.c_req + (size_t)base * R;
    int64_t* c_rk = A.c_thr + (size_t)base * R;  // scratch: the bit row of each of a claim's resources' final totals
    uint64_t* c_mask = A.c_mask + (size_t)base * W;
    uint64_t* c_rmask = A.c_rmask + (size_t)base * K;
    uint64_t* c_rflags = A.c_rflags + base;
    int64_t* c_gt = A.c_gt + (size_t)base * NB;
    int64_t* c_lt = A.c_lt + (size_t)base * NB;
    int32_t* c_tmpl = A.c_tmpl + base;
    int32_t* c_npods = A.c_npods + base;
    int32_t* g_ord = A.ord + base;
    if (is_claim) {
      // the claim the very last step changed marked its position (bit 31) with where the next sort would have put it:
      // a new claim is last, a claim that gained a pod stays where it was (one before the range it would have passed)
      const unsigned last_mv = t_end > 0 ? S.mv[(t_end + 1u) & 1u] : MV_EMPTY;
#define ORDER_SLOT(j, sl)                                                                                      \
  if ((j) * REG_NCT + tid < n) {                                                                               \
    if (sl.pos & 0x80000000u) {                                                                                \
      if (last_mv & MV_UP) sl.pos = (unsigned)(n - 1);                                                         \
      else sl.pos = (last_mv & 0xFFFFFFu) == MV_EMPTY ? (sl.pos & 0x7FFFFFFFu) : (last_mv & 0xFFFu) - 1u;       \
    }                                                                                                          \
    g_ord[sl.pos] = (j) * REG_NCT + tid;                                                                       \
  }
      FOR_SLOTS(ORDER_SLOT);
#undef ORDER_SLOT
    }
    if (is_claim) {
#define STORE_SLOT(j, sl)                                                                    \
  if ((j) * REG_NCT + tid < n) {                                                             \
    const int c = (j) * REG_NCT + tid;                                                       \
    _Pragma("unroll") for (int r = 0; r < RT; r++) if (prm.perm[r] >= 0) {                   \
      int rank;                                                                              \
      if (NWIN == RT) rank = (int)((sl.wrk >> (8 * r)) & 0xFFull);                            \
      else if (r < RT - 1) rank = s_tab[prm.toff[r] + min(sl.x[r], prm.lim[r])];             \
      else {                                                                                 \
        const int2 b = reinterpret_cast<const int2*>(s_win)[prm.toff[r] + (sl.x[r] >> prm.bsh)]; \
        rank = b.y + (b.x < sl.x[r] ? 1 : 0);                                                \
      }                                                                                      \
      c_req[(size_t)c * R + prm.perm[r]] = (int64_t)sl.x[r] * prm.scale[r];                  \
      c_rk[(size_t)c * R + prm.perm[r]] = (int64_t)(prm.row0[r] + rank);                     \
    }                                                                                        \
    c_tmpl[c] = sl.tmpl;                                                                     \
    c_npods[c] = sl.npods;                                                       \
  }
      FOR_SLOTS(STORE_SLOT);
#undef STORE_SLOT
    }
    __syncthreads();
    // InstanceTypeOptions and requirements of every claim, a warp per claim: the first pod's stage-A row inside the
    // rows of the final totals (nested rows: the AND of every row the claim moved through is the last one); a pod
    // without requirement keys leaves the template's requirements as they are (requirements.go:127-134)
    for (int c = warp; c < n; c += REG_NW) {
      const int first = c_first[c], t = c_tmpl[c];
      for (int w = lane; w < W; w += 32) {
        uint64_t m = stage_a_word(A, pods, first, t, W, w);
        for (int r = 0; r < R; r++) m &= prm.rows[(size_t)c_rk[(size_t)c * R + r] * W + w];
        c_mask[(size_t)c * W + w] = m;
      }
      for (int k = lane; k < K; k += 32) {
        c_rmask[(size_t)c * K + k] = tg->rmask[t][k];
        const int b = prm.bslot[k];
        if (b >= 0) {
          const uint32_t nib = (uint32_t)((tg->rflags[t] >> (4 * k)) & 0xFu);
          c_gt[(size_t)c * NB + b] = (nib & RF_GT) ? tg->gt[t][b] : 0;
          c_lt[(size_t)c * NB + b] = (nib & RF_LT) ? tg->lt[t][b] : 0;
        }
      }
      if (lane == 0) c_rflags[c] = tg->rflags[t];
    }
  }
  if (tid == 0) {
    pb.n_claims = n;
    pb.n_unscheduled = S.len;
    pb.evals = 0;
    pb.status = 1;
    if (PROF) { pb.prof[0] = prof[0]; pb.prof[1] = prof[1]; pb.prof[2] = prof[2]; pb.prof[3] = prof[4]; pb.prof[7] = prof[3]; pb.prof[8] = prof[5]; pb.prof[9] = prof[6]; }
  }
  if (PROF && warp == REG_ORDW && lane == 0) { pb.prof[4] = prof[0]; pb.prof[5] = prof[1]; pb.prof[6] = 0; pb.prof[10] = prof[4]; }
  if (PROF && warp == REG_QW && lane == 0) { pb.prof[11] = prof[0]; }
}

// Feasibility table of template t: bit [sum_r i_r * stride_r] = some instance type of the template's pre-filtered
// set (NewScheduler, scheduler.go:143-158) with no negative allocatable entry (resources.go:234-238) has
// Allocatable()[r] >= dv_r[i_r] for every resource r — i.e. resources.Fits of a total whose per-resource ranks are
// i.  Rank nd_r (a total above every allocatable value) is a dimension of its own and always 0.
struct FeasParams {
  int32_t cells, wpt, T, W, row_nonneg;
  int32_t nd[RT], row0[RT], stride[RT];
};

__global__ void __launch_bounds__(256) feas_table_kernel(FeasParams fp, const uint64_t* __restrict__ rows,
                                                         const uint64_t* __restrict__ tmask, uint32_t* __restrict__ bits) {
  const int t = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  bool any = false;
  if (cell < fp.cells) {
    int rowoff[RT];
    bool none = false;
#pragma unroll
    for (int r = 0; r < RT; r++) {
      rowoff[r] = -1;
      if (fp.stride[r] > 0) {
        const int i = (cell / fp.stride[r]) % (fp.nd[r] + 1);
        if (i == fp.nd[r]) none = true;
        rowoff[r] = (fp.row0[r] + i) * fp.W;
      }
    }
    for (int w = 0; w < fp.W && !any && !none; w++) {
      uint64_t m = tmask[(size_t)t * fp.W + w] & ldg(rows + (size_t)fp.row_nonneg * fp.W + w);
#pragma unroll
      for (int r = 0; r < RT; r++)
        if (rowoff[r] >= 0) m &= ldg(rows + rowoff[r] + w);
      any = m != 0;
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, any);
  if ((threadIdx.x & 31) == 0 && (cell >> 5) < fp.wpt) bits[(size_t)t * fp.wpt + (cell >> 5)] = b;
}

}  // namespace

// Plans the register-resident pack for one resident problem set: granules, table kinds and shapes, and the host
// copies of the small tables.  h_dv: the catalog's distinct allocatable values (host copy); rq[b]: problem b's request
// classes [R * nq[b]], resource-major.
bool reg_pack_plan(const KeyTab& kt, const int64_t* h_dv, int T, const int64_t (*daemon)[MAXR], int n_problems,
                   const int64_t* const* rq, const int32_t* nq, int smem_limit, RegPackPlan* plan) {
  plan->eligible = false;
  const int R = kt.n_res;
  if (T < 1 || R < 1 || R > RT || !h_dv) return false;
  int64_t g[RT], maxv[RT];
  bool direct[RT];
  size_t tab_bytes = 0;
  for (int r = 0; r < R; r++) {
    const int nd = kt.res_nd[r];
    if (nd < 1 || nd > 250) return false;
    const int64_t* dv = h_dv + kt.res_dv0[r];
    if (dv[0] < 0) return false;  // negative allocatable values shift every rank: pack_kernel's problem
    int64_t gg = 0;
    for (int i = 0; i < nd; i++) gg = std::gcd(gg, dv[i]);
    for (int t = 0; t < T; t++) {
      if (daemon[t][r] < 0) return false;
      gg = std::gcd(gg, daemon[t][r]);
    }
    for (int b = 0; b < n_problems; b++)
      for (int c = 0; c < nq[b]; c++) {
        const int64_t v = rq[b][(size_t)r * nq[b] + c];
        if (v < 0) return false;
        gg = std::gcd(gg, v);
      }
    if (gg == 0) gg = 1;
    g[r] = gg;
    maxv[r] = dv[nd - 1] / gg;
    if (maxv[r] >= (1ll << 29)) return false;
    // a direct table covers every total a claim can be asked about: its requests (<= maxv) plus a pod's (clamped to
    // maxv + 1), so the lookup needs no bound check
    direct[r] = 2 * maxv[r] + 2 <= 40960 && tab_bytes + (size_t)(2 * maxv[r] + 2) <= 65536;
    if (direct[r]) tab_bytes += (size_t)(2 * maxv[r] + 2);
  }
  // Compiled forms: (1) direct tables in slots 0 .. RT-2 and a bucket table in the last slot — for the one resource
  // whose scaled range is too large for a direct table (memory, typically), if buckets of a power-of-two width
  // separate its distinct values; (RT) every slot tracks its rank through windows.
  int nwin = 0, last = -1;
  for (int r = 0; r < R; r++)
    if (!direct[r]) { nwin++; last = r; }
  if (nwin == 0 && R == RT) last = R - 1;
  int bsh = 31;
  bool bucket_ok = nwin <= 1;
  if (bucket_ok && last >= 0) {
    const int64_t* dv = h_dv + kt.res_dv0[last];
    const int nd = kt.res_nd[last];
    bsh = 30;
    for (; bsh >= 0; bsh--) {
      bool ok = true;
      for (int i = 0; i + 1 < nd && ok; i++) ok = ((dv[i] / g[last]) >> bsh) != ((dv[i + 1] / g[last]) >> bsh);
      if (ok) break;
    }
    bucket_ok = bsh >= 0 && ((2 * maxv[last] + 1) >> bsh) + 1 <= 4096;
  }
  const int form = bucket_ok ? 1 : RT;
  int slot_of[RT];
  {
    int d = 0, w = 0;
    for (int r = 0; r < R; r++) {
      if (form == RT) slot_of[r] = w++;
      else if (r == last) slot_of[r] = RT - 1;
      else slot_of[r] = d++;
    }
  }
  plan->bsh = bsh;
  plan->nwin_form = form;
  plan->R = R;
  plan->tab.clear();
  plan->win.clear();
  for (int s = 0; s < RT; s++) {
    plan->perm[s] = -1; plan->lim[s] = 0; plan->toff[s] = 0; plan->stride[s] = 0; plan->row0[s] = 0; plan->scale[s] = 1;
    plan->nd[s] = 0; plan->maxv[s] = 0;
  }
  for (int r = 0; r < R; r++) {
    const int s = slot_of[r];
    plan->perm[s] = r;
    plan->scale[s] = g[r];
    plan->nd[s] = kt.res_nd[r];
    plan->row0[s] = kt.res_row0[r];
    plan->maxv[s] = (int32_t)maxv[r];
  }
  long long cells = 1;
  for (int s = 0; s < RT; s++) {
    if (plan->perm[s] < 0) continue;
    plan->stride[s] = (int32_t)cells;
    cells *= plan->nd[s] + 1;
    if (cells > (1ll << 22)) return false;
  }
  plan->cells = (int32_t)cells;
  plan->wpt = (int32_t)((cells + 31) / 32);
  for (int s = 0; s < RT; s++) {
    const bool is_window = form == RT, is_bucket = form == 1 && s == RT - 1;
    const int r = plan->perm[s];
    if (is_bucket) {
      plan->toff[s] = 0;
      if (r < 0) {  // padding: one bucket, rank 0 (bsh = 31)
        plan->win.push_back(INT32_MAX);
        plan->win.push_back(0);
        continue;
      }
      const int64_t* dv = h_dv + kt.res_dv0[r];
      const int nd = kt.res_nd[r];
      const int64_t nb = ((2 * maxv[r] + 1) >> bsh) + 1;
      int rank = 0;
      for (int64_t b = 0; b < nb; b++) {  // {the distinct value inside the bucket, number of distinct values below it}
        const int64_t lo = b << bsh, hi = lo + (1ll << bsh);
        while (rank < nd && dv[rank] / g[r] < lo) rank++;
        const bool in = rank < nd && dv[rank] / g[r] < hi;
        plan->win.push_back(in ? (int32_t)(dv[rank] / g[r]) : INT32_MAX);
        plan->win.push_back(rank);
      }
    } else if (!is_window) {
      plan->toff[s] = (int32_t)plan->tab.size();
      if (r < 0) {  // padding: one entry, rank 0
        plan->lim[s] = 0;
        plan->tab.push_back(0);
        continue;
      }
      const int64_t* dv = h_dv + kt.res_dv0[r];
      const int nd = kt.res_nd[r];
      plan->lim[s] = (int32_t)(2 * maxv[r] + 1);
      int rank = 0;
      for (int64_t x = 0; x <= 2 * maxv[r] + 1; x++) {  // rank(x) = number of distinct values below x
        while (rank < nd && dv[rank] / g[r] < x) rank++;
        plan->tab.push_back((uint8_t)rank);
      }
    } else {
      plan->toff[s] = (int32_t)(plan->win.size() / 4);
      const int nd = r >= 0 ? kt.res_nd[r] : 0;
      const int64_t* dv = r >= 0 ? h_dv + kt.res_dv0[r] : nullptr;
      for (int i = 0; i <= nd; i++)
        for (int k = 0; k < 4; k++) plan->win.push_back(i + k < nd ? (int32_t)(dv[i + k] / g[r]) : INT32_MAX);
    }
  }
  while (plan->tab.size() % 16) plan->tab.push_back(0);
  while (plan->win.size() % 4) plan->win.push_back(0);
  plan->daemon32.assign((size_t)MAXT * 8, 0);
  for (int t = 0; t < T; t++)
    for (int s = 0; s < RT; s++)
      if (plan->perm[s] >= 0) {
        const int r = plan->perm[s];
        plan->daemon32[(size_t)t * 8 + s] = (int32_t)std::min<int64_t>(daemon[t][r] / g[r], maxv[r] + 1);
      }
  // claims per thread: the largest compiled form shared memory allows
  for (int cpt : {4, 2, 1}) {
    const int cap = REG_NCT * cpt;
    if (reg_layout((int)plan->tab.size(), (int)(plan->win.size() / 4), T, plan->wpt, cap).total <= (size_t)smem_limit) {
      plan->cpt = cpt;
      plan->eligible = true;
      return true;
    }
  }
  return false;
}

// request classes in granules by slot: out[s * out_stride + out_off + class], clamped to one past the largest
// allocatable value (a pod that asks for more than any instance type has fits nowhere either way)
void reg_pack_scale_requests(const RegPackPlan& plan, const int64_t* rq, int nq, int32_t* out, size_t out_stride,
                             size_t out_off) {
  for (int s = 0; s < RT; s++)
    for (int c = 0; c < nq; c++) {
      int32_t v = 0;
      if (plan.perm[s] >= 0)
        v = (int32_t)std::min<int64_t>(rq[(size_t)plan.perm[s] * nq + c] / plan.scale[s], (int64_t)plan.maxv[s] + 1);
      out[(size_t)s * out_stride + out_off + c] = v;
    }
}

cudaError_t launch_feas_table(const KeyTab& kt, const uint64_t* tmask, int T, const RegPackPlan& plan, uint32_t* bits,
                              cudaStream_t stream) {
  FeasParams fp{};
  fp.cells = plan.cells; fp.wpt = plan.wpt; fp.T = T; fp.W = kt.n_words; fp.row_nonneg = kt.row_nonneg;
  for (int s = 0; s < RT; s++) { fp.nd[s] = plan.nd[s]; fp.row0[s] = plan.row0[s]; fp.stride[s] = plan.stride[s]; }
  dim3 grid((unsigned)((plan.wpt * 32 + 255) / 256), (unsigned)T);
  feas_table_kernel<<<grid, 256, 0, stream>>>(fp, kt.rows, tmask, bits);
  return cudaGetLastError();
}

cudaError_t launch_pack_reg(const KeyTab& kt, const TmplDev* tg, int T, const PodsDev& pods, PackArrays arr,
                            PackProblem* problems, int n_problems, int max_pods, const RegPackPlan& plan,
                            const RegPackDev& dev, bool* may_overflow, int* n_launched, cudaStream_t stream) {
  *may_overflow = false;
  if (n_launched) *n_launched = 0;
  if (n_problems == 0) return cudaSuccess;
  RegParams prm{};
  prm.R = kt.n_res; prm.W = kt.n_words; prm.K = kt.n_keys; prm.NB = kt.n_bounded; prm.T = T;
  if (max_pods >= (1 << 25)) {  // a claim's pod count travels in 25 bits of an event word
    *may_overflow = true;
    return cudaSuccess;
  }
  prm.wpt = plan.wpt;
  prm.tab_bytes = (int32_t)plan.tab.size();
  prm.win_entries = (int32_t)(plan.win.size() / 4);
  for (int s = 0; s < RT; s++) {
    prm.bsh = plan.bsh;
    prm.perm[s] = plan.perm[s]; prm.lim[s] = plan.lim[s]; prm.toff[s] = plan.toff[s]; prm.stride[s] = plan.stride[s];
    prm.row0[s] = plan.row0[s]; prm.scale[s] = plan.scale[s];
  }
  prm.tab = dev.tab; prm.win = reinterpret_cast<const int4*>(dev.win); prm.bits = dev.bits; prm.daemon = dev.daemon;
  prm.rq = dev.rq; prm.rows = kt.rows;
  for (int k = 0; k < MAXK; k++) prm.bslot[k] = kt.bslot[k];
  cudaError_t err = cudaSuccess;
  const bool prof = getenv("B200S_REG_PROFILE") != nullptr;
#define REG_CALL(NWIN_, CPT_)                                                                     \
  do {                                                                                            \
    auto kern = prof ? pack_reg_kernel<NWIN_, CPT_, true> : pack_reg_kernel<NWIN_, CPT_, false>; \
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);    \
    if (err != cudaSuccess) return err;                                                           \
    kern<<<n_problems, REG_THREADS, bytes, stream>>>(prm, tg, pods, arr, problems);               \
  } while (0)
  // Claims per thread: the smallest form first (less code on the step's critical path); a problem that needs more
  // claims than it holds leaves status 0 and the next form — then pack_kernel — redoes it, each skipping what an earlier
  // one finished.  A form is launched only if some problem could still need it; batches of large problems do not try
  // the small forms (what they do before giving up is wasted).
  const char* force = getenv("B200S_REG_CPT");  // experiments: the first form to try
  // (the one-slot form only where it cannot run out: every claim holds at least one pod)
  const int first = force ? atoi(force) : (max_pods <= REG_NCT ? 1 : max_pods <= 65536 ? 2 : 4);
  int launched = 0, prev_cap = 0;
  for (int cpt : {1, 2, 4}) {
    if (cpt > plan.cpt) break;
    if (cpt < first && cpt < plan.cpt) continue;
    if (launched && prev_cap >= max_pods) break;  // the smaller form cannot have overflowed
    prm.cap = REG_NCT * cpt;
    prev_cap = prm.cap;
    const size_t bytes = reg_layout(prm.tab_bytes, prm.win_entries, T, plan.wpt, prm.cap).total;
    if (plan.nwin_form == 1) {
      if (cpt == 1) REG_CALL(1, 1);
      else if (cpt == 2) REG_CALL(1, 2);
      else REG_CALL(1, 4);
    } else {
      if (cpt == 1) REG_CALL(RT, 1);
      else if (cpt == 2) REG_CALL(RT, 2);
      else REG_CALL(RT, 4);
    }
    launched++;
    *may_overflow = prm.cap < max_pods;
    if (cudaGetLastError() != cudaSuccess) return cudaErrorLaunchFailure;
  }
#undef REG_CALL
  if (n_launched) *n_launched = launched;
  return cudaGetLastError();
}

}  // namespace b200

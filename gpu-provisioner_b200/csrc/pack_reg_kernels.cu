// pack_reg_kernels.cu — the register-resident form of the pack (Scheduler.Solve's sequential loop,
// K/controllers/provisioning/scheduling/scheduler.go:394-421, :488-675) for batches in which no pod defines a
// requirement key and there are no existing nodes: the common provisioning batch, and BASELINE config[1].
//
// pack_kernels.cu walks s.newNodeClaims in order, lane group per claim, claim state in global memory.  A step there
// is a chain of dependent loads (order -> claim state -> tables -> type mask) and several CTA barriers.  On this
// machine a warp's dependent instruction stream runs at 4-5 cycles per instruction whatever it does, so a step costs
// what its longest role ISSUES; this form is built to issue little:
//
//   * every in-flight NodeClaim is owned by one thread and lives in its registers: the accumulated requests as
//     32-bit multiples of each resource's granule (the gcd of every quantity of that resource in the problem — exact;
//     a problem whose scaled values do not fit falls back to pack_kernel);
//   * the claim ORDER is never read by the scan.  Every claim carries a priority (len(Pods), tie-break) such that
//     "first feasible claim in s.newNodeClaims order" (scheduler.go:566-575) == "feasible claim of least priority":
//     one warp min-reduce + one shared-memory atomicMin + one CTA barrier per pod.  The priority is updated locally:
//     Go's sort.Slice (scheduler.go:499) puts a claim that gained a pod in front of the next bucket and a new claim
//     at the back of bucket 1 whenever it does not take the real pdqsort path; the ORDER WARP keeps the explicit
//     order beside the scan (exact pdqsort emulation, b200_core.cuh GoSort), tells the claim threads when a full
//     sort happened, and they then re-derive their tie-breaks from the explicit positions;
//   * a claim of selector-less pods keeps exactly those instance types of its template's pre-filtered set that fit
//     its requests (nodeclaim.go:144-146, :392-423; the rows "alloc[r] >= v" are nested), so CanAdd's filter result
//     is non-empty iff the FEASIBILITY TABLE cell of the new total's per-resource ranks is set: a bit table over
//     rank space, built once per Solve by feas_table_kernel from the bit rows.  The rank of a total is one byte load
//     from a direct value -> rank table (resources with a small scaled range) or one 16-byte window of the next
//     distinct allocatable values and four compares (the others).  No type mask is read or written in the loop; the
//     final InstanceTypeOptions are the first pod's stage-A row inside the rows of the final totals;
//   * the queue warp prepares pod rows 32 at a time, a batch ahead (Queue.Pop, queue.go:46-60: until the first
//     pass over the queue is over the pop order is the FFD order), and appends pushed-back pods (queue.go:63-66) to
//     a side list that the second phase then cycles through with the lastLen rule.
//
// The host twin (tests/native/emul.cpp emul_solve_reg) runs the same algorithm serially and is checked against the
// oracle on the CPU; the kernel is checked against the oracle in tests/test_gpu.py.  A problem that needs more claims
// than the threads hold leaves status 0 and pack_kernel (launched right after, skipping finished problems) redoes it.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <numeric>

#include "device_types.cuh"
#include "kernels.h"

namespace b200 {

namespace {

#ifndef B200S_REG_THREADS
#define B200S_REG_THREADS 512
#endif
constexpr int REG_THREADS = B200S_REG_THREADS;
constexpr int REG_NW = REG_THREADS / 32;
constexpr int REG_CW = REG_NW - 2;     // claim warps; then the order warp and the queue warp
constexpr int REG_NCT = REG_CW * 32;   // claim threads
constexpr int REG_ORDW = REG_CW, REG_QW = REG_CW + 1;
constexpr int RING = 64;
constexpr int RT = REG_RT;             // resource slots of a claim (catalog resources, padded)
constexpr unsigned PRIO_INF = 0xffffffffu;
constexpr int ORD_NONE = 0, ORD_ROTATE = 1, ORD_GENERIC = 2;

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
  unsigned win[3];  // the step's winner: claim id | len(Pods) before the pod << 11
  int32_t irr[2];
  int32_t head, len, n_pushed, _pad;
  RingEnt ring[RING];
};

__host__ __device__ inline size_t reg_align16(size_t x) { return (x + 15) & ~(size_t)15; }

constexpr int REG_KCAP = 2048;  // pod counts below this have an entry in the order warp's count table

struct RegLayout {
  size_t tab, win, bits, ord, key, pos, gt, total;
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
  l.gt = o; o += reg_align16((size_t)REG_KCAP * 4);
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
  for (int i = lane; i < n; i += 32) snap[i] = ord[i];
  __syncwarp();
  GoSort gs{ord, key, true};
  gs.sort(n);
  __syncwarp();
  for (int i = lane; i < n; i += 32) pos_of[ord[i]] = i;
  __syncwarp();
}

// The order warp's eager normalisation (see pack_kernels.cu warp_normalise for the derivation): what sort.Slice will
// do to an order that was sorted before one element changed.  Returns true when the full pdqsort ran.  Keeps pos_of
// (claim -> position) and how to take the normalisation back (Results.NewNodeClaims is un-normalised when no later
// add() reaches the sort).
struct OrdUndo {
  int mode, src, dst;
  // a decided but not yet executed move: the order warp executes it at the start of the NEXT step, while the claim
  // warps find that step's winner (pend: 0 none, 1 a claim that gained a pod passes the run behind it, 2 a new claim
  // walks down from the end)
  int pend, pa, pdst, pc, pk;
};

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

__device__ __forceinline__ bool ord_normalise(int32_t* ord, int32_t* key, int32_t* pos_of, int32_t* snap, int n, int pos,
                                              bool appended, OrdUndo& u) {
  const int lane = threadIdx.x & 31;
  u.mode = ORD_NONE;
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
    ord_full_sort(ord, key, pos_of, snap, n);
    u.mode = ORD_GENERIC;
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
  if (dst != pos) {
    ord_move(ord, key, pos_of, pos, dst);
    u.mode = ORD_ROTATE;
    u.src = dst;
    u.dst = pos;
  }
  return false;
}

// The same decisions from the count table G[k] = number of claims with more than k pods (k < REG_KCAP), without
// reading the order: a claim at position a that goes from k to k + 1 pods belongs at the end of the run of k's, position
// n - G[k] - 1; a new claim with one pod belongs behind the last claim with one pod, position n - G[1].  choosePivot
// can only see an inversion if the changed position is one of its nine samples, so the sample keys are read only
// then (never for an appended claim: the last position is not a sample).
__device__ __forceinline__ bool ord_win(int32_t* ord, int32_t* key, int32_t* pos_of, int32_t* G, int32_t* snap, int n, int c,
                                        int k, OrdUndo& u) {
  const int lane = threadIdx.x & 31;
  const int a = pos_of[c];
  if (k + 1 >= REG_KCAP) {  // beyond the table: the generic routine
    if (lane == 0) key[a] = k + 1;
    __syncwarp();
    return ord_normalise(ord, key, pos_of, snap, n, a, false, u);
  }
  const int gk = G[k];
  const int dst = n - gk - 1;
  __syncwarp();
  if (lane == 0) G[k] = gk + 1;
  u.mode = ORD_NONE;
  if (dst == a) {  // still sorted
    if (lane == 0) key[a] = k + 1;
    __syncwarp();
    return false;
  }
  bool generic = n < 50;
  if (!generic) {
    const int q = n >> 2;
    const int d1 = a - q, d2 = a - 2 * q, d3 = a - 3 * q;
    if ((d1 >= -1 && d1 <= 1) || (d2 >= -1 && d2 <= 1) || (d3 >= -1 && d3 <= 1)) {
      if (lane == 0) key[a] = k + 1;
      __syncwarp();
      int v = 0;
      if (lane < 9) v = key[q * (lane / 3 + 1) + (lane % 3) - 1];
      const int up = __shfl_down_sync(0xffffffffu, v, 1);
      const int mid_next = __shfl_down_sync(0xffffffffu, v, 3);
      bool bad = false;
      if (lane < 9 && (lane % 3) != 2) bad = up < v;
      if (lane == 1 || lane == 4) bad = bad || mid_next < v;
      generic = __ballot_sync(0xffffffffu, bad) != 0;
      __syncwarp();
      if (!generic && lane == 0) key[a] = k;  // the move below leaves a k at position a
    }
  }
  if (generic) {
    if (lane == 0) key[a] = k + 1;
    __syncwarp();
    ord_full_sort(ord, key, pos_of, snap, n);
    u.mode = ORD_GENERIC;
    return true;
  }
  // the claim passes the run of k's behind it (ord_apply, next step)
  u.pend = 1; u.pa = a; u.pdst = dst; u.pc = c; u.pk = k;
  u.mode = ORD_ROTATE;
  u.src = dst;
  u.dst = a;
  return false;
}

// executes the move ord_win / ord_append decided: by the order warp, before it looks at the arrays again
__device__ __forceinline__ void ord_apply(int32_t* ord, int32_t* key, int32_t* pos_of, OrdUndo& u) {
  const int lane = threadIdx.x & 31;
  if (u.pend == 1) {  // those behind move one slot down, the keys of positions [a, dst) stay k
    for (int cb = u.pa + 1; cb <= u.pdst; cb += 32) {
      const int i = cb + lane;
      int v = 0;
      if (i <= u.pdst) v = ord[i];
      __syncwarp();
      if (i <= u.pdst) { ord[i - 1] = v; pos_of[v] = i - 1; }
      __syncwarp();
    }
    if (lane == 0) { ord[u.pdst] = u.pc; key[u.pdst] = u.pk + 1; pos_of[u.pc] = u.pdst; }
    __syncwarp();
  } else if (u.pend == 2) {
    ord_move(ord, key, pos_of, u.pa, u.pdst);
  }
  u.pend = 0;
}

__device__ __forceinline__ bool ord_append(int32_t* ord, int32_t* key, int32_t* pos_of, int32_t* G, int32_t* snap, int n, int c,
                                           OrdUndo& u) {
  const int lane = threadIdx.x & 31;
  const int g1 = G[1];
  __syncwarp();
  if (lane == 0) { ord[n] = c; key[n] = 1; pos_of[c] = n; G[0] = n + 1; }
  __syncwarp();
  u.mode = ORD_NONE;
  if (g1 == 0) return false;  // every claim has one pod: sorted
  const int n1 = n + 1;
  if (n1 < 50) {
    ord_full_sort(ord, key, pos_of, snap, n1);
    u.mode = ORD_GENERIC;
    return true;
  }
  const int dst = n - g1;
  u.pend = 2; u.pa = n; u.pdst = dst;
  u.mode = ORD_ROTATE;
  u.src = dst;
  u.dst = n;
  return false;
}

struct RegParams {
  int32_t R, W, K, NB, T;
  int32_t tb;                 // tie-break bits of a priority; key = priority >> tb
  int32_t wpt;                // feasibility-table words per template
  int32_t cap;                // claims the CTA's threads can hold
  int32_t tab_bytes, win_entries;
  // per resource slot (direct-table slots first, window slots last; slots past the catalog's resources are inert)
  int32_t perm[RT];           // catalog resource of the slot, -1 for padding
  int32_t lim[RT];            // direct slot: the value -> rank table has lim + 1 entries (lim = 2 * maxv + 1)
  int32_t toff[RT];           // direct slot: offset into the byte table; window slot: offset into the window table
  int32_t stride[RT];         // of the slot's rank in the feasibility table
  int32_t row0[RT];           // first bit row "alloc >= dv[0]" of the slot's resource
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
  uint32_t prio;
  int32_t tmpl;
  bool feas;
};

__device__ __forceinline__ void slot_clear(Slot& s) {
#pragma unroll
  for (int r = 0; r < RT; r++) s.x[r] = 0;
  s.wrk = 0; s.nwrk = 0; s.prio = PRIO_INF; s.tmpl = 0; s.feas = false;
}

// NodeClaim.CanAdd for a pod without requirement keys (nodeclaim.go:114-163): whether the filter result of the new
// total is non-empty — the feasibility-table cell of its ranks — and the taints.  Returns the claim's priority when
// the pod fits, PRIO_INF otherwise.  Straight-line code: the common case takes no branch.
template <int NWIN>
__device__ __forceinline__ unsigned slot_eval(Slot& s, const RegParams& prm, const uint8_t* s_tab, const int4* s_win,
                                              const uint32_t* s_bits, const uint64_t* taints, const int32_t (&q)[RT],
                                              uint64_t tol, bool tol_all, bool live) {
  int idx = 0;
  uint64_t nw = 0;
#pragma unroll
  for (int r = 0; r < RT - NWIN; r++) {
    // the table reaches 2 * maxv + 1 >= any claim total (<= maxv) + any pod request (clamped to maxv + 1)
    idx += (int)s_tab[prm.toff[r] + s.x[r] + q[r]] * prm.stride[r];
  }
#pragma unroll
  for (int r = RT - NWIN; r < RT; r++) {
    const int sh = 8 * (r - (RT - NWIN));
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
  bool f = live && ((s_bits[s.tmpl * prm.wpt + (idx >> 5)] >> (idx & 31)) & 1u);
  if (!tol_all) f = f && (taints[s.tmpl] & ~tol) == 0;  // nodeclaim.go:116
  s.feas = f;
  s.nwrk = nw;
  return f ? s.prio : PRIO_INF;
}

// NewNodeClaim + Add (nodeclaim.go:83-109, :168-180): requests = daemon overhead + the pod; back of bucket 1
template <int NWIN>
__device__ __forceinline__ void slot_open(Slot& s, const RegParams& prm, const int4* s_win, const int32_t* daemon,
                                          const int32_t (&q)[RT], int tnew, int TB, int back) {
  uint64_t nw = 0;
#pragma unroll
  for (int r = 0; r < RT; r++) {
    s.x[r] = daemon[r] + q[r];
    if (r >= RT - NWIN) {
      int i = 0;
      while (s_win[prm.toff[r] + i].x < s.x[r]) i++;
      nw |= (uint64_t)i << (8 * (r - (RT - NWIN)));
    }
  }
  s.wrk = nw;
  s.tmpl = tnew;
  s.prio = (1u << TB) | (unsigned)back;
}

// PROF: cycle counters of the warp leaders into PackProblem::prof: [0] step start -> barrier, [1] wait at the
// barrier, [2] barrier -> end of step, [3] steps for warp 0; the same three for the order warp in [4..6] with its
// wait for the winner's id in [7] and the full pdqsorts in [10]; for the queue warp in [8], [9], [11]
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
  int32_t* s_G = reinterpret_cast<int32_t*>(smem + L.gt);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TB = prm.tb;
  const bool is_claim = warp < REG_CW;

  PackProblem& pb = problems[blockIdx.x];
  const int base = pb.pod_begin, n_pods = pb.pod_end - pb.pod_begin;
  if (pb.node_end != pb.node_begin) return;  // existing nodes: pack_kernel's problem (uniform)
  if (pb.status == 1) return;                // finished by the smaller form of this kernel (uniform)
  int32_t* push_list = reinterpret_cast<int32_t*>(A.c_hi + (size_t)base * prm.R);  // scratch: pods pushed back, a ring of n_pods
  int32_t* c_first = A.c_nopt + base;  // scratch: the pod that opened the claim
  int32_t* g_snap = A.ord_snap + base;
  const int32_t* queue = A.queue + base;

  // ---- stage tables ----
  for (int i = tid; i < MAXT; i += REG_THREADS) {
    S.taints[i] = i < tg->T ? tg->taints[i] : 0ull;
    for (int r = 0; r < 8; r++) S.daemon[i][r] = prm.daemon[i * 8 + r];
  }
  for (int i = tid; i < prm.tab_bytes; i += REG_THREADS) s_tab[i] = prm.tab[i];
  for (int i = tid; i < prm.win_entries; i += REG_THREADS) s_win[i] = prm.win[i];
  for (int i = tid; i < prm.T * prm.wpt; i += REG_THREADS) s_bits[i] = prm.bits[i];
  for (int i = tid; i < REG_KCAP; i += REG_THREADS) s_G[i] = 0;
  if (tid == 0) {
    uint64_t at = 0;
    for (int t = 0; t < tg->T; t++) at |= tg->taints[t];
    S.all_taints = at;
    for (int i = 0; i < 3; i++) { S.best[i] = PRIO_INF; S.best2[i] = PRIO_INF; S.win[i] = 0; }
    S.irr[0] = 0; S.irr[1] = 0;
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
  const int OFF = 1 << (TB - 1), TBMAX = (1 << TB) - 2;
  int n = 0, seq = 0, front = OFF - 1, back = OFF;  // uniform across the CTA
  OrdUndo undo{ORD_NONE, 0, 0, 0, 0, 0, 0, 0};       // order warp
  // queue warp: the batch being prepared (one pod per lane), staged over the steps of a batch
  int pf_p = -1, pf_rq = 0, pf_sig = 0, pf_tnew = -1;
  int32_t pf_q[RT];
  uint64_t pf_tol = 0;
#pragma unroll
  for (int r = 0; r < RT; r++) pf_q[r] = 0;
  bool overflow = false;
  unsigned long long prof[5] = {0, 0, 0, 0, 0};
  long long tick = PROF ? clock64() : 0;
#define RPROF(i)                                    \
  do {                                              \
    if (PROF && (tid & 31) == 0) {                  \
      const long long now_ = clock64();             \
      prof[i] += (unsigned long long)(now_ - tick); \
      tick = now_;                                  \
    }                                               \
  } while (0)

  auto pop_slow = [&](int t) {  // second phase (queue warp): Queue.Pop with the lastLen rule over the pushed pods
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

  int s3i = 0;  // t % 3
  for (int t = 0;; t++) {
    if (t == n_pods) {  // the first pass over the queue is over: what is left are the pods that were pushed back
      __syncthreads();
      if (S.n_pushed == 0) break;
      __syncthreads();
      if (warp == REG_QW) {
        if (lane == 0) { S.head = 0; S.len = S.n_pushed; }
        __syncwarp();
        pop_slow(t);
      }
      __syncthreads();
    }
    const bool phase2 = t >= n_pods;
    const RingEnt& e = S.ring[t & (RING - 1)];
    if (phase2 && e.p < 0) break;
    const int s3 = s3i, s3n = s3i == 2 ? 0 : s3i + 1;
    s3i = s3n;

    // ---- before the barrier: each role its own work ----
    int32_t q[RT];
    unsigned cand = PRIO_INF;
    if (is_claim) {
      // addToInflightNode: every claim decides for itself, the least priority among the feasible ones wins
      if (tid == 0) { S.best[s3n] = PRIO_INF; S.best2[s3n] = PRIO_INF; }
      const uint64_t tol = e.tol;
#pragma unroll
      for (int r = 0; r < RT; r++) q[r] = e.q[r];
      const bool tol_all = (S.all_taints & ~tol) == 0;
#define EVAL_SLOT(j, sl)                                                       \
  if ((j) * REG_NCT + (tid & ~31) < n) /* this warp holds a claim in slot j */ \
    cand = min(cand, slot_eval<NWIN>(sl, prm, s_tab, s_win, s_bits, S.taints, q, tol, tol_all, (j) * REG_NCT + tid < n))
      FOR_SLOTS(EVAL_SLOT);
#undef EVAL_SLOT
      const unsigned wmin = __reduce_min_sync(0xffffffffu, cand);
      if (lane == 0 && wmin != PRIO_INF) atomicMin(&S.best[s3], wmin);
    } else if (warp == REG_QW && !phase2) {  // rows of the next batch (the ring's other half), a stage every few steps
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
    }
    RPROF(0);
    __syncthreads();  // B1
    RPROF(1);
    unsigned g;
    const bool renorm = (t > 0 && S.irr[(t - 1) & 1] != 0) || front <= 0 || back >= TBMAX;
    if (renorm) {
      // the previous step ended in a full pdqsort (or the tie-break field ran out): tie-breaks are re-derived from the
      // explicit order, brought up to date first (a move may still be pending) and not touched before the barrier below
      if (warp == REG_ORDW) ord_apply(s_ord, s_key, s_pos, undo);
      __syncthreads();
      if (is_claim) {
        cand = PRIO_INF;
#define RENORM_SLOT(j, sl)                                                            \
  if ((j) * REG_NCT + tid < n) {                                                      \
    sl.prio = ((sl.prio >> TB) << TB) | (unsigned)(OFF + s_pos[(j) * REG_NCT + tid]); \
    if (sl.feas) cand = min(cand, sl.prio);                                           \
  }
        FOR_SLOTS(RENORM_SLOT);
#undef RENORM_SLOT
        const unsigned wmin = __reduce_min_sync(0xffffffffu, cand);
        if (lane == 0 && wmin != PRIO_INF) atomicMin(&S.best2[s3], wmin);
      }
      front = OFF - 1;
      back = OFF + n;
      __syncthreads();  // B2
      g = S.best2[s3];
    } else {
      g = S.best[s3];
    }
    // the move the previous step decided, beside the claim warps' search for this step's winner
    if (warp == REG_ORDW) ord_apply(s_ord, s_key, s_pos, undo);

    if (g != PRIO_INF) {
      // ---- NodeClaim.Add (nodeclaim.go:168-180) by the owner; the claim moves to the front of the next bucket ----
      if (is_claim) {
        // first the hand-off of (claim, pod count) to the order warp, then the owner's own update
        int mine = -1;
#define FIND_SLOT(j, sl) \
  if (sl.feas && sl.prio == g) mine = (j)
        FOR_SLOTS(FIND_SLOT);
#undef FIND_SLOT
        if (mine >= 0) S.win[s3] = (unsigned)(mine * REG_NCT + tid) | ((g >> TB) << 11);
        if (__any_sync(0xffffffffu, mine >= 0)) named_bar_sync(1, 64);
        if (mine >= 0) {
          const int p = e.p;
#define COMMIT_SLOT(j, sl)                                          \
  if (mine == (j)) {                                                \
    _Pragma("unroll") for (int r = 0; r < RT; r++) sl.x[r] += q[r]; \
    sl.wrk = sl.nwrk;                                               \
    sl.prio = (((g >> TB) + 1u) << TB) | (unsigned)front;           \
  }
          FOR_SLOTS(COMMIT_SLOT);
#undef COMMIT_SLOT
          A.pod_kind[p] = B200S_PLACE_CLAIM;
          A.pod_index[p] = mine * REG_NCT + tid;
          A.pod_seq[p] = seq;
        }
      } else if (warp == REG_ORDW) {
        named_bar_sync(1, 64);
        const unsigned wv = S.win[s3];
        RPROF(3);
        const bool gen = ord_win(s_ord, s_key, s_pos, s_G, g_snap, n, (int)(wv & 2047u), (int)(wv >> 11), undo);
        if (lane == 0) S.irr[t & 1] = gen ? 1 : 0;
        if (PROF && gen) prof[4]++;
      }
      seq++;
      front--;
    } else {
      const int tnew = e.tnew, p = e.p;
      if (tnew >= 0) {
        // ---- addToNewNodeClaim: the template walk was done for every pod by stage A (scheduler.go:587-675) ----
        if (n >= prm.cap) { overflow = true; break; }  // uniform: pack_kernel redoes this problem
        const int c = n;
        if (is_claim && tid == c % REG_NCT) {
#define OPEN_SLOT(j, sl) \
  if ((j) == c / REG_NCT) slot_open<NWIN>(sl, prm, s_win, S.daemon[tnew], q, tnew, TB, back)
          FOR_SLOTS(OPEN_SLOT);
#undef OPEN_SLOT
          c_first[c] = p;
          A.pod_kind[p] = B200S_PLACE_CLAIM;
          A.pod_index[p] = c;
          A.pod_seq[p] = seq;
        } else if (warp == REG_ORDW) {
          const bool gen = ord_append(s_ord, s_key, s_pos, s_G, g_snap, n, c, undo);
          if (lane == 0) S.irr[t & 1] = gen ? 1 : 0;
          if (PROF && gen) prof[4]++;
        }
        n++;
        back++;
        seq++;
      } else {
        // ---- Queue.Push (queue.go:63-66): the pod stays in PodErrors ----
        if (warp == REG_QW && lane == 0) {
          if (!phase2) {
            // the reference's queue holds the n_pods - (t + 1) pods not yet popped and the pods pushed so far
            push_list[S.n_pushed] = p;
            A.lastlen[p] = (n_pods - (t + 1)) + S.n_pushed + 1;
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
        } else if (warp == REG_ORDW) {
          undo.mode = ORD_NONE;  // the reference sorted at the start of this add() too: the last normalisation stands
          if (lane == 0) S.irr[t & 1] = 0;
        }
      }
    }
    if (phase2) {
      if (warp == REG_QW) pop_slow(t + 1);
      __syncthreads();
    }
    RPROF(2);
    if (PROF && tid == 0) prof[4]++;
  }

  // ---- epilogue ----
  __syncthreads();
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
    // Results.NewNodeClaims: the order after the last sort the reference ran plus the modification that followed it
    if (warp == REG_ORDW) {
      ord_apply(s_ord, s_key, s_pos, undo);
      if (undo.mode == ORD_ROTATE) ord_move(s_ord, s_key, s_pos, undo.src, undo.dst);
      else if (undo.mode == ORD_GENERIC)
        for (int i = lane; i < n; i += 32) s_ord[i] = g_snap[i];
      __syncwarp();
      for (int i = lane; i < n; i += 32) g_ord[i] = s_ord[i];
    }
    if (is_claim) {
#define STORE_SLOT(j, sl)                                                                    \
  if ((j) * REG_NCT + tid < n) {                                                             \
    const int c = (j) * REG_NCT + tid;                                                       \
    _Pragma("unroll") for (int r = 0; r < RT; r++) if (prm.perm[r] >= 0) {                   \
      int rank;                                                                              \
      if (r < RT - NWIN) rank = s_tab[prm.toff[r] + min(sl.x[r], prm.lim[r])];               \
      else rank = (int)((sl.wrk >> (8 * (r < RT - NWIN ? 0 : r - (RT - NWIN)))) & 0xFFull);  \
      c_req[(size_t)c * R + prm.perm[r]] = (int64_t)sl.x[r] * prm.scale[r];                  \
      c_rk[(size_t)c * R + prm.perm[r]] = (int64_t)(prm.row0[r] + rank);                     \
    }                                                                                        \
    c_tmpl[c] = sl.tmpl;                                                                     \
    c_npods[c] = (int)(sl.prio >> TB);                                                       \
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
    if (PROF) { pb.prof[0] = prof[0]; pb.prof[1] = prof[1]; pb.prof[2] = prof[2]; pb.prof[3] = prof[4]; }
  }
  if (PROF && warp == REG_ORDW && lane == 0) { pb.prof[4] = prof[0]; pb.prof[5] = prof[1]; pb.prof[6] = prof[2]; pb.prof[7] = prof[3]; pb.prof[10] = prof[4]; }
  if (PROF && warp == REG_QW && lane == 0) { pb.prof[8] = prof[0]; pb.prof[9] = prof[1]; pb.prof[11] = prof[2]; }
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
  // slots: direct resources first, then padding, window resources last
  int nwin = 0;
  for (int r = 0; r < R; r++) nwin += direct[r] ? 0 : 1;
  const int form = nwin <= 1 ? 1 : RT;  // compiled forms: one window slot, or every slot a window
  int slot_of[RT];
  {
    int d = 0, w = RT - form;
    for (int r = 0; r < R; r++) {
      if (form == RT) slot_of[r] = w++;
      else if (direct[r]) slot_of[r] = d++;
      else slot_of[r] = RT - 1;
    }
  }
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
    const bool is_window = s >= RT - form;
    const int r = plan->perm[s];
    if (!is_window) {
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
  plan->daemon32.assign((size_t)MAXT * 8, 0);
  for (int t = 0; t < T; t++)
    for (int s = 0; s < RT; s++)
      if (plan->perm[s] >= 0) {
        const int r = plan->perm[s];
        plan->daemon32[(size_t)t * 8 + s] = (int32_t)std::min<int64_t>(daemon[t][r] / g[r], maxv[r] + 1);
      }
  // claims per thread: the largest compiled form shared memory allows
  for (int cpt : {4, 2}) {
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
  int kb = 1;
  while ((1 << kb) <= max_pods) kb++;  // keys up to max_pods
  prm.tb = 31 - kb;
  if (prm.tb < 8) return cudaErrorInvalidValue;
  prm.wpt = plan.wpt;
  prm.tab_bytes = (int32_t)plan.tab.size();
  prm.win_entries = (int32_t)(plan.win.size() / 4);
  for (int s = 0; s < RT; s++) {
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
  // Claims per thread: the two-slot form first (less code on the step's critical path); a problem that needs more
  // claims than it holds leaves status 0 and the four-slot form — then pack_kernel — redoes it, each skipping what an
  // earlier one finished.  A form is launched only if some problem could still need it.
  const char* force = getenv("B200S_REG_CPT");  // experiments: "4" skips the two-slot form
  int launched = 0;
  for (int cpt : {2, 4}) {
    if (cpt > plan.cpt) break;
    if (cpt == 2 && force && atoi(force) == 4 && plan.cpt >= 4) continue;
    if (launched && REG_NCT * (cpt / 2) >= max_pods) break;  // the smaller form cannot have overflowed
    prm.cap = REG_NCT * cpt;
    const size_t bytes = reg_layout(prm.tab_bytes, prm.win_entries, T, plan.wpt, prm.cap).total;
    if (plan.nwin_form == 1) {
      if (cpt == 2) REG_CALL(1, 2);
      else REG_CALL(1, 4);
    } else {
      if (cpt == 2) REG_CALL(RT, 2);
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

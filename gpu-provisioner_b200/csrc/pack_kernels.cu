// pack_kernels.cu — the first-fit-decreasing bin-pack of Scheduler.Solve as a persistent kernel:
// one CTA per Solve (independent simulations of a batch run on different SMs), pods consumed one at a
// time in queue order exactly as the reference does (K/controllers/provisioning/scheduling/
// scheduler.go:377-432, queue.go:46-66), parallelism inside a step across existing nodes, across
// in-flight NodeClaims and across SKU words.
//
//   step(pod):  addToExistingNode  (scheduler.go:515-550, existingnode.go:70-119)   thread per node
//               sort.Slice by len(Pods) (scheduler.go:499)                          exact Go pdqsort order
//               addToInflightNode  (scheduler.go:552-584, nodeclaim.go:114-180)     lane group per claim
//               addToNewNodeClaim  (scheduler.go:587-675)                           stage A row of the pod
//
// "First success by lowest index" (scheduler.go:533-540, :566-575) is an atomicMin over the indices
// evaluated in parallel; candidates are visited in ascending chunks so the minimum is global.
#include <climits>
#include <cstdlib>

#include "device_types.cuh"
#include "kernels.h"

namespace b200 {

constexpr int PACK_THREADS = 512;
constexpr int ORD_SMEM_CAP = 8192;    // NodeClaims whose order / count-by-position arrays stay in shared memory
constexpr int ORD_SMALL_CAP = 12288;  // the same for the SMALL form: no problem has more pods than this, so the order
                                      // never leaves shared memory and the spill code is compiled out
constexpr int MODE_NONE = 0;

struct PackShared {
  TmplDev tm;
  // pod rows, double buffered: the next pod's row is fetched while the current one commits
  int64_t p_req[2][MAXR];
  uint64_t p_mask[2][MAXK];
  int32_t p_gt[2][MAXB], p_lt[2][MAXB];
  uint64_t p_flags[2], p_tol[2];
  uint32_t p_keys[2];  // bit k: the pod's requirements define key k
  uint32_t p_ract[2];  // bit r: the pod requests resource r (non-zero)
  int32_t p_id[2];     // pod index, -1 when the queue is done
  int64_t zero_b[MAXB];
  int32_t best[2];
  int32_t hit_c[PACK_THREADS / 32];  // claim id behind each warp's first feasible index of the current scan
  int32_t n_claims, seq, ord_global, _pad0;
  int32_t head, len, n_pushed, _pad1;  // the queue warp's own words, kept apart from what the other warps read
  int32_t mode, rot_src, rot_dst, rot_orig;   // order normalisation left for the whole CTA after this step's commit
  int32_t undo_mode, undo_src, undo_dst;      // how to take the last normalisation back (see the end of the kernel)
  unsigned long long evals;
  uint64_t all_taints;  // union of the templates' taints
  int64_t remaining[MAXT][MAXR];  // NodePool limits: s.remainingResources per template (scheduler.go:170-172)
  int32_t t_new, _pad2;           // the template a limit-aware walk chose for the current pod, -1 none
  unsigned long long wprof[4];  // diagnostic: cycles of the commit phase by role (next pod's row, claim order, claim
                                // state) and, in [3], full pdqsort runs (count << 40 | cycles)
};

constexpr int MODE_ROTATE = 1, MODE_GENERIC = 2;

__host__ __device__ inline size_t pack_align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t pack_smem_bytes(int G, int W, int n_rows, int n_dv, int tables_in_smem, size_t* scratch_off,
                                                  size_t* tables_off, size_t* ord_off = nullptr, int ord_cap = ORD_SMEM_CAP) {
  (void)G;
  size_t o = pack_align16(sizeof(PackShared));
  if (scratch_off) *scratch_off = o;
  if (ord_off) *ord_off = o;
  o += pack_align16((size_t)ord_cap * 8);  // ord[] + key[]
  if (tables_off) *tables_off = o;
  if (tables_in_smem)  // bit rows, distinct allocatable values, rank buckets
    o += pack_align16((size_t)n_rows * W * 8) + pack_align16((size_t)n_dv * 8) + pack_align16((size_t)MAXR * (RANK_BKTS + 1) * 2);
  return o;
}

// ExistingNode.CanAdd (existingnode.go:70-107) for one node by one thread: taints, Fits(pod, remaining),
// Compatible(node labels, pod requirements) with no allow-set.
template <bool KEYS>
__device__ __forceinline__ bool node_can_add(const KeyTab& kt, const PackArrays& A, int row, const PodRef& pd,
                                             const int64_t* zero_b) {
  if (A.n_taints[row] & ~pd.tol) return false;
  for (int r = 0; r < kt.n_res; r++) {
    int64_t avail = A.n_avail[(size_t)r * A.Etot + row];
    if (avail < 0 || pd.req[r] > avail) return false;  // resources.go:234-243
  }
  if constexpr (KEYS) {  // a pod without requirements is compatible with any label set
    StateRef st{A.n_rmask + row, (size_t)A.Etot, A.n_rflags[row], zero_b, zero_b, 1, nullptr, 0, 0};
    for (int k = 0; k < kt.n_keys; k++)
      if (!compatible_key(kt, k, state_req(kt, st, k), pod_req(kt, pd, k), false)) return false;
  }
  return true;
}

// Moves the element at position src of (ord, key) to position dst, shifting what lies between by one slot;
// all threads of the CTA take part (uniform arguments).
__device__ __forceinline__ void cta_rotate(int32_t* ord, int32_t* key, int src, int dst) {
  const int tid = threadIdx.x;
  const int x = ord[src], kx = key[src];
  __syncthreads();
  if (src < dst) {
    for (int cb = src + 1; cb <= dst; cb += PACK_THREADS) {
      int i = cb + tid, v = 0, kv = 0;
      if (i <= dst) { v = ord[i]; kv = key[i]; }
      __syncthreads();
      if (i <= dst) { ord[i - 1] = v; key[i - 1] = kv; }
      __syncthreads();
    }
  } else {
    for (int ce = src - 1; ce >= dst; ce -= PACK_THREADS) {
      int i = ce - tid, v = 0, kv = 0;
      if (i >= dst) { v = ord[i]; kv = key[i]; }
      __syncthreads();
      if (i >= dst) { ord[i + 1] = v; key[i + 1] = kv; }
      __syncthreads();
    }
  }
  if (tid == 0) { ord[dst] = x; key[dst] = kx; }
  __syncthreads();
}

// The same move by one warp (no CTA barrier): used inside the commit phase while the other warps do their part.
__device__ __forceinline__ void warp_rotate(int32_t* ord, int32_t* key, int src, int dst) {
  const int lane = threadIdx.x & 31;
  const int x = ord[src], kx = key[src];
  __syncwarp();
  if (src < dst) {
    for (int cb = src + 1; cb <= dst; cb += 32) {
      int i = cb + lane, v = 0, kv = 0;
      if (i <= dst) { v = ord[i]; kv = key[i]; }
      __syncwarp();
      if (i <= dst) { ord[i - 1] = v; key[i - 1] = kv; }
      __syncwarp();
    }
  } else {
    for (int ce = src - 1; ce >= dst; ce -= 32) {
      int i = ce - lane, v = 0, kv = 0;
      if (i >= dst) { v = ord[i]; kv = key[i]; }
      __syncwarp();
      if (i >= dst) { ord[i + 1] = v; key[i + 1] = kv; }
      __syncwarp();
    }
  }
  if (lane == 0) { ord[dst] = x; key[dst] = kx; }
  __syncwarp();
}

// What Go's sort.Slice will do to an order that was sorted before one element changed (position pos got one more
// pod, or a new claim with one pod was appended at pos = n-1).  Run by ONE WARP right after the change:
//   * still sorted                                   -> identity
//   * n >= 50 and choosePivot sees no inversion      -> partialInsertionSort re-inserts the element next to its
//                                                       equals (zsortfunc.go partialInsertionSort_func): the warp
//                                                       shifts the run it passes over, 32 slots per step
//   * otherwise                                      -> the full pdqsort (lane 0), order saved first for the undo
// choosePivot_func with n >= 50 reports "increasing" iff none of its order2 calls swaps, i.e. iff the three
// adjacent triples around n/4, n/2, 3n/4 are non-decreasing and so are their middles: 9 keys, read by 9 lanes.
constexpr int WARP_SHIFT_LIMIT = 1024;  // slots one warp shifts before handing the rest to the whole CTA

__device__ __forceinline__ void warp_normalise(PackShared& S, int32_t* ord, int32_t* key, int32_t* snap, int n, int pos,
                                               bool appended) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) S.mode = MODE_NONE;
  if (n <= 1) return;
  const bool sorted = appended ? !(key[n - 1] < key[n - 2]) : (pos == n - 1 || !(key[pos + 1] < key[pos]));
  if (sorted) return;
  bool increasing = false;
  if (n >= 50) {
    const int q = n / 4;
    int v = 0;
    if (lane < 9) v = key[q * (lane / 3 + 1) + (lane % 3) - 1];
    const int up = __shfl_down_sync(0xffffffffu, v, 1);
    const int mid_next = __shfl_down_sync(0xffffffffu, v, 3);
    bool bad = false;
    if (lane < 9 && (lane % 3) != 2) bad = up < v;                  // inside a triple
    if (lane == 1 || lane == 4) bad = bad || mid_next < v;          // middle of triple t vs middle of triple t+1
    increasing = __ballot_sync(0xffffffffu, bad) == 0;
  }
  if (!increasing) {
    // rare path (measured: one packed word per element instead of two arrays made this sort slower, not faster)
    for (int i = lane; i < n; i += 32) snap[i] = ord[i];  // lets the last normalisation be taken back
    __syncwarp();
    {
      const long long t_g0 = clock64();
      GoSort gs{ord, key, true};  // the whole warp: lane 0 writes, the linear scans go 32 elements per step
      gs.sort(n);
      __syncwarp();
      if (lane == 0) {
        S.mode = MODE_GENERIC;
        S.wprof[3] += (1ull << 40) + (unsigned long long)(clock64() - t_g0);  // count in the high bits, cycles below
      }
    }
    __syncwarp();
    return;
  }
  // re-insertion: the displaced element X walks over the run of smaller (when it grew) or larger (when it was
  // appended) keys; the run moves one slot the other way
  const int x = ord[pos], kx = key[pos];
  __syncwarp();
  int dst = -1, hole = pos, moved = 0;
  if (!appended) {
    for (int b0 = pos + 1; dst < 0 && moved < WARP_SHIFT_LIMIT; b0 += 32) {
      const int i = b0 + lane;
      int v = 0, kv = 0;
      bool stop = true;
      if (i < n) { v = ord[i]; kv = key[i]; stop = !(kv < kx); }
      const unsigned sb = __ballot_sync(0xffffffffu, stop);
      const int cnt = sb ? __ffs(sb) - 1 : 32;
      __syncwarp();
      if (lane < cnt) { ord[i - 1] = v; key[i - 1] = kv; }
      __syncwarp();
      hole = b0 + cnt - 1;
      moved += cnt;
      if (sb) dst = hole;
    }
  } else {
    for (int t0 = pos - 1; dst < 0 && moved < WARP_SHIFT_LIMIT; t0 -= 32) {
      const int i = t0 - lane;
      int v = 0, kv = 0;
      bool stop = true;
      if (i >= 0) { v = ord[i]; kv = key[i]; stop = !(kx < kv); }
      const unsigned sb = __ballot_sync(0xffffffffu, stop);
      const int cnt = sb ? __ffs(sb) - 1 : 32;
      __syncwarp();
      if (lane < cnt) { ord[i + 1] = v; key[i + 1] = kv; }
      __syncwarp();
      hole = t0 - cnt + 1;
      moved += cnt;
      if (sb) dst = hole;
    }
  }
  if (lane == 0) {
    ord[hole] = x;
    key[hole] = kx;
    if (dst >= 0) {  // done: remember how to take it back
      S.undo_mode = hole != pos ? MODE_ROTATE : MODE_NONE;
      S.undo_src = hole;
      S.undo_dst = pos;
    } else {  // long run: X sits at `hole`, the whole CTA finishes the move after the barrier
      int lo, hi;
      if (!appended) {
        lo = hole + 1; hi = n;  // first index > hole whose count >= kx
        while (lo < hi) { int mid = (lo + hi) >> 1; if (key[mid] < kx) lo = mid + 1; else hi = mid; }
        S.rot_dst = lo - 1;
      } else {
        lo = 0; hi = hole;      // first index whose count > kx
        while (lo < hi) { int mid = (lo + hi) >> 1; if (key[mid] <= kx) lo = mid + 1; else hi = mid; }
        S.rot_dst = lo;
      }
      S.rot_src = hole;
      S.rot_orig = pos;
      S.mode = MODE_ROTATE;
    }
  }
  __syncwarp();
}

// ---- NodePool limits (scheduler.go:587-675, :831-867), all by one warp ----
// filterByRemainingResources as a bit row: the instance types whose Capacity stays within the template's remaining
// resources on every limited resource = AND over those resources of the row "capacity[r] <= largest value <= remaining"
__device__ __forceinline__ void warp_limit_row(const KeyTab& kt, const PackShared& S, int t, int W, uint64_t* L) {
  const int lane = threadIdx.x & 31;
  for (int w = lane; w < W; w += 32) {
    uint64_t m = ~0ull;
    for (int r = 0; r < kt.n_res; r++) {
      if (!((S.tm.limit_mask[t] >> r) & 1u)) continue;
      const int64_t* cv = kt.cap_dv + kt.cap_dv0[r];
      int lo = 0, hi = kt.cap_nd[r];  // number of values <= remaining
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ldg(cv + mid) <= S.remaining[t][r]) lo = mid + 1; else hi = mid; }
      m &= lo == 0 ? 0ull : ldg(kt.rows + (size_t)(kt.cap_row0[r] + lo - 1) * W + w);
    }
    L[(size_t)t * W + w] = m;
  }
  __syncwarp();
}

// addToNewNodeClaim's template walk when NodePools have limits: the candidate set of template t is its pre-filtered
// instance types inside the limit row, an empty one is the "exceed limits" code, and the six diagnostic flags are taken
// over the limited set.  Writes the pod's attempt codes; returns the first template that works, or -1.
__device__ __forceinline__ int warp_limited_walk(const KeyTab& kt, const PackShared& S, const PackArrays& A, const PodsDev& pods,
                                                 int p, int W, const uint64_t* L) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, T = S.tm.T;
  const size_t sig = (size_t)pods.sig_id[p], rq = (size_t)pods.req_id[p];
  int chosen = -1;
  for (int t = 0; t < T; t++) {
    const uint8_t status = A.sg_status[sig * T + t];
    uint8_t code = status == B200S_FIT_DROPPED ? (uint8_t)B200S_FIT_DROPPED : (uint8_t)B200S_FIT_NONE;
    if (status != B200S_FIT_DROPPED && chosen < 0) {
      const bool lim = S.tm.limit_mask[t] != 0;
      bool any = true;
      if (lim) {
        bool a = false;
        for (int w = lane; w < W; w += 32) a = a || (A.tmask[(size_t)t * W + w] & L[(size_t)t * W + w]) != 0;
        any = __any_sync(FULL, a);
      }
      if (!any) {
        code = B200S_FIT_LIMITS;  // scheduler.go:603
      } else if (status != B200S_FIT_NONE) {
        code = status;  // taints / incompatible requirements (nodeclaim.go:116,128)
      } else {
        bool nz = false, bC = false, bF = false, bO = false, bCF = false, bCO = false, bFO = false;
        for (int w = lane; w < W; w += 32) {
          const uint64_t l = lim ? L[(size_t)t * W + w] : ~0ull;
          const uint64_t c = A.sgC[(sig * T + t) * W + w] & l, o = A.sgO[(sig * T + t) * W + w] & l,
                         f = A.rqF[(rq * T + t) * W + w] & l;
          nz = nz || (c & o & f) != 0;
          bC = bC || c; bF = bF || f; bO = bO || o;
          bCF = bCF || (c & f & ~o); bCO = bCO || (c & o & ~f); bFO = bFO || (f & o & ~c);
        }
        if (__any_sync(FULL, nz)) {
          chosen = t;
        } else {
          code = filter_flags(__any_sync(FULL, bC), __any_sync(FULL, bF), __any_sync(FULL, bO), __any_sync(FULL, bCF),
                              __any_sync(FULL, bCO), __any_sync(FULL, bFO));
        }
      }
    }
    if (lane == 0) A.a_code[(size_t)p * T + t] = code;
  }
  return chosen;
}

// word w of the instance types a NEW NodeClaim of template t keeps for pod p: its stage-A row, inside the template's
// limit row when the NodePool has limits
__device__ __forceinline__ uint64_t new_claim_word(const PackArrays& A, const PodsDev& pods, const PackShared& S, const uint64_t* L,
                                                   int p, int t, int W, int w) {
  uint64_t m = stage_a_word(A, pods, p, t, W, w);
  if (S.tm.limit_mask[t]) m &= L[(size_t)t * W + w];
  return m;
}

// subtractMax (scheduler.go:831-848) after a NodeClaim of a limited template opened: remaining -= the largest Capacity
// among the claim's instance types, per limited resource — the smallest row "capacity <= cv[i]" that holds them all
__device__ __forceinline__ void warp_subtract_max(const KeyTab& kt, PackShared& S, const PackArrays& A, const PodsDev& pods,
                                                  const uint64_t* L, int p, int t, int W) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (int r = 0; r < kt.n_res; r++) {
    if (!((S.tm.limit_mask[t] >> r) & 1u)) continue;
    const int64_t* cv = kt.cap_dv + kt.cap_dv0[r];
    int64_t mx = 0;
    for (int i = kt.cap_nd[r] - 1; i >= 0; i--) {
      bool out = false;
      for (int w = lane; w < W; w += 32)
        out = out || (new_claim_word(A, pods, S, L, p, t, W, w) & ~ldg(kt.rows + (size_t)(kt.cap_row0[r] + i) * W + w)) != 0;
      if (__any_sync(FULL, out)) break;
      mx = ldg(cv + i);
    }
    __syncwarp();
    if (lane == 0) S.remaining[t][r] -= mx;
    __syncwarp();
  }
}

// KEYS: some pod of the batch defines a requirement key; NODES: there are existing nodes.  Without them the compiler
// drops the requirement algebra / the node paths from the step loop (16.3 k -> 8.2 k SASS instructions without both),
// and this kernel's speed follows its size: one CTA, every warp in a different region of ~250 KB of code
// (config[1]: 30.4 -> 25.8 ms).
// PROFILE: the per-phase / per-role cycle counters of b200s_resident_pack_profile (clock reads on the step's critical
// path, 3-4 % of it); every shared-memory-table form is also built without them and runs that way unless profiling
// was asked for (b200s_resident_set_pack_profile).
template <int G, int WPL, bool TS, bool KEYS, bool NODES, bool SMALL, bool PROFILE>
__global__ void __launch_bounds__(PACK_THREADS, 1) pack_kernel(const __grid_constant__ KeyTab kt,
                                                               const TmplDev* __restrict__ tg, PodsDev pods, PackArrays A,
                                                               PackProblem* __restrict__ problems, int tables_in_smem) {
  // kt stays in the constant bank (never modified): the bit-row tables are read through L1 from global memory,
  // they are read-only and a few tens of KB
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NW = PACK_THREADS / 32;
  PackShared& S = *reinterpret_cast<PackShared*>(smem);
  size_t scratch_off, tables_off, ord_off;
  constexpr int ORD_CAP = SMALL ? ORD_SMALL_CAP : ORD_SMEM_CAP;
  pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, tables_in_smem, &scratch_off, &tables_off, &ord_off, ORD_CAP);
  int32_t* s_ord = reinterpret_cast<int32_t*>(smem + ord_off);
  int32_t* s_key = s_ord + ORD_CAP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gl = tid & (G - 1);
  const int R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded, W = kt.n_words;
  // TS: the resource tables of the in-flight check (bit rows, distinct values, rank buckets) are staged into shared
  // memory -- the check is a chain of four dependent table reads per claim and the L1 does not keep them
  LaneRes lres = lane_res(kt, gl);
  const uint64_t* q_rows = kt.rows;
  if (TS) {
    uint64_t* s_rows = reinterpret_cast<uint64_t*>(smem + tables_off);
    int64_t* s_dv = reinterpret_cast<int64_t*>(smem + tables_off + pack_align16((size_t)kt.n_rows * kt.n_words * 8));
    uint16_t* s_bkt = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(s_dv) + pack_align16((size_t)kt.n_dv * 8));
    for (int i = tid; i < kt.n_rows * kt.n_words; i += PACK_THREADS) s_rows[i] = kt.rows[i];
    for (int i = tid; i < kt.n_dv; i += PACK_THREADS) s_dv[i] = kt.res_dv[i];
    for (int i = tid; i < MAXR * (RANK_BKTS + 1); i += PACK_THREADS) s_bkt[i] = kt.res_bkt[i];
    q_rows = s_rows;
    if (gl < R) {
      lres.dv = s_dv + kt.res_dv0[gl];
      lres.bkt = s_bkt + (size_t)gl * (RANK_BKTS + 1);
    }
  }

  PackProblem& pb = problems[blockIdx.x];
  if (pb.status == 1) return;  // finished by the register-resident form (pack_reg_kernels.cu); uniform
  const int base = pb.pod_begin, n_pods = pb.pod_end - pb.pod_begin;
  const int E0 = pb.node_begin, E = NODES ? pb.node_end - pb.node_begin : 0;
  // this problem's slices
  uint64_t* c_mask = A.c_mask + (size_t)base * W;
  int64_t* c_req = A.c_req + (size_t)base * R;
  int64_t* c_hi = A.c_hi + (size_t)base * R;
  int64_t* c_thr = A.c_thr + (size_t)base * R;
  uint64_t* c_rmask = A.c_rmask + (size_t)base * K;
  uint64_t* c_rflags = A.c_rflags + base;
  int64_t* c_gt = A.c_gt + (size_t)base * NB;
  int64_t* c_lt = A.c_lt + (size_t)base * NB;
  int32_t* c_tmpl = A.c_tmpl + base;
  int32_t* c_npods = A.c_npods + base;
  int32_t* c_nopt = A.c_nopt + base;
  int32_t* g_ord = A.ord + base;      // the order array's home (and its spill area past ORD_SMEM_CAP claims)
  int32_t* g_key = A.ord_key + base;
  int32_t* g_snap = A.ord_snap + base;
  int32_t* queue = A.queue + base;

  {
    uint64_t* d = reinterpret_cast<uint64_t*>(&S.tm);
    const uint64_t* s = reinterpret_cast<const uint64_t*>(tg);
    for (int i = tid; i < (int)(sizeof(TmplDev) / 8); i += PACK_THREADS) d[i] = s[i];
  }
  if (tid < MAXB) S.zero_b[tid] = 0;
  if (tid == 0) {
    S.n_claims = 0; S.head = 0; S.len = n_pods; S.seq = 0; S.evals = 0;
    S.ord_global = 0; S.n_pushed = 0; S.mode = MODE_NONE; S.undo_mode = MODE_NONE;
    S.best[0] = INT_MAX; S.best[1] = INT_MAX;
    for (int i = 0; i < 4; i++) S.wprof[i] = 0;
    uint64_t at = 0;
    for (int t = 0; t < tg->T; t++) at |= tg->taints[t];
    S.all_taints = at;
  }
  uint64_t* lim_rows = A.limit_rows ? A.limit_rows + (size_t)blockIdx.x * tg->T * W : nullptr;
  for (int i = tid; i < MAXT * MAXR; i += PACK_THREADS) S.remaining[i / MAXR][i % MAXR] = tg->limit[i / MAXR][i % MAXR];
  unsigned long long my_evals = 0;
  unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tick = clock64();
#define PROF(i)                                       \
  do {                                                \
    if (PROFILE && tid == 0) {                        \
      long long now_ = clock64();                     \
      prof[i] += (unsigned long long)(now_ - tick);   \
      tick = now_;                                    \
    }                                                 \
  } while (0)
  __syncthreads();

  // Queue.Pop (queue.go:46-60) + the pod's row (updateCachedPodData's products, scheduler.go:467-486) through
  // its signature / request-class ids, by the queue warp into buffer `buf`.
  // The row is three dependent reads away from the queue (pod id -> class ids -> class rows), so the queue warp runs
  // them as a three-stage pipeline across steps: this step's row was requested one step ago and sits in registers,
  // the class ids of the pod after it were requested one step ago as well, and so was the pod id after that.  Every
  // stage carries the pod id it belongs to; a popped pod that does not match (queue shorter than the look-ahead, or
  // a pod pushed back in between) is fetched on the spot.  Entries ahead of the queue's tail never change.
  // Per-lane roles for the row: lane r < 8 the request of resource r, lanes 8..23 the value mask of key lane-8,
  // lanes 24..27 the Gt/Lt bounds of bounded key lane-24, lane 28 the flags word and the tolerated-taint set.
  constexpr int QW = NW - 1;  // the queue warp: the last one; it takes no part in the claim scan
  int pf_pod1 = -1, pf_pod2 = -1, pf_sig2 = 0, pf_rq2 = 0, pf_pod3 = -1;
  uint64_t pf_v0 = 0, pf_v1 = 0;
  // what this lane fetches, fixed for the kernel: a 64-bit word indexed by the request class or the signature, and
  // for the lanes with a second value a 64-bit or two 32-bit words
  const uint64_t* rf_a = nullptr;  // 64-bit source of v0
  const uint64_t* rf_b = nullptr;  // 64-bit source of v1
  const int32_t* rf_g = nullptr;   // 32-bit sources of (v0, v1): the Gt / Lt bounds
  const int32_t* rf_l = nullptr;
  bool rf_by_rq = false;
  if (lane < 8) {
    if (lane < R) { rf_a = reinterpret_cast<const uint64_t*>(pods.rq_req + (size_t)lane * pods.NQ); rf_by_rq = true; }
  } else if (lane < 24) {
    if (lane - 8 < K) rf_a = pods.sig_mask + (size_t)(lane - 8) * pods.NS;
  } else if (lane < 28) {
    if (lane - 24 < NB) { rf_g = pods.sig_gt + (size_t)(lane - 24) * pods.NS; rf_l = pods.sig_lt + (size_t)(lane - 24) * pods.NS; }
  } else if (lane == 28) {
    rf_a = pods.sig_flags;
    rf_b = pods.sig_tol;
  }
  uint64_t* st_a = nullptr;  // where this lane's v0 goes in the row buffer 0, and the distance to buffer 1
  int st_a_step = 0;
  if (lane < 8) {
    if (lane < R) { st_a = reinterpret_cast<uint64_t*>(&S.p_req[0][lane]); st_a_step = MAXR; }
  } else if (lane < 24) {
    if (lane - 8 < K) { st_a = &S.p_mask[0][lane - 8]; st_a_step = MAXK; }
  }
  auto row_fetch = [&](int sig, int rq, uint64_t& v0, uint64_t& v1) {
    const int idx = rf_by_rq ? rq : sig;
    v0 = 0;
    v1 = 0;
    if (rf_a) v0 = rf_a[idx];
    if (rf_b) v1 = rf_b[idx];
    if (rf_g) {
      v0 = (uint64_t)(uint32_t)rf_g[idx];
      v1 = (uint64_t)(uint32_t)rf_l[idx];
    }
  };
  auto pop_and_load = [&](int buf) -> int {
    int p = -1;
    if (lane == 0) {
      if (S.len > 0) {
        int cand = queue[S.head];
        if (S.n_pushed == 0 || A.lastlen[cand] != S.len) {  // lastLen is zero until a pod has been pushed back
          p = cand;
          S.head = S.head + 1 == n_pods ? 0 : S.head + 1;
          S.len--;
        }
      }
      S.p_id[buf] = p;
    }
    p = __shfl_sync(0xffffffffu, p, 0);
    if (p >= 0) {
      uint64_t v0 = pf_v0, v1 = pf_v1;
      if (p != pf_pod1) {  // not the row that was requested a step ago
        const int sig = p == pf_pod2 ? pf_sig2 : pods.sig_id[p], rq = p == pf_pod2 ? pf_rq2 : pods.req_id[p];
        row_fetch(sig, rq, v0, v1);
      }
      const unsigned ract = __ballot_sync(0xffffffffu, lane < R && v0 != 0);
      if (st_a) st_a[buf ? st_a_step : 0] = v0;  // the request / the key's value mask
      if (lane == 0) S.p_ract[buf] = ract;
      if (rf_g) {
        S.p_gt[buf][lane - 24] = (int32_t)(uint32_t)v0;
        S.p_lt[buf][lane - 24] = (int32_t)(uint32_t)v1;
      } else if (lane == 28) {
        // bit k of keys = the DEFINED bit (bit 0) of nibble k of the flags word: gather every fourth bit
        uint64_t x = v0 & 0x1111111111111111ull;
        x = (x | (x >> 3)) & 0x0303030303030303ull;
        x = (x | (x >> 6)) & 0x000F000F000F000Full;
        x = (x | (x >> 12)) & 0x000000FF000000FFull;
        x = (x | (x >> 24)) & 0xFFFFull;
        const uint32_t keys = (uint32_t)x & (K >= 32 ? 0xffffffffu : ((1u << K) - 1u));
        S.p_flags[buf] = v0;
        S.p_keys[buf] = keys;
        S.p_tol[buf] = v1;
      }
    }
    // advance the pipeline; nothing below is waited for in this step
    __syncwarp();
    const int len = S.len, head = S.head;  // after the pop: queue[head] is the pod of the step after next
    pf_pod1 = len > 0 ? pf_pod2 : -1;
    if (pf_pod1 >= 0) row_fetch(pf_sig2, pf_rq2, pf_v0, pf_v1);
    pf_pod2 = len > 1 ? pf_pod3 : -1;
    if (pf_pod2 >= 0) {
      pf_sig2 = pods.sig_id[pf_pod2];
      pf_rq2 = pods.req_id[pf_pod2];
    }
    pf_pod3 = -1;
    if (len > 2) {
      int at = head + 2;
      if (at >= n_pods) at -= n_pods;
      pf_pod3 = queue[at];
    }
    return p;
  };
  // prime the three stages for the first pods (dependent reads, once)
  if (warp == QW && n_pods > 0) {
    pf_pod1 = queue[0];
    row_fetch(pods.sig_id[pf_pod1], pods.req_id[pf_pod1], pf_v0, pf_v1);
    if (n_pods > 1) {
      pf_pod2 = queue[1];
      pf_sig2 = pods.sig_id[pf_pod2];
      pf_rq2 = pods.req_id[pf_pod2];
    }
    if (n_pods > 2) pf_pod3 = queue[2];
  }

  if (warp == QW) pop_and_load(0);
  __syncthreads();
  const bool any_limit = S.tm.any_limit != 0 && lim_rows != nullptr;  // uniform
  if (any_limit) {
    for (int t = warp; t < S.tm.T; t += NW) warp_limit_row(kt, S, t, W, lim_rows);
    __syncthreads();
  }
  int cur = 0;

  for (;;) {
    const int p = S.p_id[cur];
    if (p < 0) break;
    const PodRef pd{S.p_mask[cur], 1, S.p_flags[cur], S.p_gt[cur], S.p_lt[cur], 1, S.p_req[cur], 1, S.p_tol[cur]};
    const uint32_t podkeys = KEYS ? S.p_keys[cur] : 0u, ract = S.p_ract[cur];
    const bool tol_all = (S.all_taints & ~pd.tol) == 0;
    // While no pod has ever been pushed back, Queue.Pop does not depend on this pod's outcome (the lastLen rule only
    // looks at pods that were pushed, queue.go:46-60): the queue warp pops the next pod and stores its row NOW, beside
    // the scan it takes no part in, instead of in the commit phase where it was the longest role.
    const bool fast = S.n_pushed == 0;
    // read before the queue warp goes off on its own: a step without a barrier in its scan (no claims yet, or the pod
    // lands on an existing node) lets the other warps reach the commit, where warp 2 may add a claim
    const int n = S.n_claims;
    bool spec_popped = false;
    if (warp == QW && fast) {
      const long long t_p0 = PROFILE ? clock64() : 0;
      spec_popped = pop_and_load(cur ^ 1) >= 0;
      __syncwarp();
      if (PROFILE && lane == 0) S.wprof[0] += (unsigned long long)(clock64() - t_p0);
    }
    prof[7] += tid == 0;
    PROF(0);

    // ---- addToExistingNode ----
    // Every thread walks ITS nodes (tid, tid + 512, ...) in ascending order and stops at its first hit: the minimum over
    // the threads' first hits is the lowest feasible index (scheduler.go:533-540) — one barrier pair per pod instead of
    // one per 512 nodes.  A thread also stops once its next node lies behind a hit somebody has already published.
    int found = INT_MAX;
    if (E > 0) {
      for (int e = tid; e < E; e += PACK_THREADS) {
        if (e >= PACK_THREADS && e > atomicMin(&S.best[cur], INT_MAX)) break;  // an atomic read: no hazard with the atomicMin below
        if (node_can_add<KEYS>(kt, A, E0 + e, pd, S.zero_b)) {
          atomicMin(&S.best[cur], e);
          break;
        }
      }
      __syncthreads();
      found = S.best[cur];
      __syncthreads();
    }
    const bool on_node = found != INT_MAX;
    PROF(1);

    const bool og = SMALL ? false : (S.ord_global != 0);
    int32_t* ord = og ? g_ord : s_ord;
    int32_t* key = og ? g_key : s_key;
    uint64_t out0[WPL];
    int my_idx0 = -1, my_c0 = -1;
    bool my_changed0 = false;  // the winning group's claim moves to new table rows (its instance types shrink)
    int64_t my_thr0 = 0;       // lane r < R of that group: resource r's new threshold when it moved, else 0
    if (!on_node) {
      // ---- addToInflightNode ----
      // One thread per in-flight claim, s.newNodeClaims order: warp w takes claims [32(w + it*NW), +32) in
      // pass `it` and stops as soon as the CTA-wide best index lies before its next block, so every index below
      // the final minimum has been evaluated and rejected (scheduler.go:566-575).
      if (tid == 0) S.undo_mode = MODE_NONE;  // the reference sorts here too: the last normalisation stands
      // Round 0 looks at the first 64 claims with two warps (most pods land there); only if none takes the pod
      // do all warps walk the rest, 512 claims per pass.
      auto scan_block = [&](int b0) -> bool {
        const int idx = b0 + lane;
        int code = 0, c = -1;
        if (idx < n) {
          c = ord[idx];
          StateRef st{c_rmask + (size_t)c * K, 1, podkeys ? c_rflags[c] : 0ull, c_gt + (size_t)c * NB, c_lt + (size_t)c * NB, 1,
                      c_req + (size_t)c * R, 1, S.tm.taints[c_tmpl[c]]};
          code = thread_inflight(kt, st, c_hi + (size_t)c * R, c_thr + (size_t)c * R, c_mask + (size_t)c * W, pd, podkeys, ract);
          if (code) my_evals += (unsigned long long)c_nopt[c];
        }
        const unsigned hit = __ballot_sync(0xffffffffu, code == 2);
        if (!hit) return false;
        const int first = __ffs(hit) - 1;
        const int hc = __shfl_sync(0xffffffffu, c, first);
        if (lane == 0) {
          S.hit_c[warp] = hc;
          atomicMin(&S.best[cur], b0 + first);
        }
        return true;
      };
      // Lane-group rounds: the first GROUP_ROUNDS * NW * (32/G) claims of the order, one lane group each.  The order is
      // by pod count, emptiest first, so most pods land here (C2: 63 % on the first 16 claims, 93 % on the first 80).
      // Selector-less pods take the convergent threshold check and the winning group keeps the claim's new instance
      // types for the commit; other pods use one thread per claim like the later rounds.
      // (Running the first round on four warps only -- one per SM sub-partition -- was measured slower: 40.9 vs
      // 38.3 ms on C2, the extra round costs more than the issue contention.)
      // SMALL (selector-less pods, few of them): lane-group rounds all the way -- the one-thread-per-claim code and the
      // claim recomputation that goes with it are compiled out, and for this kernel less code is more speed
      constexpr int GPW = 32 / G, GROUP_ROUNDS = SMALL ? (1 << 20) : 2, SW = QW;  // SW scanning warps: all but the queue warp
      const int ROUND0 = podkeys ? 0 : 2 * SW * GPW;  // pods with selectors go straight to one thread per claim
      // every "did anyone hit" decision is taken with a barrier reduction so that it is uniform across the CTA
      int any_hit = 0;
      if (!podkeys) {
        for (int gr = 0; gr < GROUP_ROUNDS && !any_hit; gr++) {
          if (gr * SW * GPW >= n) break;  // uniform
          bool fits = false;
          if (warp != QW) {  // the queue warp has no claims to look at; it meets the others at the same barrier below
            const int idx = gr * SW * GPW + warp * GPW + (lane / G);
            const bool live = idx < n;
            const int c = live ? ord[idx] : 0;
            // a pod that tolerates every taint any template carries needs no per-claim lookup (uniform)
            const bool tolerated = (tol_all || !live) ? true : ((S.tm.taints[c_tmpl[c]] & ~pd.tol) == 0);
            uint64_t o_[WPL];
            bool chg_;
            int64_t thr_;
            fits = group_inflight_nokeys<G, WPL, TS>(kt, q_rows, lres, live, c_req + (size_t)c * R, c_hi + (size_t)c * R,
                                                     c_thr + (size_t)c * R, c_mask + (size_t)c * W, tolerated, pd.req, o_,
                                                     chg_, thr_);
            if (live && gl == 0 && chg_) my_evals += (unsigned long long)c_nopt[c];  // types actually re-filtered
            if (fits) {
              my_idx0 = idx;
              my_c0 = c;
              my_changed0 = chg_;
              my_thr0 = thr_;
#pragma unroll
              for (int i = 0; i < WPL; i++) out0[i] = o_[i];
              if (gl == 0) atomicMin(&S.best[cur], idx);
            }
          }
          any_hit = __syncthreads_or(fits ? 1 : 0);
        }
      }
      // (Measured and dropped for the long scans of config[2]-like inputs: a per-claim "refuses this signature" memo
      // -- no gain, the pass is bound by its dependent loads, not by the requirement algebra -- and four blocks per
      // warp iteration with their loads issued together -- the four inlined copies of the check made the whole kernel
      // slower, 1.9 -> 2.4 s on 100k x 1k and 33 -> 36 ms on config[1].)
      if (!SMALL && !any_hit && n > ROUND0) {
        for (int it = 0; warp != QW; it++) {
          const int b0 = ROUND0 + (it * SW + warp) * 32;
          int best_now = 0;
          if (lane == 0) best_now = *(volatile int32_t*)&S.best[cur];
          best_now = __shfl_sync(0xffffffffu, best_now, 0);  // one read per warp keeps the loop exit uniform
          if (b0 >= n || b0 > best_now) break;
          if (scan_block(b0)) break;  // later blocks of this warp come later in order
        }
        __syncthreads();
      }
      found = S.best[cur];
    }
    PROF(2);

    // a step whose scan had no barrier (existing node, or no claim yet) still has to order the queue warp's early
    // reads of this step against the commit's writes
    if (on_node || n == 0) __syncthreads();

    // ---- commit: exactly one of (node, in-flight claim, new claim, push back).  Warp 0 handles the queue and the
    // next pod's row, the first lane group of warp 1 the claim state, warp 2 the claim order. ----
    if (tid == 0) S.best[cur ^ 1] = INT_MAX;
    int t_new = -1;
    if (!on_node && found == INT_MAX) {
      if (any_limit) {  // the walk depends on the claims opened so far: one warp does it now (uniform branch)
        if (warp == 1) {
          const int tw = warp_limited_walk(kt, S, A, pods, p, W, lim_rows);
          if (lane == 0) S.t_new = tw;
        }
        __syncthreads();
        t_new = S.t_new;
      } else {
        t_new = A.a_tmpl[p];
      }
    }
    if (on_node) {
      const int row = E0 + found;  // ExistingNode.Add (existingnode.go:111-119)
      if (warp == 1) {
        if (lane < R) A.n_avail[(size_t)lane * A.Etot + row] -= pd.req[lane];
      } else if (KEYS && warp == 2) {  // Requirements.Add of a pod without requirements leaves the node's labels alone
        const int k = lane;
        uint32_t nib_lo = 0, nib_hi = 0;
        if (k < K) {
          StateRef st{A.n_rmask + row, (size_t)A.Etot, A.n_rflags[row], S.zero_b, S.zero_b, 1, nullptr, 0, 0};
          EReq m = add_req(kt, k, state_req(kt, st, k), pod_req(kt, pd, k));
          A.n_rmask[(size_t)k * A.Etot + row] = m.mask;
          uint32_t nib = m.f & 0x3u;  // nodes never keep bounds: a concrete set drops them (requirement.go:186-189)
          if (k < 8) nib_lo = nib << (4 * k); else nib_hi = nib << (4 * (k - 8));
        }
        nib_lo = __reduce_or_sync(0xffffffffu, nib_lo);
        nib_hi = __reduce_or_sync(0xffffffffu, nib_hi);
        if (k == 0) A.n_rflags[row] = (uint64_t)nib_lo | ((uint64_t)nib_hi << 32);
      }
      if (tid == 0) {
        A.n_npods[row]++;
        A.pod_kind[p] = B200S_PLACE_NODE;
        A.pod_index[p] = found;
        A.pod_seq[p] = S.seq++;
      }
    } else if (found != INT_MAX) {
      const int r0sz = podkeys ? 0 : 2 * QW * (32 / G);  // claims covered by the lane-group rounds (ROUND0 of the scan)
      const bool by_group = SMALL || (!podkeys && found < r0sz);  // found by a lane group that still holds the new instance types
      if (by_group) {
        if (my_idx0 == found) {  // NodeClaim.Add (nodeclaim.go:168-180) by the group that evaluated the claim
          const long long t_c0 = PROFILE ? clock64() : 0;
          const int c = my_c0;
          if (my_changed0) {  // some resource moved past its threshold: the claim's types shrank
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < WPL; i++) {
              int w = gl + i * G;
              if (w < W) c_mask[(size_t)c * W + w] = out0[i];
              cnt += popc64(out0[i]);
            }
            cnt = (int)__reduce_add_sync(group_mask<G>(), (unsigned)cnt);
            if (gl < R && my_thr0 != 0) c_thr[(size_t)c * R + gl] = my_thr0;
            if (gl == 0) c_nopt[c] = cnt;
          }
          for (int r = gl; r < R; r += G) c_req[(size_t)c * R + r] += pd.req[r];
          if (gl == 0) {
            c_npods[c]++;
            A.pod_kind[p] = B200S_PLACE_CLAIM;
            A.pod_index[p] = c;
            A.pod_seq[p] = S.seq++;
            if (PROFILE) S.wprof[2] += (unsigned long long)(clock64() - t_c0);
          }
        }
      } else if (!SMALL && warp == 1 && lane < G) {  // found by the one-thread-per-claim scan: one lane group redoes the claim in full
        const int c = S.hit_c[found < r0sz ? (found >> 5) : (((found - r0sz) >> 5) % QW)];  // not ord[found]: warp 2 reorders ord
        StateRef st{c_rmask + (size_t)c * K, 1, podkeys ? c_rflags[c] : 0ull, c_gt + (size_t)c * NB, c_lt + (size_t)c * NB, 1,
                    c_req + (size_t)c * R, 1, 0};
        uint64_t cand[WPL], out[WPL];
#pragma unroll
        for (int i = 0; i < WPL; i++) {
          int w = gl + i * G;
          cand[i] = w < W ? c_mask[(size_t)c * W + w] : 0ull;
        }
        int visited;
        group_inflight<G, WPL>(kt, st, c_hi + (size_t)c * R, pd, podkeys, cand, out, &visited);
        if (podkeys)
          group_store_merged_keys<G>(kt, st, pd, podkeys, c_rmask + (size_t)c * K, &c_rflags[c], c_gt + (size_t)c * NB,
                                     c_lt + (size_t)c * NB);
        int cnt = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
          int w = gl + i * G;
          if (w < W) c_mask[(size_t)c * W + w] = out[i];
          cnt += popc64(out[i]);
        }
        cnt = (int)__reduce_add_sync(group_mask<G>(), (unsigned)cnt);
        for (int r = gl; r < R; r += G) {
          const int64_t x = c_req[(size_t)c * R + r] + pd.req[r];
          c_req[(size_t)c * R + r] = x;
          // the pod fits, so some allocatable value is >= x
          c_thr[(size_t)c * R + r] = ldg(kt.res_dv + kt.res_dv0[r] +
                                         fit_rank_bkt(kt.res_dv + kt.res_dv0[r], kt.res_bkt + (size_t)r * (RANK_BKTS + 1), x));
        }
        if (gl == 0) {
          c_nopt[c] = cnt;
          c_npods[c]++;
          A.pod_kind[p] = B200S_PLACE_CLAIM;
          A.pod_index[p] = c;
          A.pod_seq[p] = S.seq++;
        }
      }
      if (warp == 2) {  // the claim moved one pod up: what sort.Slice will do with that
        const long long t_n0 = PROFILE ? clock64() : 0;
        if (lane == 0) key[found]++;
        __syncwarp();
        warp_normalise(S, ord, key, g_snap, n, found, false);
        if (PROFILE && lane == 0) S.wprof[1] += (unsigned long long)(clock64() - t_n0);
      }
    } else if (t_new >= 0) {
      // ---- addToNewNodeClaim: the template walk was done for every pod by stage A ----
      const int t = t_new, c = n;
      if (warp == 1 && lane < G) {  // NewNodeClaim + Add (nodeclaim.go:83-109, :168-180)
        StateRef st{S.tm.rmask[t], 1, S.tm.rflags[t], S.tm.gt[t], S.tm.lt[t], 1, S.tm.daemon[t], 1, 0};
        if constexpr (KEYS) {
          group_store_merged<G>(kt, st, pd, c_rmask + (size_t)c * K, 1, &c_rflags[c], c_gt + (size_t)c * NB,
                                c_lt + (size_t)c * NB, 1);
        } else {  // the pod defines no key: Requirements.Add leaves the template's requirements as they are
          for (int k = gl; k < K; k += G) {
            c_rmask[(size_t)c * K + k] = S.tm.rmask[t][k];
            const int b = kt.bslot[k];
            if (b >= 0) {
              const uint32_t nib = (uint32_t)((S.tm.rflags[t] >> (4 * k)) & 0xFu);
              c_gt[(size_t)c * NB + b] = (nib & RF_GT) ? S.tm.gt[t][b] : 0;
              c_lt[(size_t)c * NB + b] = (nib & RF_LT) ? S.tm.lt[t][b] : 0;
            }
          }
          if (gl == 0) c_rflags[c] = S.tm.rflags[t];
        }
        int cnt = 0;
        for (int w = gl; w < W; w += G) {
          uint64_t m = new_claim_word(A, pods, S, lim_rows, p, t, W, w);
          c_mask[(size_t)c * W + w] = m;
          cnt += popc64(m);
        }
        cnt = (int)__reduce_add_sync(group_mask<G>(), (unsigned)cnt);
        for (int r = gl; r < R; r += G) {
          const int64_t x = S.tm.daemon[t][r] + pd.req[r];
          c_req[(size_t)c * R + r] = x;
          c_thr[(size_t)c * R + r] = ldg(kt.res_dv + kt.res_dv0[r] +
                                         fit_rank_bkt(kt.res_dv + kt.res_dv0[r], kt.res_bkt + (size_t)r * (RANK_BKTS + 1), x));
        }
        if (gl == 0) {
          c_tmpl[c] = t;
          c_npods[c] = 1;
          c_nopt[c] = cnt;
        }
      } else if (warp >= 3 && warp < QW) {
        // largest Allocatable() among the claim's instance types, per resource: the highest table row
        // "alloc >= dv[i]" that still meets the mask (the mask only shrinks afterwards, so this stays an upper
        // bound).  One lane group per resource, beside the group that builds the claim.
        constexpr int HI_GROUPS = (QW - 3) * 32 / G;
        for (int r = ((warp - 3) * 32 + lane) / G; r < R; r += HI_GROUPS) {
          int lo = 0, hi = kt.res_nd[r] - 1;  // row 0 always meets a non-empty mask
          while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            unsigned hit = 0;
            for (int w = gl; w < W; w += G)
              hit |= (new_claim_word(A, pods, S, lim_rows, p, t, W, w) & ldg(kt.rows + (size_t)(kt.res_row0[r] + mid) * W + w)) ? 1u : 0u;
            hit = __reduce_or_sync(group_mask<G>(), hit);
            if (hit) lo = mid; else hi = mid - 1;
          }
          if (gl == 0) c_hi[(size_t)c * R + r] = ldg(kt.res_dv + kt.res_dv0[r] + lo);
        }
      } else if (warp == 2) {
        if (lane == 0) {
          ord[n] = c;
          key[n] = 1;
          S.n_claims = n + 1;
        }
        __syncwarp();
        warp_normalise(S, ord, key, g_snap, n + 1, n, true);
      }
      if (tid == 0) {
        A.pod_kind[p] = B200S_PLACE_CLAIM;
        A.pod_index[p] = c;
        A.pod_seq[p] = S.seq++;
      }
    } else if (warp == QW && lane == 0) {  // Queue.Push (queue.go:63-66), by the queue warp ahead of its pop; the pod stays in PodErrors
      int tail = S.head + S.len;
      if (tail >= n_pods) tail -= n_pods;
      queue[tail] = p;
      S.len++;
      S.n_pushed++;
      // lastLen = the queue's length right after the append (queue.go:63-66); a next pod popped ahead of time is
      // still in the reference's queue at that moment
      A.lastlen[p] = S.len + (spec_popped ? 1 : 0);
      A.pod_kind[p] = B200S_PLACE_UNSCHEDULED;
      A.pod_index[p] = -1;
    }
    if (warp == QW) {
      // pop here when pods have been pushed back before (the lastLen rule needs this pod's outcome first), or when
      // the early pop found the queue empty and this pod has just been appended to it
      const bool pushed_now = !on_node && found == INT_MAX && t_new < 0;
      __syncwarp();
      if (!fast || (pushed_now && !spec_popped)) {
        const long long t_p0 = PROFILE ? clock64() : 0;
        pop_and_load(cur ^ 1);
        __syncwarp();
        if (PROFILE && lane == 0) S.wprof[0] += (unsigned long long)(clock64() - t_p0);
      }
    }
    __syncthreads();
    PROF(3);
    if (any_limit && t_new >= 0 && S.tm.limit_mask[t_new]) {  // uniform: the new claim counts against its NodePool's limits
      if (warp == 1) {
        warp_subtract_max(kt, S, A, pods, lim_rows, p, t_new, W);
        warp_limit_row(kt, S, t_new, W, lim_rows);
      }
      __syncthreads();
    }

    // ---- what is left of the order normalisation: long rotates and the full pdqsort's bookkeeping ----
    const int mode = S.mode;  // written by warp 2 before the barrier above; nobody writes it before the next one
    if (mode != MODE_NONE) {  // uniform
      const int src = S.rot_src, dst = S.rot_dst, orig = S.rot_orig;
      if (mode == MODE_ROTATE) cta_rotate(ord, key, src, dst);
      __syncthreads();  // every thread has read S.mode: only now may it be reset (a late warp that saw MODE_NONE
                        // would skip this block and fall one barrier behind)
      if (tid == 0) {
        S.undo_mode = mode;
        S.undo_src = dst;
        S.undo_dst = orig;
        S.mode = MODE_NONE;
      }
      __syncthreads();
    }
    if (!SMALL && !S.ord_global && S.n_claims >= ORD_SMEM_CAP) {  // spill the order arrays to global memory
      for (int i = tid; i < S.n_claims; i += PACK_THREADS) { g_ord[i] = s_ord[i]; g_key[i] = s_key[i]; }
      __syncthreads();
      if (tid == 0) S.ord_global = 1;
      __syncthreads();
    }
    PROF(4);
    cur ^= 1;
  }

  // Results.NewNodeClaims is the order after the LAST sort the reference ran plus the modification that followed it:
  // if no later add() reached the sort, the eager normalisation of that last modification has to be taken back.
  {
    const bool og = SMALL ? false : (S.ord_global != 0);
    int32_t* ord = og ? g_ord : s_ord;
    int32_t* key = og ? g_key : s_key;
    const int um = S.undo_mode, n = S.n_claims;
    __syncthreads();
    if (um == MODE_ROTATE) cta_rotate(ord, key, S.undo_src, S.undo_dst);
    else if (um == MODE_GENERIC)
      for (int i = tid; i < n; i += PACK_THREADS) ord[i] = g_snap[i];
    __syncthreads();
    if (!og)
      for (int i = tid; i < n; i += PACK_THREADS) g_ord[i] = s_ord[i];
  }
  for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
  if ((tid & 31) == 0 && my_evals) atomicAdd(&S.evals, my_evals);
  __syncthreads();
  if (tid == 0) {
    pb.n_claims = S.n_claims;
    pb.n_unscheduled = S.len;
    pb.evals = S.evals;
    for (int i = 0; i < 8; i++) pb.prof[i] = prof[i];
    for (int i = 0; i < 4; i++) pb.prof[8 + i] = S.wprof[i];
  }
}

// Results.NewNodeClaims in final order + FinalizeScheduling (nodeclaim.go:252-268: only the placeholder
// hostname is dropped, which this encoding never materialises) + pod -> claim position.  gridDim.y CTAs share a
// problem's rows and pods; every one of them writes the (identical) claim -> position map before it uses it.
__global__ void __launch_bounds__(256) pack_finalize_kernel(KeyTab kt, PodsDev pods, PackArrays A,
                                                            const PackProblem* __restrict__ problems, ClaimOut out) {
  const PackProblem& pb = problems[blockIdx.x];
  const int base = pb.pod_begin, n_pods = pb.pod_end - pb.pod_begin, n = pb.n_claims;
  const int R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded, W = kt.n_words;
  const int row_words = W + R + K + 2 * NB;
  const int tid = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;
  for (int i = tid; i < n * row_words; i += nth) {
    int pos = i / row_words, j = i - pos * row_words;
    int c = A.ord[base + pos];
    size_t src = (size_t)base + c, dst = (size_t)base + pos;
    if (j < W) out.mask[dst * W + j] = A.c_mask[src * W + j];
    else if ((j -= W) < R) out.req[dst * R + j] = A.c_req[src * R + j];
    else if ((j -= R) < K) out.rmask[dst * K + j] = A.c_rmask[src * K + j];
    else if ((j -= K) < NB) out.gt[dst * NB + j] = A.c_gt[src * NB + j];
    else { j -= NB; out.lt[dst * NB + j] = A.c_lt[src * NB + j]; }
  }
  for (int pos = threadIdx.x; pos < n; pos += blockDim.x) {
    int c = A.ord[base + pos];
    if (blockIdx.y == 0) {
      out.tmpl[base + pos] = A.c_tmpl[base + c];
      out.npods[base + pos] = A.c_npods[base + c];
      out.rflags[base + pos] = A.c_rflags[base + c];
    }
    out.pos_of_claim[base + c] = pos;
  }
  __syncthreads();
  for (int i = tid; i < n_pods; i += nth) {
    int p = base + i;  // pods of a problem are a contiguous index range (the queue ring is NOT a permutation any more)
    if (A.pod_kind[p] == B200S_PLACE_CLAIM) A.pod_index[p] = out.pos_of_claim[base + A.pod_index[p]];
  }
  (void)pods;
}

static void pick_group_pack(int W, int* G, int* WPL) {
  if (W <= 8) { *G = 8; *WPL = 1; }
  else if (W <= 16) { *G = 16; *WPL = 1; }
  else if (W <= 32) { *G = 32; *WPL = 1; }
  else if (W <= 64) { *G = 32; *WPL = 2; }
  else { *G = 32; *WPL = 4; }
}

cudaError_t launch_pack(const KeyTab& kt, const TmplDev* tg, int T, const PodsDev& pods, PackArrays arr,
                        PackProblem* problems, int n_problems, bool keys, bool nodes, int max_pods, bool profile,
                        int smem_limit, cudaStream_t stream) {
  (void)T;
  if (kt.n_words > 128) return cudaErrorInvalidValue;
  if (n_problems == 0) return cudaSuccess;
  int G, WPL;
  pick_group_pack(kt.n_words, &G, &WPL);
  size_t so, to;
  // the resource tables go to shared memory when they fit next to the order arrays; a batch whose problems all have
  // few pods keeps a larger order array there and never spills it
  int in_smem = 1;
  bool small = max_pods <= ORD_SMALL_CAP && !keys && !nodes;
  size_t bytes = pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, in_smem, &so, &to, nullptr, small ? ORD_SMALL_CAP : ORD_SMEM_CAP);
  if (small && bytes > (size_t)smem_limit) {
    small = false;
    bytes = pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, in_smem, &so, &to);
  }
  if (bytes > (size_t)smem_limit || getenv("B200S_PACK_NO_TS")) {
    in_smem = 0;
    small = false;
    bytes = pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, in_smem, &so, &to);
  }
  cudaError_t err = cudaSuccess;
#define PICK(G_, WPL_, K_, N_, S_) \
  (profile ? pack_kernel<G_, WPL_, true, K_, N_, S_, true> : pack_kernel<G_, WPL_, true, K_, N_, S_, false>)
#define CALL(G_, WPL_)                                                                            \
  do {                                                                                            \
    auto kern = !in_smem ? pack_kernel<G_, WPL_, false, true, true, false, true>                 \
                : nodes  ? (keys ? PICK(G_, WPL_, true, true, false) : PICK(G_, WPL_, false, true, false)) \
                : keys   ? PICK(G_, WPL_, true, false, false)                                    \
                : small  ? PICK(G_, WPL_, false, false, true)                                    \
                         : PICK(G_, WPL_, false, false, false);                                  \
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);    \
    if (err != cudaSuccess) return err;                                                           \
    kern<<<n_problems, PACK_THREADS, bytes, stream>>>(kt, tg, pods, arr, problems, in_smem);      \
  } while (0)
  if (G == 8) CALL(8, 1);
  else if (G == 16) CALL(16, 1);
  else if (WPL == 1) CALL(32, 1);
  else if (WPL == 2) CALL(32, 2);
  else CALL(32, 4);
#undef CALL
#undef PICK
  return cudaGetLastError();
}

cudaError_t launch_pack_finalize(const KeyTab& kt, const PodsDev& pods, PackArrays arr, const PackProblem* problems,
                                 int n_problems, int max_pods, ClaimOut out, cudaStream_t stream) {
  if (n_problems == 0) return cudaSuccess;
  // a single large Solve spreads over many CTAs, a batch of small ones keeps one each
  int y = max_pods / 512;
  y = y < 1 ? 1 : y > 64 ? 64 : y;
  while (y > 1 && (long long)n_problems * y > 1184) y >>= 1;
  pack_finalize_kernel<<<dim3((unsigned)n_problems, (unsigned)y), 256, 0, stream>>>(kt, pods, arr, problems, out);
  return cudaGetLastError();
}

}  // namespace b200

// b200_core.cuh — the arithmetic of the fit-and-pack path on the bitset encoding, as plain
// host/device functions.  Kernels (fit_kernels.cu, pack_kernels.cu) wrap these in lane-cooperative
// code; the encoder (encoder.cpp) uses the same algebra on the host so that templates are built with
// exactly the semantics the kernels evaluate.
//
// Encoding recap (include/b200sched.h, Level 0): a requirement on device key k is
//   mask  : u64, bit v = v-th value of key k's universe (the In set, or the NotIn exclusion set)
//   flags : DEFINED | COMPLEMENT | HASGT | HASLT
//   gt/lt : integer bounds (only complement requirements carry them: requirement.go:186-189)
// An undefined key reads as Exists (requirements.go:154-160).
#pragma once
#include <stdint.h>

#include "../../include/b200sched.h"

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200 {

constexpr uint32_t RF_DEF = B200S_RF_DEFINED, RF_CMPL = B200S_RF_COMPLEMENT, RF_GT = B200S_RF_HASGT,
                   RF_LT = B200S_RF_HASLT;
constexpr int MAXK = B200S_MAX_KEYS, MAXR = B200S_MAX_RES, MAXT = B200S_MAX_TEMPLATES, MAXB = B200S_MAX_BOUNDED;

struct EReq {
  uint64_t mask;
  int64_t gt, lt;
  uint32_t f;
};

// Per-catalog constants every kernel needs (lives in kernel-parameter / constant space).
struct KeyTab {
  int32_t n_keys, n_res, n_words, n_skus;
  int32_t key_zone, key_ct, n_zones, n_cts;
  uint32_t wellknown;
  int32_t n_bounded;
  int32_t nvals[MAXK];
  int8_t bslot[MAXK];      // bounded slot of key k, or -1
  uint64_t isint[MAXK];    // bit v: value v parses as an integer
  const int64_t* val_int;  // [K*64]
  // tables
  const uint64_t* rows;    // [n_rows*W]
  const int64_t* res_dv;   // ascending distinct allocatable values
  int32_t res_nd[MAXR], res_dv0[MAXR], res_row0[MAXR];
  int32_t res_p2[MAXR];    // largest power of two <= res_nd[r] (0 when the resource has no values)
  int32_t max_p2;          // max over resources
  const uint16_t* res_bkt; // [MAXR][RANK_BKTS + 1]: fit_rank of each log-scale bucket's lowest value (rank_bucket)
  int32_t key_row0[MAXK];
  int32_t zc_row0, row_nonneg, n_rows, n_dv;
  const uint64_t* price;   // [n_zc*S]
  const uint64_t* price_max;  // [n_zc*S] most expensive available offering, B200S_PRICE_ABSENT if none (may be null)
  // NodePool limits: ascending distinct Capacity values per resource and the rows "capacity[r] <= cap_dv[r][i]"
  int32_t cap_nd[MAXR], cap_dv0[MAXR], cap_row0[MAXR];
  const int64_t* cap_dv;
};

// read-only table loads: through the non-coherent (L1-cached) path on the device
template <class Tp>
B2_HD Tp ldg(const Tp* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

B2_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}

B2_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

B2_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

B2_HD uint64_t universe_of(const KeyTab& kt, int k) {
  int n = kt.nvals[k];
  return n >= 64 ? ~0ull : ((1ull << n) - 1ull);
}

// withinIntPtrs over the whole universe of key k (requirement.go:311-327): bit v set iff value v is
// an integer strictly inside (gt, lt).  No bounds -> every value is within.
B2_HD uint64_t within_mask(const KeyTab& kt, int k, uint32_t f, int64_t gt, int64_t lt) {
  if (!(f & (RF_GT | RF_LT))) return ~0ull;
  uint64_t m = 0, cand = kt.isint[k];
  while (cand) {
    int v = ctz64(cand);
    cand &= cand - 1;
    int64_t x = kt.val_int[k * 64 + v];
    bool ok = true;
    if ((f & RF_GT) && gt >= x) ok = false;
    if ((f & RF_LT) && lt <= x) ok = false;
    if (ok) m |= 1ull << v;
  }
  return m;
}

// Operator() in {NotIn, DoesNotExist} (requirement.go:267-278): the two operators requirements.go
// :183 and :254-258 forgive.
B2_HD bool forgiving(const EReq& q) { return (q.f & RF_CMPL) ? (q.mask != 0) : (q.mask == 0); }

B2_HD void merge_bounds(const EReq& a, const EReq& b, uint32_t* f, int64_t* gt, int64_t* lt) {
  bool ag = a.f & RF_GT, bg = b.f & RF_GT, al = a.f & RF_LT, bl = b.f & RF_LT;
  *gt = ag ? (bg ? (a.gt > b.gt ? a.gt : b.gt) : a.gt) : b.gt;
  *lt = al ? (bl ? (a.lt < b.lt ? a.lt : b.lt) : a.lt) : b.lt;
  *f = ((ag || bg) ? RF_GT : 0u) | ((al || bl) ? RF_LT : 0u);
}

// Requirement.HasIntersection (requirement.go:197-231); both sides defined.
B2_HD bool has_intersection(const KeyTab& kt, int k, const EReq& a, const EReq& b) {
  uint32_t bf;
  int64_t gt, lt;
  merge_bounds(a, b, &bf, &gt, &lt);
  if ((bf & RF_GT) && (bf & RF_LT) && gt >= lt) return false;
  bool ca = a.f & RF_CMPL, cb = b.f & RF_CMPL;
  if (ca && cb) return true;
  uint64_t w = within_mask(kt, k, bf, gt, lt);
  if (ca) return (b.mask & ~a.mask & w) != 0;
  if (cb) return (a.mask & ~b.mask & w) != 0;
  return (a.mask & b.mask & w) != 0;
}

// Requirement.Intersection (requirement.go:158-191); both sides defined.
B2_HD EReq intersection(const KeyTab& kt, int k, const EReq& a, const EReq& b) {
  uint32_t bf;
  int64_t gt, lt;
  merge_bounds(a, b, &bf, &gt, &lt);
  EReq o;
  if ((bf & RF_GT) && (bf & RF_LT) && gt >= lt) {  // -> DoesNotExist
    o.mask = 0; o.gt = 0; o.lt = 0; o.f = RF_DEF;
    return o;
  }
  bool ca = a.f & RF_CMPL, cb = b.f & RF_CMPL;
  uint64_t values = (ca && cb) ? (a.mask | b.mask) : ca ? (b.mask & ~a.mask) : cb ? (a.mask & ~b.mask) : (a.mask & b.mask);
  values &= within_mask(kt, k, bf, gt, lt);
  o.mask = values;
  if (ca && cb) {
    o.f = RF_DEF | RF_CMPL | bf;
    o.gt = (bf & RF_GT) ? gt : 0;
    o.lt = (bf & RF_LT) ? lt : 0;
  } else {
    o.f = RF_DEF;
    o.gt = 0; o.lt = 0;
  }
  return o;
}

// Requirements.Add for one key (requirements.go:127-134): either side may be undefined.
B2_HD EReq add_req(const KeyTab& kt, int k, const EReq& existing, const EReq& incoming) {
  if (!(incoming.f & RF_DEF)) return existing;
  if (!(existing.f & RF_DEF)) return incoming;
  return intersection(kt, k, incoming, existing);
}

// One key of Requirements.Compatible(state, pod, allow) (requirements.go:175-191, :248-268).
B2_HD bool compatible_key(const KeyTab& kt, int k, const EReq& state, const EReq& pod, bool allowWellKnown) {
  if (!(pod.f & RF_DEF)) return true;
  if (!(state.f & RF_DEF)) {
    if (allowWellKnown && ((kt.wellknown >> k) & 1u)) return true;
    return forgiving(pod);
  }
  if (has_intersection(kt, k, state, pod)) return true;
  return forgiving(pod) && forgiving(state);
}

// Values of key k a (defined) requirement admits, inside the universe.
B2_HD uint64_t allowed_values(const KeyTab& kt, int k, const EReq& q) {
  if (q.f & RF_CMPL) return ~q.mask & within_mask(kt, k, q.f, q.gt, q.lt) & universe_of(kt, k);
  return q.mask;
}

// lower_bound over the ascending distinct allocatable values of one resource: index of the first
// value >= x, or nd when x exceeds them all (then no instance type fits).
// Log-scale bucket of a non-negative quantity: exponent and the two bits below the leading one, monotone in x.
// res_bkt[r][b] = fit_rank(lowest value of bucket b), so fit_rank(x) lies in [res_bkt[r][b], res_bkt[r][b + 1]] and
// the search inside a bucket is one or two steps for real catalogs (distinct allocatable values are rarely within
// 25 % of each other).
constexpr int RANK_BKTS = 256;
B2_HD int rank_bucket(int64_t x) {
  if (x <= 0) return 0;
#if defined(__CUDA_ARCH__)
  const int e = 63 - __clzll(x);
#else
  const int e = 63 - __builtin_clzll((unsigned long long)x);
#endif
  const int m = e >= 2 ? (int)((x >> (e - 2)) & 3) : (int)((x << (2 - e)) & 3);
  return e * 4 + m;
}
B2_HD int64_t rank_bucket_low(int b) {  // smallest x with rank_bucket(x) == b (b >= 1 reachable buckets only)
  const int e = b >> 2, m = b & 3;
  if (e >= 2) return ((int64_t)(4 + m)) << (e - 2);
  return ((int64_t)(4 + m)) >> (2 - e);
}

B2_HD int fit_rank(const int64_t* dv, int nd, int64_t x) {
  int lo = 0, hi = nd;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (ldg(dv + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// The bucket table of one resource (RANK_BKTS + 1 entries), built on the host at catalog upload.
inline void rank_bucket_table(const int64_t* dv, int nd, uint16_t* row) {
  row[RANK_BKTS] = (uint16_t)nd;
  for (int b = RANK_BKTS - 1; b >= 1; b--) {
    // a bucket no value maps to (the mantissa bits of 1, 2, 3 are not all used) takes the next one's rank, so
    // [row[b], row[b + 1]] always brackets the rank of every x in bucket b
    const int64_t low = rank_bucket_low(b);
    row[b] = rank_bucket(low) == b ? (uint16_t)fit_rank(dv, nd, low) : row[b + 1];
  }
  row[0] = 0;
}

// fit_rank through the bucket table of resource r
// (PLAIN: the tables were staged into shared memory, where the non-coherent load path does not exist)
template <bool PLAIN = false>
B2_HD int fit_rank_bkt(const int64_t* dv, const uint16_t* bkt, int64_t x) {
  const int b = rank_bucket(x);
  int lo = PLAIN ? bkt[b] : ldg(bkt + b), hi = PLAIN ? bkt[b + 1] : ldg(bkt + b + 1);
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((PLAIN ? dv[mid] : ldg(dv + mid)) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- one SKU word of the filter loop (nodeclaim.go:392-423), 64 instance types at a time ----------
// Scratch a group prepares once per (state, pod) pair and every word then reads.
struct EvalScratch {
  uint64_t allowed[MAXK];  // admitted values per key of the merged requirements
  int32_t fit_row[MAXR];   // table row per resource, or -1 when nothing fits
  uint32_t defined;        // bit k: merged requirements define key k
  uint32_t forgive;        // bit k: merged requirement k is NotIn / DoesNotExist
  uint64_t zone_allowed, ct_allowed;  // offering filter (all ones when undefined)
};

struct WordEval {
  uint64_t compat, fits, off;
};

// fits (nodeclaim.go:447-449 -> resources.go:232-245)
B2_HD uint64_t eval_fits_word(const KeyTab& kt, const int32_t* fit_row, int w) {
  const int W = kt.n_words;
  const uint64_t* rows = kt.rows;
  uint64_t fits = ldg(rows + (size_t)kt.row_nonneg * W + w);
  for (int r = 0; r < kt.n_res; r++) {
    int row = fit_row[r];
    fits &= row < 0 ? 0ull : ldg(rows + (size_t)row * W + w);
  }
  return fits;
}

// SKUs whose requirement on key k intersects a requirement that admits `allowed` (requirement.go:197-231 for a
// concrete SKU set), or that do not define k, or — when the requirement is NotIn/DoesNotExist — that say
// DoesNotExist (requirements.go:253-258)
B2_HD uint64_t eval_key_word(const KeyTab& kt, int k, uint64_t allowed, bool forgive, int w) {
  const int W = kt.n_words;
  const uint64_t* rows = kt.rows;
  int row0 = kt.key_row0[k], nv = kt.nvals[k];
  uint64_t acc = ldg(rows + (size_t)(row0 + nv) * W + w);
  if (forgive) acc |= ldg(rows + (size_t)(row0 + nv + 1) * W + w);
  while (allowed) {
    int v = ctz64(allowed);
    allowed &= allowed - 1;
    acc |= ldg(rows + (size_t)(row0 + v) * W + w);
  }
  return acc;
}

// compatible (nodeclaim.go:443-445 -> requirements.go:248-268)
B2_HD uint64_t eval_compat_word(const KeyTab& kt, const EvalScratch& sc, int w) {
  uint64_t compat = ~0ull;
  uint32_t def = sc.defined;
  while (def) {
    int k = ctz32(def);
    def &= def - 1;
    compat &= eval_key_word(kt, k, sc.allowed[k], (sc.forgive >> k) & 1u, w);
  }
  return compat;
}

// offering scan (nodeclaim.go:400-406): some Available offering whose zone and capacity type are admitted
B2_HD uint64_t eval_off_word(const KeyTab& kt, uint64_t zone_allowed, uint64_t ct_allowed, int w) {
  const int W = kt.n_words;
  const uint64_t* rows = kt.rows;
  uint64_t off = 0;
  uint64_t za = zone_allowed & (kt.n_zones >= 64 ? ~0ull : ((1ull << kt.n_zones) - 1ull));
  const uint64_t cmask = kt.n_cts >= 64 ? ~0ull : ((1ull << kt.n_cts) - 1ull);
  while (za) {
    int z = ctz64(za);
    za &= za - 1;
    uint64_t ca = ct_allowed & cmask;
    while (ca) {
      int c = ctz64(ca);
      ca &= ca - 1;
      off |= ldg(rows + (size_t)(kt.zc_row0 + z * kt.n_cts + c) * W + w);
    }
  }
  return off;
}

B2_HD WordEval eval_word(const KeyTab& kt, const EvalScratch& sc, int w) {
  WordEval e;
  e.fits = eval_fits_word(kt, sc.fit_row, w);
  e.compat = eval_compat_word(kt, sc, w);
  e.off = eval_off_word(kt, sc.zone_allowed, sc.ct_allowed, w);
  return e;
}

// The six diagnostic flags over a candidate set (nodeclaim.go:409-416), from OR-reduced words.
B2_HD uint8_t filter_flags(uint64_t anyC, uint64_t anyF, uint64_t anyO, uint64_t anyCFnO, uint64_t anyCOnF, uint64_t anyFOnC) {
  return (uint8_t)(B200S_FIT_FILTER | (anyC ? B200S_FF_REQUIREMENTS : 0) | (anyF ? B200S_FF_FITS : 0) |
                   (anyO ? B200S_FF_OFFERING : 0) | (anyCFnO ? B200S_FF_REQ_AND_FITS : 0) |
                   (anyCOnF ? B200S_FF_REQ_AND_OFF : 0) | (anyFOnC ? B200S_FF_FITS_AND_OFF : 0));
}

// ---- requirement rows as the kernels store them -----------------------------------------------
B2_HD EReq make_ereq(uint64_t mask, uint64_t flags_word, int k, int64_t gt, int64_t lt) {
  EReq q;
  q.mask = mask;
  q.f = (uint32_t)((flags_word >> (4 * k)) & 0xFu);
  q.gt = (q.f & RF_GT) ? gt : 0;
  q.lt = (q.f & RF_LT) ? lt : 0;
  return q;
}

// ---- offering prices under a requirement set's admitted zones / capacity types (za, ca: value bit sets) ----------
// OrderByPrice's key (types.go:221-240): cheapest available offering the requirements admit, MaxFloat64 if none
B2_HD uint64_t cheapest_price(const KeyTab& kt, uint64_t za, uint64_t ca, int s) {
  uint64_t p = B200S_PRICE_NONE;
  for (uint64_t z = za; z; z &= z - 1) {
    const int zi = ctz64(z);
    for (uint64_t c = ca; c; c &= c - 1) {
      uint64_t q = kt.price[(size_t)(zi * kt.n_cts + ctz64(c)) * kt.n_skus + s];
      p = q < p ? q : p;
    }
  }
  return p;
}

// Offerings.Available().WorstLaunchPrice(reqs) (types.go:460-474): the most expensive compatible offering of the
// first capacity type, in the order reserved, spot, on-demand, that has one
B2_HD uint64_t worst_launch_price(const KeyTab& kt, uint64_t za, uint64_t ca, int ct0, int ct1, int ct2,
                                                       int s) {
  const int order[3] = {ct0, ct1, ct2};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int c = order[i];
    if (c < 0 || !((ca >> c) & 1ull)) continue;
    uint64_t top = 0;
    bool any = false;
    for (uint64_t z = za; z; z &= z - 1) {
      uint64_t q = kt.price_max[(size_t)(ctz64(z) * kt.n_cts + c) * kt.n_skus + s];
      if (q != B200S_PRICE_ABSENT) { any = true; top = q > top ? q : top; }
    }
    if (any) return top;
  }
  return B200S_PRICE_NONE;
}

// ---- Go sort.Slice on the in-flight NodeClaim order (scheduler.go:499) ---------------------------
// ord[i] = claim id at position i of s.newNodeClaims, key[i] = len(Pods) of that claim (kept by
// position so a comparison is two adjacent-array reads).  Exact restatement of pdqsort_func
// (Go 1.24 src/sort/zsortfunc.go) specialised to less(i,j) = key[i] < key[j]; swaps move both arrays.
struct GoSort {
  int32_t* ord;
  int32_t* key;
  // coop (device only): ALL 32 lanes of a warp run the algorithm with identical control flow on arrays in shared or
  // global memory; lane 0 alone writes, and the long linear scans of partition / partialInsertionSort -- where a
  // nearly sorted input spends its time -- test 32 elements per step.
  bool coop = false;
  B2_HD bool less(int i, int j) const { return key[i] < key[j]; }
  B2_HD void swap(int i, int j) const {
#if defined(__CUDA_ARCH__)
    if (coop) {
      __syncwarp();  // every lane has finished the reads that led here
      if ((threadIdx.x & 31) == 0) swap1(i, j);
      __syncwarp();
      return;
    }
#endif
    swap1(i, j);
  }
  B2_HD void swap1(int i, int j) const {
    int32_t t = ord[i];
    ord[i] = ord[j];
    ord[j] = t;
    t = key[i];
    key[i] = key[j];
    key[j] = t;
  }
  // for (; i <= j && P(key[i], kv); i++);  P = "<" (le false) or "<=" (le true)
  B2_HD int fwd_while(int i, int j, int32_t kv, bool le) const {
#if defined(__CUDA_ARCH__)
    if (coop) {
      const int lane = threadIdx.x & 31;
      for (;; i += 32) {
        const int x = i + lane;
        const bool stop = x > j || !(le ? key[x] <= kv : key[x] < kv);
        const unsigned sb = __ballot_sync(0xffffffffu, stop);
        if (sb) return i + __ffs((int)sb) - 1;
      }
    }
#endif
    while (i <= j && (le ? key[i] <= kv : key[i] < kv)) i++;
    return i;
  }
  // for (; i <= j && P(key[j], kv); j--);  P = ">=" (gt false) or ">" (gt true)
  B2_HD int bwd_while(int i, int j, int32_t kv, bool gt) const {
#if defined(__CUDA_ARCH__)
    if (coop) {
      const int lane = threadIdx.x & 31;
      for (;; j -= 32) {
        const int x = j - lane;
        const bool stop = x < i || !(gt ? key[x] > kv : key[x] >= kv);
        const unsigned sb = __ballot_sync(0xffffffffu, stop);
        if (sb) return j - (__ffs((int)sb) - 1);
      }
    }
#endif
    while (i <= j && (gt ? key[j] > kv : key[j] >= kv)) j--;
    return j;
  }
  // for (; i < b && !(key[i] < key[i-1]); i++);
  B2_HD int fwd_while_nondecreasing(int i, int b) const {
#if defined(__CUDA_ARCH__)
    if (coop) {
      const int lane = threadIdx.x & 31;
      for (;; i += 32) {
        const int x = i + lane;
        const bool stop = x >= b || key[x] < key[x - 1];
        const unsigned sb = __ballot_sync(0xffffffffu, stop);
        if (sb) return i + __ffs((int)sb) - 1;
      }
    }
#endif
    while (i < b && !less(i, i - 1)) i++;
    return i;
  }

  B2_HD void insertion(int a, int b) const {
#if defined(__CUDA_ARCH__)
    if (coop && b - a <= 32) {
      // insertion sort with a strict comparison is the stable sort: element L goes to a + (number of elements with a
      // smaller key, or an equal key and a smaller index)
      const int lane = threadIdx.x & 31, len = b - a;
      int32_t v = 0, kv = 0;
      if (lane < len) { v = ord[a + lane]; kv = key[a + lane]; }
      int rank = 0;
      for (int m = 0; m < len; m++) {
        const int32_t km = __shfl_sync(0xffffffffu, kv, m);
        rank += (km < kv || (km == kv && m < lane)) ? 1 : 0;
      }
      __syncwarp();
      if (lane < len) { ord[a + rank] = v; key[a + rank] = kv; }
      __syncwarp();
      return;
    }
#endif
    for (int i = a + 1; i < b; i++)
      for (int j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
  }
  B2_HD void sift_down(int lo, int hi, int first) const {
    int root = lo;
    for (;;) {
      int child = 2 * root + 1;
      if (child >= hi) return;
      if (child + 1 < hi && less(first + child, first + child + 1)) child++;
      if (!less(first + root, first + child)) return;
      swap(first + root, first + child);
      root = child;
    }
  }
  B2_HD void heap_sort(int a, int b) const {
    int first = a, lo = 0, hi = b - a;
    for (int i = (hi - 1) / 2; i >= 0; i--) sift_down(i, hi, first);
    for (int i = hi - 1; i >= 0; i--) {
      swap(first, first + i);
      sift_down(lo, i, first);
    }
  }
  B2_HD void order2(int& a, int& b, int& swaps) const {
    if (less(b, a)) {
      swaps++;
      int t = a; a = b; b = t;
    }
  }
  B2_HD int median(int a, int b, int c, int& swaps) const {
    order2(a, b, swaps);
    order2(b, c, swaps);
    order2(a, b, swaps);
    return b;
  }
  // hint: 0 unknown, 1 increasing, 2 decreasing
  B2_HD int choose_pivot(int a, int b, int* hint) const {
    int l = b - a, swaps = 0;
    int i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
#if defined(__CUDA_ARCH__)
    if (coop && l >= 8) {
      // the same medians on (index, key) pairs in registers: the nine sample keys are read by nine lanes at once
      const int lane = threadIdx.x & 31;
      int32_t mine = 0;
      if (lane < 9) mine = key[a + l / 4 * (lane / 3 + 1) + (lane % 3) - 1];
      int ix[9];
      int32_t kx[9];
#pragma unroll
      for (int m = 0; m < 9; m++) {
        kx[m] = __shfl_sync(0xffffffffu, mine, m);
        ix[m] = a + l / 4 * (m / 3 + 1) + (m % 3) - 1;
      }
      auto order2r = [&](int& ia, int32_t& ka, int& ib, int32_t& kb) {
        if (kb < ka) {
          swaps++;
          const int ti = ia; ia = ib; ib = ti;
          const int32_t tk = ka; ka = kb; kb = tk;
        }
      };
      auto medianr = [&](int& ia, int32_t& ka, int& ib, int32_t& kb, int& ic, int32_t& kc) {  // result in (ib, kb)
        order2r(ia, ka, ib, kb);
        order2r(ib, kb, ic, kc);
        order2r(ia, ka, ib, kb);
      };
      if (l >= 50) {
        medianr(ix[0], kx[0], ix[1], kx[1], ix[2], kx[2]);
        medianr(ix[3], kx[3], ix[4], kx[4], ix[5], kx[5]);
        medianr(ix[6], kx[6], ix[7], kx[7], ix[8], kx[8]);
      }
      medianr(ix[1], kx[1], ix[4], kx[4], ix[7], kx[7]);
      *hint = swaps == 0 ? 1 : swaps == 12 ? 2 : 0;
      return ix[4];
    }
#endif
    if (l >= 8) {
      if (l >= 50) {
        i = median(i - 1, i, i + 1, swaps);
        j = median(j - 1, j, j + 1, swaps);
        k = median(k - 1, k, k + 1, swaps);
      }
      j = median(i, j, k, swaps);
    }
    *hint = swaps == 0 ? 1 : swaps == 12 ? 2 : 0;
    return j;
  }
  B2_HD void reverse_range(int a, int b) const {
    int i = a, j = b - 1;
    while (i < j) { swap(i, j); i++; j--; }
  }
  B2_HD bool partial_insertion(int a, int b) const {
    int i = a + 1;
    for (int s = 0; s < 5; s++) {
      i = fwd_while_nondecreasing(i, b);
      if (i == b) return true;
      if (b - a < 50) return false;
      swap(i, i - 1);
#if defined(__CUDA_ARCH__)
      if (coop) {  // the two shifts below as block moves: a run of adjacent swaps carries one element across its run
        if (i - a >= 2) shift_left_coop(i - 1);
        if (b - i >= 2) shift_right_coop(i, b);
        continue;
      }
#endif
      if (i - a >= 2)
        for (int j = i - 1; j >= 1; j--) {
          if (!less(j, j - 1)) break;
          swap(j, j - 1);
        }
      if (b - i >= 2)
        for (int j = i + 1; j < b; j++) {
          if (!less(j, j - 1)) break;
          swap(j, j - 1);
        }
    }
    return false;
  }
#if defined(__CUDA_ARCH__)
  // for (j = i + 1; j < b && key[j] < key[j-1]; j++) swap(j, j-1): the element at i passes every following element
  // smaller than itself
  __device__ void shift_right_coop(int i, int b) const {
    const int lane = threadIdx.x & 31;
    const int32_t kx = key[i], x = ord[i];
    int stop = b;
    for (int j0 = i + 1; j0 < b; j0 += 32) {
      const int j = j0 + lane;
      const unsigned sb = __ballot_sync(0xffffffffu, j >= b || !(key[j] < kx));
      if (sb) { stop = j0 + __ffs((int)sb) - 1; break; }
    }
    if (stop > b) stop = b;
    const int dst = stop - 1;  // where the element ends up
    if (dst == i) return;
    __syncwarp();
    for (int c0 = i + 1; c0 <= dst; c0 += 32) {
      const int j = c0 + lane;
      int32_t v = 0, kv = 0;
      if (j <= dst) { v = ord[j]; kv = key[j]; }
      __syncwarp();
      if (j <= dst) { ord[j - 1] = v; key[j - 1] = kv; }
      __syncwarp();
    }
    if (lane == 0) { ord[dst] = x; key[dst] = kx; }
    __syncwarp();
  }
  // for (j = i; j >= 1 && key[j] < key[j-1]; j--) swap(j, j-1): the element at i passes every preceding element
  // greater than itself (down to index 1: the routine's own bound, not the range's)
  __device__ void shift_left_coop(int i) const {
    const int lane = threadIdx.x & 31;
    const int32_t ky = key[i], y = ord[i];
    int dst = 0;  // where the element ends up
    for (int j0 = i; j0 >= 1; j0 -= 32) {
      const int j = j0 - lane;  // candidate: the element stops at j when j < 1 or !(ky < key[j-1])
      const unsigned sb = __ballot_sync(0xffffffffu, j < 1 || !(ky < key[j - 1]));
      if (sb) { dst = j0 - (__ffs((int)sb) - 1); break; }
    }
    if (dst < 0) dst = 0;
    if (dst == i) return;
    __syncwarp();
    for (int c0 = i - 1; c0 >= dst; c0 -= 32) {
      const int j = c0 - lane;
      int32_t v = 0, kv = 0;
      if (j >= dst) { v = ord[j]; kv = key[j]; }
      __syncwarp();
      if (j >= dst) { ord[j + 1] = v; key[j + 1] = kv; }
      __syncwarp();
    }
    if (lane == 0) { ord[dst] = y; key[dst] = ky; }
    __syncwarp();
  }
#endif
  B2_HD void break_patterns(int a, int b) const {
    int length = b - a;
    if (length >= 8) {
      uint64_t random = (uint64_t)length;
      int bl = 0;
      for (unsigned x = (unsigned)length; x; x >>= 1) bl++;
      unsigned modulus = 1u << bl;
      int idx = a + (length / 4) * 2 - 1;
      for (int i = 0; i < 3; i++) {
        random ^= random << 13;
        random ^= random >> 7;
        random ^= random << 17;
        int other = (int)((unsigned)random & (modulus - 1));
        if (other >= length) other -= length;
        swap(idx - 1 + i, a + other);
      }
    }
  }
  B2_HD int partition_equal(int a, int b, int pivot) const {
    swap(a, pivot);
    int i = a + 1, j = b - 1;
    for (;;) {
      const int32_t kv = key[a];
      i = fwd_while(i, j, kv, true);   // while (i <= j && !less(a, i)) i++
      j = bwd_while(i, j, kv, true);   // while (i <= j && less(a, j)) j--
      if (i > j) break;
      swap(i, j);
      i++; j--;
    }
    return i;
  }
  B2_HD int partition(int a, int b, int pivot, bool* already) const {
    swap(a, pivot);
    int i = a + 1, j = b - 1;
    const int32_t kv = key[a];  // data[a] holds the pivot until the final swap
    i = fwd_while(i, j, kv, false);  // while (i <= j && less(i, a)) i++
    j = bwd_while(i, j, kv, false);  // while (i <= j && !less(j, a)) j--
    if (i > j) {
      swap(j, a);
      *already = true;
      return j;
    }
    swap(i, j);
    i++; j--;
    for (;;) {
      i = fwd_while(i, j, kv, false);
      j = bwd_while(i, j, kv, false);
      if (i > j) break;
      swap(i, j);
      i++; j--;
    }
    swap(j, a);
    *already = false;
    return j;
  }

  // pdqsort_func(data, 0, n, bits.Len(n)).  The Go routine recurses into the smaller side and then
  // keeps looping on the larger one; here the larger side's loop state is parked on an explicit stack
  // while the smaller side runs, which is the same order of operations with depth <= log2(n).
  B2_HD void sort(int n) const {
    struct Frame { int a, b, limit; bool wasBalanced, wasPartitioned; };
    Frame stack[36];
    int sp = 0;
    int limit0 = 0;
    for (unsigned x = (unsigned)n; x; x >>= 1) limit0++;
    stack[sp++] = Frame{0, n, limit0, true, true};
    while (sp > 0) {
      Frame fr = stack[--sp];
      int a = fr.a, b = fr.b, limit = fr.limit;
      bool wasBalanced = fr.wasBalanced, wasPartitioned = fr.wasPartitioned;
      for (;;) {
        int length = b - a;
        if (length <= 12) { insertion(a, b); break; }
        if (limit == 0) { heap_sort(a, b); break; }
        if (!wasBalanced) { break_patterns(a, b); limit--; }
        int hint;
        int pivot = choose_pivot(a, b, &hint);
        if (hint == 2) {
          reverse_range(a, b);
          pivot = (b - 1) - (pivot - a);
          hint = 1;
        }
        if (wasBalanced && wasPartitioned && hint == 1) {
          if (partial_insertion(a, b)) break;
        }
        if (a > 0 && !less(a - 1, pivot)) {
          a = partition_equal(a, b, pivot);
          continue;
        }
        bool already;
        int mid = partition(a, b, pivot, &already);
        wasPartitioned = already;
        int leftLen = mid - a, rightLen = b - mid;
        int balanceThreshold = length / 8;
        if (leftLen < rightLen) {
          wasBalanced = leftLen >= balanceThreshold;
          stack[sp++] = Frame{mid + 1, b, limit, wasBalanced, wasPartitioned};  // resume here afterwards
          b = mid;  // pdqsort_func(data, a, mid, limit)
        } else {
          wasBalanced = rightLen >= balanceThreshold;
          stack[sp++] = Frame{a, mid, limit, wasBalanced, wasPartitioned};
          a = mid + 1;  // pdqsort_func(data, mid+1, b, limit)
        }
        wasBalanced = true;
        wasPartitioned = true;
      }
    }
  }
};

}  // namespace b200

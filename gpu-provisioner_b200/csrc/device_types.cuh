// device_types.cuh — device-side views of the SoA encoding and the lane-group evaluation that both
// the dense stage-A kernel and the pack kernel call.
#pragma once
#include <cuda_runtime.h>

#include "b200_core.cuh"

namespace b200 {

// Dictionary-encoded pods (include/b200sched.h b200s_pods_soa_t): two class ids per pod, class tables
// for requirement signatures (key-major over NS) and request vectors (resource-major over NQ).
struct PodsDev {
  int32_t P, NS, NQ;
  int32_t res_cpu, res_mem;
  const int32_t* sig_id;    // [P]
  const int32_t* req_id;    // [P]
  const int64_t* ts;        // [P]
  const uint64_t* uid_hi;
  const uint64_t* uid_lo;
  const uint64_t* sig_mask;  // [K*NS]
  const uint64_t* sig_flags; // [NS]
  const int32_t* sig_gt;     // [NB*NS]
  const int32_t* sig_lt;     // [NB*NS]
  const uint64_t* sig_tol;   // [NS]
  const int64_t* rq_req;     // [R*NQ]
};

// NodeClaimTemplates, weight order; lives in global memory, copied to shared memory by each CTA.
struct TmplDev {
  int32_t T;
  int32_t _pad;
  uint64_t taints[MAXT];
  uint64_t rflags[MAXT];
  uint64_t rmask[MAXT][MAXK];
  int64_t gt[MAXT][MAXB];
  int64_t lt[MAXT][MAXB];
  int64_t daemon[MAXT][MAXR];
  int64_t limit[MAXT][MAXR];   // s.remainingResources of the template's NodePool at the start of the Solve
  uint32_t limit_mask[MAXT];   // bit r: the NodePool limits resource r
  uint32_t any_limit, _pad2;
};

// A requirement set + request vector + taints somewhere in (generic) memory, with element strides.
struct StateRef {
  const uint64_t* mask;  // element k at mask[k*ms]
  size_t ms;
  uint64_t flags;
  const int64_t* gt;     // element b at gt[b*bs]
  const int64_t* lt;
  size_t bs;
  const int64_t* req;    // element r at req[r*rs]
  size_t rs;
  uint64_t taints;
};

struct PodRef {
  const uint64_t* mask;
  size_t ms;
  uint64_t flags;
  const int32_t* gt;
  const int32_t* lt;
  size_t bs;
  const int64_t* req;
  size_t rs;
  uint64_t tol;
};

constexpr int ST_OK = 0, ST_TAINT = 1, ST_INCOMPATIBLE = 2, ST_FILTER = 3;

__device__ __forceinline__ EReq state_req(const KeyTab& kt, const StateRef& s, int k) {
  int b = kt.bslot[k];
  return make_ereq(s.mask[(size_t)k * s.ms], s.flags, k, b >= 0 ? s.gt[(size_t)b * s.bs] : 0, b >= 0 ? s.lt[(size_t)b * s.bs] : 0);
}

__device__ __forceinline__ EReq pod_req(const KeyTab& kt, const PodRef& p, int k) {
  int b = kt.bslot[k];
  return make_ereq(p.mask[(size_t)k * p.ms], p.flags, k, b >= 0 ? (int64_t)p.gt[(size_t)b * p.bs] : 0,
                   b >= 0 ? (int64_t)p.lt[(size_t)b * p.bs] : 0);
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
  }
}

// NodeClaim.CanAdd for one (state, pod) pair, cooperatively by a group of G lanes
// (K/controllers/provisioning/scheduling/nodeclaim.go:114-163 with filterInstanceTypesByRequirements
// :373-441 inlined).  Lane gl owns SKU words gl, gl+G, ...; cand[] is the state's InstanceTypeOptions.
// Returns ST_*; on ST_OK out[] holds the surviving instance types, on ST_FILTER *code the six flags.
// *n_visited = instance types the filter loop visited (0 when it never ran).
template <int G, int WPL>
__device__ __forceinline__ int group_can_add(const KeyTab& kt, EvalScratch* sc, const StateRef& st, const PodRef& pd,
                                             bool check_taints, const uint64_t (&cand)[WPL], uint64_t (&out)[WPL],
                                             uint8_t* code, int* n_visited) {
  const unsigned gm = group_mask<G>();
  const int gl = threadIdx.x & (G - 1);
  const int K = kt.n_keys, R = kt.n_res, W = kt.n_words;
  *n_visited = 0;
  // taints (nodeclaim.go:116; taints.go:53-65 precomputed into the tolerated bitset)
  if (check_taints && (st.taints & ~pd.tol)) return ST_TAINT;
  // requirement compatibility and merge, one key per lane (nodeclaim.go:125-131)
  bool ok = true;
  uint32_t my_def = 0, my_forg = 0;
  for (int k = gl; k < K; k += G) {
    EReq s = state_req(kt, st, k);
    EReq p = pod_req(kt, pd, k);
    ok = ok && compatible_key(kt, k, s, p, true);
    EReq m = add_req(kt, k, s, p);
    uint64_t a = 0;
    if (m.f & RF_DEF) {
      a = allowed_values(kt, k, m);
      my_def |= 1u << k;
      if (forgiving(m)) my_forg |= 1u << k;
    }
    sc->allowed[k] = a;
  }
  // total requests and their table rows, one resource per lane (nodeclaim.go:144)
  for (int r = gl; r < R; r += G) {
    int64_t x = st.req[(size_t)r * st.rs] + pd.req[(size_t)r * pd.rs];
    int nd = kt.res_nd[r];
    int i = fit_rank(kt.res_dv + kt.res_dv0[r], nd, x);
    sc->fit_row[r] = i < nd ? kt.res_row0[r] + i : -1;
  }
  unsigned bad = __ballot_sync(gm, !ok) & gm;
  uint32_t def = __reduce_or_sync(gm, my_def);
  uint32_t forg = __reduce_or_sync(gm, my_forg);
  if (bad) return ST_INCOMPATIBLE;
  __syncwarp(gm);
  if (gl == 0) {
    sc->defined = def;
    sc->forgive = forg;
    sc->zone_allowed = (kt.key_zone >= 0 && ((def >> kt.key_zone) & 1u)) ? sc->allowed[kt.key_zone] : ~0ull;
    sc->ct_allowed = (kt.key_ct >= 0 && ((def >> kt.key_ct) & 1u)) ? sc->allowed[kt.key_ct] : ~0ull;
  }
  __syncwarp(gm);
  // the filter loop, 64 instance types per word
  uint32_t bits = 0;
  int visited = 0;
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    int w = gl + i * G;
    out[i] = 0;
    if (w < W) {
      uint64_t c = cand[i];
      visited += popc64(c);
      WordEval e = eval_word(kt, *sc, w);
      uint64_t keep = c & e.compat & e.fits & e.off;
      out[i] = keep;
      bits |= (keep ? 1u : 0u) | ((c & e.compat) ? 2u : 0u) | ((c & e.fits) ? 4u : 0u) | ((c & e.off) ? 8u : 0u) |
              ((c & e.compat & e.fits & ~e.off) ? 16u : 0u) | ((c & e.compat & e.off & ~e.fits) ? 32u : 0u) |
              ((c & e.fits & e.off & ~e.compat) ? 64u : 0u);
    }
  }
  bits = __reduce_or_sync(gm, bits);
  *n_visited = (int)__reduce_add_sync(gm, (unsigned)visited);
  __syncwarp(gm);  // scratch may be rewritten by the next call
  if (bits & 1u) return ST_OK;
  *code = filter_flags(bits & 2u, bits & 4u, bits & 8u, bits & 16u, bits & 32u, bits & 64u);
  return ST_FILTER;
}

// Requirements of a new or updated NodeClaim: per-key Add of the pod's requirements
// (nodeclaim.go:131,173), written by the group's lanes to dst.
template <int G>
__device__ __forceinline__ void group_store_merged(const KeyTab& kt, const StateRef& st, const PodRef& pd, uint64_t* dmask,
                                                   size_t dms, uint64_t* dflags, int64_t* dgt, int64_t* dlt, size_t dbs) {
  const unsigned gm = group_mask<G>();
  const int gl = threadIdx.x & (G - 1);
  uint32_t flo = 0, fhi = 0;
  for (int k = gl; k < kt.n_keys; k += G) {
    EReq m = add_req(kt, k, state_req(kt, st, k), pod_req(kt, pd, k));
    dmask[(size_t)k * dms] = m.mask;
    uint32_t nib = m.f & 0xFu;
    if (k < 8) flo |= nib << (4 * k); else fhi |= nib << (4 * (k - 8));
    int b = kt.bslot[k];
    if (b >= 0) {
      dgt[(size_t)b * dbs] = (m.f & RF_GT) ? m.gt : 0;
      dlt[(size_t)b * dbs] = (m.f & RF_LT) ? m.lt : 0;
    }
  }
  flo = __reduce_or_sync(gm, flo);
  fhi = __reduce_or_sync(gm, fhi);
  if (gl == 0) *dflags = (uint64_t)flo | ((uint64_t)fhi << 32);
}

// Requirement part of CanAdd for a (template, signature) pair, without resources: taints, Compatible, and the
// two SKU sets the filter loop derives from the merged requirements — C = cand & compatible(merged),
// O = cand & hasOffering(merged).  The resource side (fits) depends only on the request vector and is
// tabulated separately, so the filter result of a fresh NodeClaim is C & O & F (nodeclaim.go:392-423).
template <int G, int WPL>
__device__ __forceinline__ int group_label_masks(const KeyTab& kt, EvalScratch* sc, const StateRef& st, const PodRef& pd,
                                                 const uint64_t (&cand)[WPL], uint64_t (&C)[WPL], uint64_t (&O)[WPL]) {
  const unsigned gm = group_mask<G>();
  const int gl = threadIdx.x & (G - 1);
  const int K = kt.n_keys, W = kt.n_words;
  if (st.taints & ~pd.tol) return ST_TAINT;
  bool ok = true;
  uint32_t my_def = 0, my_forg = 0;
  for (int k = gl; k < K; k += G) {
    EReq s = state_req(kt, st, k);
    EReq p = pod_req(kt, pd, k);
    ok = ok && compatible_key(kt, k, s, p, true);
    EReq m = add_req(kt, k, s, p);
    uint64_t a = 0;
    if (m.f & RF_DEF) {
      a = allowed_values(kt, k, m);
      my_def |= 1u << k;
      if (forgiving(m)) my_forg |= 1u << k;
    }
    sc->allowed[k] = a;
  }
  unsigned bad = __ballot_sync(gm, !ok) & gm;
  uint32_t def = __reduce_or_sync(gm, my_def);
  uint32_t forg = __reduce_or_sync(gm, my_forg);
  if (bad) return ST_INCOMPATIBLE;
  __syncwarp(gm);
  if (gl == 0) {
    sc->defined = def;
    sc->forgive = forg;
  }
  __syncwarp(gm);
  const uint64_t za = (kt.key_zone >= 0 && ((def >> kt.key_zone) & 1u)) ? sc->allowed[kt.key_zone] : ~0ull;
  const uint64_t ca = (kt.key_ct >= 0 && ((def >> kt.key_ct) & 1u)) ? sc->allowed[kt.key_ct] : ~0ull;
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    int w = gl + i * G;
    C[i] = 0;
    O[i] = 0;
    if (w < W) {
      C[i] = cand[i] & eval_compat_word(kt, *sc, w);
      O[i] = cand[i] & eval_off_word(kt, za, ca, w);
    }
  }
  __syncwarp(gm);
  return ST_OK;
}

// NodeClaim.CanAdd against an IN-FLIGHT claim (scheduler.go:552-584 -> nodeclaim.go:114-163), specialised
// on the keys the pod defines.  Invariant of a claim's InstanceTypeOptions (by induction over Add): every
// remaining SKU already satisfies compatible() for the claim's requirement on each key, has an offering the
// claim's zone / capacity-type requirements admit, and fits the claim's requests.  Adding a pod changes the
// merged requirement only on keys the pod defines, so only those keys' SKU rows (and the offering rows when
// the pod touches zone or capacity type) have to be re-evaluated, plus the resource rows for the new total.
// The diagnostic flags are not needed here: addToInflightNode discards the error.  All lanes run the scalar
// key algebra redundantly (it is uniform across the group), which avoids any cross-lane traffic.
template <int G, int WPL>
__device__ __forceinline__ bool group_inflight(const KeyTab& kt, const StateRef& st, const int64_t* hi, const PodRef& pd,
                                               uint32_t podkeys, const uint64_t (&cand)[WPL], uint64_t (&out)[WPL],
                                               int* n_visited) {
  const unsigned gm = group_mask<G>();
  const int gl = threadIdx.x & (G - 1);
  const int R = kt.n_res, W = kt.n_words;
  *n_visited = 0;
  if (st.taints & ~pd.tol) return false;  // nodeclaim.go:116
  // resource rows of the new total (nodeclaim.go:144), one resource per lane, then broadcast.  A total above
  // the largest Allocatable() the claim's instance types ever had cannot fit (cheap exact reject).
  int my_row = 0;
  bool over = false;
  if (gl < R) {
    int64_t x = st.req[(size_t)gl * st.rs] + pd.req[(size_t)gl * pd.rs];
    over = x > hi[gl];
    if (!over) {
      int nd = kt.res_nd[gl];
      int i = fit_rank(kt.res_dv + kt.res_dv0[gl], nd, x);
      my_row = i < nd ? kt.res_row0[gl] + i : -1;
      over = my_row < 0;
    }
  }
  if (__ballot_sync(gm, over) & gm) return false;
  uint64_t keep[WPL];
  int visited = 0;
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    keep[i] = cand[i];
    visited += popc64(cand[i]);
  }
  for (int r = 0; r < R; r++) {
    int row = __shfl_sync(gm, my_row, r, G);
#pragma unroll
    for (int i = 0; i < WPL; i++) {
      int w = gl + i * G;
      if (w < W) keep[i] &= ldg(kt.rows + (size_t)row * W + w);
    }
  }
  // keys the pod defines (nodeclaim.go:125-131): Compatible, merge, and the merged key's SKU rows
  bool ok = true, touch_off = false;
  uint32_t keys = podkeys;
  while (keys) {
    int k = ctz32(keys);
    keys &= keys - 1;
    EReq s = state_req(kt, st, k);
    EReq p = pod_req(kt, pd, k);
    ok = ok && compatible_key(kt, k, s, p, true);
    EReq m = add_req(kt, k, s, p);
    uint64_t a = allowed_values(kt, k, m);
    bool fg = forgiving(m);
#pragma unroll
    for (int i = 0; i < WPL; i++) {
      int w = gl + i * G;
      if (w < W) keep[i] &= eval_key_word(kt, k, a, fg, w);
    }
    touch_off = touch_off || k == kt.key_zone || k == kt.key_ct;
  }
  if (!ok) return false;  // incompatible requirements (uniform across the group)
  if (touch_off) {
    uint64_t za = ~0ull, ca = ~0ull;
    if (kt.key_zone >= 0) {
      EReq m = add_req(kt, kt.key_zone, state_req(kt, st, kt.key_zone), pod_req(kt, pd, kt.key_zone));
      if (m.f & RF_DEF) za = allowed_values(kt, kt.key_zone, m);
    }
    if (kt.key_ct >= 0) {
      EReq m = add_req(kt, kt.key_ct, state_req(kt, st, kt.key_ct), pod_req(kt, pd, kt.key_ct));
      if (m.f & RF_DEF) ca = allowed_values(kt, kt.key_ct, m);
    }
#pragma unroll
    for (int i = 0; i < WPL; i++) {
      int w = gl + i * G;
      if (w < W) keep[i] &= eval_off_word(kt, za, ca, w);
    }
  }
  unsigned any = 0;
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    out[i] = keep[i];
    any |= keep[i] ? 1u : 0u;
  }
  any = __reduce_or_sync(gm, any);
  *n_visited = (int)__reduce_add_sync(gm, (unsigned)visited);
  return any != 0;
}

// group_inflight for a pod WITHOUT requirement keys (no nodeSelector / affinity), the common case: only the resource
// rows can change the claim's instance types.  Written so that all 32 lanes of the warp stay converged — no early
// exits, full-mask shuffles and votes only (sub-warp masks compile to shuffle loops) — with `live` predicating the
// groups that have no claim to look at.  Lane gl < R of a group owns resource gl, lane w owns SKU word w (+ i*G).
// Returns whether the pod fits; out[] is the claim's new InstanceTypeOptions when it does.
// Per-lane constants of the resource a lane owns (lane gl < R owns resource gl for the whole kernel): hoisted
// once, because indexing the constant bank with a per-lane index serialises into one access per distinct lane.
struct LaneRes {
  const int64_t* dv;
  const uint16_t* bkt;
  int32_t nd, p2, row0;
};

__device__ __forceinline__ LaneRes lane_res(const KeyTab& kt, int r) {
  LaneRes l{nullptr, nullptr, 0, 0, 0};
  if (r < kt.n_res) {
    l.dv = kt.res_dv + kt.res_dv0[r];
    l.bkt = kt.res_bkt + (size_t)r * (RANK_BKTS + 1);
    l.nd = kt.res_nd[r];
    l.p2 = kt.res_p2[r];
    l.row0 = kt.res_row0[r];
  }
  return l;
}

template <int G, int WPL, bool PLAIN = false>
__device__ __forceinline__ bool group_inflight_nokeys(const KeyTab& kt, const uint64_t* rows, const LaneRes& lr, bool live,
                                                      const int64_t* creq, const int64_t* chi, const int64_t* cthr,
                                                      const uint64_t* cmask, bool tolerated, const int64_t* preq,
                                                      uint64_t (&out)[WPL], bool& types_changed, int64_t& new_thr) {
  // Invariant: the claim's types already lie inside row "alloc[r] >= thr[r]" of every resource r, thr[r] being the
  // smallest distinct allocatable value >= requests[r].  While requests[r] + pod[r] <= thr[r] that row is still the one
  // resources.Fits selects, so the claim keeps every type and the pod fits (nodeclaim.go:144-146 with an unchanged
  // filter result).  Only a total that moves past thr[r] needs the rank search and the row; one beyond the largest
  // allocatable value among the claim's types (hi[r]) cannot fit at all.
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned GM = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
  const unsigned lane = threadIdx.x & 31u, gl = threadIdx.x & (G - 1), gbase = lane & ~(unsigned)(G - 1);
  const int R = kt.n_res, W = kt.n_words;
  int64_t x = 0;
  bool dead = !tolerated, moved = false;
  if (live && gl < (unsigned)R) {
    x = creq[gl] + preq[gl];
    dead = dead || x > chi[gl];
    moved = x > cthr[gl];
  }
  const unsigned deadb = __ballot_sync(FULL, dead), movedb = __ballot_sync(FULL, moved);
  bool gdead = ((deadb >> gbase) & GM) != 0;
  const unsigned gmoved = gdead ? 0u : ((movedb >> gbase) & GM);  // resources whose row changes, for my group's claim
  types_changed = gmoved != 0;
  new_thr = 0;
  if (!__any_sync(FULL, live && gmoved != 0)) return live && !gdead;  // the common case: nothing to look up
  // some claim of this warp moves to a new row: rank search by the lanes that own a moved resource
  int row = -1;
  bool dead2 = false;
  if (moved && !gdead) {
    const int lo = fit_rank_bkt<PLAIN>(lr.dv, lr.bkt, x);
    dead2 = lo >= lr.nd;
    if (!dead2) {
      row = lr.row0 + lo;
      new_thr = PLAIN ? lr.dv[lo] : ldg(lr.dv + lo);
    }
  }
  const unsigned dead2b = __ballot_sync(FULL, dead2);
  gdead = gdead || ((dead2b >> gbase) & GM) != 0;
  uint64_t keep[WPL];
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    const int w = (int)gl + i * G;
    keep[i] = (live && !gdead && w < W) ? cmask[w] : 0ull;
  }
  unsigned any_moved = movedb;  // union over the warp's groups, folded onto the low G bits: a uniform loop bound
#pragma unroll
  for (int o = 16; o >= G && G < 32; o >>= 1) any_moved |= any_moved >> o;
  any_moved &= GM;
  for (unsigned m = any_moved; m; m &= m - 1) {
    const int r = __ffs((int)m) - 1;
    const int rr = __shfl_sync(FULL, row, (int)gbase + r);  // -1: this group's claim keeps resource r's row
#pragma unroll
    for (int i = 0; i < WPL; i++) {
      const int w = (int)gl + i * G;
      if (rr >= 0 && w < W) keep[i] &= PLAIN ? rows[(size_t)rr * W + w] : ldg(rows + (size_t)rr * W + w);
    }
  }
  bool nz = false;
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    out[i] = keep[i];
    nz = nz || keep[i] != 0;
  }
  const unsigned nzb = __ballot_sync(FULL, nz);
  return live && !gdead && ((nzb >> gbase) & GM) != 0;
}

// The same in-flight CanAdd as group_inflight, decided by ONE thread (used by the scan, 512 claims per pass): only
// the yes/no answer is needed there, so the SKU words are walked until the first one that keeps an instance type.
constexpr int MAX_FAST_KEYS = 4;  // pod-defined keys handled from registers; more than that re-derives per word

// Returns 0 when CanAdd fails before the filter loop (taints, Compatible), 1 when the filter loop leaves no
// instance type, 2 when the pod fits.  `ract` has bit r set for the resources the pod actually requests: a
// resource with a zero request leaves the claim's total unchanged, and the claim's instance types already fit
// that total, so its row is skipped (exact by the invariant above).  Written branch-free over resources and
// words so that the independent table lookups overlap.
__device__ __forceinline__ int thread_inflight(const KeyTab& kt, const StateRef& st, const int64_t* hi, const int64_t* thr,
                                               const uint64_t* cmask, const PodRef& pd, uint32_t podkeys, uint32_t ract) {
  const int W = kt.n_words;
  if (st.taints & ~pd.tol) return 0;
  // New totals first: a claim whose largest allocatable value cannot hold the pod is rejected for a few loads -- on a
  // long scan most claims are full -- and a total that stays at or below the claim's threshold (device_types.cuh,
  // group_inflight_nokeys) keeps the resource's row, so only the resources in `mv` need a rank and a row at all.
  int64_t x[MAXR];
  uint32_t mv = 0;
  bool dead = false;
#pragma unroll
  for (int r = 0; r < MAXR; r++) {
    x[r] = 0;
    if ((ract >> r) & 1u) {
      x[r] = st.req[(size_t)r * st.rs] + pd.req[(size_t)r * pd.rs];
      dead = dead || x[r] > hi[r];
      if (x[r] > thr[r]) mv |= 1u << r;
    }
  }
  if (dead) return 1;
  // keys the pod defines: Compatible + merged admitted values (nodeclaim.go:125-131)
  int kk[MAX_FAST_KEYS];
  uint64_t ka[MAX_FAST_KEYS];
  uint32_t kfg = 0;
  int nk = 0;
  bool touch_off = false;
  uint64_t za = ~0ull, ca = ~0ull;
  if (podkeys) {
    uint32_t keys = podkeys;
    while (keys) {
      int k = ctz32(keys);
      keys &= keys - 1;
      EReq s = state_req(kt, st, k);
      EReq p = pod_req(kt, pd, k);
      if (!compatible_key(kt, k, s, p, true)) return 0;
      EReq m = add_req(kt, k, s, p);
      uint64_t a = allowed_values(kt, k, m);
#pragma unroll
      for (int j = 0; j < MAX_FAST_KEYS; j++)
        if (j == nk) { kk[j] = k; ka[j] = a; }
      if (nk < MAX_FAST_KEYS && forgiving(m)) kfg |= 1u << nk;
      nk++;
      touch_off = touch_off || k == kt.key_zone || k == kt.key_ct;
    }
    if (touch_off) {
      if (kt.key_zone >= 0) {
        EReq m = add_req(kt, kt.key_zone, state_req(kt, st, kt.key_zone), pod_req(kt, pd, kt.key_zone));
        if (m.f & RF_DEF) za = allowed_values(kt, kt.key_zone, m);
      }
      if (kt.key_ct >= 0) {
        EReq m = add_req(kt, kt.key_ct, state_req(kt, st, kt.key_ct), pod_req(kt, pd, kt.key_ct));
        if (m.f & RF_DEF) ca = allowed_values(kt, kt.key_ct, m);
      }
    }
  }
  if (!podkeys && !mv) return 2;  // nothing can change the claim's (non-empty) instance types
  int rowoff[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; r++) {
    rowoff[r] = 0;
    if ((mv >> r) & 1u) {
      const int lo = fit_rank_bkt(kt.res_dv + kt.res_dv0[r], kt.res_bkt + (size_t)r * (RANK_BKTS + 1), x[r]);
      dead = dead || lo >= kt.res_nd[r];
      rowoff[r] = (kt.res_row0[r] + lo) * W;
    }
  }
  if (dead) return 1;
  bool any = false;
  if (!podkeys) {
#pragma unroll 4
    for (int w = 0; w < W; w++) {
      uint64_t m = cmask[w];
#pragma unroll
      for (int r = 0; r < MAXR; r++)
        if ((mv >> r) & 1u) m &= ldg(kt.rows + rowoff[r] + w);
      any = any || m != 0;
    }
    return any ? 2 : 1;
  }
  for (int w = 0; w < W && !any; w++) {
    uint64_t m = cmask[w];
    if (m == 0) continue;
#pragma unroll
    for (int r = 0; r < MAXR; r++)
      if ((mv >> r) & 1u) m &= ldg(kt.rows + rowoff[r] + w);
    if (m == 0) continue;
    if (nk <= MAX_FAST_KEYS) {
#pragma unroll
      for (int j = 0; j < MAX_FAST_KEYS; j++)
        if (j < nk) m &= eval_key_word(kt, kk[j], ka[j], (kfg >> j) & 1u, w);
    } else {
      uint32_t keys = podkeys;
      while (keys) {
        int k = ctz32(keys);
        keys &= keys - 1;
        EReq mm = add_req(kt, k, state_req(kt, st, k), pod_req(kt, pd, k));
        m &= eval_key_word(kt, k, allowed_values(kt, k, mm), forgiving(mm), w);
      }
    }
    if (m && touch_off) m &= eval_off_word(kt, za, ca, w);
    any = m != 0;
  }
  return any ? 2 : 1;
}

// NodeClaim.Add's requirement update (nodeclaim.go:173) for the keys the pod defines; the rest are unchanged.
template <int G>
__device__ __forceinline__ void group_store_merged_keys(const KeyTab& kt, const StateRef& st, const PodRef& pd,
                                                        uint32_t podkeys, uint64_t* dmask, uint64_t* dflags, int64_t* dgt,
                                                        int64_t* dlt) {
  const int gl = threadIdx.x & (G - 1);
  uint64_t flags = st.flags;
  uint32_t keys = podkeys;
  while (keys) {
    int k = ctz32(keys);
    keys &= keys - 1;
    EReq m = add_req(kt, k, state_req(kt, st, k), pod_req(kt, pd, k));
    flags = (flags & ~(0xFull << (4 * k))) | ((uint64_t)(m.f & 0xFu) << (4 * k));
    if (gl == 0) {
      dmask[k] = m.mask;
      int b = kt.bslot[k];
      if (b >= 0) {
        dgt[b] = (m.f & RF_GT) ? m.gt : 0;
        dlt[b] = (m.f & RF_LT) ? m.lt : 0;
      }
    }
  }
  if (gl == 0) *dflags = flags;
}

__device__ __forceinline__ uint64_t tail_mask(int S, int w) {
  int W = (S + 63) >> 6;
  if (w != W - 1 || (S & 63) == 0) return ~0ull;
  return (1ull << (S & 63)) - 1ull;
}

}  // namespace b200

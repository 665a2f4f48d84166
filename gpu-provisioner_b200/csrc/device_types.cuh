// device_types.cuh — device-side views of the SoA encoding and the lane-group evaluation that both
// the dense stage-A kernel and the pack kernel call.
#pragma once
#include <cuda_runtime.h>

#include "b200_core.cuh"

namespace b200 {

struct PodsDev {
  int32_t P;
  int32_t res_cpu, res_mem;
  const int64_t* req;      // [R*P]
  const uint64_t* rmask;   // [K*P]
  const uint64_t* rflags;  // [P]
  const int32_t* gt;       // [NB*P]
  const int32_t* lt;       // [NB*P]
  const uint64_t* tol;     // [P]
  const int64_t* ts;
  const uint64_t* uid_hi;
  const uint64_t* uid_lo;
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

__device__ __forceinline__ uint64_t tail_mask(int S, int w) {
  int W = (S + 63) >> 6;
  if (w != W - 1 || (S & 63) == 0) return ~0ull;
  return (1ull << (S & 63)) - 1ull;
}

}  // namespace b200

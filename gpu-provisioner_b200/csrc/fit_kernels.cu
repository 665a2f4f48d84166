// fit_kernels.cu — sm_100a kernels of the dense side of the path:
//   template_prefilter_kernel  NewScheduler's per-NodePool instance-type pre-filter (scheduler.go:143-158)
//   class_tables_kernel        the filter loop of a fresh NodeClaim (nodeclaim.go:114-163, :373-441) factored by
//                              what it depends on: per (requirement signature, template) the SKU sets
//                              C = compatible, O = hasOffering and the taint / Compatible status; per (request
//                              vector, template) the SKU set F = fits
//   fit_dense_kernel           addToNewNodeClaim's template walk for every pending pod (scheduler.go:587-675):
//                              per-pod feasibility mask C & O & F over SKUs, the first feasible template, the
//                              per-template attempt codes (six diagnostic flags on failure) — a pure stream
//   ffd_* kernels              NewQueue's order (queue.go:37-43, :72-108): bitonic sort of pod indices on
//                              (cpu desc, memory desc, creationTimestamp asc, UID asc) — a total order
//   select_cheapest_kernel     OrderByPrice(reqs)[0] (cloudprovider/types.go:221-240) per SKU set
//
// All integer / compare / bitset work: no tensor cores, no floating-point arithmetic (prices are compared
// as their IEEE-754 bit patterns, which is order-preserving for the non-negative finite prices the
// encoder admits).  Bit-row tables are staged in shared memory; pods stream through in SoA tiles.
#include <cstdlib>
#include <cstring>

#include "device_types.cuh"
#include <algorithm>

#include "kernels.h"

namespace b200 {

// -------------------------------------------------------------------------------------------------
// shared-memory staging helpers
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_words(void* dst, const void* src, size_t bytes) {
  // bytes is a multiple of 8; both 8-byte aligned
  uint64_t* d = reinterpret_cast<uint64_t*>(dst);
  const uint64_t* s = reinterpret_cast<const uint64_t*>(src);
  for (size_t i = threadIdx.x; i < bytes / 8; i += blockDim.x) d[i] = s[i];
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Carves the catalog tables (bit rows + distinct values) into shared memory and repoints kt at them.
__device__ __forceinline__ unsigned char* stage_tables(KeyTab& kt, unsigned char* smem, int tables_in_smem) {
  if (!tables_in_smem) return smem;
  size_t row_bytes = (size_t)kt.n_rows * kt.n_words * 8;
  size_t dv_bytes = (size_t)kt.n_dv * 8;
  copy_words(smem, kt.rows, row_bytes);
  copy_words(smem + row_bytes, kt.res_dv, dv_bytes);
  kt.rows = reinterpret_cast<const uint64_t*>(smem);
  kt.res_dv = reinterpret_cast<const int64_t*>(smem + row_bytes);
  return smem + row_bytes + dv_bytes;
}

// -------------------------------------------------------------------------------------------------
// template pre-filter: one lane group per template, "pod" = no requirements, zero requests
// -------------------------------------------------------------------------------------------------
template <int G, int WPL>
__global__ void __launch_bounds__(32 * MAXT) template_prefilter_kernel(KeyTab kt, const TmplDev* __restrict__ tg,
                                                                       const uint64_t* __restrict__ type_admit,
                                                                       uint64_t* __restrict__ tmask,
                                                                       int32_t* __restrict__ alive,
                                                                       int32_t* __restrict__ tm_count) {
  __shared__ EvalScratch sc[MAXT];
  __shared__ int64_t zero_req[MAXR];
  __shared__ uint64_t zero_mask[MAXK];
  __shared__ int32_t zero_b[MAXB];
  if (threadIdx.x < MAXR) zero_req[threadIdx.x] = 0;
  if (threadIdx.x < MAXK) zero_mask[threadIdx.x] = 0;
  if (threadIdx.x < MAXB) zero_b[threadIdx.x] = 0;
  __syncthreads();
  const int t = threadIdx.x >> 5;  // one warp per template
  const int gl = threadIdx.x & (G - 1);
  if (t >= tg->T || (threadIdx.x & 31) >= G) return;
  StateRef st{tg->rmask[t], 1, tg->rflags[t], tg->gt[t], tg->lt[t], 1, zero_req, 1, 0};
  PodRef pd{zero_mask, 1, 0, zero_b, zero_b, 1, zero_req, 1, 0};
  uint64_t cand[WPL], out[WPL];
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    int w = gl + i * G;
    cand[i] = 0;
    if (w < kt.n_words) cand[i] = tail_mask(kt.n_skus, w) & (type_admit ? type_admit[(size_t)t * kt.n_words + w] : ~0ull);
  }
  uint8_t code = 0;
  int visited;
  int stt = group_can_add<G, WPL>(kt, &sc[t], st, pd, false, cand, out, &code, &visited);
  unsigned cnt = 0;  // the template's instance types after the pre-filter: what one fresh-claim evaluation visits
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    int w = gl + i * G;
    if (w < kt.n_words) {
      uint64_t v = stt == ST_OK ? out[i] : 0ull;
      tmask[(size_t)t * kt.n_words + w] = v;
      cnt += (unsigned)popc64(v);
    }
  }
  cnt = __reduce_add_sync(group_mask<G>(), cnt);
  if (gl == 0) {
    alive[t] = cnt ? 1 : 0;
    tm_count[t] = (int32_t)cnt;
  }
}

// -------------------------------------------------------------------------------------------------
// class tables: one lane group per (signature, template) and per (request vector, template)
// -------------------------------------------------------------------------------------------------
constexpr int CLS_THREADS = 256;


template <int G, int WPL>
__global__ void __launch_bounds__(CLS_THREADS) class_tables_kernel(KeyTab kt, const TmplDev* __restrict__ tg,
                                                                   const uint64_t* __restrict__ tmask_g,
                                                                   const int32_t* __restrict__ alive_g, PodsDev pods,
                                                                   uint8_t* __restrict__ sg_status,
                                                                   uint64_t* __restrict__ sgC, uint64_t* __restrict__ sgO,
                                                                   uint64_t* __restrict__ sgCO,
                                                                   uint64_t* __restrict__ rqF, int tables_in_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NG = CLS_THREADS / G;
  TmplDev* tm = reinterpret_cast<TmplDev*>(smem);
  EvalScratch* scratch = reinterpret_cast<EvalScratch*>(smem + align16(sizeof(TmplDev)));
  unsigned char* tab = smem + align16(sizeof(TmplDev)) + align16(sizeof(EvalScratch) * NG);
  copy_words(tm, tg, sizeof(TmplDev));
  stage_tables(kt, tab, tables_in_smem);
  __syncthreads();
  const int T = tm->T, W = kt.n_words, R = kt.n_res;
  const int g = threadIdx.x / G, gl = threadIdx.x & (G - 1);
  const unsigned gm = group_mask<G>();
  const long long n_sig = (long long)pods.NS * T, n_items = n_sig + (long long)pods.NQ * T;
  for (long long item = (long long)blockIdx.x * NG + g; item < n_items; item += (long long)gridDim.x * NG) {
    if (item < n_sig) {
      const int sig = (int)(item / T), t = (int)(item % T);
      uint8_t status = B200S_FIT_DROPPED;
      uint64_t C[WPL], O[WPL];
#pragma unroll
      for (int i = 0; i < WPL; i++) { C[i] = 0; O[i] = 0; }
      if (alive_g[t]) {
        StateRef st{tm->rmask[t], 1, tm->rflags[t], tm->gt[t], tm->lt[t], 1, tm->daemon[t], 1, tm->taints[t]};
        PodRef pd{pods.sig_mask + sig, (size_t)pods.NS, pods.sig_flags[sig], pods.sig_gt + sig, pods.sig_lt + sig,
                  (size_t)pods.NS, nullptr, 0, pods.sig_tol[sig]};
        uint64_t cand[WPL];
#pragma unroll
        for (int i = 0; i < WPL; i++) {
          int w = gl + i * G;
          cand[i] = w < W ? tmask_g[(size_t)t * W + w] : 0ull;
        }
        int stt = group_label_masks<G, WPL>(kt, &scratch[g], st, pd, cand, C, O);
        status = stt == ST_OK ? B200S_FIT_NONE : stt == ST_TAINT ? B200S_FIT_TAINT : B200S_FIT_INCOMPATIBLE;
      }
      if (gl == 0) sg_status[item] = status;
#pragma unroll
      for (int i = 0; i < WPL; i++) {
        int w = gl + i * G;
        if (w < W) {
          sgC[(size_t)item * W + w] = C[i];
          sgO[(size_t)item * W + w] = O[i];
          sgCO[(size_t)item * W + w] = C[i] & O[i];
        }
      }
    } else {
      const long long j = item - n_sig;
      const int q = (int)(j / T), t = (int)(j % T);
      // F = template's instance types that fit daemon overhead + the request vector (nodeclaim.go:144,447-449)
      int my_row = 0;
      if (gl < R) {
        int64_t x = tm->daemon[t][gl] + pods.rq_req[(size_t)gl * pods.NQ + q];
        int nd = kt.res_nd[gl];
        int i = fit_rank(kt.res_dv + kt.res_dv0[gl], nd, x);
        my_row = i < nd ? kt.res_row0[gl] + i : -1;
      }
      uint64_t F[WPL];
#pragma unroll
      for (int i = 0; i < WPL; i++) {
        int w = gl + i * G;
        F[i] = w < W ? (tmask_g[(size_t)t * W + w] & kt.rows[(size_t)kt.row_nonneg * W + w]) : 0ull;
      }
      for (int r = 0; r < R; r++) {
        int row = __shfl_sync(gm, my_row, r, G);
#pragma unroll
        for (int i = 0; i < WPL; i++) {
          int w = gl + i * G;
          if (w < W) F[i] &= row < 0 ? 0ull : kt.rows[(size_t)row * W + w];
        }
      }
#pragma unroll
      for (int i = 0; i < WPL; i++) {
        int w = gl + i * G;
        if (w < W) rqF[(size_t)j * W + w] = F[i];
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------
// dense stage A: stream the pods, combine the class rows
// -------------------------------------------------------------------------------------------------
constexpr int FIT_THREADS = 256;

template <int G, int WPL>
__global__ void __launch_bounds__(FIT_THREADS) fit_dense_kernel(int P, int T, int W, const int32_t* __restrict__ sig_id,
                                                                const int32_t* __restrict__ req_id,
                                                                const uint8_t* __restrict__ sg_status,
                                                                const uint64_t* __restrict__ sgC,
                                                                const uint64_t* __restrict__ sgO,
                                                                const uint64_t* __restrict__ rqF,
                                                                const int32_t* __restrict__ tm_count,
                                                                int32_t* __restrict__ tmpl_out,
                                                                uint64_t* __restrict__ mask_out,
                                                                uint8_t* __restrict__ code_out,
                                                                unsigned long long* __restrict__ evals) {
  constexpr int NG = FIT_THREADS / G;
  const unsigned gm = group_mask<G>();
  const int gl = threadIdx.x & (G - 1);
  unsigned long long my_evals = 0;
  for (int p = blockIdx.x * NG + threadIdx.x / G; p < P; p += gridDim.x * NG) {
    const int sig = sig_id[p], rq = req_id[p];
    int chosen = -1;
    uint64_t keep[WPL];
#pragma unroll
    for (int i = 0; i < WPL; i++) keep[i] = 0;
    uint32_t code_lo = 0, code_hi = 0;  // byte t of the pair = attempt code of template t (T <= 16... 8 per word)
    uint32_t code_2 = 0, code_3 = 0;
    for (int t = 0; t < T; t++) {
      uint32_t code = sg_status[(size_t)sig * T + t];
      if (chosen >= 0 && code != B200S_FIT_DROPPED) code = B200S_FIT_NONE;  // templates after the first success are never tried
      if (code == B200S_FIT_NONE && chosen < 0) {  // addToNewNodeClaim stops at the first template that works
        const size_t cb = ((size_t)sig * T + t) * W, fb = ((size_t)rq * T + t) * W;
        uint64_t c[WPL], o[WPL], f[WPL];
        unsigned any = 0;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
          int w = gl + i * G;
          c[i] = 0; o[i] = 0; f[i] = 0;
          if (w < W) {
            c[i] = sgC[cb + w];
            o[i] = sgO[cb + w];
            f[i] = rqF[fb + w];
          }
          any |= (c[i] & o[i] & f[i]) ? 1u : 0u;
        }
        any = __reduce_or_sync(gm, any);
        if (gl == 0) my_evals += (unsigned long long)tm_count[t];
        if (any) {
          chosen = t;
#pragma unroll
          for (int i = 0; i < WPL; i++) keep[i] = c[i] & o[i] & f[i];
        } else {  // InstanceTypeFilterError: the six flags (nodeclaim.go:409-416)
          uint32_t bits = 0;
#pragma unroll
          for (int i = 0; i < WPL; i++)
            bits |= (c[i] ? 2u : 0u) | (f[i] ? 4u : 0u) | (o[i] ? 8u : 0u) | ((c[i] & f[i] & ~o[i]) ? 16u : 0u) |
                    ((c[i] & o[i] & ~f[i]) ? 32u : 0u) | ((f[i] & o[i] & ~c[i]) ? 64u : 0u);
          bits = __reduce_or_sync(gm, bits);
          code = filter_flags(bits & 2u, bits & 4u, bits & 8u, bits & 16u, bits & 32u, bits & 64u);
        }
      }
      const uint32_t sh = (uint32_t)(t & 3) * 8u;
      if (t < 4) code_lo |= code << sh;
      else if (t < 8) code_hi |= code << sh;
      else if (t < 12) code_2 |= code << sh;
      else code_3 |= code << sh;
    }
    if (gl == 0) tmpl_out[p] = chosen;
    if (gl < T) {
      uint32_t word = gl < 4 ? code_lo : gl < 8 ? code_hi : gl < 12 ? code_2 : code_3;
      code_out[(size_t)p * T + gl] = (uint8_t)(word >> ((gl & 3) * 8));
    }
    if (G < MAXT && gl + G < T) {
      int t = gl + G;
      uint32_t word = t < 4 ? code_lo : t < 8 ? code_hi : t < 12 ? code_2 : code_3;
      code_out[(size_t)p * T + t] = (uint8_t)(word >> ((t & 3) * 8));
    }
#pragma unroll
    for (int i = 0; i < WPL; i++) {
      int w = gl + i * G;
      if (w < W && mask_out) mask_out[(size_t)p * W + w] = keep[i];
    }
  }
  for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
  if ((threadIdx.x & 31) == 0 && my_evals) atomicAdd(evals, my_evals);
}

// Vector form of fit_dense_kernel for an even number of SKU words: a lane owns 16-byte pieces of the row
// (two SKU words), LPG lanes own one pod, 32/LPG pods per warp.  Lane j of a group also carries the attempt code of
// template j (in chunks of LPG templates), so the codes cost one byte load and one byte store per pod.
// Control flow is kept uniform across the WARP so that every vote is a single full-mask instruction from which a
// group extracts its own bits (sub-warp masks make the compiler emit shuffle loops).  Indices are 32-bit: the
// launcher checks that the tables and the output stay below 2^31 vectors.
struct DenseArgs {
  int P, T, W2;
  const int32_t* __restrict__ sig_id;
  const int32_t* __restrict__ req_id;
  const uint8_t* __restrict__ sg_status;
  const ulonglong2* __restrict__ sgC;
  const ulonglong2* __restrict__ sgO;
  const ulonglong2* __restrict__ sgCO;  // C & O
  const ulonglong2* __restrict__ rqF;
  const int32_t* __restrict__ tm_count;
  int32_t* __restrict__ tmpl_out;
  ulonglong2* __restrict__ mask_out;
  uint8_t* __restrict__ code_out;
  unsigned long long* __restrict__ evals;
};

// One warp iteration (32/LPG pods) of the template walk in full generality: any number of templates, later
// candidates after a failed one, the six diagnostic flags.  Must be entered by the whole warp.
template <int LPG, int VPL>
__device__ __forceinline__ void dense_general_iter(const DenseArgs& a, int p, bool live, unsigned sig, unsigned rq,
                                                   unsigned& my_evals, ulonglong2* srow = nullptr) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr unsigned GLOW = LPG == 32 ? 0xffffffffu : ((1u << LPG) - 1u);
  const unsigned lane = threadIdx.x & 31u, gl = threadIdx.x & (LPG - 1), gshift = lane & ~(unsigned)(LPG - 1);
  const int T = a.T, W2 = a.W2;
  int chosen = -1;
  ulonglong2 keep[VPL];
#pragma unroll
  for (int i = 0; i < VPL; i++) keep[i] = make_ulonglong2(0ull, 0ull);
  for (int t0 = 0; t0 < T; t0 += LPG) {
    const int t = t0 + (int)gl;
    uint32_t code = (live && t < T) ? a.sg_status[sig * (unsigned)T + (unsigned)t] : 0xFFu;
    const uint32_t status = code;
    unsigned ok = (__ballot_sync(FULL, status == B200S_FIT_NONE) >> gshift) & GLOW;
    for (;;) {  // addToNewNodeClaim walks the templates in order and stops at the first that works
      const bool active = chosen < 0 && ok != 0;
      if (!__any_sync(FULL, active)) break;
      const int j = active ? __ffs(ok) - 1 : 0;
      if (active) ok &= ok - 1;
      const unsigned tt = (unsigned)(t0 + j);
      const unsigned cb = (sig * (unsigned)T + tt) * (unsigned)W2, fb = (rq * (unsigned)T + tt) * (unsigned)W2;
      ulonglong2 co[VPL], f[VPL];
      bool nz = false;
#pragma unroll
      for (int i = 0; i < VPL; i++) {
        const unsigned v = gl + i * LPG;
        co[i] = f[i] = make_ulonglong2(0ull, 0ull);
        if (active && v < (unsigned)W2) {
          co[i] = __ldg(a.sgCO + cb + v);
          f[i] = __ldg(a.rqF + fb + v);
        }
        nz = nz || ((co[i].x & f[i].x) | (co[i].y & f[i].y)) != 0;
      }
      if (active && gl == 0) my_evals += (unsigned)a.tm_count[tt];
      const bool fits = ((__ballot_sync(FULL, nz) >> gshift) & GLOW) != 0;
      if (active && fits) {
        chosen = (int)tt;
#pragma unroll
        for (int i = 0; i < VPL; i++) keep[i] = make_ulonglong2(co[i].x & f[i].x, co[i].y & f[i].y);
      }
      if (__any_sync(FULL, active && !fits)) {  // InstanceTypeFilterError: the six flags (nodeclaim.go:409-416)
        bool bC = false, bF = false, bO = false, bCF = false, bCO = false, bFO = false;
#pragma unroll
        for (int i = 0; i < VPL; i++) {
          const unsigned v = gl + i * LPG;
          ulonglong2 c = make_ulonglong2(0ull, 0ull), o = c;
          if (active && !fits && v < (unsigned)W2) {  // the separate C and O rows are only read on this path
            c = __ldg(a.sgC + cb + v);
            o = __ldg(a.sgO + cb + v);
          }
          bC = bC || (c.x | c.y);
          bF = bF || (f[i].x | f[i].y);
          bO = bO || (o.x | o.y);
          bCF = bCF || ((c.x & f[i].x & ~o.x) | (c.y & f[i].y & ~o.y));
          bCO = bCO || ((c.x & o.x & ~f[i].x) | (c.y & o.y & ~f[i].y));
          bFO = bFO || ((f[i].x & o.x & ~c.x) | (f[i].y & o.y & ~c.y));
        }
        const unsigned vC = (__ballot_sync(FULL, bC) >> gshift) & GLOW, vF = (__ballot_sync(FULL, bF) >> gshift) & GLOW,
                       vO = (__ballot_sync(FULL, bO) >> gshift) & GLOW, vCF = (__ballot_sync(FULL, bCF) >> gshift) & GLOW,
                       vCO = (__ballot_sync(FULL, bCO) >> gshift) & GLOW, vFO = (__ballot_sync(FULL, bFO) >> gshift) & GLOW;
        if (active && !fits && (int)gl == j) code = filter_flags(vC, vF, vO, vCF, vCO, vFO);
      }
    }
    // templates after the first success are never tried; dropped ones stay marked
    if (chosen >= 0 && t > chosen && status != B200S_FIT_DROPPED) code = B200S_FIT_NONE;
    if (live && t < T) a.code_out[(unsigned)p * (unsigned)T + (unsigned)t] = (uint8_t)code;
  }
  if (live && gl == 0) a.tmpl_out[p] = chosen;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const unsigned v = gl + i * LPG;
    if (srow) {  // staged: the warp's rows leave through one bulk copy (fit_dense_bulk_kernel)
      if (v < (unsigned)W2) srow[v] = keep[i];
    } else if (live && v < (unsigned)W2 && a.mask_out) {
      a.mask_out[(unsigned)p * (unsigned)W2 + v] = keep[i];  // null: rows on demand
    }
  }
}

// any T: the general walk for every warp iteration
template <int LPG, int VPL>
__global__ void __launch_bounds__(FIT_THREADS) fit_dense_vec_kernel(const __grid_constant__ DenseArgs a) {
  constexpr int NG = FIT_THREADS / LPG;
  const unsigned lane = threadIdx.x & 31u;
  unsigned my_evals = 0;
  const int stride = gridDim.x * NG;
  // all lanes of a warp run the same number of iterations (idle groups are predicated off)
  const int warp_first = blockIdx.x * NG + threadIdx.x / LPG - (int)(lane / LPG);
  for (int pw = warp_first; pw < a.P; pw += stride) {
    const int p = pw + (int)(lane / LPG);
    const bool live = p < a.P;
    const unsigned sig = live ? (unsigned)a.sig_id[p] : 0u, rq = live ? (unsigned)a.req_id[p] : 0u;
    dense_general_iter<LPG, VPL>(a, p, live, sig, rq, my_evals);
  }
  for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
  if (lane == 0 && my_evals) atomicAdd(a.evals, (unsigned long long)my_evals);
}

// The same walk with the rows staged in shared memory: the 32/LPG pods of a warp iteration are consecutive, so their
// rows are one contiguous piece of the output, written by ONE bulk asynchronous copy (cp.async.bulk, the TMA unit)
// issued by lane 0 — the store stream no longer shares the load/store unit's queues with the class-row gathers.
// Two buffers per warp: a buffer is reused once the copy before the last has finished reading it.
template <int LPG, int VPL>
__global__ void __launch_bounds__(FIT_THREADS) fit_dense_bulk_kernel(const __grid_constant__ DenseArgs a) {
  constexpr int NG = FIT_THREADS / LPG, PPW = 32 / LPG;
  extern __shared__ __align__(128) unsigned char dense_smem[];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int W2 = a.W2;
  const unsigned buf_vecs = (unsigned)(PPW * W2);  // 16-byte vectors of one buffer
  ulonglong2* wbuf = reinterpret_cast<ulonglong2*>(dense_smem) + (size_t)warp * 2 * buf_vecs;
  unsigned my_evals = 0;
  const int stride = gridDim.x * NG;
  const int warp_first = blockIdx.x * NG + threadIdx.x / LPG - (int)(lane / LPG);
  int it = 0;
  for (int pw = warp_first; pw < a.P; pw += stride, it++) {
    const int p = pw + (int)(lane / LPG);
    const bool live = p < a.P;
    const unsigned sig = live ? (unsigned)a.sig_id[p] : 0u, rq = live ? (unsigned)a.req_id[p] : 0u;
    ulonglong2* buf = wbuf + (size_t)(it & 1) * buf_vecs;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the copy out of this buffer is done
    __syncwarp();
    dense_general_iter<LPG, VPL>(a, p, live, sig, rq, my_evals, buf + (size_t)(lane / LPG) * W2);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
      const int rows = min(PPW, a.P - pw);
      const unsigned bytes = (unsigned)rows * (unsigned)W2 * 16u;
      const unsigned long long dst = (unsigned long long)(a.mask_out + (size_t)pw * W2);
      const unsigned src = (unsigned)__cvta_generic_to_shared(buf);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
  if (lane == 0 && my_evals) atomicAdd(a.evals, (unsigned long long)my_evals);
}

// -------------------------------------------------------------------------------------------------
// FFD order: bitonic sort of pod indices by NewQueue's comparison (queue.go:72-108): cpu descending, memory descending,
// creation timestamp ascending, UID ascending; indices >= P are +infinity.  The leading part of the key is folded
// into one word per pod, k1 = (problem of a batched solve << 32) | rank of the pod's request class among the classes
// in (cpu desc, memory desc) order (equal cpu and memory: equal rank), so that a comparison reads k1, the timestamp
// and the UID's high word — which the chunk kernels keep in shared memory beside the index — and the UID's low word
// only on a tie of all of those.
// -------------------------------------------------------------------------------------------------
struct FfdKeys {
  const uint64_t* k1;   // [P]
  const int64_t* ts;    // [P]
  const uint64_t* uh;   // [P]
  const uint64_t* ul;   // [P]
  int P;
};

// rank[q] = number of request classes that sort strictly before class q
__global__ void __launch_bounds__(256) ffd_class_rank_kernel(PodsDev pods, int32_t* __restrict__ rank) {
  __shared__ int64_t s_cpu[256], s_mem[256];
  const int q = blockIdx.x * 256 + threadIdx.x, NQ = pods.NQ;
  const int64_t* cpu = pods.rq_req + (size_t)pods.res_cpu * NQ;
  const int64_t* mem = pods.rq_req + (size_t)pods.res_mem * NQ;
  const int64_t c = q < NQ ? cpu[q] : 0, m = q < NQ ? mem[q] : 0;
  int before = 0;
  for (int t0 = 0; t0 < NQ; t0 += 256) {
    const int o = t0 + threadIdx.x;
    s_cpu[threadIdx.x] = o < NQ ? cpu[o] : INT64_MIN;
    s_mem[threadIdx.x] = o < NQ ? mem[o] : INT64_MIN;
    __syncthreads();
    const int lim = min(256, NQ - t0);
    for (int j = 0; j < lim; j++) before += (s_cpu[j] > c || (s_cpu[j] == c && s_mem[j] > m)) ? 1 : 0;
    __syncthreads();
  }
  if (q < NQ) rank[q] = before;
}

__global__ void ffd_init_kernel(PodsDev pods, const int32_t* __restrict__ seg, const int32_t* __restrict__ rank,
                                int32_t* __restrict__ idx, uint64_t* __restrict__ k1, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  idx[i] = i;
  if (i < pods.P) k1[i] = ((uint64_t)(uint32_t)(seg ? seg[i] : 0) << 32) | (uint32_t)rank[pods.req_id[i]];
}

__device__ __forceinline__ bool ffd_less_keys(const FfdKeys& K, int a, uint64_t ka, int64_t ta, uint64_t ha, int b,
                                              uint64_t kb, int64_t tb, uint64_t hb) {
  if (ka != kb) return ka < kb;   // padding carries the largest k1
  if (ta != tb) return ta < tb;
  if (ha != hb) return ha < hb;
  if (a < K.P && b < K.P) {
    const uint64_t la = K.ul[a], lb = K.ul[b];
    if (la != lb) return la < lb;
  }
  return a < b;  // identical UIDs cannot occur in a cluster; keeps the order total anyway (and orders the padding)
}

constexpr int SORT_THREADS = 1024;
constexpr int SORT_CHUNK = 2 * SORT_THREADS;  // elements a CTA sorts in shared memory (more per CTA: fewer launches, but the stages of a chunk run on one SM)
constexpr size_t SORT_SMEM = (size_t)SORT_CHUNK * (4 + 8 + 8 + 8);

// all bitonic stages with stride < SORT_CHUNK for merge size k (k_lo..k_hi doubling) inside shared memory
__global__ void __launch_bounds__(SORT_THREADS) ffd_local_kernel(FfdKeys K, int32_t* __restrict__ idx, int k_lo, int k_hi) {
  extern __shared__ __align__(16) unsigned char ffd_smem[];
  uint64_t* s_k = reinterpret_cast<uint64_t*>(ffd_smem);
  int64_t* s_t = reinterpret_cast<int64_t*>(s_k + SORT_CHUNK);
  uint64_t* s_h = reinterpret_cast<uint64_t*>(s_t + SORT_CHUNK);
  int32_t* s_i = reinterpret_cast<int32_t*>(s_h + SORT_CHUNK);
  const int base = blockIdx.x * SORT_CHUNK;
  for (int e = threadIdx.x; e < SORT_CHUNK; e += SORT_THREADS) {
    const int p = idx[base + e];
    s_i[e] = p;
    const bool real = p < K.P;
    s_k[e] = real ? K.k1[p] : ~0ull;
    s_t[e] = real ? K.ts[p] : 0;
    s_h[e] = real ? K.uh[p] : 0ull;
  }
  __syncthreads();
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = min(k >> 1, SORT_CHUNK >> 1); j > 0; j >>= 1) {
#pragma unroll
      for (int rep = 0; rep < SORT_CHUNK / (2 * SORT_THREADS); rep++) {
        const int t = threadIdx.x + rep * SORT_THREADS;
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower element of the pair
        const int l = i + j;
        const bool up = (((base + i) & k) == 0);
        const int a = s_i[i], b = s_i[l];
        const uint64_t ka = s_k[i], kb = s_k[l], ha = s_h[i], hb = s_h[l];
        const int64_t ta = s_t[i], tb = s_t[l];
        const bool sw = up ? ffd_less_keys(K, b, kb, tb, hb, a, ka, ta, ha) : ffd_less_keys(K, a, ka, ta, ha, b, kb, tb, hb);
        if (sw) {
          s_i[i] = b; s_i[l] = a;
          s_k[i] = kb; s_k[l] = ka;
          s_t[i] = tb; s_t[l] = ta;
          s_h[i] = hb; s_h[l] = ha;
        }
      }
      __syncthreads();
    }
  }
  for (int e = threadIdx.x; e < SORT_CHUNK; e += SORT_THREADS) idx[base + e] = s_i[e];
}

// one global bitonic stage (stride j >= SORT_CHUNK) of merge size k
__global__ void ffd_global_kernel(FfdKeys K, int32_t* __restrict__ idx, int n_pad, int j, int k) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad / 2) return;
  int i = ((t / j) * 2 * j) + (t % j);
  int l = i + j;
  bool up = ((i & k) == 0);
  const int a = idx[i], b = idx[l];
  const bool ra = a < K.P, rb = b < K.P;
  const uint64_t ka = ra ? K.k1[a] : ~0ull, kb = rb ? K.k1[b] : ~0ull;
  const int64_t ta = ra ? K.ts[a] : 0, tb = rb ? K.ts[b] : 0;
  const uint64_t ha = ra ? K.uh[a] : 0ull, hb = rb ? K.uh[b] : 0ull;
  const bool sw = up ? ffd_less_keys(K, b, kb, tb, hb, a, ka, ta, ha) : ffd_less_keys(K, a, ka, ta, ha, b, kb, tb, hb);
  if (sw) { idx[i] = b; idx[l] = a; }
}

// -------------------------------------------------------------------------------------------------
// cheapest SKU of a set under the set's zone / capacity-type requirements: one warp per set
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) select_cheapest_kernel(KeyTab kt, int n, const uint64_t* __restrict__ masks,
                                                              const uint64_t* __restrict__ rmask, size_t rms, size_t rks,
                                                              const uint64_t* __restrict__ rflags,
                                                              const int64_t* __restrict__ rgt,
                                                              const int64_t* __restrict__ rlt, size_t bms, size_t bks,
                                                              int32_t* __restrict__ sku_out,
                                                              uint64_t* __restrict__ price_out,
                                                              int32_t* __restrict__ nopt_out,
                                                              const PackProblem* __restrict__ problems,
                                                              const int32_t* __restrict__ row_problem) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  if (problems) {  // rows are NodeClaim slots of a (batched) solve: only the first n_claims of a problem exist
    const PackProblem& pb = problems[row_problem ? row_problem[row] : 0];
    if (row - pb.pod_begin >= pb.n_claims) return;
  }
  const int W = kt.n_words, S = kt.n_skus;
  // allowed zones / capacity types of this set's requirements (element k of row i at rmask[i*rms + k*rks])
  uint64_t za = ~0ull, ca = ~0ull;
  uint64_t flags = rflags[row];
  auto allowed_of = [&](int k) -> uint64_t {
    int b = kt.bslot[k];
    EReq q = make_ereq(rmask[(size_t)row * rms + (size_t)k * rks], flags, k, b >= 0 ? rgt[(size_t)row * bms + (size_t)b * bks] : 0,
                       b >= 0 ? rlt[(size_t)row * bms + (size_t)b * bks] : 0);
    return (q.f & RF_DEF) ? allowed_values(kt, k, q) : ~0ull;
  };
  if (kt.key_zone >= 0) za = allowed_of(kt.key_zone);
  if (kt.key_ct >= 0) ca = allowed_of(kt.key_ct);
  za &= kt.n_zones >= 64 ? ~0ull : ((1ull << kt.n_zones) - 1ull);
  ca &= kt.n_cts >= 64 ? ~0ull : ((1ull << kt.n_cts) - 1ull);
  uint64_t best_p = ~0ull;  // strictly above any price incl. MaxFloat64
  int best_s = -1, count = 0;
  for (int w = lane; w < W; w += 32) {
    uint64_t m = masks[(size_t)row * W + w];
    count += popc64(m);
    while (m) {
      int s = (w << 6) + ctz64(m);
      m &= m - 1;
      uint64_t p = B200S_PRICE_NONE;  // math.MaxFloat64 when no compatible available offering (types.go:224)
      uint64_t z = za;
      while (z) {
        int zi = ctz64(z);
        z &= z - 1;
        uint64_t c = ca;
        while (c) {
          int ci = ctz64(c);
          c &= c - 1;
          uint64_t q = kt.price[(size_t)(zi * kt.n_cts + ci) * S + s];
          p = q < p ? q : p;
        }
      }
      if (p < best_p || (p == best_p && s < best_s)) { best_p = p; best_s = s; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t op = __shfl_down_sync(0xffffffffu, best_p, o);
    int os = __shfl_down_sync(0xffffffffu, best_s, o);
    count += __shfl_down_sync(0xffffffffu, count, o);
    if (os >= 0 && (best_s < 0 || op < best_p || (op == best_p && os < best_s))) { best_p = op; best_s = os; }
  }
  if (lane == 0) {
    sku_out[row] = best_s;
    price_out[row] = best_s >= 0 ? best_p : B200S_PRICE_NONE;
    if (nopt_out) nopt_out[row] = count;
  }
}

// -------------------------------------------------------------------------------------------------
// consolidation commands: computeConsolidation (K/controllers/disruption/consolidation.go:133-307) on
// finished simulations, one warp per simulation.  Prices are compared as IEEE-754 bit patterns (all
// non-negative and finite, so the integer order is the float order).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) consolidation_kernel(KeyTab kt, ConsolidationArgs a, int key_cap) {
  extern __shared__ __align__(16) unsigned char cons_smem[];  // key_cap price keys per warp, indexed by SKU (0: none)
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= a.n) return;
  const int W = kt.n_words, K = kt.n_keys, NB = kt.n_bounded;
  uint64_t* out_mask = a.mask + (size_t)row * W;
  for (int w = lane; w < W; w += 32) out_mask[w] = 0;
  if (lane < B200S_S2S_MIN_TYPES) a.head[(size_t)row * B200S_S2S_MIN_TYPES + lane] = -1;
  int decision = B200S_CMD_NONE, why = B200S_WHY_OK, n_options = 0;
  bool spot_required = false, tie = false;
  if (a.n_unscheduled[row] > 0) {
    why = B200S_WHY_UNSCHEDULABLE;
  } else if (a.n_claims[row] == 0) {
    decision = B200S_CMD_DELETE;
  } else if (a.n_claims[row] != 1) {
    why = B200S_WHY_MANY_CLAIMS;
  } else {
    // zones / capacity types the replacement NodeClaim's requirements admit
    const uint64_t flags = a.rflags[row];
    auto allowed_of = [&](int k) -> uint64_t {
      const int b = kt.bslot[k];
      EReq q = make_ereq(a.rmask[(size_t)row * K + k], flags, k, b >= 0 ? a.rgt[(size_t)row * NB + b] : 0,
                         b >= 0 ? a.rlt[(size_t)row * NB + b] : 0);
      return (q.f & RF_DEF) ? allowed_values(kt, k, q) : ~0ull;
    };
    uint64_t za = kt.key_zone >= 0 ? allowed_of(kt.key_zone) : ~0ull;
    uint64_t ca = allowed_of(kt.key_ct);
    za &= kt.n_zones >= 64 ? ~0ull : ((1ull << kt.n_zones) - 1ull);
    ca &= kt.n_cts >= 64 ? ~0ull : ((1ull << kt.n_cts) - 1ull);
    const bool has_spot = (ca >> a.ct_spot) & 1ull, has_od = (ca >> a.ct_on_demand) & 1ull;
    const bool s2s = a.all_spot[row] && has_spot;  // consolidation.go:189-192
    if (s2s && !a.spot_to_spot) {
      why = B200S_WHY_S2S_DISABLED;
    } else {
      // spot-to-spot first narrows the requirements to In{spot} (:240); InstanceTypes.Compatible (:242) is implied by
      // the price filter, whose MaxFloat64 for "no compatible offering" is never below a finite candidate price
      const uint64_t cf = s2s ? (1ull << a.ct_spot) : ca;
      const uint64_t limit = a.candidate_price[row];
      int count = 0;
      for (int w = lane; w < W; w += 32) {
        uint64_t keep = 0;
        for (uint64_t m = a.claim_mask[(size_t)row * W + w]; m; m &= m - 1) {
          const int b = ctz64(m);
          if (worst_launch_price(kt, za, cf, a.ct_reserved, a.ct_spot, a.ct_on_demand, (w << 6) + b) < limit) keep |= 1ull << b;
        }
        out_mask[w] = keep;
        count += popc64(keep);
      }
      for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
      if (count == 0) {
        why = B200S_WHY_NOT_CHEAPER;
      } else if (s2s && a.n_candidates[row] == 1) {
        if (count < B200S_S2S_MIN_TYPES) {
          why = B200S_WHY_S2S_TOO_FEW;
          for (int w = lane; w < W; w += 32) out_mask[w] = 0;
        } else {
          // keep the 15 cheapest in OrderByPrice order, which was taken under the requirements BEFORE the narrowing
          // (:187): 16 rounds of "smallest (price, index) above the previous pick"; lane i remembers pick i
          uint64_t last_p = 0, my_p = 0;
          int last_s = -1, my_s = -1;
          const int rounds = count > B200S_S2S_MIN_TYPES ? B200S_S2S_MIN_TYPES + 1 : B200S_S2S_MIN_TYPES;
          // every round looks at every survivor's key: computed once into this warp's slice of shared memory when the
          // catalog fits (a lane only ever reads the keys of its own words, so no warp synchronisation is involved)
          uint64_t* keys = (uint64_t*)cons_smem + (size_t)(threadIdx.x >> 5) * key_cap;
          const bool cached = key_cap >= W * 64;
          if (cached)
            for (int w = lane; w < W; w += 32)
              for (uint64_t m = out_mask[w]; m; m &= m - 1) {
                const int s = (w << 6) + ctz64(m);
                keys[s] = cheapest_price(kt, za, ca, s);
              }
          for (int rd = 0; rd < rounds; rd++) {
            uint64_t best_p = ~0ull;
            int best_s = -1;
            for (int w = lane; w < W; w += 32)
              for (uint64_t m = out_mask[w]; m; m &= m - 1) {
                const int s = (w << 6) + ctz64(m);
                const uint64_t p = cached ? keys[s] : cheapest_price(kt, za, ca, s);
                const bool above = rd == 0 || p > last_p || (p == last_p && s > last_s);
                if (above && (p < best_p || (p == best_p && s < best_s))) { best_p = p; best_s = s; }
              }
            for (int o = 16; o > 0; o >>= 1) {
              uint64_t op = __shfl_xor_sync(0xffffffffu, best_p, o);
              int os = __shfl_xor_sync(0xffffffffu, best_s, o);
              if (os >= 0 && (best_s < 0 || op < best_p || (op == best_p && os < best_s))) { best_p = op; best_s = os; }
            }
            last_p = best_p; last_s = best_s;
            if (lane == rd) { my_p = best_p; my_s = best_s; }
          }
          const uint64_t next_p = __shfl_down_sync(0xffffffffu, my_p, 1);
          tie = __ballot_sync(0xffffffffu, lane + 1 < rounds && my_p == next_p) != 0;
          __syncwarp();
          for (int w = lane; w < W; w += 32) out_mask[w] = 0;
          __syncwarp();
          if (lane < B200S_S2S_MIN_TYPES) {
            atomicOr((unsigned long long*)&out_mask[my_s >> 6], 1ull << (my_s & 63));
            a.head[(size_t)row * B200S_S2S_MIN_TYPES + lane] = my_s;
          }
          decision = B200S_CMD_REPLACE;
          n_options = B200S_S2S_MIN_TYPES;
        }
      } else {
        decision = B200S_CMD_REPLACE;
        n_options = count;
      }
      // OD -> [OD, spot]: the filter assumed the spot price, so the launch must not fall back to on-demand (:218-221)
      spot_required = decision == B200S_CMD_REPLACE && (s2s || (has_spot && has_od));
    }
  }
  if (lane == 0) {
    a.decision[row] = decision;
    a.reason[row] = why;
    a.n_options[row] = n_options;
    a.spot_required[row] = spot_required ? 1 : 0;
    a.price_tie[row] = tie ? 1 : 0;
  }
}

// -------------------------------------------------------------------------------------------------
// NodeClaimTemplate.ToNodeClaim's instance-type list (nodeclaimtemplate.go:77-82): the whole OrderByPrice(reqs)
// order of a set, cut at max_types.  One CTA per set: compact the set bits, one key per type, bitonic sort of
// (price bits, SKU index) pairs in shared memory.
// -------------------------------------------------------------------------------------------------
constexpr int ORDER_THREADS = 256;

__global__ void __launch_bounds__(ORDER_THREADS) order_by_price_kernel(KeyTab kt, int n, int n_pad, int max_types,
                                                                       const uint64_t* __restrict__ masks,
                                                                       const uint64_t* __restrict__ rmask,
                                                                       const uint64_t* __restrict__ rflags,
                                                                       const int64_t* __restrict__ rgt,
                                                                       const int64_t* __restrict__ rlt,
                                                                       int32_t* __restrict__ order_out,
                                                                       int32_t* __restrict__ count_out,
                                                                       uint8_t* __restrict__ tie_out) {
  extern __shared__ __align__(16) unsigned char order_smem[];
  uint64_t* key = (uint64_t*)order_smem;        // [n_pad]
  int32_t* sku = (int32_t*)(key + n_pad);        // [n_pad]
  __shared__ int word_off[65];                   // exclusive prefix of the words' popcounts, per 64-word chunk
  __shared__ int s_count, s_tie;
  const int row = blockIdx.x, tid = threadIdx.x;
  const int W = kt.n_words, K = kt.n_keys, NB = kt.n_bounded;
  const uint64_t flags = rflags[row];
  auto allowed_of = [&](int k) -> uint64_t {
    const int b = kt.bslot[k];
    EReq q = make_ereq(rmask[(size_t)row * K + k], flags, k, b >= 0 ? rgt[(size_t)row * NB + b] : 0,
                       b >= 0 ? rlt[(size_t)row * NB + b] : 0);
    return (q.f & RF_DEF) ? allowed_values(kt, k, q) : ~0ull;
  };
  uint64_t za = kt.key_zone >= 0 ? allowed_of(kt.key_zone) : ~0ull;
  uint64_t ca = kt.key_ct >= 0 ? allowed_of(kt.key_ct) : ~0ull;
  za &= kt.n_zones >= 64 ? ~0ull : ((1ull << kt.n_zones) - 1ull);
  ca &= kt.n_cts >= 64 ? ~0ull : ((1ull << kt.n_cts) - 1ull);
  if (tid == 0) { s_count = 0; s_tie = 0; }
  for (int i = tid; i < n_pad; i += ORDER_THREADS) { key[i] = ~0ull; sku[i] = 0x7fffffff; }
  __syncthreads();
  // compaction, 64 words at a time: thread w of the chunk owns one word
  for (int w0 = 0; w0 < W; w0 += 64) {
    const int w = w0 + tid;
    const uint64_t m = (tid < 64 && w < W) ? masks[(size_t)row * W + w] : 0;
    if (tid < 64) word_off[tid + 1] = popc64(m);
    __syncthreads();
    if (tid == 0) {
      word_off[0] = s_count;
      for (int i = 0; i < 64; i++) word_off[i + 1] += word_off[i];
      s_count = word_off[64];
    }
    __syncthreads();
    if (tid < 64) {
      int at = word_off[tid];
      for (uint64_t mm = m; mm; mm &= mm - 1) sku[at++] = (w << 6) + ctz64(mm);
    }
    __syncthreads();
  }
  const int count = s_count;
  for (int i = tid; i < count; i += ORDER_THREADS) key[i] = cheapest_price(kt, za, ca, sku[i]);
  __syncthreads();
  // bitonic sort, ascending by (price, index); the padding sorts last
  for (int k = 2; k <= n_pad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < n_pad / 2; t += ORDER_THREADS) {
        const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
        const bool up = (i & k) == 0;
        const uint64_t ka = key[i], kb = key[l];
        const int sa = sku[i], sb = sku[l];
        const bool a_gt_b = ka > kb || (ka == kb && sa > sb);
        if (a_gt_b == up) { key[i] = kb; key[l] = ka; sku[i] = sb; sku[l] = sa; }
      }
      __syncthreads();
    }
  const int listed = count < max_types ? count : max_types;
  bool tie = false;
  for (int i = tid; i < max_types; i += ORDER_THREADS) {
    order_out[(size_t)row * max_types + i] = i < listed ? sku[i] : -1;
    if (i < listed && i + 1 < count && key[i] == key[i + 1]) tie = true;  // sort.Slice is not stable: Go's order is its own
  }
  if (tie) s_tie = 1;
  __syncthreads();
  if (tid == 0) { count_out[row] = listed; tie_out[row] = (uint8_t)s_tie; }
}

// -------------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------------
static void pick_group(int W, int* G, int* WPL) {
  if (W <= 8) { *G = 8; *WPL = 1; }
  else if (W <= 16) { *G = 16; *WPL = 1; }
  else if (W <= 32) { *G = 32; *WPL = 1; }
  else if (W <= 64) { *G = 32; *WPL = 2; }
  else { *G = 32; *WPL = 4; }
}

#define DISPATCH_GROUP(W, CALL)                                  \
  do {                                                           \
    int G_, WPL_;                                                \
    pick_group(W, &G_, &WPL_);                                   \
    if (G_ == 8) { CALL(8, 1); }                                 \
    else if (G_ == 16) { CALL(16, 1); }                          \
    else if (WPL_ == 1) { CALL(32, 1); }                         \
    else if (WPL_ == 2) { CALL(32, 2); }                         \
    else { CALL(32, 4); }                                        \
  } while (0)

cudaError_t launch_template_prefilter(const KeyTab& kt, const TmplDev* tg, const uint64_t* type_admit, uint64_t* tmask,
                                      int32_t* alive, int32_t* tm_count, cudaStream_t stream) {
  if (kt.n_words > 128) return cudaErrorInvalidValue;
#define CALL(G, WPL) template_prefilter_kernel<G, WPL><<<1, 32 * MAXT, 0, stream>>>(kt, tg, type_admit, tmask, alive, tm_count)
  DISPATCH_GROUP(kt.n_words, CALL);
#undef CALL
  return cudaGetLastError();
}

static size_t class_smem_bytes(const KeyTab& kt, int G, int in_smem) {
  size_t o = align16(sizeof(TmplDev)) + align16(sizeof(EvalScratch) * (CLS_THREADS / G));
  if (in_smem) o += align16((size_t)kt.n_rows * kt.n_words * 8 + (size_t)kt.n_dv * 8);
  return o;
}

cudaError_t launch_class_tables(const KeyTab& kt, const TmplDev* tg, int T, const uint64_t* tmask, const int32_t* alive,
                                const PodsDev& pods, uint8_t* sg_status, uint64_t* sgC, uint64_t* sgO, uint64_t* sgCO,
                                uint64_t* rqF, int32_t* tm_count, int n_sm, int smem_limit, cudaStream_t stream) {
  if (kt.n_words > 128) return cudaErrorInvalidValue;
  (void)tmask;
  (void)tm_count;  // filled by the template pre-filter
  long long items = ((long long)pods.NS + pods.NQ) * T;
  if (items == 0) return cudaGetLastError();
  int G, WPL;
  pick_group(kt.n_words, &G, &WPL);
  (void)smem_limit;
  int in_smem = 0;  // tables are read through the non-coherent L1 path (ldg), which shared-memory copies cannot use
  size_t bytes = class_smem_bytes(kt, G, in_smem);
  cudaError_t err = cudaSuccess;
#define CALL(G_, WPL_)                                                                                      \
  do {                                                                                                      \
    auto kern = class_tables_kernel<G_, WPL_>;                                                              \
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);              \
    if (err != cudaSuccess) return err;                                                                     \
    long long need = (items + (CLS_THREADS / G_) - 1) / (CLS_THREADS / G_);                                 \
    int grid = (int)(need < (long long)n_sm * 4 ? need : (long long)n_sm * 4);                              \
    kern<<<grid, CLS_THREADS, bytes, stream>>>(kt, tg, tmask, alive, pods, sg_status, sgC, sgO, sgCO, rqF, in_smem); \
  } while (0)
  DISPATCH_GROUP(kt.n_words, CALL);
#undef CALL
  return cudaGetLastError();
}

cudaError_t launch_fit_dense(const KeyTab& kt, int T, const PodsDev& pods, const uint8_t* sg_status, const uint64_t* sgC,
                             const uint64_t* sgO, const uint64_t* sgCO, const uint64_t* rqF, const int32_t* tm_count,
                             int32_t* tmpl_out,
                             uint64_t* mask_out, uint8_t* code_out, unsigned long long* evals, int n_sm,
                             cudaStream_t stream) {
  if (kt.n_words > 128) return cudaErrorInvalidValue;
  if (pods.P == 0) return cudaSuccess;
  const long long W2l = kt.n_words / 2;
  const bool small = (long long)pods.P * W2l < (1ll << 31) && (long long)pods.NS * T * W2l < (1ll << 31) &&
                     (long long)pods.NQ * T * W2l < (1ll << 31) && (long long)pods.P * T < (1ll << 31);
  if ((kt.n_words & 1) == 0 && small) {  // 16-byte rows, 32-bit indices: the vector kernel
    const int W2 = kt.n_words / 2;
    // Rows through bulk asynchronous copies (fit_dense_bulk_kernel): measured slower than plain 16-byte stores on
    // 1M x 2k (stage 163 us against 153 us) — kept as an opt-in experiment, see profiles/README.md
    const bool bulk = getenv("B200S_DENSE_BULK") != nullptr;
    DenseArgs da{pods.P, T, W2, pods.sig_id, pods.req_id, sg_status, reinterpret_cast<const ulonglong2*>(sgC),
                 reinterpret_cast<const ulonglong2*>(sgO), reinterpret_cast<const ulonglong2*>(sgCO),
                 reinterpret_cast<const ulonglong2*>(rqF), tm_count, tmpl_out, reinterpret_cast<ulonglong2*>(mask_out),
                 code_out, evals};
#define VCALL(LPG_, VPL_)                                                                                        \
  do {                                                                                                           \
    long long need = ((long long)pods.P + (FIT_THREADS / LPG_) - 1) / (FIT_THREADS / LPG_);                      \
    int grid = (int)(need < (long long)n_sm * 8 ? need : (long long)n_sm * 8);                                   \
    const size_t bulk_smem = (size_t)(FIT_THREADS / 32) * 2 * (32 / LPG_) * W2 * 16;                             \
    if (mask_out && bulk && bulk_smem * 8 <= 200 * 1024) {                                                       \
      fit_dense_bulk_kernel<LPG_, VPL_><<<grid, FIT_THREADS, bulk_smem, stream>>>(da);                           \
    } else {                                                                                                     \
      fit_dense_vec_kernel<LPG_, VPL_><<<grid, FIT_THREADS, 0, stream>>>(da);                                    \
    }                                                                                                            \
  } while (0)
    // few lanes per pod, two 16-byte pieces per lane; measured on 1M x 2k (W2 = 16): 8x2 78 us, 4x4 94 us,
    // 16x1 118 us, 2x8 147 us (profiles/README.md)
    if (W2 <= 4) VCALL(4, 1);
    else if (W2 <= 8) VCALL(4, 2);
    else if (W2 <= 16) VCALL(8, 2);
    else if (W2 <= 32) VCALL(16, 2);
    else VCALL(32, 2);
#undef VCALL
    return cudaGetLastError();
  }
#define CALL(G_, WPL_)                                                                                          \
  do {                                                                                                          \
    long long need = ((long long)pods.P + (FIT_THREADS / G_) - 1) / (FIT_THREADS / G_);                         \
    int grid = (int)(need < (long long)n_sm * 8 ? need : (long long)n_sm * 8);                                  \
    fit_dense_kernel<G_, WPL_><<<grid, FIT_THREADS, 0, stream>>>(pods.P, T, kt.n_words, pods.sig_id, pods.req_id, \
                                                                 sg_status, sgC, sgO, rqF, tm_count, tmpl_out,  \
                                                                 mask_out, code_out, evals);                    \
  } while (0)
  DISPATCH_GROUP(kt.n_words, CALL);
#undef CALL
  return cudaGetLastError();
}

static int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

int ffd_padded(int P) { return next_pow2(P < SORT_CHUNK ? SORT_CHUNK : P); }

// ExistingNodes of a batch of simulations that share one cluster (b200s_solve_batch_shared): problem b gets rows
// [b*E0, (b+1)*E0) of the batch-wide, resource/key-major arrays — a copy of the shared rows
__global__ void __launch_bounds__(256) replicate_nodes_kernel(int R, int K, int E0, long long Etot, const int64_t* __restrict__ s_avail,
                                                              const uint64_t* __restrict__ s_rmask, const uint64_t* __restrict__ s_rflags,
                                                              const uint64_t* __restrict__ s_taints, int64_t* __restrict__ d_avail,
                                                              uint64_t* __restrict__ d_rmask, uint64_t* __restrict__ d_rflags,
                                                              uint64_t* __restrict__ d_taints) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Etot; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % E0);
    for (int x = 0; x < R; x++) d_avail[(size_t)x * Etot + i] = s_avail[(size_t)x * E0 + e];
    for (int k = 0; k < K; k++) d_rmask[(size_t)k * Etot + i] = s_rmask[(size_t)k * E0 + e];
    d_rflags[i] = s_rflags[e];
    d_taints[i] = s_taints[e];
  }
}

// a node a simulation removes: nothing fits a negative Available entry (resources.go:234-243), so the row is skipped
// exactly as if it were not in the list
__global__ void exclude_nodes_kernel(int R, long long Etot, int n, const int64_t* __restrict__ rows, int64_t* __restrict__ d_avail) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  for (int x = 0; x < R; x++) d_avail[(size_t)x * Etot + rows[j]] = -1;
}

cudaError_t launch_replicate_nodes(int R, int K, int E0, long long Etot, const int64_t* s_avail, const uint64_t* s_rmask,
                                   const uint64_t* s_rflags, const uint64_t* s_taints, int64_t* d_avail, uint64_t* d_rmask,
                                   uint64_t* d_rflags, uint64_t* d_taints, int n_excl, const int64_t* excl_rows, int n_sm,
                                   cudaStream_t stream) {
  if (Etot <= 0) return cudaSuccess;
  const int grid = (int)std::min<long long>((Etot + 255) / 256, (long long)n_sm * 8);
  replicate_nodes_kernel<<<grid, 256, 0, stream>>>(R, K, E0, Etot, s_avail, s_rmask, s_rflags, s_taints, d_avail, d_rmask, d_rflags, d_taints);
  if (n_excl > 0) exclude_nodes_kernel<<<(n_excl + 255) / 256, 256, 0, stream>>>(R, Etot, n_excl, excl_rows, d_avail);
  return cudaGetLastError();
}

cudaError_t launch_ffd_class_rank(const PodsDev& pods, int32_t* rank, cudaStream_t stream) {
  if (pods.NQ <= 0) return cudaSuccess;
  ffd_class_rank_kernel<<<(pods.NQ + 255) / 256, 256, 0, stream>>>(pods, rank);
  return cudaGetLastError();
}

cudaError_t launch_ffd_sort(const PodsDev& pods, const int32_t* seg, const int32_t* rank, uint64_t* k1, int32_t* idx,
                            int* launches, cudaStream_t stream) {
  const int n_pad = ffd_padded(pods.P);
  int nl = 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ffd_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SORT_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  ffd_init_kernel<<<(n_pad + 255) / 256, 256, 0, stream>>>(pods, seg, rank, idx, k1, n_pad);
  nl++;
  const FfdKeys K{k1, pods.ts, pods.uid_hi, pods.uid_lo, pods.P};
  // sort every chunk in shared memory (merge sizes 2 .. SORT_CHUNK)
  ffd_local_kernel<<<n_pad / SORT_CHUNK, SORT_THREADS, SORT_SMEM, stream>>>(K, idx, 2, SORT_CHUNK);
  nl++;
  for (int k = SORT_CHUNK * 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j >= SORT_CHUNK; j >>= 1) {
      ffd_global_kernel<<<(n_pad / 2 + 255) / 256, 256, 0, stream>>>(K, idx, n_pad, j, k);
      nl++;
    }
    ffd_local_kernel<<<n_pad / SORT_CHUNK, SORT_THREADS, SORT_SMEM, stream>>>(K, idx, k, k);
    nl++;
  }
  if (launches) *launches = nl;
  return cudaGetLastError();
}

cudaError_t launch_select_cheapest(const KeyTab& kt, int n, const uint64_t* masks, const uint64_t* rmask, size_t rms,
                                   size_t rks, const uint64_t* rflags, const int64_t* rgt, const int64_t* rlt, size_t bms,
                                   size_t bks, int32_t* sku_out, uint64_t* price_out, int32_t* nopt_out,
                                   const PackProblem* problems, const int32_t* row_problem, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  int blocks = (int)(((long long)n * 32 + 255) / 256);
  select_cheapest_kernel<<<blocks, 256, 0, stream>>>(kt, n, masks, rmask, rms, rks, rflags, rgt, rlt, bms, bks, sku_out,
                                                     price_out, nopt_out, problems, row_problem);
  return cudaGetLastError();
}

int order_by_price_padded(int S) {
  int p = 2;
  while (p < S) p <<= 1;
  return p;
}

cudaError_t launch_order_by_price(const KeyTab& kt, int n, int max_types, const uint64_t* masks, const uint64_t* rmask,
                                  const uint64_t* rflags, const int64_t* rgt, const int64_t* rlt, int32_t* order_out,
                                  int32_t* count_out, uint8_t* tie_out, int smem_limit, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int n_pad = order_by_price_padded(kt.n_skus);
  const size_t bytes = (size_t)n_pad * 12;
  if ((long long)bytes > smem_limit - 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(order_by_price_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  order_by_price_kernel<<<n, ORDER_THREADS, bytes, stream>>>(kt, n, n_pad, max_types, masks, rmask, rflags, rgt, rlt, order_out,
                                                             count_out, tie_out);
  return cudaGetLastError();
}

cudaError_t launch_consolidation(const KeyTab& kt, const ConsolidationArgs& a, int smem_limit, cudaStream_t stream) {
  if (a.n <= 0) return cudaSuccess;
  const int blocks = (a.n * 32 + 255) / 256;
  int key_cap = kt.n_words * 64;  // per warp; 8 warps per CTA
  size_t bytes = (size_t)key_cap * 8 * 8;
  if ((long long)bytes > smem_limit - 2048) { key_cap = 0; bytes = 0; }  // large catalog: keys recomputed per round
  cudaError_t e = cudaFuncSetAttribute(consolidation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  consolidation_kernel<<<blocks, 256, bytes, stream>>>(kt, a, key_cap);
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// compact stage-A rows: what another GPU needs of a pod's row to run the pack — the first feasible template and the
// per-template attempt codes.  The instance-type row itself is a function of (signature, request class, template)
// and the replicated class tables, so it is never sent (one all-gather of the result rows).
// row = [int8 template | T code bytes | padding], `rb` bytes (compact_row_bytes(T))
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_export_kernel(int P, int T, int rb, const int32_t* __restrict__ tmpl,
                                                             const uint8_t* __restrict__ code, uint8_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)P * rb; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / rb), j = (int)(i % rb);
    uint8_t v = 0;
    if (j == 0) v = (uint8_t)(int8_t)tmpl[p];
    else if (j <= T) v = code[(size_t)p * T + (j - 1)];
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) compact_import_kernel(int P, int T, int rb, const uint8_t* __restrict__ in,
                                                             int32_t* __restrict__ tmpl, uint8_t* __restrict__ code) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)P * (T + 1); i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / (T + 1)), j = (int)(i % (T + 1));
    const uint8_t v = in[(size_t)p * rb + j];
    if (j == 0) tmpl[p] = (int32_t)(int8_t)v;
    else code[(size_t)p * T + (j - 1)] = v;
  }
}

int compact_row_bytes(int T) { return T <= 7 ? 8 : 24; }

// The same export fused with its all-gather: every rank stores its shard's rows straight into every rank's gathered
// buffer over NVLink peer mappings (8 bytes a pod: the whole exchange is smaller than one collective's fixed cost), then
// the ranks meet at a flag barrier — rank `me` publishes `epoch` in slot `me` of every rank's flag array once its stores
// are out, and waits until its own array shows `epoch` in every slot.  peer.out[q] / peer.flags[q] are rank q's buffer and
// flag array as mapped into THIS process (own memory for q == me).
struct CompactPeers {
  uint8_t* out[8];
  unsigned* flags[8];
};

__global__ void __launch_bounds__(256) compact_export_peers_kernel(int P, int T, int rb, const int32_t* __restrict__ tmpl,
                                                                   const uint8_t* __restrict__ code, CompactPeers peer, int n_peers,
                                                                   long long row_off) {
  const int words = rb / 8;  // 8-byte words of a row
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    unsigned long long w[3] = {0ull, 0ull, 0ull};
    w[0] = (unsigned long long)(uint8_t)(int8_t)tmpl[p];
    for (int j = 1; j <= T; j++) w[j >> 3] |= (unsigned long long)code[(size_t)p * T + (j - 1)] << (8 * (j & 7));
    for (int q = 0; q < n_peers; q++) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(peer.out[q] + (size_t)(row_off + p) * rb);
      for (int i = 0; i < words; i++) dst[i] = w[i];
    }
  }
}

// one thread per peer: signal it, then wait for it.  *timed_out is set when a peer does not show up (the caller falls back
// to a collective); the wait is bounded so that a missing peer cannot hang the GPU.
__global__ void compact_peers_barrier_kernel(CompactPeers peer, int n_peers, int me, unsigned epoch, int* __restrict__ timed_out) {
  const int q = threadIdx.x;
  if (q >= n_peers) return;
  __threadfence_system();  // this rank's row stores (the kernel before this one) are visible system-wide before the flag
  unsigned* theirs = peer.flags[q] + me;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const unsigned* mine = peer.flags[me] + q;
  for (unsigned spins = 0;; spins++) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (spins > (1u << 22)) {  // ~ a second
      *timed_out = 1;
      break;
    }
    __nanosleep(200);
  }
}

cudaError_t launch_compact_export_peers(int P, int T, const int32_t* tmpl, const uint8_t* code, int n_peers, int me,
                                        void* const* peer_out, void* const* peer_flags, long long row_off, unsigned epoch,
                                        int* timed_out, int n_sm, cudaStream_t stream) {
  if (n_peers < 1 || n_peers > 8 || me < 0 || me >= n_peers) return cudaErrorInvalidValue;
  CompactPeers cp{};
  for (int q = 0; q < n_peers; q++) {
    cp.out[q] = static_cast<uint8_t*>(peer_out[q]);
    cp.flags[q] = static_cast<unsigned*>(peer_flags[q]);
  }
  const int rb = compact_row_bytes(T);
  if (P > 0) {
    const int grid = (int)std::min<long long>(((long long)P + 255) / 256, (long long)n_sm * 8);
    compact_export_peers_kernel<<<grid, 256, 0, stream>>>(P, T, rb, tmpl, code, cp, n_peers, row_off);
  }
  compact_peers_barrier_kernel<<<1, 32, 0, stream>>>(cp, n_peers, me, epoch, timed_out);
  return cudaGetLastError();
}



cudaError_t launch_compact_export(int P, int T, const int32_t* tmpl, const uint8_t* code, uint8_t* out, int n_sm, cudaStream_t stream) {
  if (P == 0) return cudaSuccess;
  const int rb = compact_row_bytes(T);
  const long long n = (long long)P * rb;
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)n_sm * 8);
  compact_export_kernel<<<grid, 256, 0, stream>>>(P, T, rb, tmpl, code, out);
  return cudaGetLastError();
}

cudaError_t launch_compact_import(int P, int T, const uint8_t* in, int32_t* tmpl, uint8_t* code, int n_sm, cudaStream_t stream) {
  if (P == 0) return cudaSuccess;
  const int rb = compact_row_bytes(T);
  const long long n = (long long)P * (T + 1);
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)n_sm * 8);
  compact_import_kernel<<<grid, 256, 0, stream>>>(P, T, rb, in, tmpl, code);
  return cudaGetLastError();
}

}  // namespace b200

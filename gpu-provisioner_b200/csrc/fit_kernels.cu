// fit_kernels.cu — sm_100a kernels of the dense side of the path:
//   template_prefilter_kernel  NewScheduler's per-NodePool instance-type pre-filter (scheduler.go:143-158)
//   fit_matrix_kernel          addToNewNodeClaim's template walk for every pending pod (scheduler.go:587-675,
//                              nodeclaim.go:114-163, :373-441): per-pod feasibility mask over SKUs, the first
//                              feasible template, the per-template attempt codes
//   ffd_* kernels              NewQueue's order (queue.go:37-43, :72-108): bitonic sort of pod indices on
//                              (cpu desc, memory desc, creationTimestamp asc, UID asc) — a total order
//   select_cheapest_kernel     OrderByPrice(reqs)[0] (cloudprovider/types.go:221-240) per SKU set
//
// All integer / compare / bitset work: no tensor cores, no floating-point arithmetic (prices are compared
// as their IEEE-754 bit patterns, which is order-preserving for the non-negative finite prices the
// encoder admits).  Bit-row tables are staged in shared memory; pods stream through in SoA tiles.
#include "device_types.cuh"
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
                                                                       int32_t* __restrict__ alive) {
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
  unsigned any = 0;
#pragma unroll
  for (int i = 0; i < WPL; i++) {
    int w = gl + i * G;
    if (w < kt.n_words) {
      uint64_t v = stt == ST_OK ? out[i] : 0ull;
      tmask[(size_t)t * kt.n_words + w] = v;
      any |= v ? 1u : 0u;
    }
  }
  any = __reduce_or_sync(group_mask<G>(), any);
  if (gl == 0) alive[t] = any ? 1 : 0;
}

// -------------------------------------------------------------------------------------------------
// dense stage A
// -------------------------------------------------------------------------------------------------
constexpr int FIT_THREADS = 256;
constexpr int FIT_TILE = 128;  // pods staged per CTA iteration

struct FitSmemLayout {
  size_t tmpl, scratch, tmask, alive, tile, tables, total;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline FitSmemLayout fit_layout(int G, int R, int K, int NB, int T, int W, int n_rows, int n_dv,
                                                    int tables_in_smem) {
  FitSmemLayout L;
  size_t o = 0;
  L.tmpl = o; o += align16(sizeof(TmplDev));
  L.scratch = o; o += align16(sizeof(EvalScratch) * (FIT_THREADS / G));
  L.tmask = o; o += align16((size_t)T * W * 8);
  L.alive = o; o += align16((size_t)MAXT * 4);
  L.tile = o;
  // per pod: R requests, K masks, flags, tolerated (8 B each) + 2*NB int32 bounds
  o += align16((size_t)FIT_TILE * ((size_t)(R + K + 2) * 8 + (size_t)NB * 8));
  L.tables = o;
  if (tables_in_smem) o += align16((size_t)n_rows * W * 8 + (size_t)n_dv * 8);
  L.total = o;
  return L;
}

template <int G, int WPL>
__global__ void __launch_bounds__(FIT_THREADS) fit_matrix_kernel(KeyTab kt, const TmplDev* __restrict__ tg,
                                                                 const uint64_t* __restrict__ tmask_g,
                                                                 const int32_t* __restrict__ alive_g, PodsDev pods,
                                                                 int32_t* __restrict__ tmpl_out,
                                                                 uint64_t* __restrict__ mask_out,
                                                                 uint8_t* __restrict__ code_out,
                                                                 unsigned long long* __restrict__ evals,
                                                                 int tables_in_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NG = FIT_THREADS / G;
  const int R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded, W = kt.n_words;
  const int T = tg->T;
  const FitSmemLayout L = fit_layout(G, R, K, NB, T, W, kt.n_rows, kt.n_dv, tables_in_smem);
  TmplDev* tm = reinterpret_cast<TmplDev*>(smem + L.tmpl);
  EvalScratch* scratch = reinterpret_cast<EvalScratch*>(smem + L.scratch);
  uint64_t* tmask = reinterpret_cast<uint64_t*>(smem + L.tmask);
  int32_t* alive = reinterpret_cast<int32_t*>(smem + L.alive);
  // tile arrays (SoA inside the tile, stride FIT_TILE)
  int64_t* t_req = reinterpret_cast<int64_t*>(smem + L.tile);
  uint64_t* t_mask = reinterpret_cast<uint64_t*>(t_req + (size_t)R * FIT_TILE);
  uint64_t* t_flags = t_mask + (size_t)K * FIT_TILE;
  uint64_t* t_tol = t_flags + FIT_TILE;
  int32_t* t_gt = reinterpret_cast<int32_t*>(t_tol + FIT_TILE);
  int32_t* t_lt = t_gt + (size_t)NB * FIT_TILE;

  copy_words(tm, tg, sizeof(TmplDev));
  copy_words(tmask, tmask_g, (size_t)T * W * 8);
  if (threadIdx.x < MAXT) alive[threadIdx.x] = threadIdx.x < T ? alive_g[threadIdx.x] : 0;
  stage_tables(kt, smem + L.tables, tables_in_smem);
  __syncthreads();

  const int g = threadIdx.x / G, gl = threadIdx.x & (G - 1);
  const int P = pods.P;
  const int n_tiles = (P + FIT_TILE - 1) / FIT_TILE;
  unsigned long long my_evals = 0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int p0 = tile * FIT_TILE;
    const int np = min(FIT_TILE, P - p0);
    // coalesced SoA tile load: each array row is np contiguous words in HBM
    for (int i = threadIdx.x; i < (R + K + 2) * FIT_TILE; i += FIT_THREADS) {
      int row = i / FIT_TILE, j = i - row * FIT_TILE;
      if (j < np) {
        uint64_t v;
        if (row < R) v = (uint64_t)pods.req[(size_t)row * P + p0 + j];
        else if (row < R + K) v = pods.rmask[(size_t)(row - R) * P + p0 + j];
        else if (row == R + K) v = pods.rflags[p0 + j];
        else v = pods.tol[p0 + j];
        reinterpret_cast<uint64_t*>(t_req)[i] = v;
      }
    }
    for (int i = threadIdx.x; i < 2 * NB * FIT_TILE; i += FIT_THREADS) {
      int row = i / FIT_TILE, j = i - row * FIT_TILE;
      if (j < np) t_gt[i] = row < NB ? pods.gt[(size_t)row * P + p0 + j] : pods.lt[(size_t)(row - NB) * P + p0 + j];
    }
    __syncthreads();

    for (int j = g; j < np; j += NG) {
      const int p = p0 + j;
      PodRef pd{t_mask + j, FIT_TILE, t_flags[j], t_gt + j, t_lt + j, FIT_TILE, t_req + j, FIT_TILE, t_tol[j]};
      int chosen = -1;
      uint64_t keep[WPL];
#pragma unroll
      for (int i = 0; i < WPL; i++) keep[i] = 0;
      for (int t = 0; t < T; t++) {
        uint8_t code = alive[t] ? B200S_FIT_NONE : B200S_FIT_DROPPED;
        if (alive[t] && chosen < 0) {
          StateRef st{tm->rmask[t], 1, tm->rflags[t], tm->gt[t], tm->lt[t], 1, tm->daemon[t], 1, tm->taints[t]};
          uint64_t cand[WPL], out[WPL];
#pragma unroll
          for (int i = 0; i < WPL; i++) {
            int w = gl + i * G;
            cand[i] = w < W ? tmask[(size_t)t * W + w] : 0ull;
          }
          int visited;
          int stt = group_can_add<G, WPL>(kt, &scratch[g], st, pd, true, cand, out, &code, &visited);
          if (gl == 0) my_evals += (unsigned long long)visited;
          if (stt == ST_OK) {
            chosen = t;
            code = B200S_FIT_NONE;
#pragma unroll
            for (int i = 0; i < WPL; i++) keep[i] = out[i];
          } else if (stt == ST_TAINT) {
            code = B200S_FIT_TAINT;
          } else if (stt == ST_INCOMPATIBLE) {
            code = B200S_FIT_INCOMPATIBLE;
          }
        }
        if (gl == 0) code_out[(size_t)p * T + t] = code;
      }
      if (gl == 0) tmpl_out[p] = chosen;
#pragma unroll
      for (int i = 0; i < WPL; i++) {
        int w = gl + i * G;
        if (w < W) mask_out[(size_t)p * W + w] = keep[i];
      }
    }
    __syncthreads();
  }
  // one atomic per warp
  for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
  if ((threadIdx.x & 31) == 0 && my_evals) atomicAdd(evals, my_evals);
}

// -------------------------------------------------------------------------------------------------
// FFD order: bitonic sort of pod indices.  less(a,b) is queue.go:72-108; indices >= P are +infinity.
// `seg` (may be null) is a leading ascending key: the problem a pod belongs to in a batched solve.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ffd_less(const PodsDev& pods, const int32_t* __restrict__ seg, int a, int b) {
  const int P = pods.P;
  if (a >= P) return false;  // padding sorts last
  if (b >= P) return true;
  if (seg) {
    int sa = seg[a], sb = seg[b];
    if (sa != sb) return sa < sb;
  }
  int64_t ca = pods.req[(size_t)pods.res_cpu * P + a], cb = pods.req[(size_t)pods.res_cpu * P + b];
  if (ca != cb) return ca > cb;
  int64_t ma = pods.req[(size_t)pods.res_mem * P + a], mb = pods.req[(size_t)pods.res_mem * P + b];
  if (ma != mb) return ma > mb;
  int64_t ta = pods.ts[a], tb = pods.ts[b];
  if (ta != tb) return ta < tb;
  uint64_t ha = pods.uid_hi[a], hb = pods.uid_hi[b];
  if (ha != hb) return ha < hb;
  uint64_t la = pods.uid_lo[a], lb = pods.uid_lo[b];
  if (la != lb) return la < lb;
  return a < b;  // identical UIDs cannot occur in a cluster; keeps the order total anyway
}

constexpr int SORT_THREADS = 512;
constexpr int SORT_CHUNK = 2 * SORT_THREADS;  // elements a CTA sorts in shared memory

__global__ void ffd_init_kernel(int32_t* __restrict__ idx, int n_pad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) idx[i] = i;
}

// all bitonic stages with stride < SORT_CHUNK for merge size k (k_lo..k_hi doubling) inside shared memory
__global__ void __launch_bounds__(SORT_THREADS) ffd_local_kernel(PodsDev pods, const int32_t* __restrict__ seg,
                                                                 int32_t* __restrict__ idx, int k_lo, int k_hi) {
  __shared__ int32_t s[SORT_CHUNK];
  const int base = blockIdx.x * SORT_CHUNK;
  s[threadIdx.x] = idx[base + threadIdx.x];
  s[threadIdx.x + SORT_THREADS] = idx[base + threadIdx.x + SORT_THREADS];
  __syncthreads();
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    for (int j = min(k >> 1, SORT_CHUNK >> 1); j > 0; j >>= 1) {
      int t = threadIdx.x;
      int i = ((t / j) * 2 * j) + (t % j);  // lower element of the pair
      int l = i + j;
      bool up = (((base + i) & k) == 0);
      int a = s[i], b = s[l];
      bool sw = up ? ffd_less(pods, seg, b, a) : ffd_less(pods, seg, a, b);
      if (sw) { s[i] = b; s[l] = a; }
      __syncthreads();
    }
  }
  idx[base + threadIdx.x] = s[threadIdx.x];
  idx[base + threadIdx.x + SORT_THREADS] = s[threadIdx.x + SORT_THREADS];
}

// one global bitonic stage (stride j >= SORT_CHUNK) of merge size k
__global__ void ffd_global_kernel(PodsDev pods, const int32_t* __restrict__ seg, int32_t* __restrict__ idx, int n_pad,
                                  int j, int k) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad / 2) return;
  int i = ((t / j) * 2 * j) + (t % j);
  int l = i + j;
  bool up = ((i & k) == 0);
  int a = idx[i], b = idx[l];
  bool sw = up ? ffd_less(pods, seg, b, a) : ffd_less(pods, seg, a, b);
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
                                      int32_t* alive, cudaStream_t stream) {
  if (kt.n_words > 128) return cudaErrorInvalidValue;
#define CALL(G, WPL) template_prefilter_kernel<G, WPL><<<1, 32 * MAXT, 0, stream>>>(kt, tg, type_admit, tmask, alive)
  DISPATCH_GROUP(kt.n_words, CALL);
#undef CALL
  return cudaGetLastError();
}

int fit_tables_fit_in_smem(const KeyTab& kt, int T, int smem_limit) {
  int G, WPL;
  pick_group(kt.n_words, &G, &WPL);
  FitSmemLayout L = fit_layout(G, kt.n_res, kt.n_keys, kt.n_bounded, T, kt.n_words, kt.n_rows, kt.n_dv, 1);
  return L.total <= (size_t)smem_limit;
}

cudaError_t launch_fit_matrix(const KeyTab& kt, const TmplDev* tg, int T, const uint64_t* tmask, const int32_t* alive,
                              const PodsDev& pods, int32_t* tmpl_out, uint64_t* mask_out, uint8_t* code_out,
                              unsigned long long* evals, int n_sm, int smem_limit, cudaStream_t stream) {
  if (kt.n_words > 128) return cudaErrorInvalidValue;
  if (pods.P == 0) return cudaSuccess;
  int G, WPL;
  pick_group(kt.n_words, &G, &WPL);
  int in_smem = fit_tables_fit_in_smem(kt, T, smem_limit);
  FitSmemLayout L = fit_layout(G, kt.n_res, kt.n_keys, kt.n_bounded, T, kt.n_words, kt.n_rows, kt.n_dv, in_smem);
  int n_tiles = (pods.P + FIT_TILE - 1) / FIT_TILE;
  cudaError_t err = cudaSuccess;
#define CALL(G, WPL)                                                                                              \
  do {                                                                                                            \
    auto kern = fit_matrix_kernel<G, WPL>;                                                                        \
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);                  \
    if (err != cudaSuccess) return err;                                                                           \
    int per_sm = 1;                                                                                               \
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FIT_THREADS, L.total);                     \
    if (err != cudaSuccess) return err;                                                                           \
    if (per_sm < 1) per_sm = 1;                                                                                   \
    int grid = n_tiles < n_sm * per_sm ? n_tiles : n_sm * per_sm;                                                 \
    kern<<<grid, FIT_THREADS, L.total, stream>>>(kt, tg, tmask, alive, pods, tmpl_out, mask_out, code_out, evals, \
                                                 in_smem);                                                        \
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

cudaError_t launch_ffd_sort(const PodsDev& pods, const int32_t* seg, int32_t* idx, int* launches, cudaStream_t stream) {
  const int n_pad = ffd_padded(pods.P);
  int nl = 0;
  ffd_init_kernel<<<(n_pad + 255) / 256, 256, 0, stream>>>(idx, n_pad);
  nl++;
  // sort every chunk in shared memory (merge sizes 2 .. SORT_CHUNK)
  ffd_local_kernel<<<n_pad / SORT_CHUNK, SORT_THREADS, 0, stream>>>(pods, seg, idx, 2, SORT_CHUNK);
  nl++;
  for (int k = SORT_CHUNK * 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j >= SORT_CHUNK; j >>= 1) {
      ffd_global_kernel<<<(n_pad / 2 + 255) / 256, 256, 0, stream>>>(pods, seg, idx, n_pad, j, k);
      nl++;
    }
    ffd_local_kernel<<<n_pad / SORT_CHUNK, SORT_THREADS, 0, stream>>>(pods, seg, idx, k, k);
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

}  // namespace b200

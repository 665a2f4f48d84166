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

#include "device_types.cuh"
#include "kernels.h"

namespace b200 {

constexpr int PACK_THREADS = 512;
constexpr int ORD_SMEM_CAP = 8192;  // NodeClaims whose order / count-by-position arrays stay in shared memory
constexpr int KIND_INC = 1, KIND_APPEND = 2;
constexpr int MODE_NONE = 0, MODE_ROT_INC = 1, MODE_ROT_APPEND = 2;

struct PackShared {
  TmplDev tm;
  int64_t p_req[MAXR];
  uint64_t p_mask[MAXK];
  int64_t zero_b[MAXB];
  int32_t p_gt[MAXB], p_lt[MAXB];
  uint64_t p_flags, p_tol;
  int32_t pod, best, n_claims, head, len, seq, dirty_pos, dirty_kind, mode, rot_from, rot_to, rot_x, rot_k, ord_global,
      n_pushed;
  unsigned long long evals;
};

__host__ __device__ inline size_t pack_align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t pack_smem_bytes(int G, int W, int n_rows, int n_dv, int tables_in_smem, size_t* scratch_off,
                                                  size_t* tables_off, size_t* ord_off = nullptr) {
  size_t o = pack_align16(sizeof(PackShared));
  if (scratch_off) *scratch_off = o;
  o += pack_align16(sizeof(EvalScratch) * (PACK_THREADS / G));
  if (ord_off) *ord_off = o;
  o += pack_align16((size_t)ORD_SMEM_CAP * 8);  // ord[] + key[]
  if (tables_off) *tables_off = o;
  if (tables_in_smem) o += pack_align16((size_t)n_rows * W * 8 + (size_t)n_dv * 8);
  return o;
}

// ExistingNode.CanAdd (existingnode.go:70-107) for one node by one thread: taints, Fits(pod, remaining),
// Compatible(node labels, pod requirements) with no allow-set.
__device__ __forceinline__ bool node_can_add(const KeyTab& kt, const PackArrays& A, int row, const PodRef& pd,
                                             const int64_t* zero_b) {
  if (A.n_taints[row] & ~pd.tol) return false;
  for (int r = 0; r < kt.n_res; r++) {
    int64_t avail = A.n_avail[(size_t)r * A.Etot + row];
    if (avail < 0 || pd.req[r] > avail) return false;  // resources.go:234-243
  }
  StateRef st{A.n_rmask + row, (size_t)A.Etot, A.n_rflags[row], zero_b, zero_b, 1, nullptr, 0, 0};
  for (int k = 0; k < kt.n_keys; k++)
    if (!compatible_key(kt, k, state_req(kt, st, k), pod_req(kt, pd, k), false)) return false;
  return true;
}

template <int G, int WPL>
__global__ void __launch_bounds__(PACK_THREADS, 1) pack_kernel(KeyTab kt, const TmplDev* __restrict__ tg, PodsDev pods,
                                                               PackArrays A, PackProblem* __restrict__ problems,
                                                               int tables_in_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NG = PACK_THREADS / G;
  PackShared& S = *reinterpret_cast<PackShared*>(smem);
  size_t scratch_off, tables_off, ord_off;
  pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, tables_in_smem, &scratch_off, &tables_off, &ord_off);
  EvalScratch* scratch = reinterpret_cast<EvalScratch*>(smem + scratch_off);
  int32_t* s_ord = reinterpret_cast<int32_t*>(smem + ord_off);
  int32_t* s_key = s_ord + ORD_SMEM_CAP;
  const int tid = threadIdx.x, g = tid / G, gl = tid & (G - 1);
  const int R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded, W = kt.n_words, P = pods.P;

  PackProblem& pb = problems[blockIdx.x];
  const int base = pb.pod_begin, n_pods = pb.pod_end - pb.pod_begin;
  const int E0 = pb.node_begin, E = pb.node_end - pb.node_begin;
  // this problem's slices
  uint64_t* c_mask = A.c_mask + (size_t)base * W;
  int64_t* c_req = A.c_req + (size_t)base * R;
  uint64_t* c_rmask = A.c_rmask + (size_t)base * K;
  uint64_t* c_rflags = A.c_rflags + base;
  int64_t* c_gt = A.c_gt + (size_t)base * NB;
  int64_t* c_lt = A.c_lt + (size_t)base * NB;
  int32_t* c_tmpl = A.c_tmpl + base;
  int32_t* c_npods = A.c_npods + base;
  int32_t* g_ord = A.ord + base;      // the order array's home (and its spill area past ORD_SMEM_CAP claims)
  int32_t* g_key = A.ord_key + base;
  int32_t* queue = A.queue + base;

  {
    uint64_t* d = reinterpret_cast<uint64_t*>(&S.tm);
    const uint64_t* s = reinterpret_cast<const uint64_t*>(tg);
    for (int i = tid; i < (int)(sizeof(TmplDev) / 8); i += PACK_THREADS) d[i] = s[i];
  }
  if (tables_in_smem) {
    size_t row_words = (size_t)kt.n_rows * W, dv_words = (size_t)kt.n_dv;
    uint64_t* d = reinterpret_cast<uint64_t*>(smem + tables_off);
    for (size_t i = tid; i < row_words; i += PACK_THREADS) d[i] = kt.rows[i];
    for (size_t i = tid; i < dv_words; i += PACK_THREADS) d[row_words + i] = (uint64_t)kt.res_dv[i];
    kt.rows = d;
    kt.res_dv = reinterpret_cast<const int64_t*>(d + row_words);
  }
  if (tid < MAXB) S.zero_b[tid] = 0;
  if (tid == 0) {
    S.n_claims = 0; S.head = 0; S.len = n_pods; S.seq = 0; S.dirty_pos = -1; S.dirty_kind = 0; S.evals = 0;
    S.ord_global = 0; S.n_pushed = 0;
  }
  unsigned long long my_evals = 0;
  __syncthreads();

  for (;;) {
    // ---- Queue.Pop (queue.go:46-60) ----
    if (tid == 0) {
      int p = -1;
      if (S.len > 0) {
        int cand = queue[S.head];
        if (S.n_pushed == 0 || A.lastlen[cand] != S.len) {  // lastLen is zero until a pod has been pushed back
          p = cand;
          S.head = S.head + 1 == n_pods ? 0 : S.head + 1;
          S.len--;
        }
      }
      S.pod = p;
      S.best = INT_MAX;
    }
    __syncthreads();
    const int p = S.pod;
    if (p < 0) break;
    // ---- pod row -> shared memory (updateCachedPodData's products, scheduler.go:467-486) ----
    if (tid < R) S.p_req[tid] = pods.req[(size_t)tid * P + p];
    if (tid >= 32 && tid < 32 + K) S.p_mask[tid - 32] = pods.rmask[(size_t)(tid - 32) * P + p];
    if (tid >= 64 && tid < 64 + NB) {
      S.p_gt[tid - 64] = pods.gt[(size_t)(tid - 64) * P + p];
      S.p_lt[tid - 64] = pods.lt[(size_t)(tid - 64) * P + p];
    }
    if (tid == 96) { S.p_flags = pods.rflags[p]; S.p_tol = pods.tol[p]; }
    __syncthreads();
    const PodRef pd{S.p_mask, 1, S.p_flags, S.p_gt, S.p_lt, 1, S.p_req, 1, S.p_tol};

    // ---- addToExistingNode ----
    int found = INT_MAX;
    for (int eb = 0; eb < E; eb += PACK_THREADS) {
      int e = eb + tid;
      if (e < E && node_can_add(kt, A, E0 + e, pd, S.zero_b)) atomicMin(&S.best, e);
      __syncthreads();
      found = S.best;
      __syncthreads();
      if (found != INT_MAX) break;
    }
    if (found != INT_MAX) {
      const int row = E0 + found;  // ExistingNode.Add (existingnode.go:111-119)
      if (tid < R) A.n_avail[(size_t)tid * A.Etot + row] -= S.p_req[tid];
      if (tid >= 32 && tid < 64) {
        int k = tid - 32;
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
      __syncthreads();
      continue;
    }

    // ---- sort.Slice(newNodeClaims, len(Pods) ascending) with Go's exact tie order ----
    // the order (claim ids) and the pod counts by position live in shared memory up to ORD_SMEM_CAP claims
    if (!S.ord_global && S.n_claims >= ORD_SMEM_CAP) {
      for (int i = tid; i < S.n_claims; i += PACK_THREADS) { g_ord[i] = s_ord[i]; g_key[i] = s_key[i]; }
      __syncthreads();
      if (tid == 0) S.ord_global = 1;
      __syncthreads();
    }
    int32_t* ord = S.ord_global ? g_ord : s_ord;
    int32_t* key = S.ord_global ? g_key : s_key;
    if (tid == 0) {
      int mode = MODE_NONE;
      const int n = S.n_claims;
      if (S.dirty_pos >= 0 && n > 1) {
        GoSort gs{ord, key};
        const int pos = S.dirty_pos;
        bool sorted = S.dirty_kind == KIND_INC ? (pos == n - 1 || !gs.less(pos + 1, pos)) : !gs.less(n - 1, n - 2);
        if (!sorted) {
          int hint = 0;
          if (n >= 50) gs.choose_pivot(0, n, &hint);
          if (n >= 50 && hint == 1) {
            // pdqsort takes partialInsertionSort: the single displaced element is re-inserted next to
            // its equals (see DESIGN.md "tie order"); done by all threads as a rotate
            const int cx = key[pos];
            if (S.dirty_kind == KIND_INC) {
              int lo = pos + 1, hi = n;  // first index > pos whose count >= cx
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (key[mid] < cx) lo = mid + 1; else hi = mid;
              }
              mode = MODE_ROT_INC;
              S.rot_from = pos;
              S.rot_to = lo - 1;
            } else {
              int lo = 0, hi = n - 1;  // first index whose count > cx
              while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (key[mid] <= cx) lo = mid + 1; else hi = mid;
              }
              mode = MODE_ROT_APPEND;
              S.rot_from = n - 1;
              S.rot_to = lo;
            }
            S.rot_x = ord[pos];
            S.rot_k = cx;
          } else {
            gs.sort(n);
          }
        }
      }
      S.dirty_pos = -1;
      S.mode = mode;
    }
    __syncthreads();
    if (S.mode == MODE_ROT_INC) {  // [from+1 .. to] move one slot left, X lands at `to`
      const int from = S.rot_from, to = S.rot_to;
      for (int cb = from + 1; cb <= to; cb += PACK_THREADS) {
        int i = cb + tid;
        int v = 0, kv = 0;
        if (i <= to) { v = ord[i]; kv = key[i]; }
        __syncthreads();
        if (i <= to) { ord[i - 1] = v; key[i - 1] = kv; }
        __syncthreads();
      }
      if (tid == 0) { ord[to] = S.rot_x; key[to] = S.rot_k; }
      __syncthreads();
    } else if (S.mode == MODE_ROT_APPEND) {  // [to .. from-1] move one slot right, X lands at `to`
      const int from = S.rot_from, to = S.rot_to;
      for (int ce = from - 1; ce >= to; ce -= PACK_THREADS) {
        int i = ce - tid;
        int v = 0, kv = 0;
        if (i >= to) { v = ord[i]; kv = key[i]; }
        __syncthreads();
        if (i >= to) { ord[i + 1] = v; key[i + 1] = kv; }
        __syncthreads();
      }
      if (tid == 0) { ord[to] = S.rot_x; key[to] = S.rot_k; }
      __syncthreads();
    }

    // ---- addToInflightNode ----
    const int n = S.n_claims;
    found = INT_MAX;
    for (int cb = 0; cb < n; cb += NG) {
      const int idx = cb + g;
      uint64_t out[WPL];
      int c = -1;
      if (idx < n) {
        c = ord[idx];
        const int t = c_tmpl[c];
        StateRef st{c_rmask + (size_t)c * K, 1, c_rflags[c], c_gt + (size_t)c * NB, c_lt + (size_t)c * NB, 1,
                    c_req + (size_t)c * R, 1, S.tm.taints[t]};
        uint64_t cand[WPL];
#pragma unroll
        for (int i = 0; i < WPL; i++) {
          int w = gl + i * G;
          cand[i] = w < W ? c_mask[(size_t)c * W + w] : 0ull;
        }
        uint8_t code;
        int visited;
        int stt = group_can_add<G, WPL>(kt, &scratch[g], st, pd, true, cand, out, &code, &visited);
        if (gl == 0) {
          my_evals += (unsigned long long)visited;
          if (stt == ST_OK) atomicMin(&S.best, idx);
        }
      }
      __syncthreads();
      found = S.best;
      __syncthreads();
      if (found != INT_MAX) {
        if (idx == found) {  // NodeClaim.Add (nodeclaim.go:168-180)
          StateRef st{c_rmask + (size_t)c * K, 1, c_rflags[c], c_gt + (size_t)c * NB, c_lt + (size_t)c * NB, 1,
                      c_req + (size_t)c * R, 1, 0};
          group_store_merged<G>(kt, st, pd, c_rmask + (size_t)c * K, 1, &c_rflags[c], c_gt + (size_t)c * NB,
                                c_lt + (size_t)c * NB, 1);
#pragma unroll
          for (int i = 0; i < WPL; i++) {
            int w = gl + i * G;
            if (w < W) c_mask[(size_t)c * W + w] = out[i];
          }
          for (int r = gl; r < R; r += G) c_req[(size_t)c * R + r] += S.p_req[r];
          if (gl == 0) {
            c_npods[c]++;
            key[found]++;
            A.pod_kind[p] = B200S_PLACE_CLAIM;
            A.pod_index[p] = c;
            A.pod_seq[p] = S.seq++;
            S.dirty_pos = found;
            S.dirty_kind = KIND_INC;
          }
        }
        break;
      }
    }
    if (found != INT_MAX) {
      __syncthreads();
      continue;
    }

    // ---- addToNewNodeClaim: the template walk was done for every pod by stage A ----
    const int t = A.a_tmpl[p];
    if (t >= 0) {
      const int c = n;  // NewNodeClaim + Add (nodeclaim.go:83-109, :168-180)
      if (g == 0) {
        StateRef st{S.tm.rmask[t], 1, S.tm.rflags[t], S.tm.gt[t], S.tm.lt[t], 1, S.tm.daemon[t], 1, 0};
        group_store_merged<G>(kt, st, pd, c_rmask + (size_t)c * K, 1, &c_rflags[c], c_gt + (size_t)c * NB,
                              c_lt + (size_t)c * NB, 1);
        for (int w = gl; w < W; w += G) c_mask[(size_t)c * W + w] = A.a_mask[(size_t)p * W + w];
        for (int r = gl; r < R; r += G) c_req[(size_t)c * R + r] = S.tm.daemon[t][r] + S.p_req[r];
        if (gl == 0) {
          c_tmpl[c] = t;
          c_npods[c] = 1;
          ord[n] = c;
          key[n] = 1;
          S.n_claims = n + 1;
          S.dirty_pos = n;
          S.dirty_kind = KIND_APPEND;
          A.pod_kind[p] = B200S_PLACE_CLAIM;
          A.pod_index[p] = c;
          A.pod_seq[p] = S.seq++;
        }
      }
    } else if (tid == 0) {  // Queue.Push (queue.go:63-66); the pod stays in PodErrors
      int tail = S.head + S.len;
      if (tail >= n_pods) tail -= n_pods;
      queue[tail] = p;
      S.len++;
      S.n_pushed++;
      A.lastlen[p] = S.len;
      A.pod_kind[p] = B200S_PLACE_UNSCHEDULED;
      A.pod_index[p] = -1;
    }
    __syncthreads();
  }

  if (!S.ord_global)
    for (int i = tid; i < S.n_claims; i += PACK_THREADS) g_ord[i] = s_ord[i];
  for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(0xffffffffu, my_evals, o);
  if ((tid & 31) == 0 && my_evals) atomicAdd(&S.evals, my_evals);
  __syncthreads();
  if (tid == 0) {
    pb.n_claims = S.n_claims;
    pb.n_unscheduled = S.len;
    pb.evals = S.evals;
  }
}

// Results.NewNodeClaims in final order + FinalizeScheduling (nodeclaim.go:252-268: only the placeholder
// hostname is dropped, which this encoding never materialises) + pod -> claim position.
__global__ void __launch_bounds__(256) pack_finalize_kernel(KeyTab kt, PodsDev pods, PackArrays A,
                                                            const PackProblem* __restrict__ problems, ClaimOut out) {
  const PackProblem& pb = problems[blockIdx.x];
  const int base = pb.pod_begin, n_pods = pb.pod_end - pb.pod_begin, n = pb.n_claims;
  const int R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded, W = kt.n_words;
  const int row_words = W + R + K + 2 * NB;
  for (int i = threadIdx.x; i < n * row_words; i += blockDim.x) {
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
    out.tmpl[base + pos] = A.c_tmpl[base + c];
    out.npods[base + pos] = A.c_npods[base + c];
    out.rflags[base + pos] = A.c_rflags[base + c];
    out.pos_of_claim[base + c] = pos;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_pods; i += blockDim.x) {
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
                        PackProblem* problems, int n_problems, int smem_limit, cudaStream_t stream) {
  (void)T;
  if (kt.n_words > 128) return cudaErrorInvalidValue;
  if (n_problems == 0) return cudaSuccess;
  int G, WPL;
  pick_group_pack(kt.n_words, &G, &WPL);
  size_t so, to;
  int in_smem = pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, 1, &so, &to) <= (size_t)smem_limit;
  size_t bytes = pack_smem_bytes(G, kt.n_words, kt.n_rows, kt.n_dv, in_smem, &so, &to);
  cudaError_t err = cudaSuccess;
#define CALL(G_, WPL_)                                                                            \
  do {                                                                                            \
    auto kern = pack_kernel<G_, WPL_>;                                                            \
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
  return cudaGetLastError();
}

cudaError_t launch_pack_finalize(const KeyTab& kt, const PodsDev& pods, PackArrays arr, const PackProblem* problems,
                                 int n_problems, ClaimOut out, cudaStream_t stream) {
  if (n_problems == 0) return cudaSuccess;
  pack_finalize_kernel<<<n_problems, 256, 0, stream>>>(kt, pods, arr, problems, out);
  return cudaGetLastError();
}

}  // namespace b200

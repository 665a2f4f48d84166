// kernels.h — host-callable launchers of the sm_100a kernels (fit_kernels.cu, pack_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "device_types.cuh"

namespace b200 {

// One Solve inside a (possibly batched) launch.  Pods of problem b are the entries
// [pod_begin, pod_end) of the FFD-sorted index array (problem id is the leading sort key).
struct PackProblem {
  int32_t pod_begin, pod_end;
  int32_t node_begin, node_end;
  // outputs
  int32_t n_claims, n_unscheduled;
  unsigned long long evals;
  // SM cycles thread 0 spent per phase (pop, pod row, nodes, sort, pre-filter, evaluate, commit, steps)
  unsigned long long prof[12];
  int32_t status, _pad;  // 1: the register-resident pack (pack_reg_kernels.cu) finished this problem; pack_kernel skips it
};

// Mutable state of all problems of a launch (device pointers).
struct PackArrays {
  // in-flight NodeClaims; problem b owns rows [pod_begin, pod_end) (a claim holds at least one pod)
  uint64_t* c_mask;    // [Ptot*W]  InstanceTypeOptions
  int64_t* c_req;      // [Ptot*R]  Spec.Resources.Requests
  int64_t* c_hi;       // [Ptot*R]  upper bound of Allocatable over the claim's instance types (pre-filter only)
  int64_t* c_thr;      // [Ptot*R]  smallest distinct Allocatable value >= the claim's requests: while requests + pod
                       //           stay at or below it the resource's table row, hence the claim's types, cannot change
  uint64_t* c_rmask;   // [Ptot*K]  Requirements
  uint64_t* c_rflags;  // [Ptot]
  int64_t* c_gt;       // [Ptot*NB]
  int64_t* c_lt;       // [Ptot*NB]
  int32_t* c_tmpl;     // [Ptot]
  int32_t* c_npods;    // [Ptot]    len(Pods)
  int32_t* c_nopt;     // [Ptot]    len(InstanceTypeOptions)
  int32_t* ord;        // [Ptot]    s.newNodeClaims order, local claim ids
  int32_t* ord_key;    // [Ptot]    len(Pods) by position (spill area of the shared-memory copy)
  int32_t* ord_snap;   // [Ptot]    order before the last full pdqsort (so the final normalisation can be undone)
  // existing nodes (mutable copies), resource/key-major with stride Etot
  int32_t Etot;
  int64_t* n_avail;    // [R*Etot]
  uint64_t* n_rmask;   // [K*Etot]
  uint64_t* n_rflags;  // [Etot]
  const uint64_t* n_taints;
  int32_t* n_npods;    // [Etot]
  // queue (the FFD-sorted index array is used in place as the ring buffer)
  int32_t* queue;      // [Ppad]
  int32_t* lastlen;    // [Ptot]
  // stage A results
  const int32_t* a_tmpl;   // [Ptot]
  const uint64_t* a_mask;  // [Ptot*W]; null when the rows are not materialised: the row of pod p on template t is then
                           // sgCO[sig(p), t] & rqF[rq(p), t] (stage_a_word)
  const uint64_t* sgCO;    // [(NS*T)*W] class rows C & O
  const uint64_t* rqF;     // [(NQ*T)*W] class rows F
  int32_t T;
  // NodePool limits: the new-claim template walk happens in the pack (the limit-filtered candidate set depends on the
  // claims opened so far), so it needs what the dense stage reads
  const uint8_t* sg_status;  // [NS*T]
  const uint64_t* sgC;       // [(NS*T)*W]
  const uint64_t* sgO;       // [(NS*T)*W]
  const uint64_t* tmask;     // [T*W] templates' pre-filtered instance types
  uint8_t* a_code;           // [Ptot*T] per-template attempt codes (rewritten by a limited walk)
  uint64_t* limit_rows;      // [n_problems*T*W] filterByRemainingResources rows, null without limits
  // per pod outputs
  int32_t* pod_kind;   // [Ptot]
  int32_t* pod_index;  // [Ptot] local claim id / node row
  int32_t* pod_seq;    // [Ptot]
};

// word w of pod p's stage-A row on template t (the instance types a fresh NodeClaim of template t keeps for the pod)
__device__ __forceinline__ uint64_t stage_a_word(const PackArrays& A, const PodsDev& pods, int p, int t, int W, int w) {
  if (A.a_mask) return A.a_mask[(size_t)p * W + w];
  return A.sgCO[((size_t)pods.sig_id[p] * A.T + t) * W + w] & A.rqF[((size_t)pods.req_id[p] * A.T + t) * W + w];
}

// final, ordered NodeClaim rows (what b200s_result_t carries)
struct ClaimOut {
  int32_t* tmpl;
  int32_t* npods;
  int64_t* req;      // [*R]
  uint64_t* mask;    // [*W]
  uint64_t* rmask;   // [*K]
  uint64_t* rflags;
  int64_t* gt;       // [*NB]
  int64_t* lt;
  int32_t* pos_of_claim;  // [Ptot] position in Results.NewNodeClaims of local claim id
};

// consolidation_kernel's arguments: DEVICE pointers, row i = simulation i (b200s_consolidation_sims_t / _cmds_t)
struct ConsolidationArgs {
  int n, spot_to_spot, ct_spot, ct_on_demand, ct_reserved;
  const int32_t *n_unscheduled, *n_claims, *n_candidates;
  const uint8_t* all_spot;
  const uint64_t* candidate_price;  // float64 bits
  const uint64_t *claim_mask, *rmask, *rflags;
  const int64_t *rgt, *rlt;
  int32_t *decision, *reason, *n_options, *head;
  uint64_t* mask;
  uint8_t *spot_required, *price_tie;
};
cudaError_t launch_consolidation(const KeyTab& kt, const ConsolidationArgs& a, int smem_limit, cudaStream_t stream);
// whole OrderByPrice order per set, claim-major rows (rmask [n*K], rgt/rlt [n*NB]); order_out [n*max_types]
cudaError_t launch_order_by_price(const KeyTab& kt, int n, int max_types, const uint64_t* masks, const uint64_t* rmask,
                                  const uint64_t* rflags, const int64_t* rgt, const int64_t* rlt, int32_t* order_out,
                                  int32_t* count_out, uint8_t* tie_out, int smem_limit, cudaStream_t stream);

cudaError_t launch_template_prefilter(const KeyTab& kt, const TmplDev* tg, const uint64_t* type_admit, uint64_t* tmask,
                                      int32_t* alive, int32_t* tm_count, cudaStream_t stream);
cudaError_t launch_class_tables(const KeyTab& kt, const TmplDev* tg, int T, const uint64_t* tmask, const int32_t* alive,
                                const PodsDev& pods, uint8_t* sg_status, uint64_t* sgC, uint64_t* sgO, uint64_t* sgCO,
                                uint64_t* rqF, int32_t* tm_count, int n_sm, int smem_limit, cudaStream_t stream);
cudaError_t launch_fit_dense(const KeyTab& kt, int T, const PodsDev& pods, const uint8_t* sg_status, const uint64_t* sgC,
                             const uint64_t* sgO, const uint64_t* sgCO, const uint64_t* rqF, const int32_t* tm_count,
                             int32_t* tmpl_out,
                             uint64_t* mask_out, uint8_t* code_out, unsigned long long* evals, int n_sm,
                             cudaStream_t stream);
int compact_row_bytes(int T);
cudaError_t launch_compact_export(int P, int T, const int32_t* tmpl, const uint8_t* code, uint8_t* out, int n_sm, cudaStream_t stream);
cudaError_t launch_compact_export_peers(int P, int T, const int32_t* tmpl, const uint8_t* code, int n_peers, int me,
                                        void* const* peer_out, void* const* peer_flags, long long row_off, unsigned epoch,
                                        int* timed_out, int n_sm, cudaStream_t stream);
cudaError_t launch_compact_import(int P, int T, const uint8_t* in, int32_t* tmpl, uint8_t* code, int n_sm, cudaStream_t stream);
int ffd_padded(int P);
// b200s_solve_batch_shared: every problem's ExistingNodes = a copy of the shared rows, minus the rows it removes
cudaError_t launch_replicate_nodes(int R, int K, int E0, long long Etot, const int64_t* s_avail, const uint64_t* s_rmask,
                                   const uint64_t* s_rflags, const uint64_t* s_taints, int64_t* d_avail, uint64_t* d_rmask,
                                   uint64_t* d_rflags, uint64_t* d_taints, int n_excl, const int64_t* excl_rows, int n_sm,
                                   cudaStream_t stream);
// rank[q]: request classes that sort strictly before class q in (cpu desc, memory desc) order; O(NQ^2 / threads)
cudaError_t launch_ffd_class_rank(const PodsDev& pods, int32_t* rank, cudaStream_t stream);
cudaError_t launch_ffd_sort(const PodsDev& pods, const int32_t* seg, const int32_t* rank, uint64_t* k1, int32_t* idx,
                            int* launches, cudaStream_t stream);
cudaError_t launch_select_cheapest(const KeyTab& kt, int n, const uint64_t* masks, const uint64_t* rmask, size_t rms,
                                   size_t rks, const uint64_t* rflags, const int64_t* rgt, const int64_t* rlt, size_t bms,
                                   size_t bks, int32_t* sku_out, uint64_t* price_out, int32_t* nopt_out,
                                   const PackProblem* problems, const int32_t* row_problem, cudaStream_t stream);
cudaError_t launch_pack(const KeyTab& kt, const TmplDev* tg, int T, const PodsDev& pods, PackArrays arr,
                        PackProblem* problems, int n_problems, bool keys, bool nodes, int max_pods, bool profile,
                        int smem_limit, cudaStream_t stream);
// pack_reg_kernels.cu: the register-resident pack's plan for one resident problem set — granule per resource, table
// kind per resource slot, feasibility-table shape, and host copies of the small tables it stages
constexpr int REG_RT = 5;  // resource slots of a claim
struct RegPackPlan {
  bool eligible = false;
  int32_t R = 0, cpt = 0, nwin_form = 1, cells = 0, wpt = 0, bsh = 31;
  int32_t perm[REG_RT], lim[REG_RT], toff[REG_RT], stride[REG_RT], row0[REG_RT], nd[REG_RT], maxv[REG_RT];
  int64_t scale[REG_RT];
  std::vector<uint8_t> tab;       // direct value -> rank tables
  std::vector<int32_t> win;       // four values per window entry (form RT) / {value, rank} per bucket (form 1)
  std::vector<int32_t> daemon32;  // [MAXT][8]
};
struct RegPackDev {  // device copies
  const uint8_t* tab;
  const int32_t* win;
  const uint32_t* bits;
  const int32_t* daemon;
  const int32_t* rq;  // [REG_RT][NQ]
};
bool reg_pack_plan(const KeyTab& kt, const int64_t* h_dv, int T, const int64_t (*daemon)[MAXR], int n_problems,
                   const int64_t* const* rq, const int32_t* nq, int smem_limit, RegPackPlan* plan);
void reg_pack_scale_requests(const RegPackPlan& plan, const int64_t* rq, int nq, int32_t* out, size_t out_stride,
                             size_t out_off);
cudaError_t launch_feas_table(const KeyTab& kt, const uint64_t* tmask, int T, const RegPackPlan& plan, uint32_t* bits,
                              cudaStream_t stream);
cudaError_t launch_pack_reg(const KeyTab& kt, const TmplDev* tg, int T, const PodsDev& pods, PackArrays arr,
                            PackProblem* problems, int n_problems, int max_pods, const RegPackPlan& plan,
                            const RegPackDev& dev, bool* may_overflow, int* n_launched, cudaStream_t stream);
cudaError_t launch_pack_finalize(const KeyTab& kt, const PodsDev& pods, PackArrays arr, const PackProblem* problems,
                                 int n_problems, int max_pods, ClaimOut out, cudaStream_t stream);

}  // namespace b200

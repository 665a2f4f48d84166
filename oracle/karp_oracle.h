/* karp_oracle.h — C entry points of the CPU oracle (TEST INFRASTRUCTURE; see karp_oracle.cpp).
 * The oracle consumes the same semantic model (b200s_model_t) the product's encoder does and
 * reports results in model terms: label values as a u64 per MODEL key whose bit i is the i-th
 * value of that key in ascending value-id order. */
#ifndef KARP_ORACLE_H
#define KARP_ORACLE_H
#include <stdint.h>
#include "../include/b200sched.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_opts {
  int32_t stable_sort; /* 1: replace Go's pdqsort tie order by a stable sort (H1 sensitivity probe) */
  int32_t reserved[7];
} oracle_opts_t;

typedef struct oracle_result {
  int32_t cap_claims; /* in */
  int32_t n_claims, n_unscheduled, n_templates;
  int32_t template_pool[B200S_MAX_TEMPLATES]; /* model pool of each surviving template, weight order */
  int64_t fit_evals, prefilter_evals, canadd_claim, canadd_node, new_claim_attempts, sorts, sorts_unsorted, price_ties,
      newclaim_evals; /* part of fit_evals spent in addToNewNodeClaim's template walk */
  int32_t value_overflow; /* a key had more than 64 values: claim_req_mask is truncated */
  int32_t _pad;
  int32_t* queue_order;     /* [P] FFD order (may be NULL) */
  int32_t* node_order;      /* [E] model node index per sorted row (may be NULL) */
  int32_t* node_npods;      /* [E] (may be NULL) */
  int32_t* pod_place_kind;  /* [P] */
  int32_t* pod_place_index; /* [P] */
  int32_t* pod_seq;         /* [P] */
  uint8_t* pod_fit_code;    /* [P*n_templates] */
  int32_t* claim_template;  /* [cap] index into the surviving template list */
  int32_t* claim_npods;
  int64_t* claim_requests;  /* [cap*R] */
  uint64_t* claim_mask;     /* [cap*W] */
  int32_t* claim_n_options;
  uint64_t* claim_req_mask; /* [cap*n_keys] */
  uint8_t* claim_req_flags; /* [cap*n_keys] B200S_RF_* */
  int64_t* claim_req_gt;    /* [cap*n_keys] */
  int64_t* claim_req_lt;    /* [cap*n_keys] */
  int32_t* claim_sku;
  uint64_t* claim_price;
} oracle_result_t;

/* computeConsolidation over one simulation (K/controllers/disruption/consolidation.go:133-307); value ids are MODEL
 * value ids of the capacity-type key */
typedef struct oracle_consolidation_in {
  int32_t n_candidates, all_spot, spot_to_spot;
  int32_t val_spot, val_on_demand, val_reserved; /* val_reserved may be -1 */
  double candidate_price;
  /* multi-node probes only (multinodeconsolidation.go:139-150): apply filterOutSameInstanceType (:172-224) to a
   * replacement and report validDecision.  One row per candidate: its instance type and the cheapest offering its
   * labels admit, < 0 when there is none. */
  int32_t filter_same_type, n_cand_rows;
  const int32_t* cand_sku;
  const double* cand_price;
} oracle_consolidation_in_t;

typedef struct oracle_consolidation_out {
  int32_t decision, reason, n_options, spot_required, price_tie; /* B200S_CMD_*, B200S_WHY_* */
  int32_t valid; /* validDecision of firstNConsolidationOption, set when filter_same_type */
  int32_t head[B200S_S2S_MIN_TYPES];
  uint64_t* mask;  /* [W] */
  int32_t* order;  /* [n_skus] the replacement's InstanceTypeOptions in their final order (may be NULL) */
} oracle_consolidation_out_t;

int oracle_version(void);
int oracle_solve(const b200s_model_t* model, const oracle_opts_t* opts, oracle_result_t* out);
/* Solve, then NodeClaimTemplate.ToNodeClaim's instance-type list for every new NodeClaim
 * (nodeclaimtemplate.go:77-82): OrderByPrice(requirements) cut at max_types.  order_out [cap_claims*max_types]
 * (-1 padded), count_out / tie_out [cap_claims]; tie = two listed neighbours, or the last listed and the first cut,
 * have the same price (Go's unstable sort decides their order). */
int oracle_solve_ordered(const b200s_model_t* model, const oracle_opts_t* opts, oracle_result_t* out, int32_t max_types,
                         int32_t* order_out, int32_t* count_out, uint8_t* tie_out);
int oracle_consolidate(const b200s_model_t* model, const oracle_opts_t* opts, oracle_result_t* out,
                       const oracle_consolidation_in_t* cin, oracle_consolidation_out_t* cout);
int oracle_fit_matrix(const b200s_model_t* model, int32_t pod_begin, int32_t pod_end, int32_t* tmpl_out,
                      uint64_t* mask_out, uint8_t* code_out, int64_t* evals_out, int32_t* n_templates_out);
int oracle_go_sort_perm(const int64_t* keys, int32_t n, int32_t* perm_out);
int oracle_req_algebra(const b200s_model_t* model, int32_t a0, int32_t a1, int32_t b0, int32_t b1,
                       int32_t allow_wellknown, uint64_t* mask_out, uint8_t* flags_out, int64_t* gt_out,
                       int64_t* lt_out);
int oracle_tolerates(const b200s_model_t* model, int32_t taint0, int32_t taint1, int32_t tol0, int32_t tol1);
int oracle_create_select(const char* name, int32_t n_types, const char* const* instance_types, const double* prices,
                         int64_t storage_bytes, int32_t* chosen, int32_t* disk_gb, int32_t* is_gpu);

int oracle_pod_requests(int32_t n_res, int32_t res_pods, int32_t n_containers, const int64_t* req, const uint8_t* req_has,
                        const int64_t* lim, const uint8_t* lim_has, int32_t n_init, const int64_t* init_req,
                        const uint8_t* init_req_has, const int64_t* init_lim, const uint8_t* init_lim_has,
                        const uint8_t* init_restart_always, const int64_t* overhead, int64_t* out);
int oracle_allocatable(int32_t n_res, int32_t res_memory, const int64_t* capacity, const uint8_t* capacity_has,
                       const int64_t* kube, const int64_t* system, const int64_t* eviction, const uint8_t* is_hugepages,
                       int64_t* out);

#ifdef __cplusplus
}
#endif
#endif

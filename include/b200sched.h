/*
 * b200sched.h — C-ABI of libb200sched: the B200-native fit-and-pack core that sits
 * behind Karpenter's provisioning scheduler as vendored by Azure/gpu-provisioner.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference
 * tree, K/ = vendor/sigs.k8s.io/karpenter/pkg/):
 *
 *   b200s_encode            <- scheduling.NewScheduler's per-NodePool template build and the
 *                              data the scheduler reads out of Pods / InstanceTypes
 *                              (K/controllers/provisioning/scheduling/scheduler.go:116-182,
 *                               :467-486; K/cloudprovider/types.go:105-121,181-219)
 *   b200s_catalog_upload    <- CloudProvider.GetInstanceTypes consumer side
 *                              (K/cloudprovider/types.go:84-88; call site
 *                               K/controllers/provisioning/provisioner.go:252-271)
 *   b200s_fit_matrix        <- filterInstanceTypesByRequirements over fresh NodeClaims
 *                              (K/controllers/provisioning/scheduling/nodeclaim.go:373-441,
 *                               reached through scheduler.go:587-675 and :143-158)
 *   b200s_select_cheapest   <- InstanceTypes.OrderByPrice(reqs)[0]
 *                              (K/cloudprovider/types.go:221-240)
 *   b200s_solve             <- (*Scheduler).Solve
 *                              (K/controllers/provisioning/scheduling/scheduler.go:377-432)
 *   b200s_solve_batch       <- the per-candidate SimulateScheduling loop
 *                              (K/controllers/disruption/helpers.go:52-143,
 *                               singlenodeconsolidation.go:68-101)
 *   b200s_solve_batch_shared <- the same loop with the cluster's nodes handed over once
 *   b200s_create_select     <- instance.Provider.Create's vmSize pick
 *                              (pkg/providers/instance/instance.go:76-95, :321-369)
 *
 * Conventions: every function returns 0 or a negative
 * B200S_E* code and never aborts; the caller owns and allocates every output buffer
 * and passes capacities; the library keeps no caller pointer after return; no
 * callbacks; one CUDA stream per context; a context is not re-entrant.  Any non-zero
 * return means "fall back to the vendored CPU Solve" on the Go side.  There is NO CPU
 * backend in this library: b200s_init fails with B200S_ENODEVICE when no sm_100
 * device is usable.
 */
#ifndef B200SCHED_H
#define B200SCHED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200S_VERSION 100 /* 0.1.0 */

/* ---- limits of the exact encoding (inputs beyond them are rejected, never approximated) ---- */
#define B200S_MAX_KEYS 16        /* label keys that take part in requirements              */
#define B200S_MAX_VALUES 64      /* distinct values per label key (one u64 word)           */
#define B200S_MAX_RES 8          /* resource dimensions                                    */
#define B200S_MAX_TEMPLATES 16   /* NodePools                                              */
#define B200S_MAX_TAINTS 64      /* distinct taints over NodePools + existing nodes        */
#define B200S_MAX_BOUNDED 4      /* label keys that carry Gt/Lt bounds anywhere            */
#define B200S_MAX_ZC 64          /* zones x capacity-types                                 */

/* ---- error codes ---- */
#define B200S_OK 0
#define B200S_EINVAL (-1)           /* malformed argument                                   */
#define B200S_ENODEVICE (-2)        /* no usable sm_100 device / CUDA runtime failure at init */
#define B200S_ECUDA (-3)            /* CUDA runtime error (see b200s_last_error)             */
#define B200S_ENOTREPRESENTABLE (-4) /* input outside the exact encoding: use the CPU path   */
#define B200S_ECAPACITY (-5)        /* caller-provided output capacity too small             */
#define B200S_ENOMEM (-6)
#define B200S_ESTATE (-7)           /* unknown catalog id / context misuse                   */
#define B200S_ESTALE (-8)           /* pods / nodes were encoded against another key layout than the uploaded catalog
                                       (layout_hash differs): upload this encode's catalog and call again */

/* ---- requirement operators (corev1.NodeSelectorOperator; K/scheduling/requirement.go:47-87) ---- */
#define B200S_OP_IN 0
#define B200S_OP_NOTIN 1
#define B200S_OP_EXISTS 2
#define B200S_OP_DOESNOTEXIST 3
#define B200S_OP_GT 4
#define B200S_OP_LT 5

/* ---- taint effects / toleration operators (vendor/k8s.io/api/core/v1/toleration.go:38-57) ---- */
#define B200S_EFFECT_NONE 0
#define B200S_EFFECT_NOSCHEDULE 1
#define B200S_EFFECT_PREFERNOSCHEDULE 2
#define B200S_EFFECT_NOEXECUTE 3
#define B200S_TOLOP_EMPTY 0  /* "" means Equal */
#define B200S_TOLOP_EQUAL 1
#define B200S_TOLOP_EXISTS 2
#define B200S_TOLOP_OTHER 3  /* anything else never tolerates */

/* =====================================================================================
 * Level 1 — the semantic model.  This is what the cgo shim flattens Karpenter objects
 * into: strings are interned to small integers, everything else keeps the reference's
 * meaning (operators, sets, exact integer quantities).  The CPU oracle consumes exactly
 * this structure too, so the encoder below is itself under test.
 * ===================================================================================== */

/* one NodeSelectorRequirement: key op values (K/scheduling/requirement.go:36-43) */
typedef struct b200s_req {
  int32_t key;        /* label key id, 0..n_keys-1                                        */
  int32_t op;         /* B200S_OP_*                                                      */
  int32_t val_off;    /* In/NotIn: range in req_values (value ids)                       */
  int32_t val_cnt;
  int64_t bound;      /* Gt/Lt: the integer operand (strconv.Atoi of values[0])          */
  int32_t min_values; /* -1 = nil; anything else is rejected by the encoder              */
  int32_t _pad;
} b200s_req_t;

typedef struct b200s_taint {
  int32_t key, value; /* interned strings, 0 = "" */
  int32_t effect;     /* B200S_EFFECT_* */
} b200s_taint_t;

typedef struct b200s_toleration {
  int32_t key, value; /* interned strings, 0 = "" */
  int32_t op;         /* B200S_TOLOP_* */
  int32_t effect;     /* B200S_EFFECT_*, NONE = matches every effect */
} b200s_toleration_t;

typedef struct b200s_model {
  /* label universe */
  int32_t n_keys;
  const uint8_t* key_wellknown; /* [n_keys] 1 if in v1.WellKnownLabels (K/apis/v1/labels.go:83-92) */
  int32_t key_zone;             /* topology.kubernetes.io/zone, or -1 */
  int32_t key_capacity_type;    /* karpenter.sh/capacity-type, or -1 */
  int32_t key_instance_type;    /* node.kubernetes.io/instance-type, or -1.  Its values name SKUs, so it is
                                   exempt from the 64-values-per-key limit: every SKU must carry
                                   `In {its own name}`, NodePools may constrain it, pods may not. */
  int32_t n_vals;
  const int32_t* val_key;       /* [n_vals] key each interned value belongs to */
  const int64_t* val_int;       /* [n_vals] strconv.Atoi(value) when val_isint */
  const uint8_t* val_isint;     /* [n_vals] */

  /* resources: exact fixed-scale int64 (cpu in milli-units, the rest in base units) */
  int32_t n_res;
  int32_t res_cpu, res_memory;  /* indices used by the FFD order (queue.go:72-108) */

  /* requirement / taint / toleration pools */
  int32_t n_reqs;
  const b200s_req_t* reqs;
  int32_t n_req_values;
  const int32_t* req_values;
  int32_t n_taints;
  const b200s_taint_t* taints;
  int32_t n_tolerations;
  const b200s_toleration_t* tolerations;

  /* instance types (the GetInstanceTypes rows; K/cloudprovider/types.go:105-121) */
  int32_t n_skus;
  const int32_t* sku_req_off;   /* [n_skus+1] into reqs                                   */
  const int64_t* sku_alloc;     /* [n_skus*n_res] Allocatable() (types.go:181-219)        */
  const int64_t* sku_capacity;  /* [n_skus*n_res] Capacity (pool limits, scheduler.go:851) */
  const int32_t* sku_off_off;   /* [n_skus+1] into offerings                              */
  int32_t n_offerings;
  const int32_t* off_req_off;   /* [n_offerings+1] into reqs                              */
  const double* off_price;      /* [n_offerings]                                          */
  const uint8_t* off_available; /* [n_offerings]                                          */

  /* NodePools, unsorted: OrderByWeight (K/utils/nodepool/nodepool.go:121-131) is part of the path */
  int32_t n_pools;
  const int32_t* pool_req_off;    /* [n_pools+1] spec requirements + labels as In reqs     */
  const int32_t* pool_taint_off;  /* [n_pools+1] into taints                               */
  const int32_t* pool_weight;     /* [n_pools]                                             */
  const int32_t* pool_name_rank;  /* [n_pools] rank of the pool name in ascending string order */
  const int64_t* pool_daemon;     /* [n_pools*n_res] daemon overhead (scheduler.go:773-783) */
  const uint8_t* pool_has_limit;  /* [n_pools*n_res] limits present per resource           */
  const int64_t* pool_limit;      /* [n_pools*n_res]                                       */

  /* pending pods */
  int32_t n_pods;
  const int32_t* pod_req_off;     /* [n_pods+1] NewPodRequirements (requirements.go:74-110) */
  const int32_t* pod_tol_off;     /* [n_pods+1] into tolerations                           */
  const int64_t* pod_requests;    /* [n_pods*n_res] RequestsForPods incl. pods:1 (resources.go:30-38) */
  const int64_t* pod_ts;          /* [n_pods] creationTimestamp, seconds                   */
  const uint64_t* pod_uid_hi;     /* [n_pods] UID as a 128-bit big-endian integer          */
  const uint64_t* pod_uid_lo;

  /* existing nodes, unsorted: sortExistingNodes (scheduler.go:729-742) is part of the path */
  int32_t n_nodes;
  const int32_t* node_req_off;    /* [n_nodes+1] node labels as In reqs                    */
  const int32_t* node_taint_off;  /* [n_nodes+1]                                           */
  const int64_t* node_available;  /* [n_nodes*n_res] remainingResources (existingnode.go:59) */
  const uint8_t* node_initialized;/* [n_nodes]                                             */
  const int32_t* node_name_rank;  /* [n_nodes]                                             */
} b200s_model_t;

/* =====================================================================================
 * Level 0 — the device-facing structure-of-arrays encoding.  Label sets are one u64
 * word per key, taints one u64 word, SKU sets ceil(S/64) words.  All pointers are HOST
 * pointers; b200s_catalog_upload / b200s_fit_matrix / b200s_solve copy them to HBM.
 * ===================================================================================== */

/* per-key flag nibble inside a flags word: key k uses bits [4k, 4k+4) */
#define B200S_RF_DEFINED 1u
#define B200S_RF_COMPLEMENT 2u
#define B200S_RF_HASGT 4u
#define B200S_RF_HASLT 8u

typedef struct b200s_catalog_soa {
  int32_t n_skus;   /* S */
  int32_t n_words;  /* W = ceil(S/64) */
  int32_t n_res;    /* R */
  int32_t n_keys;   /* K device key slots */
  int32_t n_zones, n_cts, key_zone, key_ct; /* device key slots of zone / capacity-type, -1 if absent */
  uint32_t wellknown;                 /* bit k: device key k is a well-known label             */
  int32_t key_nvals[B200S_MAX_KEYS];  /* size of each key's value universe                     */
  int32_t n_bounded;
  int32_t bounded_key[B200S_MAX_BOUNDED]; /* device keys that may carry Gt/Lt                 */
  /* value universe numerics: within(gt,lt) needs strconv.Atoi of the value */
  const int64_t* val_int;    /* [K*64] */
  const uint64_t* val_isint; /* [K] bit v: value v of key k parses as an integer */
  /* bit-row tables, each row W words over SKUs */
  int32_t n_rows;
  const uint64_t* rows;      /* [n_rows*W] */
  int32_t row_nonneg;        /* SKUs whose allocatable has no negative entry (resources.go:234-238) */
  int32_t res_nd[B200S_MAX_RES];    /* distinct allocatable values per resource               */
  int32_t res_dv0[B200S_MAX_RES];   /* offset of resource r in res_dv                         */
  int32_t res_row0[B200S_MAX_RES];  /* row of "alloc[r] >= dv[r][0]"; row0+i for dv[r][i]     */
  const int64_t* res_dv;     /* ascending distinct allocatable values, concatenated          */
  int32_t n_dv;
  int32_t key_row0[B200S_MAX_KEYS]; /* row of admit[k][0]; +v; then undef[k] at +nvals, dne[k] at +nvals+1 */
  int32_t zc_row0;           /* row of offering-available[z*n_cts+c]                         */
  const uint64_t* price;     /* [n_zones*n_cts*S] cheapest available offering price (float64 bits) per (zone,ct,SKU), B200S_PRICE_NONE if none */
  const int64_t* alloc;      /* [R*S] raw allocatable, resource-major                        */
  const int64_t* capacity;   /* [R*S] raw capacity, resource-major                           */
  const uint64_t* price_max; /* [n_zones*n_cts*S] MOST expensive available offering per (zone,ct,SKU), B200S_PRICE_ABSENT if
                                none: Offerings.MostExpensive for WorstLaunchPrice (types.go:454-474) */
  /* NodePool limits (scheduler.go:831-867): per resource the ascending distinct Capacity values and the bit rows
   * "capacity[r] <= cap_dv[r][i]" (row cap_row0[r] + i): filterByRemainingResources is the AND of one row per limited
   * resource, subtractMax the smallest row that contains the new NodeClaim's instance types */
  int32_t cap_nd[B200S_MAX_RES], cap_dv0[B200S_MAX_RES], cap_row0[B200S_MAX_RES];
  const int64_t* cap_dv;
  int32_t n_cap_dv;
  /* The device key slots, each key's value bits, the well-known mask and the bounded-key slots are chosen by
   * b200s_encode from the keys and Gt/Lt bounds of ONE model (SKUs, offerings, pools AND that batch's pods), so a pods /
   * nodes SoA is only meaningful next to the catalog of the same encode.  layout_hash digests that choice; it is
   * stamped into all three structures and the solve entry points answer B200S_ESTALE when they differ (0 = unchecked). */
  uint64_t layout_hash;
} b200s_catalog_soa_t;

#define B200S_PRICE_NONE 0x7FEFFFFFFFFFFFFFull /* math.MaxFloat64 (types.go:224-225) */
#define B200S_PRICE_ABSENT 0xFFFFFFFFFFFFFFFFull /* price_max cell without an available offering */

typedef struct b200s_templates {
  int32_t n_templates;       /* T, already in OrderByWeight order */
  int32_t pool_index[B200S_MAX_TEMPLATES]; /* model pool each template came from */
  uint64_t taints[B200S_MAX_TEMPLATES];
  uint64_t req_mask[B200S_MAX_TEMPLATES][B200S_MAX_KEYS];
  uint64_t req_flags[B200S_MAX_TEMPLATES];
  int64_t req_gt[B200S_MAX_TEMPLATES][B200S_MAX_BOUNDED];
  int64_t req_lt[B200S_MAX_TEMPLATES][B200S_MAX_BOUNDED];
  int64_t daemon[B200S_MAX_TEMPLATES][B200S_MAX_RES];
  int64_t limit[B200S_MAX_TEMPLATES][B200S_MAX_RES];    /* s.remainingResources[pool] at the start of the Solve */
  uint8_t has_limit[B200S_MAX_TEMPLATES][B200S_MAX_RES]; /* 1: the NodePool limits this resource (scheduler.go:170-172) */
  const uint64_t* type_admit; /* [T*W] SKUs each template's instance-type requirement admits; NULL = all */
} b200s_templates_t;

/* Pods are dictionary-encoded: replicas of one workload share a requirement signature (requirement
 * set + tolerated taints) and a request vector, so a pod row is two class ids (8 bytes) plus the FFD
 * tie-break keys.  The class tables are tiny next to the pod arrays and stay L2-resident. */
typedef struct b200s_pods_soa {
  int32_t n_pods;            /* P */
  int32_t n_res, n_keys, n_bounded;
  int32_t res_cpu, res_memory; /* resource rows the FFD order sorts on (queue.go:72-108) */
  int32_t n_sigs;            /* NS distinct requirement signatures */
  int32_t n_reqcls;          /* NQ distinct request vectors        */
  /* per pod */
  const int32_t* sig_id;     /* [P] */
  const int32_t* req_id;     /* [P] */
  const int64_t* ts;         /* [P] creationTimestamp, seconds */
  const uint64_t* uid_hi;    /* [P] */
  const uint64_t* uid_lo;    /* [P] */
  /* signature table, key-major over NS */
  const uint64_t* sig_mask;  /* [K*NS] */
  const uint64_t* sig_flags; /* [NS]   */
  const int32_t* sig_gt;     /* [NB*NS] */
  const int32_t* sig_lt;     /* [NB*NS] */
  const uint64_t* sig_tolerated; /* [NS] bit t: tolerates taint t (taints.go:53-65) */
  /* request-class table, resource-major over NQ */
  const int64_t* rq_requests; /* [R*NQ] RequestsForPods incl. pods:1 (resources.go:30-38) */
  uint64_t layout_hash;      /* see b200s_catalog_soa_t */
} b200s_pods_soa_t;

typedef struct b200s_nodes_soa {
  int32_t n_nodes;           /* E, already in sortExistingNodes order */
  int32_t n_res, n_keys;
  const int32_t* node_index; /* [E] model node each row came from */
  const int64_t* available;  /* [R*E] resource-major */
  const uint64_t* req_mask;  /* [K*E] key-major (node labels: single-value In sets) */
  const uint64_t* req_flags; /* [E] */
  const uint64_t* taints;    /* [E] */
  uint64_t layout_hash;      /* see b200s_catalog_soa_t */
} b200s_nodes_soa_t;

/* ---- per-template attempt code written by fit_matrix / solve (parity is on the code, never on message text) ---- */
#define B200S_FIT_NONE 0x00          /* template not evaluated (an earlier one succeeded) or success */
#define B200S_FIT_TAINT 0x01         /* did not tolerate taint (nodeclaim.go:116)                   */
#define B200S_FIT_INCOMPATIBLE 0x02  /* incompatible requirements (nodeclaim.go:128)               */
#define B200S_FIT_LIMITS 0x03        /* all instance types exceed nodepool limits (scheduler.go:603) */
#define B200S_FIT_DROPPED 0x7F       /* template had no instance types after NewScheduler's pre-filter and
                                        does not exist for the scheduler (scheduler.go:147-156)       */
#define B200S_FIT_FILTER 0x80        /* InstanceTypeFilterError; low 6 bits are the flags below     */
#define B200S_FF_REQUIREMENTS 0x01   /* requirementsMet        (nodeclaim.go:409) */
#define B200S_FF_FITS 0x02           /* fits                   (:410)             */
#define B200S_FF_OFFERING 0x04       /* hasOffering            (:411)             */
#define B200S_FF_REQ_AND_FITS 0x08   /* requirementsAndFits    (:414)             */
#define B200S_FF_REQ_AND_OFF 0x10    /* requirementsAndOffering(:415)             */
#define B200S_FF_FITS_AND_OFF 0x20   /* fitsAndOffering        (:416)             */

/* ---- where a pod ended up ---- */
#define B200S_PLACE_UNSCHEDULED (-1)
#define B200S_PLACE_NODE 0   /* index = row of the existing node (sorted order) */
#define B200S_PLACE_CLAIM 1  /* index = position in Results.NewNodeClaims       */

typedef struct b200s_solve_opts {
  int32_t want_masks;        /* fill claim_mask */
  int32_t reserved[7];
} b200s_solve_opts_t;

typedef struct b200s_result {
  /* capacities, set by the caller */
  int32_t cap_claims;        /* rows available in the claim_* arrays */
  /* counts, set by the library */
  int32_t n_claims;          /* len(Results.NewNodeClaims) */
  int32_t n_unscheduled;     /* len(Results.PodErrors) */
  int64_t fit_evals;         /* instance types the device path actually re-filtered (informational: claims that keep
                              * their types under the threshold rule are decided without touching them) */
  /* per pod [P] */
  int32_t* pod_place_kind;   /* B200S_PLACE_* */
  int32_t* pod_place_index;
  int32_t* pod_seq;          /* order in which pods were placed (Pods append order) */
  uint8_t* pod_fit_code;     /* [P*T] last new-NodeClaim attempt per template */
  /* per new NodeClaim, in Results.NewNodeClaims order [cap_claims] */
  int32_t* claim_template;
  int32_t* claim_npods;
  int64_t* claim_requests;   /* [cap_claims*R] */
  uint64_t* claim_mask;      /* [cap_claims*W] InstanceTypeOptions; may be NULL unless want_masks */
  uint64_t* claim_req_mask;  /* [cap_claims*K] merged requirements (FinalizeScheduling applied) */
  uint64_t* claim_req_flags; /* [cap_claims] */
  int64_t* claim_req_gt;     /* [cap_claims*NB] */
  int64_t* claim_req_lt;     /* [cap_claims*NB] */
  int32_t* claim_sku;        /* OrderByPrice(reqs)[0] */
  uint64_t* claim_price;     /* its price, float64 bits */
  int32_t* claim_n_options;  /* len(InstanceTypeOptions) */
} b200s_result_t;

typedef struct b200s_ctx b200s_ctx_t;
typedef struct b200s_encoded b200s_encoded_t;

/* ---- lifecycle ---- */
int b200s_version(void);
int b200s_init(int device, b200s_ctx_t** out);
void b200s_destroy(b200s_ctx_t* ctx);
const char* b200s_last_error(const b200s_ctx_t* ctx);
/* CUDA stream the context launches on (a cudaStream_t), for callers that time with events */
void* b200s_stream(b200s_ctx_t* ctx);
int b200s_synchronize(b200s_ctx_t* ctx);

/* ---- device timing on the context's stream (CUDA events) and the L2 flush the timing rules need ---- */
void* b200s_event_create(b200s_ctx_t* ctx);
int b200s_event_record(b200s_ctx_t* ctx, void* ev);
int b200s_event_elapsed_ms(b200s_ctx_t* ctx, void* start, void* stop, float* ms);
void b200s_event_destroy(void* ev);
int b200s_flush_l2(b200s_ctx_t* ctx, size_t bytes);

/* ---- host-side encoder: semantic model -> SoA (total or reject; no device needed) ---- */
int b200s_encode(const b200s_model_t* model, b200s_encoded_t** out, char* err, size_t err_cap);
void b200s_encoded_free(b200s_encoded_t* enc);
const b200s_catalog_soa_t* b200s_encoded_catalog(const b200s_encoded_t* enc);
const b200s_templates_t* b200s_encoded_templates(const b200s_encoded_t* enc);
const b200s_pods_soa_t* b200s_encoded_pods(const b200s_encoded_t* enc);
const b200s_nodes_soa_t* b200s_encoded_nodes(const b200s_encoded_t* enc);
/* model key / value ids behind a device key slot and bit (for decoding results) */
int b200s_encoded_key_model_id(const b200s_encoded_t* enc, int device_key);
int b200s_encoded_value_model_id(const b200s_encoded_t* enc, int device_key, int bit);

/* ---- device path ---- */
int b200s_catalog_upload(b200s_ctx_t* ctx, const b200s_catalog_soa_t* cat, uint64_t* catalog_id);
int b200s_catalog_release(b200s_ctx_t* ctx, uint64_t catalog_id);

/* Dense stage A.  For every pod, walk the templates in weight order exactly as
 * addToNewNodeClaim does and record the first template whose filtered instance-type set
 * is non-empty.  Outputs (caller-allocated HOST buffers):
 *   tmpl_out  [P]    first feasible template or -1
 *   mask_out  [P*W]  surviving instance types of that template (zeros if none)
 *   code_out  [P*T]  per-template attempt code (B200S_FIT_*)
 * Instance types evaluated = sum over pods of the |InstanceTypeOptions| of every
 * template whose filter loop ran; returned through evals_out (may be NULL). */
int b200s_fit_matrix(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_pods_soa_t* pods,
                     const b200s_templates_t* templates, int32_t* tmpl_out, uint64_t* mask_out,
                     uint8_t* code_out, int64_t* evals_out);

/* argmin-price SKU of n SKU sets under per-set requirement rows (zone / capacity-type
 * constraints are read from req_mask/req_flags/req_gt/req_lt, K-major like the pod SoA).
 * sku_out[i] = -1 and price_out[i] = B200S_PRICE_NONE when set i is empty. */
int b200s_select_cheapest(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n, const uint64_t* masks,
                          const uint64_t* req_mask, const uint64_t* req_flags, const int64_t* req_gt,
                          const int64_t* req_lt, int32_t* sku_out, uint64_t* price_out);

/* The instance-type list NodeClaimTemplate.ToNodeClaim writes into a NodeClaim
 * (K/controllers/provisioning/scheduling/nodeclaimtemplate.go:77-82): InstanceTypeOptions.OrderByPrice(reqs)
 * (types.go:221-240) of each of n SKU sets, cut at max_types (MaxInstanceTypes = 600, nodeclaimtemplate.go:41).
 * Rows are claim-major as in b200s_result_t: masks [n*W], req_mask [n*K], req_flags [n], req_gt / req_lt [n*NB].
 * order_out [n*max_types] SKU indices in price order, -1 after count_out[i] entries; tie_out[i] = 1 when two listed
 * neighbours (or the last listed and the first one cut) share a price -- Go's sort is unstable there, this library
 * orders equal prices by SKU index. */
#define B200S_MAX_INSTANCE_TYPES 600
int b200s_order_by_price(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n, const uint64_t* masks,
                         const uint64_t* req_mask, const uint64_t* req_flags, const int64_t* req_gt,
                         const int64_t* req_lt, int32_t max_types, int32_t* order_out, int32_t* count_out,
                         uint8_t* tie_out);

/* Full Solve: stage A, FFD order, pack, finalize, cheapest SKU per NodeClaim. */
int b200s_solve(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_pods_soa_t* pods,
                const b200s_nodes_soa_t* existing, const b200s_templates_t* templates,
                const b200s_solve_opts_t* opts, b200s_result_t* out);

/* B independent Solves sharing one catalog and one template set (consolidation simulations).
 * Problem b uses pods[b], existing[b] (may be NULL) and fills out[b]. */
int b200s_solve_batch(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n_problems,
                      const b200s_pods_soa_t* const* pods, const b200s_nodes_soa_t* const* existing,
                      const b200s_templates_t* templates, const b200s_solve_opts_t* opts,
                      b200s_result_t* const* out);

/* The same for simulations that share one cluster — SimulateScheduling per candidate
 * (K/controllers/disruption/helpers.go:52-143) always schedules "the candidates' pods against every other node", so
 * the B problems differ only in which pods are pending and which few nodes are gone.  ONE pod set (one b200s_encode,
 * one dictionary of signatures / request classes) whose rows pod_off[b] .. pod_off[b+1] are problem b's pods
 * (pod_off[0] = 0, pod_off[n_problems] = pods->n_pods), ONE node set, of which problem b removes the rows
 * excl_rows[excl_off[b] .. excl_off[b+1]) (excl_off may be NULL: nothing removed).  The node rows are uploaded
 * once and copied per problem on the device; a removed row gets a negative Available entry, which nothing fits
 * (resources.go:234-243), so every other node keeps its place in the order.  Results are those of b200s_solve_batch
 * on the B separate problems, except that existing-node indices (pod_place_index of B200S_PLACE_NODE pods, the
 * node rows of out[b]) count rows of the SHARED node set. */
int b200s_solve_batch_shared(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n_problems, const b200s_pods_soa_t* pods,
                             const int32_t* pod_off, const b200s_nodes_soa_t* nodes, const int32_t* excl_off,
                             const int32_t* excl_rows, const b200s_templates_t* templates,
                             const b200s_solve_opts_t* opts, b200s_result_t* const* out);

/* ---- consolidation commands over a batch of simulations (SURVEY section 8f N2, config 4's consumer) ----
 * computeConsolidation (K/controllers/disruption/consolidation.go:133-307) applied to n finished simulations at
 * once, one warp per simulation: the decision on the counts of the Solve result, OrderByPrice, the price filter
 * RemoveInstanceTypeOptionsByPriceAndMinValues (nodeclaim.go:270-279; minValues is nil in this scope) with
 * Offerings.WorstLaunchPrice (types.go:460-474), the spot-to-spot rules (:230-307) and the capacity-type
 * requirement the command adds.  Inputs are HOST arrays, row i = simulation i; the claim_* rows are
 * Results.NewNodeClaims[0] as b200s_solve / b200s_solve_batch returned it and are read only when n_claims[i] == 1. */
typedef struct b200s_consolidation_sims {
  int32_t n;
  int32_t spot_to_spot;            /* the SpotToSpotConsolidation feature gate (consolidation.go:233) */
  int32_t ct_spot, ct_on_demand, ct_reserved; /* DEVICE value bits of "spot" / "on-demand" / "reserved" inside the
                                      capacity-type key (b200s_encoded_value_model_id maps bits to model values);
                                      ct_reserved may be -1 when no offering and no requirement names it */
  const int32_t* n_unscheduled;    /* [n] PodErrors left after OmitBy(IsProvisionable) (scheduler.go:326-330) */
  const int32_t* n_claims;         /* [n] len(Results.NewNodeClaims) */
  const int32_t* n_candidates;     /* [n] nodes the command would remove */
  const uint8_t* all_spot;         /* [n] every candidate is a spot node (consolidation.go:178-183) */
  const double* candidate_price;   /* [n] getCandidatePrices (consolidation.go:310-328): finite, >= 0 */
  const uint64_t* claim_mask;      /* [n*W] InstanceTypeOptions */
  const uint64_t* claim_req_mask;  /* [n*K] */
  const uint64_t* claim_req_flags; /* [n] */
  const int64_t* claim_req_gt;     /* [n*NB] */
  const int64_t* claim_req_lt;     /* [n*NB] */
} b200s_consolidation_sims_t;

#define B200S_CMD_NONE 0     /* Command{}: leave the candidates alone */
#define B200S_CMD_DELETE 1   /* the pods fit on what is already there (consolidation.go:155-160) */
#define B200S_CMD_REPLACE 2  /* one cheaper replacement NodeClaim */
#define B200S_WHY_OK 0
#define B200S_WHY_UNSCHEDULABLE 1    /* !AllNonPendingPodsScheduled (:147)                  */
#define B200S_WHY_MANY_CLAIMS 2      /* len(NewNodeClaims) != 1 (:163)                      */
#define B200S_WHY_NOT_CHEAPER 3      /* nothing left after the price filter (:205, :258)    */
#define B200S_WHY_S2S_DISABLED 4     /* spot-to-spot with the gate off (:233)               */
#define B200S_WHY_S2S_TOO_FEW 5      /* fewer than 15 cheaper spot types, one candidate (:279) */
#define B200S_S2S_MIN_TYPES 15       /* MinInstanceTypesForSpotToSpotConsolidation (consolidation.go:49) */

typedef struct b200s_consolidation_cmds {   /* caller-allocated HOST arrays */
  int32_t* decision;      /* [n] B200S_CMD_*  */
  int32_t* reason;        /* [n] B200S_WHY_*  */
  int32_t* n_options;     /* [n] len(Replacements[0].InstanceTypeOptions), 0 unless REPLACE */
  uint64_t* mask;         /* [n*W] those instance types as a set (zeros unless REPLACE) */
  uint8_t* spot_required; /* [n] 1: the command narrowed the NodeClaim's capacity-type requirement to In{spot}
                             (consolidation.go:218-221, :240) */
  int32_t* head;          /* [n*B200S_S2S_MIN_TYPES] the options in OrderByPrice order when the command truncated them to
                             the 15 cheapest (single-candidate spot-to-spot, :303), else filled with -1 */
  uint8_t* price_tie;     /* [n] 1: two of the 16 cheapest share a price, so `head` depends on Go's unstable sort */
} b200s_consolidation_cmds_t;

int b200s_consolidation_commands(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_consolidation_sims_t* sims,
                                 b200s_consolidation_cmds_t* out);

/* ---- resident (device-pointer) variants used for device-timed measurement ---- */
typedef struct b200s_resident b200s_resident_t;
int b200s_resident_create(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_pods_soa_t* pods,
                          const b200s_nodes_soa_t* existing, const b200s_templates_t* templates,
                          b200s_resident_t** out);
void b200s_resident_free(b200s_ctx_t* ctx, b200s_resident_t* r);
/* enqueue only, no host copies, no synchronisation */
int b200s_resident_fit_matrix(b200s_ctx_t* ctx, b200s_resident_t* r);
int b200s_resident_solve(b200s_ctx_t* ctx, b200s_resident_t* r);
/* copy the resident results back (synchronises) */
int b200s_resident_fetch_fit(b200s_ctx_t* ctx, b200s_resident_t* r, int32_t* tmpl_out, uint64_t* mask_out,
                             uint8_t* code_out, int64_t* evals_out);
int b200s_resident_fetch_solve(b200s_ctx_t* ctx, b200s_resident_t* r, b200s_result_t* out);
/* device-to-device copy of the resident dense results into caller-owned DEVICE buffers (e.g. the send buffers of
 * an NCCL all-gather when the pod rows are sharded across GPUs); enqueued on the context's stream */
int b200s_resident_copy_fit_device(b200s_ctx_t* ctx, b200s_resident_t* r, void* d_tmpl, void* d_mask, void* d_code);
/* Plumbing for the peer-store export below: a device allocation of the library's own (zero-initialised) whose CUDA IPC
 * handle (64 bytes, cudaIpcMemHandle_t) another process of the same node opens on ITS device with peer access enabled;
 * b200s_copy_device enqueues a device-to-device copy on the context's stream (to look at such a buffer from a framework
 * tensor).  One process per GPU exchanges the handles with whatever it has (the bench: torch.distributed). */
int b200s_device_alloc(b200s_ctx_t* ctx, uint64_t bytes, void** out);
int b200s_device_free(b200s_ctx_t* ctx, void* p);
int b200s_ipc_export(b200s_ctx_t* ctx, void* dptr, uint8_t handle[64]);
int b200s_ipc_open(b200s_ctx_t* ctx, const uint8_t handle[64], void** out);
int b200s_ipc_close(b200s_ctx_t* ctx, void* p);
int b200s_copy_device(b200s_ctx_t* ctx, void* dst, const void* src, uint64_t bytes);

/* The export fused with its all-gather, for ranks of one NVLink domain that have mapped each other's memory (CUDA IPC;
 * torch's storage sharing does it in the bench): peer_out[q] is rank q's gathered buffer [P_total * row_bytes] and
 * peer_flags[q] rank q's flag array (uint32[n_peers], zero-initialised) AS MAPPED INTO THIS PROCESS (q == me: own
 * memory).  The shard's rows are stored at row_offset in every peer's buffer, then the ranks meet at a flag barrier:
 * rank `me` publishes `epoch` (a counter the caller increments every call, starting at 1) in slot `me` of every flag
 * array and waits for `epoch` in every slot of its own.  When the call's stream work completes, this rank's buffer holds
 * every rank's rows of this epoch.  The wait is bounded (about a second): *d_timed_out (device int, caller-zeroed) is
 * set instead of hanging when a peer never arrives.  The caller must not start epoch e+1 while a peer may still read
 * epoch e's rows (the bench has a step barrier).  n_peers <= 8. */
int b200s_resident_export_compact_peers(b200s_ctx_t* ctx, b200s_resident_t* r, int32_t n_peers, int32_t me,
                                        void* const* peer_out, void* const* peer_flags, int64_t row_offset,
                                        uint32_t epoch, int32_t* d_timed_out);

/* ---- pod rows sharded across GPUs ----
 * Each GPU evaluates its block of pod rows (b200s_resident_fit_matrix on a resident built from that block); what the
 * GPU that runs the pack needs of a pod is its first feasible template and the per-template attempt codes — the
 * instance-type row is a function of (signature, request class, template) and of the class tables every GPU builds from
 * the replicated catalog, so it is never sent.  export writes the compact rows of this resident's pods to DEVICE memory
 * (the send buffer of ONE all-gather); import takes the compact rows of ALL of a resident's pods; pack then runs FFD
 * order + bin-pack + cheapest SKU on them without evaluating a pod again ("merge the global packing").
 * Row: [int8 template | n_templates code bytes | padding], b200s_compact_row_bytes(n_templates) bytes. */
int b200s_compact_row_bytes(int32_t n_templates);
int b200s_resident_export_compact(b200s_ctx_t* ctx, b200s_resident_t* r, void* d_out);
int b200s_resident_import_compact(b200s_ctx_t* ctx, b200s_resident_t* r, const void* d_in);
int b200s_resident_pack(b200s_ctx_t* ctx, b200s_resident_t* r);
/* number of kernels the last resident_* call enqueued */
int b200s_resident_last_launches(const b200s_resident_t* r);
/* device time of the dense fit kernel alone in the last stage A (CUDA events on the context's stream, recorded on
 * every launch); blocks until that kernel has finished */
int b200s_resident_dense_kernel_ms(b200s_ctx_t* ctx, b200s_resident_t* r, float* ms);
/* diagnostic: SM cycles thread 0 of problem 0's pack CTA spent per phase of the last solve -- out[0..4] = step start,
 * existing nodes, in-flight scan, commit + next pod's row, rest of the order normalisation; out[7] = pod steps;
 * out[8..10] = cycles inside the commit phase by role (queue warp: next pod's row, warp 2: claim order, winning lane
 * group: claim state); out[11] = full pdqsort runs << 40 | their cycles; out[5..6] reserved */
int b200s_resident_pack_profile(b200s_ctx_t* ctx, b200s_resident_t* r, uint64_t out[12]);
/* The cycle counters cost 3-4 % of a step: solves run without them unless they are switched on here beforehand
 * (without them only out[7], the step count, and out[11] are filled). */
int b200s_resident_set_pack_profile(b200s_resident_t* r, int on);

/* ---- config 1: the vmSize pick of instance.Provider.Create (host only, no device) ----
 * name: NodeClaim name; instance_types: the values of the node.kubernetes.io/instance-type
 * requirement in the caller's order; prices: optional per-name price (NULL = keep the
 * reference's values[0] rule); storage_bytes: spec.resources.requests["storage"].
 * Outputs: chosen index, OS disk size in GB, 1 if the SKU is labelled gpu (instance.go:335-339). */
#define B200S_CREATE_EBADNAME (-20)
#define B200S_CREATE_ENOTYPE (-21)
#define B200S_CREATE_ENOSTORAGE (-22)
int b200s_create_select(const char* name, int32_t n_types, const char* const* instance_types,
                        const double* prices, int64_t storage_bytes, int32_t* chosen, int32_t* disk_gb,
                        int32_t* is_gpu);

/* ---- host helpers for the two per-object computations the model's int64 vectors come from ----
 * b200s_pod_requests: resources.RequestsForPods for one pod (K/utils/resources/resources.go:30-38, :114-144):
 *   per container requests, falling back to limits where a request is absent (:204-217); sum over containers;
 *   max with the init-container high-water mark (restartPolicy Always init containers are sidecars: they add to
 *   the sum and to every later init container's need); plus pod overhead; then pods is SET to 1.
 *   Arrays are dense [count * n_res] int64 with a presence byte per entry; absent means "not in the ResourceList".
 * b200s_allocatable: InstanceType.Allocatable (K/cloudprovider/types.go:181-214): capacity minus
 *   (kubeReserved + systemReserved + evictionThreshold) for the resources capacity lists, then every hugepages-*
 *   capacity is taken off memory, clamped at zero. */
int b200s_pod_requests(int32_t n_res, int32_t res_pods, int32_t n_containers, const int64_t* req, const uint8_t* req_has,
                       const int64_t* lim, const uint8_t* lim_has, int32_t n_init, const int64_t* init_req,
                       const uint8_t* init_req_has, const int64_t* init_lim, const uint8_t* init_lim_has,
                       const uint8_t* init_restart_always, const int64_t* overhead, int64_t* out);
int b200s_allocatable(int32_t n_res, int32_t res_memory, const int64_t* capacity, const uint8_t* capacity_has,
                      const int64_t* kube_reserved, const int64_t* system_reserved, const int64_t* eviction_threshold,
                      const uint8_t* is_hugepages, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* B200SCHED_H */

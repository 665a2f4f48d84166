// encoder.cpp — b200s_encode: semantic model (b200s_model_t) -> device-facing SoA + bit-row tables.
//
// Host code, no CUDA calls.  Total or reject: any input the exact encoding cannot represent returns
// B200S_ENOTREPRESENTABLE with a reason, and the Go side keeps using the vendored CPU Solve for that
// batch.  Nothing is ever approximated.
//
// What is built here, and the reference code each piece stands in for:
//   * NodeClaimTemplates in OrderByWeight order with their merged requirement sets
//     (K/controllers/provisioning/scheduling/nodeclaimtemplate.go:56-75,
//      K/utils/nodepool/nodepool.go:121-131)
//   * existing nodes in sortExistingNodes order (scheduler.go:729-742)
//   * per-pod request vectors, requirement sets and the tolerated-taint bitset
//     (scheduler.go:467-486; K/scheduling/taints.go:53-65; vendor/k8s.io/api/core/v1/toleration.go:38-57)
//   * catalog index tables: per resource the ascending distinct Allocatable() values with
//     "alloc >= value" SKU bit-rows, per (label key, value) the SKUs admitting it, per (zone,
//     capacity-type) the SKUs with an available offering, and the cheapest such offering's price
//     (K/cloudprovider/types.go:105-121,181-219,355-362)
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "b200_core.cuh"

using namespace b200;

struct b200s_encoded {
  b200s_catalog_soa_t cat{};
  b200s_templates_t tmpl{};
  b200s_pods_soa_t pods{};
  b200s_nodes_soa_t nodes{};
  // owned storage
  std::vector<int64_t> val_int, res_dv, cap_dv, alloc_rm, cap_rm, rq_requests, pod_ts, node_available;
  std::vector<uint64_t> val_isint, rows, price, price_max, type_admit, sig_mask, sig_flags, sig_tolerated, pod_uid_hi, pod_uid_lo,
      node_req_mask, node_req_flags, node_taints;
  std::vector<int32_t> sig_gt, sig_lt, pod_sig, pod_reqcls, node_index;
  // decode maps
  std::vector<int> devkey_model;               // device key slot -> model key
  std::vector<std::vector<int>> devkey_values; // device key slot -> model value ids by bit
};

namespace {

struct Reject {
  std::string why;
};

[[noreturn]] void reject(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Reject{buf};
}

struct Ctx {
  const b200s_model_t* M;
  b200s_encoded* E;
  KeyTab kt{};                    // host copy used for the algebra
  std::vector<int> keyslot;       // model key -> device slot or -1
  std::vector<int> valbit;        // model value id -> bit within its key
  std::vector<int> type_value_sku;// value id of the instance-type key -> SKU index or -1
};

// requirement list [r0,r1) of the model -> per-slot EReq via Requirements.Add semantics
// (requirements.go:127-134).  Keys without a device slot are handled by `on_unslotted`.
struct ReqSet {
  EReq q[MAXK];
  ReqSet() {
    for (auto& e : q) e = EReq{0, 0, 0, 0};
  }
};

EReq ereq_from_model(Ctx& c, const b200s_req_t& r, int slot) {
  const b200s_model_t* M = c.M;
  EReq e{0, 0, 0, RF_DEF};
  if (r.min_values >= 0) reject("minValues on key %d is not representable", r.key);
  switch (r.op) {
    case B200S_OP_IN:
    case B200S_OP_NOTIN:
      for (int i = 0; i < r.val_cnt; i++) {
        int v = M->req_values[r.val_off + i];
        if (v < 0 || v >= M->n_vals || M->val_key[v] != r.key) reject("requirement value %d does not belong to key %d", v, r.key);
        e.mask |= 1ull << c.valbit[v];
      }
      if (r.op == B200S_OP_NOTIN) e.f |= RF_CMPL;
      break;
    case B200S_OP_EXISTS: e.f |= RF_CMPL; break;
    case B200S_OP_DOESNOTEXIST: break;
    case B200S_OP_GT: e.f |= RF_CMPL | RF_GT; e.gt = r.bound; break;
    case B200S_OP_LT: e.f |= RF_CMPL | RF_LT; e.lt = r.bound; break;
    default: reject("unknown operator %d", r.op);
  }
  (void)slot;
  return e;
}

void encode_ereq(const ReqSet& rs, int K, const KeyTab& kt, uint64_t* mask_out, size_t mask_stride, uint64_t* flags_out) {
  uint64_t f = 0;
  for (int k = 0; k < K; k++) {
    mask_out[(size_t)k * mask_stride] = rs.q[k].mask;
    f |= (uint64_t)(rs.q[k].f & 0xFu) << (4 * k);
  }
  *flags_out = f;
  (void)kt;
}

bool tolerates_taint(const b200s_toleration_t& t, const b200s_taint_t& taint) {  // toleration.go:38-57
  if (t.effect != B200S_EFFECT_NONE && t.effect != taint.effect) return false;
  if (t.key != 0 && t.key != taint.key) return false;
  switch (t.op) {
    case B200S_TOLOP_EMPTY:
    case B200S_TOLOP_EQUAL: return t.value == taint.value;
    case B200S_TOLOP_EXISTS: return true;
    default: return false;
  }
}

void do_encode(Ctx& c) {
  const b200s_model_t* M = c.M;
  b200s_encoded* E = c.E;
  const int S = M->n_skus, R = M->n_res, P = M->n_pods, NE = M->n_nodes;
  if (S <= 0) reject("empty catalog");
  if (R <= 0 || R > MAXR) reject("%d resource dimensions (max %d)", R, MAXR);
  if (M->n_pools > MAXT) reject("%d NodePools (max %d)", M->n_pools, MAXT);
  const int W = (S + 63) / 64;
  const int KT = M->key_instance_type;

  // ---- device key slots: keys used by SKUs, offerings, pools or pods (node-only labels cannot
  // influence any check: Intersects only visits keys present on both sides and the undefined-key
  // rule iterates the pod's keys) ----
  c.keyslot.assign(M->n_keys, -1);
  std::vector<uint8_t> used(M->n_keys, 0);
  auto mark = [&](int r0, int r1) {
    for (int i = r0; i < r1; i++) {
      int k = M->reqs[i].key;
      if (k < 0 || k >= M->n_keys) reject("requirement key %d out of range", k);
      used[k] = 1;
    }
  };
  mark(M->sku_req_off[0], M->sku_req_off[S]);
  if (M->n_offerings) mark(M->off_req_off[0], M->off_req_off[M->n_offerings]);
  if (M->n_pools) mark(M->pool_req_off[0], M->pool_req_off[M->n_pools]);
  if (P) mark(M->pod_req_off[0], M->pod_req_off[P]);
  int K = 0;
  for (int k = 0; k < M->n_keys; k++) {
    if (!used[k] || k == KT) continue;
    if (K >= MAXK) reject("more than %d label keys take part in requirements", MAXK);
    c.keyslot[k] = K++;
    E->devkey_model.push_back(k);
  }
  E->devkey_values.assign(K, {});
  c.valbit.assign(M->n_vals, 0);
  {
    std::vector<int> cnt(M->n_keys, 0);
    for (int v = 0; v < M->n_vals; v++) {
      int k = M->val_key[v];
      if (k < 0 || k >= M->n_keys) reject("value %d has key %d out of range", v, k);
      c.valbit[v] = cnt[k]++;
      int slot = c.keyslot[k];
      if (slot >= 0) {
        if (c.valbit[v] >= B200S_MAX_VALUES) reject("label key %d has more than %d values", k, B200S_MAX_VALUES);
        E->devkey_values[slot].push_back(v);
      }
    }
  }

  KeyTab& kt = c.kt;
  kt.n_keys = K; kt.n_res = R; kt.n_words = W; kt.n_skus = S;
  kt.key_zone = M->key_zone >= 0 ? c.keyslot[M->key_zone] : -1;
  kt.key_ct = M->key_capacity_type >= 0 ? c.keyslot[M->key_capacity_type] : -1;
  kt.wellknown = 0;
  E->val_int.assign((size_t)std::max(K, 1) * 64, 0);
  E->val_isint.assign(std::max(K, 1), 0);
  for (int s = 0; s < K; s++) {
    int mk = E->devkey_model[s];
    if (M->key_wellknown[mk]) kt.wellknown |= 1u << s;
    kt.nvals[s] = (int)E->devkey_values[s].size();
    kt.bslot[s] = -1;
    for (size_t b = 0; b < E->devkey_values[s].size(); b++) {
      int v = E->devkey_values[s][b];
      if (M->val_isint[v]) {
        E->val_isint[s] |= 1ull << b;
        E->val_int[(size_t)s * 64 + b] = M->val_int[v];
      }
    }
    kt.isint[s] = E->val_isint[s];
  }
  kt.val_int = E->val_int.data();
  kt.n_zones = kt.key_zone >= 0 ? kt.nvals[kt.key_zone] : 0;
  kt.n_cts = kt.key_ct >= 0 ? kt.nvals[kt.key_ct] : 0;
  if ((int64_t)kt.n_zones * kt.n_cts > B200S_MAX_ZC) reject("%d zones x %d capacity types exceeds %d", kt.n_zones, kt.n_cts, B200S_MAX_ZC);

  // ---- bounded keys: any Gt/Lt in pools or pods ----
  int NB = 0;
  int bounded_key[MAXB];
  auto scan_bounds = [&](int r0, int r1, bool is_pod) {
    for (int i = r0; i < r1; i++) {
      const b200s_req_t& r = M->reqs[i];
      if (r.op != B200S_OP_GT && r.op != B200S_OP_LT) continue;
      int slot = c.keyslot[r.key];
      if (slot < 0) reject("Gt/Lt on the instance-type key");
      if (is_pod && (r.bound < INT32_MIN || r.bound > INT32_MAX)) reject("pod Gt/Lt bound %lld outside int32", (long long)r.bound);
      if (kt.bslot[slot] < 0) {
        if (NB >= MAXB) reject("more than %d label keys carry Gt/Lt bounds", MAXB);
        kt.bslot[slot] = (int8_t)NB;
        bounded_key[NB++] = slot;
      }
    }
  };
  if (M->n_pools) scan_bounds(M->pool_req_off[0], M->pool_req_off[M->n_pools], false);
  if (P) scan_bounds(M->pod_req_off[0], M->pod_req_off[P], true);
  kt.n_bounded = NB;

  auto build_set = [&](int r0, int r1, ReqSet* rs, const char* what, std::vector<EReq>* type_reqs) {
    for (int i = r0; i < r1; i++) {
      const b200s_req_t& r = M->reqs[i];
      if (r.key == KT) {
        if (!type_reqs) reject("%s carries an instance-type requirement", what);
        if (r.min_values >= 0) reject("minValues is not representable");
        // kept in model value ids; resolved against the catalog below
        EReq marker{(uint64_t)i, 0, 0, RF_DEF};
        type_reqs->push_back(marker);
        continue;
      }
      int slot = c.keyslot[r.key];
      if (slot < 0) continue;  // node-only label
      EReq e = ereq_from_model(c, r, slot);
      rs->q[slot] = add_req(kt, slot, rs->q[slot], e);
    }
  };

  // ---- catalog rows ----
  std::vector<ReqSet> sku_req(S);
  c.type_value_sku.assign(M->n_vals, -1);
  for (int s = 0; s < S; s++) {
    std::vector<EReq> treqs;
    build_set(M->sku_req_off[s], M->sku_req_off[s + 1], &sku_req[s], "an instance type", &treqs);
    for (int k = 0; k < K; k++)
      if ((sku_req[s].q[k].f & RF_DEF) && (sku_req[s].q[k].f & RF_CMPL))
        reject("instance type %d has a NotIn/Exists/Gt/Lt requirement on key %d", s, E->devkey_model[k]);
    if (KT >= 0) {
      if (treqs.size() != 1) reject("instance type %d must carry exactly one instance-type requirement", s);
      const b200s_req_t& r = M->reqs[treqs[0].mask];
      if (r.op != B200S_OP_IN || r.val_cnt != 1) reject("instance type %d: instance-type requirement must be In {name}", s);
      int v = M->req_values[r.val_off];
      if (c.type_value_sku[v] >= 0) reject("instance types %d and %d share a name", c.type_value_sku[v], s);
      c.type_value_sku[v] = s;
    }
  }

  // distinct allocatable values and suffix rows
  std::vector<std::vector<int64_t>> dv(R);
  int n_rows = 1;  // row 0 = nonneg
  E->cat.row_nonneg = 0;
  for (int r = 0; r < R; r++) {
    for (int s = 0; s < S; s++) dv[r].push_back(M->sku_alloc[(size_t)s * R + r]);
    std::sort(dv[r].begin(), dv[r].end());
    dv[r].erase(std::unique(dv[r].begin(), dv[r].end()), dv[r].end());
    E->cat.res_nd[r] = (int)dv[r].size();
    E->cat.res_dv0[r] = (int)E->res_dv.size();
    E->cat.res_row0[r] = n_rows;
    n_rows += (int)dv[r].size();
    E->res_dv.insert(E->res_dv.end(), dv[r].begin(), dv[r].end());
  }
  for (int k = 0; k < K; k++) {
    E->cat.key_row0[k] = n_rows;
    n_rows += kt.nvals[k] + 2;
  }
  E->cat.zc_row0 = n_rows;
  const int ZC = kt.n_zones * kt.n_cts;
  n_rows += ZC;
  // distinct Capacity values and prefix rows "capacity[r] <= cv[i]" (NodePool limits, scheduler.go:831-867)
  std::vector<std::vector<int64_t>> cv(R);
  for (int r = 0; r < R; r++) {
    for (int s = 0; s < S; s++) cv[r].push_back(M->sku_capacity[(size_t)s * R + r]);
    std::sort(cv[r].begin(), cv[r].end());
    cv[r].erase(std::unique(cv[r].begin(), cv[r].end()), cv[r].end());
    E->cat.cap_nd[r] = (int)cv[r].size();
    E->cat.cap_dv0[r] = (int)E->cap_dv.size();
    E->cat.cap_row0[r] = n_rows;
    n_rows += (int)cv[r].size();
    E->cap_dv.insert(E->cap_dv.end(), cv[r].begin(), cv[r].end());
  }
  E->rows.assign((size_t)n_rows * W, 0);
  auto setbit = [&](int row, int s) { E->rows[(size_t)row * W + s / 64] |= 1ull << (s % 64); };
  E->alloc_rm.assign((size_t)R * S, 0);
  E->cap_rm.assign((size_t)R * S, 0);
  for (int s = 0; s < S; s++) {
    bool nonneg = true;
    for (int r = 0; r < R; r++) {
      int64_t a = M->sku_alloc[(size_t)s * R + r];
      E->alloc_rm[(size_t)r * S + s] = a;
      E->cap_rm[(size_t)r * S + s] = M->sku_capacity[(size_t)s * R + r];
      if (a < 0) nonneg = false;
      int idx = (int)(std::lower_bound(dv[r].begin(), dv[r].end(), a) - dv[r].begin());
      for (int i = 0; i <= idx; i++) setbit(E->cat.res_row0[r] + i, s);  // alloc >= dv[i] for i <= idx
      int cidx = (int)(std::lower_bound(cv[r].begin(), cv[r].end(), M->sku_capacity[(size_t)s * R + r]) - cv[r].begin());
      for (int i = cidx; i < (int)cv[r].size(); i++) setbit(E->cat.cap_row0[r] + i, s);  // capacity <= cv[i] for i >= cidx
    }
    if (nonneg) setbit(0, s);
    for (int k = 0; k < K; k++) {
      const EReq& q = sku_req[s].q[k];
      int row0 = E->cat.key_row0[k], nv = kt.nvals[k];
      if (!(q.f & RF_DEF)) { setbit(row0 + nv, s); continue; }
      if (q.mask == 0) setbit(row0 + nv + 1, s);  // DoesNotExist
      uint64_t m = q.mask;
      while (m) {
        int v = ctz64(m);
        m &= m - 1;
        setbit(row0 + v, s);
      }
    }
  }
  // offerings (types.go:355-362): exactly zone In{z} and capacity-type In{c}
  E->price.assign((size_t)std::max(ZC, 1) * S, B200S_PRICE_NONE);
  E->price_max.assign((size_t)std::max(ZC, 1) * S, B200S_PRICE_ABSENT);
  for (int s = 0; s < S; s++) {
    for (int o = M->sku_off_off[s]; o < M->sku_off_off[s + 1]; o++) {
      int z = -1, ct = -1;
      for (int i = M->off_req_off[o]; i < M->off_req_off[o + 1]; i++) {
        const b200s_req_t& r = M->reqs[i];
        if (r.op != B200S_OP_IN || r.val_cnt != 1) reject("offering %d: requirement on key %d is not a single-value In", o, r.key);
        int v = M->req_values[r.val_off];
        if (r.key == M->key_zone && z < 0) z = c.valbit[v];
        else if (r.key == M->key_capacity_type && ct < 0) ct = c.valbit[v];
        else reject("offering %d: only one zone and one capacity-type requirement are representable (key %d)", o, r.key);
      }
      if (z < 0 || ct < 0) reject("offering %d lacks a zone or capacity-type requirement", o);
      double p = M->off_price[o];
      if (!(p >= 0.0) || !std::isfinite(p)) reject("offering %d: price must be finite and non-negative", o);
      if (!M->off_available[o]) continue;
      int zc = z * kt.n_cts + ct;
      setbit(E->cat.zc_row0 + zc, s);
      uint64_t bits;
      std::memcpy(&bits, &p, 8);
      uint64_t& slot = E->price[(size_t)zc * S + s];
      if (bits < slot) slot = bits;
      uint64_t& top = E->price_max[(size_t)zc * S + s];  // lo.MaxBy keeps the first of equals: same price either way
      if (top == B200S_PRICE_ABSENT || bits > top) top = bits;
    }
  }

  // ---- catalog struct ----
  b200s_catalog_soa_t& cat = E->cat;
  cat.n_skus = S; cat.n_words = W; cat.n_res = R; cat.n_keys = K;
  cat.n_zones = kt.n_zones; cat.n_cts = kt.n_cts; cat.key_zone = kt.key_zone; cat.key_ct = kt.key_ct;
  cat.wellknown = kt.wellknown;
  for (int k = 0; k < K; k++) cat.key_nvals[k] = kt.nvals[k];
  cat.n_bounded = NB;
  for (int b = 0; b < NB; b++) cat.bounded_key[b] = bounded_key[b];
  cat.val_int = E->val_int.data();
  cat.val_isint = E->val_isint.data();
  cat.n_rows = n_rows;
  cat.rows = E->rows.data();
  cat.res_dv = E->res_dv.data();
  cat.n_dv = (int)E->res_dv.size();
  cat.price = E->price.data();
  cat.price_max = E->price_max.data();
  cat.alloc = E->alloc_rm.data();
  cat.capacity = E->cap_rm.data();
  cat.cap_dv = E->cap_dv.data();
  cat.n_cap_dv = (int)E->cap_dv.size();
  {  // the key layout this encode chose (FNV-1a): slots, value bits, numerics, well-known mask, bounded slots
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) {
      for (int i = 0; i < 8; i++) { h ^= (v >> (8 * i)) & 0xFFu; h *= 1099511628211ull; }
    };
    mix((uint64_t)K); mix((uint64_t)R); mix((uint64_t)(int64_t)kt.key_zone); mix((uint64_t)(int64_t)kt.key_ct);
    mix(kt.wellknown); mix((uint64_t)NB);
    for (int b = 0; b < NB; b++) mix((uint64_t)bounded_key[b]);
    for (int k = 0; k < K; k++) {
      mix((uint64_t)E->devkey_model[k]);
      mix((uint64_t)E->devkey_values[k].size());
      for (int v : E->devkey_values[k]) mix((uint64_t)v);
      mix(E->val_isint[k]);
      for (size_t b = 0; b < E->devkey_values[k].size(); b++) mix((uint64_t)E->val_int[(size_t)k * 64 + b]);
    }
    if (h == 0) h = 1;
    cat.layout_hash = h;
  }

  // ---- taint universe over pools + nodes ----
  std::vector<b200s_taint_t> taint_u;
  auto taint_bit = [&](const b200s_taint_t& t) {
    for (size_t i = 0; i < taint_u.size(); i++)
      if (taint_u[i].key == t.key && taint_u[i].value == t.value && taint_u[i].effect == t.effect) return (int)i;
    if (taint_u.size() >= B200S_MAX_TAINTS) reject("more than %d distinct taints", B200S_MAX_TAINTS);
    taint_u.push_back(t);
    return (int)taint_u.size() - 1;
  };

  // ---- templates: OrderByWeight (nodepool.go:121-131) ----
  std::vector<int> order(M->n_pools);
  for (int i = 0; i < M->n_pools; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) {  // a total order: names are unique
    if (M->pool_weight[a] == M->pool_weight[b]) return M->pool_name_rank[a] > M->pool_name_rank[b];
    return M->pool_weight[a] > M->pool_weight[b];
  });
  b200s_templates_t& T = E->tmpl;
  T.n_templates = M->n_pools;
  bool any_type_req = false;
  E->type_admit.assign((size_t)std::max(M->n_pools, 1) * W, 0);
  for (int t = 0; t < M->n_pools; t++) {
    int pi = order[t];
    T.pool_index[t] = pi;
    for (int r = 0; r < MAXR; r++) { T.limit[t][r] = 0; T.has_limit[t][r] = 0; }
    for (int r = 0; r < R; r++) {
      if (M->pool_has_limit && M->pool_has_limit[(size_t)pi * R + r]) {  // s.remainingResources[pool] (scheduler.go:170-172)
        T.has_limit[t][r] = 1;
        T.limit[t][r] = M->pool_limit[(size_t)pi * R + r];
      }
      int64_t d = M->pool_daemon[(size_t)pi * R + r];
      if (d < 0) reject("negative daemon overhead");
      T.daemon[t][r] = d;
    }
    uint64_t tb = 0;
    for (int i = M->pool_taint_off[pi]; i < M->pool_taint_off[pi + 1]; i++) {
      if (M->taints[i].effect == B200S_EFFECT_PREFERNOSCHEDULE)
        reject("NodePool %d has a PreferNoSchedule taint (preference relaxation is not kernelised)", pi);
      tb |= 1ull << taint_bit(M->taints[i]);
    }
    T.taints[t] = tb;
    ReqSet rs;
    std::vector<EReq> treqs;
    build_set(M->pool_req_off[pi], M->pool_req_off[pi + 1], &rs, "a NodePool", &treqs);
    encode_ereq(rs, K, kt, &T.req_mask[t][0], 1, &T.req_flags[t]);
    for (int b = 0; b < MAXB; b++) { T.req_gt[t][b] = 0; T.req_lt[t][b] = 0; }
    for (int k = 0; k < K; k++) {
      if (rs.q[k].f & (RF_GT | RF_LT)) {
        int b = kt.bslot[k];
        T.req_gt[t][b] = rs.q[k].gt;
        T.req_lt[t][b] = rs.q[k].lt;
      }
    }
    // instance-type requirement of the pool -> SKU admit mask.  Evaluated with the reference's own
    // set algebra on value ids: the SKU side is In {name}, so HasIntersection reduces to "name is
    // admitted" and the NotIn/DoesNotExist forgiveness never applies to an In set.
    uint64_t* adm = &E->type_admit[(size_t)t * W];
    for (int w = 0; w < W; w++) adm[w] = ~0ull;
    if (!treqs.empty()) {
      any_type_req = true;
      // fold the pool's instance-type requirements with Add semantics over SKU-index sets
      bool cmpl = true;
      std::vector<uint8_t> set(S, 0);  // In set, or exclusion set when cmpl
      bool first = true;
      for (auto& mk : treqs) {
        const b200s_req_t& r = M->reqs[mk.mask];
        bool rc;
        std::vector<uint8_t> rset(S, 0);
        switch (r.op) {
          case B200S_OP_IN: rc = false; break;
          case B200S_OP_NOTIN: rc = true; break;
          case B200S_OP_EXISTS: rc = true; break;
          case B200S_OP_DOESNOTEXIST: rc = false; break;
          default: reject("Gt/Lt on the instance-type key");
        }
        for (int i = 0; i < r.val_cnt; i++) {
          int s = c.type_value_sku[M->req_values[r.val_off + i]];
          if (s >= 0) rset[s] = 1;  // names outside the catalog cannot match any SKU
        }
        if (first) { cmpl = rc; set = rset; first = false; continue; }
        std::vector<uint8_t> out(S, 0);
        for (int s = 0; s < S; s++)
          out[s] = (cmpl && rc) ? (set[s] | rset[s]) : cmpl ? (rset[s] & !set[s]) : rc ? (set[s] & !rset[s]) : (set[s] & rset[s]);
        cmpl = cmpl && rc;
        set = out;
      }
      for (int s = 0; s < S; s++) {
        bool admitted = cmpl ? !set[s] : set[s];
        if (!admitted) adm[s / 64] &= ~(1ull << (s % 64));
      }
    }
  }
  T.type_admit = any_type_req ? E->type_admit.data() : nullptr;

  // ---- existing nodes: sortExistingNodes (scheduler.go:729-742) ----
  std::vector<int> norder(NE);
  for (int i = 0; i < NE; i++) norder[i] = i;
  std::stable_sort(norder.begin(), norder.end(), [&](int a, int b) {
    bool ia = M->node_initialized[a], ib = M->node_initialized[b];
    if (ia && !ib) return true;
    if (!ia && ib) return false;
    return M->node_name_rank[a] < M->node_name_rank[b];
  });
  E->node_index.assign(norder.begin(), norder.end());
  E->node_available.assign((size_t)R * std::max(NE, 1), 0);
  E->node_req_mask.assign((size_t)std::max(K, 1) * std::max(NE, 1), 0);
  E->node_req_flags.assign(std::max(NE, 1), 0);
  E->node_taints.assign(std::max(NE, 1), 0);
  for (int e = 0; e < NE; e++) {
    int ni = norder[e];
    for (int r = 0; r < R; r++) E->node_available[(size_t)r * NE + e] = M->node_available[(size_t)ni * R + r];
    ReqSet rs;
    std::vector<EReq> ignored;
    build_set(M->node_req_off[ni], M->node_req_off[ni + 1], &rs, "a node", &ignored);
    for (int k = 0; k < K; k++)
      if (rs.q[k].f & (RF_GT | RF_LT)) reject("node %d carries a bounded requirement", ni);
    encode_ereq(rs, K, kt, &E->node_req_mask[e], (size_t)NE, &E->node_req_flags[e]);
    uint64_t tb = 0;
    for (int i = M->node_taint_off[ni]; i < M->node_taint_off[ni + 1]; i++) tb |= 1ull << taint_bit(M->taints[i]);
    E->node_taints[e] = tb;
  }
  b200s_nodes_soa_t& N = E->nodes;
  N.n_nodes = NE; N.n_res = R; N.n_keys = K;
  N.node_index = E->node_index.data();
  N.available = E->node_available.data();
  N.req_mask = E->node_req_mask.data();
  N.req_flags = E->node_req_flags.data();
  N.taints = E->node_taints.data();

  // ---- pods: dictionary-encode requirement signatures and request vectors ----
  // Replicas of one workload share both, so the pod arrays shrink to two ids per pod and the kernels
  // evaluate each distinct signature / request vector once.  Exact: a class is a bit-for-bit equal row.
  // A dictionary of fixed-width rows: open addressing over the row hash, rows kept in one flat array (the per-pod cost of
  // the encode is these two look-ups; a node-based map with vector keys was half of b200s_encode's time).
  struct RowDict {
    size_t width = 0;
    std::vector<uint64_t> rows;   // [n * width]
    std::vector<int32_t> slot;    // id or -1
    size_t n = 0, mask = 0;
    void init(size_t w, size_t expect) {
      width = w;
      size_t cap = 64;
      while (cap < 2 * expect) cap <<= 1;
      slot.assign(cap, -1);
      mask = cap - 1;
    }
    static uint64_t hash(const uint64_t* v, size_t w) {
      uint64_t h = 0x9E3779B97F4A7C15ull;
      for (size_t i = 0; i < w; i++) { h ^= v[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xBF58476D1CE4E5B9ull; }
      return h ^ (h >> 29);
    }
    void grow() {
      std::vector<int32_t> ns(slot.size() * 2, -1);
      const size_t nm = ns.size() - 1;
      for (size_t id = 0; id < n; id++) {
        size_t i = (size_t)hash(rows.data() + id * width, width) & nm;
        while (ns[i] >= 0) i = (i + 1) & nm;
        ns[i] = (int32_t)id;
      }
      slot.swap(ns);
      mask = nm;
    }
    int32_t id_of(const uint64_t* v) {
      size_t i = (size_t)hash(v, width) & mask;
      for (;; i = (i + 1) & mask) {
        const int32_t id = slot[i];
        if (id < 0) break;
        if (std::memcmp(rows.data() + (size_t)id * width, v, width * 8) == 0) return id;
      }
      slot[i] = (int32_t)n;
      rows.insert(rows.end(), v, v + width);
      n++;
      if (2 * n > slot.size()) grow();
      return (int32_t)(n - 1);
    }
    const uint64_t* row(size_t id) const { return rows.data() + id * width; }
  };
  RowDict sig_of, rq_of;
  rq_of.init((size_t)R, 1024);
  sig_of.init((size_t)K + 2 + 2 * (size_t)NB, 1024);
  const size_t Pz = (size_t)std::max(P, 1);
  E->pod_sig.assign(Pz, 0);
  E->pod_reqcls.assign(Pz, 0);
  E->pod_ts.assign(M->pod_ts, M->pod_ts + P);
  E->pod_ts.resize(Pz);
  E->pod_uid_hi.assign(M->pod_uid_hi, M->pod_uid_hi + P);
  E->pod_uid_hi.resize(Pz);
  E->pod_uid_lo.assign(M->pod_uid_lo, M->pod_uid_lo + P);
  E->pod_uid_lo.resize(Pz);
  std::vector<long double> sums(R, 0);
  std::vector<uint64_t> row;
  int32_t plain_sig = -1;  // signature id of a pod without requirements and tolerations, once one has been seen
  for (int p = 0; p < P; p++) {
    row.assign((size_t)R, 0);
    for (int r = 0; r < R; r++) {
      int64_t q = M->pod_requests[(size_t)p * R + r];
      if (q < 0) reject("pod %d has a negative request", p);
      row[r] = (uint64_t)q;
      sums[r] += (long double)q;
    }
    E->pod_reqcls[p] = rq_of.id_of(row.data());

    int r0 = M->pod_req_off[p], r1 = M->pod_req_off[p + 1];
    if (r1 == r0 && M->pod_tol_off[p + 1] == M->pod_tol_off[p] && plain_sig >= 0) {
      E->pod_sig[p] = plain_sig;  // no requirement, no toleration: the signature the first such pod got
      continue;
    }
    ReqSet rs;
    if (r1 > r0) build_set(r0, r1, &rs, "a pod", nullptr);
    row.assign((size_t)K + 2 + 2 * (size_t)NB, 0);
    uint64_t f = 0;
    for (int k = 0; k < K; k++) {
      row[k] = rs.q[k].mask;
      f |= (uint64_t)(rs.q[k].f & 0xFu) << (4 * k);
      if (rs.q[k].f & (RF_GT | RF_LT)) {
        int b = kt.bslot[k];
        row[(size_t)K + 2 + b] = (uint64_t)(int64_t)(int32_t)rs.q[k].gt;
        row[(size_t)K + 2 + NB + b] = (uint64_t)(int64_t)(int32_t)rs.q[k].lt;
        if ((rs.q[k].f & RF_CMPL) && rs.q[k].mask != 0 && NE > 0)
          reject("pod %d combines NotIn with Gt/Lt on one key while existing nodes are present", p);
      }
    }
    row[K] = f;
    uint64_t tol = 0;
    int t0 = M->pod_tol_off[p], t1 = M->pod_tol_off[p + 1];
    for (size_t ti = 0; ti < taint_u.size(); ti++) {
      bool ok = false;
      for (int i = t0; i < t1 && !ok; i++) ok = tolerates_taint(M->tolerations[i], taint_u[ti]);
      if (ok) tol |= 1ull << ti;
    }
    row[(size_t)K + 1] = tol;
    E->pod_sig[p] = sig_of.id_of(row.data());
    if (r1 == r0 && t1 == t0) plain_sig = E->pod_sig[p];
  }
  for (int r = 0; r < R; r++) {
    long double mx = 0;
    for (int t = 0; t < M->n_pools; t++) mx = std::max(mx, (long double)T.daemon[t][r]);
    if (sums[r] + mx > 9.0e18L) reject("accumulated requests of resource %d could overflow int64", r);
  }
  const size_t NS = sig_of.n, NQ = rq_of.n, NS1 = std::max<size_t>(NS, 1), NQ1 = std::max<size_t>(NQ, 1);
  E->sig_mask.assign((size_t)std::max(K, 1) * NS1, 0);
  E->sig_flags.assign(NS1, 0);
  E->sig_tolerated.assign(NS1, 0);
  E->sig_gt.assign((size_t)std::max(NB, 1) * NS1, 0);
  E->sig_lt.assign((size_t)std::max(NB, 1) * NS1, 0);
  for (size_t i = 0; i < NS; i++) {
    const uint64_t* rw = sig_of.row(i);
    for (int k = 0; k < K; k++) E->sig_mask[(size_t)k * NS + i] = rw[k];
    E->sig_flags[i] = rw[K];
    E->sig_tolerated[i] = rw[(size_t)K + 1];
    for (int b = 0; b < NB; b++) {
      E->sig_gt[(size_t)b * NS + i] = (int32_t)(int64_t)rw[(size_t)K + 2 + b];
      E->sig_lt[(size_t)b * NS + i] = (int32_t)(int64_t)rw[(size_t)K + 2 + NB + b];
    }
  }
  E->rq_requests.assign((size_t)R * NQ1, 0);
  for (size_t i = 0; i < NQ; i++)
    for (int r = 0; r < R; r++) E->rq_requests[(size_t)r * NQ + i] = (int64_t)rq_of.row(i)[r];
  b200s_pods_soa_t& Pd = E->pods;
  Pd.n_pods = P; Pd.n_res = R; Pd.n_keys = K; Pd.n_bounded = NB;
  Pd.res_cpu = M->res_cpu; Pd.res_memory = M->res_memory;
  Pd.n_sigs = (int32_t)NS; Pd.n_reqcls = (int32_t)NQ;
  Pd.sig_id = E->pod_sig.data();
  Pd.req_id = E->pod_reqcls.data();
  Pd.ts = E->pod_ts.data();
  Pd.uid_hi = E->pod_uid_hi.data();
  Pd.uid_lo = E->pod_uid_lo.data();
  Pd.sig_mask = E->sig_mask.data();
  Pd.sig_flags = E->sig_flags.data();
  Pd.sig_gt = E->sig_gt.data();
  Pd.sig_lt = E->sig_lt.data();
  Pd.sig_tolerated = E->sig_tolerated.data();
  Pd.rq_requests = E->rq_requests.data();
  Pd.layout_hash = E->cat.layout_hash;
  E->nodes.layout_hash = E->cat.layout_hash;
}

}  // namespace

extern "C" {

int b200s_encode(const b200s_model_t* model, b200s_encoded_t** out, char* err, size_t err_cap) {
  if (!model || !out) return B200S_EINVAL;
  *out = nullptr;
  std::unique_ptr<b200s_encoded> E(new (std::nothrow) b200s_encoded());
  if (!E) return B200S_ENOMEM;
  Ctx c;
  c.M = model;
  c.E = E.get();
  try {
    do_encode(c);
  } catch (const Reject& r) {
    if (err && err_cap) snprintf(err, err_cap, "%s", r.why.c_str());
    return B200S_ENOTREPRESENTABLE;
  } catch (const std::bad_alloc&) {
    return B200S_ENOMEM;
  }
  if (err && err_cap) err[0] = 0;
  *out = E.release();
  return B200S_OK;
}

void b200s_encoded_free(b200s_encoded_t* enc) { delete enc; }
const b200s_catalog_soa_t* b200s_encoded_catalog(const b200s_encoded_t* enc) { return enc ? &enc->cat : nullptr; }
const b200s_templates_t* b200s_encoded_templates(const b200s_encoded_t* enc) { return enc ? &enc->tmpl : nullptr; }
const b200s_pods_soa_t* b200s_encoded_pods(const b200s_encoded_t* enc) { return enc ? &enc->pods : nullptr; }
const b200s_nodes_soa_t* b200s_encoded_nodes(const b200s_encoded_t* enc) { return enc ? &enc->nodes : nullptr; }
int b200s_encoded_key_model_id(const b200s_encoded_t* enc, int device_key) {
  if (!enc || device_key < 0 || device_key >= (int)enc->devkey_model.size()) return -1;
  return enc->devkey_model[device_key];
}
int b200s_encoded_value_model_id(const b200s_encoded_t* enc, int device_key, int bit) {
  if (!enc || device_key < 0 || device_key >= (int)enc->devkey_values.size()) return -1;
  if (bit < 0 || bit >= (int)enc->devkey_values[device_key].size()) return -1;
  return enc->devkey_values[device_key][bit];
}

// ---- config 1: instance.Provider.Create's selection (pkg/providers/instance/instance.go:76-95,
// newAgentPoolObject :321-369); host only ----
int b200s_create_select(const char* name, int32_t n_types, const char* const* instance_types, const double* prices,
                        int64_t storage_bytes, int32_t* chosen, int32_t* disk_gb, int32_t* is_gpu) {
  if (!name || !chosen || !disk_gb || !is_gpu) return B200S_EINVAL;
  // AgentPoolNameRegex ^[a-z][a-z0-9]{0,11}$ (instance.go:50,81)
  size_t n = 0;
  for (; name[n]; n++) {
    char ch = name[n];
    bool lower = ch >= 'a' && ch <= 'z', digit = ch >= '0' && ch <= '9';
    if (n == 0 ? !lower : !(lower || digit)) return B200S_CREATE_EBADNAME;
    if (n >= 12) return B200S_CREATE_EBADNAME;
  }
  if (n == 0) return B200S_CREATE_EBADNAME;
  if (n_types <= 0 || !instance_types) return B200S_CREATE_ENOTYPE;  // instance.go:91-93
  int best = 0;  // instance.go:95 takes element 0; with prices the cheapest wins, first on ties
  if (prices)
    for (int i = 1; i < n_types; i++)
      if (prices[i] < prices[best]) best = i;
  if (storage_bytes <= 0) return B200S_CREATE_ENOSTORAGE;  // instance.go:348-350
  *chosen = best;
  *disk_gb = (int32_t)(storage_bytes >> 30);
  *is_gpu = std::strstr(instance_types[best], "Standard_N") ? 1 : 0;
  return B200S_OK;
}

int b200s_pod_requests(int32_t R, int32_t res_pods, int32_t n_containers, const int64_t* req, const uint8_t* req_has,
                       const int64_t* lim, const uint8_t* lim_has, int32_t n_init, const int64_t* init_req,
                       const uint8_t* init_req_has, const int64_t* init_lim, const uint8_t* init_lim_has,
                       const uint8_t* init_restart_always, const int64_t* overhead, int64_t* out) {
  if (R <= 0 || R > B200S_MAX_RES || res_pods < 0 || res_pods >= R || n_containers < 0 || n_init < 0 || !out) return B200S_EINVAL;
  // all quantities are non-negative, so a dense vector with 0 for "absent" is equivalent to the ResourceList maps
  // under Merge (sum) and MaxResources (max)
  int64_t total[B200S_MAX_RES] = {0}, sidecars[B200S_MAX_RES] = {0}, high[B200S_MAX_RES] = {0};
  auto effective = [&](const int64_t* q, const uint8_t* qh, const int64_t* l, const uint8_t* lh, int i, int r) -> int64_t {
    if (qh[(size_t)i * R + r]) return q[(size_t)i * R + r];          // MergeResourceLimitsIntoRequests (resources.go:204-217)
    if (lh && lh[(size_t)i * R + r]) return l[(size_t)i * R + r];
    return 0;
  };
  for (int i = 0; i < n_containers; i++)
    for (int r = 0; r < R; r++) {
      int64_t v = effective(req, req_has, lim, lim_has, i, r);
      if (v < 0) return B200S_ENOTREPRESENTABLE;
      total[r] += v;
    }
  for (int i = 0; i < n_init; i++) {
    const bool always = init_restart_always && init_restart_always[i];
    for (int r = 0; r < R; r++) {
      int64_t v = effective(init_req, init_req_has, init_lim, init_lim_has, i, r);
      if (v < 0) return B200S_ENOTREPRESENTABLE;
      if (always) {  // resources.go:125-128
        total[r] += v;
        sidecars[r] += v;
        high[r] = std::max(high[r], sidecars[r]);
      } else {       // resources.go:129-132
        high[r] = std::max(high[r], v + sidecars[r]);
      }
    }
  }
  for (int r = 0; r < R; r++) {
    out[r] = std::max(total[r], high[r]);  // resources.go:135
    if (overhead) {
      if (overhead[r] < 0) return B200S_ENOTREPRESENTABLE;
      out[r] += overhead[r];               // resources.go:137-139
    }
  }
  out[res_pods] = 1;  // RequestsForPods (resources.go:36)
  return B200S_OK;
}

int b200s_allocatable(int32_t R, int32_t res_memory, const int64_t* capacity, const uint8_t* capacity_has,
                      const int64_t* kube_reserved, const int64_t* system_reserved, const int64_t* eviction_threshold,
                      const uint8_t* is_hugepages, int64_t* out) {
  if (R <= 0 || R > B200S_MAX_RES || res_memory < 0 || res_memory >= R || !capacity || !capacity_has || !out) return B200S_EINVAL;
  for (int r = 0; r < R; r++) {  // resources.Subtract keeps exactly the keys of the left side (resources.go:81-94)
    int64_t ov = (kube_reserved ? kube_reserved[r] : 0) + (system_reserved ? system_reserved[r] : 0) +
                 (eviction_threshold ? eviction_threshold[r] : 0);
    out[r] = capacity_has[r] ? capacity[r] - ov : 0;
  }
  if (is_hugepages)
    for (int r = 0; r < R; r++)
      if (is_hugepages[r] && capacity_has[r]) {  // types.go:189-198
        out[res_memory] -= capacity[r];
        if (out[res_memory] < 0) out[res_memory] = 0;
      }
  return B200S_OK;
}

}  // extern "C"

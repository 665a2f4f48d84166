// emul.cpp — TEST-ONLY host emulation of the device algorithm.
//
// Runs the arithmetic the kernels run (b200_core.cuh: requirement algebra on bitsets, table lookups,
// eval_word, GoSort) with plain serial control flow, on the SoA the product encoder emits.  It lets
// the CPU test suite (no GPU in the authoring container) check encoder + tables + algebra + sort port
// against the oracle; the lane-cooperative wrappers and the kernels themselves are only exercised by
// the `-m gpu` tests.  Never linked into libb200sched.so and never measured.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../gpu-provisioner_b200/csrc/b200_core.cuh"

using namespace b200;

namespace {

KeyTab keytab_from(const b200s_catalog_soa_t* c) {
  KeyTab kt{};
  kt.n_keys = c->n_keys; kt.n_res = c->n_res; kt.n_words = c->n_words; kt.n_skus = c->n_skus;
  kt.key_zone = c->key_zone; kt.key_ct = c->key_ct; kt.n_zones = c->n_zones; kt.n_cts = c->n_cts;
  kt.wellknown = c->wellknown; kt.n_bounded = c->n_bounded;
  for (int k = 0; k < MAXK; k++) { kt.nvals[k] = 0; kt.bslot[k] = -1; kt.isint[k] = 0; }
  for (int k = 0; k < c->n_keys; k++) { kt.nvals[k] = c->key_nvals[k]; kt.isint[k] = c->val_isint[k]; kt.key_row0[k] = c->key_row0[k]; }
  for (int b = 0; b < c->n_bounded; b++) kt.bslot[c->bounded_key[b]] = (int8_t)b;
  for (int r = 0; r < c->n_res; r++) { kt.res_nd[r] = c->res_nd[r]; kt.res_dv0[r] = c->res_dv0[r]; kt.res_row0[r] = c->res_row0[r]; }
  kt.val_int = c->val_int; kt.rows = c->rows; kt.res_dv = c->res_dv; kt.price = c->price;
  kt.price_max = c->price_max;
  kt.zc_row0 = c->zc_row0; kt.row_nonneg = c->row_nonneg; kt.n_rows = c->n_rows; kt.n_dv = c->n_dv;
  for (int r = 0; r < c->n_res; r++) { kt.cap_nd[r] = c->cap_nd[r]; kt.cap_dv0[r] = c->cap_dv0[r]; kt.cap_row0[r] = c->cap_row0[r]; }
  kt.cap_dv = c->cap_dv;
  return kt;
}

struct ReqRow {  // one requirement set, row-major
  uint64_t mask[MAXK];
  uint64_t flags;
  int64_t gt[MAXB], lt[MAXB];
};

EReq get(const KeyTab& kt, const ReqRow& r, int k) {
  int b = kt.bslot[k];
  return make_ereq(r.mask[k], r.flags, k, b >= 0 ? r.gt[b] : 0, b >= 0 ? r.lt[b] : 0);
}

void put(const KeyTab& kt, ReqRow& r, int k, const EReq& q) {
  r.mask[k] = q.mask;
  r.flags = (r.flags & ~(0xFull << (4 * k))) | ((uint64_t)(q.f & 0xFu) << (4 * k));
  int b = kt.bslot[k];
  if (b >= 0) { r.gt[b] = (q.f & RF_GT) ? q.gt : 0; r.lt[b] = (q.f & RF_LT) ? q.lt : 0; }
}

ReqRow pod_row(const KeyTab& kt, const b200s_pods_soa_t* p, int i) {
  ReqRow r{};
  const int sig = p->sig_id[i], ns = p->n_sigs;
  for (int k = 0; k < kt.n_keys; k++) r.mask[k] = p->sig_mask[(size_t)k * ns + sig];
  r.flags = p->sig_flags[sig];
  for (int b = 0; b < kt.n_bounded; b++) { r.gt[b] = p->sig_gt[(size_t)b * ns + sig]; r.lt[b] = p->sig_lt[(size_t)b * ns + sig]; }
  return r;
}

int64_t pod_request(const b200s_pods_soa_t* p, int r, int i) { return p->rq_requests[(size_t)r * p->n_reqcls + p->req_id[i]]; }
uint64_t pod_tolerated(const b200s_pods_soa_t* p, int i) { return p->sig_tolerated[p->sig_id[i]]; }

ReqRow tmpl_row(const KeyTab& kt, const b200s_templates_t* tm, int t) {
  ReqRow r{};
  for (int k = 0; k < kt.n_keys; k++) r.mask[k] = tm->req_mask[t][k];
  r.flags = tm->req_flags[t];
  for (int b = 0; b < kt.n_bounded; b++) { r.gt[b] = tm->req_gt[t][b]; r.lt[b] = tm->req_lt[t][b]; }
  return r;
}

constexpr int ST_OK = 0, ST_TAINT = 1, ST_INCOMPATIBLE = 2, ST_FILTER = 3;

// serial twin of device_types.cuh::group_can_add
int can_add(const KeyTab& kt, const ReqRow& st, const int64_t* st_req, uint64_t st_taints, const ReqRow& pd,
            const int64_t* pd_req, uint64_t pd_tol, bool check_taints, const uint64_t* cand, uint64_t* out, uint8_t* code,
            int64_t* visited, ReqRow* merged) {
  if (check_taints && (st_taints & ~pd_tol)) return ST_TAINT;
  EvalScratch sc{};
  bool ok = true;
  uint32_t def = 0, forg = 0;
  ReqRow m{};
  for (int k = 0; k < kt.n_keys; k++) {
    EReq s = get(kt, st, k), p = get(kt, pd, k);
    ok = ok && compatible_key(kt, k, s, p, true);
    EReq q = add_req(kt, k, s, p);
    put(kt, m, k, q);
    sc.allowed[k] = 0;
    if (q.f & RF_DEF) {
      sc.allowed[k] = allowed_values(kt, k, q);
      def |= 1u << k;
      if (forgiving(q)) forg |= 1u << k;
    }
  }
  if (!ok) return ST_INCOMPATIBLE;
  for (int r = 0; r < kt.n_res; r++) {
    int64_t x = st_req[r] + pd_req[r];
    int i = fit_rank(kt.res_dv + kt.res_dv0[r], kt.res_nd[r], x);
    sc.fit_row[r] = i < kt.res_nd[r] ? kt.res_row0[r] + i : -1;
  }
  sc.defined = def;
  sc.forgive = forg;
  sc.zone_allowed = (kt.key_zone >= 0 && ((def >> kt.key_zone) & 1u)) ? sc.allowed[kt.key_zone] : ~0ull;
  sc.ct_allowed = (kt.key_ct >= 0 && ((def >> kt.key_ct) & 1u)) ? sc.allowed[kt.key_ct] : ~0ull;
  uint64_t any = 0, aC = 0, aF = 0, aO = 0, aCF = 0, aCO = 0, aFO = 0;
  for (int w = 0; w < kt.n_words; w++) {
    uint64_t c = cand[w];
    *visited += popc64(c);
    WordEval e = eval_word(kt, sc, w);
    out[w] = c & e.compat & e.fits & e.off;
    any |= out[w];
    aC |= c & e.compat; aF |= c & e.fits; aO |= c & e.off;
    aCF |= c & e.compat & e.fits & ~e.off;
    aCO |= c & e.compat & e.off & ~e.fits;
    aFO |= c & e.fits & e.off & ~e.compat;
  }
  if (merged) *merged = m;
  if (any) return ST_OK;
  *code = filter_flags(aC, aF, aO, aCF, aCO, aFO);
  return ST_FILTER;
}

uint64_t tailmask(int S, int w) {
  int W = (S + 63) / 64;
  if (w != W - 1 || (S & 63) == 0) return ~0ull;
  return (1ull << (S & 63)) - 1ull;
}

void prefilter(const KeyTab& kt, const b200s_templates_t* tm, std::vector<uint64_t>* tmask, std::vector<int>* alive) {
  const int W = kt.n_words, T = tm->n_templates;
  tmask->assign((size_t)std::max(T, 1) * W, 0);
  alive->assign(MAXT, 0);
  std::vector<int64_t> zero(MAXR, 0);
  for (int t = 0; t < T; t++) {
    std::vector<uint64_t> cand(W), out(W);
    for (int w = 0; w < W; w++) cand[w] = tailmask(kt.n_skus, w) & (tm->type_admit ? tm->type_admit[(size_t)t * W + w] : ~0ull);
    ReqRow st = tmpl_row(kt, tm, t), pd{};
    uint8_t code = 0;
    int64_t vis = 0;
    int s = can_add(kt, st, zero.data(), 0, pd, zero.data(), 0, false, cand.data(), out.data(), &code, &vis, nullptr);
    if (s == ST_OK) {
      for (int w = 0; w < W; w++) (*tmask)[(size_t)t * W + w] = out[w];
      (*alive)[t] = 1;
    }
  }
}

}  // namespace

extern "C" {

int emul_fit_matrix(const b200s_catalog_soa_t* cat, const b200s_templates_t* tm, const b200s_pods_soa_t* pods,
                    int32_t* tmpl_out, uint64_t* mask_out, uint8_t* code_out, int64_t* evals_out) {
  KeyTab kt = keytab_from(cat);
  const int W = kt.n_words, T = tm->n_templates, P = pods->n_pods, R = kt.n_res;
  std::vector<uint64_t> tmask;
  std::vector<int> alive;
  prefilter(kt, tm, &tmask, &alive);
  int64_t evals = 0;
  std::vector<uint64_t> out(W);
  std::vector<int64_t> preq(R);
  for (int p = 0; p < P; p++) {
    ReqRow pd = pod_row(kt, pods, p);
    for (int r = 0; r < R; r++) preq[r] = pod_request(pods, r, p);
    int chosen = -1;
    for (int w = 0; w < W; w++) mask_out[(size_t)p * W + w] = 0;
    for (int t = 0; t < T; t++) {
      uint8_t code = alive[t] ? B200S_FIT_NONE : B200S_FIT_DROPPED;
      if (alive[t] && chosen < 0) {
        ReqRow st = tmpl_row(kt, tm, t);
        int s = can_add(kt, st, tm->daemon[t], tm->taints[t], pd, preq.data(), pod_tolerated(pods, p), true,
                        &tmask[(size_t)t * W], out.data(), &code, &evals, nullptr);
        if (s == ST_OK) {
          chosen = t;
          code = B200S_FIT_NONE;
          for (int w = 0; w < W; w++) mask_out[(size_t)p * W + w] = out[w];
        } else if (s == ST_TAINT) code = B200S_FIT_TAINT;
        else if (s == ST_INCOMPATIBLE) code = B200S_FIT_INCOMPATIBLE;
      }
      code_out[(size_t)p * T + t] = code;
    }
    tmpl_out[p] = chosen;
  }
  if (evals_out) *evals_out = evals;
  return 0;
}

int emul_go_sort(const int32_t* cnt, int32_t n, int32_t* ord) {
  std::vector<int32_t> key(n);
  for (int i = 0; i < n; i++) key[i] = cnt[ord[i]];
  GoSort gs{ord, key.data()};
  gs.sort(n);
  return 0;
}

// Serial twin of ffd sort + pack_kernel + finalize + select_cheapest.
int emul_solve(const b200s_catalog_soa_t* cat, const b200s_templates_t* tm, const b200s_pods_soa_t* pods,
               const b200s_nodes_soa_t* nodes, b200s_result_t* out) {
  KeyTab kt = keytab_from(cat);
  const int W = kt.n_words, T = tm->n_templates, P = pods->n_pods, R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded;
  const int E = nodes ? nodes->n_nodes : 0;
  std::vector<int32_t> a_tmpl(std::max(P, 1));
  std::vector<uint64_t> a_mask((size_t)std::max(P, 1) * W);
  std::vector<uint8_t> a_code((size_t)std::max(P, 1) * std::max(T, 1));
  int64_t evals = 0;
  emul_fit_matrix(cat, tm, pods, a_tmpl.data(), a_mask.data(), a_code.data(), &evals);
  // FFD order (queue.go:72-108)
  std::vector<int> q(P);
  for (int i = 0; i < P; i++) q[i] = i;
  auto req = [&](int r, int p) { return pod_request(pods, r, p); };
  std::sort(q.begin(), q.end(), [&](int a, int b) {
    if (req(pods->res_cpu, a) != req(pods->res_cpu, b)) return req(pods->res_cpu, a) > req(pods->res_cpu, b);
    if (req(pods->res_memory, a) != req(pods->res_memory, b)) return req(pods->res_memory, a) > req(pods->res_memory, b);
    if (pods->ts[a] != pods->ts[b]) return pods->ts[a] < pods->ts[b];
    if (pods->uid_hi[a] != pods->uid_hi[b]) return pods->uid_hi[a] < pods->uid_hi[b];
    if (pods->uid_lo[a] != pods->uid_lo[b]) return pods->uid_lo[a] < pods->uid_lo[b];
    return a < b;
  });
  struct Claim { ReqRow reqs; std::vector<int64_t> req; std::vector<uint64_t> mask; int tmpl; };
  std::vector<Claim> claims;
  std::vector<int32_t> npods, ord;
  struct Node { ReqRow reqs; std::vector<int64_t> avail; uint64_t taints; };
  std::vector<Node> nd(E);
  for (int e = 0; e < E; e++) {
    nd[e].reqs = ReqRow{};
    for (int k = 0; k < K; k++) nd[e].reqs.mask[k] = nodes->req_mask[(size_t)k * E + e];
    nd[e].reqs.flags = nodes->req_flags[e];
    nd[e].avail.resize(R);
    for (int r = 0; r < R; r++) nd[e].avail[r] = nodes->available[(size_t)r * E + e];
    nd[e].taints = nodes->taints[e];
  }
  std::vector<int> lastlen(P, 0);
  std::vector<int> pod_claim(P, -1);
  size_t head = 0;
  int seq = 0;
  std::vector<int64_t> preq(R);
  std::vector<uint64_t> outw(W);
  for (int p = 0; p < P; p++) { out->pod_place_kind[p] = B200S_PLACE_UNSCHEDULED; out->pod_place_index[p] = -1; out->pod_seq[p] = -1; }
  // NodePool limits (scheduler.go:587-675, :831-867): s.remainingResources per template; with limits the template walk of
  // a new NodeClaim depends on the claims opened so far, so it happens here instead of in stage A
  bool any_limit = false;
  std::vector<std::vector<int64_t>> remaining(std::max(T, 1), std::vector<int64_t>(R, 0));
  for (int t = 0; t < T; t++)
    for (int r = 0; r < R; r++) {
      remaining[t][r] = tm->limit[t][r];
      any_limit = any_limit || tm->has_limit[t][r];
    }
  std::vector<uint64_t> tmask_l;
  std::vector<int> alive_l;
  if (any_limit) prefilter(kt, tm, &tmask_l, &alive_l);
  // filterByRemainingResources as one bit row: instance types whose capacity stays within every limited resource
  auto limit_row = [&](int t, std::vector<uint64_t>& L) {
    L.assign(W, ~0ull);
    for (int r = 0; r < R; r++) {
      if (!tm->has_limit[t][r]) continue;
      const int64_t* cv = kt.cap_dv + kt.cap_dv0[r];
      int idx = -1;
      for (int i = 0; i < kt.cap_nd[r]; i++) if (cv[i] <= remaining[t][r]) idx = i;
      for (int w = 0; w < W; w++) L[w] &= idx < 0 ? 0ull : kt.rows[(size_t)(kt.cap_row0[r] + idx) * W + w];
    }
  };
  for (;;) {
    size_t len = q.size() - head;
    if (len == 0) break;
    int p = q[head];
    if (lastlen[p] == (int)len) break;
    head++;
    ReqRow pd = pod_row(kt, pods, p);
    uint64_t tol = pod_tolerated(pods, p);
    for (int r = 0; r < R; r++) preq[r] = req(r, p);
    bool placed = false;
    for (int e = 0; e < E && !placed; e++) {  // addToExistingNode
      if (nd[e].taints & ~tol) continue;
      bool fits = true;
      for (int r = 0; r < R; r++) if (nd[e].avail[r] < 0 || preq[r] > nd[e].avail[r]) fits = false;
      if (!fits) continue;
      bool ok = true;
      for (int k = 0; k < K; k++) ok = ok && compatible_key(kt, k, get(kt, nd[e].reqs, k), get(kt, pd, k), false);
      if (!ok) continue;
      for (int k = 0; k < K; k++) put(kt, nd[e].reqs, k, add_req(kt, k, get(kt, nd[e].reqs, k), get(kt, pd, k)));
      for (int r = 0; r < R; r++) nd[e].avail[r] -= preq[r];
      out->pod_place_kind[p] = B200S_PLACE_NODE; out->pod_place_index[p] = e; out->pod_seq[p] = seq++;
      placed = true;
    }
    if (placed) continue;
    if (!ord.empty()) {
      std::vector<int32_t> key(ord.size());
      for (size_t i = 0; i < ord.size(); i++) key[i] = npods[ord[i]];
      GoSort gs{ord.data(), key.data()};
      gs.sort((int)ord.size());
    }
    for (size_t i = 0; i < ord.size() && !placed; i++) {  // addToInflightNode
      Claim& c = claims[ord[i]];
      uint8_t code; ReqRow merged;
      int s = can_add(kt, c.reqs, c.req.data(), tm->taints[c.tmpl], pd, preq.data(), tol, true, c.mask.data(), outw.data(), &code, &evals, &merged);
      if (s == ST_OK) {
        c.reqs = merged; c.mask = outw;
        for (int r = 0; r < R; r++) c.req[r] += preq[r];
        npods[ord[i]]++;
        pod_claim[p] = ord[i];
        out->pod_place_kind[p] = B200S_PLACE_CLAIM; out->pod_seq[p] = seq++;
        placed = true;
      }
    }
    if (placed) continue;
    int t = a_tmpl[p];
    std::vector<uint64_t> new_row(a_mask.begin() + (size_t)p * W, a_mask.begin() + (size_t)(p + 1) * W);
    if (any_limit) {  // the limit-aware walk; its codes replace stage A's for this pod
      t = -1;
      std::vector<uint64_t> L, cand(W), got(W);
      for (int i = 0; i < T; i++) {
        uint8_t code = alive_l[i] ? B200S_FIT_NONE : B200S_FIT_DROPPED;
        if (alive_l[i] && t < 0) {
          bool lim = false;
          for (int r = 0; r < R; r++) lim = lim || tm->has_limit[i][r];
          bool empty = false;
          for (int w = 0; w < W; w++) cand[w] = tmask_l[(size_t)i * W + w];
          if (lim) {
            limit_row(i, L);
            uint64_t any = 0;
            for (int w = 0; w < W; w++) { cand[w] &= L[w]; any |= cand[w]; }
            empty = any == 0;
          }
          if (empty) {
            code = B200S_FIT_LIMITS;
          } else {
            ReqRow st = tmpl_row(kt, tm, i);
            int s = can_add(kt, st, tm->daemon[i], tm->taints[i], pd, preq.data(), tol, true, cand.data(), got.data(), &code, &evals, nullptr);
            if (s == ST_OK) { t = i; code = B200S_FIT_NONE; new_row = got; }
            else if (s == ST_TAINT) code = B200S_FIT_TAINT;
            else if (s == ST_INCOMPATIBLE) code = B200S_FIT_INCOMPATIBLE;
          }
        }
        a_code[(size_t)p * T + i] = code;
      }
    }
    if (t >= 0) {  // addToNewNodeClaim
      Claim c;
      ReqRow st = tmpl_row(kt, tm, t);
      c.reqs = ReqRow{};
      for (int k = 0; k < K; k++) put(kt, c.reqs, k, add_req(kt, k, get(kt, st, k), get(kt, pd, k)));
      c.req.resize(R);
      for (int r = 0; r < R; r++) c.req[r] = tm->daemon[t][r] + preq[r];
      c.mask = new_row;
      c.tmpl = t;
      claims.push_back(c);
      npods.push_back(1);
      ord.push_back((int)claims.size() - 1);
      pod_claim[p] = (int)claims.size() - 1;
      out->pod_place_kind[p] = B200S_PLACE_CLAIM; out->pod_seq[p] = seq++;
      for (int r = 0; r < R; r++) {  // subtractMax (scheduler.go:831-848): the largest capacity among the claim's types
        if (!tm->has_limit[t][r]) continue;
        const int64_t* cv = kt.cap_dv + kt.cap_dv0[r];
        int64_t mx = 0;
        for (int i = kt.cap_nd[r] - 1; i >= 0; i--) {  // smallest row "capacity <= cv[i]" that holds every type
          bool sub = true;
          for (int w = 0; w < W; w++) sub = sub && !(new_row[w] & ~kt.rows[(size_t)(kt.cap_row0[r] + i) * W + w]);
          if (!sub) break;
          mx = cv[i];
        }
        remaining[t][r] -= mx;
      }
    } else {
      q.push_back(p);
      lastlen[p] = (int)(q.size() - head);
      out->pod_place_kind[p] = B200S_PLACE_UNSCHEDULED;
    }
  }
  int n = (int)claims.size();
  if (n > out->cap_claims) return B200S_ECAPACITY;
  out->n_claims = n;
  out->n_unscheduled = (int)(q.size() - head);
  out->fit_evals = evals;
  std::vector<int> pos(n);
  for (int i = 0; i < n; i++) pos[ord[i]] = i;
  for (int p = 0; p < P; p++) {
    if (out->pod_place_kind[p] == B200S_PLACE_CLAIM) out->pod_place_index[p] = pos[pod_claim[p]];
    for (int t = 0; t < T; t++) {
      uint8_t c = a_code[(size_t)p * T + t];
      out->pod_fit_code[(size_t)p * T + t] = (out->pod_place_kind[p] == B200S_PLACE_UNSCHEDULED || c == B200S_FIT_DROPPED) ? c : (uint8_t)B200S_FIT_NONE;
    }
  }
  const int ZC = kt.n_zones * kt.n_cts;
  for (int i = 0; i < n; i++) {
    const Claim& c = claims[ord[i]];
    out->claim_template[i] = c.tmpl;
    out->claim_npods[i] = npods[ord[i]];
    for (int r = 0; r < R; r++) out->claim_requests[(size_t)i * R + r] = c.req[r];
    if (out->claim_mask) for (int w = 0; w < W; w++) out->claim_mask[(size_t)i * W + w] = c.mask[w];
    for (int k = 0; k < K; k++) out->claim_req_mask[(size_t)i * K + k] = c.reqs.mask[k];
    out->claim_req_flags[i] = c.reqs.flags;
    for (int b = 0; b < NB; b++) { out->claim_req_gt[(size_t)i * NB + b] = c.reqs.gt[b]; out->claim_req_lt[(size_t)i * NB + b] = c.reqs.lt[b]; }
    // OrderByPrice head (types.go:221-240)
    uint64_t za = ~0ull, ca = ~0ull;
    if (kt.key_zone >= 0) { EReq qz = get(kt, c.reqs, kt.key_zone); if (qz.f & RF_DEF) za = allowed_values(kt, kt.key_zone, qz); }
    if (kt.key_ct >= 0) { EReq qc = get(kt, c.reqs, kt.key_ct); if (qc.f & RF_DEF) ca = allowed_values(kt, kt.key_ct, qc); }
    uint64_t best = ~0ull; int bs = -1, cnt = 0;
    for (int s = 0; s < kt.n_skus; s++) {
      if (!((c.mask[s / 64] >> (s % 64)) & 1ull)) continue;
      cnt++;
      uint64_t pr = B200S_PRICE_NONE;
      for (int zc = 0; zc < ZC; zc++) {
        int z = zc / kt.n_cts, ct = zc % kt.n_cts;
        if (!((za >> z) & 1ull) || !((ca >> ct) & 1ull)) continue;
        pr = std::min(pr, kt.price[(size_t)zc * kt.n_skus + s]);
      }
      if (pr < best) { best = pr; bs = s; }
    }
    out->claim_sku[i] = bs;
    out->claim_price[i] = bs >= 0 ? best : B200S_PRICE_NONE;
    out->claim_n_options[i] = cnt;
  }
  return 0;
}

// fit_rank directly and through the log-bucket table the catalog upload builds, on the same queries
// the two price keys the consolidation and price-order kernels compute per instance type (b200_core.cuh), for every SKU
void emul_price_keys(const b200s_catalog_soa_t* cat, uint64_t za, uint64_t ca, int32_t ct_reserved, int32_t ct_spot,
                     int32_t ct_on_demand, uint64_t* cheapest, uint64_t* worst) {
  KeyTab kt = keytab_from(cat);
  for (int s = 0; s < kt.n_skus; s++) {
    cheapest[s] = cheapest_price(kt, za, ca, s);
    worst[s] = worst_launch_price(kt, za, ca, ct_reserved, ct_spot, ct_on_demand, s);
  }
}

// values of key k a requirement row admits, as the kernels compute them (undefined keys admit everything)
uint64_t emul_allowed_values(const b200s_catalog_soa_t* cat, int32_t k, uint64_t mask, uint64_t flags_word, int64_t gt,
                             int64_t lt) {
  KeyTab kt = keytab_from(cat);
  EReq q = make_ereq(mask, flags_word, k, gt, lt);
  return (q.f & RF_DEF) ? allowed_values(kt, k, q) : ~0ull;
}

void emul_fit_rank_both(const int64_t* dv, int32_t nd, const int64_t* xs, int32_t n, int32_t* plain, int32_t* bucketed) {
  std::vector<uint16_t> row(RANK_BKTS + 1);
  rank_bucket_table(dv, nd, row.data());
  for (int i = 0; i < n; i++) {
    plain[i] = fit_rank(dv, nd, xs[i]);
    bucketed[i] = fit_rank_bkt(dv, row.data(), xs[i]);
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Serial twin of pack_reg_kernel (pack_reg_kernels.cu): the register-resident form of the pack for batches
// without pod selectors and without existing nodes.  What differs from emul_solve, and is therefore checked here
// against the oracle on the CPU: (1) the claim order is never read by the scan -- every claim carries a priority
// (len(Pods), tie-break) such that "first feasible claim in s.newNodeClaims order" == "feasible claim of least
// priority", updated locally on a win (front of the next bucket) or a new claim (back of bucket 1) and re-derived
// from the explicit order only after a full pdqsort; (2) the in-flight check is the threshold / upper-bound /
// row-AND form with the claim's types recomputed only when a resource moves to a new table row.
namespace {

struct OrderEmu {  // the order warp's arrays and its eager normalisation with undo (pack_kernels.cu warp_normalise)
  std::vector<int32_t> ord, key, pos_of, snap;
  int undo_mode = 0, undo_src = 0, undo_dst = 0;  // 0 none, 1 rotate, 2 generic
  bool generic = false;
  void move(int src, int dst) {
    int x = ord[src], kx = key[src];
    if (src < dst) for (int i = src; i < dst; i++) { ord[i] = ord[i + 1]; key[i] = key[i + 1]; pos_of[ord[i]] = i; }
    else for (int i = src; i > dst; i--) { ord[i] = ord[i - 1]; key[i] = key[i - 1]; pos_of[ord[i]] = i; }
    ord[dst] = x; key[dst] = kx; pos_of[x] = dst;
  }
  void normalise(int pos, bool appended) {
    const int n = (int)ord.size();
    generic = false;
    undo_mode = 0;
    if (n <= 1) return;
    const bool sorted = appended ? !(key[n - 1] < key[n - 2]) : (pos == n - 1 || !(key[pos + 1] < key[pos]));
    if (sorted) return;
    bool increasing = false;
    if (n >= 50) {
      const int q = n / 4;
      int v[9];
      for (int l = 0; l < 9; l++) v[l] = key[q * (l / 3 + 1) + (l % 3) - 1];
      bool bad = false;
      for (int l = 0; l < 9; l++) {
        if ((l % 3) != 2 && v[l + 1] < v[l]) bad = true;
        if ((l == 1 || l == 4) && v[l + 3] < v[l]) bad = true;
      }
      increasing = !bad;
    }
    if (!increasing) {
      snap = ord;
      GoSort gs{ord.data(), key.data()};
      gs.sort(n);
      for (int i = 0; i < n; i++) pos_of[ord[i]] = i;
      generic = true;
      undo_mode = 2;
      return;
    }
    const int kx = key[pos];
    int dst = pos;
    if (!appended) { while (dst + 1 < n && key[dst + 1] < kx) dst++; }
    else { while (dst - 1 >= 0 && kx < key[dst - 1]) dst--; }
    move(pos, dst);
    undo_mode = dst != pos ? 1 : 0;
    undo_src = dst;
    undo_dst = pos;
  }
  void undo() {
    if (undo_mode == 1) move(undo_src, undo_dst);
    else if (undo_mode == 2) { ord = snap; }
  }
};

}  // namespace

extern "C" int emul_solve_reg(const b200s_catalog_soa_t* cat, const b200s_templates_t* tm, const b200s_pods_soa_t* pods,
                              b200s_result_t* out, int64_t* stats, int32_t kcap) {
  KeyTab kt = keytab_from(cat);
  const int W = kt.n_words, T = tm->n_templates, P = pods->n_pods, R = kt.n_res, K = kt.n_keys, NB = kt.n_bounded;
  for (int s = 0; s < pods->n_sigs; s++)
    if (pods->sig_flags[s] & 0x1111111111111111ull) return B200S_EINVAL;  // a pod defines a requirement key
  std::vector<int32_t> a_tmpl(std::max(P, 1));
  std::vector<uint64_t> a_mask((size_t)std::max(P, 1) * W);
  std::vector<uint8_t> a_code((size_t)std::max(P, 1) * std::max(T, 1));
  int64_t evals = 0;
  emul_fit_matrix(cat, tm, pods, a_tmpl.data(), a_mask.data(), a_code.data(), &evals);
  std::vector<int> q(P);
  for (int i = 0; i < P; i++) q[i] = i;
  auto req = [&](int r, int p) { return pod_request(pods, r, p); };
  std::sort(q.begin(), q.end(), [&](int a, int b) {
    if (req(pods->res_cpu, a) != req(pods->res_cpu, b)) return req(pods->res_cpu, a) > req(pods->res_cpu, b);
    if (req(pods->res_memory, a) != req(pods->res_memory, b)) return req(pods->res_memory, a) > req(pods->res_memory, b);
    if (pods->ts[a] != pods->ts[b]) return pods->ts[a] < pods->ts[b];
    if (pods->uid_hi[a] != pods->uid_hi[b]) return pods->uid_hi[a] < pods->uid_hi[b];
    if (pods->uid_lo[a] != pods->uid_lo[b]) return pods->uid_lo[a] < pods->uid_lo[b];
    return a < b;
  });
  // feasibility table per template (pack_reg_kernels.cu feas_table_kernel): bit [i_0, .., i_{R-1}] = some instance
  // type of the template's pre-filtered set has Allocatable()[r] >= dv_r[i_r] for every r.  A claim of selector-less
  // pods keeps exactly the template's types that fit its requests, so CanAdd's filter result is non-empty iff the
  // cell of the new total's ranks is set (nodeclaim.go:144-146, :392-423).
  std::vector<uint64_t> tmask;
  std::vector<int> alive;
  prefilter(kt, tm, &tmask, &alive);
  int64_t stride[MAXR], cells = 1;
  for (int r = 0; r < R; r++) { stride[r] = cells; cells *= std::max(kt.res_nd[r], 1); }
  std::vector<std::vector<uint8_t>> table(std::max(T, 1), std::vector<uint8_t>((size_t)cells, 0));
  for (int t = 0; t < T; t++)
    for (int64_t cell = 0; cell < cells; cell++) {
      bool any = false;
      for (int w = 0; w < W && !any; w++) {
        uint64_t m = tmask[(size_t)t * W + w] & kt.rows[(size_t)kt.row_nonneg * W + w];
        for (int r = 0; r < R; r++) {
          int i = (int)((cell / stride[r]) % std::max(kt.res_nd[r], 1));
          m &= kt.rows[(size_t)(kt.res_row0[r] + i) * W + w];
        }
        any = m != 0;
      }
      table[t][(size_t)cell] = any;
    }
  // The order, as pack_reg_kernel keeps it: every claim knows its POSITION in s.newNodeClaims; a step's change of the
  // order is one range of positions moving by one (the shift word), applied by everyone one step late; the minimum of a
  // step is taken over priorities 2 * pos + 1 that do not have the previous step's shift yet, in which the claim that
  // changed took a value between its new neighbours' old ones; only when sort.Slice could take the real pdqsort path is
  // the explicit order laid out and sorted (OrderEmu::normalise), lazily, at the start of the NEXT add() — which is when
  // the reference sorts, so the last step's change is simply never applied.
  struct Claim { std::vector<int64_t> req, thr; std::vector<int> rk; int tmpl; uint32_t pos, prio; int npods; int first; };
  std::vector<Claim> claims;
  OrderEmu oe;
  const int KCAP = kcap > 0 ? kcap : 2048;  // pod counts with an entry in the count table G[x] = claims with more than x pods
  std::vector<int> G(KCAP, 0);
  struct Shift { int lo = 1, hi = 0; bool up = false, meet = false; int meet_a = 0, meet_n = 0; bool meet_app = false; } rec;
  const uint32_t MARK = 0x80000000u;
  std::vector<int> lastlen(P, 0), pod_claim(P, -1);
  size_t head = 0;
  int seq = 0;
  std::vector<int64_t> preq(R), x(R);
  int64_t n_generic = 0, n_resolve = 0, n_steps = 0;
  for (int p = 0; p < P; p++) { out->pod_place_kind[p] = B200S_PLACE_UNSCHEDULED; out->pod_place_index[p] = -1; out->pod_seq[p] = -1; }
  int64_t n_meets = 0;
  // the sort at the start of an add(): the previous step's shift, or — when it asked for it — the real sort.Slice
  auto start_of_add = [&]() {
    const int n = (int)claims.size();
    if (rec.meet) {
      n_meets++;
      oe.ord.assign(n, 0); oe.key.assign(n, 0); oe.pos_of.assign(n, 0);
      for (int c = 0; c < n; c++) {  // the order as it is: positions before the change, the new pod count
        oe.ord[claims[c].pos] = c; oe.key[claims[c].pos] = claims[c].npods; oe.pos_of[c] = (int)claims[c].pos;
      }
      oe.normalise(rec.meet_a, rec.meet_app);
      if (oe.generic) n_generic++;
      for (int c = 0; c < n; c++) claims[c].pos = (uint32_t)oe.pos_of[c];
    } else {
      for (int c = 0; c < n; c++) {
        uint32_t& ps = claims[c].pos;
        if (!(ps & MARK) && (int)ps >= rec.lo && (int)ps <= rec.hi) ps += rec.up ? 1u : 0xFFFFFFFFu;
        ps &= ~MARK;  // the claim that moved had marked where it went
      }
    }
    for (int c = 0; c < n; c++) claims[c].prio = 2u * claims[c].pos + 1u;
    rec = Shift{};
  };
  auto positions_are_a_permutation = [&]() -> bool {
    const int n = (int)claims.size();
    std::vector<char> seen(n, 0);
    for (int c = 0; c < n; c++) {
      const uint32_t ps = claims[c].pos & ~MARK;
      if (ps >= (uint32_t)n || seen[ps]) return false;
      seen[ps] = 1;
    }
    return true;
  };
  for (;;) {
    size_t len = q.size() - head;
    if (len == 0) break;
    int p = q[head];
    if (lastlen[p] == (int)len) break;
    head++;
    n_steps++;
    const uint64_t tol = pod_tolerated(pods, p);
    for (int r = 0; r < R; r++) preq[r] = req(r, p);
    // in-flight: every claim classified independently, least priority among the feasible ones wins
    int win = -1;
    std::vector<int> win_rk, nrk(R);
    bool win_moved = false;
    std::vector<char> feasible(claims.size(), 0), is_moved(claims.size(), 0);
    std::vector<std::vector<int>> moved_rk(claims.size());
    for (int c = 0; c < (int)claims.size(); c++) {
      Claim& cl = claims[c];
      if (tm->taints[cl.tmpl] & ~tol) continue;
      bool moved = false;
      for (int r = 0; r < R; r++) {
        x[r] = cl.req[r] + preq[r];
        if (preq[r] != 0 && x[r] > cl.thr[r]) moved = true;  // a resource the pod does not request keeps its row
      }
      if (moved) {
        n_resolve++;
        bool ok = true;
        int64_t cell = 0;
        for (int r = 0; r < R && ok; r++) {
          int i = cl.rk[r];
          if (preq[r] != 0 && x[r] > cl.thr[r]) {
            const int64_t* dv = kt.res_dv + kt.res_dv0[r];
            i++;
            while (i < kt.res_nd[r] && dv[i] < x[r]) i++;
            if (i >= kt.res_nd[r]) { ok = false; break; }
          }
          nrk[r] = i;
          cell += (int64_t)i * stride[r];
        }
        if (!ok || !table[cl.tmpl][(size_t)cell]) continue;
      }
      feasible[c] = 1;
      if (moved) moved_rk[c] = nrk;
      is_moved[c] = moved;
    }
    // the kernel takes the minimum before it knows the previous step's shift; when that step asked for the real sort
    // the minimum is taken again afterwards over the re-read positions
    auto least = [&]() {
      int w = -1;
      for (int c = 0; c < (int)claims.size(); c++)
        if (feasible[c] && (w < 0 || claims[c].prio < claims[w].prio)) w = c;
      return w;
    };
    const bool resorted = rec.meet;
    win = least();
    start_of_add();
    if (resorted) win = least();
    else if (win >= 0 && least() != win) return -3000;  // the stale priorities and the true positions must agree
    if (!positions_are_a_permutation()) return -3001;
    if (win >= 0) {
      win_moved = is_moved[win];
      if (win_moved) win_rk = moved_rk[win];
      Claim& cl = claims[win];
      if (win_moved) {
        cl.rk = win_rk;
        for (int r = 0; r < R; r++) cl.thr[r] = kt.res_dv[kt.res_dv0[r] + cl.rk[r]];
      }
      for (int r = 0; r < R; r++) cl.req[r] += preq[r];
      // the claim now has k + 1 pods and belongs at the end of the run of claims with k pods (pack_reg_kernel's winner)
      const int n = (int)claims.size(), a = (int)cl.pos, k = cl.npods;
      const bool big = k + 1 >= KCAP;
      const int kk = big ? KCAP - 1 : k;
      const int gk = G[kk];
      if (!big) G[kk] = gk + 1;
      int dst = n - gk - 1;
      const int qq = n >> 2;
      const bool near = (unsigned)(a - qq + 1) <= 2u || (unsigned)(a - 2 * qq + 1) <= 2u || (unsigned)(a - 3 * qq + 1) <= 2u;
      const bool meet = big || (dst != a && (n < 50 || near));
      if (big) dst = a;
      rec.lo = a + 1; rec.hi = dst; rec.up = false; rec.meet = meet;
      rec.meet_a = a; rec.meet_app = false; rec.meet_n = n;
      cl.npods = k + 1;
      cl.prio = 2u * (uint32_t)dst + 2u;
      cl.pos = meet ? (uint32_t)a : (MARK | (uint32_t)dst);
      pod_claim[p] = win;
      out->pod_place_kind[p] = B200S_PLACE_CLAIM; out->pod_seq[p] = seq++;
      continue;
    }
    int t = a_tmpl[p];
    if (t >= 0) {
      Claim c;
      c.req.resize(R); c.thr.resize(R); c.rk.resize(R);
      for (int r = 0; r < R; r++) {
        c.req[r] = tm->daemon[t][r] + preq[r];
        c.rk[r] = fit_rank(kt.res_dv + kt.res_dv0[r], kt.res_nd[r], c.req[r]);
        c.thr[r] = kt.res_dv[kt.res_dv0[r] + c.rk[r]];
      }
      c.first = p;
      c.tmpl = t;
      c.npods = 1;
      // appended (scheduler.go:672); the next sort puts it behind the last claim with one pod
      const int n = (int)claims.size();
      const int dst = n - (KCAP > 1 ? G[1] : 0);
      const bool meet = (n + 1 < 50 && dst != n) || KCAP <= 1;
      c.pos = meet ? (uint32_t)n : (MARK | (uint32_t)dst);
      c.prio = 2u * (uint32_t)dst;
      rec.lo = dst; rec.hi = n - 1; rec.up = true; rec.meet = meet;
      rec.meet_a = n; rec.meet_app = true; rec.meet_n = n + 1;
      claims.push_back(c);
      const int cid = (int)claims.size() - 1;
      pod_claim[p] = cid;
      out->pod_place_kind[p] = B200S_PLACE_CLAIM; out->pod_seq[p] = seq++;
    } else {
      q.push_back(p);
      lastlen[p] = (int)(q.size() - head);
      out->pod_place_kind[p] = B200S_PLACE_UNSCHEDULED;
    }
  }
  // Results.NewNodeClaims: the order after the last sort the reference ran plus what the last add() did — the last
  // step's change is not applied: its claim keeps its place (one before the range it would have passed), a new claim is last
  const int n = (int)claims.size();
  for (int c = 0; c < n; c++) {
    uint32_t& ps = claims[c].pos;
    if (ps & MARK) ps = rec.up ? (uint32_t)(n - 1) : (rec.lo <= rec.hi ? (uint32_t)(rec.lo - 1) : (ps & ~MARK));
  }
  if (!positions_are_a_permutation()) return -3002;
  oe.ord.assign(n, 0);
  for (int c = 0; c < n; c++) oe.ord[claims[c].pos] = c;
  if (stats) { stats[0] = n_steps; stats[1] = n_generic; stats[2] = n_resolve; stats[3] = n_meets; }
  if (n > out->cap_claims) return B200S_ECAPACITY;
  out->n_claims = n;
  out->n_unscheduled = (int)(q.size() - head);
  out->fit_evals = evals;
  std::vector<int> pos(n);
  for (int i = 0; i < n; i++) pos[oe.ord[i]] = i;
  for (int p = 0; p < P; p++) {
    if (out->pod_place_kind[p] == B200S_PLACE_CLAIM) out->pod_place_index[p] = pos[pod_claim[p]];
    for (int t = 0; t < T; t++) {
      uint8_t c = a_code[(size_t)p * T + t];
      out->pod_fit_code[(size_t)p * T + t] = (out->pod_place_kind[p] == B200S_PLACE_UNSCHEDULED || c == B200S_FIT_DROPPED) ? c : (uint8_t)B200S_FIT_NONE;
    }
  }
  const int ZC = kt.n_zones * kt.n_cts;
  for (int i = 0; i < n; i++) {
    const Claim& c = claims[oe.ord[i]];
    // InstanceTypeOptions: the first pod's stage-A row inside the rows of the final totals (the rows are nested, so
    // the AND of every row the claim ever moved to is the last one)
    std::vector<uint64_t> cmask(W);
    for (int w = 0; w < W; w++) {
      uint64_t m = a_mask[(size_t)c.first * W + w];
      for (int r = 0; r < R; r++) m &= kt.rows[(size_t)(kt.res_row0[r] + c.rk[r]) * W + w];
      cmask[w] = m;
    }
    out->claim_template[i] = c.tmpl;
    out->claim_npods[i] = c.npods;
    for (int r = 0; r < R; r++) out->claim_requests[(size_t)i * R + r] = c.req[r];
    if (out->claim_mask) for (int w = 0; w < W; w++) out->claim_mask[(size_t)i * W + w] = cmask[w];
    for (int k = 0; k < K; k++) out->claim_req_mask[(size_t)i * K + k] = tm->req_mask[c.tmpl][k];
    out->claim_req_flags[i] = tm->req_flags[c.tmpl];
    for (int b = 0; b < NB; b++) {
      const uint32_t nib = (uint32_t)((tm->req_flags[c.tmpl] >> (4 * cat->bounded_key[b])) & 0xFu);
      out->claim_req_gt[(size_t)i * NB + b] = (nib & RF_GT) ? tm->req_gt[c.tmpl][b] : 0;
      out->claim_req_lt[(size_t)i * NB + b] = (nib & RF_LT) ? tm->req_lt[c.tmpl][b] : 0;
    }
    ReqRow rr = tmpl_row(kt, tm, c.tmpl);
    uint64_t za = ~0ull, ca = ~0ull;
    if (kt.key_zone >= 0) { EReq qz = get(kt, rr, kt.key_zone); if (qz.f & RF_DEF) za = allowed_values(kt, kt.key_zone, qz); }
    if (kt.key_ct >= 0) { EReq qc = get(kt, rr, kt.key_ct); if (qc.f & RF_DEF) ca = allowed_values(kt, kt.key_ct, qc); }
    uint64_t best = ~0ull; int bs = -1, cnt = 0;
    for (int s = 0; s < kt.n_skus; s++) {
      if (!((cmask[s / 64] >> (s % 64)) & 1ull)) continue;
      cnt++;
      uint64_t pr = B200S_PRICE_NONE;
      for (int zc = 0; zc < ZC; zc++) {
        int z = zc / kt.n_cts, ct = zc % kt.n_cts;
        if (!((za >> z) & 1ull) || !((ca >> ct) & 1ull)) continue;
        pr = std::min(pr, kt.price[(size_t)zc * kt.n_skus + s]);
      }
      if (pr < best) { best = pr; bs = s; }
    }
    out->claim_sku[i] = bs;
    out->claim_price[i] = bs >= 0 ? best : B200S_PRICE_NONE;
    out->claim_n_options[i] = cnt;
  }
  return 0;
}

// api.cu — the C-ABI of libb200sched (include/b200sched.h): contexts, catalog upload, and the
// host-buffer / resident entry points that enqueue the kernels of fit_kernels.cu and pack_kernels.cu
// on the context's stream.  No CPU fallback lives here: without a usable sm_100 device b200s_init
// fails and every other entry point needs a context.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"

using namespace b200;

namespace {

struct Block {
  void* ptr;
  size_t size;
  bool in_use;
};

struct Catalog {
  KeyTab kt{};  // device pointers
  std::vector<void*> bufs;
  uint64_t layout_hash = 0;   // b200s_catalog_soa_t::layout_hash of the uploaded encode
  std::vector<int64_t> h_dv;  // host copy of the distinct allocatable values (the register-resident pack's plan reads them)
  int S = 0, W = 0, R = 0, K = 0, NB = 0;
};

}  // namespace

struct b200s_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n_sm = 0;
  int smem_limit = 0;
  std::map<uint64_t, std::unique_ptr<Catalog>> catalogs;
  uint64_t next_id = 1;
  std::string err;
  std::vector<Block> cache;  // grow-only device block cache so repeated calls do not hit cudaMalloc

  void* alloc(size_t bytes) {
    if (bytes == 0) bytes = 8;
    bytes = (bytes + 255) & ~(size_t)255;
    Block* best = nullptr;
    for (auto& b : cache)
      if (!b.in_use && b.size >= bytes && b.size <= bytes * 2 + 4096 && (!best || b.size < best->size)) best = &b;
    if (best) {
      best->in_use = true;
      return best->ptr;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    cache.push_back(Block{p, bytes, true});
    return p;
  }
  void release(void* p) {
    for (auto& b : cache)
      if (b.ptr == p) b.in_use = false;
  }
};

struct b200s_resident {
  b200s_ctx* ctx = nullptr;
  Catalog* cat = nullptr;
  int P = 0, Ppad = 0, E = 0, T = 0, n_problems = 0;
  bool profile_pack = false;  // b200s_resident_set_pack_profile
  bool any_limit = false;  // some NodePool limits a resource: the new-claim template walk happens in the pack
  bool plain = true;  // no pod defines a requirement key (selects the pack kernel's form together with E == 0)
  std::vector<void*> owned;
  PodsDev pods{};
  int32_t* seg = nullptr;          // problem of each pod (batch only)
  int32_t* row_problem = nullptr;  // problem of each claim row (batch only)
  TmplDev* tmpl = nullptr;
  uint64_t* type_admit = nullptr;
  uint64_t* tmask = nullptr;
  int32_t* alive = nullptr;
  // nodes: pristine copies + the mutable state the pack works on
  int64_t* n_avail0 = nullptr;
  uint64_t* n_rmask0 = nullptr;
  uint64_t* n_rflags0 = nullptr;
  uint64_t* n_taints = nullptr;
  // class tables (per signature x template, per request vector x template)
  uint8_t* sg_status = nullptr;
  int32_t* rq_rank = nullptr;   // [NQ] FFD rank of a request class (launch_ffd_class_rank)
  uint64_t* ffd_k1 = nullptr;   // [P] leading sort key of a pod
  uint64_t* sgC = nullptr;
  uint64_t* sgO = nullptr;
  uint64_t* sgCO = nullptr;
  uint64_t* rqF = nullptr;
  int32_t* tm_count = nullptr;
  // stage A
  int32_t* a_tmpl = nullptr;
  uint64_t* a_mask = nullptr;
  uint8_t* a_code = nullptr;
  unsigned long long* a_evals = nullptr;
  // pack
  PackArrays arr{};
  PackProblem* problems = nullptr;
  std::vector<PackProblem> problems_host;
  ClaimOut cout{};
  int32_t* sel_sku = nullptr;
  uint64_t* sel_price = nullptr;
  int32_t* sel_nopt = nullptr;
  RegPackPlan reg_plan{};      // register-resident pack: table shape (eligible only without pod keys / nodes)
  uint32_t* feas_bits = nullptr;
  RegPackDev reg_dev{};
  int last_launches = 0;
  bool fit_done = false, solve_done = false, rows_done = false, imported = false;
  cudaEvent_t ev_dense0 = nullptr, ev_dense1 = nullptr;  // bracket the dense kernel of the last stage A on ctx->stream

  template <class Tp>
  Tp* dalloc(size_t n) {
    void* p = ctx->alloc(n * sizeof(Tp));
    if (p) owned.push_back(p);
    return reinterpret_cast<Tp*>(p);
  }
};

namespace {

int fail(b200s_ctx* ctx, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (ctx) {
    ctx->err = what;
    if (e != cudaSuccess) {
      ctx->err += ": ";
      ctx->err += cudaGetErrorString(e);
    }
  }
  cudaGetLastError();
  return code;
}

#define CUDA_TRY(ctx, expr, what)                                     \
  do {                                                                \
    cudaError_t e_ = (expr);                                          \
    if (e_ != cudaSuccess) return fail(ctx, B200S_ECUDA, what, e_);   \
  } while (0)

template <class Tp>
int upload(b200s_ctx* ctx, Tp* dst, const Tp* src, size_t n) {
  if (n == 0) return 0;
  if (!dst || !src) return fail(ctx, B200S_EINVAL, "null buffer");
  CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, n * sizeof(Tp), cudaMemcpyHostToDevice, ctx->stream), "H2D copy");
  return 0;
}

// instance-type sets handed in by the caller: no bit at or above n_skus (the kernels index price rows by bit position)
bool masks_in_range(const uint64_t* masks, size_t n, int S, int W) {
  if ((S & 63) == 0) return true;
  const uint64_t tail = ~((1ull << (S & 63)) - 1ull);
  for (size_t i = 0; i < n; i++)
    if (masks[i * W + (W - 1)] & tail) return false;
  return true;
}

int check_shapes(b200s_ctx* ctx, const Catalog* cat, const b200s_pods_soa_t* pods, const b200s_nodes_soa_t* nodes) {
  if (!pods) return fail(ctx, B200S_EINVAL, "pods is null");
  if (cat->layout_hash && pods->layout_hash && cat->layout_hash != pods->layout_hash)
    return fail(ctx, B200S_ESTALE, "pods were encoded against another key layout than the uploaded catalog: upload the catalog of the same b200s_encode");
  if (nodes && nodes->n_nodes > 0 && cat->layout_hash && nodes->layout_hash && cat->layout_hash != nodes->layout_hash)
    return fail(ctx, B200S_ESTALE, "nodes were encoded against another key layout than the uploaded catalog");
  if (pods->n_res != cat->R || pods->n_keys != cat->K || pods->n_bounded != cat->NB)
    return fail(ctx, B200S_EINVAL, "pods SoA does not match the catalog's R/K/NB");
  if (pods->n_pods > 0 && (pods->n_sigs <= 0 || pods->n_reqcls <= 0)) return fail(ctx, B200S_EINVAL, "pods without class tables");
  if (pods->res_cpu < 0 || pods->res_cpu >= cat->R || pods->res_memory < 0 || pods->res_memory >= cat->R)
    return fail(ctx, B200S_EINVAL, "pods res_cpu/res_memory out of range");
  if (nodes && nodes->n_nodes > 0 && (nodes->n_res != cat->R || nodes->n_keys != cat->K))
    return fail(ctx, B200S_EINVAL, "nodes SoA does not match the catalog's R/K");
  return 0;
}

void free_resident(b200s_resident* r) {
  if (!r) return;
  for (void* p : r->owned) r->ctx->release(p);
  if (r->ev_dense0) cudaEventDestroy(r->ev_dense0);
  if (r->ev_dense1) cudaEventDestroy(r->ev_dense1);
  delete r;
}

// b200s_solve_batch_shared: the problems are ranges of ONE pod set (pods_in[0]) and share ONE node set, of which
// problem b removes the rows excl_rows[excl_off[b] .. excl_off[b+1])
struct SharedBatch {
  const int32_t* pod_off;   // [n_problems + 1]
  const b200s_nodes_soa_t* nodes;
  const int32_t* excl_off;  // [n_problems + 1], may be null (nothing removed)
  const int32_t* excl_rows;
};

// Builds every device buffer a fit / solve over `n_problems` pod sets needs.
int build_resident(b200s_ctx* ctx, uint64_t catalog_id, int n_problems, const b200s_pods_soa_t* const* pods_in,
                   const b200s_nodes_soa_t* const* nodes_in, const b200s_templates_t* tm, b200s_resident** out,
                   const SharedBatch* sh = nullptr) {
  auto it = ctx->catalogs.find(catalog_id);
  if (it == ctx->catalogs.end()) return fail(ctx, B200S_ESTATE, "unknown catalog id");
  Catalog* cat = it->second.get();
  if (!tm || tm->n_templates < 0 || tm->n_templates > MAXT) return fail(ctx, B200S_EINVAL, "bad templates");
  if (n_problems < 1) return fail(ctx, B200S_EINVAL, "no problems");
  const int R = cat->R, K = cat->K, NB = cat->NB, W = cat->W, T = tm->n_templates;
  long long Ptot = 0, Etot = 0;
  const int n_sets = sh ? 1 : n_problems;  // pod / class-table sets to upload
  const int E0 = sh && sh->nodes ? sh->nodes->n_nodes : 0;
  if (sh) {
    int rc = check_shapes(ctx, cat, pods_in[0], sh->nodes);
    if (rc) return rc;
    Ptot = pods_in[0]->n_pods;
    Etot = (long long)n_problems * E0;
    if (!sh->pod_off || sh->pod_off[0] != 0 || sh->pod_off[n_problems] != Ptot) return fail(ctx, B200S_EINVAL, "pod_off must run from 0 to n_pods");
    for (int b = 0; b < n_problems; b++)
      if (sh->pod_off[b + 1] < sh->pod_off[b]) return fail(ctx, B200S_EINVAL, "pod_off must not decrease");
    if (sh->excl_off) {
      if (sh->excl_off[0] != 0) return fail(ctx, B200S_EINVAL, "excl_off must start at 0");
      for (int b = 0; b < n_problems; b++)
        if (sh->excl_off[b + 1] < sh->excl_off[b]) return fail(ctx, B200S_EINVAL, "excl_off must not decrease");
      for (int j = 0; j < sh->excl_off[n_problems]; j++)
        if (!sh->excl_rows || sh->excl_rows[j] < 0 || sh->excl_rows[j] >= E0) return fail(ctx, B200S_EINVAL, "excluded node row out of range");
    }
  } else {
    for (int b = 0; b < n_problems; b++) {
      const b200s_nodes_soa_t* nd = nodes_in ? nodes_in[b] : nullptr;
      int rc = check_shapes(ctx, cat, pods_in[b], nd);
      if (rc) return rc;
      Ptot += pods_in[b]->n_pods;
      if (nd) Etot += nd->n_nodes;
    }
  }
  if (Ptot > (1ll << 30) || Etot > (1ll << 30)) return fail(ctx, B200S_EINVAL, "problem too large");
  std::unique_ptr<b200s_resident, void (*)(b200s_resident*)> r(new (std::nothrow) b200s_resident(), free_resident);
  if (!r) return fail(ctx, B200S_ENOMEM, "host allocation");
  r->ctx = ctx;
  r->cat = cat;
  r->P = (int)Ptot;
  r->E = (int)Etot;
  r->T = T;
  r->n_problems = n_problems;
  r->Ppad = ffd_padded(r->P);
  const size_t P = (size_t)std::max<long long>(Ptot, 1), E = (size_t)std::max<long long>(Etot, 1);

  // ---- pods: per-pod ids / tie-break keys concatenated; class tables concatenated with id offsets ----
  long long NStot = 0, NQtot = 0;
  for (int b = 0; b < n_sets; b++) { NStot += pods_in[b]->n_sigs; NQtot += pods_in[b]->n_reqcls; }
  const size_t NS = (size_t)std::max<long long>(NStot, 1), NQ = (size_t)std::max<long long>(NQtot, 1);
  int32_t* d_sig = r->dalloc<int32_t>(P);
  int32_t* d_rq = r->dalloc<int32_t>(P);
  int64_t* d_ts = r->dalloc<int64_t>(P);
  uint64_t* d_uhi = r->dalloc<uint64_t>(P);
  uint64_t* d_ulo = r->dalloc<uint64_t>(P);
  uint64_t* d_smask = r->dalloc<uint64_t>((size_t)std::max(K, 1) * NS);
  uint64_t* d_sflags = r->dalloc<uint64_t>(NS);
  int32_t* d_sgt = r->dalloc<int32_t>((size_t)std::max(NB, 1) * NS);
  int32_t* d_slt = r->dalloc<int32_t>((size_t)std::max(NB, 1) * NS);
  uint64_t* d_stol = r->dalloc<uint64_t>(NS);
  int64_t* d_rq_req = r->dalloc<int64_t>((size_t)R * NQ);
  if (!d_sig || !d_rq || !d_ts || !d_uhi || !d_ulo || !d_smask || !d_sflags || !d_sgt || !d_slt || !d_stol || !d_rq_req)
    return fail(ctx, B200S_ENOMEM, "device allocation (pods)");
  {
    size_t off = 0, soff = 0, qoff = 0;
    std::vector<int32_t> tmp;
    for (int b = 0; b < n_sets; b++) {
      const b200s_pods_soa_t* ps = pods_in[b];
      const size_t n = (size_t)ps->n_pods, ns = (size_t)ps->n_sigs, nq = (size_t)ps->n_reqcls;
      int rc = 0;
      if (n_sets == 1) {
        rc = upload(ctx, d_sig, ps->sig_id, n);
        if (!rc) rc = upload(ctx, d_rq, ps->req_id, n);
      } else {  // class ids of problem b are shifted by the classes of the problems before it
        tmp.resize(2 * n);
        for (size_t i = 0; i < n; i++) { tmp[i] = ps->sig_id[i] + (int32_t)soff; tmp[n + i] = ps->req_id[i] + (int32_t)qoff; }
        rc = upload(ctx, d_sig + off, tmp.data(), n);
        if (!rc) rc = upload(ctx, d_rq + off, tmp.data() + n, n);
        if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "sync (ids)");
      }
      if (!rc) rc = upload(ctx, d_ts + off, ps->ts, n);
      if (!rc) rc = upload(ctx, d_uhi + off, ps->uid_hi, n);
      if (!rc) rc = upload(ctx, d_ulo + off, ps->uid_lo, n);
      for (int x = 0; x < K && !rc; x++) rc = upload(ctx, d_smask + (size_t)x * NStot + soff, ps->sig_mask + (size_t)x * ns, ns);
      for (int x = 0; x < NB && !rc; x++) {
        rc = upload(ctx, d_sgt + (size_t)x * NStot + soff, ps->sig_gt + (size_t)x * ns, ns);
        if (!rc) rc = upload(ctx, d_slt + (size_t)x * NStot + soff, ps->sig_lt + (size_t)x * ns, ns);
      }
      if (!rc) rc = upload(ctx, d_sflags + soff, ps->sig_flags, ns);
      for (size_t i = 0; i < ns; i++)
        if (ps->sig_flags[i] & 0x1111111111111111ull) r->plain = false;  // a DEFINED bit in some key's nibble
      if (!rc) rc = upload(ctx, d_stol + soff, ps->sig_tolerated, ns);
      for (int x = 0; x < R && !rc; x++) rc = upload(ctx, d_rq_req + (size_t)x * NQtot + qoff, ps->rq_requests + (size_t)x * nq, nq);
      if (rc) return rc;
      off += n; soff += ns; qoff += nq;
    }
  }
  r->pods = PodsDev{(int32_t)Ptot, (int32_t)NStot, (int32_t)NQtot, pods_in[0]->res_cpu, pods_in[0]->res_memory, d_sig, d_rq,
                    d_ts, d_uhi, d_ulo, d_smask, d_sflags, d_sgt, d_slt, d_stol, d_rq_req};

  // ---- FFD: rank of every request class in (cpu desc, memory desc) order ----
  r->rq_rank = r->dalloc<int32_t>((size_t)std::max<long long>((long long)NQtot, 1));
  r->ffd_k1 = r->dalloc<uint64_t>((size_t)std::max<long long>((long long)Ptot, 1));
  if (!r->rq_rank || !r->ffd_k1) return fail(ctx, B200S_ENOMEM, "device allocation (FFD keys)");
  if (NQtot <= 16384) {
    CUDA_TRY(ctx, launch_ffd_class_rank(r->pods, r->rq_rank, ctx->stream), "ffd_class_rank launch");
  } else {  // many classes: sorted on the host, once per resident set
    const size_t nqt = (size_t)NQtot;
    std::vector<int64_t> cpu(nqt), mem(nqt);
    size_t qo = 0;
    for (int b = 0; b < n_sets; b++) {
      const b200s_pods_soa_t* ps = pods_in[b];
      const size_t nq = (size_t)ps->n_reqcls;
      for (size_t q = 0; q < nq; q++) {
        cpu[qo + q] = ps->rq_requests[(size_t)ps->res_cpu * nq + q];
        mem[qo + q] = ps->rq_requests[(size_t)ps->res_memory * nq + q];
      }
      qo += nq;
    }
    std::vector<int32_t> ordq(nqt), rank(nqt);
    for (size_t q = 0; q < nqt; q++) ordq[q] = (int32_t)q;
    std::sort(ordq.begin(), ordq.end(), [&](int32_t x, int32_t y) {
      return cpu[x] != cpu[y] ? cpu[x] > cpu[y] : mem[x] > mem[y];
    });
    for (size_t i = 0; i < nqt; i++) {
      const int32_t q = ordq[i];
      const bool tie = i > 0 && cpu[ordq[i - 1]] == cpu[q] && mem[ordq[i - 1]] == mem[q];
      rank[q] = tie ? rank[ordq[i - 1]] : (int32_t)i;
    }
    int rc = upload(ctx, r->rq_rank, rank.data(), nqt);
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "sync (ranks)");
    if (rc) return rc;
  }

  // ---- problems, segment ids ----
  r->problems_host.resize(n_problems);
  {
    int po = 0, eo = 0;
    std::vector<int32_t> seg, rowp;
    if (n_problems > 1) { seg.resize(P); rowp.resize(P); }
    for (int b = 0; b < n_problems; b++) {
      PackProblem& pb = r->problems_host[b];
      pb.pod_begin = po;
      pb.pod_end = sh ? sh->pod_off[b + 1] : po + pods_in[b]->n_pods;
      pb.node_begin = eo;
      pb.node_end = eo + (sh ? E0 : (nodes_in && nodes_in[b]) ? nodes_in[b]->n_nodes : 0);
      pb.n_claims = 0; pb.n_unscheduled = 0; pb.evals = 0;
      pb.status = 0; pb._pad = 0;
      for (int i = 0; i < 12; i++) pb.prof[i] = 0;
      if (n_problems > 1)
        for (int i = pb.pod_begin; i < pb.pod_end; i++) { seg[i] = b; rowp[i] = b; }
      po = pb.pod_end;
      eo = pb.node_end;
    }
    r->problems = r->dalloc<PackProblem>(n_problems);
    if (!r->problems) return fail(ctx, B200S_ENOMEM, "device allocation (problems)");
    if (n_problems > 1) {
      r->seg = r->dalloc<int32_t>(P);
      r->row_problem = r->dalloc<int32_t>(P);
      if (!r->seg || !r->row_problem) return fail(ctx, B200S_ENOMEM, "device allocation (segments)");
      int rc = upload(ctx, r->seg, seg.data(), (size_t)Ptot);
      if (!rc) rc = upload(ctx, r->row_problem, rowp.data(), (size_t)Ptot);
      if (rc) return rc;
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (segments)");  // host vectors go out of scope
    }
  }

  // ---- templates ----
  {
    std::unique_ptr<TmplDev> h(new TmplDev());
    std::memset(h.get(), 0, sizeof(TmplDev));
    h->T = T;
    for (int t = 0; t < T; t++) {
      h->taints[t] = tm->taints[t];
      h->rflags[t] = tm->req_flags[t];
      for (int k = 0; k < MAXK; k++) h->rmask[t][k] = tm->req_mask[t][k];
      for (int b = 0; b < MAXB; b++) { h->gt[t][b] = tm->req_gt[t][b]; h->lt[t][b] = tm->req_lt[t][b]; }
      for (int x = 0; x < MAXR; x++) h->daemon[t][x] = tm->daemon[t][x];
      for (int x = 0; x < cat->R; x++)
        if (tm->has_limit[t][x]) {
          h->limit[t][x] = tm->limit[t][x];
          h->limit_mask[t] |= 1u << x;
          h->any_limit = 1;
        }
    }
    r->any_limit = h->any_limit != 0;
    if (r->any_limit && !cat->kt.cap_dv) return fail(ctx, B200S_EINVAL, "templates carry NodePool limits but the catalog has no capacity rows");
    r->tmpl = r->dalloc<TmplDev>(1);
    r->tmask = r->dalloc<uint64_t>((size_t)std::max(T, 1) * W);
    r->alive = r->dalloc<int32_t>(MAXT);
    if (!r->tmpl || !r->tmask || !r->alive) return fail(ctx, B200S_ENOMEM, "device allocation (templates)");
    CUDA_TRY(ctx, cudaMemcpyAsync(r->tmpl, h.get(), sizeof(TmplDev), cudaMemcpyHostToDevice, ctx->stream), "H2D templates");
    if (tm->type_admit && T > 0) {
      r->type_admit = r->dalloc<uint64_t>((size_t)T * W);
      if (!r->type_admit) return fail(ctx, B200S_ENOMEM, "device allocation (type_admit)");
      int rc = upload(ctx, r->type_admit, tm->type_admit, (size_t)T * W);
      if (rc) return rc;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (templates)");
  }

  // ---- existing nodes (concatenated; resource/key-major over the total) ----
  r->n_avail0 = r->dalloc<int64_t>((size_t)R * E);
  r->n_rmask0 = r->dalloc<uint64_t>((size_t)std::max(K, 1) * E);
  r->n_rflags0 = r->dalloc<uint64_t>(E);
  r->n_taints = r->dalloc<uint64_t>(E);
  if (!r->n_avail0 || !r->n_rmask0 || !r->n_rflags0 || !r->n_taints) return fail(ctx, B200S_ENOMEM, "device allocation (nodes)");
  if (sh && E0 > 0) {  // the shared rows once, then a copy per problem on the device
    const b200s_nodes_soa_t* nd = sh->nodes;
    const size_t n = (size_t)E0;
    int64_t* s_avail = r->dalloc<int64_t>((size_t)R * n);
    uint64_t* s_rmask = r->dalloc<uint64_t>((size_t)std::max(K, 1) * n);
    uint64_t* s_rflags = r->dalloc<uint64_t>(n);
    uint64_t* s_taints = r->dalloc<uint64_t>(n);
    std::vector<int64_t> rows;  // batch-wide rows of the removed nodes
    if (sh->excl_off)
      for (int b = 0; b < n_problems; b++)
        for (int j = sh->excl_off[b]; j < sh->excl_off[b + 1]; j++) rows.push_back((int64_t)b * E0 + sh->excl_rows[j]);
    int64_t* d_rows = r->dalloc<int64_t>(std::max<size_t>(rows.size(), 1));
    if (!s_avail || !s_rmask || !s_rflags || !s_taints || !d_rows) return fail(ctx, B200S_ENOMEM, "device allocation (shared nodes)");
    int rc = upload(ctx, s_avail, nd->available, (size_t)R * n);
    if (!rc) rc = upload(ctx, s_rmask, nd->req_mask, (size_t)K * n);
    if (!rc) rc = upload(ctx, s_rflags, nd->req_flags, n);
    if (!rc) rc = upload(ctx, s_taints, nd->taints, n);
    if (!rc) rc = upload(ctx, d_rows, rows.data(), rows.size());
    if (rc) return rc;
    CUDA_TRY(ctx, launch_replicate_nodes(R, K, E0, Etot, s_avail, s_rmask, s_rflags, s_taints, r->n_avail0, r->n_rmask0, r->n_rflags0,
                                         r->n_taints, (int)rows.size(), d_rows, ctx->n_sm, ctx->stream), "replicate_nodes launch");
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (shared nodes)");  // rows goes out of scope
  } else {
    size_t off = 0;
    for (int b = 0; b < n_problems; b++) {
      const b200s_nodes_soa_t* nd = nodes_in ? nodes_in[b] : nullptr;
      if (!nd || nd->n_nodes == 0) continue;
      const size_t n = (size_t)nd->n_nodes;
      int rc = 0;
      for (int x = 0; x < R && !rc; x++) rc = upload(ctx, r->n_avail0 + (size_t)x * Etot + off, nd->available + (size_t)x * n, n);
      for (int x = 0; x < K && !rc; x++) rc = upload(ctx, r->n_rmask0 + (size_t)x * Etot + off, nd->req_mask + (size_t)x * n, n);
      if (!rc) rc = upload(ctx, r->n_rflags0 + off, nd->req_flags, n);
      if (!rc) rc = upload(ctx, r->n_taints + off, nd->taints, n);
      if (rc) return rc;
      off += n;
    }
  }

  // ---- stage A outputs, sort, pack state, outputs ----
  const size_t T1 = (size_t)std::max(T, 1);
  r->sg_status = r->dalloc<uint8_t>(NS * T1);
  r->sgC = r->dalloc<uint64_t>(NS * T1 * W);
  r->sgO = r->dalloc<uint64_t>(NS * T1 * W);
  r->sgCO = r->dalloc<uint64_t>(NS * T1 * W);
  r->rqF = r->dalloc<uint64_t>(NQ * T1 * W);
  r->tm_count = r->dalloc<int32_t>(MAXT);
  r->a_tmpl = r->dalloc<int32_t>(P);
  r->a_mask = nullptr;  // the P x W rows are materialised only for b200s_(resident_)fit_matrix: the Solve reads a row as
                        // class row & class row when it opens a claim
  r->a_code = r->dalloc<uint8_t>(P * (size_t)std::max(T, 1));
  r->a_evals = r->dalloc<unsigned long long>(1);
  PackArrays& A = r->arr;
  A.c_mask = r->dalloc<uint64_t>(P * W);
  A.c_req = r->dalloc<int64_t>(P * R);
  A.c_hi = r->dalloc<int64_t>(P * R);
  A.c_thr = r->dalloc<int64_t>(P * R);
  A.c_rmask = r->dalloc<uint64_t>(P * (size_t)std::max(K, 1));
  A.c_rflags = r->dalloc<uint64_t>(P);
  A.c_gt = r->dalloc<int64_t>(P * (size_t)std::max(NB, 1));
  A.c_lt = r->dalloc<int64_t>(P * (size_t)std::max(NB, 1));
  A.c_tmpl = r->dalloc<int32_t>(P);
  A.c_npods = r->dalloc<int32_t>(P);
  A.c_nopt = r->dalloc<int32_t>(P);
  A.ord = r->dalloc<int32_t>(P);
  A.ord_key = r->dalloc<int32_t>(P);
  A.ord_snap = r->dalloc<int32_t>(P);
  A.Etot = (int32_t)Etot;
  A.n_avail = r->dalloc<int64_t>((size_t)R * E);
  A.n_rmask = r->dalloc<uint64_t>((size_t)std::max(K, 1) * E);
  A.n_rflags = r->dalloc<uint64_t>(E);
  A.n_taints = r->n_taints;
  A.n_npods = r->dalloc<int32_t>(E);
  A.queue = r->dalloc<int32_t>((size_t)r->Ppad);
  A.lastlen = r->dalloc<int32_t>(P);
  A.a_tmpl = r->a_tmpl;
  A.a_mask = nullptr;
  A.sgCO = nullptr; A.rqF = nullptr; A.T = T;
  A.pod_kind = r->dalloc<int32_t>(P);
  A.pod_index = r->dalloc<int32_t>(P);
  A.pod_seq = r->dalloc<int32_t>(P);
  ClaimOut& O = r->cout;
  O.tmpl = r->dalloc<int32_t>(P);
  O.npods = r->dalloc<int32_t>(P);
  O.req = r->dalloc<int64_t>(P * R);
  O.mask = r->dalloc<uint64_t>(P * W);
  O.rmask = r->dalloc<uint64_t>(P * (size_t)std::max(K, 1));
  O.rflags = r->dalloc<uint64_t>(P);
  O.gt = r->dalloc<int64_t>(P * (size_t)std::max(NB, 1));
  O.lt = r->dalloc<int64_t>(P * (size_t)std::max(NB, 1));
  O.pos_of_claim = r->dalloc<int32_t>(P);
  r->sel_sku = r->dalloc<int32_t>(P);
  r->sel_price = r->dalloc<uint64_t>(P);
  r->sel_nopt = r->dalloc<int32_t>(P);
  void* must[] = {r->sg_status, r->sgC, r->sgO, r->sgCO, r->rqF, r->tm_count, r->a_tmpl, r->a_code, r->a_evals, A.c_mask, A.c_req, A.c_hi, A.c_thr, A.c_rmask, A.c_rflags, A.c_gt, A.c_lt,
                  A.c_tmpl, A.c_npods, A.c_nopt, A.ord, A.ord_key, A.ord_snap, A.n_avail, A.n_rmask, A.n_rflags, A.n_npods, A.queue, A.lastlen, A.pod_kind,
                  A.pod_index, A.pod_seq, O.tmpl, O.npods, O.req, O.mask, O.rmask, O.rflags, O.gt, O.lt, O.pos_of_claim,
                  r->sel_sku, r->sel_price, r->sel_nopt};
  for (void* p : must)
    if (!p) return fail(ctx, B200S_ENOMEM, "device allocation (state)");
  r->reg_plan.eligible = false;
  if (r->plain && Etot == 0 && T > 0 && Ptot > 0 && !r->any_limit && !getenv("B200S_PACK_NO_REG")) {
    // selector-less pods, no existing nodes: plan the register-resident pack (granules, tables) on the host
    std::vector<const int64_t*> rqs(n_sets);
    std::vector<int32_t> nqs(n_sets);
    for (int b = 0; b < n_sets; b++) { rqs[b] = pods_in[b]->rq_requests; nqs[b] = pods_in[b]->n_reqcls; }
    RegPackPlan& plan = r->reg_plan;
    if (reg_pack_plan(cat->kt, cat->h_dv.data(), T, tm->daemon, n_sets, rqs.data(), nqs.data(), ctx->smem_limit, &plan)) {
      std::vector<int32_t> rq32((size_t)REG_RT * NQ, 0);
      size_t qoff = 0;
      for (int b = 0; b < n_sets; b++) {
        reg_pack_scale_requests(plan, rqs[b], nqs[b], rq32.data(), NQ, qoff);
        qoff += (size_t)nqs[b];
      }
      uint8_t* d_tab = r->dalloc<uint8_t>(plan.tab.size());
      int32_t* d_win = r->dalloc<int32_t>(plan.win.size());
      int32_t* d_dm = r->dalloc<int32_t>(plan.daemon32.size());
      int32_t* d_rq = r->dalloc<int32_t>(rq32.size());
      r->feas_bits = r->dalloc<uint32_t>((size_t)T * plan.wpt);
      if (!d_tab || !d_win || !d_dm || !d_rq || !r->feas_bits) return fail(ctx, B200S_ENOMEM, "device allocation (register pack tables)");
      int rc = upload(ctx, d_tab, plan.tab.data(), plan.tab.size());
      if (!rc) rc = upload(ctx, d_win, plan.win.data(), plan.win.size());
      if (!rc) rc = upload(ctx, d_dm, plan.daemon32.data(), plan.daemon32.size());
      if (!rc) rc = upload(ctx, d_rq, rq32.data(), rq32.size());
      if (rc) return rc;
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (register pack tables)");  // rq32 goes out of scope
      r->reg_dev = RegPackDev{d_tab, d_win, r->feas_bits, d_dm, d_rq};
    }
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (upload)");
  *out = r.release();
  return B200S_OK;
}

// stage A: template pre-filter, class tables, dense fit matrix
// stage A.  want_rows: materialise every pod's instance-type row (b200s_fit_matrix's output); dense: run the per-pod
// kernel at all (false when the per-pod results were imported from other GPUs: only the class tables are needed)
int enqueue_fit(b200s_ctx* ctx, b200s_resident* r, int* launches, bool want_rows, bool dense = true) {
  if (want_rows && !r->a_mask) {
    r->a_mask = r->dalloc<uint64_t>((size_t)std::max(r->P, 1) * r->cat->W);
    if (!r->a_mask) return fail(ctx, B200S_ENOMEM, "device allocation (fit rows)");
  }
  r->arr.a_mask = nullptr;  // the pack always reads rows through the class tables
  r->arr.sgCO = r->sgCO;
  r->arr.rqF = r->rqF;
  r->arr.T = r->T;
  r->arr.sg_status = r->sg_status;
  r->arr.sgC = r->sgC;
  r->arr.sgO = r->sgO;
  r->arr.tmask = r->tmask;
  r->arr.a_code = r->a_code;
  if (r->any_limit && !r->arr.limit_rows) {
    r->arr.limit_rows = r->dalloc<uint64_t>((size_t)r->n_problems * std::max(r->T, 1) * r->cat->W);
    if (!r->arr.limit_rows) return fail(ctx, B200S_ENOMEM, "device allocation (limit rows)");
  }
  CUDA_TRY(ctx, cudaMemsetAsync(r->a_evals, 0, sizeof(unsigned long long), ctx->stream), "memset");
  CUDA_TRY(ctx, launch_template_prefilter(r->cat->kt, r->tmpl, r->type_admit, r->tmask, r->alive, r->tm_count, ctx->stream),
           "template_prefilter launch");
  CUDA_TRY(ctx, launch_class_tables(r->cat->kt, r->tmpl, r->T, r->tmask, r->alive, r->pods, r->sg_status, r->sgC, r->sgO,
                                    r->sgCO, r->rqF, r->tm_count, ctx->n_sm, ctx->smem_limit, ctx->stream),
           "class_tables launch");
  if (!r->ev_dense0) {
    CUDA_TRY(ctx, cudaEventCreate(&r->ev_dense0), "event");
    CUDA_TRY(ctx, cudaEventCreate(&r->ev_dense1), "event");
  }
  CUDA_TRY(ctx, cudaEventRecord(r->ev_dense0, ctx->stream), "event record");
  if (dense)
    CUDA_TRY(ctx, launch_fit_dense(r->cat->kt, r->T, r->pods, r->sg_status, r->sgC, r->sgO, r->sgCO, r->rqF, r->tm_count, r->a_tmpl,
                                   want_rows ? r->a_mask : nullptr, r->a_code, r->a_evals, ctx->n_sm, ctx->stream),
             "fit_dense launch");
  CUDA_TRY(ctx, cudaEventRecord(r->ev_dense1, ctx->stream), "event record");
  *launches += 1 + (((long long)r->pods.NS + r->pods.NQ) * r->T > 0 ? 1 : 0) + (r->P > 0 ? 1 : 0);
  return 0;
}

int enqueue_solve(b200s_ctx* ctx, b200s_resident* r, int* launches, bool imported = false) {
  const Catalog* cat = r->cat;
  const size_t E = (size_t)std::max(r->E, 1), P = (size_t)std::max(r->P, 1);
  int rc = enqueue_fit(ctx, r, launches, false, !imported);
  if (rc) return rc;
  // fresh mutable state
  CUDA_TRY(ctx, cudaMemcpyAsync(r->arr.n_avail, r->n_avail0, (size_t)cat->R * E * 8, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
  CUDA_TRY(ctx, cudaMemcpyAsync(r->arr.n_rmask, r->n_rmask0, (size_t)std::max(cat->K, 1) * E * 8, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
  CUDA_TRY(ctx, cudaMemcpyAsync(r->arr.n_rflags, r->n_rflags0, E * 8, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
  CUDA_TRY(ctx, cudaMemsetAsync(r->arr.n_npods, 0, E * 4, ctx->stream), "memset");
  CUDA_TRY(ctx, cudaMemsetAsync(r->arr.lastlen, 0, P * 4, ctx->stream), "memset");
  CUDA_TRY(ctx, cudaMemsetAsync(r->arr.pod_kind, 0xFF, P * 4, ctx->stream), "memset");
  CUDA_TRY(ctx, cudaMemsetAsync(r->arr.pod_index, 0xFF, P * 4, ctx->stream), "memset");
  CUDA_TRY(ctx, cudaMemsetAsync(r->arr.pod_seq, 0xFF, P * 4, ctx->stream), "memset");
  // the host copy doubles as the fetch buffer: a Solve before this one left its outputs (and status 1) in it
  for (PackProblem& pb : r->problems_host) {
    pb.n_claims = 0; pb.n_unscheduled = 0; pb.evals = 0; pb.status = 0;
    for (unsigned long long& c : pb.prof) c = 0;
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(r->problems, r->problems_host.data(), sizeof(PackProblem) * r->n_problems,
                                cudaMemcpyHostToDevice, ctx->stream), "H2D problems");
  int max_pods = 0;  // the largest problem of the batch
  for (const PackProblem& pb : r->problems_host) max_pods = std::max(max_pods, pb.pod_end - pb.pod_begin);
  int nl = 0;
  CUDA_TRY(ctx, launch_ffd_sort(r->pods, r->seg, r->rq_rank, r->ffd_k1, r->arr.queue, &nl, ctx->stream), "ffd_sort launch");
  *launches += nl;
  bool generic_pack = true;
  if (r->reg_plan.eligible && !r->profile_pack) {
    // selector-less pods, no existing nodes: the register-resident form; problems it cannot hold (more claims than
    // its threads keep) are left to the generic kernel, which skips the finished ones
    bool may_overflow = false;
    int n_reg = 0;
    CUDA_TRY(ctx, launch_feas_table(cat->kt, r->tmask, r->T, r->reg_plan, r->feas_bits, ctx->stream), "feas_table launch");
    CUDA_TRY(ctx, launch_pack_reg(cat->kt, r->tmpl, r->T, r->pods, r->arr, r->problems, r->n_problems, max_pods, r->reg_plan,
                                  r->reg_dev, &may_overflow, &n_reg, ctx->stream),
             "pack_reg launch");
    *launches += 1 + n_reg;
    generic_pack = may_overflow;
  }
  if (generic_pack) {
    CUDA_TRY(ctx, launch_pack(cat->kt, r->tmpl, r->T, r->pods, r->arr, r->problems, r->n_problems, !r->plain, r->E > 0, max_pods,
                              r->profile_pack, ctx->smem_limit, ctx->stream),
             "pack launch");
    *launches += 1;
  }
  CUDA_TRY(ctx, launch_pack_finalize(cat->kt, r->pods, r->arr, r->problems, r->n_problems, max_pods, r->cout, ctx->stream),
           "pack_finalize launch");
  CUDA_TRY(ctx, launch_select_cheapest(cat->kt, r->P, r->cout.mask, r->cout.rmask, (size_t)cat->K, 1, r->cout.rflags,
                                       r->cout.gt, r->cout.lt, (size_t)cat->NB, 1, r->sel_sku, r->sel_price, r->sel_nopt,
                                       r->problems, r->row_problem, ctx->stream),
           "select_cheapest launch");
  *launches += 2;
  return 0;
}

int fetch_fit(b200s_ctx* ctx, b200s_resident* r, int32_t* tmpl_out, uint64_t* mask_out, uint8_t* code_out, int64_t* evals_out) {
  const size_t P = (size_t)r->P;
  if (P && mask_out && !r->rows_done) return fail(ctx, B200S_ESTATE, "the fit rows were not materialised: call b200s_resident_fit_matrix first");
  if (P) {
    CUDA_TRY(ctx, cudaMemcpyAsync(tmpl_out, r->a_tmpl, P * 4, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(mask_out, r->a_mask, P * r->cat->W * 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
    if (r->T) CUDA_TRY(ctx, cudaMemcpyAsync(code_out, r->a_code, P * r->T, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
  }
  unsigned long long ev = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&ev, r->a_evals, 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (fetch fit)");
  if (evals_out) *evals_out = (int64_t)ev;
  return 0;
}

// copies problem b's results into out
int fetch_problem(b200s_ctx* ctx, b200s_resident* r, int b, const PackProblem& pb, unsigned long long fit_evals,
                  b200s_result_t* out) {
  const Catalog* cat = r->cat;
  const int R = cat->R, K = cat->K, NB = cat->NB, W = cat->W, T = r->T;
  const size_t base = (size_t)pb.pod_begin, n = (size_t)(pb.pod_end - pb.pod_begin), nc = (size_t)pb.n_claims;
  (void)b;
  out->n_claims = pb.n_claims;
  out->n_unscheduled = pb.n_unscheduled;
  out->fit_evals = (int64_t)(pb.evals + fit_evals);
  if ((int)nc > out->cap_claims) return fail(ctx, B200S_ECAPACITY, "cap_claims too small");
  cudaStream_t s = ctx->stream;
  if (n) {
    CUDA_TRY(ctx, cudaMemcpyAsync(out->pod_place_kind, r->arr.pod_kind + base, n * 4, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->pod_place_index, r->arr.pod_index + base, n * 4, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->pod_seq, r->arr.pod_seq + base, n * 4, cudaMemcpyDeviceToHost, s), "D2H");
    if (T) CUDA_TRY(ctx, cudaMemcpyAsync(out->pod_fit_code, r->a_code + base * T, n * T, cudaMemcpyDeviceToHost, s), "D2H");
  }
  if (nc) {
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_template, r->cout.tmpl + base, nc * 4, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_npods, r->cout.npods + base, nc * 4, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_requests, r->cout.req + base * R, nc * R * 8, cudaMemcpyDeviceToHost, s), "D2H");
    if (out->claim_mask)
      CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_mask, r->cout.mask + base * W, nc * W * 8, cudaMemcpyDeviceToHost, s), "D2H");
    if (K) CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_req_mask, r->cout.rmask + base * K, nc * K * 8, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_req_flags, r->cout.rflags + base, nc * 8, cudaMemcpyDeviceToHost, s), "D2H");
    if (NB) {
      CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_req_gt, r->cout.gt + base * NB, nc * NB * 8, cudaMemcpyDeviceToHost, s), "D2H");
      CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_req_lt, r->cout.lt + base * NB, nc * NB * 8, cudaMemcpyDeviceToHost, s), "D2H");
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_sku, r->sel_sku + base, nc * 4, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_price, r->sel_price + base, nc * 8, cudaMemcpyDeviceToHost, s), "D2H");
    CUDA_TRY(ctx, cudaMemcpyAsync(out->claim_n_options, r->sel_nopt + base, nc * 4, cudaMemcpyDeviceToHost, s), "D2H");
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(s), "sync (fetch solve)");
  // attempt codes are reported for pods left in PodErrors only (scheduler.go:411,419)
  for (size_t i = 0; i < n; i++)
    if (out->pod_place_kind[i] != B200S_PLACE_UNSCHEDULED)
      for (int t = 0; t < T; t++)
        if (out->pod_fit_code[i * T + t] != B200S_FIT_DROPPED) out->pod_fit_code[i * T + t] = B200S_FIT_NONE;
  return 0;
}

int fetch_solve_all(b200s_ctx* ctx, b200s_resident* r, b200s_result_t* const* outs) {
  CUDA_TRY(ctx, cudaMemcpyAsync(r->problems_host.data(), r->problems, sizeof(PackProblem) * r->n_problems,
                                cudaMemcpyDeviceToHost, ctx->stream), "D2H problems");
  unsigned long long fit_ev = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&fit_ev, r->a_evals, 8, cudaMemcpyDeviceToHost, ctx->stream), "D2H");
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (solve)");
  if (r->n_problems > 1) {
    // A batch: every array once for all problems (their rows are ranges of the same device arrays), then host copies
    // of each problem's slices — not fifteen device copies and a synchronisation per problem.
    const Catalog* cat = r->cat;
    const size_t R = (size_t)cat->R, K = (size_t)cat->K, NB = (size_t)cat->NB, W = (size_t)cat->W, T = (size_t)r->T, P = (size_t)r->P;
    bool want_mask = false;
    for (int b = 0; b < r->n_problems; b++) {
      if (!outs[b]) return fail(ctx, B200S_EINVAL, "null result");
      if (r->problems_host[b].n_claims > outs[b]->cap_claims) return fail(ctx, B200S_ECAPACITY, "cap_claims too small");
      want_mask = want_mask || outs[b]->claim_mask != nullptr;
    }
    std::vector<int32_t> h_kind(P), h_index(P), h_seq(P), h_tmpl(P), h_npods(P), h_sku(P), h_nopt(P);
    std::vector<uint8_t> h_code(P * T);
    std::vector<int64_t> h_req(P * R), h_gt(P * NB), h_lt(P * NB);
    std::vector<uint64_t> h_mask(want_mask ? P * W : 0), h_rmask(P * K), h_rflags(P), h_price(P);
    cudaStream_t st = ctx->stream;
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
      return bytes == 0 ? cudaSuccess : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    CUDA_TRY(ctx, d2h(h_kind.data(), r->arr.pod_kind, P * 4), "D2H");
    CUDA_TRY(ctx, d2h(h_index.data(), r->arr.pod_index, P * 4), "D2H");
    CUDA_TRY(ctx, d2h(h_seq.data(), r->arr.pod_seq, P * 4), "D2H");
    CUDA_TRY(ctx, d2h(h_code.data(), r->a_code, P * T), "D2H");
    CUDA_TRY(ctx, d2h(h_tmpl.data(), r->cout.tmpl, P * 4), "D2H");
    CUDA_TRY(ctx, d2h(h_npods.data(), r->cout.npods, P * 4), "D2H");
    CUDA_TRY(ctx, d2h(h_req.data(), r->cout.req, P * R * 8), "D2H");
    if (want_mask) CUDA_TRY(ctx, d2h(h_mask.data(), r->cout.mask, P * W * 8), "D2H");
    CUDA_TRY(ctx, d2h(h_rmask.data(), r->cout.rmask, P * K * 8), "D2H");
    CUDA_TRY(ctx, d2h(h_rflags.data(), r->cout.rflags, P * 8), "D2H");
    CUDA_TRY(ctx, d2h(h_gt.data(), r->cout.gt, P * NB * 8), "D2H");
    CUDA_TRY(ctx, d2h(h_lt.data(), r->cout.lt, P * NB * 8), "D2H");
    CUDA_TRY(ctx, d2h(h_sku.data(), r->sel_sku, P * 4), "D2H");
    CUDA_TRY(ctx, d2h(h_price.data(), r->sel_price, P * 8), "D2H");
    CUDA_TRY(ctx, d2h(h_nopt.data(), r->sel_nopt, P * 4), "D2H");
    CUDA_TRY(ctx, cudaStreamSynchronize(st), "sync (fetch batch)");
    for (int b = 0; b < r->n_problems; b++) {
      const PackProblem& pb = r->problems_host[b];
      b200s_result_t* out = outs[b];
      const size_t base = (size_t)pb.pod_begin, n = (size_t)(pb.pod_end - pb.pod_begin), nc = (size_t)pb.n_claims;
      out->n_claims = pb.n_claims;
      out->n_unscheduled = pb.n_unscheduled;
      out->fit_evals = (int64_t)pb.evals;
      if (n) {
        std::memcpy(out->pod_place_kind, h_kind.data() + base, n * 4);
        std::memcpy(out->pod_place_index, h_index.data() + base, n * 4);
        std::memcpy(out->pod_seq, h_seq.data() + base, n * 4);
        if (T) std::memcpy(out->pod_fit_code, h_code.data() + base * T, n * T);
      }
      if (nc) {
        std::memcpy(out->claim_template, h_tmpl.data() + base, nc * 4);
        std::memcpy(out->claim_npods, h_npods.data() + base, nc * 4);
        std::memcpy(out->claim_requests, h_req.data() + base * R, nc * R * 8);
        if (out->claim_mask) std::memcpy(out->claim_mask, h_mask.data() + base * W, nc * W * 8);
        if (K) std::memcpy(out->claim_req_mask, h_rmask.data() + base * K, nc * K * 8);
        std::memcpy(out->claim_req_flags, h_rflags.data() + base, nc * 8);
        if (NB) {
          std::memcpy(out->claim_req_gt, h_gt.data() + base * NB, nc * NB * 8);
          std::memcpy(out->claim_req_lt, h_lt.data() + base * NB, nc * NB * 8);
        }
        std::memcpy(out->claim_sku, h_sku.data() + base, nc * 4);
        std::memcpy(out->claim_price, h_price.data() + base, nc * 8);
        std::memcpy(out->claim_n_options, h_nopt.data() + base, nc * 4);
      }
      // attempt codes are reported for pods left in PodErrors only (scheduler.go:411,419)
      for (size_t i = 0; i < n; i++)
        if (out->pod_place_kind[i] != B200S_PLACE_UNSCHEDULED)
          for (size_t t = 0; t < T; t++)
            if (out->pod_fit_code[i * T + t] != B200S_FIT_DROPPED) out->pod_fit_code[i * T + t] = B200S_FIT_NONE;
    }
    if (outs[0]) outs[0]->fit_evals += (int64_t)fit_ev;  // stage A ran once for the whole batch
    return 0;
  }
  for (int b = 0; b < r->n_problems; b++) {
    int rc = fetch_problem(ctx, r, b, r->problems_host[b], r->n_problems == 1 ? fit_ev : 0ull, outs[b]);
    if (rc) return rc;
  }
  if (r->n_problems > 1 && outs[0]) outs[0]->fit_evals += (int64_t)fit_ev;  // stage A ran once for the whole batch
  return 0;
}

}  // namespace

extern "C" {

int b200s_version(void) { return B200S_VERSION; }

int b200s_init(int device, b200s_ctx_t** out) {
  if (!out) return B200S_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return B200S_ENODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200S_ENODEVICE;
  if (prop.major != 10) return B200S_ENODEVICE;  // kernels are built for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return B200S_ENODEVICE;
  std::unique_ptr<b200s_ctx> ctx(new (std::nothrow) b200s_ctx());
  if (!ctx) return B200S_ENOMEM;
  ctx->device = device;
  ctx->n_sm = prop.multiProcessorCount;
  ctx->smem_limit = (int)prop.sharedMemPerBlockOptin;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return B200S_ENODEVICE;
  *out = ctx.release();
  return B200S_OK;
}

void b200s_destroy(b200s_ctx_t* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->catalogs)
    for (void* p : kv.second->bufs) cudaFree(p);
  for (auto& b : ctx->cache) cudaFree(b.ptr);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* b200s_last_error(const b200s_ctx_t* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
void* b200s_stream(b200s_ctx_t* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int b200s_synchronize(b200s_ctx_t* ctx) {
  if (!ctx) return B200S_EINVAL;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "synchronize");
  return B200S_OK;
}

void* b200s_event_create(b200s_ctx_t* ctx) {
  if (!ctx) return nullptr;
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return (void*)e;
}
int b200s_event_record(b200s_ctx_t* ctx, void* ev) {
  if (!ctx || !ev) return B200S_EINVAL;
  CUDA_TRY(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->stream), "event record");
  return B200S_OK;
}
int b200s_event_elapsed_ms(b200s_ctx_t* ctx, void* a, void* b, float* ms) {
  if (!ctx || !a || !b || !ms) return B200S_EINVAL;
  CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)b), "event sync");
  CUDA_TRY(ctx, cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b), "event elapsed");
  return B200S_OK;
}
void b200s_event_destroy(void* ev) {
  if (ev) cudaEventDestroy((cudaEvent_t)ev);
}
/* writes `bytes` of a pattern over a scratch buffer on the context's stream: the L2 flush the timing
 * rules ask for between timed iterations */
int b200s_flush_l2(b200s_ctx_t* ctx, size_t bytes) {
  if (!ctx) return B200S_EINVAL;
  void* p = ctx->alloc(bytes);
  if (!p) return fail(ctx, B200S_ENOMEM, "flush buffer");
  cudaError_t e = cudaMemsetAsync(p, 0x5A, bytes, ctx->stream);
  ctx->release(p);
  if (e != cudaSuccess) return fail(ctx, B200S_ECUDA, "flush memset", e);
  return B200S_OK;
}

int b200s_catalog_upload(b200s_ctx_t* ctx, const b200s_catalog_soa_t* c, uint64_t* catalog_id) {
  if (!ctx || !c || !catalog_id) return B200S_EINVAL;
  if (c->n_skus <= 0 || c->n_words != (c->n_skus + 63) / 64 || c->n_res <= 0 || c->n_res > MAXR || c->n_keys < 0 ||
      c->n_keys > MAXK || c->n_bounded < 0 || c->n_bounded > MAXB || c->n_rows <= 0)
    return fail(ctx, B200S_EINVAL, "bad catalog shape");
  cudaSetDevice(ctx->device);
  std::unique_ptr<Catalog> cat(new (std::nothrow) Catalog());
  if (!cat) return B200S_ENOMEM;
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(bytes, 8)) != cudaSuccess) return nullptr;
    cat->bufs.push_back(d);
    if (bytes && cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return nullptr;
    return d;
  };
  const int S = c->n_skus, W = c->n_words, K = c->n_keys, ZC = std::max(c->n_zones * c->n_cts, 1);
  KeyTab& kt = cat->kt;
  kt.n_keys = K; kt.n_res = c->n_res; kt.n_words = W; kt.n_skus = S;
  kt.key_zone = c->key_zone; kt.key_ct = c->key_ct; kt.n_zones = c->n_zones; kt.n_cts = c->n_cts;
  kt.wellknown = c->wellknown;
  kt.n_bounded = c->n_bounded;
  for (int k = 0; k < MAXK; k++) { kt.nvals[k] = 0; kt.bslot[k] = -1; kt.isint[k] = 0; kt.key_row0[k] = 0; }
  for (int k = 0; k < K; k++) { kt.nvals[k] = c->key_nvals[k]; kt.isint[k] = c->val_isint[k]; kt.key_row0[k] = c->key_row0[k]; }
  for (int b = 0; b < c->n_bounded; b++) {
    if (c->bounded_key[b] < 0 || c->bounded_key[b] >= K) return fail(ctx, B200S_EINVAL, "bad bounded key");
    kt.bslot[c->bounded_key[b]] = (int8_t)b;
  }
  for (int r = 0; r < MAXR; r++) { kt.res_nd[r] = 0; kt.res_dv0[r] = 0; kt.res_row0[r] = 0; }
  kt.max_p2 = 0;
  for (int r = 0; r < MAXR; r++) kt.res_p2[r] = 0;
  for (int r = 0; r < c->n_res; r++) {
    kt.res_nd[r] = c->res_nd[r]; kt.res_dv0[r] = c->res_dv0[r]; kt.res_row0[r] = c->res_row0[r];
    int p2 = 0;
    for (int x = 1; x <= c->res_nd[r]; x <<= 1) p2 = x;
    kt.res_p2[r] = p2;
    kt.max_p2 = std::max(kt.max_p2, p2);
  }
  kt.zc_row0 = c->zc_row0; kt.row_nonneg = c->row_nonneg; kt.n_rows = c->n_rows; kt.n_dv = c->n_dv;
  kt.val_int = (const int64_t*)put(c->val_int, (size_t)std::max(K, 1) * 64 * 8);
  kt.rows = (const uint64_t*)put(c->rows, (size_t)c->n_rows * W * 8);
  kt.res_dv = (const int64_t*)put(c->res_dv, (size_t)c->n_dv * 8);
  kt.price = (const uint64_t*)put(c->price, (size_t)ZC * S * 8);
  kt.price_max = c->price_max ? (const uint64_t*)put(c->price_max, (size_t)ZC * S * 8) : nullptr;
  for (int r = 0; r < MAXR; r++) { kt.cap_nd[r] = 0; kt.cap_dv0[r] = 0; kt.cap_row0[r] = 0; }
  kt.cap_dv = nullptr;
  if (c->cap_dv && c->n_cap_dv > 0) {
    for (int r = 0; r < c->n_res; r++) { kt.cap_nd[r] = c->cap_nd[r]; kt.cap_dv0[r] = c->cap_dv0[r]; kt.cap_row0[r] = c->cap_row0[r]; }
    kt.cap_dv = (const int64_t*)put(c->cap_dv, (size_t)c->n_cap_dv * 8);
  }
  {  // rank bucket table: fit_rank of every bucket's lowest value, per resource
    std::vector<uint16_t> bkt((size_t)MAXR * (RANK_BKTS + 1), 0);
    for (int r = 0; r < c->n_res; r++) {
      if (c->res_nd[r] > 65535) return fail(ctx, B200S_EINVAL, "too many distinct allocatable values");
      const int64_t* dv = c->res_dv + c->res_dv0[r];
      uint16_t* row = bkt.data() + (size_t)r * (RANK_BKTS + 1);
      rank_bucket_table(dv, c->res_nd[r], row);
    }
    kt.res_bkt = (const uint16_t*)put(bkt.data(), bkt.size() * 2);
    if (!kt.res_bkt) {
      for (void* p : cat->bufs) cudaFree(p);
      return fail(ctx, B200S_ENOMEM, "catalog upload");
    }
  }
  if (!kt.val_int || !kt.rows || !kt.res_dv || !kt.price) {
    for (void* p : cat->bufs) cudaFree(p);
    return fail(ctx, B200S_ENOMEM, "catalog upload");
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync (catalog)");
  cat->S = S; cat->W = W; cat->R = c->n_res; cat->K = K; cat->NB = c->n_bounded;
  cat->h_dv.assign(c->res_dv, c->res_dv + c->n_dv);
  cat->layout_hash = c->layout_hash;
  *catalog_id = ctx->next_id++;
  ctx->catalogs[*catalog_id] = std::move(cat);
  return B200S_OK;
}

int b200s_catalog_release(b200s_ctx_t* ctx, uint64_t catalog_id) {
  if (!ctx) return B200S_EINVAL;
  auto it = ctx->catalogs.find(catalog_id);
  if (it == ctx->catalogs.end()) return fail(ctx, B200S_ESTATE, "unknown catalog id");
  cudaStreamSynchronize(ctx->stream);
  for (void* p : it->second->bufs) cudaFree(p);
  ctx->catalogs.erase(it);
  return B200S_OK;
}

int b200s_resident_create(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_pods_soa_t* pods,
                          const b200s_nodes_soa_t* existing, const b200s_templates_t* templates, b200s_resident_t** out) {
  if (!ctx || !out || !pods) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  const b200s_pods_soa_t* pp[1] = {pods};
  const b200s_nodes_soa_t* nn[1] = {existing};
  return build_resident(ctx, catalog_id, 1, pp, nn, templates, out);
}

void b200s_resident_free(b200s_ctx_t* ctx, b200s_resident_t* r) {
  if (!ctx || !r) return;
  cudaStreamSynchronize(ctx->stream);
  free_resident(r);
}

int b200s_resident_fit_matrix(b200s_ctx_t* ctx, b200s_resident_t* r) {
  if (!ctx || !r) return B200S_EINVAL;
  r->last_launches = 0;
  int rc = enqueue_fit(ctx, r, &r->last_launches, true);
  if (!rc) { r->fit_done = true; r->rows_done = true; }
  return rc;
}

int b200s_resident_solve(b200s_ctx_t* ctx, b200s_resident_t* r) {
  if (!ctx || !r) return B200S_EINVAL;
  r->last_launches = 0;
  int rc = enqueue_solve(ctx, r, &r->last_launches);
  if (!rc) { r->fit_done = true; r->solve_done = true; }
  return rc;
}

int b200s_resident_fetch_fit(b200s_ctx_t* ctx, b200s_resident_t* r, int32_t* tmpl_out, uint64_t* mask_out, uint8_t* code_out,
                             int64_t* evals_out) {
  if (!ctx || !r || !r->fit_done) return B200S_ESTATE;
  return fetch_fit(ctx, r, tmpl_out, mask_out, code_out, evals_out);
}

int b200s_resident_fetch_solve(b200s_ctx_t* ctx, b200s_resident_t* r, b200s_result_t* out) {
  if (!ctx || !r || !out || !r->solve_done) return B200S_ESTATE;
  b200s_result_t* outs[1] = {out};
  return fetch_solve_all(ctx, r, outs);
}

int b200s_resident_copy_fit_device(b200s_ctx_t* ctx, b200s_resident_t* r, void* d_tmpl, void* d_mask, void* d_code) {
  if (!ctx || !r || !r->fit_done) return B200S_ESTATE;
  const size_t P = (size_t)r->P;
  if (P == 0) return B200S_OK;
  if (d_tmpl) CUDA_TRY(ctx, cudaMemcpyAsync(d_tmpl, r->a_tmpl, P * 4, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
  if (d_mask) {
    if (!r->rows_done) return fail(ctx, B200S_ESTATE, "the fit rows were not materialised: call b200s_resident_fit_matrix first");
    CUDA_TRY(ctx, cudaMemcpyAsync(d_mask, r->a_mask, P * r->cat->W * 8, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
  }
  if (d_code && r->T) CUDA_TRY(ctx, cudaMemcpyAsync(d_code, r->a_code, P * r->T, cudaMemcpyDeviceToDevice, ctx->stream), "D2D");
  return B200S_OK;
}

int b200s_compact_row_bytes(int32_t n_templates) { return n_templates < 0 || n_templates > MAXT ? B200S_EINVAL : compact_row_bytes(n_templates); }

/* the per-pod results another GPU needs to pack these pods: [int8 first feasible template | T attempt codes | pad] per
 * pod, b200s_compact_row_bytes(T) bytes each, written to DEVICE memory d_out (the send buffer of the one all-gather of
 * SURVEY.md section 8e).  The instance-type rows stay where they were computed. */
int b200s_resident_export_compact(b200s_ctx_t* ctx, b200s_resident_t* r, void* d_out) {
  if (!ctx || !r || !d_out) return B200S_EINVAL;
  if (!r->fit_done) return B200S_ESTATE;
  CUDA_TRY(ctx, launch_compact_export(r->P, r->T, r->a_tmpl, r->a_code, (uint8_t*)d_out, ctx->n_sm, ctx->stream), "compact_export launch");
  r->last_launches += r->P > 0 ? 1 : 0;
  return B200S_OK;
}

/* device memory another process of the node can map (CUDA IPC): what the peer-store export writes into */
int b200s_device_alloc(b200s_ctx_t* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out || bytes == 0) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(ctx, B200S_ENOMEM, "cudaMalloc");
  CUDA_TRY(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream), "memset");
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync");
  *out = p;
  return B200S_OK;
}

int b200s_device_free(b200s_ctx_t* ctx, void* p) {
  if (!ctx) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  if (p && cudaFree(p) != cudaSuccess) return fail(ctx, B200S_ECUDA, "cudaFree");
  return B200S_OK;
}

int b200s_ipc_export(b200s_ctx_t* ctx, void* dptr, uint8_t handle[64]) {
  if (!ctx || !dptr || !handle) return B200S_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, dptr), "cudaIpcGetMemHandle");
  std::memcpy(handle, &h, 64);
  return B200S_OK;
}

int b200s_ipc_open(b200s_ctx_t* ctx, const uint8_t handle[64], void** out) {
  if (!ctx || !handle || !out) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  CUDA_TRY(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return B200S_OK;
}

int b200s_ipc_close(b200s_ctx_t* ctx, void* p) {
  if (!ctx || !p) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaIpcCloseMemHandle(p), "cudaIpcCloseMemHandle");
  return B200S_OK;
}

int b200s_copy_device(b200s_ctx_t* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx || !dst || !src) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream), "D2D copy");
  return B200S_OK;
}

/* export fused with the exchange: the rows go straight into every rank's gathered buffer through peer mappings, then a
 * flag barrier (see include/b200sched.h) */
int b200s_resident_export_compact_peers(b200s_ctx_t* ctx, b200s_resident_t* r, int32_t n_peers, int32_t me, void* const* peer_out,
                                        void* const* peer_flags, int64_t row_offset, uint32_t epoch, int32_t* d_timed_out) {
  if (!ctx || !r || !peer_out || !peer_flags || !d_timed_out || n_peers < 1 || n_peers > 8 || me < 0 || me >= n_peers || row_offset < 0)
    return B200S_EINVAL;
  for (int q = 0; q < n_peers; q++)
    if (!peer_out[q] || !peer_flags[q]) return B200S_EINVAL;
  if (!r->fit_done) return B200S_ESTATE;
  CUDA_TRY(ctx, launch_compact_export_peers(r->P, r->T, r->a_tmpl, r->a_code, n_peers, me, peer_out, peer_flags, row_offset, epoch,
                                            d_timed_out, ctx->n_sm, ctx->stream), "compact_export_peers launch");
  r->last_launches += r->P > 0 ? 2 : 1;
  return B200S_OK;
}

/* the inverse on the GPU that packs: d_in holds compact rows for ALL of this resident's pods (gathered from the ranks
 * that evaluated them); afterwards b200s_resident_pack runs the Solve without evaluating the pods again */
int b200s_resident_import_compact(b200s_ctx_t* ctx, b200s_resident_t* r, const void* d_in) {
  if (!ctx || !r || !d_in) return B200S_EINVAL;
  CUDA_TRY(ctx, launch_compact_import(r->P, r->T, (const uint8_t*)d_in, r->a_tmpl, r->a_code, ctx->n_sm, ctx->stream), "compact_import launch");
  r->fit_done = true;
  r->imported = true;
  r->last_launches = r->P > 0 ? 1 : 0;
  return B200S_OK;
}

/* FFD order + pack + cheapest SKU on per-pod stage-A results that are already in place (imported): "merge the global
 * packing" of the row-sharded evaluation.  The class tables are (re)built locally; rows are read through them. */
int b200s_resident_pack(b200s_ctx_t* ctx, b200s_resident_t* r) {
  if (!ctx || !r) return B200S_EINVAL;
  if (!r->imported) return fail(ctx, B200S_ESTATE, "b200s_resident_pack needs b200s_resident_import_compact first");
  int rc = enqueue_solve(ctx, r, &r->last_launches, true);
  if (!rc) r->solve_done = true;
  return rc;
}

int b200s_resident_last_launches(const b200s_resident_t* r) { return r ? r->last_launches : 0; }

int b200s_resident_set_pack_profile(b200s_resident_t* r, int on) {
  if (!r) return B200S_EINVAL;
  r->profile_pack = on != 0;
  return B200S_OK;
}

int b200s_resident_dense_kernel_ms(b200s_ctx_t* ctx, b200s_resident_t* r, float* ms) {
  if (!ctx || !r || !ms || !r->fit_done || !r->ev_dense0) return B200S_ESTATE;
  CUDA_TRY(ctx, cudaEventSynchronize(r->ev_dense1), "event sync");
  CUDA_TRY(ctx, cudaEventElapsedTime(ms, r->ev_dense0, r->ev_dense1), "event elapsed");
  return B200S_OK;
}

/* diagnostic: SM cycles thread 0 of problem 0's pack CTA spent per phase in the last solve
 * (step start, existing nodes, in-flight scan, commit + next pod, order normalisation) and the step count */
int b200s_resident_pack_profile(b200s_ctx_t* ctx, b200s_resident_t* r, uint64_t out[12]) {
  if (!ctx || !r || !out || !r->solve_done) return B200S_ESTATE;
  CUDA_TRY(ctx, cudaMemcpyAsync(r->problems_host.data(), r->problems, sizeof(PackProblem) * r->n_problems,
                                cudaMemcpyDeviceToHost, ctx->stream), "D2H problems");
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream), "sync");
  for (int i = 0; i < 12; i++) out[i] = r->problems_host[0].prof[i];
  return B200S_OK;
}

int b200s_fit_matrix(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_pods_soa_t* pods, const b200s_templates_t* templates,
                     int32_t* tmpl_out, uint64_t* mask_out, uint8_t* code_out, int64_t* evals_out) {
  if (!ctx || !pods || !templates || !tmpl_out || !mask_out || !code_out) return B200S_EINVAL;
  b200s_resident_t* r = nullptr;
  int rc = b200s_resident_create(ctx, catalog_id, pods, nullptr, templates, &r);
  if (rc) return rc;
  rc = b200s_resident_fit_matrix(ctx, r);
  if (!rc) rc = fetch_fit(ctx, r, tmpl_out, mask_out, code_out, evals_out);
  b200s_resident_free(ctx, r);
  return rc;
}

int b200s_select_cheapest(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n, const uint64_t* masks, const uint64_t* req_mask,
                          const uint64_t* req_flags, const int64_t* req_gt, const int64_t* req_lt, int32_t* sku_out,
                          uint64_t* price_out) {
  if (!ctx || n < 0 || !masks || !req_flags || !sku_out || !price_out) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  auto it = ctx->catalogs.find(catalog_id);
  if (it == ctx->catalogs.end()) return fail(ctx, B200S_ESTATE, "unknown catalog id");
  Catalog* cat = it->second.get();
  if (n == 0) return B200S_OK;
  if (!masks_in_range(masks, (size_t)n, cat->S, cat->W)) return fail(ctx, B200S_EINVAL, "select_cheapest: a mask has bits beyond n_skus");
  const size_t N = (size_t)n, K = (size_t)std::max(cat->K, 1), NB = (size_t)std::max(cat->NB, 1), W = (size_t)cat->W;
  uint64_t* d_masks = (uint64_t*)ctx->alloc(N * W * 8);
  uint64_t* d_rmask = (uint64_t*)ctx->alloc(N * K * 8);
  uint64_t* d_rflags = (uint64_t*)ctx->alloc(N * 8);
  int64_t* d_gt = (int64_t*)ctx->alloc(N * NB * 8);
  int64_t* d_lt = (int64_t*)ctx->alloc(N * NB * 8);
  int32_t* d_sku = (int32_t*)ctx->alloc(N * 4);
  uint64_t* d_price = (uint64_t*)ctx->alloc(N * 8);
  void* all[] = {d_masks, d_rmask, d_rflags, d_gt, d_lt, d_sku, d_price};
  auto cleanup = [&]() {
    for (void* p : all)
      if (p) ctx->release(p);
  };
  for (void* p : all)
    if (!p) { cleanup(); return fail(ctx, B200S_ENOMEM, "device allocation"); }
  int rc = upload(ctx, d_masks, masks, N * W);
  if (!rc && cat->K) rc = upload(ctx, d_rmask, req_mask, N * cat->K);   // key-major [K*n]
  if (!rc) rc = upload(ctx, d_rflags, req_flags, N);
  if (!rc && cat->NB) rc = upload(ctx, d_gt, req_gt, N * cat->NB);
  if (!rc && cat->NB) rc = upload(ctx, d_lt, req_lt, N * cat->NB);
  if (!rc) {
    cudaError_t e = launch_select_cheapest(cat->kt, n, d_masks, d_rmask, 1, N, d_rflags, d_gt, d_lt, 1, N, d_sku, d_price,
                                           nullptr, nullptr, nullptr, ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "select_cheapest launch", e);
  }
  if (!rc && cudaMemcpyAsync(sku_out, d_sku, N * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "D2H");
  if (!rc && cudaMemcpyAsync(price_out, d_price, N * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "D2H");
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = fail(ctx, B200S_ECUDA, "sync");
  cleanup();
  return rc;
}

int b200s_order_by_price(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n, const uint64_t* masks, const uint64_t* req_mask,
                         const uint64_t* req_flags, const int64_t* req_gt, const int64_t* req_lt, int32_t max_types,
                         int32_t* order_out, int32_t* count_out, uint8_t* tie_out) {
  if (!ctx || n < 0 || max_types < 1 || !order_out || !count_out || !tie_out) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  auto it = ctx->catalogs.find(catalog_id);
  if (it == ctx->catalogs.end()) return fail(ctx, B200S_ESTATE, "unknown catalog id");
  Catalog* cat = it->second.get();
  if (n == 0) return B200S_OK;
  if (!masks || !masks_in_range(masks, (size_t)n, cat->S, cat->W)) return fail(ctx, B200S_EINVAL, "order_by_price: a mask has bits beyond n_skus");
  const size_t N = (size_t)n, K = (size_t)cat->K, NB = (size_t)cat->NB, W = (size_t)cat->W, M = (size_t)max_types;
  std::vector<void*> held;
  auto dev = [&](size_t bytes) -> void* {
    void* p = ctx->alloc(bytes);
    if (p) held.push_back(p);
    return p;
  };
  auto cleanup = [&]() {
    for (void* p : held) ctx->release(p);
  };
  uint64_t* d_masks = (uint64_t*)dev(N * W * 8);
  uint64_t* d_rmask = (uint64_t*)dev(N * std::max<size_t>(K, 1) * 8);
  uint64_t* d_rflags = (uint64_t*)dev(N * 8);
  int64_t* d_gt = (int64_t*)dev(N * std::max<size_t>(NB, 1) * 8);
  int64_t* d_lt = (int64_t*)dev(N * std::max<size_t>(NB, 1) * 8);
  int32_t* d_order = (int32_t*)dev(N * M * 4);
  int32_t* d_count = (int32_t*)dev(N * 4);
  uint8_t* d_tie = (uint8_t*)dev(N);
  if (held.size() != 8) { cleanup(); return fail(ctx, B200S_ENOMEM, "device allocation"); }
  int rc = upload(ctx, d_masks, masks, N * W);
  if (!rc && K) rc = upload(ctx, d_rmask, req_mask, N * K);
  if (!rc) rc = upload(ctx, d_rflags, req_flags, N);
  if (!rc && NB) rc = upload(ctx, d_gt, req_gt, N * NB);
  if (!rc && NB) rc = upload(ctx, d_lt, req_lt, N * NB);
  if (!rc) {
    cudaError_t e = launch_order_by_price(cat->kt, n, max_types, d_masks, d_rmask, d_rflags, d_gt, d_lt, d_order, d_count,
                                          d_tie, ctx->smem_limit, ctx->stream);
    if (e == cudaErrorInvalidConfiguration) rc = fail(ctx, B200S_ENOTREPRESENTABLE, "order_by_price: catalog too large for one CTA's shared memory");
    else if (e != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "order_by_price launch", e);
  }
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (!rc && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "D2H");
  };
  down(order_out, d_order, N * M * 4);
  down(count_out, d_count, N * 4);
  down(tie_out, d_tie, N);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = fail(ctx, B200S_ECUDA, "sync");
  cleanup();
  return rc;
}

int b200s_consolidation_commands(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_consolidation_sims_t* sims,
                                 b200s_consolidation_cmds_t* out) {
  if (!ctx || !sims || !out || sims->n < 0) return B200S_EINVAL;
  if (!out->decision || !out->reason || !out->n_options || !out->mask || !out->spot_required || !out->head || !out->price_tie)
    return fail(ctx, B200S_EINVAL, "consolidation: null output array");
  cudaSetDevice(ctx->device);
  auto it = ctx->catalogs.find(catalog_id);
  if (it == ctx->catalogs.end()) return fail(ctx, B200S_ESTATE, "unknown catalog id");
  Catalog* cat = it->second.get();
  const KeyTab& kt = cat->kt;
  if (!kt.price_max) return fail(ctx, B200S_EINVAL, "consolidation: the catalog was uploaded without price_max");
  // Has(spot) / Has(on-demand) are evaluated on the catalog's value universe: both values must be in it
  if (kt.key_ct < 0 || sims->ct_spot < 0 || sims->ct_spot >= kt.n_cts || sims->ct_on_demand < 0 ||
      sims->ct_on_demand >= kt.n_cts || sims->ct_reserved >= kt.n_cts || sims->ct_spot == sims->ct_on_demand)
    return fail(ctx, B200S_ENOTREPRESENTABLE, "consolidation: spot / on-demand are not values of the catalog's capacity-type key");
  if (sims->ct_reserved < -1 || sims->ct_reserved == sims->ct_spot || sims->ct_reserved == sims->ct_on_demand)
    return fail(ctx, B200S_EINVAL, "consolidation: ct_reserved must be -1 or a capacity type of its own");
  if (sims->n == 0) return B200S_OK;
  if (!sims->claim_mask || !masks_in_range(sims->claim_mask, (size_t)sims->n, cat->S, cat->W))
    return fail(ctx, B200S_EINVAL, "consolidation: a claim mask has bits beyond n_skus");
  const size_t N = (size_t)sims->n, K = (size_t)cat->K, NB = (size_t)cat->NB, W = (size_t)cat->W;
  std::vector<uint64_t> limit(N);
  for (size_t i = 0; i < N; i++) {
    const double p = sims->candidate_price[i];
    if (!(p >= 0.0) || !std::isfinite(p)) return fail(ctx, B200S_EINVAL, "consolidation: candidate price must be finite and non-negative");
    std::memcpy(&limit[i], &p, 8);
  }
  std::vector<void*> held;
  auto dev = [&](size_t bytes) -> void* {
    void* p = ctx->alloc(bytes);
    if (p) held.push_back(p);
    return p;
  };
  auto cleanup = [&]() {
    for (void* p : held) ctx->release(p);
  };
  ConsolidationArgs a;
  a.n = sims->n; a.spot_to_spot = sims->spot_to_spot != 0;
  a.ct_spot = sims->ct_spot; a.ct_on_demand = sims->ct_on_demand; a.ct_reserved = sims->ct_reserved;
  int32_t* d_unsched = (int32_t*)dev(N * 4);
  int32_t* d_claims = (int32_t*)dev(N * 4);
  int32_t* d_cand = (int32_t*)dev(N * 4);
  uint8_t* d_spot = (uint8_t*)dev(N);
  uint64_t* d_limit = (uint64_t*)dev(N * 8);
  uint64_t* d_cmask = (uint64_t*)dev(N * W * 8);
  uint64_t* d_rmask = (uint64_t*)dev(N * std::max<size_t>(K, 1) * 8);
  uint64_t* d_rflags = (uint64_t*)dev(N * 8);
  int64_t* d_gt = (int64_t*)dev(N * std::max<size_t>(NB, 1) * 8);
  int64_t* d_lt = (int64_t*)dev(N * std::max<size_t>(NB, 1) * 8);
  a.decision = (int32_t*)dev(N * 4);
  a.reason = (int32_t*)dev(N * 4);
  a.n_options = (int32_t*)dev(N * 4);
  a.head = (int32_t*)dev(N * B200S_S2S_MIN_TYPES * 4);
  a.mask = (uint64_t*)dev(N * W * 8);
  a.spot_required = (uint8_t*)dev(N);
  a.price_tie = (uint8_t*)dev(N);
  if (held.size() != 17) { cleanup(); return fail(ctx, B200S_ENOMEM, "device allocation"); }
  a.n_unscheduled = d_unsched; a.n_claims = d_claims; a.n_candidates = d_cand; a.all_spot = d_spot;
  a.candidate_price = d_limit; a.claim_mask = d_cmask; a.rmask = d_rmask; a.rflags = d_rflags; a.rgt = d_gt; a.rlt = d_lt;
  int rc = upload(ctx, d_unsched, sims->n_unscheduled, N);
  if (!rc) rc = upload(ctx, d_claims, sims->n_claims, N);
  if (!rc) rc = upload(ctx, d_cand, sims->n_candidates, N);
  if (!rc) rc = upload(ctx, d_spot, sims->all_spot, N);
  if (!rc) rc = upload(ctx, d_limit, (const uint64_t*)limit.data(), N);
  if (!rc) rc = upload(ctx, d_cmask, sims->claim_mask, N * W);
  if (!rc) rc = upload(ctx, d_rmask, sims->claim_req_mask, N * K);
  if (!rc) rc = upload(ctx, d_rflags, sims->claim_req_flags, N);
  if (!rc && NB) rc = upload(ctx, d_gt, sims->claim_req_gt, N * NB);
  if (!rc && NB) rc = upload(ctx, d_lt, sims->claim_req_lt, N * NB);
  if (!rc) {
    cudaError_t e = launch_consolidation(kt, a, ctx->smem_limit, ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "consolidation launch", e);
  }
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (!rc && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = fail(ctx, B200S_ECUDA, "D2H");
  };
  down(out->decision, a.decision, N * 4);
  down(out->reason, a.reason, N * 4);
  down(out->n_options, a.n_options, N * 4);
  down(out->head, a.head, N * B200S_S2S_MIN_TYPES * 4);
  down(out->mask, a.mask, N * W * 8);
  down(out->spot_required, a.spot_required, N);
  down(out->price_tie, a.price_tie, N);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && !rc) rc = fail(ctx, B200S_ECUDA, "sync");
  cleanup();
  return rc;
}

int b200s_solve_batch(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n_problems, const b200s_pods_soa_t* const* pods,
                      const b200s_nodes_soa_t* const* existing, const b200s_templates_t* templates,
                      const b200s_solve_opts_t* opts, b200s_result_t* const* out) {
  (void)opts;
  if (!ctx || n_problems < 1 || !pods || !templates || !out) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  b200s_resident_t* r = nullptr;
  int rc = build_resident(ctx, catalog_id, n_problems, pods, existing, templates, &r);
  if (rc) return rc;
  r->last_launches = 0;
  rc = enqueue_solve(ctx, r, &r->last_launches);
  if (!rc) rc = fetch_solve_all(ctx, r, out);
  cudaStreamSynchronize(ctx->stream);
  free_resident(r);
  return rc;
}

int b200s_solve_batch_shared(b200s_ctx_t* ctx, uint64_t catalog_id, int32_t n_problems, const b200s_pods_soa_t* pods,
                             const int32_t* pod_off, const b200s_nodes_soa_t* nodes, const int32_t* excl_off,
                             const int32_t* excl_rows, const b200s_templates_t* templates, const b200s_solve_opts_t* opts,
                             b200s_result_t* const* out) {
  (void)opts;
  if (!ctx || n_problems < 1 || !pods || !pod_off || !templates || !out) return B200S_EINVAL;
  cudaSetDevice(ctx->device);
  const b200s_pods_soa_t* pp[1] = {pods};
  const SharedBatch sh{pod_off, nodes, excl_off, excl_rows};
  b200s_resident_t* r = nullptr;
  int rc = build_resident(ctx, catalog_id, n_problems, pp, nullptr, templates, &r, &sh);
  if (rc) return rc;
  r->last_launches = 0;
  rc = enqueue_solve(ctx, r, &r->last_launches);
  if (!rc) rc = fetch_solve_all(ctx, r, out);
  cudaStreamSynchronize(ctx->stream);
  free_resident(r);
  return rc;
}

int b200s_solve(b200s_ctx_t* ctx, uint64_t catalog_id, const b200s_pods_soa_t* pods, const b200s_nodes_soa_t* existing,
                const b200s_templates_t* templates, const b200s_solve_opts_t* opts, b200s_result_t* out) {
  if (!pods || !out) return B200S_EINVAL;
  const b200s_pods_soa_t* pp[1] = {pods};
  const b200s_nodes_soa_t* nn[1] = {existing};
  b200s_result_t* oo[1] = {out};
  return b200s_solve_batch(ctx, catalog_id, 1, pp, nn, templates, opts, oo);
}

}  // extern "C"

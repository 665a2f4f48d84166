// karp_oracle.cpp — CPU oracle: a restatement of the Karpenter v1.7.0 provisioning scheduler as
// vendored by Azure/gpu-provisioner (module sigs.k8s.io/karpenter v1.7.0, reference go.mod:29;
// source under /root/reference/vendor/sigs.k8s.io/karpenter/pkg = "K/").
//
// THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, load or call it.  The product library (libb200sched.so) never
// links or loads anything from oracle/.
//
// PARITY UNPINNED at the scheduler boundary: `go mod vendor` stripped every Karpenter _test.go and
// no Go toolchain exists in the build image, so there is no golden vector and no way to run the
// reference here.  Fidelity is argued function by function (each function below cites the Go it
// follows) and by the property tests in tests/.  The one piece the reference's own tests DO pin —
// instance.Provider.Create's vmSize / disk-size / name checks — is restated in create_select() and
// checked against those vectors (tests/test_create_select.py).
//
// The Go stdlib `sort` (pdqsort, Go 1.24) is outside the reference tree; sort_slice() below restates
// src/sort/zsortfunc.go from its published algorithm.  It decides the tie order of in-flight
// NodeClaims (scheduler.go:499), which is part of the observable result.
//
// Deliberately reference-shaped: requirement sets are key->Requirement maps holding value sets, the
// filter loop visits one instance type at a time and each offering of it, resources are compared one
// name at a time.  Nothing here shares code or data layout with the device path.

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/b200sched.h"
#include "karp_oracle.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Go sort.Slice — pdqsort_func (Go 1.24 src/sort/zsortfunc.go, src/sort/sort.go)
// ------------------------------------------------------------------------------------------------
struct LessSwap {
  std::function<bool(int, int)> Less;
  std::function<void(int, int)> Swap;
};

enum SortedHint { unknownHint = 0, increasingHint, decreasingHint };

inline int bitsLen(unsigned long long x) {
  int n = 0;
  while (x) { n++; x >>= 1; }
  return n;
}

void insertionSort_func(LessSwap& d, int a, int b) {
  for (int i = a + 1; i < b; i++)
    for (int j = i; j > a && d.Less(j, j - 1); j--) d.Swap(j, j - 1);
}

void siftDown_func(LessSwap& d, int lo, int hi, int first) {
  int root = lo;
  for (;;) {
    int child = 2 * root + 1;
    if (child >= hi) return;
    if (child + 1 < hi && d.Less(first + child, first + child + 1)) child++;
    if (!d.Less(first + root, first + child)) return;
    d.Swap(first + root, first + child);
    root = child;
  }
}

void heapSort_func(LessSwap& d, int a, int b) {
  int first = a, lo = 0, hi = b - a;
  for (int i = (hi - 1) / 2; i >= 0; i--) siftDown_func(d, i, hi, first);
  for (int i = hi - 1; i >= 0; i--) {
    d.Swap(first, first + i);
    siftDown_func(d, lo, i, first);
  }
}

void order2_func(LessSwap& d, int& a, int& b, int* swaps) {
  if (d.Less(b, a)) {
    (*swaps)++;
    std::swap(a, b);
  }
}

int median_func(LessSwap& d, int a, int b, int c, int* swaps) {
  order2_func(d, a, b, swaps);
  order2_func(d, b, c, swaps);
  order2_func(d, a, b, swaps);
  return b;
}

int medianAdjacent_func(LessSwap& d, int a, int* swaps) { return median_func(d, a - 1, a, a + 1, swaps); }

int choosePivot_func(LessSwap& d, int a, int b, SortedHint* hint) {
  const int shortestNinther = 50, maxSwaps = 4 * 3;
  int l = b - a;
  int swaps = 0;
  int i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
  if (l >= 8) {
    if (l >= shortestNinther) {
      i = medianAdjacent_func(d, i, &swaps);
      j = medianAdjacent_func(d, j, &swaps);
      k = medianAdjacent_func(d, k, &swaps);
    }
    j = median_func(d, i, j, k, &swaps);
  }
  if (swaps == 0) *hint = increasingHint;
  else if (swaps == maxSwaps) *hint = decreasingHint;
  else *hint = unknownHint;
  return j;
}

void reverseRange_func(LessSwap& d, int a, int b) {
  int i = a, j = b - 1;
  while (i < j) { d.Swap(i, j); i++; j--; }
}

bool partialInsertionSort_func(LessSwap& d, int a, int b) {
  const int maxSteps = 5, shortestShifting = 50;
  int i = a + 1;
  for (int j = 0; j < maxSteps; j++) {
    while (i < b && !d.Less(i, i - 1)) i++;
    if (i == b) return true;
    if (b - a < shortestShifting) return false;
    d.Swap(i, i - 1);
    if (i - a >= 2) {
      for (int jj = i - 1; jj >= 1; jj--) {
        if (!d.Less(jj, jj - 1)) break;
        d.Swap(jj, jj - 1);
      }
    }
    if (b - i >= 2) {
      for (int jj = i + 1; jj < b; jj++) {
        if (!d.Less(jj, jj - 1)) break;
        d.Swap(jj, jj - 1);
      }
    }
  }
  return false;
}

void breakPatterns_func(LessSwap& d, int a, int b) {
  int length = b - a;
  if (length >= 8) {
    uint64_t random = (uint64_t)length;  // xorshift seeded with the length
    unsigned modulus = 1u << bitsLen((unsigned)length);
    int idx = a + (length / 4) * 2 - 1;
    for (int i = 0; i < 3; i++) {
      random ^= random << 13;
      random ^= random >> 7;
      random ^= random << 17;
      int other = (int)((unsigned)random & (modulus - 1));
      if (other >= length) other -= length;
      d.Swap(idx - 1 + i, a + other);
    }
  }
}

int partitionEqual_func(LessSwap& d, int a, int b, int pivot) {
  d.Swap(a, pivot);
  int i = a + 1, j = b - 1;
  for (;;) {
    while (i <= j && !d.Less(a, i)) i++;
    while (i <= j && d.Less(a, j)) j--;
    if (i > j) break;
    d.Swap(i, j);
    i++; j--;
  }
  return i;
}

int partition_func(LessSwap& d, int a, int b, int pivot, bool* alreadyPartitioned) {
  d.Swap(a, pivot);
  int i = a + 1, j = b - 1;
  while (i <= j && d.Less(i, a)) i++;
  while (i <= j && !d.Less(j, a)) j--;
  if (i > j) {
    d.Swap(j, a);
    *alreadyPartitioned = true;
    return j;
  }
  d.Swap(i, j);
  i++; j--;
  for (;;) {
    while (i <= j && d.Less(i, a)) i++;
    while (i <= j && !d.Less(j, a)) j--;
    if (i > j) break;
    d.Swap(i, j);
    i++; j--;
  }
  d.Swap(j, a);
  *alreadyPartitioned = false;
  return j;
}

void pdqsort_func(LessSwap& d, int a, int b, int limit) {
  const int maxInsertion = 12;
  bool wasBalanced = true, wasPartitioned = true;
  for (;;) {
    int length = b - a;
    if (length <= maxInsertion) { insertionSort_func(d, a, b); return; }
    if (limit == 0) { heapSort_func(d, a, b); return; }
    if (!wasBalanced) { breakPatterns_func(d, a, b); limit--; }
    SortedHint hint;
    int pivot = choosePivot_func(d, a, b, &hint);
    if (hint == decreasingHint) {
      reverseRange_func(d, a, b);
      pivot = (b - 1) - (pivot - a);
      hint = increasingHint;
    }
    if (wasBalanced && wasPartitioned && hint == increasingHint) {
      if (partialInsertionSort_func(d, a, b)) return;
    }
    if (a > 0 && !d.Less(a - 1, pivot)) {
      int mid = partitionEqual_func(d, a, b, pivot);
      a = mid;
      continue;
    }
    bool alreadyPartitioned;
    int mid = partition_func(d, a, b, pivot, &alreadyPartitioned);
    wasPartitioned = alreadyPartitioned;
    int leftLen = mid - a, rightLen = b - mid;
    int balanceThreshold = length / 8;
    if (leftLen < rightLen) {
      wasBalanced = leftLen >= balanceThreshold;
      pdqsort_func(d, a, mid, limit);
      a = mid + 1;
    } else {
      wasBalanced = rightLen >= balanceThreshold;
      pdqsort_func(d, mid + 1, b, limit);
      b = mid;
    }
  }
}

// sort.Slice(x, less) on a vector; `stable` swaps in a stable sort (the H1 sensitivity probe)
template <class T, class Less>
void sort_slice(std::vector<T>& v, Less less, bool stable = false) {
  if (stable) {
    std::stable_sort(v.begin(), v.end(), less);
    return;
  }
  LessSwap d;
  d.Less = [&](int i, int j) { return less(v[i], v[j]); };
  d.Swap = [&](int i, int j) { std::swap(v[i], v[j]); };
  int n = (int)v.size();
  pdqsort_func(d, 0, n, bitsLen((unsigned)n));
}

// ------------------------------------------------------------------------------------------------
// scheduling.Requirement — K/scheduling/requirement.go
// ------------------------------------------------------------------------------------------------
thread_local const b200s_model_t* g_model = nullptr;  // value numerics for withinIntPtrs; set per call, per thread (the
                                                      // bench's reference arm runs oracle_fit_matrix on pod ranges
                                                      // from several threads)

enum Operator { OpIn, OpNotIn, OpExists, OpDoesNotExist };

struct Requirement {
  int key = -1;
  bool complement = false;
  std::vector<int> values;  // sorted value ids (sets.Set[string])
  bool hasGt = false, hasLt = false;
  int64_t gt = 0, lt = 0;
  // MinValues: always nil here (generator constraint A.12; the encoder rejects it)

  bool has(int v) const { return std::binary_search(values.begin(), values.end(), v); }
  // requirement.go:267-278 (Len() < MaxInt64 <=> the excluded set is non-empty)
  Operator op() const {
    if (complement) return values.empty() ? OpExists : OpNotIn;
    return values.empty() ? OpDoesNotExist : OpIn;
  }
};

// requirement.go:47-87
Requirement NewRequirement(int key, int op, const int32_t* vals, int nvals, int64_t bound) {
  Requirement r;
  r.key = key;
  if (op == B200S_OP_IN) {
    r.complement = false;
    r.values.assign(vals, vals + nvals);
    std::sort(r.values.begin(), r.values.end());
    r.values.erase(std::unique(r.values.begin(), r.values.end()), r.values.end());
    return r;
  }
  r.complement = true;
  if (op == B200S_OP_DOESNOTEXIST) r.complement = false;
  if (op == B200S_OP_NOTIN) {
    r.values.assign(vals, vals + nvals);
    std::sort(r.values.begin(), r.values.end());
    r.values.erase(std::unique(r.values.begin(), r.values.end()), r.values.end());
  }
  if (op == B200S_OP_GT) { r.hasGt = true; r.gt = bound; }
  if (op == B200S_OP_LT) { r.hasLt = true; r.lt = bound; }
  return r;
}

// requirement.go:311-327
bool withinIntPtrs(int valueId, bool hasGt, int64_t gt, bool hasLt, int64_t lt) {
  if (!hasGt && !hasLt) return true;
  if (!g_model->val_isint[valueId]) return false;
  int64_t value = g_model->val_int[valueId];
  if (hasGt && gt >= value) return false;
  if (hasLt && lt <= value) return false;
  return true;
}

// requirement.go:158-191
Requirement Intersection(const Requirement& r, const Requirement& o) {
  bool complement = r.complement && o.complement;
  bool hasGt = r.hasGt || o.hasGt, hasLt = r.hasLt || o.hasLt;
  int64_t gt = r.hasGt ? (o.hasGt ? std::max(r.gt, o.gt) : r.gt) : o.gt;
  int64_t lt = r.hasLt ? (o.hasLt ? std::min(r.lt, o.lt) : r.lt) : o.lt;
  if (hasGt && hasLt && gt >= lt) return NewRequirement(r.key, B200S_OP_DOESNOTEXIST, nullptr, 0, 0);
  std::vector<int> values;
  if (r.complement && o.complement) {
    std::set_union(r.values.begin(), r.values.end(), o.values.begin(), o.values.end(), std::back_inserter(values));
  } else if (r.complement && !o.complement) {
    std::set_difference(o.values.begin(), o.values.end(), r.values.begin(), r.values.end(), std::back_inserter(values));
  } else if (!r.complement && o.complement) {
    std::set_difference(r.values.begin(), r.values.end(), o.values.begin(), o.values.end(), std::back_inserter(values));
  } else {
    std::set_intersection(r.values.begin(), r.values.end(), o.values.begin(), o.values.end(), std::back_inserter(values));
  }
  values.erase(std::remove_if(values.begin(), values.end(), [&](int v) { return !withinIntPtrs(v, hasGt, gt, hasLt, lt); }),
               values.end());
  Requirement out;
  out.key = r.key;
  out.values = std::move(values);
  out.complement = complement;
  if (complement) { out.hasGt = hasGt; out.gt = gt; out.hasLt = hasLt; out.lt = lt; }
  return out;
}

// requirement.go:197-231
bool HasIntersection(const Requirement& r, const Requirement& o) {
  bool hasGt = r.hasGt || o.hasGt, hasLt = r.hasLt || o.hasLt;
  int64_t gt = r.hasGt ? (o.hasGt ? std::max(r.gt, o.gt) : r.gt) : o.gt;
  int64_t lt = r.hasLt ? (o.hasLt ? std::min(r.lt, o.lt) : r.lt) : o.lt;
  if (hasGt && hasLt && gt >= lt) return false;
  if (r.complement && o.complement) return true;
  if (r.complement && !o.complement) {
    for (int v : o.values)
      if (!r.has(v) && withinIntPtrs(v, hasGt, gt, hasLt, lt)) return true;
    return false;
  }
  if (!r.complement && o.complement) {
    for (int v : r.values)
      if (!o.has(v) && withinIntPtrs(v, hasGt, gt, hasLt, lt)) return true;
    return false;
  }
  for (int v : r.values)
    if (o.has(v) && withinIntPtrs(v, hasGt, gt, hasLt, lt)) return true;
  return false;
}

// ------------------------------------------------------------------------------------------------
// scheduling.Requirements — K/scheduling/requirements.go
// ------------------------------------------------------------------------------------------------
struct Requirements {
  std::map<int, Requirement> m;

  // requirements.go:127-134
  void Add(const Requirement& req) {
    auto it = m.find(req.key);
    if (it != m.end()) {
      Requirement merged = Intersection(req, it->second);
      it->second = std::move(merged);
    } else {
      m.emplace(req.key, req);
    }
  }
  void AddAll(const Requirements& o) {
    for (auto& kv : o.m) Add(kv.second);
  }
  bool Has(int key) const { return m.count(key) != 0; }
  // requirements.go:154-160: undefined reads as Exists
  Requirement Get(int key) const {
    auto it = m.find(key);
    if (it == m.end()) return NewRequirement(key, B200S_OP_EXISTS, nullptr, 0, 0);
    return it->second;
  }
};

// requirements.go:248-268
bool Intersects(const Requirements& r, const Requirements& incoming) {
  const Requirements* smallest = &r;
  const Requirements* largest = &incoming;
  if (smallest->m.size() > largest->m.size()) std::swap(smallest, largest);
  for (auto& kv : smallest->m) {
    int key = kv.first;
    auto eit = r.m.find(key), iit = incoming.m.find(key);
    if (eit == r.m.end() || iit == incoming.m.end()) continue;
    const Requirement& existing = eit->second;
    const Requirement& inc = iit->second;
    if (!HasIntersection(existing, inc)) {
      Operator oi = inc.op();
      if (oi == OpNotIn || oi == OpDoesNotExist) {
        Operator oe = existing.op();
        if (oe == OpNotIn || oe == OpDoesNotExist) continue;
      }
      return false;
    }
  }
  return true;
}

// requirements.go:175-191.  allowUndefinedWellKnown selects v1.WellKnownLabels as the allow-set.
bool Compatible(const Requirements& r, const Requirements& incoming, bool allowUndefinedWellKnown) {
  for (auto& kv : incoming.m) {
    int key = kv.first;
    if (allowUndefinedWellKnown && g_model->key_wellknown[key]) continue;
    Operator op = kv.second.op();
    if (r.Has(key) || op == OpNotIn || op == OpDoesNotExist) continue;
    return false;  // label does not have known values
  }
  return Intersects(r, incoming);
}

Requirements RequirementsFromModel(const b200s_model_t* M, int off0, int off1) {
  Requirements out;
  for (int i = off0; i < off1; i++) {
    const b200s_req_t& q = M->reqs[i];
    out.Add(NewRequirement(q.key, q.op, M->req_values + q.val_off, q.val_cnt, q.bound));
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// Taints — K/scheduling/taints.go:53-65, vendor/k8s.io/api/core/v1/toleration.go:38-57
// ------------------------------------------------------------------------------------------------
bool ToleratesTaint(const b200s_toleration_t& t, const b200s_taint_t& taint) {
  if (t.effect != B200S_EFFECT_NONE && t.effect != taint.effect) return false;
  if (t.key != 0 && t.key != taint.key) return false;
  switch (t.op) {
    case B200S_TOLOP_EMPTY:
    case B200S_TOLOP_EQUAL: return t.value == taint.value;
    case B200S_TOLOP_EXISTS: return true;
    default: return false;
  }
}

bool Tolerates(const std::vector<b200s_taint_t>& taints, const std::vector<b200s_toleration_t>& tolerations) {
  bool ok = true;
  for (auto& taint : taints) {
    bool tolerates = false;
    for (auto& t : tolerations) tolerates = tolerates || ToleratesTaint(t, taint);
    if (!tolerates) ok = false;
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// resources — K/utils/resources/resources.go.  ResourceLists are dense int64 vectors over the
// model's resource names; "missing" reads as 0 exactly as in Fits (:239-243), and the negative-total
// rule (:234-238) is kept.
// ------------------------------------------------------------------------------------------------
using ResourceList = std::vector<int64_t>;

ResourceList Merge(const ResourceList& a, const ResourceList& b) {  // resources.go:53-67
  ResourceList r(a.size());
  for (size_t i = 0; i < a.size(); i++) r[i] = a[i] + b[i];
  return r;
}

bool Fits(const ResourceList& candidate, const ResourceList& total) {  // resources.go:232-245
  for (int64_t q : total)
    if (0 > q) return false;
  for (size_t i = 0; i < candidate.size(); i++)
    if (candidate[i] > total[i]) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// cloudprovider.InstanceType / Offering — K/cloudprovider/types.go
// ------------------------------------------------------------------------------------------------
struct Offering {
  Requirements requirements;
  double price;
  bool available;
};

struct InstanceType {
  int index;
  Requirements requirements;
  std::vector<Offering> offerings;
  ResourceList allocatable, capacity;
};

struct Counters {
  int64_t fit_evals = 0;       // instance types visited by the filter loop inside Solve
  int64_t prefilter_evals = 0; // ... by NewScheduler's template pre-filter
  int64_t canadd_claim = 0, canadd_node = 0, new_claim_attempts = 0;
  int64_t sorts = 0, sorts_unsorted = 0;
  int64_t price_ties = 0;
  int64_t newclaim_evals = 0;
};

struct FilterError {
  bool requirementsMet = false, fits = false, hasOffering = false;
  bool requirementsAndFits = false, requirementsAndOffering = false, fitsAndOffering = false;
  uint8_t code() const {
    return (uint8_t)(B200S_FIT_FILTER | (requirementsMet ? B200S_FF_REQUIREMENTS : 0) | (fits ? B200S_FF_FITS : 0) |
                     (hasOffering ? B200S_FF_OFFERING : 0) | (requirementsAndFits ? B200S_FF_REQ_AND_FITS : 0) |
                     (requirementsAndOffering ? B200S_FF_REQ_AND_OFF : 0) | (fitsAndOffering ? B200S_FF_FITS_AND_OFF : 0));
  }
};

// nodeclaim.go:373-441 (minValues is nil everywhere: A.12)
std::vector<const InstanceType*> filterInstanceTypesByRequirements(const std::vector<const InstanceType*>& instanceTypes,
                                                                    const Requirements& requirements,
                                                                    const ResourceList& totalRequests, FilterError* err,
                                                                    int64_t* evals) {
  std::vector<const InstanceType*> remaining;
  for (const InstanceType* it : instanceTypes) {
    (*evals)++;
    bool itCompat = Intersects(it->requirements, requirements);  // nodeclaim.go:443-445
    bool itFits = Fits(totalRequests, it->allocatable);          // nodeclaim.go:447-449
    bool itHasOffering = false;
    for (const Offering& of : it->offerings) {
      if (of.available && Compatible(requirements, of.requirements, true)) {
        itHasOffering = true;
        break;
      }
    }
    err->requirementsMet = err->requirementsMet || itCompat;
    err->fits = err->fits || itFits;
    err->hasOffering = err->hasOffering || itHasOffering;
    err->requirementsAndFits = err->requirementsAndFits || (itCompat && itFits && !itHasOffering);
    err->requirementsAndOffering = err->requirementsAndOffering || (itCompat && itHasOffering && !itFits);
    err->fitsAndOffering = err->fitsAndOffering || (itFits && itHasOffering && !itCompat);
    if (itCompat && itFits && itHasOffering) remaining.push_back(it);
  }
  return remaining;
}

// types.go:221-240.  Returns the head of the price-ordered list (the selection the parity unit names).
const InstanceType* OrderByPriceHead(std::vector<const InstanceType*> its, const Requirements& reqs, double* priceOut,
                                     Counters* C, bool stable) {
  auto priceOf = [&](const InstanceType* it) {
    double p = 1.7976931348623157e308;  // math.MaxFloat64
    for (const Offering& of : it->offerings)
      if (of.available && Compatible(reqs, of.requirements, true) && of.price < p) p = of.price;
    return p;
  };
  std::vector<std::pair<double, const InstanceType*>> keyed;
  keyed.reserve(its.size());
  for (auto* it : its) keyed.emplace_back(priceOf(it), it);
  sort_slice(keyed, [](const std::pair<double, const InstanceType*>& a, const std::pair<double, const InstanceType*>& b) {
    return a.first < b.first;
  }, stable);
  if (keyed.empty()) { *priceOut = 1.7976931348623157e308; return nullptr; }
  if (keyed.size() > 1 && keyed[0].first == keyed[1].first) C->price_ties++;
  *priceOut = keyed[0].first;
  return keyed[0].second;
}

// types.go:221-240, the whole list (sort.Slice: Go's pdqsort, so equal prices land in its order)
std::vector<const InstanceType*> OrderByPrice(const std::vector<const InstanceType*>& its, const Requirements& reqs,
                                              bool stable, std::vector<double>* keysOut) {
  std::vector<std::pair<double, const InstanceType*>> keyed;
  keyed.reserve(its.size());
  for (auto* it : its) {
    double p = 1.7976931348623157e308;
    for (const Offering& of : it->offerings)
      if (of.available && Compatible(reqs, of.requirements, true) && of.price < p) p = of.price;
    keyed.emplace_back(p, it);
  }
  sort_slice(keyed, [](const std::pair<double, const InstanceType*>& a, const std::pair<double, const InstanceType*>& b) {
    return a.first < b.first;
  }, stable);
  std::vector<const InstanceType*> out;
  for (auto& kv : keyed) {
    out.push_back(kv.second);
    if (keysOut) keysOut->push_back(kv.first);
  }
  return out;
}

// requirement.go:252-257
bool RequirementHas(const Requirement& r, int valueId) {
  bool in = r.has(valueId);
  return (r.complement ? !in : in) && withinIntPtrs(valueId, r.hasGt, r.gt, r.hasLt, r.lt);
}

// Offerings.Available().WorstLaunchPrice(reqs) — types.go:460-474; ctReqs = Reserved, Spot, OnDemand (types.go:45-47)
double WorstLaunchPrice(const InstanceType* it, const Requirements& reqs, const std::vector<Requirements>& ctReqs) {
  for (const Requirements& ct : ctReqs) {
    bool any = false;
    double top = 0;
    for (const Offering& of : it->offerings) {
      if (!of.available) continue;
      if (!Compatible(reqs, of.requirements, true) || !Compatible(ct, of.requirements, true)) continue;
      if (!any || of.price > top) top = of.price;  // MostExpensive, types.go:454-458
      any = true;
    }
    if (any) return top;
  }
  return 1.7976931348623157e308;
}

// ------------------------------------------------------------------------------------------------
// NodeClaimTemplate / NodeClaim / ExistingNode — K/controllers/provisioning/scheduling
// ------------------------------------------------------------------------------------------------
struct Pod {
  int index;
  Requirements requirements;  // NewPodRequirements (requirements.go:74-110), computed by the caller
  std::vector<b200s_toleration_t> tolerations;
  ResourceList requests;  // RequestsForPods (resources.go:30-38)
  int64_t ts;
  uint64_t uidHi, uidLo;
};

struct NodeClaimTemplate {  // nodeclaimtemplate.go:46-75
  int poolIndex;
  int weight, nameRank;
  Requirements requirements;
  std::vector<b200s_taint_t> taints;
  std::vector<const InstanceType*> instanceTypeOptions;
  ResourceList daemon;
};

struct NodeClaim {
  const NodeClaimTemplate* tmpl;
  int templateOrder;
  Requirements requirements;
  std::vector<const InstanceType*> instanceTypeOptions;
  ResourceList requests;  // Spec.Resources.Requests
  std::vector<int> pods;
};

struct ExistingNode {
  int modelIndex;
  bool initialized;
  int nameRank;
  Requirements requirements;
  std::vector<b200s_taint_t> taints;
  ResourceList remaining;
  std::vector<int> pods;
};

enum AddErr { ErrNone = 0, ErrTaint, ErrIncompatible, ErrFilter, ErrLimits };

struct Scheduler {
  const b200s_model_t* M;
  std::vector<InstanceType> catalog;
  std::vector<std::unique_ptr<NodeClaimTemplate>> templates;
  std::vector<std::unique_ptr<NodeClaim>> claimStore;
  std::vector<NodeClaim*> newNodeClaims;
  std::vector<ExistingNode> existingNodes;
  std::vector<std::vector<uint8_t>> hasLimit;    // remainingResources presence per template
  std::vector<ResourceList> remainingResources;  // per template (keyed by pool in Go)
  bool toleratePreferNoSchedule = false;
  bool stableSort = false;
  Counters C;
  int seq = 0;
  std::vector<int> podSeq, podKind, podIndexOut;
  std::vector<NodeClaim*> podClaim;
  std::vector<uint8_t> podFitCode;  // [P*T]

  // nodeclaim.go:114-163 (topology / host ports / reserved offerings are no-ops: A.12)
  AddErr ClaimCanAdd(const NodeClaim& n, const Pod& pod, const std::vector<b200s_toleration_t>& tolerations,
                     Requirements* updatedReqs, std::vector<const InstanceType*>* updatedITs, FilterError* ferr) {
    if (!Tolerates(n.tmpl->taints, tolerations)) return ErrTaint;
    Requirements nodeClaimRequirements = n.requirements;
    if (!Compatible(nodeClaimRequirements, pod.requirements, true)) return ErrIncompatible;
    nodeClaimRequirements.AddAll(pod.requirements);
    ResourceList requests = Merge(n.requests, pod.requests);
    auto remaining = filterInstanceTypesByRequirements(n.instanceTypeOptions, nodeClaimRequirements, requests, ferr, &C.fit_evals);
    if (remaining.empty()) return ErrFilter;
    *updatedReqs = std::move(nodeClaimRequirements);
    *updatedITs = std::move(remaining);
    return ErrNone;
  }

  // existingnode.go:70-107
  bool NodeCanAdd(const ExistingNode& n, const Pod& pod, const std::vector<b200s_toleration_t>& tolerations,
                  Requirements* updated) {
    C.canadd_node++;
    if (!Tolerates(n.taints, tolerations)) return false;
    if (!Fits(pod.requests, n.remaining)) return false;
    if (!Compatible(n.requirements, pod.requirements, false)) return false;
    Requirements nodeRequirements = n.requirements;
    nodeRequirements.AddAll(pod.requirements);
    *updated = std::move(nodeRequirements);
    return true;
  }

  // scheduler.go:515-550 (one worker: pieces are visited in index order and the first success wins)
  bool addToExistingNode(const Pod& pod, const std::vector<b200s_toleration_t>& tolerations) {
    for (size_t i = 0; i < existingNodes.size(); i++) {
      Requirements r;
      if (NodeCanAdd(existingNodes[i], pod, tolerations, &r)) {
        ExistingNode& n = existingNodes[i];  // existingnode.go:111-119
        n.pods.push_back(pod.index);
        for (size_t k = 0; k < n.remaining.size(); k++) n.remaining[k] -= pod.requests[k];
        n.requirements = std::move(r);
        podKind[pod.index] = B200S_PLACE_NODE;
        podIndexOut[pod.index] = (int)i;
        podSeq[pod.index] = seq++;
        return true;
      }
    }
    return false;
  }

  // scheduler.go:552-584
  bool addToInflightNode(const Pod& pod, const std::vector<b200s_toleration_t>& tolerations) {
    for (size_t i = 0; i < newNodeClaims.size(); i++) {
      Requirements r;
      std::vector<const InstanceType*> its;
      FilterError fe;
      C.canadd_claim++;
      if (ClaimCanAdd(*newNodeClaims[i], pod, tolerations, &r, &its, &fe) == ErrNone) {
        NodeClaim* n = newNodeClaims[i];  // nodeclaim.go:168-180
        n->pods.push_back(pod.index);
        n->instanceTypeOptions = std::move(its);
        n->requests = Merge(n->requests, pod.requests);
        n->requirements = std::move(r);
        podKind[pod.index] = B200S_PLACE_CLAIM;
        podClaim[pod.index] = n;
        podSeq[pod.index] = seq++;
        return true;
      }
    }
    return false;
  }

  // scheduler.go:851-867
  std::vector<const InstanceType*> filterByRemainingResources(const std::vector<const InstanceType*>& its, int t) {
    std::vector<const InstanceType*> filtered;
    for (auto* it : its) {
      bool viable = true;
      for (int r = 0; r < M->n_res; r++)
        if (hasLimit[t][r] && it->capacity[r] > remainingResources[t][r]) viable = false;
      if (viable) filtered.push_back(it);
    }
    return filtered;
  }

  // scheduler.go:587-675
  bool addToNewNodeClaim(const Pod& pod, const std::vector<b200s_toleration_t>& tolerations) {
    int T = (int)templates.size();
    for (int t = 0; t < T; t++) podFitCode[(size_t)pod.index * T + t] = B200S_FIT_NONE;
    for (int i = 0; i < T; i++) {
      NodeClaimTemplate* tmpl = templates[i].get();
      std::vector<const InstanceType*> its = tmpl->instanceTypeOptions;
      bool anyLimit = false;
      for (int r = 0; r < M->n_res; r++) anyLimit = anyLimit || hasLimit[i][r];
      if (anyLimit) {
        its = filterByRemainingResources(its, i);
        if (its.empty()) {
          podFitCode[(size_t)pod.index * T + i] = B200S_FIT_LIMITS;
          continue;
        }
      }
      // NewNodeClaim (nodeclaim.go:83-109): template requirements (+ a unique hostname, which no pod
      // may select on: A.12), requests start at the daemon overhead
      NodeClaim nc;
      nc.tmpl = tmpl;
      nc.templateOrder = i;
      nc.requirements = tmpl->requirements;
      nc.instanceTypeOptions = std::move(its);
      nc.requests = tmpl->daemon;
      Requirements r;
      std::vector<const InstanceType*> upd;
      FilterError fe;
      C.new_claim_attempts++;
      int64_t before = C.fit_evals;
      AddErr e = ClaimCanAdd(nc, pod, tolerations, &r, &upd, &fe);
      C.newclaim_evals += C.fit_evals - before;
      if (e != ErrNone) {
        podFitCode[(size_t)pod.index * T + i] =
            e == ErrTaint ? B200S_FIT_TAINT : e == ErrIncompatible ? B200S_FIT_INCOMPATIBLE : fe.code();
        continue;
      }
      auto owned = std::make_unique<NodeClaim>(std::move(nc));
      NodeClaim* n = owned.get();
      claimStore.push_back(std::move(owned));
      n->pods.push_back(pod.index);
      n->instanceTypeOptions = std::move(upd);
      n->requests = Merge(n->requests, pod.requests);
      n->requirements = std::move(r);
      newNodeClaims.push_back(n);
      podKind[pod.index] = B200S_PLACE_CLAIM;
      podClaim[pod.index] = n;
      podSeq[pod.index] = seq++;
      // subtractMax (scheduler.go:831-848)
      if (anyLimit && !n->instanceTypeOptions.empty()) {
        for (int rr = 0; rr < M->n_res; rr++) {
          if (!hasLimit[i][rr]) continue;
          int64_t mx = n->instanceTypeOptions[0]->capacity[rr];
          for (auto* it : n->instanceTypeOptions) mx = std::max(mx, it->capacity[rr]);
          remainingResources[i][rr] -= mx;
        }
      }
      return true;
    }
    return false;
  }

  // scheduler.go:488-513
  bool add(const Pod& pod, const std::vector<b200s_toleration_t>& tolerations) {
    if (addToExistingNode(pod, tolerations)) return true;
    C.sorts++;
    {
      bool sorted = true;
      for (size_t i = 1; i < newNodeClaims.size(); i++)
        if (newNodeClaims[i]->pods.size() < newNodeClaims[i - 1]->pods.size()) sorted = false;
      if (!sorted) C.sorts_unsorted++;
    }
    sort_slice(newNodeClaims, [](NodeClaim* a, NodeClaim* b) { return a->pods.size() < b->pods.size(); }, stableSort);
    if (addToInflightNode(pod, tolerations)) return true;
    if (templates.empty()) return false;
    return addToNewNodeClaim(pod, tolerations);
  }

  // scheduler.go:434-465 + preferences.go:38-57,132-146.  The only relaxation that can apply to pods
  // without affinities / spreads is the PreferNoSchedule toleration.
  bool trySchedule(const Pod& pod) {
    std::vector<b200s_toleration_t> tolerations = pod.tolerations;  // pod.DeepCopy()
    for (;;) {
      if (add(pod, tolerations)) return true;
      if (!toleratePreferNoSchedule) return false;
      b200s_toleration_t tol{0, 0, B200S_TOLOP_EXISTS, B200S_EFFECT_PREFERNOSCHEDULE};
      bool present = false;
      for (auto& t : tolerations)  // Toleration.MatchToleration
        if (t.key == tol.key && t.effect == tol.effect && t.op == tol.op && t.value == tol.value) present = true;
      if (present) return false;
      tolerations.push_back(tol);
    }
  }
};

void fill_requirements_out(const Requirements& reqs, const b200s_model_t* M, const std::vector<int>& valLocal,
                           uint64_t* maskOut, uint8_t* flagsOut, int64_t* gtOut, int64_t* ltOut, int* overflow) {
  for (int k = 0; k < M->n_keys; k++) { maskOut[k] = 0; flagsOut[k] = 0; gtOut[k] = 0; ltOut[k] = 0; }
  for (auto& kv : reqs.m) {
    const Requirement& r = kv.second;
    uint8_t f = B200S_RF_DEFINED;
    if (r.complement) f |= B200S_RF_COMPLEMENT;
    if (r.hasGt) { f |= B200S_RF_HASGT; gtOut[r.key] = r.gt; }
    if (r.hasLt) { f |= B200S_RF_HASLT; ltOut[r.key] = r.lt; }
    flagsOut[r.key] = f;
    uint64_t m = 0;
    for (int v : r.values) {
      if (valLocal[v] >= 64) { *overflow = 1; continue; }
      m |= 1ull << valLocal[v];
    }
    maskOut[r.key] = m;
  }
}

}  // namespace

// ================================================================================================
// C entry points
// ================================================================================================
extern "C" {

int oracle_version(void) { return 100; }

static void computeConsolidation(Scheduler& S, int nUnscheduled, const oracle_consolidation_in_t* in,
                                 oracle_consolidation_out_t* o);

struct OrderedOut {
  int32_t max_types;
  int32_t* order;
  int32_t* count;
  uint8_t* tie;
};

static int solve_then(const b200s_model_t* M, const oracle_opts_t* opts, oracle_result_t* out,
                      const oracle_consolidation_in_t* cin, oracle_consolidation_out_t* cout,
                      const OrderedOut* ordered = nullptr) {
  g_model = M;
  Scheduler S;
  S.M = M;
  S.stableSort = opts && opts->stable_sort;
  const int R = M->n_res, P = M->n_pods;

  // local bit index of every value inside its key (ascending value id)
  std::vector<int> valLocal(M->n_vals, 0);
  {
    std::vector<int> cnt(M->n_keys, 0);
    for (int v = 0; v < M->n_vals; v++) valLocal[v] = cnt[M->val_key[v]]++;
  }

  // ---- catalog (GetInstanceTypes rows) ----
  S.catalog.resize(M->n_skus);
  for (int s = 0; s < M->n_skus; s++) {
    InstanceType& it = S.catalog[s];
    it.index = s;
    it.requirements = RequirementsFromModel(M, M->sku_req_off[s], M->sku_req_off[s + 1]);
    it.allocatable.assign(M->sku_alloc + (size_t)s * R, M->sku_alloc + (size_t)(s + 1) * R);
    it.capacity.assign(M->sku_capacity + (size_t)s * R, M->sku_capacity + (size_t)(s + 1) * R);
    for (int o = M->sku_off_off[s]; o < M->sku_off_off[s + 1]; o++) {
      Offering of;
      of.requirements = RequirementsFromModel(M, M->off_req_off[o], M->off_req_off[o + 1]);
      of.price = M->off_price[o];
      of.available = M->off_available[o] != 0;
      it.offerings.push_back(std::move(of));
    }
  }
  std::vector<const InstanceType*> allITs;
  for (auto& it : S.catalog) allITs.push_back(&it);

  // ---- NewScheduler (scheduler.go:116-182): OrderByWeight, templates, pre-filter ----
  std::vector<int> poolOrder(M->n_pools);
  for (int i = 0; i < M->n_pools; i++) poolOrder[i] = i;
  sort_slice(poolOrder, [&](int a, int b) {  // nodepool.go:121-131
    if (M->pool_weight[a] == M->pool_weight[b]) return M->pool_name_rank[a] > M->pool_name_rank[b];
    return M->pool_weight[a] > M->pool_weight[b];
  });
  for (int pi : poolOrder) {
    for (int t = M->pool_taint_off[pi]; t < M->pool_taint_off[pi + 1]; t++)
      if (M->taints[t].effect == B200S_EFFECT_PREFERNOSCHEDULE) S.toleratePreferNoSchedule = true;  // scheduler.go:133-140
    auto tmpl = std::make_unique<NodeClaimTemplate>();
    tmpl->poolIndex = pi;
    tmpl->weight = M->pool_weight[pi];
    tmpl->nameRank = M->pool_name_rank[pi];
    tmpl->requirements = RequirementsFromModel(M, M->pool_req_off[pi], M->pool_req_off[pi + 1]);
    tmpl->taints.assign(M->taints + M->pool_taint_off[pi], M->taints + M->pool_taint_off[pi + 1]);
    tmpl->daemon.assign(M->pool_daemon + (size_t)pi * R, M->pool_daemon + (size_t)(pi + 1) * R);
    FilterError fe;
    ResourceList zero(R, 0);
    tmpl->instanceTypeOptions =
        filterInstanceTypesByRequirements(allITs, tmpl->requirements, zero, &fe, &S.C.prefilter_evals);  // scheduler.go:146
    if (tmpl->instanceTypeOptions.empty()) continue;  // scheduler.go:147-156
    std::vector<uint8_t> hl(M->pool_has_limit + (size_t)pi * R, M->pool_has_limit + (size_t)(pi + 1) * R);
    S.hasLimit.push_back(hl);
    S.remainingResources.emplace_back(M->pool_limit + (size_t)pi * R, M->pool_limit + (size_t)(pi + 1) * R);
    S.templates.push_back(std::move(tmpl));
  }
  const int T = (int)S.templates.size();

  // ---- existing nodes (scheduler.go:677-686, :729-742; existingnode.go:40-66) ----
  S.existingNodes.resize(M->n_nodes);
  for (int e = 0; e < M->n_nodes; e++) {
    ExistingNode& n = S.existingNodes[e];
    n.modelIndex = e;
    n.initialized = M->node_initialized[e] != 0;
    n.nameRank = M->node_name_rank[e];
    n.requirements = RequirementsFromModel(M, M->node_req_off[e], M->node_req_off[e + 1]);
    n.taints.assign(M->taints + M->node_taint_off[e], M->taints + M->node_taint_off[e + 1]);
    n.remaining.assign(M->node_available + (size_t)e * R, M->node_available + (size_t)(e + 1) * R);
  }
  std::stable_sort(S.existingNodes.begin(), S.existingNodes.end(), [](const ExistingNode& a, const ExistingNode& b) {
    if (a.initialized && !b.initialized) return true;
    if (!a.initialized && b.initialized) return false;
    return a.nameRank < b.nameRank;
  });

  // ---- pods: updateCachedPodData (scheduler.go:467-486) ----
  std::vector<Pod> pods(P);
  for (int p = 0; p < P; p++) {
    Pod& pod = pods[p];
    pod.index = p;
    pod.requirements = RequirementsFromModel(M, M->pod_req_off[p], M->pod_req_off[p + 1]);
    pod.tolerations.assign(M->tolerations + M->pod_tol_off[p], M->tolerations + M->pod_tol_off[p + 1]);
    pod.requests.assign(M->pod_requests + (size_t)p * R, M->pod_requests + (size_t)(p + 1) * R);
    pod.ts = M->pod_ts[p];
    pod.uidHi = M->pod_uid_hi[p];
    pod.uidLo = M->pod_uid_lo[p];
  }
  S.podSeq.assign(P, -1);
  S.podKind.assign(P, B200S_PLACE_UNSCHEDULED);
  S.podIndexOut.assign(P, -1);
  S.podClaim.assign(P, nullptr);
  S.podFitCode.assign((size_t)P * std::max(T, 1), B200S_FIT_NONE);

  // ---- NewQueue (queue.go:37-43, :72-108) ----
  std::vector<int> queue(P);
  for (int p = 0; p < P; p++) queue[p] = p;
  const int rc = M->res_cpu, rm = M->res_memory;
  sort_slice(queue, [&](int i, int j) {
    const Pod& l = pods[i];
    const Pod& r = pods[j];
    if (l.requests[rc] < r.requests[rc]) return false;
    if (l.requests[rc] > r.requests[rc]) return true;
    if (l.requests[rm] < r.requests[rm]) return false;
    if (l.requests[rm] > r.requests[rm]) return true;
    if (l.ts != r.ts) return l.ts < r.ts;
    if (l.uidHi != r.uidHi) return l.uidHi < r.uidHi;
    return l.uidLo < r.uidLo;
  });
  if (out->queue_order)
    for (int i = 0; i < P; i++) out->queue_order[i] = queue[i];

  // ---- Solve loop (scheduler.go:394-421; queue.go:46-66) ----
  std::vector<int> lastLen(P, 0);
  std::vector<uint8_t> inErr(P, 0);
  size_t head = 0;
  std::vector<int> q = queue;  // q[head:] is q.pods
  for (;;) {
    size_t len = q.size() - head;
    if (len == 0) break;
    int p = q[head];
    if (lastLen[p] == (int)len) break;
    head++;
    if (!S.trySchedule(pods[p])) {
      inErr[p] = 1;
      q.push_back(p);
      lastLen[p] = (int)(q.size() - head);
    } else {
      inErr[p] = 0;
    }
  }

  // ---- results ----
  out->n_templates = T;
  for (int t = 0; t < T && t < B200S_MAX_TEMPLATES; t++) out->template_pool[t] = S.templates[t]->poolIndex;
  for (int e = 0; e < M->n_nodes; e++) {
    if (out->node_order) out->node_order[e] = S.existingNodes[e].modelIndex;
    if (out->node_npods) out->node_npods[e] = (int)S.existingNodes[e].pods.size();
  }
  int nClaims = (int)S.newNodeClaims.size();
  out->n_claims = nClaims;
  out->n_unscheduled = 0;
  for (int p = 0; p < P; p++) out->n_unscheduled += inErr[p];
  std::unordered_map<NodeClaim*, int> claimPos;
  for (int c = 0; c < nClaims; c++) claimPos[S.newNodeClaims[c]] = c;
  const int W = (M->n_skus + 63) / 64;
  int overflow = 0;
  for (int p = 0; p < P; p++) {
    out->pod_place_kind[p] = S.podKind[p];
    out->pod_place_index[p] = S.podKind[p] == B200S_PLACE_CLAIM ? claimPos[S.podClaim[p]] : S.podIndexOut[p];
    out->pod_seq[p] = S.podSeq[p];
    for (int t = 0; t < T; t++) out->pod_fit_code[(size_t)p * T + t] = inErr[p] ? S.podFitCode[(size_t)p * T + t] : (uint8_t)B200S_FIT_NONE;
  }
  if (nClaims > out->cap_claims) {
    g_model = nullptr;
    return B200S_ECAPACITY;
  }
  for (int c = 0; c < nClaims; c++) {
    NodeClaim* n = S.newNodeClaims[c];  // FinalizeScheduling (nodeclaim.go:252-268): nothing but the hostname to drop
    out->claim_template[c] = n->templateOrder;
    out->claim_npods[c] = (int)n->pods.size();
    for (int r = 0; r < R; r++) out->claim_requests[(size_t)c * R + r] = n->requests[r];
    for (int w = 0; w < W; w++) out->claim_mask[(size_t)c * W + w] = 0;
    for (auto* it : n->instanceTypeOptions) out->claim_mask[(size_t)c * W + it->index / 64] |= 1ull << (it->index % 64);
    out->claim_n_options[c] = (int)n->instanceTypeOptions.size();
    fill_requirements_out(n->requirements, M, valLocal, out->claim_req_mask + (size_t)c * M->n_keys,
                          out->claim_req_flags + (size_t)c * M->n_keys, out->claim_req_gt + (size_t)c * M->n_keys,
                          out->claim_req_lt + (size_t)c * M->n_keys, &overflow);
    double price;
    const InstanceType* head = OrderByPriceHead(n->instanceTypeOptions, n->requirements, &price, &S.C, S.stableSort);
    out->claim_sku[c] = head ? head->index : -1;
    uint64_t bits;
    std::memcpy(&bits, &price, 8);
    out->claim_price[c] = bits;
  }
  out->fit_evals = S.C.fit_evals;
  out->prefilter_evals = S.C.prefilter_evals;
  out->canadd_claim = S.C.canadd_claim;
  out->canadd_node = S.C.canadd_node;
  out->new_claim_attempts = S.C.new_claim_attempts;
  out->sorts = S.C.sorts;
  out->sorts_unsorted = S.C.sorts_unsorted;
  out->price_ties = S.C.price_ties;
  out->newclaim_evals = S.C.newclaim_evals;
  out->value_overflow = overflow;
  if (ordered) {  // ToNodeClaim (nodeclaimtemplate.go:77-82)
    for (int c = 0; c < nClaims; c++) {
      NodeClaim* n = S.newNodeClaims[c];
      std::vector<double> keys;
      std::vector<const InstanceType*> its = OrderByPrice(n->instanceTypeOptions, n->requirements, S.stableSort, &keys);
      const int listed = std::min<int>((int)its.size(), ordered->max_types);
      ordered->count[c] = listed;
      ordered->tie[c] = 0;
      for (int i = 0; i < ordered->max_types; i++)
        ordered->order[(size_t)c * ordered->max_types + i] = i < listed ? its[i]->index : -1;
      for (int i = 0; i < listed && i + 1 < (int)its.size(); i++)
        if (keys[i] == keys[i + 1]) ordered->tie[c] = 1;
    }
  }
  if (cin && cout) computeConsolidation(S, out->n_unscheduled, cin, cout);
  g_model = nullptr;
  return 0;
}

int oracle_solve(const b200s_model_t* M, const oracle_opts_t* opts, oracle_result_t* out) {
  return solve_then(M, opts, out, nullptr, nullptr);
}

int oracle_solve_ordered(const b200s_model_t* M, const oracle_opts_t* opts, oracle_result_t* out, int32_t max_types,
                         int32_t* order_out, int32_t* count_out, uint8_t* tie_out) {
  if (max_types < 1 || !order_out || !count_out || !tie_out) return B200S_EINVAL;
  OrderedOut oo{max_types, order_out, count_out, tie_out};
  return solve_then(M, opts, out, nullptr, nullptr, &oo);
}

/* SimulateScheduling's Solve followed by computeConsolidation (K/controllers/disruption/consolidation.go:133-307).
 * Every pod of the model counts as a non-pending pod of the candidates (AllNonPendingPodsScheduled, scheduler.go:326). */
int oracle_consolidate(const b200s_model_t* M, const oracle_opts_t* opts, oracle_result_t* out,
                       const oracle_consolidation_in_t* cin, oracle_consolidation_out_t* cout) {
  if (!cin || !cout || cin->val_spot < 0 || cin->val_on_demand < 0 || M->key_capacity_type < 0) return B200S_EINVAL;
  return solve_then(M, opts, out, cin, cout);
}

static void computeConsolidation(Scheduler& S, int nUnscheduled, const oracle_consolidation_in_t* in,
                                 oracle_consolidation_out_t* o) {
  const b200s_model_t* M = S.M;
  const int W = (M->n_skus + 63) / 64, ctKey = M->key_capacity_type;
  o->decision = B200S_CMD_NONE;
  o->reason = B200S_WHY_OK;
  o->n_options = 0;
  o->spot_required = 0;
  o->price_tie = 0;
  for (int i = 0; i < B200S_S2S_MIN_TYPES; i++) o->head[i] = -1;
  for (int w = 0; w < W; w++) o->mask[w] = 0;
  o->valid = 0;
  std::vector<Requirements> ctReqs;  // reserved, spot, on-demand (types.go:464-468)
  auto report = [&](const std::vector<const InstanceType*>& its) {
    o->n_options = (int)its.size();
    for (int w = 0; w < W; w++) o->mask[w] = 0;
    for (size_t i = 0; i < its.size(); i++) {
      o->mask[its[i]->index / 64] |= 1ull << (its[i]->index % 64);
      if (o->order) o->order[i] = its[i]->index;
    }
  };
  auto finish = [&](std::vector<const InstanceType*> its) {
    o->decision = B200S_CMD_REPLACE;
    if (in->filter_same_type) {
      // filterOutSameInstanceType (multinodeconsolidation.go:172-224): a replacement of one of the removed types must be
      // cheaper than the cheapest removed node of that type, or the command is just a deletion in disguise
      std::unordered_map<int, double> priceByType;
      std::unordered_map<int, bool> existing;
      for (int i = 0; i < in->n_cand_rows; i++) {
        existing[in->cand_sku[i]] = true;
        if (in->cand_price[i] < 0) continue;  // no compatible offering (:185-187)
        auto f = priceByType.find(in->cand_sku[i]);
        double have = f == priceByType.end() ? 1.7976931348623157e308 : f->second;
        if (in->cand_price[i] < have) priceByType[in->cand_sku[i]] = in->cand_price[i];
      }
      double maxPrice = 1.7976931348623157e308;
      for (auto* it : its)
        if (existing.count(it->index)) {
          double p = priceByType.count(it->index) ? priceByType[it->index] : 0.0;  // Go map zero value
          if (p < maxPrice) maxPrice = p;
        }
      std::vector<const InstanceType*> keep;
      for (auto* it : its)
        if (WorstLaunchPrice(it, S.newNodeClaims[0]->requirements, ctReqs) < maxPrice) keep.push_back(it);
      its = std::move(keep);
      o->valid = !its.empty();  // multinodeconsolidation.go:145-148
    }
    report(its);
  };
  if (nUnscheduled > 0) { o->reason = B200S_WHY_UNSCHEDULABLE; return; }          // :147
  if (S.newNodeClaims.empty()) { o->decision = B200S_CMD_DELETE; o->valid = 1; return; }  // :155
  if (S.newNodeClaims.size() != 1) { o->reason = B200S_WHY_MANY_CLAIMS; return; }  // :163
  NodeClaim* n = S.newNodeClaims[0];
  auto single = [&](int v) {
    Requirements r;
    int32_t vv = v;
    r.Add(NewRequirement(ctKey, B200S_OP_IN, &vv, 1, 0));
    return r;
  };
  if (in->val_reserved >= 0) ctReqs.push_back(single(in->val_reserved));
  ctReqs.push_back(single(in->val_spot));
  ctReqs.push_back(single(in->val_on_demand));
  std::vector<double> keys;
  std::vector<const InstanceType*> its = OrderByPrice(n->instanceTypeOptions, n->requirements, S.stableSort, &keys);  // :187
  std::unordered_map<const InstanceType*, double> keyOf;
  for (size_t i = 0; i < its.size(); i++) keyOf[its[i]] = keys[i];
  auto byPrice = [&](const std::vector<const InstanceType*>& in_, const Requirements& reqs) {  // nodeclaim.go:270-274
    std::vector<const InstanceType*> keep;
    for (auto* it : in_)
      if (WorstLaunchPrice(it, reqs, ctReqs) < in->candidate_price) keep.push_back(it);
    return keep;
  };
  if (in->all_spot && RequirementHas(n->requirements.Get(ctKey), in->val_spot)) {  // :189-192 -> :230-307
    if (!in->spot_to_spot) { o->reason = B200S_WHY_S2S_DISABLED; return; }
    int32_t spot = in->val_spot;
    n->requirements.Add(NewRequirement(ctKey, B200S_OP_IN, &spot, 1, 0));  // :240
    std::vector<const InstanceType*> compat;                              // :242, types.go:243-251
    for (auto* it : its) {
      bool has = false;
      for (const Offering& of : it->offerings)
        if (of.available && Compatible(n->requirements, of.requirements, true)) { has = true; break; }
      if (has) compat.push_back(it);
    }
    std::vector<const InstanceType*> cheaper = byPrice(compat, n->requirements);  // :246
    if (cheaper.empty()) { o->reason = B200S_WHY_NOT_CHEAPER; return; }          // :253
    if (in->n_candidates > 1) { o->spot_required = 1; finish(cheaper); return; }  // :262
    if ((int)cheaper.size() < B200S_S2S_MIN_TYPES) { o->reason = B200S_WHY_S2S_TOO_FEW; return; }  // :279
    for (size_t i = 0; i + 1 < cheaper.size() && i < (size_t)B200S_S2S_MIN_TYPES; i++)
      if (keyOf[cheaper[i]] == keyOf[cheaper[i + 1]]) o->price_tie = 1;
    cheaper.resize(B200S_S2S_MIN_TYPES);  // :303 (no minValues)
    for (int i = 0; i < B200S_S2S_MIN_TYPES; i++) o->head[i] = cheaper[i]->index;
    o->spot_required = 1;
    finish(cheaper);
    return;
  }
  std::vector<const InstanceType*> cheaper = byPrice(its, n->requirements);  // :198
  if (cheaper.empty()) { o->reason = B200S_WHY_NOT_CHEAPER; return; }       // :205
  Requirement ctReq = n->requirements.Get(ctKey);                          // :218-221
  if (RequirementHas(ctReq, in->val_spot) && RequirementHas(ctReq, in->val_on_demand)) {
    int32_t spot = in->val_spot;
    n->requirements.Add(NewRequirement(ctKey, B200S_OP_IN, &spot, 1, 0));
    o->spot_required = 1;
  }
  finish(cheaper);
}

// Dense stage A on the CPU, reference-shaped: for every pod run addToNewNodeClaim's template walk
// (scheduler.go:587-675) against fresh NodeClaims.  Same outputs as b200s_fit_matrix.
int oracle_fit_matrix(const b200s_model_t* M, int32_t pod_begin, int32_t pod_end, int32_t* tmpl_out, uint64_t* mask_out,
                      uint8_t* code_out, int64_t* evals_out, int32_t* n_templates_out) {
  g_model = M;
  const int R = M->n_res;
  std::vector<InstanceType> catalog(M->n_skus);
  for (int s = 0; s < M->n_skus; s++) {
    InstanceType& it = catalog[s];
    it.index = s;
    it.requirements = RequirementsFromModel(M, M->sku_req_off[s], M->sku_req_off[s + 1]);
    it.allocatable.assign(M->sku_alloc + (size_t)s * R, M->sku_alloc + (size_t)(s + 1) * R);
    for (int o = M->sku_off_off[s]; o < M->sku_off_off[s + 1]; o++) {
      Offering of;
      of.requirements = RequirementsFromModel(M, M->off_req_off[o], M->off_req_off[o + 1]);
      of.price = M->off_price[o];
      of.available = M->off_available[o] != 0;
      it.offerings.push_back(std::move(of));
    }
  }
  std::vector<const InstanceType*> allITs;
  for (auto& it : catalog) allITs.push_back(&it);
  std::vector<int> poolOrder(M->n_pools);
  for (int i = 0; i < M->n_pools; i++) poolOrder[i] = i;
  sort_slice(poolOrder, [&](int a, int b) {
    if (M->pool_weight[a] == M->pool_weight[b]) return M->pool_name_rank[a] > M->pool_name_rank[b];
    return M->pool_weight[a] > M->pool_weight[b];
  });
  std::vector<NodeClaimTemplate> templates;
  int64_t pre = 0;
  for (int pi : poolOrder) {
    NodeClaimTemplate t;
    t.poolIndex = pi;
    t.requirements = RequirementsFromModel(M, M->pool_req_off[pi], M->pool_req_off[pi + 1]);
    t.taints.assign(M->taints + M->pool_taint_off[pi], M->taints + M->pool_taint_off[pi + 1]);
    t.daemon.assign(M->pool_daemon + (size_t)pi * R, M->pool_daemon + (size_t)(pi + 1) * R);
    FilterError fe;
    ResourceList zero(R, 0);
    t.instanceTypeOptions = filterInstanceTypesByRequirements(allITs, t.requirements, zero, &fe, &pre);
    if (t.instanceTypeOptions.empty()) continue;
    templates.push_back(std::move(t));
  }
  const int T = (int)templates.size();
  const int W = (M->n_skus + 63) / 64;
  if (n_templates_out) *n_templates_out = T;
  int64_t evals = 0;
  for (int p = pod_begin; p < pod_end; p++) {
    size_t o = (size_t)(p - pod_begin);
    Requirements podReqs = RequirementsFromModel(M, M->pod_req_off[p], M->pod_req_off[p + 1]);
    std::vector<b200s_toleration_t> tol(M->tolerations + M->pod_tol_off[p], M->tolerations + M->pod_tol_off[p + 1]);
    ResourceList preq(M->pod_requests + (size_t)p * R, M->pod_requests + (size_t)(p + 1) * R);
    tmpl_out[o] = -1;
    for (int w = 0; w < W; w++) mask_out[o * W + w] = 0;
    for (int t = 0; t < T; t++) code_out[o * T + t] = B200S_FIT_NONE;
    for (int t = 0; t < T; t++) {
      const NodeClaimTemplate& tm = templates[t];
      if (!Tolerates(tm.taints, tol)) { code_out[o * T + t] = B200S_FIT_TAINT; continue; }
      Requirements reqs = tm.requirements;
      if (!Compatible(reqs, podReqs, true)) { code_out[o * T + t] = B200S_FIT_INCOMPATIBLE; continue; }
      reqs.AddAll(podReqs);
      ResourceList total = Merge(tm.daemon, preq);
      FilterError fe;
      auto remaining = filterInstanceTypesByRequirements(tm.instanceTypeOptions, reqs, total, &fe, &evals);
      if (remaining.empty()) { code_out[o * T + t] = fe.code(); continue; }
      tmpl_out[o] = t;
      for (auto* it : remaining) mask_out[o * W + it->index / 64] |= 1ull << (it->index % 64);
      break;
    }
  }
  if (evals_out) *evals_out = evals;
  g_model = nullptr;
  return 0;
}

// Exposed pieces for the property tests ---------------------------------------------------------

// Runs Go's sort.Slice on keys[] (ascending by key) and returns the resulting permutation.
int oracle_go_sort_perm(const int64_t* keys, int32_t n, int32_t* perm_out) {
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  sort_slice(idx, [&](int a, int b) { return keys[a] < keys[b]; });
  for (int i = 0; i < n; i++) perm_out[i] = idx[i];
  return 0;
}

// Requirement algebra on two requirement lists of the model: returns bit0 = Compatible(a,b,allowWK),
// bit1 = Intersects(a,b), and writes Intersection (a Add b) as masks like oracle_solve does.
int oracle_req_algebra(const b200s_model_t* M, int32_t a0, int32_t a1, int32_t b0, int32_t b1, int32_t allow_wellknown,
                       uint64_t* mask_out, uint8_t* flags_out, int64_t* gt_out, int64_t* lt_out) {
  g_model = M;
  std::vector<int> valLocal(M->n_vals, 0);
  {
    std::vector<int> cnt(M->n_keys, 0);
    for (int v = 0; v < M->n_vals; v++) valLocal[v] = cnt[M->val_key[v]]++;
  }
  Requirements A = RequirementsFromModel(M, a0, a1), B = RequirementsFromModel(M, b0, b1);
  int res = (Compatible(A, B, allow_wellknown != 0) ? 1 : 0) | (Intersects(A, B) ? 2 : 0);
  Requirements Cm = A;
  Cm.AddAll(B);
  int overflow = 0;
  fill_requirements_out(Cm, M, valLocal, mask_out, flags_out, gt_out, lt_out, &overflow);
  g_model = nullptr;
  return res;
}

// taints.go:53-65 over model ranges
int oracle_tolerates(const b200s_model_t* M, int32_t taint0, int32_t taint1, int32_t tol0, int32_t tol1) {
  std::vector<b200s_taint_t> ts(M->taints + taint0, M->taints + taint1);
  std::vector<b200s_toleration_t> tl(M->tolerations + tol0, M->tolerations + tol1);
  return Tolerates(ts, tl) ? 1 : 0;
}

// ---- config 1: instance.Provider.Create's selection (pkg/providers/instance/instance.go:76-95,
// newAgentPoolObject :321-369).  AgentPoolNameRegex = ^[a-z][a-z0-9]{0,11}$ (:50).
int oracle_create_select(const char* name, int32_t n_types, const char* const* instance_types, const double* prices,
                         int64_t storage_bytes, int32_t* chosen, int32_t* disk_gb, int32_t* is_gpu) {
  size_t len = std::strlen(name);
  bool ok = len >= 1 && len <= 12 && name[0] >= 'a' && name[0] <= 'z';
  for (size_t i = 1; ok && i < len; i++)
    ok = (name[i] >= 'a' && name[i] <= 'z') || (name[i] >= '0' && name[i] <= '9');
  if (!ok) return B200S_CREATE_EBADNAME;                 // instance.go:81-84
  if (n_types <= 0) return B200S_CREATE_ENOTYPE;          // instance.go:91-93
  int pick = 0;                                           // instance.go:95: vmSize := instanceTypes[0]
  if (prices) {                                           // the price-ordered variant
    for (int i = 1; i < n_types; i++)
      if (prices[i] < prices[pick]) pick = i;
  }
  if (storage_bytes <= 0) return B200S_CREATE_ENOSTORAGE; // instance.go:348-350
  *chosen = pick;
  *disk_gb = (int32_t)(storage_bytes >> 30);              // instance.go:351
  *is_gpu = std::strstr(instance_types[pick], "Standard_N") != nullptr ? 1 : 0;  // instance.go:335-339
  return 0;
}

// ---- resources.RequestsForPods for one pod, reference-shaped: ResourceLists as name->quantity maps
// (K/utils/resources/resources.go:30-38 RequestsForPods, :114-144 podRequests, :187-201 MaxResources,
//  :204-217 MergeResourceLimitsIntoRequests)
int oracle_pod_requests(int32_t R, int32_t res_pods, int32_t n_containers, const int64_t* req, const uint8_t* req_has,
                        const int64_t* lim, const uint8_t* lim_has, int32_t n_init, const int64_t* init_req,
                        const uint8_t* init_req_has, const int64_t* init_lim, const uint8_t* init_lim_has,
                        const uint8_t* init_restart_always, const int64_t* overhead, int64_t* out) {
  using RL = std::map<int, int64_t>;
  auto mergeInto = [](RL& dest, const RL& src) { for (auto& kv : src) dest[kv.first] += kv.second; };
  auto maxResources = [](const RL& a, const RL& b) {
    RL r;
    for (const RL* l : {&a, &b})
      for (auto& kv : *l) {
        auto it = r.find(kv.first);
        if (it == r.end() || kv.second > it->second) r[kv.first] = kv.second;
      }
    return r;
  };
  auto container = [&](const int64_t* q, const uint8_t* qh, const int64_t* l, const uint8_t* lh, int i) {
    RL ret;
    for (int r = 0; r < R; r++) if (qh[(size_t)i * R + r]) ret[r] = q[(size_t)i * R + r];
    if (lh) for (int r = 0; r < R; r++) if (lh[(size_t)i * R + r] && !qh[(size_t)i * R + r]) ret[r] = l[(size_t)i * R + r];
    return ret;
  };
  RL requests, restartable, maxInit;
  for (int i = 0; i < n_containers; i++) mergeInto(requests, container(req, req_has, lim, lim_has, i));
  for (int i = 0; i < n_init; i++) {
    RL c = container(init_req, init_req_has, init_lim, init_lim_has, i);
    if (init_restart_always && init_restart_always[i]) {
      mergeInto(requests, c);
      mergeInto(restartable, c);
      maxInit = maxResources(maxInit, restartable);
    } else {
      RL sum = c;
      mergeInto(sum, restartable);
      maxInit = maxResources(maxInit, sum);
    }
  }
  requests = maxResources(requests, maxInit);
  if (overhead) { RL ov; for (int r = 0; r < R; r++) if (overhead[r]) ov[r] = overhead[r]; mergeInto(requests, ov); }
  requests[res_pods] = 1;
  for (int r = 0; r < R; r++) { auto it = requests.find(r); out[r] = it == requests.end() ? 0 : it->second; }
  return 0;
}

// ---- InstanceType.Allocatable (K/cloudprovider/types.go:181-214)
int oracle_allocatable(int32_t R, int32_t res_memory, const int64_t* capacity, const uint8_t* capacity_has,
                       const int64_t* kube, const int64_t* system, const int64_t* eviction, const uint8_t* is_hugepages,
                       int64_t* out) {
  std::map<int, int64_t> total, alloc;  // Overhead.Total() = Merge(kube, system, eviction) (types.go:343-345)
  for (int r = 0; r < R; r++) total[r] = (kube ? kube[r] : 0) + (system ? system[r] : 0) + (eviction ? eviction[r] : 0);
  for (int r = 0; r < R; r++) if (capacity_has[r]) alloc[r] = capacity[r] - total[r];  // resources.Subtract
  for (int r = 0; r < R; r++)
    if (capacity_has[r] && is_hugepages && is_hugepages[r]) {
      int64_t current = alloc.count(res_memory) ? alloc[res_memory] : 0;
      current -= capacity[r];
      if (current < 0) current = 0;
      alloc[res_memory] = current;
    }
  for (int r = 0; r < R; r++) out[r] = alloc.count(r) ? alloc[r] : 0;
  return 0;
}

}  // extern "C"

"""A second restatement of Scheduler.Solve, written in Python from the Go text and in the reference's own shape —
requirement maps of value sets, one instance type and one offering at a time, lists of NodeClaim objects re-sorted with a
port of Go's pdqsort before every pod.  It shares no code with oracle/karp_oracle.cpp; tests/test_solve_ref.py fuzzes the
two against each other, which is what stands in for the reference's own (vendored-away) scheduler tests.

Paths: K/ = vendor/sigs.k8s.io/karpenter/pkg/ of the reference tree.  Scope = what the model can express: no topology,
host ports, volumes, reserved offerings, minValues or preference relaxation.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np

from b200sched import model as M

MAX_INT64 = (1 << 63) - 1


# ---- scheduling.Requirement (K/scheduling/requirement.go) ---------------------------------------------------------------
class Requirement:
    __slots__ = ("key", "values", "complement", "gt", "lt")

    def __init__(self, key, values=(), complement=False, gt=None, lt=None):
        self.key, self.values, self.complement, self.gt, self.lt = key, set(values), complement, gt, lt


class Ctx:
    """what turns a value id into the integer Atoi would give (withinIntPtrs, requirement.go:309-327)"""

    def __init__(self, model: M.Model):
        a = model.arrays
        self.isint = a["val_isint"]
        self.int = a["val_int"]
        self.wellknown = set(int(k) for k in np.nonzero(a["key_wellknown"])[0])

    def within(self, value, gt, lt) -> bool:
        if gt is None and lt is None:
            return True
        if not self.isint[value]:  # "If bounds are set, non integer values are invalid"
            return False
        v = int(self.int[value])
        if gt is not None and gt >= v:
            return False
        if lt is not None and lt <= v:
            return False
        return True


def new_requirement(key: int, op: int, values=(), bound: Optional[int] = None) -> Requirement:
    """NewRequirementWithFlexibility (requirement.go:47-87), minValues nil"""
    if op == M.OP_IN:
        return Requirement(key, values, False)
    r = Requirement(key, (), True)
    if op == M.OP_DOESNOTEXIST:
        r.complement = False
    if op == M.OP_NOTIN:
        r.values.update(values)
    if op == M.OP_GT:
        r.gt = bound
    if op == M.OP_LT:
        r.lt = bound
    return r


def _max_ptr(a, b):
    return b if a is None else a if b is None else max(a, b)


def _min_ptr(a, b):
    return b if a is None else a if b is None else min(a, b)


def req_len(r: Requirement) -> int:  # requirement.go:279-284
    return MAX_INT64 - len(r.values) if r.complement else len(r.values)


def operator(r: Requirement) -> int:  # requirement.go:266-277
    if r.complement:
        return M.OP_NOTIN if req_len(r) < MAX_INT64 else M.OP_EXISTS
    return M.OP_IN if req_len(r) > 0 else M.OP_DOESNOTEXIST


def intersection(cx: Ctx, r: Requirement, o: Requirement) -> Requirement:  # requirement.go:146-182
    complement = r.complement and o.complement
    gt, lt = _max_ptr(r.gt, o.gt), _min_ptr(r.lt, o.lt)
    if gt is not None and lt is not None and gt >= lt:
        return new_requirement(r.key, M.OP_DOESNOTEXIST)
    if r.complement and o.complement:
        values = r.values | o.values
    elif r.complement and not o.complement:
        values = o.values - r.values
    elif not r.complement and o.complement:
        values = r.values - o.values
    else:
        values = r.values & o.values
    values = {v for v in values if cx.within(v, gt, lt)}
    if not complement:
        gt, lt = None, None
    return Requirement(r.key, values, complement, gt, lt)


def has_intersection(cx: Ctx, r: Requirement, o: Requirement) -> bool:  # requirement.go:188-227
    gt, lt = _max_ptr(r.gt, o.gt), _min_ptr(r.lt, o.lt)
    if gt is not None and lt is not None and gt >= lt:
        return False
    if r.complement and o.complement:
        return True
    if r.complement and not o.complement:
        return any(v not in r.values and cx.within(v, gt, lt) for v in o.values)
    if not r.complement and o.complement:
        return any(v not in o.values and cx.within(v, gt, lt) for v in r.values)
    return any(v in o.values and cx.within(v, gt, lt) for v in r.values)


# ---- scheduling.Requirements (K/scheduling/requirements.go) -------------------------------------------------------------
class Requirements(dict):
    def add(self, cx: Ctx, *reqs: Requirement):  # :127-134
        for r in reqs:
            if r.key in self:
                r = intersection(cx, r, self[r.key])
            self[r.key] = r

    def get(self, key) -> Requirement:  # :154-160: "If not defined, allow any values with the exists operator"
        return self[key] if key in self else new_requirement(key, M.OP_EXISTS)

    def copy_of(self, cx: Ctx) -> "Requirements":  # scheduling.NewRequirements(r.Values()...)
        out = Requirements()
        out.add(cx, *self.values())
        return out


def intersects(cx: Ctx, r: Requirements, o: Requirements) -> bool:  # :259-280 (nil error <=> True)
    for key in (k for k in r if k in o):
        existing, incoming = r.get(key), o.get(key)
        if not has_intersection(cx, existing, incoming):
            if operator(incoming) in (M.OP_NOTIN, M.OP_DOESNOTEXIST) and operator(existing) in (M.OP_NOTIN, M.OP_DOESNOTEXIST):
                continue
            return False
    return True


def compatible(cx: Ctx, r: Requirements, o: Requirements, allow_undefined_wellknown: bool) -> bool:  # :176-192
    for key in o:
        if allow_undefined_wellknown and key in cx.wellknown:
            continue
        if key in r or operator(o.get(key)) in (M.OP_NOTIN, M.OP_DOESNOTEXIST):
            continue
        return False  # "label does not have known values"
    return intersects(cx, r, o)


def requirements_from(cx: Ctx, model: M.Model, lo: int, hi: int) -> Requirements:
    a = model.arrays
    out = Requirements()
    for i in range(lo, hi):
        q = a["reqs"][i]
        vals = [int(v) for v in a["req_values"][int(q["val_off"]):int(q["val_off"]) + int(q["val_cnt"])]]
        out.add(cx, new_requirement(int(q["key"]), int(q["op"]), vals, int(q["bound"])))
    return out


# ---- taints (K/scheduling/taints.go:48-64, k8s.io/api/core/v1 Toleration.ToleratesTaint) --------------------------------
def tolerates_taint(tol, taint) -> bool:
    tkey, tval, top, teff = tol
    key, val, eff = taint
    if teff != M.EFFECT_NONE and teff != eff:
        return False
    if tkey != 0 and tkey != key:  # len(t.Key) > 0 (the model interns "" as string 0)
        return False
    if top in (M.TOLOP_EMPTY, M.TOLOP_EQUAL):
        return tval == val
    return top == M.TOLOP_EXISTS


def tolerates(taints, tolerations) -> bool:
    return all(any(tolerates_taint(t, taint) for t in tolerations) for taint in taints)


# ---- resources (K/utils/resources/resources.go:232-247) -----------------------------------------------------------------
def fits(candidate, total) -> bool:
    if any(q < 0 for q in total):
        return False
    return all(c <= t for c, t in zip(candidate, total))


# ---- sort.Slice: Go's pdqsort_func (src/sort/zsortfunc.go; the Go standard library is not in the reference tree) ---------
class _Sorter:
    def __init__(self, data: list, less):
        self.d, self.lessf = data, less

    def less(self, i, j):
        return self.lessf(self.d[i], self.d[j])

    def swap(self, i, j):
        self.d[i], self.d[j] = self.d[j], self.d[i]

    def insertion(self, a, b):
        for i in range(a + 1, b):
            j = i
            while j > a and self.less(j, j - 1):
                self.swap(j, j - 1)
                j -= 1

    def sift_down(self, lo, hi, first):
        root = lo
        while True:
            child = 2 * root + 1
            if child >= hi:
                return
            if child + 1 < hi and self.less(first + child, first + child + 1):
                child += 1
            if not self.less(first + root, first + child):
                return
            self.swap(first + root, first + child)
            root = child

    def heap_sort(self, a, b):
        first, lo, hi = a, 0, b - a
        for i in range((hi - 1) // 2, -1, -1):
            self.sift_down(i, hi, first)
        for i in range(hi - 1, -1, -1):
            self.swap(first, first + i)
            self.sift_down(lo, i, first)

    def order2(self, a, b, swaps):
        if self.less(b, a):
            swaps[0] += 1
            return b, a
        return a, b

    def median(self, a, b, c, swaps):
        a, b = self.order2(a, b, swaps)
        b, c = self.order2(b, c, swaps)
        a, b = self.order2(a, b, swaps)
        return b

    def choose_pivot(self, a, b):
        ln = b - a
        swaps = [0]
        i, j, k = a + ln // 4 * 1, a + ln // 4 * 2, a + ln // 4 * 3
        if ln >= 8:
            if ln >= 50:
                i = self.median(i - 1, i, i + 1, swaps)
                j = self.median(j - 1, j, j + 1, swaps)
                k = self.median(k - 1, k, k + 1, swaps)
            j = self.median(i, j, k, swaps)
        return j, ("inc" if swaps[0] == 0 else "dec" if swaps[0] == 12 else "unknown")

    def reverse_range(self, a, b):
        i, j = a, b - 1
        while i < j:
            self.swap(i, j)
            i, j = i + 1, j - 1

    def partial_insertion(self, a, b) -> bool:
        i = a + 1
        for _ in range(5):
            while i < b and not self.less(i, i - 1):
                i += 1
            if i == b:
                return True
            if b - a < 50:
                return False
            self.swap(i, i - 1)
            if i - a >= 2:
                j = i - 1
                while j >= 1:
                    if not self.less(j, j - 1):
                        break
                    self.swap(j, j - 1)
                    j -= 1
            if b - i >= 2:
                j = i + 1
                while j < b:
                    if not self.less(j, j - 1):
                        break
                    self.swap(j, j - 1)
                    j += 1
        return False

    def break_patterns(self, a, b):
        length = b - a
        if length >= 8:
            rnd = length
            modulus = 1 << length.bit_length()
            idx = a + (length // 4) * 2 - 1
            for i in range(3):
                rnd ^= (rnd << 13) & 0xFFFFFFFFFFFFFFFF
                rnd ^= rnd >> 7
                rnd ^= (rnd << 17) & 0xFFFFFFFFFFFFFFFF
                other = rnd & (modulus - 1)
                if other >= length:
                    other -= length
                self.swap(idx - 1 + i, a + other)

    def partition_equal(self, a, b, pivot):
        self.swap(a, pivot)
        i, j = a + 1, b - 1
        while True:
            while i <= j and not self.less(a, i):
                i += 1
            while i <= j and self.less(a, j):
                j -= 1
            if i > j:
                break
            self.swap(i, j)
            i, j = i + 1, j - 1
        return i

    def partition(self, a, b, pivot):
        self.swap(a, pivot)
        i, j = a + 1, b - 1
        while i <= j and self.less(i, a):
            i += 1
        while i <= j and not self.less(j, a):
            j -= 1
        if i > j:
            self.swap(j, a)
            return j, True
        self.swap(i, j)
        i, j = i + 1, j - 1
        while True:
            while i <= j and self.less(i, a):
                i += 1
            while i <= j and not self.less(j, a):
                j -= 1
            if i > j:
                break
            self.swap(i, j)
            i, j = i + 1, j - 1
        self.swap(j, a)
        return j, False

    def pdqsort(self, a, b, limit):
        was_balanced, was_partitioned = True, True
        while True:
            length = b - a
            if length <= 12:
                self.insertion(a, b)
                return
            if limit == 0:
                self.heap_sort(a, b)
                return
            if not was_balanced:
                self.break_patterns(a, b)
                limit -= 1
            pivot, hint = self.choose_pivot(a, b)
            if hint == "dec":
                self.reverse_range(a, b)
                pivot = (b - 1) - (pivot - a)
                hint = "inc"
            if was_balanced and was_partitioned and hint == "inc":
                if self.partial_insertion(a, b):
                    return
            if a > 0 and not self.less(a - 1, pivot):
                a = self.partition_equal(a, b, pivot)
                continue
            mid, already = self.partition(a, b, pivot)
            was_partitioned = already
            left, right = mid - a, b - mid
            balance = length // 8
            if left < right:
                was_balanced = left >= balance
                self.pdqsort(a, mid, limit)
                a = mid + 1
            else:
                was_balanced = right >= balance
                self.pdqsort(mid + 1, b, limit)
                b = mid


def sort_slice(data: list, less) -> None:
    n = len(data)
    _Sorter(data, less).pdqsort(0, n, n.bit_length())


# ---- the scheduler ------------------------------------------------------------------------------------------------------
class InstanceType:
    def __init__(self, idx, requirements, allocatable, capacity, offerings):
        self.idx, self.requirements, self.allocatable, self.capacity, self.offerings = idx, requirements, allocatable, capacity, offerings


class Template:
    def __init__(self, pool, requirements, taints, daemon):
        self.pool, self.requirements, self.taints, self.daemon = pool, requirements, taints, daemon
        self.instance_types: List[InstanceType] = []


class NodeClaim:
    def __init__(self, template: Template, instance_types, requirements):
        self.template, self.instance_types, self.requirements = template, instance_types, requirements
        self.requests = list(template.daemon)
        self.pods: List[int] = []


class ExistingNode:
    def __init__(self, idx, requirements, taints, remaining):
        self.idx, self.requirements, self.taints, self.remaining = idx, requirements, taints, list(remaining)
        self.pods: List[int] = []


def filter_instance_types(cx: Ctx, its, requirements: Requirements, total_requests):
    """filterInstanceTypesByRequirements (K/controllers/provisioning/scheduling/nodeclaim.go:373-441), minValues aside"""
    remaining = []
    for it in its:
        it_compat = intersects(cx, it.requirements, requirements)                      # compatible(), :443-445
        it_fits = fits(total_requests, it.allocatable)                                   # fits(), :447-449
        it_offering = any(avail and compatible(cx, requirements, oreq, True) for oreq, _, avail in it.offerings)
        if it_compat and it_fits and it_offering:
            remaining.append(it)
    return remaining


def solve(model: M.Model):
    cx = Ctx(model)
    a = model.arrays
    flat = {k: a[k].reshape(-1) for k in ("sku_alloc", "sku_capacity", "pool_daemon", "pool_limit", "pool_has_limit",
                                         "node_available", "pod_requests")}
    c = model.c
    R, P, S, T, E = int(c.n_res), int(c.n_pods), int(c.n_skus), int(c.n_pools), int(c.n_nodes)
    its = []
    for s in range(S):
        offs = []
        for o in range(int(a["sku_off_off"][s]), int(a["sku_off_off"][s + 1])):
            offs.append((requirements_from(cx, model, int(a["off_req_off"][o]), int(a["off_req_off"][o + 1])),
                         float(a["off_price"][o]), bool(a["off_available"][o])))
        its.append(InstanceType(s, requirements_from(cx, model, int(a["sku_req_off"][s]), int(a["sku_req_off"][s + 1])),
                                [int(x) for x in flat["sku_alloc"][s * R:(s + 1) * R]],
                                [int(x) for x in flat["sku_capacity"][s * R:(s + 1) * R]], offs))
    taint_rows = [(int(t["key"]), int(t["value"]), int(t["effect"])) for t in a["taints"]]
    tol_rows = [(int(t["key"]), int(t["value"]), int(t["op"]), int(t["effect"])) for t in a["tolerations"]]

    # nodepool.OrderByWeight (K/utils/nodepool/nodepool.go:121-131): weight descending, then name descending
    pools = sorted(range(T), key=lambda p: (-int(a["pool_weight"][p]), -int(a["pool_name_rank"][p])))
    # NewScheduler (scheduler.go:143-158): templates whose requirements leave no instance type are dropped
    templates: List[Template] = []
    for p in pools:
        t = Template(p, requirements_from(cx, model, int(a["pool_req_off"][p]), int(a["pool_req_off"][p + 1])),
                     taint_rows[int(a["pool_taint_off"][p]):int(a["pool_taint_off"][p + 1])],
                     [int(x) for x in flat["pool_daemon"][p * R:(p + 1) * R]])
        t.instance_types = filter_instance_types(cx, its, t.requirements, [0] * R)
        if t.instance_types:
            templates.append(t)
    # remainingResources (scheduler.go:170-172): the limited resources of every NodePool
    remaining: Dict[int, Dict[int, int]] = {
        p: {r: int(flat["pool_limit"][p * R + r]) for r in range(R) if flat["pool_has_limit"][p * R + r]} for p in range(T)}

    # existing nodes, sortExistingNodes (scheduler.go:729-742): initialized first, then by name
    node_order = sorted(range(E), key=lambda e: (0 if a["node_initialized"][e] else 1, int(a["node_name_rank"][e])))
    nodes = [ExistingNode(e, requirements_from(cx, model, int(a["node_req_off"][e]), int(a["node_req_off"][e + 1])),
                          taint_rows[int(a["node_taint_off"][e]):int(a["node_taint_off"][e + 1])],
                          [int(x) for x in flat["node_available"][e * R:(e + 1) * R]]) for e in node_order]

    pod_reqs = [requirements_from(cx, model, int(a["pod_req_off"][p]), int(a["pod_req_off"][p + 1])) for p in range(P)]
    pod_tols = [tol_rows[int(a["pod_tol_off"][p]):int(a["pod_tol_off"][p + 1])] for p in range(P)]
    pod_requests = [[int(x) for x in flat["pod_requests"][p * R:(p + 1) * R]] for p in range(P)]
    cpu, mem = int(c.res_cpu), int(c.res_memory)
    # NewQueue (queue.go:37-43, :72-108): a total order, so the sorting algorithm does not matter
    queue = sorted(range(P), key=lambda p: (-pod_requests[p][cpu], -pod_requests[p][mem], int(a["pod_ts"][p]),
                                            int(a["pod_uid_hi"][p]), int(a["pod_uid_lo"][p])))
    last_len: Dict[int, int] = {}
    claims: List[NodeClaim] = []
    kind = np.full(P, -1, np.int32)
    seq = np.full(P, -1, np.int32)
    placed_on: Dict[int, object] = {}
    n_seq = 0

    def claim_can_add(nc: NodeClaim, p: int):  # NodeClaim.CanAdd (nodeclaim.go:114-163)
        if not tolerates(nc.template.taints, pod_tols[p]):
            return None
        reqs = nc.requirements.copy_of(cx)
        if not compatible(cx, reqs, pod_reqs[p], True):
            return None
        reqs.add(cx, *pod_reqs[p].values())
        total = [x + y for x, y in zip(nc.requests, pod_requests[p])]
        left = filter_instance_types(cx, nc.instance_types, reqs, total)
        if not left:
            return None
        return reqs, left

    def add(p: int) -> bool:  # Scheduler.add (scheduler.go:488-516)
        # addToExistingNode (:518-553)
        for n in nodes:
            if not tolerates(n.taints, pod_tols[p]):
                continue
            if not fits(pod_requests[p], n.remaining):
                continue
            if not compatible(cx, n.requirements, pod_reqs[p], False):
                continue
            reqs = n.requirements.copy_of(cx)
            reqs.add(cx, *pod_reqs[p].values())
            n.pods.append(p)
            n.remaining = [x - y for x, y in zip(n.remaining, pod_requests[p])]
            n.requirements = reqs
            placed_on[p] = n
            return True
        sort_slice(claims, lambda x, y: len(x.pods) < len(y.pods))  # :499
        # addToInflightNode (:555-585)
        for nc in claims:
            got = claim_can_add(nc, p)
            if got is not None:
                nc.pods.append(p)
                nc.requirements, nc.instance_types = got
                nc.requests = [x + y for x, y in zip(nc.requests, pod_requests[p])]
                placed_on[p] = nc
                return True
        # addToNewNodeClaim (:587-675)
        for t in templates:
            cand = t.instance_types
            rem = remaining[t.pool]
            cand = [it for it in cand if not any(it.capacity[r] > q for r, q in rem.items())]  # filterByRemainingResources (:849-867)
            if not cand:
                continue
            nc = NodeClaim(t, cand, t.requirements.copy_of(cx))
            got = claim_can_add(nc, p)
            if got is None:
                continue
            nc.pods.append(p)
            nc.requirements, nc.instance_types = got
            nc.requests = [x + y for x, y in zip(nc.requests, pod_requests[p])]
            claims.append(nc)
            # subtractMax (:831-847)
            remaining[t.pool] = {r: q - max(it.capacity[r] for it in nc.instance_types) for r, q in rem.items()}
            placed_on[p] = nc
            return True
        return False

    # Solve (:377-432) with Queue.Pop / Push (queue.go:46-66)
    while queue:
        p = queue[0]
        if last_len.get(p, 0) == len(queue):
            break
        queue.pop(0)
        if add(p):
            seq[p] = n_seq
            n_seq += 1
        else:
            queue.append(p)
            last_len[p] = len(queue)

    index = np.full(P, -1, np.int32)
    pos_of_claim = {id(nc): i for i, nc in enumerate(claims)}
    pos_of_node = {id(n): i for i, n in enumerate(nodes)}
    for p, where in placed_on.items():
        if isinstance(where, NodeClaim):
            kind[p], index[p] = 1, pos_of_claim[id(where)]
        else:
            kind[p], index[p] = 0, pos_of_node[id(where)]
    W = (S + 63) // 64
    mask = np.zeros((len(claims), W), np.uint64)
    sku = np.full(len(claims), -1, np.int32)
    orders = []
    for i, nc in enumerate(claims):
        for it in nc.instance_types:
            mask[i, it.idx // 64] |= np.uint64(1) << np.uint64(it.idx % 64)
        # InstanceTypes.OrderByPrice (K/cloudprovider/types.go:221-240): sort.Slice — the unstable pdqsort again — by
        # the cheapest available offering the requirements admit (math.MaxFloat64 when there is none); equal prices end
        # up where pdqsort leaves them.  NodeClaimTemplate.ToNodeClaim keeps the first MaxInstanceTypes
        # (nodeclaimtemplate.go:77-82), the parity unit's cheapest SKU is element 0.
        keyed = [(min([pr for oreq, pr, avail in it.offerings if avail and compatible(cx, nc.requirements, oreq, True)],
                      default=1.7976931348623157e308), it.idx) for it in nc.instance_types]
        sort_slice(keyed, lambda x, y: x[0] < y[0])
        orders.append([idx for _, idx in keyed])
        sku[i] = keyed[0][1] if keyed else -1
    return dict(n_claims=len(claims), n_unscheduled=int((kind < 0).sum()), pod_place_kind=kind, pod_place_index=index,
                pod_seq=seq, claim_pool=np.array([nc.template.pool for nc in claims], np.int32),
                claim_npods=np.array([len(nc.pods) for nc in claims], np.int32),
                claim_requests=np.array([nc.requests for nc in claims], np.int64).reshape(len(claims), R),
                claim_mask=mask, claim_sku=sku, claim_order=orders, node_npods=np.array([len(n.pods) for n in nodes], np.int32))

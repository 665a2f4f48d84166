"""The oracle itself: Go sort port, requirement algebra, tolerations, queue semantics, and the one piece the
reference's own tests pin — instance.Provider.Create's selection (config 1)."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
from b200sched import model as M
from b200sched import synth

RES = ["cpu", "memory", "pods"]


def test_go_sort_is_a_sort_and_a_permutation():
    rng = np.random.default_rng(1)
    for n in [0, 1, 2, 5, 12, 13, 20, 49, 50, 51, 100, 257, 1000, 5000]:
        for hi in (2, 5, 1000):
            keys = rng.integers(0, hi, n)
            perm = ob.go_sort_perm(keys)
            assert sorted(perm.tolist()) == list(range(n))
            assert (np.diff(keys[perm]) >= 0).all()


def test_go_sort_insertion_range_is_stable():
    # n <= 12 takes insertionSort (zsortfunc.go maxInsertion = 12), which is stable
    rng = np.random.default_rng(2)
    for n in range(0, 13):
        keys = rng.integers(0, 3, n)
        perm = ob.go_sort_perm(keys)
        assert perm.tolist() == np.argsort(keys, kind="stable").tolist()


def test_go_sort_single_displacement_reinserts_next_to_equals_when_n_ge_50():
    # one element bumped by one in a sorted array, n >= 50, the bump away from the sampled pivot triples:
    # pdqsort takes partialInsertionSort and the result equals a stable sort (DESIGN.md "tie order")
    n = 200
    keys = np.repeat(np.arange(1, 11), 20).astype(np.int64)
    for pos in (3, 17, 33, 61, 77, 123, 161, 188):
        k = keys.copy()
        k[pos] += 1
        perm = ob.go_sort_perm(k)
        assert perm.tolist() == np.argsort(k, kind="stable").tolist()


def test_go_sort_ties_can_permute_in_the_mid_window():
    # 13 <= n < 50 and unsorted: real partitioning happens and ties are NOT stable in general.  This pins the
    # port's behaviour on one fixed input so that accidental edits to the sort show up.
    keys = np.array([1] * 10 + [2] * 10 + [1], dtype=np.int64)
    perm = ob.go_sort_perm(keys)
    assert (np.diff(keys[perm]) >= 0).all()
    assert perm.tolist() != np.argsort(keys, kind="stable").tolist()


def _tiny_builder():
    b = M.ModelBuilder(RES)
    b.add_sku("a", [(M.KEY_ARCH, M.OP_IN, ["amd64"])], [1000, 1 << 30, 10], [1000, 1 << 30, 10],
              [([(M.KEY_ZONE, M.OP_IN, ["z1"]), (M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand"])], 1.0, True)])
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    return b


def _algebra(b, reqs_a, reqs_b, allow=True):
    """Compatible / Intersects / Add of two requirement lists through the oracle."""
    pa = b.add_pod(reqs_a, [0, 0, 1])
    pb = b.add_pod(reqs_b, [0, 0, 1])
    m = b.build()
    off = m.arrays["pod_req_off"]
    K = len(m.key_names)
    mask = np.zeros(K, np.uint64)
    flags = np.zeros(K, np.uint8)
    gt = np.zeros(K, np.int64)
    lt = np.zeros(K, np.int64)
    rc = ob.lib().oracle_req_algebra(C.byref(m.c), int(off[pa]), int(off[pa + 1]), int(off[pb]), int(off[pb + 1]),
                                     1 if allow else 0, ob._p(mask, C.c_uint64), ob._p(flags, C.c_uint8),
                                     ob._p(gt, C.c_int64), ob._p(lt, C.c_int64))
    return bool(rc & 1), bool(rc & 2), m, mask, flags, gt, lt


KEY = "example.com/k"


def test_requirement_algebra_cases():
    # In ∩ In
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_IN, ["a", "b"])], [(KEY, M.OP_IN, ["b", "c"])])
    assert c and i
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_IN, ["a"])], [(KEY, M.OP_IN, ["b"])])
    assert not c and not i
    # NotIn vs In (requirement.go:207-223)
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_NOTIN, ["a"])], [(KEY, M.OP_IN, ["a"])])
    assert not i
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_NOTIN, ["a"])], [(KEY, M.OP_IN, ["a", "b"])])
    assert i
    # both complement always intersect unless the bounds cross (requirement.go:198-206)
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_GT, 5)], [(KEY, M.OP_LT, 3)])
    assert not i
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_GT, 5)], [(KEY, M.OP_LT, 7)])
    assert i
    # Gt against a concrete set: values are parsed, non-integers never match (requirement.go:311-327)
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_GT, 5)], [(KEY, M.OP_IN, ["4", "x"])])
    assert not i
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_GT, 5)], [(KEY, M.OP_IN, ["6"])])
    assert i
    # DoesNotExist vs NotIn is forgiven by Intersects (requirements.go:253-258)
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_DOESNOTEXIST, None)], [(KEY, M.OP_NOTIN, ["a"])])
    assert i
    c, i, *_ = _algebra(_tiny_builder(), [(KEY, M.OP_DOESNOTEXIST, None)], [(KEY, M.OP_IN, ["a"])])
    assert not i
    # undefined custom key is denied unless the incoming operator is NotIn / DoesNotExist (requirements.go:178-188)
    c, i, *_ = _algebra(_tiny_builder(), [], [(KEY, M.OP_IN, ["a"])])
    assert not c and i
    c, i, *_ = _algebra(_tiny_builder(), [], [(KEY, M.OP_NOTIN, ["a"])])
    assert c
    c, i, *_ = _algebra(_tiny_builder(), [], [(KEY, M.OP_EXISTS, None)])
    assert not c
    # ... but a well-known key may be undefined
    c, i, *_ = _algebra(_tiny_builder(), [], [(M.KEY_ZONE, M.OP_IN, ["z1"])])
    assert c
    c, i, *_ = _algebra(_tiny_builder(), [], [(M.KEY_ZONE, M.OP_IN, ["z1"])], allow=False)
    assert not c


def test_intersection_prunes_and_drops_bounds():
    # NotIn {a} ∩ Gt 3: the non-integer exclusion is pruned (requirement.go:181-185), bounds stay
    c, i, m, mask, flags, gt, lt = _algebra(_tiny_builder(), [(KEY, M.OP_NOTIN, ["a"])], [(KEY, M.OP_GT, 3)])
    k = m.key_names.index(KEY)
    assert flags[k] == (1 | 2 | 4) and mask[k] == 0 and gt[k] == 3
    # In {2,5,x} ∩ Gt 3 -> In {5}, bounds dropped for concrete sets (requirement.go:186-189)
    c, i, m, mask, flags, gt, lt = _algebra(_tiny_builder(), [(KEY, M.OP_IN, ["2", "5", "x"])], [(KEY, M.OP_GT, 3)])
    k = m.key_names.index(KEY)
    vals = m.key_value_bits()[k]
    assert flags[k] == 1
    assert [m.val_names[vals[b]] for b in range(64) if (int(mask[k]) >> b) & 1] == ["5"]
    # Gt 5 ∩ Lt 3 -> DoesNotExist (requirement.go:166-168)
    c, i, m, mask, flags, gt, lt = _algebra(_tiny_builder(), [(KEY, M.OP_GT, 5)], [(KEY, M.OP_LT, 3)])
    k = m.key_names.index(KEY)
    assert flags[k] == 1 and mask[k] == 0


def test_tolerations():
    b = _tiny_builder()
    NS, NE = M.EFFECT_NOSCHEDULE, M.EFFECT_NOEXECUTE
    b.add_pool("t", [], taints=[("k1", "v1", NS), ("k2", "v2", NE)])
    cases = [
        ([("k1", M.TOLOP_EQUAL, "v1", NS), ("k2", M.TOLOP_EXISTS, "", M.EFFECT_NONE)], True),
        ([("k1", M.TOLOP_EQUAL, "v1", NS)], False),
        ([("", M.TOLOP_EXISTS, "", M.EFFECT_NONE)], True),                 # empty key + Exists: everything
        ([("k1", M.TOLOP_EMPTY, "v1", M.EFFECT_NONE), ("k2", M.TOLOP_EMPTY, "v2", NE)], True),  # "" means Equal
        ([("k1", M.TOLOP_EQUAL, "other", NS), ("k2", M.TOLOP_EXISTS, "", NE)], False),
        ([("k1", M.TOLOP_OTHER, "v1", NS), ("k2", M.TOLOP_EXISTS, "", NE)], False),
        ([("k1", M.TOLOP_EXISTS, "", NE), ("k2", M.TOLOP_EXISTS, "", NE)], False),   # effect mismatch
    ]
    pods = [b.add_pod([], [0, 0, 1], tolerations=t) for t, _ in cases]
    m = b.build()
    t0, t1 = int(m.arrays["pool_taint_off"][1]), int(m.arrays["pool_taint_off"][2])
    for p, (_, want) in zip(pods, cases):
        o0, o1 = int(m.arrays["pod_tol_off"][p]), int(m.arrays["pod_tol_off"][p + 1])
        assert bool(ob.lib().oracle_tolerates(C.byref(m.c), t0, t1, o0, o1)) == want


def test_queue_order_and_unschedulable_retry():
    m = synth.config("c3", 600, 60)
    o = ob.solve(m)
    req = m.arrays["pod_requests"]
    q = o.queue_order
    cpu, mem = req[q, 0], req[q, 1]
    ts, hi, lo = m.arrays["pod_ts"][q], m.arrays["pod_uid_hi"][q], m.arrays["pod_uid_lo"][q]
    for i in range(len(q) - 1):  # cpu desc, memory desc, creation asc, uid asc (queue.go:72-108)
        a = (-int(cpu[i]), -int(mem[i]), int(ts[i]), int(hi[i]), int(lo[i]))
        b = (-int(cpu[i + 1]), -int(mem[i + 1]), int(ts[i + 1]), int(hi[i + 1]), int(lo[i + 1]))
        assert a < b
    # every pod is either placed or left in PodErrors with an attempt code for some template
    unsched = o.pod_place_kind == -1
    assert unsched.sum() == o.n_unscheduled
    assert (o.pod_fit_code[unsched] != 0).any(axis=1).all()
    assert (o.pod_fit_code[~unsched] == 0).all()
    # pods of a claim: counts agree, requests are the sum of the pods' requests (daemon overhead is zero here)
    for c in range(o.n_claims):
        pods = np.nonzero((o.pod_place_kind == 1) & (o.pod_place_index == c))[0]
        assert len(pods) == o.claim_npods[c]
        assert (req[pods].sum(axis=0) == o.claim_requests[c]).all()
    assert o.counters["price_ties"] == 0


# ---- config 1: vectors taken from the reference's own tests -------------------------------------------------
GiB30 = 30 * 1024 * 1024 * 1024


def _create(name, types, storage, prices=None):
    from b200sched import capi
    arr = (C.c_char_p * max(len(types), 1))(*[t.encode() for t in types])
    pr = (C.c_double * max(len(types), 1))(*prices) if prices is not None else None
    ch, disk, gpu = C.c_int32(-1), C.c_int32(0), C.c_int32(0)
    rc = ob.lib().oracle_create_select(name.encode(), len(types), arr, pr, storage, C.byref(ch), C.byref(disk), C.byref(gpu))
    got = (rc, ch.value, disk.value, gpu.value)
    assert got == capi.create_select(name, types, storage, prices), "product and oracle disagree"
    return got


def test_create_select_reference_vectors():
    # pkg/providers/instance/instance_test.go:40-83 TestNewAgentPoolObject: 30Gi storage -> OSDiskSizeGB 30
    assert _create("agentpool0", ["Standard_NC6s_v3"], GiB30) == (0, 0, 30, 1)
    # ... "NodeClaim with no Storage requirement" -> error (instance.go:348-350)
    assert _create("agentpool0", ["Standard_NC6s_v3"], 0)[0] == -22
    # instance_test.go:537-668 TestCreateSuccess: instance.Type == Requirements[0].Values[0]
    assert _create("agentpool0", ["Standard_NC6s_v3"], GiB30)[1] == 0
    # instance_test.go:814-821: "invalid-name" fails ^[a-z][a-z0-9]{0,11}$ (instance.go:50,81)
    assert _create("invalid-name", ["Standard_D4s_v4"], GiB30)[0] == -20
    assert _create("a" * 13, ["Standard_D4s_v4"], GiB30)[0] == -20
    assert _create("a" * 12, ["Standard_D4s_v4"], GiB30)[0] == 0
    assert _create("9abc", ["Standard_D4s_v4"], GiB30)[0] == -20
    # instance_test.go:805-813: no instance-type requirement -> error (instance.go:91-93)
    assert _create("agentpool0", [], GiB30)[0] == -21
    # Standard_N* is labelled gpu, everything else cpu (instance.go:335-339)
    assert _create("agentpool0", ["Standard_D4s_v4"], GiB30) == (0, 0, 30, 0)


def test_create_select_config1_cheapest_of_ten():
    types = synth.REAL_GPU_SKUS[:10]
    prices = [3.0, 2.5, 9.0, 1.75, 4.0, 1.75, 6.0, 8.0, 2.0, 5.0]
    rc, chosen, disk, gpu = _create("kaito0", types, 120 << 30, prices)
    assert (rc, chosen, disk, gpu) == (0, 3, 120, 1)   # first of the two cheapest wins
    assert _create("kaito0", types, 120 << 30)[1] == 0  # without prices: the reference's values[0]

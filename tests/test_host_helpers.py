"""Rows a3 / a11 of the scope table: the per-pod request vector (resources.RequestsForPods) and the per-instance-type
allocatable vector (InstanceType.Allocatable) — host helpers of the library, checked against hand-worked cases that
follow the Go line by line and against the map-shaped oracle restatement on random inputs."""
import ctypes as C

import numpy as np

import oracle_binding as ob
import pytest

from b200sched import capi, synth

R, PODS, MEM = 4, 3, 1  # cpu, memory, hugepages-2Mi, pods


def _arr(a, dt):
    return np.ascontiguousarray(np.array(a, dtype=dt))


def _call(fn, containers, inits, overhead):
    def pack(items):
        n = len(items)
        q = np.zeros((max(n, 1), R), np.int64); qh = np.zeros((max(n, 1), R), np.uint8)
        l = np.zeros((max(n, 1), R), np.int64); lh = np.zeros((max(n, 1), R), np.uint8)
        always = np.zeros(max(n, 1), np.uint8)
        for i, it in enumerate(items):
            for r, v in it.get("req", {}).items():
                q[i, r], qh[i, r] = v, 1
            for r, v in it.get("lim", {}).items():
                l[i, r], lh[i, r] = v, 1
            always[i] = 1 if it.get("always") else 0
        return n, q, qh, l, lh, always
    nc, q, qh, l, lh, _ = pack(containers)
    ni, iq, iqh, il, ilh, always = pack(inits)
    ov = _arr(overhead if overhead is not None else [0] * R, np.int64)
    out = np.zeros(R, np.int64)
    P = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))  # noqa: E731
    rc = fn(R, PODS, nc, P(q, C.c_int64), P(qh, C.c_uint8), P(l, C.c_int64), P(lh, C.c_uint8), ni, P(iq, C.c_int64),
            P(iqh, C.c_uint8), P(il, C.c_int64), P(ilh, C.c_uint8), P(always, C.c_uint8),
            P(ov, C.c_int64) if overhead is not None else None, P(out, C.c_int64))
    assert rc == 0
    return out.tolist()


def pod_requests(containers, inits=(), overhead=None):
    a = _call(capi.lib().b200s_pod_requests, list(containers), list(inits), overhead)
    b = _call(ob.lib().oracle_pod_requests, list(containers), list(inits), overhead)
    assert a == b, (a, b)
    return a


def test_pod_requests_hand_worked():
    # two containers: requests add; a limit stands in only where the request is missing (resources.go:204-217)
    assert pod_requests([{"req": {0: 500, 1: 100}}, {"req": {0: 250}, "lim": {0: 999, 1: 64}}]) == [750, 164, 0, 1]
    # a plain init container only matters when it exceeds the running sum (resources.go:129-135)
    assert pod_requests([{"req": {0: 500}}], [{"req": {0: 300}}]) == [500, 0, 0, 1]
    assert pod_requests([{"req": {0: 500}}], [{"req": {0: 2000, 1: 7}}]) == [2000, 7, 0, 1]
    # a sidecar (restartPolicy Always) adds to the sum and to every later init container's need (:125-132)
    assert pod_requests([{"req": {0: 500}}], [{"req": {0: 100}, "always": True}, {"req": {0: 450}}]) == [600, 0, 0, 1]
    assert pod_requests([{"req": {0: 100}}], [{"req": {0: 100}, "always": True}, {"req": {0: 450}}]) == [550, 0, 0, 1]
    # pod overhead is added after the max (:137-139); pods is set, not added (:36)
    assert pod_requests([{"req": {0: 100, 3: 5}}], overhead=[10, 20, 0, 0]) == [110, 20, 0, 1]


def test_pod_requests_random_against_oracle():
    rng = np.random.default_rng(5)
    for _ in range(300):
        mk = lambda: {"req": {int(r): int(rng.integers(0, 1000)) for r in rng.choice(3, rng.integers(0, 4), replace=False)},  # noqa: E731
                      "lim": {int(r): int(rng.integers(0, 1000)) for r in rng.choice(3, rng.integers(0, 4), replace=False)},
                      "always": bool(rng.random() < 0.3)}
        pod_requests([mk() for _ in range(rng.integers(0, 4))], [mk() for _ in range(rng.integers(0, 4))],
                     [int(x) for x in rng.integers(0, 50, R)] if rng.random() < 0.5 else None)


def allocatable(cap, cap_has, kube, system, evict, huge):
    outs = []
    for fn in (capi.lib().b200s_allocatable, ob.lib().oracle_allocatable):
        out = np.zeros(R, np.int64)
        P = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))  # noqa: E731
        rc = fn(R, MEM, P(_arr(cap, np.int64), C.c_int64), P(_arr(cap_has, np.uint8), C.c_uint8), P(_arr(kube, np.int64), C.c_int64),
                P(_arr(system, np.int64), C.c_int64), P(_arr(evict, np.int64), C.c_int64), P(_arr(huge, np.uint8), C.c_uint8),
                P(out, C.c_int64))
        assert rc == 0
        outs.append(out.tolist())
    assert outs[0] == outs[1], outs
    return outs[0]


def test_allocatable_hand_worked():
    GiB = 1 << 30
    # capacity - (kube + system + eviction) per resource (types.go:182)
    assert allocatable([4000, 16 * GiB, 0, 110], [1, 1, 0, 1], [100, GiB, 0, 0], [50, 0, 0, 0], [0, 100 << 20, 0, 0],
                       [0, 0, 1, 0]) == [3850, 15 * GiB - (100 << 20), 0, 110]
    # hugepages capacity comes off memory (types.go:189-198), clamped at zero
    assert allocatable([4000, 16 * GiB, 2 * GiB, 110], [1, 1, 1, 1], [0, GiB, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0],
                       [0, 0, 1, 0]) == [4000, 13 * GiB, 2 * GiB, 110]
    assert allocatable([4000, GiB, 2 * GiB, 110], [1, 1, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0],
                       [0, 0, 1, 0])[1] == 0


def test_allocatable_random_against_oracle():
    rng = np.random.default_rng(6)
    for _ in range(300):
        allocatable(rng.integers(0, 1 << 20, R), rng.integers(0, 2, R), rng.integers(0, 1 << 10, R), rng.integers(0, 1 << 10, R),
                    rng.integers(0, 1 << 10, R), [0, 0, 1, 0])


def _team_model(team_key):
    """one catalog, one pool, one pod whose only requirement is NotIn on `team_key` -- a key nothing else uses (an
    undefined key is forgiven for NotIn, requirements.go:183, so the pod still schedules)"""
    from b200sched import model as M
    b = synth.new_builder()
    synth.add_catalog(b, 30)
    for k in ("example.com/team-a", "example.com/team-b"):
        b.key(k)
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    b.add_pod([(team_key, M.OP_NOTIN, ["x"])], [500, 1 << 30, 0, 0, 1])
    return b.build()


def test_layout_hash_tells_encodes_with_different_key_slots_apart():
    """the device key slots come from the keys one model uses, pods included.  Two batches over the same
    catalog whose pods select on different keys encode to the same K / NB but different slots; the stamp in the catalog,
    pods and nodes SoA differs, and b200s_solve answers B200S_ESTALE instead of evaluating against stale tables
    (tests/test_gpu.py::test_stale_catalog_is_refused)."""
    ea, eb = capi.Encoded(_team_model("example.com/team-a")), capi.Encoded(_team_model("example.com/team-b"))
    ea2 = capi.Encoded(_team_model("example.com/team-a"))
    assert ea.layout_hash != 0 and ea.layout_hash == ea2.layout_hash
    assert ea.pods.contents.layout_hash == ea.layout_hash == ea.nodes.contents.layout_hash
    assert ea.key_model != eb.key_model and ea.K == eb.K and ea.NB == eb.NB  # same counts, different slots
    assert ea.layout_hash != eb.layout_hash


def test_candidate_price_errors_surface_where_the_reference_would_hit_them():
    """getCandidatePrices (consolidation.go:310-328): a reserved node without an offering makes the total 0.0, any other
    unpriced candidate is an error -- but only for a probe whose simulation reaches the price comparison, and only when
    the binary search visits that probe."""
    from b200sched import consolidation as cons
    facts = [(0, 0.5, False, False), (1, -1.0, False, True), (2, 0.25, False, False), (3, -1.0, False, False)]
    assert cons.prefix_price(facts, 0) == (0.5, False)
    assert cons.prefix_price(facts, 2) == (0.0, False)      # the reserved candidate ends the walk: free
    assert cons.prefix_price(facts[2:], 1) == (0.0, True)   # unpriced and not reserved
    visited = []

    def valid_of(mid):
        visited.append(mid)
        if mid == 5:
            raise cons.CandidatePriceError("x")
        return mid < 3
    assert cons.replay_search(1, 4, valid_of) == 2 and 5 not in visited  # the search never looks at prefix 5
    with pytest.raises(cons.CandidatePriceError):
        cons.replay_search(1, 9, valid_of)                   # (1 + 9) // 2 == 5 is the first probe


def test_batch_result_buffers_alias_their_arrays():
    """capi.BatchResultBuffers lays the b200s_result_t structs of a batch out as one structured array whose pointer fields
    aim into one allocation per field, problem b's rows starting at its first pod: what the library writes through a struct
    is what raws() hands back."""
    import ctypes as C
    pods = [3, 0, 5]
    T, R, W, K, NB = 2, 4, 3, 2, 1
    b = capi.BatchResultBuffers(pods, T, R, W, K, NB)
    ptrs = b.out_pointers()
    for i, P in enumerate(pods):
        res = ptrs[i].contents
        assert res.cap_claims == max(P, 1)
        res.n_claims, res.n_unscheduled, res.fit_evals = min(P, 2), P - min(P, 2), 1000 + i
        for p in range(P):
            res.pod_place_kind[p] = 10 * i + p
            for t in range(T):
                res.pod_fit_code[p * T + t] = i + p + t
        for c in range(min(P, 2)):
            res.claim_npods[c] = 7 * i + c
            for w in range(W):
                res.claim_mask[c * W + w] = (i << 8) | (c << 4) | w
            for r in range(R):
                res.claim_requests[c * R + r] = -(i * 100 + c * 10 + r)
    raws = b.raws()
    for i, (P, raw) in enumerate(zip(pods, raws)):
        N = min(P, 2)
        assert (raw.n_claims, raw.n_unscheduled, raw.fit_evals) == (N, P - N, 1000 + i)
        assert raw.pod_place_kind.tolist() == [10 * i + p for p in range(P)]
        assert raw.pod_fit_code.shape == (P, T) and raw.pod_fit_code.tolist() == [[i + p + t for t in range(T)] for p in range(P)]
        assert raw.claim_npods.tolist() == [7 * i + c for c in range(N)]
        assert raw.claim_mask.tolist() == [[(i << 8) | (c << 4) | w for w in range(W)] for c in range(N)]
        assert raw.claim_requests.tolist() == [[-(i * 100 + c * 10 + r) for r in range(R)] for c in range(N)]
        assert raw.claim_req_mask.shape == (N, K) and raw.claim_req_gt.shape == (N, NB)

"""CPU checks of everything that does not need a GPU: the encoder and the bit-row tables, the shared core
arithmetic (tests/native/emul.cpp runs b200_core.cuh on the host), the Go-sort port used by the pack kernel,
encoder rejections and the exported C-ABI."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import fuzz_models
import oracle_binding as ob
import parity
from b200sched import capi, synth
from b200sched import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compare(model, what):
    enc = capi.Encoded(model)
    o_tmpl, o_mask, o_code, o_evals, _ = ob.fit_matrix(model)
    ef = parity.emul_fit_matrix(enc)
    cf = parity.canon_fit(enc, ef)
    assert np.array_equal(cf["tmpl"], o_tmpl), what
    assert np.array_equal(cf["mask"], o_mask), what
    assert np.array_equal(cf["code"], o_code), what
    assert ef.evals == o_evals, what
    o = ob.solve(model)
    parity.assert_same(parity.canon_device(enc, parity.emul_solve(enc)), parity.canon_oracle(model, o), what)
    return o


@pytest.mark.parametrize("kind,P,S", [("c2", 1500, 100), ("c3", 2500, 200), ("c3", 1200, 700), ("c2", 300, 1300)])
def test_emulation_matches_oracle_on_synthetic_configs(kind, P, S):
    o = _compare(synth.config(kind, P, S), f"{kind} {P}x{S}")
    assert o.counters["price_ties"] == 0


def test_emulation_matches_oracle_on_fuzzed_models():
    ran = 0
    rejected = 0
    for seed in range(400):
        m = fuzz_models.random_model(seed)
        try:
            _compare(m, f"fuzz seed {seed}")
            ran += 1
        except capi.NotRepresentable:
            rejected += 1
    assert ran >= 350, (ran, rejected)


@pytest.mark.parametrize("kind,P,S", [("c2", 1500, 100), ("c2", 4000, 500), ("c2", 300, 1300), ("c2", 3000, 64)])
def test_register_pack_twin_matches_oracle_on_config2_shapes(kind, P, S):
    """tests/native/emul.cpp emul_solve_reg, the serial twin of pack_reg_kernel: claims know their positions, a step's
    change of the order is one shift word applied a step late, the minimum is taken over priorities that do not have it
    yet, sort.Slice runs only when it could take the real pdqsort path, and the last change is never applied.  The twin
    also checks at every step that the stale priorities pick the claim the true positions pick."""
    m = synth.config(kind, P, S)
    enc = capi.Encoded(m)
    o = ob.solve(m)
    raw, stats = parity.emul_solve_reg(enc)
    parity.assert_same(parity.canon_device(enc, raw), parity.canon_oracle(m, o), f"reg twin {kind} {P}x{S}")
    assert stats[0] == P
    if o.n_claims > 60:  # most steps need no sort at all once there are 50 claims
        assert 0 < stats[3] < P // 3 and stats[1] <= stats[3], stats
    # a count table of three entries: every claim with two or more pods takes the "beyond the table" path (real sort)
    raw, stats = parity.emul_solve_reg(enc, kcap=3)
    parity.assert_same(parity.canon_device(enc, raw), parity.canon_oracle(m, o), f"reg twin {kind} {P}x{S} kcap 3")


def test_register_pack_twin_reproduces_config2_digest_and_the_kernels_counters():
    """BASELINE config[1] at full size on the serial twin: the golden digest, and the two counters pack_reg_kernel reports
    for the same workload (bench.py step_breakdown: 318 of the 10 000 steps meet the order warp, 284 run the full
    pdqsort) -- the kernel and its twin take the same path, not only reach the same result."""
    enc = capi.Encoded(synth.c2())
    raw, stats = parity.emul_solve_reg(enc)
    assert parity.digest(parity.canon_device(enc, raw))[:16] == "1423b60e3481b574"
    assert (int(stats[0]), int(stats[1]), int(stats[3])) == (10_000, 284, 318)


def test_register_pack_twin_matches_oracle_on_fuzzed_keyless_models():
    ran = 0
    for seed in range(300):
        m = fuzz_models.random_model(seed, with_nodes=False, max_pods=160, keyless=True)
        try:
            enc = capi.Encoded(m)
        except capi.NotRepresentable:
            continue
        o = parity.canon_oracle(m, ob.solve(m))
        for kcap in (0, 3, 5):  # 0: the kernel's 2 048; small tables send claims down the "beyond the table" path (3 is the least that still counts single-pod claims)
            raw, _ = parity.emul_solve_reg(enc, kcap=kcap)
            parity.assert_same(parity.canon_device(enc, raw), o, f"reg twin fuzz {seed} kcap {kcap}")
        ran += 1
    assert ran >= 250, ran


def test_nodepool_limits_match_oracle_on_fuzzed_models():
    """scheduler.go:587-675, :831-867: filterByRemainingResources before every new NodeClaim, subtractMax after it, the
    "exceed limits" attempt code — the device algorithm's host twin (bit rows over Capacity values) against the oracle"""
    ran = limited = 0
    for seed in range(300):
        m = fuzz_models.random_model(seed, with_nodes=seed % 3 == 0, max_pods=60, limits=True)
        try:
            enc = capi.Encoded(m)
        except capi.NotRepresentable:
            continue
        o = ob.solve(m)
        parity.assert_same(parity.canon_device(enc, parity.emul_solve(enc)), parity.canon_oracle(m, o), f"limits fuzz {seed}")
        ran += 1
        limited += int((o.pod_fit_code == capi.FIT_LIMITS).any())
    assert ran >= 250 and limited >= 20, (ran, limited)


def test_go_sort_port_of_the_pack_kernel_matches_the_oracle_port():
    rng = np.random.default_rng(7)
    for n in [1, 2, 12, 13, 30, 49, 50, 64, 200, 1000, 3000]:
        for hi in (2, 4, 50):
            cnt = rng.integers(1, hi + 1, n).astype(np.int32)
            assert parity.emul_go_sort(cnt).tolist() == ob.go_sort_perm(cnt.astype(np.int64)).tolist()
        # the shapes the pack produces: sorted, then one increment or one append
        cnt = np.sort(rng.integers(1, 6, n)).astype(np.int32)
        for _ in range(20):
            c = cnt.copy()
            c[rng.integers(n)] += 1
            assert parity.emul_go_sort(c).tolist() == ob.go_sort_perm(c.astype(np.int64)).tolist()
        c = np.concatenate([cnt, [1]]).astype(np.int32)
        assert parity.emul_go_sort(c).tolist() == ob.go_sort_perm(c.astype(np.int64)).tolist()


def test_single_displacement_fast_path_claim():
    """pack_kernel re-inserts the one displaced NodeClaim next to its equals when n >= 50 and choosePivot saw no
    inversion.  Whenever that condition holds, Go's sort must agree with a stable sort."""
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(400):
        n = int(rng.integers(50, 400))
        cnt = np.sort(rng.integers(1, 8, n)).astype(np.int64)
        c = cnt.copy()
        if rng.random() < 0.5:
            c[rng.integers(n)] += 1
        else:
            c = np.concatenate([c, [1]])
        l = len(c)
        # choosePivot's samples (zsortfunc.go choosePivot_func, l >= 50: three adjacent triples)
        idx = [l // 4 * 1, l // 4 * 2, l // 4 * 3]
        tri = [c[i - 1:i + 2] for i in idx]
        med = [sorted(t)[1] for t in tri]
        no_inv = all((np.diff(t) >= 0).all() for t in tri) and med[0] <= med[1] <= med[2]
        if no_inv:
            checked += 1
            assert ob.go_sort_perm(c).tolist() == np.argsort(c, kind="stable").tolist()
    assert checked > 100


def _reject(model):
    with pytest.raises(capi.NotRepresentable):
        capi.Encoded(model)


def test_encoder_rejects_what_it_cannot_represent_exactly():
    def base():
        b = synth.new_builder()
        synth.add_catalog(b, 20)
        return b
    # minValues
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"]), (synth.KEY_FAMILY, M.OP_IN, ["D", "E"])], min_values=[-1, 2])
    b.add_pod([], [100, 1 << 20, 0, 0, 1])
    _reject(b.build())
    # PreferNoSchedule taint (preference relaxation path)
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])], taints=[("k", "v", M.EFFECT_PREFERNOSCHEDULE)])
    b.add_pod([], [100, 1 << 20, 0, 0, 1])
    _reject(b.build())
    # NodePool limits are representable since round 2 (bit rows over Capacity values)
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])], limits={"cpu": 100000})
    b.add_pod([], [100, 1 << 20, 0, 0, 1])
    e = capi.Encoded(b.build())
    assert e.templates.contents.has_limit[0][0] == 1 and e.templates.contents.limit[0][0] == 100000
    # a pod selecting on the instance type
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    b.add_pod([(M.KEY_INSTANCE_TYPE, M.OP_IN, ["Standard_NC6"])], [100, 1 << 20, 0, 0, 1])
    _reject(b.build())
    # more than 64 values on one key
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"]), ("example.com/wide", M.OP_IN, [f"v{i}" for i in range(65)])])
    b.add_pod([], [100, 1 << 20, 0, 0, 1])
    _reject(b.build())
    # negative request
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    b.add_pod([], [-1, 1 << 20, 0, 0, 1])
    _reject(b.build())
    # the unrejected twin encodes
    b = base()
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    b.add_pod([], [100, 1 << 20, 0, 0, 1])
    capi.Encoded(b.build())


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200sched.h")).read()
    names = set(re.findall(r"\b(b200s_[a-z0-9_]+)\s*\(", hdr))
    names = {n for n in names if not n.endswith("_t")}
    assert len(names) >= 30
    L = capi.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert L.b200s_version() == 100


def test_no_device_means_no_context(monkeypatch):
    """There is no CPU backend: without a GPU b200s_init must fail rather than fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.B200SError) as ei:
        capi.Context(0)
    assert ei.value.code == capi.ENODEVICE


def test_golden_digests():
    """Oracle results on fixed inputs, committed with the script that made them (tests/golden/make_golden.py).
    They pin the oracle against accidental change; they are NOT reference outputs (parity unpinned)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_digests.json")))
    for name, g in gold["cases"].items():
        kind, P, S = g["kind"], g["pods"], g["skus"]
        m = synth.config(kind, P, S)
        o = ob.solve(m)
        assert parity.digest(parity.canon_oracle(m, o)) == g["digest"], name
        assert [o.n_claims, o.n_unscheduled, o.counters["fit_evals"]] == [g["n_claims"], g["n_unscheduled"], g["fit_evals"]]


def test_config4_generator_and_simulations_on_the_host_emulation():
    from b200sched import model as M
    model, node_of_pod, util = synth.c4(1500, 120, solve=ob.solve)
    assert (node_of_pod >= 0).sum() > 1400
    for i, sim in synth.c4_simulations(model, node_of_pod, util, 6):
        o = ob.solve(sim)
        enc = capi.Encoded(sim)
        parity.assert_same(parity.canon_device(enc, parity.emul_solve(enc)), parity.canon_oracle(sim, o), f"node {i}")
        assert sim.n_pods == int((node_of_pod == i).sum()) and int(sim.c.n_nodes) == int(model.c.n_nodes) - 1


def test_rank_bucket_table_matches_binary_search():
    """The pack kernel finds "how many distinct allocatable values are below x" through a log-scale bucket table
    (b200_core.cuh rank_bucket / fit_rank_bkt); it must agree with the plain lower bound for every x."""
    rng = np.random.default_rng(7)
    for trial in range(40):
        kind = trial % 4
        if kind == 0:    # cpu-like millicores, geometric
            dv = np.unique((rng.choice([2, 4, 8, 16, 32, 48, 64, 96, 128], 9) * 1000 - rng.integers(0, 400, 9)))
        elif kind == 1:  # memory-like bytes, clustered
            dv = np.unique(rng.integers(1, 2048, 60) * (1 << 30) - rng.integers(0, 1 << 28, 60))
        elif kind == 2:  # tiny values incl. 0..8 (gpu counts, pods)
            dv = np.unique(rng.integers(0, 9, 6))
        else:            # dense random up to 2^62
            dv = np.unique(rng.integers(0, 1 << 62, 300, dtype=np.int64) >> int(rng.integers(0, 50)))
        xs = np.concatenate([dv, dv + 1, dv - 1, rng.integers(0, int(dv.max()) + 5, 400), [0, 1, 2, 3, 4, 5, 7, 8, -5],
                             [np.iinfo(np.int64).max]]).astype(np.int64)
        a, b = parity.emul_fit_rank_both(dv, xs)
        assert np.array_equal(a, b), (trial, dv[:8], xs[a != b][:8])
        assert np.array_equal(a, np.searchsorted(dv, xs, side="left"))

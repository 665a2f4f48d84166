"""The C++ oracle against a second, independent restatement of Scheduler.Solve (tests/solve_ref.py: Python, written from
the Go text, reference-shaped).  Neither is the reference — there is no Go toolchain here and the vendored tree carries no
scheduler tests — but two restatements by different routes that agree on every placement, on the unstable sort's tie
order and on every NodeClaim's instance types over hundreds of adversarial models leave little room for a shared
misreading."""
import numpy as np
import pytest

import fuzz_models
import oracle_binding as ob
import solve_ref
from b200sched import synth


def _same(model, what):
    o = ob.solve(model)
    r = solve_ref.solve(model)
    assert (o.n_claims, o.n_unscheduled) == (r["n_claims"], r["n_unscheduled"]), what
    for f in ("pod_place_kind", "pod_place_index", "pod_seq", "claim_npods", "claim_requests", "claim_mask", "claim_sku"):
        assert np.array_equal(np.asarray(getattr(o, f)), r[f]), (what, f)
    assert np.array_equal(o.template_pool[o.claim_template] if o.n_claims else o.claim_template, r["claim_pool"]), what
    assert np.array_equal(o.node_npods, r["node_npods"]), what
    return o


@pytest.mark.parametrize("block", range(10))
def test_python_restatement_agrees_with_the_oracle_on_fuzzed_models(block):
    placed = unsched = on_nodes = 0
    for seed in range(block * 100, block * 100 + 100):
        m = fuzz_models.random_model(7000 + seed, with_nodes=seed % 3 != 0, max_pods=60, keyless=seed % 5 == 0,
                                     limits=seed % 2 == 0)
        o = _same(m, f"fuzz seed {seed}")
        placed += int((o.pod_place_kind >= 0).sum())
        unsched += o.n_unscheduled
        on_nodes += int((o.pod_place_kind == 0).sum())
    assert placed > 100 and unsched > 0 and (on_nodes > 0 or block == 0)


@pytest.mark.parametrize("kind,P,S", [("c2", 1500, 8), ("c3", 900, 16), ("c2", 1200, 100), ("c3", 1500, 150)])
def test_python_restatement_agrees_on_the_synthetic_configs(kind, P, S):
    """more than 50 NodeClaims: the re-sort before every pod takes pdqsort's ninther / partialInsertionSort / partition
    paths, and both sides must permute the ties the same way for the later placements to agree"""
    o = _same(synth.config(kind, P, S), f"{kind} {P}x{S}")
    assert o.n_claims >= 50


def test_python_restatement_agrees_on_consolidation_simulations():
    """existing nodes by the hundred: a candidate's pods against every other node of a loosely packed cluster"""
    model, node_of_pod, util = synth.c4(1500, 120, solve=ob.solve)
    on_nodes = 0
    for i, sim in synth.c4_simulations(model, node_of_pod, util, 8, spread=True):
        o = _same(sim, f"simulation of node {i}")
        on_nodes += int((o.pod_place_kind == 0).sum())
    assert on_nodes > 0


@pytest.mark.parametrize("block", range(3))
def test_price_ties_fall_where_pdqsort_leaves_them(block):
    """Equal prices make OrderByPrice's result — hence the cheapest SKU of a NodeClaim and ToNodeClaim's instance-type
    list — depend on Go's unstable sort.  With the prices of fuzzed catalogs squeezed onto a handful of values, the
    oracle's list and the Python restatement's (a third port of pdqsort) must still be the same list."""
    ties = 0
    for seed in range(block * 40, block * 40 + 40):
        m = fuzz_models.random_model(9000 + seed, with_nodes=False, max_pods=50, keyless=seed % 2 == 0)
        pr = m.arrays["off_price"]
        pr[:] = np.round(pr * 1.5, 0) / 1.5 + 0.25   # a few distinct prices per catalog (in place: the C model aliases it)
        o, order, count, tie = ob.solve_ordered(m, 600)
        r = solve_ref.solve(m)
        assert o.n_claims == r["n_claims"] and np.array_equal(o.pod_place_index, r["pod_place_index"]), seed
        assert np.array_equal(o.claim_sku, r["claim_sku"]), seed
        for i in range(o.n_claims):
            assert order[i, :count[i]].tolist() == r["claim_order"][i][:600], (seed, i)
        ties += int(tie.sum())
    assert ties >= 10

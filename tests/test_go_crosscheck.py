"""The off-box Go cross-check's two Python ends (tools/dump_model.py, tools/compare_go_result.py): the dump is
deterministic and complete enough to rebuild the objects, and the comparison accepts the oracle's own view and
notices a changed field.  baseline/go/main.go itself needs a Go toolchain and is not run here."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import compare_go_result as cg  # noqa: E402
import dump_model  # noqa: E402
import oracle_binding as ob  # noqa: E402
from b200sched import synth  # noqa: E402


def test_dump_is_deterministic_and_names_everything():
    gen = {"kind": "c3", "pods": 300, "skus": 80}
    a = dump_model.dump(synth.config("c3", 300, 80), gen)
    b = dump_model.dump(synth.config("c3", 300, 80), gen)
    assert json.dumps(a, sort_keys=True) == json.dumps(b, sort_keys=True)
    assert len(a["instance_types"]) == 80 and len(a["pods"]) == 300 and len(a["node_pools"]) == 4
    it = a["instance_types"][0]
    assert set(it["capacity"]) == set(a["resources"]) and len(it["offerings"]) == 6
    assert it["capacity"]["cpu"].endswith("m") and not it["capacity"]["memory"].endswith("m")
    ops = {r["operator"] for p in a["pods"] for r in p["requirements"]} | {r["operator"] for np in a["node_pools"] for r in np["requirements"]}
    assert {"In", "NotIn", "Gt"} <= ops
    assert len({p["uid"] for p in a["pods"]}) == 300 and all(len(p["uid"]) == 36 for p in a["pods"])


def test_comparison_accepts_the_oracles_view_and_sees_a_change():
    m = synth.config("c3", 500, 100)
    want = cg.oracle_view(m, ob.solve(m))
    got = json.loads(json.dumps(want))
    got["solve_seconds"] = 0.0
    # what the Go side adds and the model does not know: the NodeClass label NewNodeClaimTemplate injects
    got["new_node_claims"][0]["requirements"]["karpenter.test.sh/testnodeclass"] = {"operator": "In", "values": ["default"]}
    assert cg.compare(m, want, got) == []
    for field, value in (("cheapest", "Standard_X"), ("pods", [0]), ("node_pool", "other")):
        bad = copy.deepcopy(got)
        bad["new_node_claims"][2][field] = value
        assert any(field in d for d in cg.compare(m, want, bad))
    bad = copy.deepcopy(got)
    bad["unschedulable"] = bad["unschedulable"] + [12345]
    assert cg.compare(m, want, bad)


def test_models_with_existing_nodes_round_trip_through_the_dump_and_the_comparison():
    """config[3]-style simulations and fuzzed models: the dump names every node (labels, remaining resources, taints,
    initialized), limits and daemon overheads; the oracle's view lists the pods each existing node took; the comparison
    accepts it and notices a pod moved to another node."""
    checked = on_nodes = 0
    for gen in ([{"kind": "c4sim", "pods": 600, "skus": 60, "sim": j} for j in range(3)] +
                [{"kind": "fuzz", "seed": 1000 + s} for s in range(40)]):
        m = dump_model.generate(gen)
        if dump_model.expressible(m) is not None:
            continue
        d = json.loads(json.dumps(dump_model.dump(m, gen)))
        assert len(d["nodes"]) == int(m.c.n_nodes) and all(set(n) == {"name", "labels", "available", "taints", "initialized"} for n in d["nodes"])
        assert all("limits" in p and "daemon_overhead" in p for p in d["node_pools"])
        want = cg.oracle_view(m, ob.solve(m))
        got = json.loads(json.dumps(want))
        assert cg.compare(m, want, got) == []
        checked += 1
        if want["existing_nodes"]:
            on_nodes += 1
            bad = copy.deepcopy(got)
            name = next(iter(bad["existing_nodes"]))
            bad["existing_nodes"]["some-other-node"] = bad["existing_nodes"].pop(name)
            assert cg.compare(m, want, bad)
    assert checked >= 15 and on_nodes >= 5, (checked, on_nodes)

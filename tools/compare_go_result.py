"""Hold the Go scheduler's result (baseline/go/main.go, run off-box) against the CPU oracle on the same model.

  python tools/compare_go_result.py model.json go_result.json

model.json is what tools/dump_model.py wrote; its generator arguments are used to rebuild the model here, the dump is
re-derived from it and must equal the file (so both sides provably saw the same input).  Compared per new NodeClaim in
Results.NewNodeClaims order: NodePool, pod set, instance-type set, cheapest type, accumulated requests, and the merged
requirements of the model's keys in the projection Requirement.NodeSelectorRequirement() makes
(K/scheduling/requirement.go:93-160); plus the set of unschedulable pods.  Exit status 0 = the oracle is pinned on
this input, and DESIGN.md section 4's "parity unpinned" can be lifted for it.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import dump_model  # noqa: E402
import oracle_binding as ob  # noqa: E402
from b200sched import capi, synth  # noqa: E402


def mask_bits(words):
    out = []
    for w, word in enumerate(int(x) for x in words):
        while word:
            b = (word & -word).bit_length() - 1
            out.append(w * 64 + b)
            word &= word - 1
    return out


def oracle_view(model, o):
    """the oracle's Solve result in the JSON shape of baseline/go/main.go"""
    bits = model.key_value_bits()
    claims = []
    for c in range(o.n_claims):
        reqs = {}
        for k, name in enumerate(model.key_names):
            f = int(o.claim_req_flags[c, k])
            if not f & capi.RF_DEFINED:
                continue
            vals = sorted(model.val_names[v] for i, v in enumerate(bits[k]) if (int(o.claim_req_mask[c, k]) >> i) & 1)
            if f & capi.RF_HASGT:
                reqs[name] = {"operator": "Gt", "values": [str(int(o.claim_req_gt[c, k]))]}
            elif f & capi.RF_HASLT:
                reqs[name] = {"operator": "Lt", "values": [str(int(o.claim_req_lt[c, k]))]}
            elif f & capi.RF_COMPLEMENT:
                reqs[name] = {"operator": "NotIn" if vals else "Exists", "values": vals}
            else:
                reqs[name] = {"operator": "In" if vals else "DoesNotExist", "values": vals}
        pods = sorted(int(p) for p in range(model.n_pods)
                      if o.pod_place_kind[p] == capi.PLACE_CLAIM and o.pod_place_index[p] == c)
        claims.append({
            "node_pool": model.pool_names[int(o.template_pool[int(o.claim_template[c])])], "pods": pods,
            "instance_types": sorted(model.sku_names[s] for s in mask_bits(o.claim_mask[c])),
            "cheapest": model.sku_names[int(o.claim_sku[c])] if o.claim_sku[c] >= 0 else "",
            "requests": {r: dump_model.quantity(r, o.claim_requests[c, i]) for i, r in enumerate(model.res_names)},
            "requirements": reqs})
    unsched = sorted(int(p) for p in range(model.n_pods) if o.pod_place_kind[p] == capi.PLACE_UNSCHEDULED)
    # Results.ExistingNodes: the pods each existing node took (rows of the oracle's result are sortExistingNodes order)
    on_nodes = {}
    for p in range(model.n_pods):
        if o.pod_place_kind[p] == capi.PLACE_NODE:
            on_nodes.setdefault(model.node_names[int(o.node_order[int(o.pod_place_index[p])])], []).append(int(p))
    return {"new_node_claims": claims, "unschedulable": unsched,
            "existing_nodes": {n: sorted(ps) for n, ps in sorted(on_nodes.items())}}


def compare(model, want, got):
    """list of differences between the oracle's view and the Go result"""
    diffs = []
    if want["unschedulable"] != got["unschedulable"]:
        diffs.append(f"unschedulable pods differ: oracle {len(want['unschedulable'])}, go {len(got['unschedulable'])}")
    if want.get("existing_nodes", {}) != {n: sorted(ps) for n, ps in (got.get("existing_nodes") or {}).items() if ps}:
        diffs.append("pods placed on existing nodes differ")
    if len(want["new_node_claims"]) != len(got["new_node_claims"]):
        diffs.append(f"NewNodeClaims: oracle {len(want['new_node_claims'])}, go {len(got['new_node_claims'])}")
    keys = set(model.key_names)
    for i, (w, g) in enumerate(zip(want["new_node_claims"], got["new_node_claims"])):
        for f in ("node_pool", "pods", "instance_types", "cheapest"):
            if w[f] != g[f]:
                diffs.append(f"claim {i}: {f} differs")
        for r, q in w["requests"].items():  # a resource the Go side never saw requested is absent there
            if g["requests"].get(r, dump_model.quantity(r, 0)) != q:
                diffs.append(f"claim {i}: requests[{r}] oracle {q}, go {g['requests'].get(r)}")
        g_reqs = {k: v for k, v in g["requirements"].items() if k in keys}  # the NodeClass label key is not in the model
        for k in sorted(set(w["requirements"]) | set(g_reqs)):
            a, b = w["requirements"].get(k), g_reqs.get(k)
            if a is None or b is None or a["operator"] != b["operator"] or a["values"] != sorted(b.get("values") or []):
                diffs.append(f"claim {i}: requirement {k}: oracle {a}, go {b}")
    return diffs


if __name__ == "__main__":
    dumped = json.load(open(sys.argv[1]))
    got = json.load(open(sys.argv[2]))
    gen = dumped["generator"]
    model = dump_model.generate(gen)
    again = json.loads(json.dumps(dump_model.dump(model, gen)))
    if again != dumped:
        raise SystemExit("the model regenerated here differs from the dump the Go harness read")
    o = ob.solve(model)
    diffs = compare(model, oracle_view(model, o), got)
    for d in diffs[:50]:
        print(d)
    print(f"{len(got['new_node_claims'])} NodeClaims, {len(got['unschedulable'])} unschedulable; go Solve "
          f"{got.get('solve_seconds', float('nan')):.3f} s; {'IDENTICAL' if not diffs else str(len(diffs)) + ' differences'}")
    sys.exit(1 if diffs else 0)

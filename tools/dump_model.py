"""Write a synthetic scheduling model as string-level JSON for the off-box Go cross-check (baseline/go/main.go).

  python tools/dump_model.py {c2|c3} PODS SKUS out.json
  python tools/dump_model.py c4sim PODS SKUS out.json [SIMULATION]     # a consolidation simulation: existing nodes
  python tools/dump_model.py fuzz SEED 0 out.json                      # tests/fuzz_models.py, with existing nodes

The dump carries what the Go harness needs to rebuild Karpenter objects (InstanceTypes with Offerings, NodePools,
Pods, existing nodes as state.StateNode fixtures) from the very arrays the encoder and the oracle consume, plus the
generator arguments, so that tools/compare_go_result.py can regenerate the same model here and hold the Go scheduler's
result against the oracle's.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
from b200sched import model as M, synth  # noqa: E402

OPS = {M.OP_IN: "In", M.OP_NOTIN: "NotIn", M.OP_EXISTS: "Exists", M.OP_DOESNOTEXIST: "DoesNotExist", M.OP_GT: "Gt",
       M.OP_LT: "Lt"}
EFFECTS = {M.EFFECT_NONE: "", M.EFFECT_NOSCHEDULE: "NoSchedule", M.EFFECT_PREFERNOSCHEDULE: "PreferNoSchedule",
           M.EFFECT_NOEXECUTE: "NoExecute"}
TOLOPS = {M.TOLOP_EMPTY: "", M.TOLOP_EQUAL: "Equal", M.TOLOP_EXISTS: "Exists", M.TOLOP_OTHER: "Other"}


def quantity(res: str, v: int) -> str:
    """the model's exact fixed-scale int64 back to a resource.Quantity string (cpu is kept in milli-units)"""
    return f"{int(v)}m" if res == "cpu" else str(int(v))


def tstr(i: int) -> str:
    """taint / toleration KEYS are interned to integers in the model; only equality matters to ToleratesTaint"""
    return "" if i == 0 else f"example.com/t{int(i)}"


def vstr(i: int) -> str:
    """... and so are their values (same intern table; a key never meets a value)"""
    return "" if i == 0 else f"v{int(i)}"


def dump(model: M.Model, generator: dict) -> dict:
    a = model.arrays

    def reqs(lo, hi):
        out = []
        for i in range(lo, hi):
            r = a["reqs"][i]
            op = int(r["op"])
            vals = [model.val_names[int(v)] for v in a["req_values"][int(r["val_off"]):int(r["val_off"]) + int(r["val_cnt"])]]
            if op in (M.OP_GT, M.OP_LT):
                vals = [str(int(r["bound"]))]
            out.append({"key": model.key_names[int(r["key"])], "operator": OPS[op], "values": vals})
        return out

    def resources(row):
        return {res: quantity(res, v) for res, v in zip(model.res_names, row)}

    def taints(lo, hi):
        return [{"key": tstr(t["key"]), "value": vstr(t["value"]), "effect": EFFECTS[int(t["effect"])]}
                for t in a["taints"][lo:hi]]

    kz, kc = model.key_names.index(M.KEY_ZONE), model.key_names.index(M.KEY_CAPACITY_TYPE)
    its = []
    for s in range(model.n_skus):
        offs = []
        for o in range(a["sku_off_off"][s], a["sku_off_off"][s + 1]):
            labels = {r["key"]: r["values"][0] for r in reqs(a["off_req_off"][o], a["off_req_off"][o + 1])}
            offs.append({"zone": labels[model.key_names[kz]], "capacity_type": labels[model.key_names[kc]],
                         "price": float(a["off_price"][o]), "available": bool(a["off_available"][o])})
        its.append({"name": model.sku_names[s], "requirements": reqs(a["sku_req_off"][s], a["sku_req_off"][s + 1]),
                    "capacity": resources(a["sku_capacity"][s]), "allocatable": resources(a["sku_alloc"][s]),
                    "offerings": offs})
    pools = []
    for p in range(int(model.c.n_pools)):
        pools.append({"name": model.pool_names[p], "weight": int(a["pool_weight"][p]),
                      "requirements": reqs(a["pool_req_off"][p], a["pool_req_off"][p + 1]),
                      "taints": taints(a["pool_taint_off"][p], a["pool_taint_off"][p + 1])})
    pods = []
    for p in range(model.n_pods):
        hi, lo = int(a["pod_uid_hi"][p]), int(a["pod_uid_lo"][p])
        h = f"{hi:016x}{lo:016x}"
        tols = [{"key": tstr(t["key"]), "operator": TOLOPS[int(t["op"])], "value": vstr(t["value"]),
                 "effect": EFFECTS[int(t["effect"])]} for t in a["tolerations"][a["pod_tol_off"][p]:a["pod_tol_off"][p + 1]]]
        pods.append({"index": p, "uid": f"{h[:8]}-{h[8:12]}-{h[12:16]}-{h[16:20]}-{h[20:]}",
                     "creation_timestamp": int(a["pod_ts"][p]), "requests": resources(a["pod_requests"][p]),
                     "requirements": reqs(a["pod_req_off"][p], a["pod_req_off"][p + 1]), "tolerations": tols})
    # existing nodes: labels (the model keeps them as single-value In requirements), what is left of their allocatable
    # (ExistingNode.remainingResources, existingnode.go:57), taints, and whether the node is initialized
    nodes = []
    for e in range(int(model.c.n_nodes)):
        labels = {r["key"]: r["values"][0] for r in reqs(a["node_req_off"][e], a["node_req_off"][e + 1])}
        nodes.append({"name": model.node_names[e], "labels": labels, "available": resources(a["node_available"][e]),
                      "taints": taints(a["node_taint_off"][e], a["node_taint_off"][e + 1]),
                      "initialized": bool(a["node_initialized"][e])})
    limits = []
    for p in range(int(model.c.n_pools)):
        limits.append({res: quantity(res, a["pool_limit"][p][r]) for r, res in enumerate(model.res_names) if a["pool_has_limit"][p][r]})
    for pool, lim in zip(pools, limits):
        pool["limits"] = lim
        pool["daemon_overhead"] = resources(a["pool_daemon"][model.pool_names.index(pool["name"])])
    return {"generator": generator, "resources": list(model.res_names),
            "wellknown_extra": sorted(model.wellknown_extra), "instance_types": its, "node_pools": pools, "pods": pods,
            "nodes": nodes}


def expressible(model: M.Model):
    """None, or why baseline/go/main.go cannot rebuild this model as Karpenter objects"""
    a = model.arrays
    pods_res = model.res_names.index("pods") if "pods" in model.res_names else -1
    for p in range(int(model.c.n_pools)):
        d = [int(x) for x in a["pool_daemon"][p]]
        if any(d) and (pods_res < 0 or d[pods_res] != 1):
            return (f"pool {model.pool_names[p]}: a daemon overhead is the requests of daemonset pods, and RequestsForPods counts "
                    "each of them as pods: 1 (resources.go:36) -- an overhead whose pods entry is not 1 has no pod set")
    for e in range(int(model.c.n_nodes)):
        if any(int(x) < 0 for x in a["node_available"][e]):
            return f"node {model.node_names[e]}: negative available resources need tracked pods on the StateNode fixture"
    return None


def generate(gen: dict) -> M.Model:
    """the model a dump's generator arguments name"""
    kind = gen["kind"]
    if kind in ("c2", "c3"):
        return synth.config(kind, gen["pods"], gen["skus"])
    if kind == "fuzz":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import fuzz_models
        return fuzz_models.random_model(gen["seed"], with_nodes=True, max_pods=gen.get("max_pods", 60))
    if kind == "c4sim":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding as ob
        model, node_of_pod, util = synth.c4(gen["pods"], gen["skus"], solve=ob.solve)
        return synth.c4_simulations(model, node_of_pod, util, gen["sim"] + 1)[gen["sim"]][1]
    raise SystemExit(f"unknown generator {kind}")


if __name__ == "__main__":
    kind, P, S, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    gen = {"kind": kind, "pods": P, "skus": S}
    if kind == "fuzz":
        gen = {"kind": "fuzz", "seed": P}
    elif kind == "c4sim":
        gen["sim"] = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    m = generate(gen)
    why = expressible(m)
    if why:
        raise SystemExit("not expressible for the Go harness: " + why)
    with open(out, "w") as f:
        json.dump(dump(m, gen), f)
    print(f"{out}: {m.n_pods} pods, {m.n_skus} instance types, {int(m.c.n_pools)} node pools, {int(m.c.n_nodes)} existing nodes")

"""Writes tests/golden/oracle_digests.json: SHA-256 digests of the CPU oracle's canonical Solve result on fixed
synthetic inputs.  Run from the repo root: python tests/golden/make_golden.py

These vectors come from OUR oracle (oracle/karp_oracle.cpp), not from the reference: gpu-provisioner is Go, no Go
toolchain exists in the build image and the vendored Karpenter tree carries no scheduler tests, so there is no
reference output to record (DESIGN.md "Oracle")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))

import oracle_binding as ob  # noqa: E402
import parity  # noqa: E402
from b200sched import synth  # noqa: E402

CASES = {"c2_small": ("c2", 1000, 100), "c2_mid": ("c2", 4000, 500), "c3_small": ("c3", 1500, 150),
         "c3_mid": ("c3", 6000, 1000)}
out = {"generator": "tests/golden/make_golden.py", "source": "oracle/karp_oracle.cpp", "cases": {}}
for name, (kind, P, S) in CASES.items():
    m = synth.config(kind, P, S)
    o = ob.solve(m)
    out["cases"][name] = dict(kind=kind, pods=P, skus=S, digest=parity.digest(parity.canon_oracle(m, o)),
                              n_claims=int(o.n_claims), n_unscheduled=int(o.n_unscheduled),
                              fit_evals=int(o.counters["fit_evals"]))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_digests.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

# ---- the widened rows: ToNodeClaim lists, consolidation commands, the multi-node search ----------
import hashlib  # noqa: E402

import numpy as np  # noqa: E402

import consolidation_ref as ref  # noqa: E402
from b200sched import consolidation as cons  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def widened():
    w = {"generator": "tests/golden/make_golden.py", "source": "oracle/karp_oracle.cpp"}
    m = synth.config("c3", 1500, 150)
    o, order, count, tie = ob.solve_ordered(m, 600)
    w["to_nodeclaim_c3_1500x150"] = dict(n_claims=int(o.n_claims), listed=int(count.sum()), ties=int(tie.sum()),
                                         digest=sha(order.astype(np.int32), count.astype(np.int32)))
    model, node_of_pod, util = synth.c4(1500, 120, solve=ob.solve)
    by_util = np.argsort(util, kind="stable")
    sub, nop = ref.sub_cluster(model, node_of_pod, by_util[:40])
    cand = list(range(40))
    ctv = parity.capacity_type_values(sub)
    facts = cons.node_facts(sub, cand)
    best, _ = ref.first_n_consolidation_option(sub, nop, cand, ctv)
    rows = []
    for mid in range(1, 40):
        _, cmd = ref.probe(sub, nop, cand, facts, mid, ctv, False)
        rows.append([mid, int(cmd["decision"]), int(cmd["reason"]), int(cmd["valid"]), int(cmd["n_options"]),
                     sha(cmd["mask"].astype(np.uint64))[:16]])
    w["multi_node_search_c4_1500x120_first40"] = dict(removes_first=None if best is None else best + 1, probes=rows)
    return w


json.dump(widened(), open(os.path.join(ROOT, "tests", "golden", "widened_digests.json"), "w"), indent=1)

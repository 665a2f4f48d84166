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

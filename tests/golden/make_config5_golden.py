"""Golden digest of BASELINE.json configs[4]'s global pack (1M pods x 2k SKUs), made with the device algorithm's host
twin (tests/native/emul.cpp emul_solve, one CPU thread, tens of minutes).  The reference-shaped oracle would need hours
at this size; the twin is checked against the oracle bit for bit on every smaller configuration (tests/test_emulation.py),
and the GPU Solve is checked against this digest in tests/test_gpu.py.

    python tests/golden/make_config5_golden.py        # rewrites tests/golden/config5_digest.json
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity  # noqa: E402
from b200sched import capi, synth  # noqa: E402

if __name__ == "__main__":
    P, S = 1_000_000, 2_000
    m = synth.config("c3", P, S)
    enc = capi.Encoded(m)
    t0 = time.time()
    raw = parity.emul_solve(enc)
    out = {"config": f"synth.config('c3', {P}, {S})", "n_claims": int(raw.n_claims), "n_unscheduled": int(raw.n_unscheduled),
           "digest": parity.digest(parity.canon_device(enc, raw))[:16], "digest_chars": 16, "made_by": "tests/native/emul.cpp emul_solve",
           "seconds": round(time.time() - t0, 1)}
    json.dump(out, open(os.path.join(HERE, "config5_digest.json"), "w"), indent=1)
    print(out)

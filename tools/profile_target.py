"""Small driver for ncu captures: python tools/profile_target.py {fit|solve} {c2|c3|c5} [pods] [skus] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
import __graft_entry__ as g  # noqa: E402

g.build()
from b200sched import capi, synth  # noqa: E402

mode, kind = sys.argv[1], sys.argv[2]
defaults = {"c2": (10_000, 500), "c3": (100_000, 1_000), "c5": (1_000_000, 2_000)}
P = int(sys.argv[3]) if len(sys.argv) > 3 else defaults[kind][0]
S = int(sys.argv[4]) if len(sys.argv) > 4 else defaults[kind][1]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
model = synth.config("c2" if kind == "c2" else "c3", P, S)
if os.environ.get("B200S_DEPLOY"):  # the pod list as Deployments of N consecutive replicas (bench.as_deployments)
    import bench
    model = bench.as_deployments(model, int(os.environ["B200S_DEPLOY"]))
enc = capi.Encoded(model)
ctx = capi.Context(0)
cid = ctx.upload_catalog(enc)
res = ctx.resident(cid, enc)
ev0, ev1 = ctx.event(), ctx.event()
if mode == "solve" and not os.environ.get("B200S_NO_PROFILE"):
    res.set_pack_profile(True)
for i in range(iters):
    ctx.flush_l2()
    ctx.record(ev0)
    (res.fit_matrix if mode == "fit" else res.solve)()
    ctx.record(ev1)
    print(f"{mode} {kind} {P}x{S} iter {i}: {ctx.elapsed_ms(ev0, ev1):.4f} ms", flush=True)
    if mode == "solve":
        pr = res.pack_profile()
        steps = max(pr["steps"], 1)
        print("   cycles/step:", {k: round(v / steps, 1) for k, v in pr.items() if k not in ("steps", "full_sorts")},
              "steps", steps, "full pdqsorts", pr["full_sorts"], flush=True)
res.close()
ctx.close()

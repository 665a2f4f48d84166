"""Small driver for ncu / compute-sanitizer captures:
  python tools/profile_target.py {fit|solve} {c2|c3|c5} [pods] [skus] [iters]
  python tools/profile_target.py consolidate c2 [unused] [skus]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
import __graft_entry__ as g  # noqa: E402

g.build()
from b200sched import capi, synth  # noqa: E402

mode, kind = sys.argv[1], sys.argv[2]
if mode == "consolidate":
    # a few consolidation simulations on a small catalog and every branch of the command kernel (sanitizer runs)
    import numpy as np
    from b200sched import model as M
    S = int(sys.argv[4]) if len(sys.argv) > 4 else 150
    b = synth.new_builder()
    synth.add_catalog(b, S)
    synth.add_pools(b, "c2")
    pods = synth.bulk_pods(b, 64, "c2")
    base = synth.build_model(b, pods)
    none = np.array([], np.int64)
    sims = [M.subset(base, np.arange(i, i + n), none) for i, n in ((0, 1), (1, 2), (3, 3), (6, 8), (14, 20), (34, 30))]
    encs = [capi.Encoded(m) for m in sims]
    ctx = capi.Context(0)
    cid = ctx.upload_catalog(encs[0])
    raws = ctx.solve_batch(cid, encs)
    key = base.key_names.index(M.KEY_CAPACITY_TYPE)
    ids = {base.val_names[v]: v for v, k in enumerate(base.arrays["val_key"]) if int(k) == key}
    spot, od = encs[0].value_bit(key, ids["spot"]), encs[0].value_bit(key, ids["on-demand"])
    rows, prices, spots, cands = [], [], [], []
    for r in raws:
        for price, all_spot, n_cand in ((1e3, True, 1), (1e3, True, 2), (1e3, False, 1), (1e-6, False, 1), (0.4, True, 1)):
            rows.append(r); prices.append(price); spots.append(all_spot); cands.append(n_cand)
    for gate in (False, True):
        c = ctx.consolidation_commands(cid, encs[0], rows, prices, spots, cands, spot, od, -1, gate)
        print("gate", gate, "decisions", np.bincount(c["decision"], minlength=3).tolist(), "reasons",
              np.bincount(c["reason"], minlength=6).tolist(), "truncated", int((c["head"][:, 0] >= 0).sum()), flush=True)
    for r, e in zip(raws, encs):  # ToNodeClaim's price-ordered lists, cut short and at MaxInstanceTypes
        for cut in (5, capi.MAX_INSTANCE_TYPES):
            order, count, tie = ctx.order_by_price(cid, e, r, cut)
            print("order_by_price", r.n_claims, "claims, cut", cut, "listed", count.tolist(), "ties", int(tie.sum()), flush=True)
    ctx.release_catalog(cid)
    ctx.close()
    sys.exit(0)
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
    if mode == "solve" and os.environ.get("B200S_REG_PROFILE"):
        import ctypes as C
        raw = (C.c_uint64 * 12)()
        capi.lib().b200s_resident_pack_profile(ctx.h, res.h, raw)
        steps = max(raw[3], 1)
        print("   reg pack cycles/step:", {k: round(raw[i] / steps, 1) for i, k in enumerate(
            ["claim_eval", "claim_barrier", "claim_commit", "steps", "ordw_wait", "ordw_work", "-", "claim_post",
             "claim_meet", "meets", "full_sorts", "qw_pre"])}, "steps", raw[3], "full sorts", raw[10], "meets", raw[9], flush=True)
    if mode == "solve" and not os.environ.get("B200S_NO_PROFILE"):
        pr = res.pack_profile()
        steps = max(pr["steps"], 1)
        print("   cycles/step:", {k: round(v / steps, 1) for k, v in pr.items() if k not in ("steps", "full_sorts")},
              "steps", steps, "full pdqsorts", pr["full_sorts"], flush=True)
res.close()
ctx.close()

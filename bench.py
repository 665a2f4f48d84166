#!/usr/bin/env python
"""bench.py — pod x SKU fit evaluations/s and NodeClaims scheduled/s of the B200 path, next to the CPU oracle.

Default workload = BASELINE.json configs[1]: 10 000 synthetic pending pods x 500 Azure SKUs, one NodePool.
One step = one pass of the hot path over that batch = one b200s_solve: class tables + dense pod x SKU fit matrix
(stage A), FFD order, bin-pack into NodeClaims, cheapest SKU per NodeClaim.

  value     evaluations/s, device-timed with CUDA events on the library's stream, inputs resident in HBM, L2 flushed
            between iterations.  Evaluations per step = the P x S-after-prefilter instance types the dense stage
            decides + the instance types the reference's in-flight CanAdd calls visit (the ORACLE's count, credited;
            the device's own count is larger because it evaluates claims in parallel).
  e2e       the same through the host-buffer C-ABI call (b200s_solve): H2D of the pod / class / node SoA and D2H of
            the result inside the timed region.
  roofline  the dense kernel (the HBM-bound kernel of the path) alone: algorithmic bytes / its own CUDA-event time
            vs MEASURED_PEAKS.json hbm_gbs.  config[1] moves < 1 MB per launch, so the same kernel is also timed on
            configs[2] (100k x 1k) and configs[4]'s size (1M x 2k, one GPU): `roofline_at_scale`, `roofline_large`.
  --impl reference   the CPU oracle (reference-shaped restatement of the Go scheduler; the reference itself is Go and
            cannot be built here) on the same workload, same metric, one host core (the reference's default: one worker).
  --workload c5      1M pods x 2k SKUs, dense stage row-sharded over the ranks + one NCCL all-gather of the result rows.

Under torchrun with the default workload every rank runs an independent replica ("replicas only": the pack is
sequential per Solve) and rank 0 prints the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "pod x SKU fit evaluations/s"
WORKLOAD = "config[1]: 10k pending pods x 500 Azure SKUs, 1 NodePool; dense fit matrix + FFD pack + cheapest SKU (one Solve)"


_REAL_STDOUT = None


def emit(obj) -> None:
    """The run's one JSON line, on the process's original stdout."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def dense_bytes(enc) -> int:
    """Algorithmic bytes of one dense launch with this repo's layout: every pod row (two class
    ids) read once, every output (SKU mask, first feasible template, per-template codes) written once, every class
    row (C&O per signature x template; F per request vector x template; status) read once."""
    pod_row = 4 + 4
    row = (enc.S + 7) // 8  # S bits; the padding up to W 64-bit words is not credited
    out_row = row + 4 + enc.T
    tables = (enc.NS + enc.NQ) * enc.T * row + enc.NS * enc.T  # C&O and F rows + status (C, O alone: failures only)
    return enc.P * (pod_row + out_row) + tables


def time_device(ctx, fn, steps, warmup, inner=None):
    """Average CUDA-event time of fn over `steps` launches on the library's stream.  `inner` (optional) is read after
    every timed launch and averaged too: the library's own event pair around one kernel inside fn."""
    ev0, ev1 = ctx.event(), ctx.event()
    for _ in range(warmup):
        ctx.flush_l2()
        fn()
    ctx.synchronize()
    total = inner_total = 0.0
    for _ in range(steps):
        ctx.flush_l2()  # L2 flush between timed iterations (256 MiB memset)
        ctx.record(ev0)
        fn()
        ctx.record(ev1)
        total += ctx.elapsed_ms(ev0, ev1)
        if inner is not None:
            inner_total += inner()
    return (total / steps, inner_total / steps) if inner is not None else total / steps


def config1_latency():
    """BASELINE.json configs[0] / BASELINE.md section 3: the vmSize pick of instance.Provider.Create for one NodeClaim x
    10 Azure GPU SKUs -- CPU-only by design: b200s_create_select (host C++) next to the oracle's restatement
    of pkg/providers/instance/instance.go:76-95, microseconds per call."""
    import ctypes as C
    import oracle_binding as ob
    from b200sched import capi, synth
    names = synth.REAL_GPU_SKUS[:10]
    prices = [3.0 + 0.37 * ((7 * i) % 10) for i in range(10)]
    storage = 120 << 30
    n = 20000
    got = capi.create_select("ws0123456789", names, storage, prices)
    arr = (C.c_char_p * 10)(*[s_.encode() for s_ in names])
    pr = (C.c_double * 10)(*prices)
    ch, disk, gpu = C.c_int32(-1), C.c_int32(0), C.c_int32(0)
    f = capi.lib().b200s_create_select
    t0 = time.perf_counter()
    for _ in range(n):
        rc = f(b"ws0123456789", 10, arr, pr, storage, C.byref(ch), C.byref(disk), C.byref(gpu))
    us = 1e6 * (time.perf_counter() - t0) / n
    assert (rc, ch.value, disk.value, gpu.value) == got and rc == 0
    f = ob.lib().oracle_create_select
    t0 = time.perf_counter()
    for _ in range(n):
        rc = f(b"ws0123456789", 10, arr, pr, storage, C.byref(ch), C.byref(disk), C.byref(gpu))
    us_ref = 1e6 * (time.perf_counter() - t0) / n
    assert (rc, ch.value, disk.value, gpu.value) == got
    return {"metric": "NodeClaim -> vmSize selection latency", "unit": "us", "value": us, "reference_port_us": us_ref,
            "selected": names[got[1]], "calls": n,
            "note": "through ctypes both ways (the call overhead dominates: the selection itself is ten comparisons); "
                    "config[0] has no device part"}


def oracle_counts(model, threads=1):
    """One Solve (sequential: every placement depends on the previous ones; the reference's own in-Solve workers
    default to one) and one dense pass (independent rows: split over `threads` host threads)."""
    import oracle_binding as ob
    t0 = time.perf_counter()
    o = ob.solve(model)
    t_solve = time.perf_counter() - t0
    t0 = time.perf_counter()
    if threads > 1:
        f = (None, None, None, ob.fit_matrix_threads(model, threads))
    else:
        f = ob.fit_matrix(model)
    t_fit = time.perf_counter() - t0
    credited = f[3] + o.counters["fit_evals"] - o.counters["newclaim_evals"]
    return o, f, credited, t_solve, t_fit


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank):
    import oracle_binding as ob
    from b200sched import synth
    if rank != 0:
        return
    model = synth.c2()
    ob.lib()
    times = []
    credited = nclaims = 0
    solve_s = fit_s = 0.0
    nthreads = host_threads()
    for it in range(args.warmup + args.steps):
        o, f, credited, t_solve, t_fit = oracle_counts(model, nthreads)
        if it >= args.warmup:
            times.append(t_solve + t_fit)
            solve_s += t_solve
            fit_s += t_fit
        nclaims = o.n_claims
    ms = 1e3 * sum(times) / len(times)
    value = credited / (ms / 1e3)
    emit({
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/u64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "pods": model.n_pods, "skus": model.n_skus},
        "nodeclaims_per_s": nclaims / (ms / 1e3), "evals_per_step": int(credited),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": nthreads, "kind": "port",
                         "sample": "the whole workload per step: one Solve on one thread (a sequential chain; the "
                                   "reference's in-Solve workers default to one, K/operator/options/options.go:122) + "
                                   "one dense pass split over all host threads",
                         "solve_ms": 1e3 * solve_s / max(len(times), 1), "dense_ms": 1e3 * fit_s / max(len(times), 1),
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def as_deployments(model, replicas: int, seed: int = 7):
    """The same pod population as Deployments: P/replicas pods drawn at random, each listed `replicas` times in a row
    (replicas of one Deployment share spec and arrive consecutively in a real pod list).  Only the dense stage is
    measured on it: UIDs repeat, which the fit matrix does not look at."""
    import numpy as _np
    from b200sched import model as M
    P = model.n_pods
    rng = _np.random.default_rng(seed)
    base = rng.integers(0, P, size=(P + replicas - 1) // replicas)
    perm = _np.repeat(base, replicas)[:P]
    return M.subset(model, perm, _np.zeros(0, _np.int64))


def dense_leg(ctx, capi, synth, kind, P, S, steps, warmup, peak, run=0):
    """The dense stage alone on a larger config: the whole stage (template pre-filter, class tables, dense kernel)
    timed by events around the call, the dense kernel by the library's own event pair around that one launch."""
    model = synth.config(kind, P, S)
    enc = capi.Encoded(model)
    if run:
        model = as_deployments(model, run)
        enc = capi.Encoded(model)
    cid = ctx.upload_catalog(enc)
    res = ctx.resident(cid, enc)
    ms, ms_k = time_device(ctx, res.fit_matrix, steps, warmup, inner=res.dense_kernel_ms)
    ab = dense_bytes(enc)
    out = {"workload": f"{P} pods x {S} SKUs ({'4 tainted NodePools, selectors' if kind == 'c3' else '1 NodePool'}), one GPU"
                       + (f"; the pod list is {P // run} Deployments of {run} replicas each, replicas consecutive" if run else
                          "; every pod's class drawn independently (worst case for the class-row gathers)"),
           "signatures": enc.NS, "request_classes": enc.NQ, "algorithmic_bytes": ab,
           "kernel": "fit_dense_vec_kernel", "kernel_ms": ms_k, "achieved": ab / (ms_k * 1e-3) / 1e9,
           "frac": ab / (ms_k * 1e-3) / 1e9 / peak, "evals_per_s_kernel": P * S / (ms_k * 1e-3),
           "stage_ms": ms, "evals_per_s_stage": P * S / (ms * 1e-3), "achieved_gbs_stage": ab / (ms * 1e-3) / 1e9,
           "frac_stage": ab / (ms * 1e-3) / 1e9 / peak,
           "note": "kernel = the dense kernel alone (library event pair on its stream); stage = template pre-filter + "
                   "class tables + dense kernel (3 launches)"}
    res.close()
    ctx.release_catalog(cid)
    return out


C5_WORKLOAD = ("config[4]: 1M pods x 2k SKUs (4 tainted NodePools, selectors), pod rows sharded over the ranks: per rank "
               "class tables + dense pod x SKU rows of its shard, ONE NCCL all-gather of the compact per-pod results")


def c5_config(P, S):
    return {"workload": C5_WORKLOAD, "pods": P, "skus": S}


def run_c5(args, rank, world, local):
    """configs[4] (north_star's multi-GPU partitioning): 1M pods x 2k SKUs, the pod x SKU matrix sharded by pod rows, one
    all-gather of the compact per-pod results (first feasible template + attempt codes: 8 B/pod -- the instance-type rows
    stay on the GPU that computed them and are re-derived from the replicated class tables where the pack opens a claim),
    then the global packing of all pods on rank 0.  STRONG scaling: the total work is fixed.  The timed step is the
    sharded stage + the all-gather + the import on the packing rank; the same stage on ONE GPU is timed in the same run
    (rank 0, the others idle) so that the ratio can be read off one JSON line; the 1M-pod pack runs once, untimed steps
    aside, and is reported with its invariants."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from b200sched import capi, model as M, sharding, synth
    import oracle_binding as ob
    import parity
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, S = args.pods or 1_000_000, args.skus or 2_000
    full = synth.config("c3", P, S)
    lo, hi = sharding.row_shard(P, rank, world)
    shard = M.slice_pods(full, lo, hi) if world > 1 else full
    enc = capi.Encoded(shard)
    ctx = capi.Context(local)
    cid = ctx.upload_catalog(enc)
    res = ctx.resident(cid, enc)
    T, W = enc.T, enc.W
    rb = capi.compact_row_bytes(T)
    rows = max(sharding.row_shard(P, r, world)[1] - sharding.row_shard(P, r, world)[0] for r in range(world))
    d_send = torch.zeros(rows * rb, dtype=torch.uint8, device="cuda")
    g_all = torch.zeros(world * rows * rb, dtype=torch.uint8, device="cuda")
    lib_stream = torch.cuda.ExternalStream(ctx.stream())  # the collective is enqueued behind the kernels, no host sync
    # the packing rank: all pods, per-pod results imported
    enc_full = res_full = cid_full = None
    if rank == 0:
        enc_full = enc if world == 1 else capi.Encoded(full)
        cid_full = cid if world == 1 else ctx.upload_catalog(enc_full)
        res_full = res if world == 1 else ctx.resident(cid_full, enc_full)
    ev0, ev1, ev2, ev3 = ctx.event(), ctx.event(), ctx.event(), ctx.event()

    # The exchange fused with the export: every rank maps every rank's gathered buffer (CUDA IPC) and stores its rows
    # straight into all of them, then a flag barrier (b200s_resident_export_compact_peers).  Falls back to the NCCL
    # collective if the mapping fails, a peer times out, or the rows differ from the collective's.
    p2p = None
    if world > 1 and world <= 8 and not args.no_p2p:
        try:
            buf, flg, tmo = ctx.device_alloc(P * rb), ctx.device_alloc(64), ctx.device_alloc(4)
            handles = [None] * world
            dist.all_gather_object(handles, (ctx.ipc_export(buf), ctx.ipc_export(flg)))
            p2p = {"buf": buf, "tmo": tmo, "epoch": 0,
                   "out": [buf if q == rank else ctx.ipc_open(handles[q][0]) for q in range(world)],
                   "flags": [flg if q == rank else ctx.ipc_open(handles[q][1]) for q in range(world)]}
        except Exception as e:  # noqa: BLE001 -- any failure here means "use the collective"
            print(f"[bench] rank {rank}: peer mapping unavailable ({e}); NCCL all-gather only", file=sys.stderr)
            p2p = None
        ok = torch.tensor([1 if p2p else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            p2p = None

    def gathered_rows():
        if P % world == 0:
            return g_all
        v = g_all.view(world, rows, rb)
        return torch.cat([v[r, :sharding.row_shard(P, r, world)[1] - sharding.row_shard(P, r, world)[0]] for r in range(world)]).contiguous()

    def step(peers=False):
        ctx.flush_l2()
        ctx.record(ev0)
        res.fit_matrix()                      # prefilter + class tables + dense rows of this rank's shard
        if peers:                             # rows stored into every rank's buffer + flag barrier: export and exchange in one
            ctx.record(ev1)
            p2p["epoch"] += 1
            res.export_compact_peers(p2p["out"], p2p["flags"], rank, lo, p2p["epoch"], p2p["tmo"])
        else:
            res.export_compact(d_send.data_ptr())
            ctx.record(ev1)
            if world > 1:
                with torch.cuda.stream(lib_stream):
                    dist.all_gather_into_tensor(g_all, d_send)
        ctx.record(ev2)
        if rank == 0 and world > 1:
            res_full.import_compact(p2p["buf"] if peers else gathered_rows().data_ptr())
        ctx.record(ev3)
        return np.array([ctx.elapsed_ms(ev0, ev3), ctx.elapsed_ms(ev0, ev1), ctx.elapsed_ms(ev1, ev2), ctx.elapsed_ms(ev2, ev3)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    acc = np.zeros(4)
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            acc += step()
            if world > 1:
                dist.barrier()  # every step starts together: the time of a step is the max over the ranks
    acc /= args.steps
    acc_nccl = acc.copy()
    collective = "one NCCL all_gather_into_tensor of the compact rows" if world > 1 else "none (one GPU)"
    if p2p:
        # the same step with the peer-store export; valid only if no rank timed out and every rank's buffer equals the
        # collective's result
        for _ in range(3):
            step(peers=True)
            dist.barrier()
        check = torch.zeros(P * rb, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.copy_device(check.data_ptr(), p2p["buf"], P * rb)
        tmo_host = torch.zeros(1, dtype=torch.int32, device="cuda")
        ctx.copy_device(tmo_host.data_ptr(), p2p["tmo"], 4)
        ctx.synchronize()
        same = torch.equal(check.view(P, rb)[:, :T + 1], gathered_rows().view(P, rb)[:, :T + 1]) and int(tmo_host.item()) == 0
        ok = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            acc_p = np.zeros(4)
            for _ in range(args.steps):
                acc_p += step(peers=True)
                dist.barrier()
            acc_p /= args.steps
            t = torch.tensor(acc_p, device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            acc_p = t.cpu().numpy()
            t = torch.tensor(acc_nccl, device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if acc_p[0] < float(t[0].item()):
                acc = acc_p
                collective = ("none: every rank stores its 8 B/pod rows into every rank's buffer through CUDA-IPC peer mappings "
                              "(NVLink) and the ranks meet at a flag barrier (b200s_resident_export_compact_peers)")
            else:
                p2p["slower"] = acc_p
            p2p["acc"] = acc_p
        else:
            print(f"[bench] rank {rank}: peer-store export rejected (rows differ or a peer timed out)", file=sys.stderr)
            p2p = None
    dense_k = res.dense_kernel_ms()
    launches_stage = res.launches()
    # end to end through host buffers: this rank's shard in, its rows out (b200s_fit_matrix), then the same all-gather
    t_e2e = []
    for it in range(2 + 3):
        barrier()
        t0 = time.perf_counter()
        fit = ctx.fit_matrix(cid, enc)
        if world > 1:
            with torch.cuda.stream(lib_stream):
                dist.all_gather_into_tensor(g_all, d_send)
            ctx.synchronize()
        if it >= 2:
            t_e2e.append(1e3 * (time.perf_counter() - t0))
    ms_e2e = float(np.mean(t_e2e))
    acc_nccl_max = acc_nccl
    if world > 1:
        t = torch.tensor(list(acc) + [ms_e2e, dense_k] + list(acc_nccl), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v = t.cpu().numpy()
        acc, ms_e2e, dense_k, acc_nccl_max = v[:4], float(v[4]), float(v[5]), v[6:10]
    # parity: this rank's first and last rows (the shard boundaries) against the oracle
    fit = res.fetch_fit()
    cf = parity.canon_fit(enc, fit)
    for s0, s1 in ((0, min(1000, enc.P)), (max(enc.P - 1000, 0), enc.P)):
        o_tmpl, o_mask, o_code, _, _ = ob.fit_matrix(shard, s0, s1)
        assert np.array_equal(cf["tmpl"][s0:s1], o_tmpl) and np.array_equal(cf["mask"][s0:s1], o_mask) and \
            np.array_equal(cf["code"][s0:s1], o_code), f"rank {rank}: shard boundary rows differ from the oracle"
    single = pack = None
    cpu = None
    if rank == 0:
        # the same stage on ONE GPU, same run (the other ranks idle at the barrier below)
        def one():
            ctx.flush_l2()
            ctx.record(ev0)
            res_full.fit_matrix()
            res_full.export_compact(g_all.data_ptr() if world == 1 else gathered_rows().data_ptr())
            ctx.record(ev1)
            return ctx.elapsed_ms(ev0, ev1)
        if world > 1:
            g_one = torch.zeros(P * rb, dtype=torch.uint8, device="cuda")
            def one():  # noqa: F811
                ctx.flush_l2()
                ctx.record(ev0)
                res_full.fit_matrix()
                res_full.export_compact(g_one.data_ptr())
                ctx.record(ev1)
                return ctx.elapsed_ms(ev0, ev1)
        for _ in range(3):
            one()
        ms_one = float(np.mean([one() for _ in range(args.steps)]))
        single = {"ms_per_step": ms_one, "value": P * S / (ms_one * 1e-3), "dense_kernel_ms": res_full.dense_kernel_ms(),
                  "note": "prefilter + class tables + dense rows of ALL pods + compact export on rank 0's GPU alone"}
        if world > 1:  # the all-gathered per-pod results equal what one GPU computes for all pods
            assert torch.equal(gathered_rows().view(P, rb)[:, :T + 1], g_one.view(P, rb)[:, :T + 1]), "gathered rows differ"
        # "merge the global packing": FFD order + pack + cheapest SKU over all pods from the imported per-pod results
        if args.no_global_pack:
            pack = {"skipped": "--no-global-pack"}
        else:
            res_full.import_compact(p2p["buf"] if (p2p and "acc" in p2p) else (gathered_rows() if world > 1 else g_all).data_ptr())
            ctx.synchronize()
            t0 = time.perf_counter()
            res_full.pack()
            ctx.synchronize()
            pack_s = time.perf_counter() - t0
            raw = res_full.fetch_solve()
            placed = int((raw.pod_place_kind == capi.PLACE_CLAIM).sum())
            assert placed + raw.n_unscheduled == P and int(raw.claim_npods.sum()) == placed
            assert (raw.claim_sku >= 0).all() and (raw.claim_n_options > 0).all()
            pack = {"ms": 1e3 * pack_s, "pods": P, "nodeclaims": int(raw.n_claims), "unschedulable": int(raw.n_unscheduled),
                    "nodeclaims_per_s": raw.n_claims / pack_s, "digest": parity.digest(parity.canon_device(enc_full, raw))[:16],
                    "note": "one Solve over all pods on rank 0 from the imported per-pod results (pods with selectors: the "
                            "generic pack kernel, one CTA); checked here by counts, bit-for-bit in tests/test_gpu.py"}
        # CPU baseline: the reference-shaped dense pass over a bounded sample of the rows, all host threads
        nthreads = host_threads()
        sample = M.slice_pods(full, 0, min(P, 40_000))
        t0 = time.perf_counter()
        ob.fit_matrix_threads(sample, nthreads)
        dt = time.perf_counter() - t0
        cpu = {"value": sample.n_pods * S / dt, "unit": "evals/s", "cores": nthreads, "kind": "port",
               "sample": f"the dense pass (filterInstanceTypesByRequirements per pod x template) over the first {sample.n_pods} "
                         "of the pods, split over all host threads", "ms": 1e3 * dt, "host_cores": os.cpu_count()}
    barrier()
    peak, src = peaks()
    ms = float(acc[0])
    ab = dense_bytes(enc)
    line = {"metric": METRIC, "value": P * S / (ms * 1e-3), "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": c5_config(P, S),
            "l2": "flushed between timed iterations (256 MiB memset)",
            "stage_ms_max_over_ranks": float(acc[1]), "allgather_ms": float(acc[2]), "import_ms": float(acc[3]),
            "exchange": collective,
            "exchange_nccl": None if world == 1 else {"ms_per_step": float(acc_nccl_max[0]), "stage_ms": float(acc_nccl_max[1]),
                                                      "allgather_ms": float(acc_nccl_max[2]), "import_ms": float(acc_nccl_max[3]),
                                                      "note": "stage = prefilter + class tables + dense rows + compact export; then one all_gather_into_tensor"},
            "exchange_peer_store": None if not (p2p and "acc" in p2p) else {
                "ms_per_step": float(p2p["acc"][0]), "stage_ms": float(p2p["acc"][1]), "export_and_exchange_ms": float(p2p["acc"][2]),
                "import_ms": float(p2p["acc"][3]),
                "note": "stage = prefilter + class tables + dense rows; then ONE launch pair stores the rows into every rank's buffer and runs the flag barrier; rows checked equal to the collective's on every rank"},
            "dense_kernel_ms_max_over_ranks": dense_k,
            "allgather_bytes_per_rank": int(rows * rb), "allgather_bytes_per_pod": rb,
            "single_gpu_same_run": single, "global_pack": pack,
            "roofline": {"bound": "hbm", "kernel": "fit_dense_vec_kernel on this rank's shard", "achieved": ab / (dense_k * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": ab / (dense_k * 1e-3) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes": ab, "kernel_ms": dense_k, "peak_source": src,
                         "note": "per-shard launch: 1/N of the rows, so the launch shrinks with N (at N=8 ~11 us of a "
                                 "~60 us stage); the single-GPU figure is in the N=1 line's roofline"},
            "cpu_baseline": cpu,
            "e2e": {"value": P * S / (ms_e2e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(enc.P * 32 + enc.NS * (enc.K * 8 + 16 + enc.NB * 8) + enc.NQ * enc.R * 8),
                    "d2h_bytes_per_step": int(enc.P * (4 + W * 8 + T)),
                    "note": "per rank: b200s_fit_matrix with host buffers (shard in, its rows out) + the all-gather; max over ranks"},
            "gpu_launches": int((args.steps + max(args.warmup, 3)) * (launches_stage + 1)),
            "clocks": clk.summary(),
            "parity": "every rank's first and last 1000 rows identical to the CPU oracle; gathered per-pod results identical to the single-GPU results"}
    res.close()
    if res_full is not None and res_full is not res:
        res_full.close()
    ctx.close()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_reference_c5(args, rank):
    """the reference arm of the multi-GPU workload: the CPU oracle's dense pass (the work the sharded stage does) on a
    bounded sample of the 1M x 2k rows, all host threads, same metric and config"""
    import oracle_binding as ob
    from b200sched import model as M, synth
    if rank != 0:
        return
    P, S = args.pods or 1_000_000, args.skus or 2_000
    n = min(P, 40_000)
    sample = synth.config("c3", n, S)  # the same generator streams: the first n pods of the 1M-pod list
    ob.lib()
    nthreads = host_threads()
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        ob.fit_matrix_threads(sample, nthreads)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
        if sum(times) > 120:
            break
    dt = sum(times) / len(times)
    value = n * S / dt
    emit({"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
          "ms_per_step": 1e3 * dt * (P / n), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
          "data": "synthetic", "impl": "reference", "config": c5_config(P, S),
          "cpu_baseline": {"value": value, "unit": "evals/s", "cores": nthreads, "kind": "port",
                           "sample": f"each step: the dense pass over {n} of the {P} pods ({n * S} evaluations), split over all "
                                     "host threads; ms_per_step is scaled to the whole workload", "host_cores": os.cpu_count()},
          "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_c4(args, rank, world, local):
    """configs[3], consolidation re-pack: 50k running pods on loosely packed nodes; a batch of single-node
    consolidation simulations (one CTA per simulation, b200s_solve_batch) plus one multi-node re-pack."""
    import __graft_entry__ as g
    g.build()
    import oracle_binding as ob
    import parity
    from b200sched import capi, model as M, sharding, synth
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:  # simulations are independent: dealt round-robin to the ranks, five words per simulation all-gathered
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P, S, B = args.pods or 50_000, args.skus or 500, args.batch
    ctx = capi.Context(local)
    t0 = time.perf_counter()

    def device_solve(m):  # the loose initial packing is itself produced by the device path
        e = capi.Encoded(m)
        cid0 = ctx.upload_catalog(e)
        r = ctx.solve(cid0, e)
        ctx.release_catalog(cid0)

        class R:
            pass
        out = R()
        out.claim_sku, out.claim_requests = r.claim_sku, r.claim_requests
        out.pod_place_index = np.where(r.pod_place_kind == capi.PLACE_CLAIM, r.pod_place_index, -1)
        return out

    model, node_of_pod, util = synth.c4(P, S, solve=device_solve)
    E = int(model.c.n_nodes)
    all_sims = synth.c4_simulations(model, node_of_pod, util, B)
    mine = sharding.deal(B, rank, world)            # every rank builds the same cluster and takes its share
    sims = [all_sims[int(j)] for j in mine]
    Bm = len(sims)
    encs = [capi.Encoded(m) for _, m in sims]
    # the same simulations as ONE model (every node, the candidates' pods one candidate after the other): one encode, one
    # upload of the cluster's nodes -- what b200s_solve_batch_shared takes
    t1 = time.perf_counter()
    shared_model, pod_off, cands = synth.c4_shared(model, node_of_pod, [i for i, _ in sims])
    t_flatten_shared = time.perf_counter() - t1   # numpy subsetting of the cluster model: the Go shim's flatten, not the library
    capi.Encoded(shared_model)                    # (first call: page in)
    t1 = time.perf_counter()
    enc_all = capi.Encoded(shared_model)
    t_encode_shared = time.perf_counter() - t1    # b200s_encode of the batch's one model
    assert enc_all.layout_hash == encs[0].layout_hash
    excl = [[r] for r in enc_all.node_rows(cands)]
    sample = list(range(0, Bm, max(Bm // 32, 1)))[:32]  # a spread of this rank's simulations: parity check, CPU timing
    t_build = time.perf_counter() - t0
    cid = ctx.upload_catalog(encs[0])
    # parity of a sample of simulations against the oracle, and the reference arm's time on the same sample
    oracle_out, t_cpu = [], 0.0
    if rank == 0:
        t0 = time.perf_counter()
        oracle_out = [ob.solve(sims[j][1]) for j in sample]
        t_cpu = (time.perf_counter() - t0) / len(oracle_out)
    # what computeConsolidation needs besides the Solve results: the candidates' price and capacity type
    cand_price, cand_spot = synth.candidate_prices(model, [i for i, _ in sims])
    ct_key, v_spot, v_od, v_res = parity.capacity_type_values(model)
    ct_bits = [encs[0].value_bit(ct_key, v) if v >= 0 else -1 for v in (v_spot, v_od, v_res)]

    rows = sharding.dealt_rows(B, world)
    d_mine = torch.zeros((rows, 5), dtype=torch.int32, device="cuda")
    d_all = torch.zeros((world * rows, 5), dtype=torch.int32, device="cuda")

    def step(separate=False):
        outs = ctx.solve_batch(cid, encs) if separate else ctx.solve_batch_shared(cid, enc_all, pod_off, excl)
        cmds = ctx.consolidation_commands(cid, encs[0], outs, cand_price, cand_spot, [1] * Bm, ct_bits[0], ct_bits[1],
                                          ct_bits[2], spot_to_spot=True)
        table = None
        if world > 1:  # decision, reason, options, new NodeClaims, unschedulable pods of every simulation, to every rank
            mine_rows = np.stack([cmds["decision"], cmds["reason"], cmds["n_options"],
                                  np.array([r.n_claims for r in outs], np.int32),
                                  np.array([r.n_unscheduled for r in outs], np.int32)], axis=1).astype(np.int32)
            d_mine[:Bm].copy_(torch.from_numpy(mine_rows), non_blocking=False)
            dist.all_gather_into_tensor(d_all, d_mine)
            table = sharding.undeal(d_all.cpu().numpy().reshape(world, rows, 5), B, world)
        return outs, cmds, table

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        outs, cmds, table = step()
    barrier()
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            outs, cmds, table = step()
        barrier()
        ms = 1e3 * (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            ctx.close()
            return
        deleted_all = int((table[:, 0] == capi.CMD_DELETE).sum())
    for j in sample:
        i, m = sims[j]
        want = ob.consolidate(m, cand_price[j], cand_spot[j], 1, v_spot, v_od, v_res, True)[1]
        got = parity.device_commands_as_oracle(encs[0], cmds, j)
        for f in ("decision", "reason", "n_options", "spot_required", "price_tie"):
            assert got[f] == want[f], (f"c4 command of node {i}", f, got, want)
        assert got["mask"] == want["mask"].tolist() and got["head"] == want["head"].tolist(), f"c4 command of node {i}"
    # the separately flattened form of the same batch (round 1's step): against the oracle, and the shared form against it
    step(separate=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs_sep, _, _ = step(separate=True)
    torch.cuda.synchronize()
    ms_separate = 1e3 * (time.perf_counter() - t0)
    for j, o in zip(sample, oracle_out):
        parity.assert_same(parity.canon_device(encs[j], outs_sep[j]), parity.canon_oracle(sims[j][1], o),
                           f"c4 simulation of node {sims[j][0]}")
    for j in range(Bm):
        parity.assert_shared_matches_separate(enc_all, encs[j], sims[j][0], outs_sep[j], outs[j], f"c4 simulation of node {sims[j][0]}")
    removable = (int(((table[:, 3] == 0) & (table[:, 4] == 0)).sum()) if world > 1 else
                 sum(1 for r in outs if r.n_claims == 0 and r.n_unscheduled == 0))
    # multi-node re-pack: every pod of the candidates at once against the remaining nodes
    cand = np.array([i for i, _ in all_sims])
    keep_nodes = np.setdiff1d(np.arange(E), cand)
    pods = np.nonzero(np.isin(node_of_pod, cand))[0]
    multi = M.subset(model, pods, keep_nodes)
    em = capi.Encoded(multi)
    rm = ctx.solve(cid, em)
    om = ob.solve(multi)
    parity.assert_same(parity.canon_device(em, rm), parity.canon_oracle(multi, om), "c4 multi-node re-pack")
    # firstNConsolidationOption over the 100 cheapest-to-disrupt candidates: every prefix in one batch on the device,
    # against the reference's sequential binary search on the CPU oracle (outside the timed region; wall clock)
    import consolidation_ref as cref
    from b200sched import consolidation as cons
    ctv = (ct_key, v_spot, v_od, v_res)
    cand100 = [int(i) for i in np.argsort(util, kind="stable")[:100]]
    search_t = {}
    cons.first_n_consolidation_option(ctx, cid, model, node_of_pod, cand100, ctv)  # warm-up (device block cache)
    got = cons.first_n_consolidation_option(ctx, cid, model, node_of_pod, cand100, ctv, timing=search_t)
    # the same search with every prefix as ranges of ONE model (one encode, the nodes once: b200s_solve_batch_shared)
    search_shared = None
    try:
        ts = {}
        cons.first_n_consolidation_option(ctx, cid, model, node_of_pod, cand100, ctv, shared=True)
        one = cons.first_n_consolidation_option(ctx, cid, model, node_of_pod, cand100, ctv, timing=ts, shared=True)
        same = (one is None) == (got is None) and (one is None or (one.mid, one.decision, one.n_options, one.mask.tolist()) ==
                                                   (got.mid, got.decision, got.n_options, got.mask.tolist()))
        search_shared = {"same_result": bool(same), "one_model_numpy_s": ts["models_s"], "b200s_encode_s": ts["encode_s"],
                         "library_calls_s": ts["device_calls_s"]}
    except Exception as e:  # noqa: BLE001 -- a secondary measurement must not take the line down
        search_shared = {"error": repr(e)[:200]}
    t0 = time.perf_counter()
    want_mid, probed = cref.first_n_consolidation_option(model, node_of_pod, cand100, ctv)
    t_search_cpu = time.perf_counter() - t0
    assert (got.mid if got is not None else None) == want_mid, "multi-node search ends on a different prefix"
    if got is not None:
        w = probed[want_mid][1]
        assert (got.decision, got.n_options, got.mask.tolist()) == (w["decision"], w["n_options"], w["mask"].tolist())
    evals_cpu = sum(o.counters["fit_evals"] for o in oracle_out) / len(oracle_out)
    line = {"metric": "consolidation simulations/s", "value": B / (ms * 1e-3), "unit": "simulations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "int64/u64", "data": "synthetic",
            "config": {"workload": f"config[3]: {P} running pods on {E} loosely packed nodes ({S} SKUs); {B} single-node "
                                   "consolidation simulations per step (lowest utilisation first) through b200s_solve_batch_shared (one pod set, the cluster's nodes handed over once) and their commands through "
                                   "b200s_consolidation_commands (host buffers in and out); plus one multi-node re-pack "
                                   "outside the timed region", "build_s": t_build,
                       "multi_gpu": "one process per GPU; the simulations are dealt round-robin to the ranks and the five "
                                    "result words per simulation all-gathered (NCCL) inside the step; wall time, max over "
                                    "ranks" if world > 1 else "single GPU"},
            "separately_flattened": {"ms_per_step": ms_separate, "note": "the same batch as 256 separately flattened models through b200s_solve_batch (round 1's step): every simulation uploads its own copy of the cluster's nodes"},
            "shared_encode_ms": 1e3 * t_encode_shared, "shared_model_python_ms": 1e3 * t_flatten_shared,
            "e2e_with_encode": {"value": B / ((ms + 1e3 * t_encode_shared) * 1e-3), "unit": "simulations/s",
                                "ms_per_step": ms + 1e3 * t_encode_shared,
                                "note": "the step plus ONE b200s_encode of the batch's shared model (all nodes, every candidate's pods) -- what the shim does per batch; building that model from the cluster model in Python (shared_model_python_ms) stands in for the Go flatten and is not included"},
            "commands_all_ranks": {"delete": deleted_all} if world > 1 else None,
            "commands": {"delete": int((cmds["decision"] == capi.CMD_DELETE).sum()),
                         "replace": int((cmds["decision"] == capi.CMD_REPLACE).sum()),
                         "none": int((cmds["decision"] == capi.CMD_NONE).sum())},
            "nodes_before": E, "single_node_removable": int(removable), "batch": B,
            "multi_node_repack": {"pods": int(len(pods)), "candidates": int(len(cand)), "new_nodeclaims": int(rm.n_claims),
                                  "unschedulable": int(rm.n_unscheduled), "nodes_after": int(E - len(cand) + rm.n_claims)},
            "multi_node_search": {"candidates": len(cand100), "prefixes_in_one_batch": len(cand100) - 1,
                                  "removes_first": None if got is None else got.mid + 1,
                                  "decision": None if got is None else ["none", "delete", "replace"][got.decision],
                                  "prefix_models_numpy_s": search_t.get("models_s"),
                                  "b200s_encode_s": search_t.get("encode_s"),
                                  "library_calls_s": search_t.get("device_calls_s"),
                                  "as_one_shared_model": search_shared,
                                  "cpu_sequential_search_s": t_search_cpu, "cpu_probes": len(probed)},
            "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "simulations/s", "cores": 1, "kind": "port",
                             "sample": "32 of the same simulations (every 8th) on one host core", "evals_per_simulation": evals_cpu},
            "e2e": {"value": B / (ms * 1e-3), "unit": "simulations/s"},
            "clocks": clk.summary(), "parity": "32 simulations, their consolidation commands and the multi-node re-pack identical to the CPU oracle"}
    ctx.close()
    emit(line)


def main():
    # stdout carries exactly one JSON line.  Native libraries write to file descriptor 1 on their own (NCCL prints its
    # version banner there), so fd 1 is pointed at stderr for the whole run and the line goes out through a saved copy.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c4", "c5"],
                    help="auto: config[1] (one Solve) on one GPU, config[4] (row-sharded, strong scaling) on several")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--pods", type=int, default=0)
    ap.add_argument("--skus", type=int, default=0)
    ap.add_argument("--large", action="store_true", help="also time the dense stage on 1M pods x 2k SKUs")
    ap.add_argument("--no-p2p", dest="no_p2p", action="store_true",
                    help="multi-GPU: only the NCCL all-gather, not the peer-store export over CUDA-IPC mappings")
    ap.add_argument("--no-global-pack", dest="no_global_pack", action="store_true",
                    help="multi-GPU: skip the 1M-pod global pack after the gather (70 s on one CTA)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "auto":
        args.workload = "c2" if max(world, args.gpus) == 1 else "c5"
    if args.impl == "reference":
        (run_reference_c5 if args.workload == "c5" else run_reference)(args, rank)
        return
    if args.workload == "c5":
        run_c5(args, rank, world, local)
        return
    if args.workload == "c4":
        run_c4(args, rank, world, local)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from b200sched import capi, synth
    import parity

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA library is the product and there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    model = synth.c2()
    enc = capi.Encoded(model)
    ctx = capi.Context(local)
    cid = ctx.upload_catalog(enc)
    res = ctx.resident(cid, enc)
    P, S = enc.P, enc.S
    o, of, credited, t_solve_cpu, t_fit_cpu = oracle_counts(model)      # full outputs: the parity gate needs them
    dense_evals = of[3]
    nthreads = host_threads()
    t_solve_mt = t_fit_mt = 0.0
    if rank == 0:  # the timed baseline runs alone (the other ranks wait at the barrier below): dense pass on all host threads
        _, _, _, t_solve_mt, t_fit_mt = oracle_counts(model, nthreads)

    # parity gate before any timing
    res.solve()
    raw = res.fetch_solve()
    parity.assert_same(parity.canon_device(enc, raw), parity.canon_oracle(model, o), "bench parity")
    launches_solve = res.launches()
    res.fit_matrix()
    launches_fit = res.launches()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    with ClockSampler(local) as clk:
        ms_solve = time_device(ctx, res.solve, args.steps, args.warmup)
        ms_fit, ms_dense_k = time_device(ctx, res.fit_matrix, args.steps, args.warmup, inner=res.dense_kernel_ms)
        for _ in range(args.warmup):
            ctx.solve(cid, enc)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.solve(cid, enc)  # host buffers in, host results out, through the public C-ABI call
        ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
        # the same with the host flattening inside the timed region: semantic model -> b200s_encode (dictionary coding
        # of the pods, requirement rows, class tables) -> b200s_solve -> results; the catalog stays uploaded (its key
        # layout does not change between these batches: layout_hash)
        t_enc = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t1 = time.perf_counter()
            enc_i = capi.Encoded(model)
            t_enc += time.perf_counter() - t1
            assert enc_i.layout_hash == enc.layout_hash
            ctx.solve(cid, enc_i)
        ms_e2e_enc = 1e3 * (time.perf_counter() - t0) / args.steps
        ms_encode = 1e3 * t_enc / args.steps
    # one extra, untimed Solve with the register-resident pack's cycle counters on (its PROF form: ~5 % slower)
    os.environ["B200S_REG_PROFILE"] = "1"
    res.solve()
    prof = res.reg_pack_profile()
    del os.environ["B200S_REG_PROFILE"]
    barrier()
    if world > 1:
        t = torch.tensor([ms_solve, ms_fit, ms_e2e, ms_e2e_enc], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_solve, ms_fit, ms_e2e, ms_e2e_enc = [float(x) for x in t.tolist()]

    peak, peak_src = peaks()
    ab = dense_bytes(enc)
    h2d = enc.P * (4 + 4 + 8 + 16) + enc.NS * (enc.K * 8 + 16 + enc.NB * 8) + enc.NQ * enc.R * 8
    d2h = enc.P * (12 + enc.T) + raw.n_claims * (8 + enc.R * 8 + enc.W * 8 + enc.K * 8 + 8 + enc.NB * 16 + 16)
    if rank != 0:
        t_solve_mt = t_fit_mt = float("nan")  # only rank 0 times (and reports) the CPU baseline
    cpu_ms = 1e3 * (t_solve_mt + t_fit_mt)
    flat_ms = None
    if rank == 0:  # the device algorithm itself (class tables, bit rows, threshold rule) on one host thread
        parity.emul_solve(enc)
        t0 = time.perf_counter()
        flat = parity.emul_solve(enc)
        flat_ms = 1e3 * (time.perf_counter() - t0)
        assert flat.n_claims == raw.n_claims
    steps_total = (args.steps + args.warmup)
    line = {
        "metric": METRIC, "value": world * credited / (ms_solve * 1e-3), "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_solve, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64/u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pods": P, "skus": S},
        "workload_detail": {"signatures": enc.NS, "request_classes": enc.NQ,
                            "l2": "flushed between timed iterations (256 MiB memset)",
                            "multi_gpu": "with --workload c2 under torchrun: one independent Solve per rank (replicas); the "
                                         "default under torchrun is the row-sharded config[4] workload (--workload c5)"},
        "nodeclaims_per_s": world * raw.n_claims / (ms_solve * 1e-3), "nodeclaims": int(raw.n_claims),
        "evals_per_step": int(credited), "dense_evals_per_step": int(dense_evals),
        "step_breakdown": {"dense_stage_ms": ms_fit, "dense_stage_share": ms_fit / ms_solve,
                           "pack_and_rest_ms": ms_solve - ms_fit,
                           "note": "the FFD pack is one sequential chain per Solve (latency-bound, one CTA); "
                                   "SM cycles per pod step of claim warp 0 (one extra Solve with the kernel's counters on, ~5 % slower): "
                                   + json.dumps({k: round(v / max(prof['steps'], 1)) for k, v in prof.items()
                                                 if k in ('claim_eval', 'claim_barrier', 'claim_post', 'claim_meet', 'claim_commit', 'order_warp_sort')})
                                   + f"; {prof['steps']} steps, {prof['meets']} of them met the order warp, {prof['full_sorts']} full pdqsorts"},
        "roofline": {"bound": "hbm", "kernel": "fit_dense_vec_kernel (dense pod x SKU stage) on config[1]",
                     "achieved": ab / (ms_dense_k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": ab / (ms_dense_k * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes": ab,
                     "kernel_ms": ms_dense_k, "stage_ms": ms_fit, "evals_per_s": dense_evals / (ms_dense_k * 1e-3),
                     "peak_source": peak_src,
                     "note": "config[1] moves 0.8 MB per launch (a ~5 us kernel: launch-latency bound by construction); "
                             "the same kernel at configs[4]'s size is in roofline_at_scale; the step's time is the "
                             "sequential pack (step_breakdown), which is latency-bound, not HBM-bound"},
        "cpu_baseline": {"value": credited / (cpu_ms * 1e-3), "unit": "evals/s", "cores": nthreads, "kind": "port",
                         "sample": "the whole workload once: one Solve on one thread (a sequential chain; the "
                                   "reference's in-Solve workers default to one) + one dense pass split over all "
                                   "host threads", "ms": cpu_ms,
                         "solve_ms": 1e3 * t_solve_mt, "dense_ms": 1e3 * t_fit_mt, "dense_ms_one_thread": 1e3 * t_fit_cpu,
                         "nodeclaims_per_s_solve_only": o.n_claims / t_solve_mt,
                         "flat_port_ms": flat_ms,
                         "flat_port_note": "flat_port_ms is the whole step (dense stage + pack) of this library's own "
                                           "algorithm compiled for the host (tests/native/emul.cpp, g++ -O2, one "
                                           "thread): the best-effort CPU form, not the reference's shape",
                         "note": "solve_ms is the reference-shaped Solve alone (fresh-claim rows evaluated lazily, "
                                 "as scheduler.go does); the metric also credits the full dense pass, which the "
                                 "reference never runs", "host_cores": os.cpu_count()},
        "e2e": {"value": world * credited / (ms_e2e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "e2e_with_encode": {"value": world * credited / (ms_e2e_enc * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e_enc,
                            "encode_ms": ms_encode,
                            "note": "b200s_encode of the semantic model (what the cgo shim hands over) + b200s_solve with "
                                    "host buffers + results, per step; the catalog upload is not repeated (same layout_hash)"},
        "config1_create_select": config1_latency(),
        "gpu_launches": int(steps_total * (launches_solve + launches_fit) + steps_total * launches_solve),
        "launches_per_solve": int(launches_solve), "launches_per_dense_stage": int(launches_fit),
        "clocks": clk.summary(),
        "parity": "bit-identical to the CPU oracle (SHA-256 " + parity.digest(parity.canon_device(enc, raw))[:16] + ")",
    }
    res.close()
    if rank == 0 and (world == 1 or args.large):
        # the HBM-bound kernel at the sizes where it is HBM-relevant (single GPU; the model build dominates the time)
        big = dense_leg(ctx, capi, synth, "c3", 1_000_000, 2_000, 10, 3, peak)
        tr = os.path.join(ROOT, "profiles", "dense_traffic.json")  # dram bytes of one launch from the ncu --set full capture
        traffic = json.load(open(tr)).get("dram_bytes_per_launch") if os.path.exists(tr) else None
        at_scale = {"bound": "hbm", "achieved": big["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": big["frac"], "traffic": traffic, **big, "peak_source": peak_src}
        # SURVEY section 8(d), caveat H3: config[1] fits in L2 and moves 0.8 MB per launch, the fraction is claimed where
        # the kernel is HBM-relevant.  `roofline` is therefore the dense kernel at configs[4]'s size on this GPU, measured
        # in this run; the same kernel inside the config[1] step stays visible as `roofline_config1`.
        line["roofline_config1"] = line["roofline"]
        line["roofline"] = dict(at_scale, note="fit_dense_vec_kernel on 1M pods x 2k SKUs (configs[4]'s size, one GPU), "
                                "timed by the library's event pair around that launch in this run; config[1] itself moves "
                                "0.8 MB per launch (roofline_config1) and its step time is the sequential pack "
                                "(step_breakdown)")
        line["roofline_at_scale"] = at_scale
        line["roofline_large"] = {"c3": dense_leg(ctx, capi, synth, "c3", 100_000, 1_000, 10, 3, peak), "c5": big,
                                  "c5_deployments_of_32": dense_leg(ctx, capi, synth, "c3", 1_000_000, 2_000, 10, 3, peak, run=32)}
    ctx.close()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

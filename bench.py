#!/usr/bin/env python
"""bench.py — pod x SKU fit evaluations/s and NodeClaims scheduled/s of the B200 path, next to the CPU oracle.

One step = one pass of the hot path over BASELINE.json configs[1] (10 000 synthetic pending pods x 500 Azure
SKUs): dense fit matrix (stage A) + FFD order + bin-pack + cheapest SKU per NodeClaim, i.e. one b200s_solve.
  value    evaluations/s, device-timed (CUDA events on the library's stream), inputs resident in HBM
  e2e      the same through the host-buffer C-ABI call (b200s_solve): H2D of the pod/node SoA and D2H of the
           result inside the timed region
  roofline the dense fit-matrix kernel: algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
Evaluations per step = P x S-after-prefilter instance types the dense stage visits + the instance types the
reference's in-flight CanAdd calls visit (the oracle's count, credited — not the device's own larger count).

`--impl reference` times the CPU oracle (reference-shaped restatement of the Go scheduler; the reference itself
is Go and cannot be built here) on the same config, same metric.  Under torchrun every rank solves its own
shard-independent replica (the pack is sequential per Solve: "replicas only"; the dense stage shards by pod
rows and needs no collective) and rank 0 prints the max-over-ranks time.
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.rows = []
        self.stop = False
        self.index = index
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
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def algorithmic_bytes_fit(enc) -> int:
    """Compulsory traffic of one dense fit-matrix launch with this repo's layout (DESIGN.md "Algorithmic bytes"):
    every pod row read once, every output written once, catalog tables read once."""
    pod_row = 4 + 4                                                # signature id + request-class id
    out_row = enc.W * 8 + 4 + enc.T                                # mask, first feasible template, per-template codes
    tables = (enc.NS * 2 + enc.NQ) * enc.T * enc.W * 8 + enc.NS * enc.T  # class rows C, O, F + status, read once
    return enc.P * (pod_row + out_row) + tables


def run_reference(args, rank, world):
    import oracle_binding as ob
    from b200sched import synth
    if rank != 0:
        return
    model = synth.c2()
    P, S = model.n_pods, model.n_skus
    ob.lib()
    times = []
    evals = None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o = ob.solve(model)
        f = ob.fit_matrix(model)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
        evals = f[3] + o.counters["fit_evals"] - o.counters["newclaim_evals"]  # dense matrix + in-flight CanAdd visits
        nclaims = o.n_claims
    ms = 1e3 * sum(times) / len(times)
    value = evals / (ms / 1e3)
    line = {"metric": "pod x SKU fit evaluations/s", "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64/u64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "config[1]: 10k pending pods x 500 Azure SKUs, 1 NodePool; dense fit matrix + full Solve",
                       "pods": P, "skus": S},
            "nodeclaims_per_s": nclaims / (ms / 1e3),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": 1, "kind": "port",
                             "sample": "the whole config (one dense pass + one Solve per step), single thread as the reference's default one worker"},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    from b200sched import capi, synth
    import oracle_binding as ob
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

    # oracle counts for crediting (rank 0 only times it; every rank needs the count)
    t0 = time.perf_counter()
    o = ob.solve(model)
    t_solve_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    of = ob.fit_matrix(model)
    t_fit_cpu = time.perf_counter() - t0
    dense_evals = of[3]
    # a step evaluates the dense pod x SKU matrix (which subsumes addToNewNodeClaim's template walks) plus the
    # instance types the reference's in-flight CanAdd calls visit (the oracle's count, credited)
    credited = dense_evals + o.counters["fit_evals"] - o.counters["newclaim_evals"]

    # parity gate before any timing
    res.solve()
    raw = res.fetch_solve()
    parity.assert_same(parity.canon_device(enc, raw), parity.canon_oracle(model, o), "bench parity")
    launches = res.launches()

    ev = [ctx.event() for _ in range(4)]

    def timed_device(fn, steps, warmup):
        for _ in range(warmup):
            ctx.flush_l2()
            fn()
        ctx.synchronize()
        total = 0.0
        for _ in range(steps):
            ctx.flush_l2()          # L2 flush between timed iterations (inputs are smaller than L2)
            ctx.record(ev[0])
            fn()
            ctx.record(ev[1])
            total += ctx.elapsed_ms(ev[0], ev[1])
        return total / steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    with ClockSampler(local) as clk:
        ms_solve = timed_device(res.solve, args.steps, args.warmup)
        ms_fit = timed_device(res.fit_matrix, args.steps, args.warmup)
        # e2e: host buffers in, host results out, through the public C-ABI call
        for _ in range(args.warmup):
            ctx.solve(cid, enc)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.solve(cid, enc)
        ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
    barrier()
    if world > 1:
        t = torch.tensor([ms_solve, ms_fit, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_solve, ms_fit, ms_e2e = [float(x) for x in t.tolist()]

    peak, peak_src = peaks()
    ab = algorithmic_bytes_fit(enc)
    achieved = ab / (ms_fit * 1e-3) / 1e9
    h2d = enc.P * (4 + 4 + 8 + 16) + enc.NS * (enc.K * 8 + 16 + enc.NB * 8) + enc.NQ * enc.R * 8
    d2h = enc.P * (12 + enc.T) + raw.n_claims * (8 + enc.R * 8 + enc.W * 8 + enc.K * 8 + 8 + enc.NB * 16 + 16)

    # CPU baseline: a bounded sample — one dense pass + one Solve of this config, already timed above on this host
    cpu_ms = 1e3 * (t_solve_cpu + t_fit_cpu)
    line = {
        "metric": "pod x SKU fit evaluations/s", "value": world * credited / (ms_solve * 1e-3), "unit": "evals/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_solve,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/u64", "data": "synthetic",
        "config": {"workload": "config[1]: 10k pending pods x 500 Azure SKUs, 1 NodePool; dense fit matrix + full Solve",
                   "pods": P, "skus": S, "l2": "flushed between timed iterations (256 MiB memset)",
                   "multi_gpu": "replicas only: one independent Solve per rank, no data-path collective"},
        "nodeclaims_per_s": world * raw.n_claims / (ms_solve * 1e-3),
        "nodeclaims": int(raw.n_claims), "evals_per_step": int(credited), "dense_evals_per_step": int(dense_evals),
        "fit_matrix": {"ms": ms_fit, "evals_per_s": dense_evals / (ms_fit * 1e-3)},
        "roofline": {"bound": "hbm", "kernel": "fit_matrix_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "algorithmic_bytes": ab, "peak_source": peak_src,
                     "note": "config[1] moves ~2 MB per launch: launch-latency bound; see profiles/ for the 100k x 1k figure"},
        "cpu_baseline": {"value": credited / (cpu_ms * 1e-3), "unit": "evals/s", "cores": 1, "kind": "port",
                         "sample": "one dense pass + one Solve of the same config on one host core (reference default: one worker)",
                         "ms": cpu_ms, "host_cores": os.cpu_count()},
        "e2e": {"value": world * credited / (ms_e2e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches) * args.steps * 2 - args.steps * (launches - 2),
        "launches_per_solve": int(launches),
        "clocks": clk.summary(),
        "parity": "bit-identical to the CPU oracle (SHA-256 " + parity.digest(parity.canon_device(enc, raw))[:16] + ")",
    }
    res.close()
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

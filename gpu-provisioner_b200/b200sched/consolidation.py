"""Host side of multi-node consolidation over the C-ABI.

`MultiNodeConsolidation.firstNConsolidationOption` (K/controllers/disruption/multinodeconsolidation.go:116-170) binary-
searches the longest prefix of the cost-ordered candidates that one command can remove, paying one scheduling
simulation per probe, sequentially.  A probe's outcome depends only on its prefix, so every prefix the search could
visit (at most 100) goes to the device as ONE batch -- one CTA per simulation (`b200s_solve_batch`), one warp per
`computeConsolidation` (`b200s_consolidation_commands`), a second launch of the same kernel for
`filterOutSameInstanceType` (:172-224) -- and the search itself is replayed on the answers.  Same probes, same result,
no dependency chain of simulations.

Nothing here computes on the CPU what the device path is for: the module gathers inputs, launches, and replays the
search.  It needs the CUDA library like every other product path.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from . import model as M

MAX_FLOAT64 = 1.7976931348623157e308
MAX_PARALLEL = 100  # multinodeconsolidation.go:87


class CandidatePriceError(RuntimeError):
    """getCandidatePrices' "unable to determine offering" (consolidation.go:323), raised where the reference's search
    would have hit it: on a visited probe whose simulation scheduled everything into exactly one new NodeClaim"""


def node_facts(model: M.Model, nodes: Sequence[int]) -> List[Tuple[int, float, bool, bool]]:
    """(instance type index, cheapest offering the node's labels admit or -1, is a spot node, is a reserved-capacity
    node) per existing node: what
    getCandidatePrices (consolidation.go:310-328) and filterOutSameInstanceType (multinodeconsolidation.go:177-193) read
    off a candidate.  Offerings are not filtered by availability there."""
    a = model.arrays
    k_it, k_ct = model.key_names.index(M.KEY_INSTANCE_TYPE), model.key_names.index(M.KEY_CAPACITY_TYPE)
    sku_of = {n: i for i, n in enumerate(model.sku_names)}
    out = []
    for e in nodes:
        lab = {}
        for i in range(a["node_req_off"][e], a["node_req_off"][e + 1]):
            lab[int(a["reqs"][i]["key"])] = int(a["req_values"][a["reqs"][i]["val_off"]])
        s = sku_of[model.val_names[lab[k_it]]]
        best = -1.0
        for o in range(a["sku_off_off"][s], a["sku_off_off"][s + 1]):
            ok = True
            for i in range(a["off_req_off"][o], a["off_req_off"][o + 1]):
                key, v = int(a["reqs"][i]["key"]), int(a["req_values"][a["reqs"][i]["val_off"]])
                ok = ok and (key not in lab or lab[key] == v)
            if ok and (best < 0 or a["off_price"][o] < best):
                best = float(a["off_price"][o])
        out.append((s, best, model.val_names[lab[k_ct]] == "spot", model.val_names[lab[k_ct]] == "reserved"))
    return out


def prefix_simulation(model: M.Model, node_of_pod: np.ndarray, candidates: Sequence[int], mid: int) -> M.Model:
    """SimulateScheduling's input for candidates[0..mid] (helpers.go:52-143): their pods against every other node."""
    gone = np.asarray(candidates[:mid + 1], np.int64)
    pods = np.nonzero(np.isin(node_of_pod, gone))[0]
    keep = np.setdiff1d(np.arange(int(model.c.n_nodes)), gone)
    return M.subset(model, pods, keep)


def prefix_price(facts: Sequence[Tuple[int, float, bool, bool]], mid: int) -> Tuple[float, bool]:
    """getCandidatePrices (consolidation.go:310-328): the candidates' prices added up in candidate order -> (total,
    error).  The first candidate without a compatible offering ends the walk: a reserved-capacity node makes the total
    0.0 (free by definition, :317-320), any other is the "unable to determine offering" error (:323)."""
    price = 0.0
    for _, p, _, reserved in facts[:mid + 1]:
        if p < 0:
            return 0.0, not reserved
        price += p
    return price, False


@dataclass
class Probe:
    mid: int                 # the probe removes candidates[0..mid]
    decision: int            # capi.CMD_*
    reason: int              # capi.WHY_*
    valid: bool              # validDecision (multinodeconsolidation.go:139-150)
    n_options: int
    mask: np.ndarray         # replacement instance types after filterOutSameInstanceType
    spot_required: bool
    result: capi.SolveRaw
    price_error: bool = False  # getCandidatePrices failed on this prefix AND the reference would have called it


def shared_prefix_model(model: M.Model, node_of_pod: np.ndarray, candidates: Sequence[int], mids: Sequence[int]):
    """Every prefix of `mids` as ONE model for b200s_solve_batch_shared: all nodes of the cluster, and for each prefix, one
    after the other, the pods of candidates[0..mid] (a pod of an early candidate appears once per prefix that holds it; in
    the order prefix_simulation gives them).  Returns (model, pod_off, removed): prefix i = pods pod_off[i]..pod_off[i+1]
    against every node but the model nodes removed[i]."""
    cand = np.asarray(candidates, np.int64)
    pods, off, removed = [], [0], []
    for mid in mids:
        gone = cand[:mid + 1]
        pods.append(np.nonzero(np.isin(node_of_pod, gone))[0])
        off.append(off[-1] + len(pods[-1]))
        removed.append([int(e) for e in gone])
    allp = np.concatenate(pods) if pods else np.zeros(0, np.int64)
    return M.subset(model, allp, np.arange(int(model.c.n_nodes))), np.asarray(off, np.int32), removed


def multi_node_probes(ctx: capi.Context, cid: int, model: M.Model, node_of_pod: np.ndarray, candidates: Sequence[int],
                      mids: Sequence[int], ct_values: Tuple[int, int, int, int], spot_to_spot: bool = False,
                      timing: Optional[dict] = None, shared: bool = False) -> List[Probe]:
    """computeConsolidation + filterOutSameInstanceType for the prefixes `mids`, as one batch.
    ct_values = (model key of the capacity type, model value ids of spot, on-demand, reserved or -1).
    timing, when given, receives the wall seconds of building the prefix models ("models_s"), of b200s_encode
    ("encode_s") and of the device calls ("device_calls_s": host buffers in, results out).
    shared: hand the prefixes over as one model (shared_prefix_model, ONE b200s_encode, the cluster's nodes once:
    b200s_solve_batch_shared) instead of one model per prefix; existing-node indices in Probe.result then count rows of
    the whole cluster's node set."""
    import time
    mids = list(mids)
    if not mids:
        return []
    t0 = time.perf_counter()
    facts = node_facts(model, candidates)
    if shared:
        one, pod_off, removed = shared_prefix_model(model, node_of_pod, candidates, mids)
        t_sims = time.perf_counter()
        enc = capi.Encoded(one)
        t1 = time.perf_counter()
        uploaded = ctx.catalog_layout(cid)
        own = ctx.upload_catalog(enc) if (uploaded is not None and enc.layout_hash != uploaded) else None
        try:
            use = cid if own is None else own
            raws = ctx.solve_batch_shared(use, enc, pod_off, [enc.node_rows(r) for r in removed])
            out = _commands(ctx, use, enc, raws, mids, facts, ct_values, spot_to_spot)
        finally:
            if own is not None:
                ctx.release_catalog(own)
        if timing is not None:
            timing["models_s"], timing["encode_s"] = t_sims - t0, t1 - t_sims
            timing["device_calls_s"] = time.perf_counter() - t1
        return out
    sims = [prefix_simulation(model, node_of_pod, candidates, mid) for mid in mids]
    t_sims = time.perf_counter()
    encs = [capi.Encoded(m) for m in sims]
    t1 = time.perf_counter()
    # A prefix's pods may bring a label key, a Gt/Lt bound or a well-known flag the others do not have: b200s_encode then
    # chooses another key layout (layout_hash) and those probes cannot share the uploaded catalog.  Probes are grouped
    # by layout; the group of the caller's catalog runs as one batch on it, every other group on a catalog of its own.
    uploaded = ctx.catalog_layout(cid)
    groups = {}
    for i, e in enumerate(encs):
        groups.setdefault(e.layout_hash, []).append(i)
    out: List[Optional[Probe]] = [None] * len(mids)
    for h, idx in groups.items():
        own = None
        if uploaded is not None and h != uploaded:
            own = ctx.upload_catalog(encs[idx[0]])
        try:
            probes = _probe_group(ctx, cid if own is None else own, [encs[i] for i in idx], [mids[i] for i in idx], facts,
                                  ct_values, spot_to_spot)
        finally:
            if own is not None:
                ctx.release_catalog(own)
        for i, pr in zip(idx, probes):
            out[i] = pr
    if timing is not None:
        timing["models_s"] = t_sims - t0   # numpy row selection of the synthetic model (the cluster-state feed's job)
        timing["encode_s"] = t1 - t_sims   # b200s_encode of every prefix
        timing["device_calls_s"] = time.perf_counter() - t1
    return out


def _probe_group(ctx: capi.Context, cid: int, encs: List[capi.Encoded], mids: List[int], facts, ct_values,
                 spot_to_spot: bool) -> List[Probe]:
    """the probes of one key layout, as one batch on catalog `cid`"""
    return _commands(ctx, cid, encs[0], ctx.solve_batch(cid, encs), mids, facts, ct_values, spot_to_spot)


def _commands(ctx: capi.Context, cid: int, enc0: capi.Encoded, raws, mids: List[int], facts, ct_values,
              spot_to_spot: bool) -> List[Probe]:
    """computeConsolidation (+ filterOutSameInstanceType) on finished simulations `raws` of the prefixes `mids`"""
    key, v_spot, v_od, v_res = ct_values
    bit_spot, bit_od, bit_res = (enc0.value_bit(key, v) if v >= 0 else -1 for v in (v_spot, v_od, v_res))
    priced = [prefix_price(facts, mid) for mid in mids]
    prices = [p for p, _ in priced]
    spots = [all(f[2] for f in facts[:mid + 1]) for mid in mids]
    counts = [mid + 1 for mid in mids]
    cmds = ctx.consolidation_commands(cid, enc0, raws, prices, spots, counts, bit_spot, bit_od, bit_res, spot_to_spot)
    valid = cmds["decision"] == capi.CMD_DELETE
    masks, n_opts = cmds["mask"].copy(), cmds["n_options"].copy()
    # filterOutSameInstanceType: the price filter once more, against the cheapest removed node of a type that is also
    # among the replacement's options, under the requirements the command left on the NodeClaim
    again = [i for i in range(len(mids)) if cmds["decision"][i] == capi.CMD_REPLACE]
    if again:
        k_ct = enc0.key_model.index(key)
        rows, limits = [], []
        for i in again:
            existing, cheapest = set(), {}
            for s, p, _, _ in facts[:mids[i] + 1]:
                existing.add(s)
                if p >= 0:  # candidates without a compatible offering name a type but price nothing (:185-187)
                    cheapest[s] = min(cheapest.get(s, MAX_FLOAT64), p)
            limit = MAX_FLOAT64
            for s in existing:
                if (int(masks[i, s // 64]) >> (s % 64)) & 1:
                    limit = min(limit, cheapest.get(s, 0.0))  # Go's map zero value for an unpriced type (:203)
            r = raws[i]
            req_mask, req_flags = r.claim_req_mask[:1].copy(), r.claim_req_flags[:1].copy()
            if cmds["spot_required"][i]:  # Requirements.Add(capacity-type In [spot]) happened before (:218-221, :240)
                req_mask[0, k_ct] = np.uint64(1) << np.uint64(bit_spot)
                nib = np.uint64(0xF) << np.uint64(4 * k_ct)
                req_flags[0] = (req_flags[0] & ~nib) | (np.uint64(capi.RF_DEFINED) << np.uint64(4 * k_ct))
            rows.append(SimpleNamespace(n_claims=1, n_unscheduled=0, claim_mask=masks[i:i + 1], claim_req_mask=req_mask,
                                        claim_req_flags=req_flags, claim_req_gt=r.claim_req_gt[:1],
                                        claim_req_lt=r.claim_req_lt[:1]))
            limits.append(limit)
        second = ctx.consolidation_commands(cid, enc0, rows, limits, [False] * len(rows), [2] * len(rows), bit_spot, bit_od,
                                            bit_res, spot_to_spot)
        for j, i in enumerate(again):
            valid[i] = second["decision"][j] == capi.CMD_REPLACE  # something is left (:146)
            masks[i], n_opts[i] = second["mask"][j], second["n_options"][j]
    # the reference reaches getCandidatePrices only when everything was scheduled into exactly one new NodeClaim
    # (consolidation.go:146-168): only there does a pricing error exist
    err = [priced[i][1] and raws[i].n_unscheduled == 0 and raws[i].n_claims == 1 for i in range(len(mids))]
    return [Probe(mid=mids[i], decision=int(cmds["decision"][i]), reason=int(cmds["reason"][i]),
                  valid=bool(valid[i]) and not err[i], n_options=int(n_opts[i]), mask=masks[i],
                  spot_required=bool(cmds["spot_required"][i]), result=raws[i], price_error=bool(err[i]))
            for i in range(len(mids))]


def search_bounds(n_candidates: int, max_parallel: int = MAX_PARALLEL) -> Tuple[int, int]:
    """min / max of firstNConsolidationOption (multinodeconsolidation.go:122-125)"""
    hi = max_parallel
    if n_candidates <= hi:
        hi = n_candidates - 1
    return 1, hi


def replay_search(lo: int, hi: int, valid_of) -> Optional[int]:
    """the binary search of multinodeconsolidation.go:131-168 on known answers; returns the mid of lastSavedCommand"""
    last = None
    while lo <= hi:
        mid = (lo + hi) // 2
        if valid_of(mid):
            last = mid
            lo = mid + 1
        else:
            hi = mid - 1
    return last


def first_n_consolidation_option(ctx: capi.Context, cid: int, model: M.Model, node_of_pod: np.ndarray,
                                 candidates: Sequence[int], ct_values: Tuple[int, int, int, int],
                                 max_parallel: int = MAX_PARALLEL, spot_to_spot: bool = False,
                                 timing: Optional[dict] = None, shared: bool = False) -> Optional[Probe]:
    """firstNConsolidationOption: candidates are already in disruption-cost order and capped by the budgets."""
    if len(candidates) < 2:
        return None
    lo, hi = search_bounds(len(candidates), min(max_parallel, len(candidates)))
    probes = multi_node_probes(ctx, cid, model, node_of_pod, candidates, range(lo, hi + 1), ct_values, spot_to_spot, timing,
                               shared=shared)
    by_mid = {p.mid: p for p in probes}
    def valid_of(mid):
        if by_mid[mid].price_error:  # computeConsolidation returned an error: the search ends with it (:147)
            raise CandidatePriceError(f"unable to determine offering for a candidate of the prefix 0..{mid}")
        return by_mid[mid].valid
    best = replay_search(lo, hi, valid_of)
    return None if best is None else by_mid[best]

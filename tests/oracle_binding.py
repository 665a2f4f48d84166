"""ctypes binding of the CPU oracle (oracle/karp_oracle.cpp).  Test infrastructure only: nothing in
the product package imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libkarp_oracle.so")
MAX_TEMPLATES = 16
_P = C.POINTER


class OracleOpts(C.Structure):
    _fields_ = [("stable_sort", C.c_int32), ("reserved", C.c_int32 * 7)]


class OracleResult(C.Structure):
    _fields_ = [
        ("cap_claims", C.c_int32), ("n_claims", C.c_int32), ("n_unscheduled", C.c_int32), ("n_templates", C.c_int32),
        ("template_pool", C.c_int32 * MAX_TEMPLATES),
        ("fit_evals", C.c_int64), ("prefilter_evals", C.c_int64), ("canadd_claim", C.c_int64),
        ("canadd_node", C.c_int64), ("new_claim_attempts", C.c_int64), ("sorts", C.c_int64),
        ("sorts_unsorted", C.c_int64), ("price_ties", C.c_int64), ("newclaim_evals", C.c_int64),
        ("value_overflow", C.c_int32), ("_pad", C.c_int32),
        ("queue_order", _P(C.c_int32)), ("node_order", _P(C.c_int32)), ("node_npods", _P(C.c_int32)),
        ("pod_place_kind", _P(C.c_int32)), ("pod_place_index", _P(C.c_int32)), ("pod_seq", _P(C.c_int32)),
        ("pod_fit_code", _P(C.c_uint8)),
        ("claim_template", _P(C.c_int32)), ("claim_npods", _P(C.c_int32)), ("claim_requests", _P(C.c_int64)),
        ("claim_mask", _P(C.c_uint64)), ("claim_n_options", _P(C.c_int32)), ("claim_req_mask", _P(C.c_uint64)),
        ("claim_req_flags", _P(C.c_uint8)), ("claim_req_gt", _P(C.c_int64)), ("claim_req_lt", _P(C.c_int64)),
        ("claim_sku", _P(C.c_int32)), ("claim_price", _P(C.c_uint64)),
    ]


_lib = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_solve.restype = C.c_int
        _lib.oracle_consolidate.restype = C.c_int
        _lib.oracle_solve_ordered.restype = C.c_int
        _lib.oracle_fit_matrix.restype = C.c_int
        _lib.oracle_go_sort_perm.restype = C.c_int
        _lib.oracle_req_algebra.restype = C.c_int
        _lib.oracle_tolerates.restype = C.c_int
        _lib.oracle_create_select.restype = C.c_int
        i64p, u8p = _P(C.c_int64), _P(C.c_uint8)
        _lib.oracle_pod_requests.argtypes = [C.c_int32, C.c_int32, C.c_int32, i64p, u8p, i64p, u8p, C.c_int32, i64p, u8p, i64p,
                                             u8p, u8p, i64p, i64p]
        _lib.oracle_allocatable.argtypes = [C.c_int32, C.c_int32, i64p, u8p, i64p, i64p, i64p, u8p, i64p]
        _lib.oracle_create_select.argtypes = [C.c_char_p, C.c_int32, _P(C.c_char_p), _P(C.c_double), C.c_int64,
                                              _P(C.c_int32), _P(C.c_int32), _P(C.c_int32)]
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(_P(ct))


@dataclass
class SolveOut:
    """Canonical, model-level view of a Solve result (the parity unit of SURVEY.md section 8d)."""
    n_claims: int
    n_unscheduled: int
    n_templates: int
    template_pool: np.ndarray
    queue_order: np.ndarray
    node_order: np.ndarray
    node_npods: np.ndarray
    pod_place_kind: np.ndarray
    pod_place_index: np.ndarray
    pod_seq: np.ndarray
    pod_fit_code: np.ndarray       # [P, T]
    claim_template: np.ndarray
    claim_npods: np.ndarray
    claim_requests: np.ndarray     # [N, R]
    claim_mask: np.ndarray         # [N, W]
    claim_n_options: np.ndarray
    claim_req_mask: np.ndarray     # [N, n_keys] model keys
    claim_req_flags: np.ndarray    # [N, n_keys]
    claim_req_gt: np.ndarray
    claim_req_lt: np.ndarray
    claim_sku: np.ndarray
    claim_price: np.ndarray
    counters: Dict[str, int]


class ConsolidationIn(C.Structure):
    _fields_ = [("n_candidates", C.c_int32), ("all_spot", C.c_int32), ("spot_to_spot", C.c_int32),
                ("val_spot", C.c_int32), ("val_on_demand", C.c_int32), ("val_reserved", C.c_int32),
                ("candidate_price", C.c_double), ("filter_same_type", C.c_int32), ("n_cand_rows", C.c_int32),
                ("cand_sku", _P(C.c_int32)), ("cand_price", _P(C.c_double))]


class ConsolidationOut(C.Structure):
    _fields_ = [("decision", C.c_int32), ("reason", C.c_int32), ("n_options", C.c_int32), ("spot_required", C.c_int32),
                ("price_tie", C.c_int32), ("valid", C.c_int32), ("head", C.c_int32 * 15), ("mask", _P(C.c_uint64)),
                ("order", _P(C.c_int32))]


def consolidate(model, candidate_price: float, all_spot: bool, n_candidates: int, val_spot: int, val_on_demand: int,
                val_reserved: int = -1, spot_to_spot: bool = False, same_type_rows=None):
    """Solve + computeConsolidation (consolidation.go:133-307) on one simulation; returns (SolveOut, command dict).
    same_type_rows = [(instance type index, price or -1)] per candidate switches on the multi-node probe's
    filterOutSameInstanceType and validDecision (multinodeconsolidation.go:139-150, :172-224)."""
    cin = ConsolidationIn(n_candidates, int(all_spot), int(spot_to_spot), val_spot, val_on_demand, val_reserved,
                          float(candidate_price))
    if same_type_rows is not None:
        sku = np.asarray([r[0] for r in same_type_rows], np.int32)
        prc = np.asarray([r[1] for r in same_type_rows], np.float64)
        cin.filter_same_type, cin.n_cand_rows = 1, len(sku)
        cin.cand_sku, cin.cand_price = _p(sku, C.c_int32), _p(prc, C.c_double)
    mask = np.zeros(model.n_words, np.uint64)
    order = np.full(max(model.n_skus, 1), -1, np.int32)
    cout = ConsolidationOut()
    cout.mask, cout.order = _p(mask, C.c_uint64), _p(order, C.c_int32)
    out = solve(model, _consolidation=(cin, cout))
    cmd = dict(decision=cout.decision, reason=cout.reason, n_options=cout.n_options, spot_required=cout.spot_required,
               price_tie=cout.price_tie, valid=cout.valid, head=np.array(cout.head[:], np.int32), mask=mask,
               order=order[:cout.n_options].copy())
    return out, cmd


def solve_ordered(model, max_types: int = 600):
    """Solve + ToNodeClaim's price-ordered instance-type list per new NodeClaim (nodeclaimtemplate.go:77-82);
    returns (SolveOut, order [N, max_types], count [N], tie [N])."""
    cap = max(model.n_pods, 1)
    order = np.full(cap * max_types, -1, np.int32)
    count = np.zeros(cap, np.int32)
    tie = np.zeros(cap, np.uint8)
    out = solve(model, _ordered=(max_types, order, count, tie))
    N = out.n_claims
    return out, order.reshape(cap, max_types)[:N], count[:N], tie[:N]


def solve(model, stable_sort: bool = False, cap_claims: Optional[int] = None, _consolidation=None, _ordered=None) -> SolveOut:
    L = lib()
    P, R, W, K, E = model.n_pods, model.n_res, model.n_words, len(model.key_names), int(model.c.n_nodes)
    cap = cap_claims if cap_claims is not None else max(P, 1)
    T = max(int(model.c.n_pools), 1)
    res = OracleResult()
    res.cap_claims = cap
    arr = dict(
        queue_order=np.zeros(P, np.int32), node_order=np.zeros(max(E, 1), np.int32),
        node_npods=np.zeros(max(E, 1), np.int32), pod_place_kind=np.zeros(P, np.int32),
        pod_place_index=np.zeros(P, np.int32), pod_seq=np.zeros(P, np.int32),
        pod_fit_code=np.zeros(P * T, np.uint8), claim_template=np.zeros(cap, np.int32),
        claim_npods=np.zeros(cap, np.int32), claim_requests=np.zeros(cap * R, np.int64),
        claim_mask=np.zeros(cap * W, np.uint64), claim_n_options=np.zeros(cap, np.int32),
        claim_req_mask=np.zeros(cap * K, np.uint64), claim_req_flags=np.zeros(cap * K, np.uint8),
        claim_req_gt=np.zeros(cap * K, np.int64), claim_req_lt=np.zeros(cap * K, np.int64),
        claim_sku=np.zeros(cap, np.int32), claim_price=np.zeros(cap, np.uint64))
    ctypes_of = {np.dtype(np.int32): C.c_int32, np.dtype(np.int64): C.c_int64, np.dtype(np.uint64): C.c_uint64,
                 np.dtype(np.uint8): C.c_uint8}
    for k, a in arr.items():
        setattr(res, k, _p(a, ctypes_of[a.dtype]))
    opts = OracleOpts()
    opts.stable_sort = 1 if stable_sort else 0
    if _ordered is not None:
        rc = L.oracle_solve_ordered(C.byref(model.c), C.byref(opts), C.byref(res), C.c_int32(_ordered[0]),
                                    _p(_ordered[1], C.c_int32), _p(_ordered[2], C.c_int32), _p(_ordered[3], C.c_uint8))
    elif _consolidation is None:
        rc = L.oracle_solve(C.byref(model.c), C.byref(opts), C.byref(res))
    else:
        rc = L.oracle_consolidate(C.byref(model.c), C.byref(opts), C.byref(res), C.byref(_consolidation[0]),
                                  C.byref(_consolidation[1]))
    if rc != 0:
        raise RuntimeError(f"oracle_solve failed: {rc}")
    N, Tn = res.n_claims, res.n_templates
    return SolveOut(
        n_claims=N, n_unscheduled=res.n_unscheduled, n_templates=Tn,
        template_pool=np.array(res.template_pool[:Tn], dtype=np.int32),
        queue_order=arr["queue_order"], node_order=arr["node_order"][:E], node_npods=arr["node_npods"][:E],
        pod_place_kind=arr["pod_place_kind"], pod_place_index=arr["pod_place_index"], pod_seq=arr["pod_seq"],
        pod_fit_code=arr["pod_fit_code"][:P * Tn].reshape(P, Tn) if Tn else np.zeros((P, 0), np.uint8),
        claim_template=arr["claim_template"][:N], claim_npods=arr["claim_npods"][:N],
        claim_requests=arr["claim_requests"][:N * R].reshape(N, R), claim_mask=arr["claim_mask"][:N * W].reshape(N, W),
        claim_n_options=arr["claim_n_options"][:N], claim_req_mask=arr["claim_req_mask"][:N * K].reshape(N, K),
        claim_req_flags=arr["claim_req_flags"][:N * K].reshape(N, K), claim_req_gt=arr["claim_req_gt"][:N * K].reshape(N, K),
        claim_req_lt=arr["claim_req_lt"][:N * K].reshape(N, K), claim_sku=arr["claim_sku"][:N],
        claim_price=arr["claim_price"][:N],
        counters=dict(fit_evals=res.fit_evals, prefilter_evals=res.prefilter_evals, canadd_claim=res.canadd_claim,
                      canadd_node=res.canadd_node, new_claim_attempts=res.new_claim_attempts, sorts=res.sorts,
                      sorts_unsorted=res.sorts_unsorted, price_ties=res.price_ties, newclaim_evals=res.newclaim_evals,
                      value_overflow=res.value_overflow))


def fit_matrix(model, pod_begin: int = 0, pod_end: Optional[int] = None):
    L = lib()
    pod_end = model.n_pods if pod_end is None else pod_end
    n, W = pod_end - pod_begin, model.n_words
    T = max(int(model.c.n_pools), 1)
    tmpl = np.zeros(n, np.int32)
    mask = np.zeros(n * W, np.uint64)
    code = np.zeros(n * T, np.uint8)
    evals = C.c_int64(0)
    nt = C.c_int32(0)
    rc = L.oracle_fit_matrix(C.byref(model.c), pod_begin, pod_end, _p(tmpl, C.c_int32), _p(mask, C.c_uint64),
                             _p(code, C.c_uint8), C.byref(evals), C.byref(nt))
    if rc != 0:
        raise RuntimeError(f"oracle_fit_matrix failed: {rc}")
    Tn = nt.value
    return tmpl, mask.reshape(n, W), code[:n * Tn].reshape(n, Tn), int(evals.value), Tn


def fit_matrix_threads(model, threads: int):
    """The dense pass split into contiguous pod ranges, one per host thread (the rows are independent; ctypes
    releases the GIL).  Returns the summed evaluation count; used by bench.py's reference arm only."""
    from concurrent.futures import ThreadPoolExecutor
    lib()
    P = model.n_pods
    threads = max(1, min(threads, P))
    bounds = [P * i // threads for i in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda i: fit_matrix(model, bounds[i], bounds[i + 1])[3], range(threads)))
    return int(sum(parts))


def go_sort_perm(keys: np.ndarray) -> np.ndarray:
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    perm = np.zeros(len(keys), np.int32)
    lib().oracle_go_sort_perm(_p(keys, C.c_int64), len(keys), _p(perm, C.c_int32))
    return perm

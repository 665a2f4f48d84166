"""Parity helpers: bring a device-level result (libb200sched or the host emulation) and an oracle
result to the same canonical, model-level form and compare them field by field.

The parity unit is SURVEY.md section 8d: per new NodeClaim in Results.NewNodeClaims order — pods (by
placement order), surviving instance types, merged requirements, accumulated requests, selected SKU
and its price — plus per-pod placement and error class.  ``digest`` is the SHA-256 of it.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
import sys
from typing import Dict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-provisioner_b200"))

from b200sched import capi  # noqa: E402

EMUL_SRC = os.path.join(ROOT, "tests", "native", "emul.cpp")
EMUL_LIB = os.path.join(ROOT, "tests", "native", "_build", "libemul.so")
_emul = None


def emul():
    """Host emulation of the device algorithm (tests/native/emul.cpp), built on demand."""
    global _emul
    if _emul is None:
        os.makedirs(os.path.dirname(EMUL_LIB), exist_ok=True)
        deps = [EMUL_SRC, os.path.join(ROOT, "gpu-provisioner_b200", "csrc", "b200_core.cuh"),
                os.path.join(ROOT, "include", "b200sched.h")]
        if not os.path.exists(EMUL_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMUL_LIB) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", EMUL_SRC, "-o", EMUL_LIB])
        _emul = C.CDLL(EMUL_LIB)
    return _emul


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def emul_fit_matrix(enc: capi.Encoded) -> capi.FitOut:
    P, W, T = enc.P, enc.W, enc.T
    tmpl = np.zeros(max(P, 1), np.int32)
    mask = np.zeros(max(P, 1) * W, np.uint64)
    code = np.zeros(max(P, 1) * max(T, 1), np.uint8)
    ev = C.c_int64(0)
    rc = emul().emul_fit_matrix(enc.catalog, enc.templates, enc.pods, _p(tmpl, C.c_int32), _p(mask, C.c_uint64),
                                _p(code, C.c_uint8), C.byref(ev))
    assert rc == 0
    return capi.FitOut(tmpl[:P], mask[:P * W].reshape(P, W), code[:P * T].reshape(P, T), int(ev.value))


def emul_solve(enc: capi.Encoded) -> capi.SolveRaw:
    buf = capi.ResultBuffers(enc.P, enc.T, enc.R, enc.W, enc.K, enc.NB)
    nodes = enc.nodes if enc.E > 0 else None
    rc = emul().emul_solve(enc.catalog, enc.templates, enc.pods, nodes, C.byref(buf.c))
    assert rc == 0, rc
    return buf.raw()


def emul_fit_rank_both(dv: np.ndarray, xs: np.ndarray):
    dv = np.ascontiguousarray(dv, np.int64)
    xs = np.ascontiguousarray(xs, np.int64)
    a, b = np.zeros(len(xs), np.int32), np.zeros(len(xs), np.int32)
    emul().emul_fit_rank_both(_p(dv, C.c_int64), len(dv), _p(xs, C.c_int64), len(xs), _p(a, C.c_int32), _p(b, C.c_int32))
    return a, b


def emul_price_keys(enc: capi.Encoded, za: int, ca: int, ct_reserved: int, ct_spot: int, ct_on_demand: int):
    """(cheapest, worst-launch) price bits per SKU under the admitted zone / capacity-type bit sets, from the device
    header's own functions compiled for the host"""
    cheapest = np.zeros(enc.S, np.uint64)
    worst = np.zeros(enc.S, np.uint64)
    f = emul().emul_price_keys
    f.restype = None
    f(enc.catalog, C.c_uint64(za), C.c_uint64(ca), C.c_int32(ct_reserved), C.c_int32(ct_spot), C.c_int32(ct_on_demand),
      _p(cheapest, C.c_uint64), _p(worst, C.c_uint64))
    return cheapest, worst


def emul_go_sort(cnt: np.ndarray) -> np.ndarray:
    cnt = np.ascontiguousarray(cnt, np.int32)
    ord_ = np.arange(len(cnt), dtype=np.int32)
    emul().emul_go_sort(_p(cnt, C.c_int32), len(cnt), _p(ord_, C.c_int32))
    return ord_


# ---- canonical form ------------------------------------------------------------------------------

def alive_templates(enc: capi.Encoded, code: np.ndarray) -> np.ndarray:
    """Device template slots that survived NewScheduler's pre-filter (scheduler.go:147-156)."""
    if code.shape[0] == 0:
        return np.ones(enc.T, bool)
    return code[0] != capi.FIT_DROPPED


def canon_fit(enc: capi.Encoded, fit: capi.FitOut) -> Dict[str, np.ndarray]:
    alive = alive_templates(enc, fit.code)
    remap = np.cumsum(alive) - 1
    tmpl = np.where(fit.tmpl >= 0, remap[np.maximum(fit.tmpl, 0)], -1).astype(np.int32)
    return dict(tmpl=tmpl, mask=fit.mask, code=fit.code[:, alive])


def canon_device(enc: capi.Encoded, raw: capi.SolveRaw) -> Dict[str, np.ndarray]:
    model = enc.model
    nk = len(model.key_names)
    alive = alive_templates(enc, raw.pod_fit_code)
    remap = np.cumsum(alive) - 1
    N = raw.n_claims
    req_mask = np.zeros((N, nk), np.uint64)
    req_flags = np.zeros((N, nk), np.uint8)
    req_gt = np.zeros((N, nk), np.int64)
    req_lt = np.zeros((N, nk), np.int64)
    for k in range(enc.K):
        mk = enc.key_model[k]
        req_mask[:, mk] = raw.claim_req_mask[:, k]
        req_flags[:, mk] = ((raw.claim_req_flags >> np.uint64(4 * k)) & np.uint64(0xF)).astype(np.uint8)
    for b, k in enumerate(enc.bounded_key):
        mk = enc.key_model[k]
        req_gt[:, mk] = raw.claim_req_gt[:, b]
        req_lt[:, mk] = raw.claim_req_lt[:, b]
    # pods of each claim in placement order
    return dict(
        n_claims=np.int64(N), n_unscheduled=np.int64(raw.n_unscheduled),
        pod_place_kind=raw.pod_place_kind, pod_place_index=raw.pod_place_index, pod_seq=raw.pod_seq,
        pod_fit_code=raw.pod_fit_code[:, alive],
        claim_template=remap[raw.claim_template].astype(np.int32) if N else raw.claim_template,
        claim_npods=raw.claim_npods, claim_requests=raw.claim_requests, claim_mask=raw.claim_mask,
        claim_n_options=raw.claim_n_options, claim_req_mask=req_mask, claim_req_flags=req_flags, claim_req_gt=req_gt,
        claim_req_lt=req_lt, claim_sku=raw.claim_sku, claim_price=raw.claim_price)


def canon_oracle(model, o) -> Dict[str, np.ndarray]:
    """oracle_binding.SolveOut -> the same dictionary.  The instance-type key is dropped from the
    requirement columns: its values name SKUs, the device keeps it as the template's type_admit mask
    and no pod may touch it, so it is the template's requirement unchanged on both sides."""
    kt = int(model.c.key_instance_type)
    req_mask, req_flags = o.claim_req_mask.copy(), o.claim_req_flags.copy()
    req_gt, req_lt = o.claim_req_gt.copy(), o.claim_req_lt.copy()
    if kt >= 0 and o.n_claims:
        req_mask[:, kt] = 0
        req_flags[:, kt] = 0
        req_gt[:, kt] = 0
        req_lt[:, kt] = 0
    return dict(
        n_claims=np.int64(o.n_claims), n_unscheduled=np.int64(o.n_unscheduled),
        pod_place_kind=o.pod_place_kind, pod_place_index=o.pod_place_index, pod_seq=o.pod_seq,
        pod_fit_code=o.pod_fit_code, claim_template=o.claim_template, claim_npods=o.claim_npods,
        claim_requests=o.claim_requests, claim_mask=o.claim_mask, claim_n_options=o.claim_n_options,
        claim_req_mask=req_mask, claim_req_flags=req_flags, claim_req_gt=req_gt, claim_req_lt=req_lt,
        claim_sku=o.claim_sku, claim_price=o.claim_price)


FIELDS = ["n_claims", "n_unscheduled", "pod_place_kind", "pod_place_index", "pod_seq", "pod_fit_code", "claim_template",
          "claim_npods", "claim_requests", "claim_mask", "claim_n_options", "claim_req_mask", "claim_req_flags",
          "claim_req_gt", "claim_req_lt", "claim_sku", "claim_price"]


def digest(c: Dict[str, np.ndarray]) -> str:
    h = hashlib.sha256()
    for f in FIELDS:
        a = np.ascontiguousarray(c[f])
        h.update(f.encode())
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def assert_same(a: Dict[str, np.ndarray], b: Dict[str, np.ndarray], what: str = ""):
    for f in FIELDS:
        x, y = np.asarray(a[f]), np.asarray(b[f])
        assert x.shape == y.shape, f"{what}: {f} shape {x.shape} vs {y.shape}"
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            i = tuple(bad[0])
            raise AssertionError(f"{what}: field {f} differs at {i}: {x[i]} vs {y[i]} ({len(bad)} mismatches of {x.size})")
    assert digest(a) == digest(b)


def mask_bits(mask: np.ndarray):
    """indices of the set bits of a little-endian u64 word array"""
    out = []
    for w, word in enumerate(np.asarray(mask, np.uint64).tolist()):
        while word:
            b = (word & -word).bit_length() - 1
            out.append(w * 64 + b)
            word &= word - 1
    return out


def capacity_type_values(model):
    """(model key id of karpenter.sh/capacity-type, value ids of spot, on-demand, reserved; -1 when not interned)"""
    from b200sched import model as M
    key = model.key_names.index(M.KEY_CAPACITY_TYPE)
    ids = {model.val_names[v]: v for v, k in enumerate(model.arrays["val_key"]) if int(k) == key}
    return key, ids.get("spot", -1), ids.get("on-demand", -1), ids.get("reserved", -1)


def device_commands_as_oracle(enc: capi.Encoded, cmds: Dict[str, np.ndarray], i: int) -> Dict[str, object]:
    """row i of b200s_consolidation_commands in the shape oracle_binding.consolidate reports"""
    return dict(decision=int(cmds["decision"][i]), reason=int(cmds["reason"][i]), n_options=int(cmds["n_options"][i]),
                spot_required=int(cmds["spot_required"][i]), price_tie=int(cmds["price_tie"][i]),
                head=cmds["head"][i].tolist(), mask=cmds["mask"][i].tolist())


def emul_solve_reg(enc: capi.Encoded, kcap: int = 0):
    """host twin of the register-resident pack (selector-less pods, no existing nodes): positions, shift words, the
    real sort only on demand, exactly as pack_reg_kernel does it; (result, stats = steps, full pdqsorts, rank look-ups,
    meetings with the order warp).  kcap: size of the count table (0: the kernel's 2 048)"""
    buf = capi.ResultBuffers(enc.P, enc.T, enc.R, enc.W, enc.K, enc.NB)
    stats = np.zeros(4, np.int64)
    rc = emul().emul_solve_reg(enc.catalog, enc.templates, enc.pods, C.byref(buf.c), _p(stats, C.c_int64), C.c_int32(kcap))
    assert rc == 0, rc
    return buf.raw(), stats


def assert_shared_matches_separate(enc_all, enc_sep, removed_node: int, sep, shared, what: str = "") -> int:
    """A simulation solved through b200s_solve_batch_shared (every node of the cluster in the list, the removed one
    unusable; node indices = rows of the shared node set) against the same simulation flattened on its own (its node list
    lacks `removed_node`, every later model node sits one index lower).  Returns the number of pods placed on nodes."""
    E = enc_all.E
    node_of_row = np.array([enc_all.nodes.contents.node_index[r] for r in range(E)])
    assert (sep.n_claims, sep.n_unscheduled) == (shared.n_claims, shared.n_unscheduled), what
    for f in ("pod_place_kind", "pod_seq", "pod_fit_code", "claim_template", "claim_npods", "claim_requests", "claim_mask",
              "claim_req_mask", "claim_req_flags", "claim_req_gt", "claim_req_lt", "claim_sku", "claim_price", "claim_n_options"):
        assert np.array_equal(getattr(sep, f), getattr(shared, f)), (what, f)
    sub_node = np.array([enc_sep.nodes.contents.node_index[r] for r in range(enc_sep.E)])
    orig = np.where(sub_node >= removed_node, sub_node + 1, sub_node)
    k = sep.pod_place_kind
    on_claim, on_node = k == capi.PLACE_CLAIM, k == capi.PLACE_NODE
    assert np.array_equal(sep.pod_place_index[on_claim], shared.pod_place_index[on_claim]), what
    assert np.array_equal(orig[sep.pod_place_index[on_node]], node_of_row[shared.pod_place_index[on_node]]), what
    assert not np.any(node_of_row[shared.pod_place_index[on_node]] == removed_node), (what, "a pod landed on the removed node")
    return int(on_node.sum())

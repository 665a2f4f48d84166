"""Semantic scheduling model -> ``b200s_model_t`` (include/b200sched.h, Level 1).

This is the Python mirror of what the Go cgo shim does with Karpenter objects: intern strings,
flatten requirement lists into CSR arrays, keep quantities as exact fixed-scale int64
(``resource.Quantity`` semantics).  Both the product encoder
(``b200s_encode``) and the CPU oracle consume the resulting structure.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

OP_IN, OP_NOTIN, OP_EXISTS, OP_DOESNOTEXIST, OP_GT, OP_LT = range(6)
EFFECT_NONE, EFFECT_NOSCHEDULE, EFFECT_PREFERNOSCHEDULE, EFFECT_NOEXECUTE = range(4)
TOLOP_EMPTY, TOLOP_EQUAL, TOLOP_EXISTS, TOLOP_OTHER = range(4)

# v1.WellKnownLabels (K/apis/v1/labels.go:83-92)
KEY_NODEPOOL = "karpenter.sh/nodepool"
KEY_ZONE = "topology.kubernetes.io/zone"
KEY_REGION = "topology.kubernetes.io/region"
KEY_INSTANCE_TYPE = "node.kubernetes.io/instance-type"
KEY_ARCH = "kubernetes.io/arch"
KEY_OS = "kubernetes.io/os"
KEY_CAPACITY_TYPE = "karpenter.sh/capacity-type"
KEY_WINDOWS_BUILD = "node.kubernetes.io/windows-build"
WELL_KNOWN = {KEY_NODEPOOL, KEY_ZONE, KEY_REGION, KEY_INSTANCE_TYPE, KEY_ARCH, KEY_OS, KEY_CAPACITY_TYPE,
              KEY_WINDOWS_BUILD}
# NormalizedLabels (K/apis/v1/labels.go:129-135), applied by NewRequirementWithFlexibility (requirement.go:48-50)
NORMALIZED = {
    "failure-domain.beta.kubernetes.io/zone": KEY_ZONE,
    "beta.kubernetes.io/arch": KEY_ARCH,
    "beta.kubernetes.io/os": KEY_OS,
    "beta.kubernetes.io/instance-type": KEY_INSTANCE_TYPE,
    "failure-domain.beta.kubernetes.io/region": KEY_REGION,
}


class Req(C.Structure):
    _fields_ = [("key", C.c_int32), ("op", C.c_int32), ("val_off", C.c_int32), ("val_cnt", C.c_int32),
                ("bound", C.c_int64), ("min_values", C.c_int32), ("_pad", C.c_int32)]


class Taint(C.Structure):
    _fields_ = [("key", C.c_int32), ("value", C.c_int32), ("effect", C.c_int32)]


class Toleration(C.Structure):
    _fields_ = [("key", C.c_int32), ("value", C.c_int32), ("op", C.c_int32), ("effect", C.c_int32)]


REQ_DTYPE = np.dtype([("key", "<i4"), ("op", "<i4"), ("val_off", "<i4"), ("val_cnt", "<i4"), ("bound", "<i8"),
                      ("min_values", "<i4"), ("_pad", "<i4")])
TAINT_DTYPE = np.dtype([("key", "<i4"), ("value", "<i4"), ("effect", "<i4")])
TOL_DTYPE = np.dtype([("key", "<i4"), ("value", "<i4"), ("op", "<i4"), ("effect", "<i4")])
assert REQ_DTYPE.itemsize == C.sizeof(Req) and TAINT_DTYPE.itemsize == C.sizeof(Taint)
assert TOL_DTYPE.itemsize == C.sizeof(Toleration)

_P = C.POINTER


class CModel(C.Structure):
    _fields_ = [
        ("n_keys", C.c_int32), ("key_wellknown", _P(C.c_uint8)), ("key_zone", C.c_int32),
        ("key_capacity_type", C.c_int32), ("key_instance_type", C.c_int32), ("n_vals", C.c_int32), ("val_key", _P(C.c_int32)),
        ("val_int", _P(C.c_int64)), ("val_isint", _P(C.c_uint8)),
        ("n_res", C.c_int32), ("res_cpu", C.c_int32), ("res_memory", C.c_int32),
        ("n_reqs", C.c_int32), ("reqs", _P(Req)), ("n_req_values", C.c_int32), ("req_values", _P(C.c_int32)),
        ("n_taints", C.c_int32), ("taints", _P(Taint)), ("n_tolerations", C.c_int32),
        ("tolerations", _P(Toleration)),
        ("n_skus", C.c_int32), ("sku_req_off", _P(C.c_int32)), ("sku_alloc", _P(C.c_int64)),
        ("sku_capacity", _P(C.c_int64)), ("sku_off_off", _P(C.c_int32)), ("n_offerings", C.c_int32),
        ("off_req_off", _P(C.c_int32)), ("off_price", _P(C.c_double)), ("off_available", _P(C.c_uint8)),
        ("n_pools", C.c_int32), ("pool_req_off", _P(C.c_int32)), ("pool_taint_off", _P(C.c_int32)),
        ("pool_weight", _P(C.c_int32)), ("pool_name_rank", _P(C.c_int32)), ("pool_daemon", _P(C.c_int64)),
        ("pool_has_limit", _P(C.c_uint8)), ("pool_limit", _P(C.c_int64)),
        ("n_pods", C.c_int32), ("pod_req_off", _P(C.c_int32)), ("pod_tol_off", _P(C.c_int32)),
        ("pod_requests", _P(C.c_int64)), ("pod_ts", _P(C.c_int64)), ("pod_uid_hi", _P(C.c_uint64)),
        ("pod_uid_lo", _P(C.c_uint64)),
        ("n_nodes", C.c_int32), ("node_req_off", _P(C.c_int32)), ("node_taint_off", _P(C.c_int32)),
        ("node_available", _P(C.c_int64)), ("node_initialized", _P(C.c_uint8)), ("node_name_rank", _P(C.c_int32)),
    ]


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(_P(ctype))


# A requirement in Python terms: (key name, op, values) where values is a list of strings for In/NotIn,
# an int for Gt/Lt and ignored otherwise.
ReqT = Tuple[str, int, object]


@dataclass
class Model:
    """Array form of the model.  ``c`` is the ctypes struct whose pointers alias the numpy arrays held here."""
    key_names: List[str]
    val_names: List[str]            # value string per value id
    res_names: List[str]
    sku_names: List[str]
    pool_names: List[str]
    node_names: List[str]
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    wellknown_extra: set = field(default_factory=set)  # provider labels registered into v1.WellKnownLabels
    c: Optional[CModel] = None

    def finalize(self) -> "Model":
        a = self.arrays
        m = CModel()
        kn = self.key_names
        a["key_wellknown"] = np.array([1 if (k in WELL_KNOWN or k in self.wellknown_extra) else 0 for k in kn],
                                      dtype=np.uint8)
        m.n_keys = len(kn)
        m.key_wellknown = _ptr(a["key_wellknown"], C.c_uint8)
        m.key_zone = kn.index(KEY_ZONE) if KEY_ZONE in kn else -1
        m.key_capacity_type = kn.index(KEY_CAPACITY_TYPE) if KEY_CAPACITY_TYPE in kn else -1
        m.key_instance_type = kn.index(KEY_INSTANCE_TYPE) if KEY_INSTANCE_TYPE in kn else -1
        m.n_vals = len(a["val_key"])
        m.val_key = _ptr(a["val_key"], C.c_int32)
        m.val_int = _ptr(a["val_int"], C.c_int64)
        m.val_isint = _ptr(a["val_isint"], C.c_uint8)
        m.n_res = len(self.res_names)
        m.res_cpu = self.res_names.index("cpu")
        m.res_memory = self.res_names.index("memory")
        m.n_reqs = len(a["reqs"])
        m.reqs = C.cast(a["reqs"].ctypes.data, _P(Req))
        m.n_req_values = len(a["req_values"])
        m.req_values = _ptr(a["req_values"], C.c_int32)
        m.n_taints = len(a["taints"])
        m.taints = C.cast(a["taints"].ctypes.data, _P(Taint))
        m.n_tolerations = len(a["tolerations"])
        m.tolerations = C.cast(a["tolerations"].ctypes.data, _P(Toleration))
        m.n_skus = len(a["sku_req_off"]) - 1
        for name, ct in (("sku_req_off", C.c_int32), ("sku_alloc", C.c_int64), ("sku_capacity", C.c_int64),
                         ("sku_off_off", C.c_int32), ("off_req_off", C.c_int32), ("off_price", C.c_double),
                         ("off_available", C.c_uint8), ("pool_req_off", C.c_int32), ("pool_taint_off", C.c_int32),
                         ("pool_weight", C.c_int32), ("pool_name_rank", C.c_int32), ("pool_daemon", C.c_int64),
                         ("pool_has_limit", C.c_uint8), ("pool_limit", C.c_int64), ("pod_req_off", C.c_int32),
                         ("pod_tol_off", C.c_int32), ("pod_requests", C.c_int64), ("pod_ts", C.c_int64),
                         ("pod_uid_hi", C.c_uint64), ("pod_uid_lo", C.c_uint64), ("node_req_off", C.c_int32),
                         ("node_taint_off", C.c_int32), ("node_available", C.c_int64),
                         ("node_initialized", C.c_uint8), ("node_name_rank", C.c_int32)):
            assert a[name].flags["C_CONTIGUOUS"], name
            setattr(m, name, _ptr(a[name], ct))
        m.n_offerings = len(a["off_price"])
        m.n_pools = len(a["pool_weight"])
        m.n_pods = len(a["pod_ts"])
        m.n_nodes = len(a["node_initialized"])
        self.c = m
        return self

    # convenience -----------------------------------------------------------------------------
    @property
    def n_pods(self) -> int:
        return int(self.c.n_pods)

    @property
    def n_skus(self) -> int:
        return int(self.c.n_skus)

    @property
    def n_res(self) -> int:
        return len(self.res_names)

    @property
    def n_words(self) -> int:
        return (self.n_skus + 63) // 64

    def key_value_bits(self) -> Dict[int, List[int]]:
        """value ids of each key in ascending id order (bit i of a result mask = i-th entry)."""
        out: Dict[int, List[int]] = {k: [] for k in range(len(self.key_names))}
        for v, k in enumerate(self.arrays["val_key"]):
            out[int(k)].append(v)
        return out


def slice_pods(m: "Model", lo: int, hi: int) -> "Model":
    """A model with pods [lo, hi) of m and everything else shared (row shard for the multi-GPU dense stage)."""
    a = dict(m.arrays)
    off = a["pod_req_off"].astype(np.int64)
    r0, r1 = int(off[lo]), int(off[hi])
    p0, p1 = int(off[0]), int(off[-1])
    a["reqs"] = np.ascontiguousarray(np.concatenate([a["reqs"][:p0], a["reqs"][r0:r1], a["reqs"][p1:]]))
    a["pod_req_off"] = np.ascontiguousarray((off[lo:hi + 1] - r0 + p0).astype(np.int32))
    a["node_req_off"] = np.ascontiguousarray((a["node_req_off"].astype(np.int64) - (p1 - p0) + (r1 - r0)).astype(np.int32))
    toff = a["pod_tol_off"].astype(np.int64)
    t0, t1 = int(toff[lo]), int(toff[hi])
    a["tolerations"] = np.ascontiguousarray(a["tolerations"][t0:t1])
    a["pod_tol_off"] = np.ascontiguousarray((toff[lo:hi + 1] - t0).astype(np.int32))
    for k in ("pod_requests", "pod_ts", "pod_uid_hi", "pod_uid_lo"):
        a[k] = np.ascontiguousarray(a[k][lo:hi])
    out = Model(key_names=list(m.key_names), val_names=list(m.val_names), res_names=list(m.res_names),
                sku_names=list(m.sku_names), pool_names=list(m.pool_names), node_names=list(m.node_names), arrays=a,
                wellknown_extra=set(m.wellknown_extra))
    return out.finalize()


def _csr_take(off: np.ndarray, idx: np.ndarray):
    """Rows idx of a CSR offsets array -> (flat element indices, new offsets starting at 0)."""
    off = off.astype(np.int64)
    lens = off[idx + 1] - off[idx]
    new_off = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(lens, out=new_off[1:])
    flat = np.repeat(off[idx] - new_off[:-1], lens) + np.arange(int(new_off[-1]), dtype=np.int64)
    return flat, new_off


def subset(m: "Model", pod_idx, node_idx) -> "Model":
    """A model with the given pods and existing nodes of m (catalog, pools, universes shared): the input of one
    consolidation simulation (K/controllers/disruption/helpers.go:52-143)."""
    a = dict(m.arrays)
    pod_idx = np.asarray(pod_idx, dtype=np.int64)
    node_idx = np.asarray(node_idx, dtype=np.int64)
    p0 = int(a["pod_req_off"][0])
    pflat, poff = _csr_take(a["pod_req_off"], pod_idx)
    nflat, noff = _csr_take(a["node_req_off"], node_idx)
    head = a["reqs"][:p0]
    a["reqs"] = np.ascontiguousarray(np.concatenate([head, a["reqs"][pflat], a["reqs"][nflat]]))
    a["pod_req_off"] = np.ascontiguousarray((poff + p0).astype(np.int32))
    a["node_req_off"] = np.ascontiguousarray((noff + p0 + len(pflat)).astype(np.int32))
    tflat, toff = _csr_take(a["pod_tol_off"], pod_idx)
    a["tolerations"] = np.ascontiguousarray(a["tolerations"][tflat])
    a["pod_tol_off"] = np.ascontiguousarray(toff.astype(np.int32))
    for k in ("pod_requests", "pod_ts", "pod_uid_hi", "pod_uid_lo"):
        a[k] = np.ascontiguousarray(a[k][pod_idx])
    # taints: pools first, then nodes
    n_pool_t = int(a["pool_taint_off"][-1])
    ntflat, ntoff = _csr_take(a["node_taint_off"], node_idx)
    a["taints"] = np.ascontiguousarray(np.concatenate([a["taints"][:n_pool_t], a["taints"][ntflat]]))
    a["node_taint_off"] = np.ascontiguousarray((ntoff + n_pool_t).astype(np.int32))
    for k in ("node_available", "node_initialized", "node_name_rank"):
        a[k] = np.ascontiguousarray(a[k][node_idx])
    out = Model(key_names=list(m.key_names), val_names=list(m.val_names), res_names=list(m.res_names),
                sku_names=list(m.sku_names), pool_names=list(m.pool_names),
                node_names=[m.node_names[i] for i in node_idx], arrays=a, wellknown_extra=set(m.wellknown_extra))
    return out.finalize()


def _parse_int(s: str) -> Tuple[bool, int]:
    """strconv.Atoi: optional sign followed by decimal digits, must fit int64."""
    t = s[1:] if s[:1] in "+-" else s
    if not t or not t.isdigit() or not t.isascii():
        return False, 0
    v = int(s)
    if v < -(1 << 63) or v >= (1 << 63):
        return False, 0
    return True, v


class ModelBuilder:
    """Object-at-a-time builder (tests, small cases).  ``synth.py`` fills the arrays in bulk instead."""

    def __init__(self, res_names: Sequence[str] = ("cpu", "memory", "ephemeral-storage", "nvidia.com/gpu", "pods")):
        self.res_names = list(res_names)
        self.key_names: List[str] = []
        self._key_id: Dict[str, int] = {}
        self.val_names: List[str] = []
        self.val_key: List[int] = []
        self._val_id: Dict[Tuple[int, str], int] = {}
        self._tstr: Dict[str, int] = {"": 0}
        self.reqs: List[Tuple[int, int, int, int, int, int]] = []
        self.req_values: List[int] = []
        self.taints: List[Tuple[int, int, int]] = []
        self.tolerations: List[Tuple[int, int, int, int]] = []
        self.skus: List[dict] = []
        self.offerings: List[dict] = []
        self.pools: List[dict] = []
        self.pods: List[dict] = []
        self.nodes: List[dict] = []
        self.wellknown_extra: set = set()

    # interning ---------------------------------------------------------------------------------
    def key(self, name: str) -> int:
        name = NORMALIZED.get(name, name)
        if name not in self._key_id:
            self._key_id[name] = len(self.key_names)
            self.key_names.append(name)
        return self._key_id[name]

    def value(self, key: int, s: str) -> int:
        kv = (key, s)
        if kv not in self._val_id:
            self._val_id[kv] = len(self.val_names)
            self.val_names.append(s)
            self.val_key.append(key)
        return self._val_id[kv]

    def tstr(self, s: str) -> int:
        if s not in self._tstr:
            self._tstr[s] = len(self._tstr)
        return self._tstr[s]

    # pools -------------------------------------------------------------------------------------
    def _add_reqs(self, reqs: Sequence[ReqT], min_values: Optional[Sequence[int]] = None) -> Tuple[int, int]:
        start = len(self.reqs)
        for i, (kname, op, vals) in enumerate(reqs):
            k = self.key(kname)
            off = len(self.req_values)
            cnt, bound = 0, 0
            if op in (OP_IN, OP_NOTIN):
                ids = [self.value(k, v) for v in vals]
                self.req_values.extend(ids)
                cnt = len(ids)
            elif op in (OP_GT, OP_LT):
                bound = int(vals)
                # the operand string itself is a value of the key in Go's eyes only through Atoi; intern
                # nothing: bounds are numbers
            mv = -1 if min_values is None else min_values[i]
            self.reqs.append((k, op, off, cnt, bound, mv))
        return start, len(self.reqs)

    def _add_taints(self, taints: Sequence[Tuple[str, str, int]]) -> Tuple[int, int]:
        start = len(self.taints)
        for key, value, effect in taints:
            self.taints.append((self.tstr(key), self.tstr(value), effect))
        return start, len(self.taints)

    def add_sku(self, name: str, reqs: Sequence[ReqT], alloc: Sequence[int], capacity: Sequence[int],
                offerings: Sequence[Tuple[Sequence[ReqT], float, bool]]) -> int:
        r0, r1 = self._add_reqs(reqs)
        o0 = len(self.offerings)
        for oreqs, price, avail in offerings:
            q0, q1 = self._add_reqs(oreqs)
            self.offerings.append(dict(r0=q0, r1=q1, price=float(price), avail=bool(avail)))
        self.skus.append(dict(name=name, r0=r0, r1=r1, alloc=list(alloc), cap=list(capacity), o0=o0,
                              o1=len(self.offerings)))
        return len(self.skus) - 1

    def add_pool(self, name: str, reqs: Sequence[ReqT], taints: Sequence[Tuple[str, str, int]] = (), weight: int = 0,
                 daemon: Optional[Sequence[int]] = None, limits: Optional[Dict[str, int]] = None,
                 min_values: Optional[Sequence[int]] = None) -> int:
        r0, r1 = self._add_reqs(reqs, min_values)
        t0, t1 = self._add_taints(taints)
        self.pools.append(dict(name=name, r0=r0, r1=r1, t0=t0, t1=t1, weight=weight,
                               daemon=list(daemon) if daemon is not None else [0] * len(self.res_names),
                               limits=limits or {}))
        return len(self.pools) - 1

    def add_pod(self, reqs: Sequence[ReqT], requests: Sequence[int], tolerations: Sequence[Tuple[str, int, str, int]] = (),
                ts: int = 0, uid: int = 0) -> int:
        r0, r1 = self._add_reqs(reqs)
        t0 = len(self.tolerations)
        for key, op, value, effect in tolerations:
            self.tolerations.append((self.tstr(key), self.tstr(value), op, effect))
        self.pods.append(dict(r0=r0, r1=r1, t0=t0, t1=len(self.tolerations), requests=list(requests), ts=ts, uid=uid))
        return len(self.pods) - 1

    def add_node(self, name: str, labels: Dict[str, str], available: Sequence[int],
                 taints: Sequence[Tuple[str, str, int]] = (), initialized: bool = True) -> int:
        r0, r1 = self._add_reqs([(k, OP_IN, [v]) for k, v in labels.items()])
        t0, t1 = self._add_taints(taints)
        self.nodes.append(dict(name=name, r0=r0, r1=r1, t0=t0, t1=t1, avail=list(available), init=initialized))
        return len(self.nodes) - 1

    # -------------------------------------------------------------------------------------------
    def build(self) -> Model:
        R = len(self.res_names)
        a: Dict[str, np.ndarray] = {}
        a["val_key"] = np.array(self.val_key, dtype=np.int32).reshape(-1)
        isint, ints = [], []
        for s in self.val_names:
            ok, v = _parse_int(s)
            isint.append(1 if ok else 0)
            ints.append(v)
        a["val_isint"] = np.array(isint, dtype=np.uint8).reshape(-1)
        a["val_int"] = np.array(ints, dtype=np.int64).reshape(-1)
        reqs = np.zeros(len(self.reqs), dtype=REQ_DTYPE)
        for i, (k, op, off, cnt, bound, mv) in enumerate(self.reqs):
            reqs[i] = (k, op, off, cnt, bound, mv, 0)
        a["reqs"] = reqs
        a["req_values"] = np.array(self.req_values, dtype=np.int32).reshape(-1)
        taints = np.zeros(len(self.taints), dtype=TAINT_DTYPE)
        for i, t in enumerate(self.taints):
            taints[i] = t
        a["taints"] = taints
        tols = np.zeros(len(self.tolerations), dtype=TOL_DTYPE)
        for i, t in enumerate(self.tolerations):
            tols[i] = t
        a["tolerations"] = tols

        # requirement ranges are contiguous per owner but owners interleave; re-pack per owner class
        def repack(owner_ranges):
            idx = []
            off = [0]
            for r0, r1 in owner_ranges:
                idx.extend(range(r0, r1))
                off.append(len(idx))
            return idx, off

        order: List[int] = []
        offs: Dict[str, List[int]] = {}
        for cls, rng in (("sku", [(s["r0"], s["r1"]) for s in self.skus]),
                         ("off", [(o["r0"], o["r1"]) for o in self.offerings]),
                         ("pool", [(p["r0"], p["r1"]) for p in self.pools]),
                         ("pod", [(p["r0"], p["r1"]) for p in self.pods]),
                         ("node", [(n["r0"], n["r1"]) for n in self.nodes])):
            idx, off = repack(rng)
            base = len(order)
            order.extend(idx)
            offs[cls] = [base + o for o in off]
        a["reqs"] = reqs[np.array(order, dtype=np.int64)] if order else reqs[:0]
        a["sku_req_off"] = np.array(offs["sku"], dtype=np.int32)
        a["off_req_off"] = np.array(offs["off"], dtype=np.int32)
        a["pool_req_off"] = np.array(offs["pool"], dtype=np.int32)
        a["pod_req_off"] = np.array(offs["pod"], dtype=np.int32)
        a["node_req_off"] = np.array(offs["node"], dtype=np.int32)

        # taints: pools then nodes
        tidx: List[int] = []
        poff = [0]
        for p in self.pools:
            tidx.extend(range(p["t0"], p["t1"]))
            poff.append(len(tidx))
        noff = [len(tidx)]
        for n in self.nodes:
            tidx.extend(range(n["t0"], n["t1"]))
            noff.append(len(tidx))
        a["taints"] = taints[np.array(tidx, dtype=np.int64)] if tidx else taints[:0]
        a["pool_taint_off"] = np.array(poff, dtype=np.int32)
        a["node_taint_off"] = np.array(noff, dtype=np.int32)

        S = len(self.skus)
        a["sku_alloc"] = np.array([s["alloc"] for s in self.skus], dtype=np.int64).reshape(S, R)
        a["sku_capacity"] = np.array([s["cap"] for s in self.skus], dtype=np.int64).reshape(S, R)
        a["sku_off_off"] = np.array([s["o0"] for s in self.skus] + [len(self.offerings)], dtype=np.int32)
        a["off_price"] = np.array([o["price"] for o in self.offerings], dtype=np.float64).reshape(-1)
        a["off_available"] = np.array([1 if o["avail"] else 0 for o in self.offerings], dtype=np.uint8).reshape(-1)

        NP_ = len(self.pools)
        a["pool_weight"] = np.array([p["weight"] for p in self.pools], dtype=np.int32).reshape(-1)
        names = sorted(p["name"] for p in self.pools)
        a["pool_name_rank"] = np.array([names.index(p["name"]) for p in self.pools], dtype=np.int32).reshape(-1)
        a["pool_daemon"] = np.array([p["daemon"] for p in self.pools], dtype=np.int64).reshape(NP_, R)
        hl = np.zeros((NP_, R), dtype=np.uint8)
        lim = np.zeros((NP_, R), dtype=np.int64)
        for i, p in enumerate(self.pools):
            for rn, v in p["limits"].items():
                hl[i, self.res_names.index(rn)] = 1
                lim[i, self.res_names.index(rn)] = v
        a["pool_has_limit"], a["pool_limit"] = hl, lim

        P = len(self.pods)
        a["pod_tol_off"] = np.array([p["t0"] for p in self.pods] + [len(self.tolerations)], dtype=np.int32)
        a["pod_requests"] = np.array([p["requests"] for p in self.pods], dtype=np.int64).reshape(P, R)
        a["pod_ts"] = np.array([p["ts"] for p in self.pods], dtype=np.int64).reshape(-1)
        a["pod_uid_hi"] = np.array([(p["uid"] >> 64) & ((1 << 64) - 1) for p in self.pods], dtype=np.uint64).reshape(-1)
        a["pod_uid_lo"] = np.array([p["uid"] & ((1 << 64) - 1) for p in self.pods], dtype=np.uint64).reshape(-1)

        E = len(self.nodes)
        a["node_available"] = np.array([n["avail"] for n in self.nodes], dtype=np.int64).reshape(E, R)
        a["node_initialized"] = np.array([1 if n["init"] else 0 for n in self.nodes], dtype=np.uint8).reshape(-1)
        nn = sorted(n["name"] for n in self.nodes)
        a["node_name_rank"] = np.array([nn.index(n["name"]) for n in self.nodes], dtype=np.int32).reshape(-1)

        for k in list(a):
            a[k] = np.ascontiguousarray(a[k])
        m = Model(key_names=list(self.key_names), val_names=list(self.val_names), res_names=list(self.res_names),
                  sku_names=[s["name"] for s in self.skus], pool_names=[p["name"] for p in self.pools],
                  node_names=[n["name"] for n in self.nodes], arrays=a,
                  wellknown_extra=set(self.wellknown_extra))
        return m.finalize()

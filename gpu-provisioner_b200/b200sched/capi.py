"""ctypes binding of libb200sched.so (include/b200sched.h).

The library is the product: it is built in-tree by ``__graft_entry__.build()`` and loaded from this
directory.  There is no Python or CPU fallback — a missing library or a missing sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .model import CModel, Model

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200sched.so")

MAX_KEYS, MAX_RES, MAX_TEMPLATES, MAX_BOUNDED = 16, 8, 16, 4
PRICE_NONE = 0x7FEFFFFFFFFFFFFF
FIT_NONE, FIT_TAINT, FIT_INCOMPATIBLE, FIT_LIMITS, FIT_DROPPED, FIT_FILTER = 0x00, 0x01, 0x02, 0x03, 0x7F, 0x80
PLACE_UNSCHEDULED, PLACE_NODE, PLACE_CLAIM = -1, 0, 1
RF_DEFINED, RF_COMPLEMENT, RF_HASGT, RF_HASLT = 1, 2, 4, 8  # per-key nibble of a requirement flags word
CMD_NONE, CMD_DELETE, CMD_REPLACE = 0, 1, 2
WHY_OK, WHY_UNSCHEDULABLE, WHY_MANY_CLAIMS, WHY_NOT_CHEAPER, WHY_S2S_DISABLED, WHY_S2S_TOO_FEW = 0, 1, 2, 3, 4, 5
S2S_MIN_TYPES = 15
MAX_INSTANCE_TYPES = 600  # nodeclaimtemplate.go:41
OK, EINVAL, ENODEVICE, ECUDA, ENOTREPRESENTABLE, ECAPACITY, ENOMEM, ESTATE, ESTALE = 0, -1, -2, -3, -4, -5, -6, -7, -8
CREATE_EBADNAME, CREATE_ENOTYPE, CREATE_ENOSTORAGE = -20, -21, -22

_P = C.POINTER


class CatalogSoA(C.Structure):
    _fields_ = [
        ("n_skus", C.c_int32), ("n_words", C.c_int32), ("n_res", C.c_int32), ("n_keys", C.c_int32),
        ("n_zones", C.c_int32), ("n_cts", C.c_int32), ("key_zone", C.c_int32), ("key_ct", C.c_int32),
        ("wellknown", C.c_uint32), ("key_nvals", C.c_int32 * MAX_KEYS), ("n_bounded", C.c_int32),
        ("bounded_key", C.c_int32 * MAX_BOUNDED),
        ("val_int", _P(C.c_int64)), ("val_isint", _P(C.c_uint64)),
        ("n_rows", C.c_int32), ("rows", _P(C.c_uint64)), ("row_nonneg", C.c_int32),
        ("res_nd", C.c_int32 * MAX_RES), ("res_dv0", C.c_int32 * MAX_RES), ("res_row0", C.c_int32 * MAX_RES),
        ("res_dv", _P(C.c_int64)), ("n_dv", C.c_int32), ("key_row0", C.c_int32 * MAX_KEYS), ("zc_row0", C.c_int32),
        ("price", _P(C.c_uint64)), ("alloc", _P(C.c_int64)), ("capacity", _P(C.c_int64)),
        ("price_max", _P(C.c_uint64)),
        ("cap_nd", C.c_int32 * MAX_RES), ("cap_dv0", C.c_int32 * MAX_RES), ("cap_row0", C.c_int32 * MAX_RES),
        ("cap_dv", _P(C.c_int64)), ("n_cap_dv", C.c_int32), ("layout_hash", C.c_uint64),
    ]


class ConsolidationSims(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("spot_to_spot", C.c_int32), ("ct_spot", C.c_int32), ("ct_on_demand", C.c_int32),
        ("ct_reserved", C.c_int32),
        ("n_unscheduled", _P(C.c_int32)), ("n_claims", _P(C.c_int32)), ("n_candidates", _P(C.c_int32)),
        ("all_spot", _P(C.c_uint8)), ("candidate_price", _P(C.c_double)),
        ("claim_mask", _P(C.c_uint64)), ("claim_req_mask", _P(C.c_uint64)), ("claim_req_flags", _P(C.c_uint64)),
        ("claim_req_gt", _P(C.c_int64)), ("claim_req_lt", _P(C.c_int64)),
    ]


class ConsolidationCmds(C.Structure):
    _fields_ = [
        ("decision", _P(C.c_int32)), ("reason", _P(C.c_int32)), ("n_options", _P(C.c_int32)), ("mask", _P(C.c_uint64)),
        ("spot_required", _P(C.c_uint8)), ("head", _P(C.c_int32)), ("price_tie", _P(C.c_uint8)),
    ]


class Templates(C.Structure):
    _fields_ = [
        ("n_templates", C.c_int32), ("pool_index", C.c_int32 * MAX_TEMPLATES), ("taints", C.c_uint64 * MAX_TEMPLATES),
        ("req_mask", (C.c_uint64 * MAX_KEYS) * MAX_TEMPLATES), ("req_flags", C.c_uint64 * MAX_TEMPLATES),
        ("req_gt", (C.c_int64 * MAX_BOUNDED) * MAX_TEMPLATES), ("req_lt", (C.c_int64 * MAX_BOUNDED) * MAX_TEMPLATES),
        ("daemon", (C.c_int64 * MAX_RES) * MAX_TEMPLATES), ("limit", (C.c_int64 * MAX_RES) * MAX_TEMPLATES),
        ("has_limit", (C.c_uint8 * MAX_RES) * MAX_TEMPLATES), ("type_admit", _P(C.c_uint64)),
    ]


class PodsSoA(C.Structure):
    _fields_ = [
        ("n_pods", C.c_int32), ("n_res", C.c_int32), ("n_keys", C.c_int32), ("n_bounded", C.c_int32),
        ("res_cpu", C.c_int32), ("res_memory", C.c_int32), ("n_sigs", C.c_int32), ("n_reqcls", C.c_int32),
        ("sig_id", _P(C.c_int32)), ("req_id", _P(C.c_int32)), ("ts", _P(C.c_int64)), ("uid_hi", _P(C.c_uint64)),
        ("uid_lo", _P(C.c_uint64)),
        ("sig_mask", _P(C.c_uint64)), ("sig_flags", _P(C.c_uint64)), ("sig_gt", _P(C.c_int32)), ("sig_lt", _P(C.c_int32)),
        ("sig_tolerated", _P(C.c_uint64)), ("rq_requests", _P(C.c_int64)), ("layout_hash", C.c_uint64),
    ]


class NodesSoA(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32), ("n_res", C.c_int32), ("n_keys", C.c_int32), ("node_index", _P(C.c_int32)),
        ("available", _P(C.c_int64)), ("req_mask", _P(C.c_uint64)), ("req_flags", _P(C.c_uint64)),
        ("taints", _P(C.c_uint64)), ("layout_hash", C.c_uint64),
    ]


class SolveOpts(C.Structure):
    _fields_ = [("want_masks", C.c_int32), ("reserved", C.c_int32 * 7)]


class Result(C.Structure):
    _fields_ = [
        ("cap_claims", C.c_int32), ("n_claims", C.c_int32), ("n_unscheduled", C.c_int32), ("fit_evals", C.c_int64),
        ("pod_place_kind", _P(C.c_int32)), ("pod_place_index", _P(C.c_int32)), ("pod_seq", _P(C.c_int32)),
        ("pod_fit_code", _P(C.c_uint8)),
        ("claim_template", _P(C.c_int32)), ("claim_npods", _P(C.c_int32)), ("claim_requests", _P(C.c_int64)),
        ("claim_mask", _P(C.c_uint64)), ("claim_req_mask", _P(C.c_uint64)), ("claim_req_flags", _P(C.c_uint64)),
        ("claim_req_gt", _P(C.c_int64)), ("claim_req_lt", _P(C.c_int64)), ("claim_sku", _P(C.c_int32)),
        ("claim_price", _P(C.c_uint64)), ("claim_n_options", _P(C.c_int32)),
    ]


_lib = None


def lib():
    """Loads libb200sched.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200S_LIB", LIB_PATH)  # experiments: another build of the same library
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(path)
    L.b200s_version.restype = C.c_int
    L.b200s_init.argtypes = [C.c_int, _P(C.c_void_p)]
    L.b200s_destroy.argtypes = [C.c_void_p]
    L.b200s_destroy.restype = None
    L.b200s_last_error.argtypes = [C.c_void_p]
    L.b200s_last_error.restype = C.c_char_p
    L.b200s_stream.argtypes = [C.c_void_p]
    L.b200s_stream.restype = C.c_void_p
    L.b200s_synchronize.argtypes = [C.c_void_p]
    L.b200s_event_create.argtypes = [C.c_void_p]
    L.b200s_event_create.restype = C.c_void_p
    L.b200s_event_record.argtypes = [C.c_void_p, C.c_void_p]
    L.b200s_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _P(C.c_float)]
    L.b200s_event_destroy.argtypes = [C.c_void_p]
    L.b200s_event_destroy.restype = None
    L.b200s_flush_l2.argtypes = [C.c_void_p, C.c_size_t]
    L.b200s_encode.argtypes = [_P(CModel), _P(C.c_void_p), C.c_char_p, C.c_size_t]
    L.b200s_encoded_free.argtypes = [C.c_void_p]
    L.b200s_encoded_free.restype = None
    L.b200s_encoded_catalog.argtypes = [C.c_void_p]
    L.b200s_encoded_catalog.restype = _P(CatalogSoA)
    L.b200s_encoded_templates.argtypes = [C.c_void_p]
    L.b200s_encoded_templates.restype = _P(Templates)
    L.b200s_encoded_pods.argtypes = [C.c_void_p]
    L.b200s_encoded_pods.restype = _P(PodsSoA)
    L.b200s_encoded_nodes.argtypes = [C.c_void_p]
    L.b200s_encoded_nodes.restype = _P(NodesSoA)
    L.b200s_encoded_key_model_id.argtypes = [C.c_void_p, C.c_int]
    L.b200s_encoded_value_model_id.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.b200s_catalog_upload.argtypes = [C.c_void_p, _P(CatalogSoA), _P(C.c_uint64)]
    L.b200s_catalog_release.argtypes = [C.c_void_p, C.c_uint64]
    L.b200s_fit_matrix.argtypes = [C.c_void_p, C.c_uint64, _P(PodsSoA), _P(Templates), _P(C.c_int32), _P(C.c_uint64),
                                   _P(C.c_uint8), _P(C.c_int64)]
    L.b200s_select_cheapest.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, _P(C.c_uint64), _P(C.c_uint64),
                                        _P(C.c_uint64), _P(C.c_int64), _P(C.c_int64), _P(C.c_int32), _P(C.c_uint64)]
    L.b200s_order_by_price.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint64),
                                       _P(C.c_int64), _P(C.c_int64), C.c_int32, _P(C.c_int32), _P(C.c_int32), _P(C.c_uint8)]
    L.b200s_consolidation_commands.argtypes = [C.c_void_p, C.c_uint64, _P(ConsolidationSims), _P(ConsolidationCmds)]
    L.b200s_solve.argtypes = [C.c_void_p, C.c_uint64, _P(PodsSoA), _P(NodesSoA), _P(Templates), _P(SolveOpts),
                              _P(Result)]
    L.b200s_solve_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, _P(_P(PodsSoA)), _P(_P(NodesSoA)),
                                    _P(Templates), _P(SolveOpts), _P(_P(Result))]
    L.b200s_solve_batch_shared.argtypes = [C.c_void_p, C.c_uint64, C.c_int32, _P(PodsSoA), _P(C.c_int32), _P(NodesSoA),
                                           _P(C.c_int32), _P(C.c_int32), _P(Templates), _P(SolveOpts), _P(_P(Result))]
    L.b200s_device_alloc.argtypes = [C.c_void_p, C.c_uint64, _P(C.c_void_p)]
    L.b200s_device_free.argtypes = [C.c_void_p, C.c_void_p]
    L.b200s_ipc_export.argtypes = [C.c_void_p, C.c_void_p, _P(C.c_uint8)]
    L.b200s_ipc_open.argtypes = [C.c_void_p, _P(C.c_uint8), _P(C.c_void_p)]
    L.b200s_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
    L.b200s_copy_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.b200s_resident_export_compact_peers.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, _P(C.c_void_p), _P(C.c_void_p),
                                                      C.c_int64, C.c_uint32, C.c_void_p]
    L.b200s_resident_create.argtypes = [C.c_void_p, C.c_uint64, _P(PodsSoA), _P(NodesSoA), _P(Templates),
                                        _P(C.c_void_p)]
    L.b200s_resident_free.argtypes = [C.c_void_p, C.c_void_p]
    L.b200s_resident_free.restype = None
    L.b200s_resident_fit_matrix.argtypes = [C.c_void_p, C.c_void_p]
    L.b200s_resident_solve.argtypes = [C.c_void_p, C.c_void_p]
    L.b200s_resident_fetch_fit.argtypes = [C.c_void_p, C.c_void_p, _P(C.c_int32), _P(C.c_uint64), _P(C.c_uint8),
                                           _P(C.c_int64)]
    L.b200s_resident_fetch_solve.argtypes = [C.c_void_p, C.c_void_p, _P(Result)]
    L.b200s_resident_copy_fit_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200s_resident_last_launches.argtypes = [C.c_void_p]
    L.b200s_resident_dense_kernel_ms.argtypes = [C.c_void_p, C.c_void_p, _P(C.c_float)]
    L.b200s_resident_pack_profile.argtypes = [C.c_void_p, C.c_void_p, _P(C.c_uint64)]
    L.b200s_resident_set_pack_profile.argtypes = [C.c_void_p, C.c_int]
    i64p, u8p = _P(C.c_int64), _P(C.c_uint8)
    L.b200s_pod_requests.argtypes = [C.c_int32, C.c_int32, C.c_int32, i64p, u8p, i64p, u8p, C.c_int32, i64p, u8p, i64p, u8p, u8p,
                                     i64p, i64p]
    L.b200s_allocatable.argtypes = [C.c_int32, C.c_int32, i64p, u8p, i64p, i64p, i64p, u8p, i64p]
    L.b200s_create_select.argtypes = [C.c_char_p, C.c_int32, _P(C.c_char_p), _P(C.c_double), C.c_int64,
                                      _P(C.c_int32), _P(C.c_int32), _P(C.c_int32)]
    _lib = L
    return L


class B200SError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libb200sched error {code}: {msg}")
        self.code = code


class NotRepresentable(B200SError):
    """The encoder rejected the input: the caller keeps the vendored CPU Solve for this batch."""


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(_P(ct))


class Encoded:
    """b200s_encode(model): the SoA views plus the decode maps."""

    def __init__(self, model: Model):
        L = lib()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = L.b200s_encode(C.byref(model.c), C.byref(h), err, 512)
        if rc == ENOTREPRESENTABLE:
            raise NotRepresentable(rc, err.value.decode())
        if rc != 0:
            raise B200SError(rc, err.value.decode())
        self.model = model
        self.handle = h
        self.catalog = L.b200s_encoded_catalog(h)
        self.templates = L.b200s_encoded_templates(h)
        self.pods = L.b200s_encoded_pods(h)
        self.layout_hash = int(self.catalog.contents.layout_hash)
        self.nodes = L.b200s_encoded_nodes(h)
        c = self.catalog.contents
        self.S, self.W, self.R, self.K, self.NB = c.n_skus, c.n_words, c.n_res, c.n_keys, c.n_bounded
        self.T = self.templates.contents.n_templates
        self.P = self.pods.contents.n_pods
        self.NS = self.pods.contents.n_sigs
        self.NQ = self.pods.contents.n_reqcls
        self.E = self.nodes.contents.n_nodes
        self.key_model = [L.b200s_encoded_key_model_id(h, k) for k in range(self.K)]
        self.bounded_key = [c.bounded_key[b] for b in range(self.NB)]
        self.pool_index = [self.templates.contents.pool_index[t] for t in range(self.T)]

    def node_rows(self, model_nodes: Sequence[int]) -> List[int]:
        """rows of the encoded node set (sortExistingNodes order) that hold the given model nodes"""
        row_of = {int(self.nodes.contents.node_index[r]): r for r in range(self.E)}
        return [row_of[int(i)] for i in model_nodes]

    def value_bit(self, model_key: int, model_value: int) -> int:
        """device bit of a model value inside its key's word, -1 when the encoding does not carry it"""
        if model_key not in self.key_model:
            return -1
        k = self.key_model.index(model_key)
        L = lib()
        for b in range(self.catalog.contents.key_nvals[k]):
            if L.b200s_encoded_value_model_id(self.handle, k, b) == model_value:
                return b
        return -1

    def close(self):
        if self.handle:
            lib().b200s_encoded_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class FitOut:
    tmpl: np.ndarray   # [P] device template index (weight order over all pools) or -1
    mask: np.ndarray   # [P, W]
    code: np.ndarray   # [P, T]
    evals: int


@dataclass
class SolveRaw:
    """b200s_result_t as numpy arrays, device-level indices (see tests/parity.py for the canonical form)."""
    n_claims: int
    n_unscheduled: int
    fit_evals: int
    pod_place_kind: np.ndarray
    pod_place_index: np.ndarray
    pod_seq: np.ndarray
    pod_fit_code: np.ndarray
    claim_template: np.ndarray
    claim_npods: np.ndarray
    claim_requests: np.ndarray
    claim_mask: np.ndarray
    claim_req_mask: np.ndarray
    claim_req_flags: np.ndarray
    claim_req_gt: np.ndarray
    claim_req_lt: np.ndarray
    claim_sku: np.ndarray
    claim_price: np.ndarray
    claim_n_options: np.ndarray


class ResultBuffers:
    def __init__(self, P: int, T: int, R: int, W: int, K: int, NB: int, cap: Optional[int] = None):
        cap = max(P if cap is None else cap, 1)
        P1, T1 = max(P, 1), max(T, 1)
        self.dims = (P, T, R, W, K, NB, cap)
        self.a = dict(
            pod_place_kind=np.zeros(P1, np.int32), pod_place_index=np.zeros(P1, np.int32),
            pod_seq=np.zeros(P1, np.int32), pod_fit_code=np.zeros(P1 * T1, np.uint8),
            claim_template=np.zeros(cap, np.int32), claim_npods=np.zeros(cap, np.int32),
            claim_requests=np.zeros(cap * R, np.int64), claim_mask=np.zeros(cap * W, np.uint64),
            claim_req_mask=np.zeros(cap * max(K, 1), np.uint64), claim_req_flags=np.zeros(cap, np.uint64),
            claim_req_gt=np.zeros(cap * max(NB, 1), np.int64), claim_req_lt=np.zeros(cap * max(NB, 1), np.int64),
            claim_sku=np.zeros(cap, np.int32), claim_price=np.zeros(cap, np.uint64),
            claim_n_options=np.zeros(cap, np.int32))
        ct = {np.dtype(np.int32): C.c_int32, np.dtype(np.int64): C.c_int64, np.dtype(np.uint64): C.c_uint64,
              np.dtype(np.uint8): C.c_uint8}
        self.c = Result()
        self.c.cap_claims = cap
        for k, v in self.a.items():
            setattr(self.c, k, _p(v, ct[v.dtype]))

    def raw(self) -> SolveRaw:
        P, T, R, W, K, NB, _ = self.dims
        N = self.c.n_claims
        a = self.a
        return SolveRaw(
            n_claims=N, n_unscheduled=self.c.n_unscheduled, fit_evals=self.c.fit_evals,
            pod_place_kind=a["pod_place_kind"][:P].copy(), pod_place_index=a["pod_place_index"][:P].copy(),
            pod_seq=a["pod_seq"][:P].copy(), pod_fit_code=a["pod_fit_code"][:P * T].reshape(P, T).copy(),
            claim_template=a["claim_template"][:N].copy(), claim_npods=a["claim_npods"][:N].copy(),
            claim_requests=a["claim_requests"][:N * R].reshape(N, R).copy(),
            claim_mask=a["claim_mask"][:N * W].reshape(N, W).copy(),
            claim_req_mask=a["claim_req_mask"][:N * K].reshape(N, K).copy(),
            claim_req_flags=a["claim_req_flags"][:N].copy(),
            claim_req_gt=a["claim_req_gt"][:N * NB].reshape(N, NB).copy(),
            claim_req_lt=a["claim_req_lt"][:N * NB].reshape(N, NB).copy(),
            claim_sku=a["claim_sku"][:N].copy(), claim_price=a["claim_price"][:N].copy(),
            claim_n_options=a["claim_n_options"][:N].copy())


class BatchResultBuffers:
    """Output buffers of a batch of Solves in one allocation per field (problem b's rows start at its first pod, the way
    the library lays them out itself) and the b200s_result_t structs as one structured array: a batch of hundreds of small
    simulations spends its host time here otherwise."""
    _FIELDS = [("pod_place_kind", np.int32, "P", 1), ("pod_place_index", np.int32, "P", 1), ("pod_seq", np.int32, "P", 1),
               ("pod_fit_code", np.uint8, "P", "T"), ("claim_template", np.int32, "C", 1), ("claim_npods", np.int32, "C", 1),
               ("claim_requests", np.int64, "C", "R"), ("claim_mask", np.uint64, "C", "W"), ("claim_req_mask", np.uint64, "C", "K"),
               ("claim_req_flags", np.uint64, "C", 1), ("claim_req_gt", np.int64, "C", "NB"), ("claim_req_lt", np.int64, "C", "NB"),
               ("claim_sku", np.int32, "C", 1), ("claim_price", np.uint64, "C", 1), ("claim_n_options", np.int32, "C", 1)]

    def __init__(self, pods_per_problem: Sequence[int], T: int, R: int, W: int, K: int, NB: int):
        n = len(pods_per_problem)
        self.n = n
        self.P = np.asarray(pods_per_problem, np.int64)
        self.cap = np.maximum(self.P, 1)
        self.off = np.zeros(n + 1, np.int64)
        self.off[1:] = np.cumsum(self.cap)           # rows of problem b: off[b] .. off[b] + cap[b]
        total = int(self.off[-1])
        self.width = {"T": max(T, 1), "R": R, "W": W, "K": max(K, 1), "NB": max(NB, 1), 1: 1}
        self.dims = (T, R, W, K, NB)
        dt = np.dtype([("cap_claims", np.int32), ("n_claims", np.int32), ("n_unscheduled", np.int32), ("_pad", np.int32),
                       ("fit_evals", np.int64)] + [(f, np.uint64) for f, _, _, _ in self._FIELDS])
        assert dt.itemsize == C.sizeof(Result)
        self.res = np.zeros(n, dt)
        self.res["cap_claims"] = self.cap
        self.a = {}
        for f, t, _, w in self._FIELDS:
            wd = self.width[w]
            arr = np.zeros(total * wd, t)
            self.a[f] = arr
            self.res[f] = arr.ctypes.data + self.off[:-1] * (wd * arr.itemsize)
        self.ptrs = (self.res.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(dt.itemsize)).astype(np.uint64)

    def out_pointers(self):
        return self.ptrs.ctypes.data_as(_P(_P(Result)))

    def raws(self) -> List[SolveRaw]:
        T, R, W, K, NB = self.dims
        out = []
        a = self.a
        for b in range(self.n):
            o, P, N = int(self.off[b]), int(self.P[b]), int(self.res["n_claims"][b])

            def rows(f, n_rows, w):
                wd = self.width[w]
                v = a[f][o * wd:(o + n_rows) * wd]
                return v.copy() if w == 1 else v.reshape(n_rows, wd).copy()
            out.append(SolveRaw(
                n_claims=N, n_unscheduled=int(self.res["n_unscheduled"][b]), fit_evals=int(self.res["fit_evals"][b]),
                pod_place_kind=rows("pod_place_kind", P, 1), pod_place_index=rows("pod_place_index", P, 1),
                pod_seq=rows("pod_seq", P, 1), pod_fit_code=rows("pod_fit_code", P, "T")[:, :T],
                claim_template=rows("claim_template", N, 1), claim_npods=rows("claim_npods", N, 1),
                claim_requests=rows("claim_requests", N, "R"), claim_mask=rows("claim_mask", N, "W"),
                claim_req_mask=rows("claim_req_mask", N, "K")[:, :K], claim_req_flags=rows("claim_req_flags", N, 1),
                claim_req_gt=rows("claim_req_gt", N, "NB")[:, :NB], claim_req_lt=rows("claim_req_lt", N, "NB")[:, :NB],
                claim_sku=rows("claim_sku", N, 1), claim_price=rows("claim_price", N, 1),
                claim_n_options=rows("claim_n_options", N, 1)))
        return out


class Context:
    """b200s_init on one device.  Raises B200SError(ENODEVICE) when there is no sm_100 GPU."""

    def __init__(self, device: int = 0):
        L = lib()
        h = C.c_void_p()
        rc = L.b200s_init(device, C.byref(h))
        if rc != 0:
            raise B200SError(rc, "b200s_init failed (no usable sm_100 device?)")
        self.h = h
        self.L = L
        self._layouts = {}  # catalog id -> layout_hash of the encode it was uploaded from

    def close(self):
        if self.h:
            self.L.b200s_destroy(self.h)
            self.h = None

    def _check(self, rc: int):
        if rc != 0:
            raise B200SError(rc, self.L.b200s_last_error(self.h).decode())

    def upload_catalog(self, enc: Encoded) -> int:
        cid = C.c_uint64(0)
        self._check(self.L.b200s_catalog_upload(self.h, enc.catalog, C.byref(cid)))
        self._layouts[cid.value] = enc.layout_hash
        return cid.value

    def catalog_layout(self, cid: int):
        """layout_hash of the encode catalog `cid` was uploaded from (None when unknown)"""
        return self._layouts.get(cid)

    def release_catalog(self, cid: int):
        self._layouts.pop(cid, None)
        self._check(self.L.b200s_catalog_release(self.h, cid))

    def fit_matrix(self, cid: int, enc: Encoded) -> FitOut:
        P, W, T = enc.P, enc.W, enc.T
        tmpl = np.zeros(max(P, 1), np.int32)
        mask = np.zeros(max(P, 1) * W, np.uint64)
        code = np.zeros(max(P, 1) * max(T, 1), np.uint8)
        ev = C.c_int64(0)
        self._check(self.L.b200s_fit_matrix(self.h, cid, enc.pods, enc.templates, _p(tmpl, C.c_int32),
                                            _p(mask, C.c_uint64), _p(code, C.c_uint8), C.byref(ev)))
        return FitOut(tmpl[:P], mask[:P * W].reshape(P, W), code[:P * T].reshape(P, T), int(ev.value))

    def solve(self, cid: int, enc: Encoded) -> SolveRaw:
        buf = ResultBuffers(enc.P, enc.T, enc.R, enc.W, enc.K, enc.NB)
        opts = SolveOpts()
        opts.want_masks = 1
        nodes = enc.nodes if enc.E > 0 else None
        self._check(self.L.b200s_solve(self.h, cid, enc.pods, nodes, enc.templates, C.byref(opts), C.byref(buf.c)))
        return buf.raw()

    def solve_batch(self, cid: int, encs: Sequence[Encoded]) -> List[SolveRaw]:
        n = len(encs)
        e0 = encs[0]
        bufs = BatchResultBuffers([e.P for e in encs], e0.T, e0.R, e0.W, e0.K, e0.NB)
        pods = (_P(PodsSoA) * n)(*[e.pods for e in encs])
        nodes = (_P(NodesSoA) * n)(*[e.nodes for e in encs])
        opts = SolveOpts()
        opts.want_masks = 1
        self._check(self.L.b200s_solve_batch(self.h, cid, n, pods, nodes, e0.templates, C.byref(opts), bufs.out_pointers()))
        return bufs.raws()

    def solve_batch_shared(self, cid: int, enc: Encoded, pod_off: Sequence[int], excluded: Sequence[Sequence[int]]) -> List[SolveRaw]:
        """b200s_solve_batch_shared: problem b = pods pod_off[b]..pod_off[b+1] of `enc` against enc's nodes without the
        node ROWS excluded[b] (rows of enc.nodes, i.e. sortExistingNodes order)."""
        n = len(pod_off) - 1
        po = np.ascontiguousarray(pod_off, np.int32)
        eo = np.zeros(n + 1, np.int32)
        eo[1:] = np.cumsum([len(x) for x in excluded])
        er = np.ascontiguousarray([r for x in excluded for r in x] or [0], np.int32)
        bufs = BatchResultBuffers(np.diff(po), enc.T, enc.R, enc.W, enc.K, enc.NB)
        opts = SolveOpts()
        opts.want_masks = 1
        self._check(self.L.b200s_solve_batch_shared(self.h, cid, n, enc.pods, _p(po, C.c_int32), enc.nodes, _p(eo, C.c_int32),
                                                    _p(er, C.c_int32), enc.templates, C.byref(opts), bufs.out_pointers()))
        return bufs.raws()

    def select_cheapest(self, cid: int, enc: Encoded, masks: np.ndarray, req_mask: np.ndarray, req_flags: np.ndarray,
                        req_gt: np.ndarray, req_lt: np.ndarray):
        n = masks.shape[0]
        masks = np.ascontiguousarray(masks, np.uint64)
        req_mask = np.ascontiguousarray(req_mask, np.uint64)    # [K, n]
        req_flags = np.ascontiguousarray(req_flags, np.uint64)
        req_gt = np.ascontiguousarray(req_gt, np.int64)          # [NB, n]
        req_lt = np.ascontiguousarray(req_lt, np.int64)
        sku = np.zeros(max(n, 1), np.int32)
        price = np.zeros(max(n, 1), np.uint64)
        self._check(self.L.b200s_select_cheapest(self.h, cid, n, _p(masks, C.c_uint64), _p(req_mask, C.c_uint64),
                                                 _p(req_flags, C.c_uint64), _p(req_gt, C.c_int64), _p(req_lt, C.c_int64),
                                                 _p(sku, C.c_int32), _p(price, C.c_uint64)))
        return sku[:n], price[:n]

    def order_by_price(self, cid: int, enc: Encoded, raw: SolveRaw, max_types: int = MAX_INSTANCE_TYPES):
        """ToNodeClaim's instance-type list (nodeclaimtemplate.go:77-82) for every NodeClaim of a Solve result:
        (order [N, max_types] in price order, -1 padded; count [N]; tie [N])."""
        n = raw.n_claims
        n1, K, NB = max(n, 1), max(enc.K, 1), max(enc.NB, 1)
        masks = np.ascontiguousarray(raw.claim_mask, np.uint64)
        rm = np.zeros((n1, K), np.uint64)
        gt = np.zeros((n1, NB), np.int64)
        lt = np.zeros((n1, NB), np.int64)
        rm[:n, :enc.K], gt[:n, :enc.NB], lt[:n, :enc.NB] = raw.claim_req_mask, raw.claim_req_gt, raw.claim_req_lt
        rf = np.ascontiguousarray(raw.claim_req_flags, np.uint64)
        order = np.full((n1, max_types), -1, np.int32)
        count = np.zeros(n1, np.int32)
        tie = np.zeros(n1, np.uint8)
        self._check(self.L.b200s_order_by_price(self.h, cid, n, _p(masks, C.c_uint64), _p(rm, C.c_uint64), _p(rf, C.c_uint64),
                                                _p(gt, C.c_int64), _p(lt, C.c_int64), max_types, _p(order, C.c_int32),
                                                _p(count, C.c_int32), _p(tie, C.c_uint8)))
        return order[:n], count[:n], tie[:n]

    def consolidation_commands(self, cid: int, enc: Encoded, sims: Sequence[SolveRaw], candidate_price: Sequence[float],
                               all_spot: Sequence[bool], n_candidates: Sequence[int], ct_spot: int, ct_on_demand: int,
                               ct_reserved: int = -1, spot_to_spot: bool = False, n_unscheduled=None) -> dict:
        """computeConsolidation (consolidation.go:133-307) over finished simulations (b200s_solve_batch results)."""
        n = len(sims)
        W, K, NB = enc.W, max(enc.K, 1), max(enc.NB, 1)
        cm = np.zeros((max(n, 1), W), np.uint64)
        rm = np.zeros((max(n, 1), K), np.uint64)
        rf = np.zeros(max(n, 1), np.uint64)
        gt = np.zeros((max(n, 1), NB), np.int64)
        lt = np.zeros((max(n, 1), NB), np.int64)
        for i, r in enumerate(sims):
            if r.n_claims == 1:
                cm[i] = r.claim_mask[0]
                rm[i, :enc.K] = r.claim_req_mask[0]
                rf[i] = r.claim_req_flags[0]
                gt[i, :enc.NB] = r.claim_req_gt[0]
                lt[i, :enc.NB] = r.claim_req_lt[0]
        un = np.asarray([r.n_unscheduled for r in sims] if n_unscheduled is None else n_unscheduled, np.int32)
        nc = np.asarray([r.n_claims for r in sims], np.int32)
        cand = np.asarray(n_candidates, np.int32)
        spot = np.asarray(all_spot, np.uint8)
        price = np.asarray(candidate_price, np.float64)
        s = ConsolidationSims()
        s.n, s.spot_to_spot, s.ct_spot, s.ct_on_demand, s.ct_reserved = n, int(spot_to_spot), ct_spot, ct_on_demand, ct_reserved
        s.n_unscheduled, s.n_claims, s.n_candidates = _p(un, C.c_int32), _p(nc, C.c_int32), _p(cand, C.c_int32)
        s.all_spot, s.candidate_price = _p(spot, C.c_uint8), _p(price, C.c_double)
        s.claim_mask, s.claim_req_mask, s.claim_req_flags = _p(cm, C.c_uint64), _p(rm, C.c_uint64), _p(rf, C.c_uint64)
        s.claim_req_gt, s.claim_req_lt = _p(gt, C.c_int64), _p(lt, C.c_int64)
        n1 = max(n, 1)
        out = dict(decision=np.zeros(n1, np.int32), reason=np.zeros(n1, np.int32), n_options=np.zeros(n1, np.int32),
                   mask=np.zeros((n1, W), np.uint64), spot_required=np.zeros(n1, np.uint8),
                   head=np.zeros((n1, S2S_MIN_TYPES), np.int32), price_tie=np.zeros(n1, np.uint8))
        c = ConsolidationCmds()
        c.decision, c.reason, c.n_options = _p(out["decision"], C.c_int32), _p(out["reason"], C.c_int32), _p(out["n_options"], C.c_int32)
        c.mask, c.spot_required = _p(out["mask"], C.c_uint64), _p(out["spot_required"], C.c_uint8)
        c.head, c.price_tie = _p(out["head"], C.c_int32), _p(out["price_tie"], C.c_uint8)
        self._check(self.L.b200s_consolidation_commands(self.h, cid, C.byref(s), C.byref(c)))
        return {k: v[:n] for k, v in out.items()}

    # ---- resident (device-timed) path ----
    def resident(self, cid: int, enc: Encoded) -> "Resident":
        return Resident(self, cid, enc)

    def event(self):
        return self.L.b200s_event_create(self.h)

    def record(self, ev):
        self._check(self.L.b200s_event_record(self.h, ev))

    def elapsed_ms(self, a, b) -> float:
        ms = C.c_float(0)
        self._check(self.L.b200s_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self, nbytes: int = 256 << 20):
        self._check(self.L.b200s_flush_l2(self.h, nbytes))

    # ---- plumbing for the peer-store export (b200s_resident_export_compact_peers) ----
    def device_alloc(self, nbytes: int) -> int:
        out = C.c_void_p()
        self._check(self.L.b200s_device_alloc(self.h, C.c_uint64(int(nbytes)), C.byref(out)))
        return int(out.value)

    def device_free(self, ptr: int) -> None:
        self._check(self.L.b200s_device_free(self.h, C.c_void_p(int(ptr))))

    def ipc_export(self, ptr: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.L.b200s_ipc_export(self.h, C.c_void_p(int(ptr)), buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        self._check(self.L.b200s_ipc_open(self.h, buf, C.byref(out)))
        return int(out.value)

    def ipc_close(self, ptr: int) -> None:
        self._check(self.L.b200s_ipc_close(self.h, C.c_void_p(int(ptr))))

    def copy_device(self, dst: int, src: int, nbytes: int) -> None:
        self._check(self.L.b200s_copy_device(self.h, C.c_void_p(int(dst)), C.c_void_p(int(src)), C.c_uint64(int(nbytes))))

    def synchronize(self):
        self._check(self.L.b200s_synchronize(self.h))

    def stream(self) -> int:
        """the context's cudaStream_t as an integer (e.g. for torch.cuda.ExternalStream)"""
        return int(self.L.b200s_stream(self.h) or 0)


def compact_row_bytes(T: int) -> int:
    return int(lib().b200s_compact_row_bytes(T))


class Resident:
    def __init__(self, ctx: Context, cid: int, enc: Encoded):
        self.ctx, self.enc = ctx, enc
        h = C.c_void_p()
        nodes = enc.nodes if enc.E > 0 else None
        ctx._check(ctx.L.b200s_resident_create(ctx.h, cid, enc.pods, nodes, enc.templates, C.byref(h)))
        self.h = h

    def fit_matrix(self):
        self.ctx._check(self.ctx.L.b200s_resident_fit_matrix(self.ctx.h, self.h))

    def solve(self):
        self.ctx._check(self.ctx.L.b200s_resident_solve(self.ctx.h, self.h))

    def copy_fit_device(self, d_tmpl: int, d_mask: int, d_code: int):
        """Device pointers (ints) of caller-owned buffers, e.g. torch tensors' data_ptr()."""
        self.ctx._check(self.ctx.L.b200s_resident_copy_fit_device(self.ctx.h, self.h, C.c_void_p(d_tmpl), C.c_void_p(d_mask),
                                                                  C.c_void_p(d_code)))

    def export_compact(self, d_out: int):
        """compact rows (template + attempt codes, compact_row_bytes(T) each) of this resident's pods into device memory"""
        self.ctx._check(self.ctx.L.b200s_resident_export_compact(self.ctx.h, self.h, C.c_void_p(d_out)))

    def export_compact_peers(self, peer_out: Sequence[int], peer_flags: Sequence[int], me: int, row_offset: int, epoch: int,
                             d_timed_out: int):
        """b200s_resident_export_compact_peers: this shard's compact rows stored into every rank's gathered buffer (device
        pointers valid in this process: own memory or CUDA-IPC mappings) + the flag barrier of epoch `epoch`"""
        n = len(peer_out)
        po = (C.c_void_p * n)(*[C.c_void_p(int(x)) for x in peer_out])
        pf = (C.c_void_p * n)(*[C.c_void_p(int(x)) for x in peer_flags])
        self.ctx._check(self.ctx.L.b200s_resident_export_compact_peers(self.ctx.h, self.h, n, int(me), po, pf,
                                                                       C.c_int64(int(row_offset)), C.c_uint32(int(epoch)),
                                                                       C.c_void_p(int(d_timed_out))))

    def import_compact(self, d_in: int):
        """compact rows of ALL of this resident's pods, from device memory (gathered from the ranks that evaluated them)"""
        self.ctx._check(self.ctx.L.b200s_resident_import_compact(self.ctx.h, self.h, C.c_void_p(d_in)))

    def pack(self):
        """FFD order + pack + cheapest SKU on imported per-pod results"""
        self.ctx._check(self.ctx.L.b200s_resident_pack(self.ctx.h, self.h))

    def launches(self) -> int:
        return self.ctx.L.b200s_resident_last_launches(self.h)

    def dense_kernel_ms(self) -> float:
        """Device time of the dense fit kernel alone in the last stage A (blocks until it has finished)."""
        ms = C.c_float(0.0)
        self.ctx._check(self.ctx.L.b200s_resident_dense_kernel_ms(self.ctx.h, self.h, C.byref(ms)))
        return float(ms.value)

    def set_pack_profile(self, on: bool) -> None:
        """Run the next solves with the pack kernel's cycle counters (small selector-less batches run without)."""
        self.ctx._check(self.ctx.L.b200s_resident_set_pack_profile(self.h, 1 if on else 0))

    def pack_profile(self):
        out = (C.c_uint64 * 12)()
        self.ctx._check(self.ctx.L.b200s_resident_pack_profile(self.ctx.h, self.h, out))
        names = ["start", "nodes", "scan", "commit", "normalise", "-", "--", "steps", "queue_warp", "order_warp", "claim_group", "sorts"]
        d = dict(zip(names, [int(x) for x in out]))
        d["full_sorts"] = d["sorts"] >> 40
        d["full_sort_cycles"] = d.pop("sorts") & ((1 << 40) - 1)
        d.pop("-"), d.pop("--")
        return d

    def reg_pack_profile(self):
        """Cycle counters of the register-resident pack (pack_reg_kernels.cu), problem 0, of the last Solve run with
        B200S_REG_PROFILE set in the environment: per-role SM cycles, steps, order-warp meetings and full pdqsorts."""
        out = (C.c_uint64 * 12)()
        self.ctx._check(self.ctx.L.b200s_resident_pack_profile(self.ctx.h, self.h, out))
        names = ["claim_eval", "claim_barrier", "claim_commit", "steps", "order_warp_wait", "order_warp_sort", "-",
                 "claim_post", "claim_meet", "meets", "full_sorts", "queue_warp"]
        d = dict(zip(names, [int(x) for x in out]))
        d.pop("-")
        return d

    def fetch_fit(self) -> FitOut:
        e = self.enc
        tmpl = np.zeros(max(e.P, 1), np.int32)
        mask = np.zeros(max(e.P, 1) * e.W, np.uint64)
        code = np.zeros(max(e.P, 1) * max(e.T, 1), np.uint8)
        ev = C.c_int64(0)
        self.ctx._check(self.ctx.L.b200s_resident_fetch_fit(self.ctx.h, self.h, _p(tmpl, C.c_int32), _p(mask, C.c_uint64),
                                                            _p(code, C.c_uint8), C.byref(ev)))
        return FitOut(tmpl[:e.P], mask[:e.P * e.W].reshape(e.P, e.W), code[:e.P * e.T].reshape(e.P, e.T), int(ev.value))

    def fetch_solve(self) -> SolveRaw:
        e = self.enc
        buf = ResultBuffers(e.P, e.T, e.R, e.W, e.K, e.NB)
        self.ctx._check(self.ctx.L.b200s_resident_fetch_solve(self.ctx.h, self.h, C.byref(buf.c)))
        return buf.raw()

    def close(self):
        if self.h:
            self.ctx.L.b200s_resident_free(self.ctx.h, self.h)
            self.h = None


def create_select(name: str, instance_types: Sequence[str], storage_bytes: int, prices: Optional[Sequence[float]] = None):
    """instance.Provider.Create's vmSize pick (pkg/providers/instance/instance.go:76-95, :321-369)."""
    L = lib()
    n = len(instance_types)
    arr = (C.c_char_p * max(n, 1))(*[s.encode() for s in instance_types])
    pr = None
    if prices is not None:
        pr = (C.c_double * max(n, 1))(*prices)
    chosen, disk, gpu = C.c_int32(-1), C.c_int32(0), C.c_int32(0)
    rc = L.b200s_create_select(name.encode(), n, arr, pr, int(storage_bytes), C.byref(chosen), C.byref(disk), C.byref(gpu))
    return rc, chosen.value, disk.value, gpu.value

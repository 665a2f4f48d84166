"""Row sharding of the dense stage across ranks.

The pod x SKU matrix shards by pod rows: rows are independent, the catalog tables are replicated, and the only
collective is ONE all-gather of the compact per-pod results (first feasible template + attempt codes, 8 bytes a pod:
`to_compact` / `from_compact` mirror b200s_resident_export_compact / _import_compact) when a single pack has to see
them all; the instance-type rows stay on the GPU that computed them -- the pack re-derives a row from the replicated
class tables when it opens a claim.  The pack itself does not shard (each placement depends on every earlier one): across
GPUs it scales over independent Solves only ("replicas only") -- e.g. the consolidation simulations of config 4,
dealt round-robin to the ranks, their few result words per simulation all-gathered (`deal`, `undeal`).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def row_shard(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of rank's rows; the first n_rows % world ranks get one row more."""
    q, r = divmod(n_rows, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def pad_rows(a: np.ndarray, rows: int) -> np.ndarray:
    """all_gather_into_tensor wants equal shards: pad the first axis with zeros up to `rows`."""
    if a.shape[0] == rows:
        return a
    pad = np.zeros((rows - a.shape[0],) + a.shape[1:], a.dtype)
    return np.concatenate([a, pad], axis=0)


def unpad_gathered(g: np.ndarray, n_rows: int, world: int) -> np.ndarray:
    """g: [world, rows_max, ...] gathered shards -> [n_rows, ...] in row order."""
    parts = []
    for rank in range(world):
        lo, hi = row_shard(n_rows, rank, world)
        parts.append(g[rank, :hi - lo])
    return np.concatenate(parts, axis=0)


def deal(n_items: int, rank: int, world: int) -> np.ndarray:
    """Round-robin share of rank: items rank, rank + world, ... (simulations are ordered cheapest first, so dealing
    keeps every rank's share representative)."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def dealt_rows(n_items: int, world: int) -> int:
    """rows every rank pads its share to for an equal-sized all-gather"""
    return (n_items + world - 1) // world


def undeal(g: np.ndarray, n_items: int, world: int) -> np.ndarray:
    """g: [world, dealt_rows, ...] gathered shares -> [n_items, ...] in item order."""
    out = np.zeros((n_items,) + g.shape[2:], g.dtype)
    for rank in range(world):
        idx = deal(n_items, rank, world)
        out[idx] = g[rank, :len(idx)]
    return out


def compact_row_bytes(T: int) -> int:
    """b200s_compact_row_bytes: [int8 template | T code bytes | padding]"""
    return 8 if T <= 7 else 24


def to_compact(tmpl: np.ndarray, code: np.ndarray) -> np.ndarray:
    """per-pod (first feasible template, attempt codes [P, T]) -> compact rows [P, compact_row_bytes(T)] uint8"""
    P, T = code.shape
    out = np.zeros((P, compact_row_bytes(T)), np.uint8)
    out[:, 0] = tmpl.astype(np.int8).view(np.uint8)
    out[:, 1:T + 1] = code
    return out


def from_compact(rows: np.ndarray, T: int):
    """compact rows -> (tmpl int32 [P], code uint8 [P, T])"""
    return rows[:, 0].view(np.int8).astype(np.int32), np.ascontiguousarray(rows[:, 1:T + 1])

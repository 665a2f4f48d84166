"""TEST INFRASTRUCTURE: a stand-in for capi.Context that needs no GPU, for exercising the host-side drivers
(b200sched/consolidation.py) in the CPU suite.  solve_batch runs the device algorithm compiled for the host
(tests/native/emul.cpp); consolidation_commands follows consolidation_kernel row by row on top of the device header's own
price-key and requirement functions (parity.emul_price_keys, emul_allowed_values).  The real kernels are checked
against the oracle in tests/test_gpu.py; this file only lets the Python around them run here."""
import ctypes as C

import numpy as np

import parity
from b200sched import capi

NONE_BITS = int(np.float64(1.7976931348623157e308).view(np.uint64))


class FakeContext:
    def __init__(self):
        self._layouts, self._next = {}, 1

    def upload_catalog(self, enc):
        self._layouts[self._next] = enc.layout_hash
        self._next += 1
        return self._next - 1

    def release_catalog(self, cid):
        self._layouts.pop(cid, None)

    def catalog_layout(self, cid):
        return self._layouts.get(cid)

    def solve_batch(self, cid, encs):
        return [parity.emul_solve(e) for e in encs]

    def solve(self, cid, enc):
        return parity.emul_solve(enc)

    def solve_batch_shared(self, cid, enc, pod_off, excluded):
        """b200s_solve_batch_shared's contract on the host: problem b = pods pod_off[b]..pod_off[b+1] of enc's model
        against its nodes without the node ROWS excluded[b]; node indices of the results count rows of enc's node set."""
        from b200sched import model as M
        m = enc.model
        E = enc.E
        node_of_row = np.array([enc.nodes.contents.node_index[r] for r in range(E)])
        row_of_node = np.zeros(max(E, 1), np.int64)
        row_of_node[node_of_row] = np.arange(E)
        out = []
        for b in range(len(pod_off) - 1):
            gone = set(int(node_of_row[r]) for r in excluded[b])
            keep = np.array([e for e in range(E) if e not in gone], np.int64)
            sub = M.subset(m, np.arange(int(pod_off[b]), int(pod_off[b + 1])), keep)
            es = capi.Encoded(sub)
            assert es.layout_hash == enc.layout_hash
            raw = parity.emul_solve(es)
            sub_rows = np.array([es.nodes.contents.node_index[r] for r in range(es.E)], np.int64)  # sub-model node per row
            on_node = raw.pod_place_kind == capi.PLACE_NODE
            raw.pod_place_index[on_node] = row_of_node[keep[sub_rows[raw.pod_place_index[on_node]]]]
            out.append(raw)
        return out

    def _allowed(self, enc, k, row):
        f = parity.emul().emul_allowed_values
        f.restype = C.c_uint64
        b = enc.bounded_key.index(k) if k in enc.bounded_key else -1
        gt = int(row.claim_req_gt[0][b]) if b >= 0 else 0
        lt = int(row.claim_req_lt[0][b]) if b >= 0 else 0
        return int(f(enc.catalog, C.c_int32(k), C.c_uint64(int(row.claim_req_mask[0][k])),
                     C.c_uint64(int(row.claim_req_flags[0])), C.c_int64(gt), C.c_int64(lt)))

    def consolidation_commands(self, cid, enc, sims, candidate_price, all_spot, n_candidates, ct_spot, ct_on_demand,
                               ct_reserved=-1, spot_to_spot=False, n_unscheduled=None):
        c = enc.catalog.contents
        n, W = len(sims), enc.W
        out = dict(decision=np.zeros(n, np.int32), reason=np.zeros(n, np.int32), n_options=np.zeros(n, np.int32),
                   mask=np.zeros((n, W), np.uint64), spot_required=np.zeros(n, np.uint8),
                   head=np.full((n, capi.S2S_MIN_TYPES), -1, np.int32), price_tie=np.zeros(n, np.uint8))
        for i, r in enumerate(sims):
            unsched = r.n_unscheduled if n_unscheduled is None else n_unscheduled[i]
            if unsched > 0:
                out["reason"][i] = capi.WHY_UNSCHEDULABLE
                continue
            if r.n_claims == 0:
                out["decision"][i] = capi.CMD_DELETE
                continue
            if r.n_claims != 1:
                out["reason"][i] = capi.WHY_MANY_CLAIMS
                continue
            za = (self._allowed(enc, c.key_zone, r) if c.key_zone >= 0 else ~0) & ((1 << c.n_zones) - 1)
            ca = self._allowed(enc, c.key_ct, r) & ((1 << c.n_cts) - 1)
            has_spot, has_od = (ca >> ct_spot) & 1, (ca >> ct_on_demand) & 1
            s2s = bool(all_spot[i]) and bool(has_spot)
            if s2s and not spot_to_spot:
                out["reason"][i] = capi.WHY_S2S_DISABLED
                continue
            cf = (1 << ct_spot) if s2s else ca
            cheapest, _ = parity.emul_price_keys(enc, za, ca, ct_reserved, ct_spot, ct_on_demand)
            _, worst = parity.emul_price_keys(enc, za, cf, ct_reserved, ct_spot, ct_on_demand)
            limit = int(np.float64(candidate_price[i]).view(np.uint64))
            keep = [s for s in parity.mask_bits(r.claim_mask[0]) if int(worst[s]) < limit]
            if not keep:
                out["reason"][i] = capi.WHY_NOT_CHEAPER
                continue
            if s2s and n_candidates[i] == 1:
                if len(keep) < capi.S2S_MIN_TYPES:
                    out["reason"][i] = capi.WHY_S2S_TOO_FEW
                    continue
                ranked = sorted(keep, key=lambda s: (int(cheapest[s]), s))
                listed = ranked[:capi.S2S_MIN_TYPES + 1]
                out["price_tie"][i] = any(cheapest[a] == cheapest[b] for a, b in zip(listed, listed[1:]))
                keep = ranked[:capi.S2S_MIN_TYPES]
                out["head"][i] = keep
            out["decision"][i] = capi.CMD_REPLACE
            out["n_options"][i] = len(keep)
            for s in keep:
                out["mask"][i, s // 64] |= np.uint64(1) << np.uint64(s % 64)
            out["spot_required"][i] = 1 if (s2s or (has_spot and has_od)) else 0
        return out

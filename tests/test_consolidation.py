"""computeConsolidation (K/controllers/disruption/consolidation.go:133-307) in the oracle, on cases small enough to
decide by hand from the Go text, plus an independent numpy restatement of the price rules on the synthetic catalog.
The device side of the same rows is in test_gpu.py."""
import numpy as np

import oracle_binding as ob
import parity
from b200sched import capi
from b200sched import model as M
from b200sched import synth

RES = ["cpu", "memory", "pods"]
GiB = 1 << 30


def _offerings(spot, od, spot_ok=True, od_ok=True, zones=("z1",)):
    out = []
    for i, z in enumerate(zones):
        out.append(([(M.KEY_ZONE, M.OP_IN, [z]), (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot"])], spot * (1 + i), spot_ok))
        out.append(([(M.KEY_ZONE, M.OP_IN, [z]), (M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand"])], od * (1 + i), od_ok))
    return out


def _builder(pool_reqs=(), a_spot_ok=True, zones=("z1",)):
    b = M.ModelBuilder(RES)
    for name, cpu, spot, od, ok in [("a", 4000, 0.1, 0.3, a_spot_ok), ("b", 8000, 0.2, 0.6, True), ("c", 16000, 0.4, 1.2, True)]:
        b.add_sku(name, [(M.KEY_ARCH, M.OP_IN, ["amd64"]), (M.KEY_ZONE, M.OP_IN, list(zones)),
                         (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot", "on-demand"])],
                  [cpu, 64 * GiB, 100], [cpu, 64 * GiB, 100], _offerings(spot, od, spot_ok=ok, zones=zones))
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])] + list(pool_reqs))
    return b


def _run(model, price, all_spot=False, n_candidates=1, gate=False):
    _, spot, od, res = parity.capacity_type_values(model)
    return ob.consolidate(model, price, all_spot, n_candidates, spot, od, res, gate)


def _names(model, cmd):
    return sorted(model.sku_names[s] for s in parity.mask_bits(cmd["mask"]))


def test_on_demand_candidate_is_replaced_by_types_whose_spot_price_is_lower():
    b = _builder()
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    out, cmd = _run(m, 0.25)
    assert out.n_claims == 1 and out.n_unscheduled == 0
    # WorstLaunchPrice takes spot before on-demand (types.go:464-468): a 0.1, b 0.2, c 0.4 against 0.25
    assert (cmd["decision"], cmd["reason"], cmd["n_options"]) == (capi.CMD_REPLACE, capi.WHY_OK, 2)
    assert _names(m, cmd) == ["a", "b"]
    assert cmd["order"].tolist() == [m.sku_names.index("a"), m.sku_names.index("b")]  # OrderByPrice order survives
    # the requirements allowed both capacity types, so the command pins spot (consolidation.go:218-221)
    assert cmd["spot_required"] == 1


def test_unavailable_spot_offering_falls_back_to_the_on_demand_price():
    b = _builder(a_spot_ok=False)
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    _, cmd = _run(m, 0.25)
    assert _names(m, cmd) == ["b"]  # a: no available spot offering -> on-demand 0.3 >= 0.25


def test_worst_launch_price_is_the_most_expensive_zone():
    b = _builder(zones=("z1", "z2"))  # z2 costs twice z1
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    _, cmd = _run(m, 0.25)
    assert _names(m, cmd) == ["a"]  # a: max(0.1, 0.2) < 0.25; b: max(0.2, 0.4) is not
    b = _builder(zones=("z1", "z2"))
    b.add_pod([(M.KEY_ZONE, M.OP_IN, ["z1"])], [2000, GiB, 1])  # the claim is pinned to z1: z2's price no longer counts
    m = b.build()
    _, cmd = _run(m, 0.25)
    assert _names(m, cmd) == ["a", "b"]


def test_on_demand_only_pool_compares_on_demand_prices_and_adds_no_requirement():
    b = _builder(pool_reqs=[(M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand"])])
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    _, cmd = _run(m, 0.7)
    assert _names(m, cmd) == ["a", "b"] and cmd["spot_required"] == 0
    _, cmd = _run(m, 0.3)  # strictly cheaper only (nodeclaim.go:273)
    assert (cmd["decision"], cmd["reason"]) == (capi.CMD_NONE, capi.WHY_NOT_CHEAPER)


def test_counts_decide_before_prices():
    b = _builder()
    b.add_pod([], [64000, GiB, 1])  # fits nothing
    _, cmd = _run(b.build(), 10.0)
    assert (cmd["decision"], cmd["reason"]) == (capi.CMD_NONE, capi.WHY_UNSCHEDULABLE)
    b = _builder()
    b.add_pod([], [9000, GiB, 1])
    b.add_pod([], [9000, GiB, 1])  # 18 cpu: two NodeClaims
    out, cmd = _run(b.build(), 10.0)
    assert out.n_claims == 2 and (cmd["decision"], cmd["reason"]) == (capi.CMD_NONE, capi.WHY_MANY_CLAIMS)
    b = _builder()
    b.add_pod([], [2000, GiB, 1])
    b.add_node("n1", {M.KEY_ZONE: "z1", M.KEY_CAPACITY_TYPE: "on-demand", M.KEY_NODEPOOL: "p", M.KEY_ARCH: "amd64"},
               [4000, 8 * GiB, 10])
    out, cmd = _run(b.build(), 10.0)
    assert out.n_claims == 0 and (cmd["decision"], cmd["n_options"]) == (capi.CMD_DELETE, 0)


def test_spot_to_spot_rules():
    b = _builder()
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    _, cmd = _run(m, 0.25, all_spot=True, gate=False)
    assert (cmd["decision"], cmd["reason"]) == (capi.CMD_NONE, capi.WHY_S2S_DISABLED)
    _, cmd = _run(m, 0.25, all_spot=True, gate=True)  # one candidate needs 15 cheaper types
    assert (cmd["decision"], cmd["reason"]) == (capi.CMD_NONE, capi.WHY_S2S_TOO_FEW)
    _, cmd = _run(m, 0.25, all_spot=True, gate=True, n_candidates=2)
    assert cmd["decision"] == capi.CMD_REPLACE and _names(m, cmd) == ["a", "b"] and cmd["spot_required"] == 1
    # a spot candidate whose replacement may only be on-demand is an ordinary consolidation
    b = _builder(pool_reqs=[(M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand"])])
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    _, cmd = _run(m, 0.7, all_spot=True, gate=False)
    assert cmd["decision"] == capi.CMD_REPLACE and cmd["spot_required"] == 0


def _price_tables(model):
    """cheapest / most expensive AVAILABLE offering per (zone value, capacity-type value, SKU), straight from the model"""
    a = model.arrays
    kz, kc = model.key_names.index(M.KEY_ZONE), model.key_names.index(M.KEY_CAPACITY_TYPE)
    lo, hi = {}, {}
    for s in range(model.n_skus):
        for o in range(a["sku_off_off"][s], a["sku_off_off"][s + 1]):
            if not a["off_available"][o]:
                continue
            z = c = None
            for i in range(a["off_req_off"][o], a["off_req_off"][o + 1]):
                key, v = int(a["reqs"][i]["key"]), int(a["req_values"][a["reqs"][i]["val_off"]])
                if key == kz:
                    z = v
                elif key == kc:
                    c = v
            p = float(a["off_price"][o])
            lo[(z, c, s)] = min(lo.get((z, c, s), np.inf), p)
            hi[(z, c, s)] = max(hi.get((z, c, s), -np.inf), p)
    return lo, hi


def test_single_candidate_spot_to_spot_keeps_the_fifteen_cheapest_in_price_order():
    b = synth.new_builder()
    synth.add_catalog(b, 80)
    synth.add_pools(b, "c2")
    b.add_pod([], [500, GiB, 0, 0, 1])
    m = b.build()
    _, spot, od, res = parity.capacity_type_values(m)
    out, cmd = ob.consolidate(m, 1.0, True, 1, spot, od, res, True)
    assert out.n_claims == 1 and cmd["decision"] == capi.CMD_REPLACE and cmd["n_options"] == 15 and cmd["spot_required"] == 1
    assert cmd["price_tie"] == 0
    # independent restatement: order by the cheapest offering (any capacity type the claim allowed), keep the types
    # whose dearest SPOT offering is below the candidate price, take 15
    lo, hi = _price_tables(m)
    options = parity.mask_bits(out.claim_mask[0])
    zones = {z for (z, _, _) in lo}
    cheapest = {s: min(lo.get((z, c, s), np.inf) for z in zones for c in (spot, od)) for s in options}
    dearest_spot = {s: max(hi.get((z, spot, s), -np.inf) for z in zones) for s in options}
    want = [s for s in sorted(options, key=lambda s: cheapest[s]) if 0 <= dearest_spot[s] < 1.0][:15]
    assert cmd["head"].tolist() == want and cmd["order"].tolist() == want
    assert sorted(parity.mask_bits(cmd["mask"])) == sorted(want)


def test_the_encoder_carries_the_most_expensive_offering_per_cell():
    """price_max feeds WorstLaunchPrice on the device (types.go:454-474): check the table against the model."""
    b = synth.new_builder()
    synth.add_catalog(b, 70)
    synth.add_pools(b, "c2")
    b.add_pod([], [500, GiB, 0, 0, 1])
    m = b.build()
    enc = capi.Encoded(m)
    c = enc.catalog.contents
    S, ZC = c.n_skus, c.n_zones * c.n_cts
    top = np.ctypeslib.as_array(c.price_max, shape=(ZC * S,)).reshape(ZC, S)
    low = np.ctypeslib.as_array(c.price, shape=(ZC * S,)).reshape(ZC, S)
    lo, hi = _price_tables(m)
    kz, kc = m.key_names.index(M.KEY_ZONE), m.key_names.index(M.KEY_CAPACITY_TYPE)
    seen = 0
    for (z, ct, s), p in hi.items():
        cell = enc.value_bit(kz, z) * c.n_cts + enc.value_bit(kc, ct)
        assert top[cell, s] == np.float64(p).view(np.uint64) and low[cell, s] == np.float64(lo[(z, ct, s)]).view(np.uint64)
        seen += 1
    assert (top != np.uint64(0xFFFFFFFFFFFFFFFF)).sum() == seen


def test_candidate_prices_read_the_offering_the_node_labels_name():
    """the synthetic feed's getCandidatePrices (consolidation.go:310-328): zone and capacity type of the node pick the
    offering of its instance type, available or not"""
    model, node_of_pod, util = synth.c4(600, 60, solve=ob.solve)
    E = int(model.c.n_nodes)
    prices, spot = synth.candidate_prices(model, range(E))
    a = model.arrays
    kz, kc, ki = (model.key_names.index(k) for k in (M.KEY_ZONE, M.KEY_CAPACITY_TYPE, M.KEY_INSTANCE_TYPE))
    for e in range(E):
        lab = {int(a["reqs"][i]["key"]): model.val_names[int(a["req_values"][a["reqs"][i]["val_off"]])]
               for i in range(a["node_req_off"][e], a["node_req_off"][e + 1])}
        s = model.sku_names.index(lab[ki])
        o = a["sku_off_off"][s] + synth.ZONES.index(lab[kz]) * 2 + synth.CAPACITY_TYPES.index(lab[kc])  # add_catalog's order
        assert prices[e] == float(a["off_price"][o]) and spot[e] == (lab[kc] == "spot")
    assert any(spot) and not all(spot)


def test_to_nodeclaim_lists_every_claims_types_in_price_order():
    """ToNodeClaim (nodeclaimtemplate.go:77-82): OrderByPrice(requirements) of each NodeClaim, cut at max_types; the
    oracle's order against a numpy restatement of the price key on the synthetic catalog."""
    m = synth.config("c3", 400, 90)
    out, order, count, tie = ob.solve_ordered(m, 25)
    assert out.n_claims > 3 and not tie.any()
    lo, _ = _price_tables(m)
    kz, kc = m.key_names.index(M.KEY_ZONE), m.key_names.index(M.KEY_CAPACITY_TYPE)
    bits = m.key_value_bits()
    zones = {z for (z, _, _) in lo}
    cts = {c for (_, c, _) in lo}
    for c in range(out.n_claims):
        # values of zone / capacity type the claim's merged requirements admit (In sets in this workload, or undefined)
        def admitted(key, universe):
            f = int(out.claim_req_flags[c, key])
            if not f & capi.RF_DEFINED:
                return set(universe)
            inside = {v for i, v in enumerate(bits[key]) if (int(out.claim_req_mask[c, key]) >> i) & 1}
            return (set(universe) - inside) if f & capi.RF_COMPLEMENT else (set(universe) & inside)
        za, ca = admitted(kz, zones), admitted(kc, cts)
        options = parity.mask_bits(out.claim_mask[c])
        price = {s: min([lo.get((z, ct, s), np.inf) for z in za for ct in ca] + [np.inf]) for s in options}
        want = sorted(options, key=lambda s: (price[s], s))[:25]
        assert count[c] == len(want) and order[c, :count[c]].tolist() == want
        assert (order[c, count[c]:] == -1).all()


# ---- a second, independent restatement of computeConsolidation in plain Python, for fuzzing the oracle ----------
def _admitted(model, out, c, key, universe):
    bits = model.key_value_bits()
    f = int(out.claim_req_flags[c, key])
    if not f & capi.RF_DEFINED:
        return set(universe)
    inside = {v for i, v in enumerate(bits[key]) if (int(out.claim_req_mask[c, key]) >> i) & 1}
    return (set(universe) - inside) if f & capi.RF_COMPLEMENT else (set(universe) & inside)


def _py_consolidate(model, out, price, all_spot, n_cand, gate):
    """consolidation.go:133-307 on an oracle Solve result, written from the Go text without looking at the C++"""
    _, spot, od, _ = parity.capacity_type_values(model)
    if out.n_unscheduled:
        return ("none", "unschedulable", [])
    if out.n_claims == 0:
        return ("delete", "", [])
    if out.n_claims != 1:
        return ("none", "many", [])
    lo, hi = _price_tables(model)
    kz, kc = model.key_names.index(M.KEY_ZONE), model.key_names.index(M.KEY_CAPACITY_TYPE)
    zones = {v for v, k in enumerate(model.arrays["val_key"]) if int(k) == kz}
    cts = {v for v, k in enumerate(model.arrays["val_key"]) if int(k) == kc}
    za, ca = _admitted(model, out, 0, kz, zones), _admitted(model, out, 0, kc, cts)
    options = parity.mask_bits(out.claim_mask[0])
    order_key = {s: min([lo.get((z, c, s), np.inf) for z in za for c in ca] + [np.inf]) for s in options}
    options = sorted(options, key=lambda s: (order_key[s], s))

    def worst(s, allowed_cts):
        for ct in (spot, od):  # reserved does not exist in these catalogs: spot comes before on-demand
            if ct in allowed_cts:
                ps = [hi[(z, ct, s)] for z in za if (z, ct, s) in hi]
                if ps:
                    return max(ps)
        return np.inf

    s2s = all_spot and spot in ca
    if s2s and not gate:
        return ("none", "s2s-disabled", [])
    allowed = {spot} if s2s else ca
    keep = [s for s in options if worst(s, allowed) < price]
    if not keep:
        return ("none", "not-cheaper", [])
    if s2s and n_cand == 1:
        if len(keep) < 15:
            return ("none", "s2s-too-few", [])
        keep = keep[:15]
    narrowed = s2s or (spot in ca and od in ca)
    if _py_consolidate.rows is not None:  # filterOutSameInstanceType, multinodeconsolidation.go:172-224
        cheapest = {}
        for s, p in _py_consolidate.rows:
            if p >= 0:
                cheapest[s] = min(cheapest.get(s, np.inf), p)
            else:
                cheapest.setdefault(s, None)
        limit = min([np.inf] + [(cheapest[s] if cheapest[s] is not None else 0.0) for s in keep if s in cheapest])
        final_cts = {spot} if narrowed else ca
        keep = [s for s in keep if worst(s, final_cts) < limit]
    return ("replace", "spot" if narrowed else "", keep)


_py_consolidate.rows = None


def test_fuzzed_catalogs_agree_with_the_python_restatement():
    rng = np.random.default_rng(2024)
    names = {capi.CMD_NONE: "none", capi.CMD_DELETE: "delete", capi.CMD_REPLACE: "replace"}
    why = {capi.WHY_UNSCHEDULABLE: "unschedulable", capi.WHY_MANY_CLAIMS: "many", capi.WHY_NOT_CHEAPER: "not-cheaper",
           capi.WHY_S2S_DISABLED: "s2s-disabled", capi.WHY_S2S_TOO_FEW: "s2s-too-few", capi.WHY_OK: ""}
    seen = set()
    filtered = emptied = 0
    for trial in range(160):
        n_sku = int(rng.integers(3, 40))
        zones = ["z1", "z2", "z3"][:int(rng.integers(1, 4))]
        b = M.ModelBuilder(RES)
        prices = rng.permutation(n_sku * len(zones) * 2) + 1  # distinct prices: no sort ties
        pi = 0
        for s in range(n_sku):
            offs = []
            for z in zones:
                for ct in ("spot", "on-demand"):
                    if rng.random() < 0.85:
                        offs.append(([(M.KEY_ZONE, M.OP_IN, [z]), (M.KEY_CAPACITY_TYPE, M.OP_IN, [ct])],
                                     float(prices[pi]) / 100.0, bool(rng.random() < 0.8)))  # all distinct
                    pi += 1
            cpu = int(rng.choice([2000, 4000, 8000, 16000]))
            b.add_sku(f"s{s:02d}", [(M.KEY_ARCH, M.OP_IN, ["amd64"]), (M.KEY_ZONE, M.OP_IN, zones),
                                    (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot", "on-demand"])],
                      [cpu, 64 * GiB, 100], [cpu, 64 * GiB, 100], offs)
        pool = [[], [(M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand"])], [(M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot"])],
                [(M.KEY_CAPACITY_TYPE, M.OP_NOTIN, ["spot"])], [(M.KEY_CAPACITY_TYPE, M.OP_EXISTS, [])],
                [(M.KEY_ZONE, M.OP_NOTIN, [zones[0]])]][int(rng.integers(0, 6))]
        b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])] + pool)
        for _ in range(int(rng.integers(1, 4))):
            reqs = [(M.KEY_ZONE, M.OP_IN, [str(rng.choice(zones))])] if rng.random() < 0.3 else []
            b.add_pod(reqs, [int(rng.choice([500, 1500, 3000, 9000, 20000])), GiB, 1])
        if rng.random() < 0.2:  # a node with room: some simulations need no new NodeClaim at all
            b.add_node("n1", {M.KEY_ZONE: zones[0], M.KEY_CAPACITY_TYPE: "on-demand", M.KEY_NODEPOOL: "p",
                              M.KEY_ARCH: "amd64"}, [int(rng.choice([4000, 64000])), 64 * GiB, 50])
        m = b.build()
        _, spot, od, res = parity.capacity_type_values(m)
        if spot < 0 or od < 0:
            continue
        for price, all_spot, n_cand, gate in ((float(rng.choice([0.01, 0.5, 2.0, 50.0])), bool(rng.random() < 0.5),
                                               int(rng.integers(1, 3)), bool(rng.random() < 0.6)) for _ in range(3)):
            rows = None
            if n_cand == 2 and rng.random() < 0.7:  # a multi-node probe: the candidates' types and prices
                rows = [(int(rng.integers(0, n_sku)), float(rng.choice([-1.0, 0.02, 0.2, 1.0, 5.0]))) for _ in range(2)]
            out, cmd = ob.consolidate(m, price, all_spot, n_cand, spot, od, res, gate, same_type_rows=rows)
            _py_consolidate.rows = rows
            want = _py_consolidate(m, out, price, all_spot, n_cand, gate)
            _py_consolidate.rows = None
            got = (names[cmd["decision"]], why[cmd["reason"]] if cmd["decision"] == capi.CMD_NONE else
                   ("spot" if cmd["spot_required"] else ""), cmd["order"].tolist() if cmd["decision"] == capi.CMD_REPLACE else [])
            assert got == want, (trial, price, all_spot, n_cand, gate, rows, got, want)
            if rows is not None:  # validDecision (multinodeconsolidation.go:139-150)
                assert bool(cmd["valid"]) == (want[0] == "delete" or (want[0] == "replace" and len(want[2]) > 0))
                filtered += want[0] == "replace" and len(want[2]) < len(_py_consolidate(m, out, price, all_spot, n_cand, gate)[2])
                emptied += want[0] == "replace" and not want[2]
            seen.add(want[:2])
    assert {("delete", ""), ("replace", "spot"), ("replace", ""), ("none", "many"), ("none", "not-cheaper"),
            ("none", "s2s-disabled"), ("none", "unschedulable")} <= seen, seen
    assert filtered > 3 and emptied > 0, (filtered, emptied)


def test_device_price_keys_match_the_model():
    """cheapest_price / worst_launch_price of b200_core.cuh (what the consolidation and price-order kernels evaluate
    per instance type), compiled for the host, against the offerings of the model for every subset of zones and
    capacity types."""
    b = synth.new_builder()
    synth.add_catalog(b, 90)
    synth.add_pools(b, "c2")
    b.add_pod([], [500, GiB, 0, 0, 1])
    m = b.build()
    enc = capi.Encoded(m)
    lo, hi = _price_tables(m)
    key, spot, od, _ = parity.capacity_type_values(m)
    kz = m.key_names.index(M.KEY_ZONE)
    zones = sorted({z for (z, _, _) in lo})
    zbit = {z: enc.value_bit(kz, z) for z in zones}
    cbit = {c: enc.value_bit(key, c) for c in (spot, od)}
    none = np.float64(1.7976931348623157e308).view(np.uint64)
    for zmask in range(1, 1 << len(zones)):
        zs = [z for i, z in enumerate(zones) if (zmask >> i) & 1]
        za = sum(1 << zbit[z] for z in zs)
        for cs in ([spot], [od], [spot, od]):
            ca = sum(1 << cbit[c] for c in cs)
            cheapest, worst = parity.emul_price_keys(enc, za, ca, -1, cbit[spot], cbit[od])
            for s in range(m.n_skus):
                want_lo = min([lo[(z, c, s)] for z in zs for c in cs if (z, c, s) in lo] + [np.inf])
                assert cheapest[s] == (np.float64(want_lo).view(np.uint64) if np.isfinite(want_lo) else none)
                want_hi = np.inf
                for c in (spot, od):  # reserved, spot, on-demand: the first capacity type with an offering decides
                    ps = [hi[(z, c, s)] for z in zs if c in cs and (z, c, s) in hi]
                    if ps:
                        want_hi = max(ps)
                        break
                assert worst[s] == (np.float64(want_hi).view(np.uint64) if np.isfinite(want_hi) else none), (zs, cs, s)


def test_several_offerings_in_one_zone_and_capacity_type():
    """Nothing forbids two offerings with the same zone and capacity type: the cheapest available one prices the type
    (OrderByPrice), the dearest available one is its worst launch price; unavailable ones count for neither."""
    b = M.ModelBuilder(RES)
    cell = lambda ct, price, ok: ([(M.KEY_ZONE, M.OP_IN, ["z1"]), (M.KEY_CAPACITY_TYPE, M.OP_IN, [ct])], price, ok)  # noqa: E731
    b.add_sku("a", [(M.KEY_ARCH, M.OP_IN, ["amd64"]), (M.KEY_ZONE, M.OP_IN, ["z1"]),
                    (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot", "on-demand"])], [4000, 64 * GiB, 100], [4000, 64 * GiB, 100],
              [cell("spot", 0.30, True), cell("spot", 0.10, True), cell("spot", 0.90, False), cell("on-demand", 0.50, True)])
    b.add_sku("b", [(M.KEY_ARCH, M.OP_IN, ["amd64"]), (M.KEY_ZONE, M.OP_IN, ["z1"]),
                    (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot", "on-demand"])], [8000, 64 * GiB, 100], [8000, 64 * GiB, 100],
              [cell("spot", 0.20, True), cell("on-demand", 0.60, True)])
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    # the oracle: a's worst spot launch price is 0.30 (not 0.10, not the unavailable 0.90)
    _, cmd = _run(m, 0.25)
    assert _names(m, cmd) == ["b"]
    _, cmd = _run(m, 0.31)
    assert _names(m, cmd) == ["a", "b"] and cmd["order"].tolist() == [0, 1]  # ordered by the cheapest: a 0.10, b 0.20
    # the encoding the device reads
    enc = capi.Encoded(m)
    key, spot, od, _ = parity.capacity_type_values(m)
    bs, bo = enc.value_bit(key, spot), enc.value_bit(key, od)
    cheapest, worst = parity.emul_price_keys(enc, 1, (1 << bs) | (1 << bo), -1, bs, bo)
    f = lambda x: np.float64(x).view(np.uint64)  # noqa: E731
    assert cheapest.tolist() == [f(0.10), f(0.20)] and worst.tolist() == [f(0.30), f(0.20)]
    cheapest, worst = parity.emul_price_keys(enc, 1, 1 << bo, -1, bs, bo)
    assert cheapest.tolist() == [f(0.50), f(0.60)] and worst.tolist() == [f(0.50), f(0.60)]


def test_golden_digests_of_the_widened_rows():
    """tests/golden/widened_digests.json (written by tests/golden/make_golden.py from OUR oracle: a drift guard, not a
    reference output): ToNodeClaim lists on a config-3-shaped Solve and every probe of a multi-node search."""
    import hashlib
    import json
    import os

    import consolidation_ref as ref
    from b200sched import consolidation as cons
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "widened_digests.json")))

    def sha(*arrays):
        h = hashlib.sha256()
        for a in arrays:
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()

    g = gold["to_nodeclaim_c3_1500x150"]
    o, order, count, tie = ob.solve_ordered(synth.config("c3", 1500, 150), 600)
    assert (int(o.n_claims), int(count.sum()), int(tie.sum())) == (g["n_claims"], g["listed"], g["ties"])
    assert sha(order.astype(np.int32), count.astype(np.int32)) == g["digest"]

    g = gold["multi_node_search_c4_1500x120_first40"]
    model, node_of_pod, util = synth.c4(1500, 120, solve=ob.solve)
    sub, nop = ref.sub_cluster(model, node_of_pod, np.argsort(util, kind="stable")[:40])
    cand = list(range(40))
    ctv = parity.capacity_type_values(sub)
    facts = cons.node_facts(sub, cand)
    best, _ = ref.first_n_consolidation_option(sub, nop, cand, ctv)
    assert (None if best is None else best + 1) == g["removes_first"]
    for mid, decision, reason, valid, n_options, digest in g["probes"]:
        _, cmd = ref.probe(sub, nop, cand, facts, mid, ctv, False)
        assert [int(cmd["decision"]), int(cmd["reason"]), int(cmd["valid"]), int(cmd["n_options"]),
                sha(cmd["mask"].astype(np.uint64))[:16]] == [decision, reason, valid, n_options, digest], mid


def test_reserved_offerings_come_first_in_the_worst_launch_price():
    """types.go:464-468: reserved, then spot, then on-demand.  A type with a compatible reserved offering is priced by
    it alone, however cheap its spot offering is."""
    b = M.ModelBuilder(RES)
    off = lambda ct, price: ([(M.KEY_ZONE, M.OP_IN, ["z1"]), (M.KEY_CAPACITY_TYPE, M.OP_IN, [ct])], price, True)  # noqa: E731
    for name, cpu, offerings in (("a", 4000, [off("spot", 0.10), off("on-demand", 0.30)]),
                                 ("b", 8000, [off("reserved", 0.40), off("spot", 0.05), off("on-demand", 0.60)]),
                                 ("c", 16000, [off("reserved", 0.01), off("on-demand", 1.20)])):
        b.add_sku(name, [(M.KEY_ARCH, M.OP_IN, ["amd64"]), (M.KEY_ZONE, M.OP_IN, ["z1"]),
                         (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot", "on-demand", "reserved"])],
                  [cpu, 64 * GiB, 100], [cpu, 64 * GiB, 100], offerings)
    b.add_pool("p", [(M.KEY_NODEPOOL, M.OP_IN, ["p"])])
    b.add_pod([], [2000, GiB, 1])
    m = b.build()
    key, spot, od, res = parity.capacity_type_values(m)
    assert res >= 0
    out, cmd = ob.consolidate(m, 0.25, False, 1, spot, od, res, False)
    assert _names(m, cmd) == ["a", "c"]  # b: reserved 0.40 decides, its 0.05 spot offering does not
    # without telling the oracle which value is "reserved" the precedence is gone: b is then priced by spot
    _, cmd = ob.consolidate(m, 0.25, False, 1, spot, od, -1, False)
    assert _names(m, cmd) == ["a", "b"]  # and c has neither spot nor a cheap on-demand offering
    enc = capi.Encoded(m)
    bs, bo, br = (enc.value_bit(key, v) for v in (spot, od, res))
    f = lambda x: np.float64(x).view(np.uint64)  # noqa: E731
    _, worst = parity.emul_price_keys(enc, 1, (1 << bs) | (1 << bo) | (1 << br), br, bs, bo)
    assert worst.tolist() == [f(0.10), f(0.40), f(0.01)]
    _, worst = parity.emul_price_keys(enc, 1, (1 << bs) | (1 << bo), br, bs, bo)  # requirements exclude reserved
    assert worst.tolist() == [f(0.10), f(0.05), f(1.20)]


def test_multi_node_search_driver_on_the_host_stand_in():
    """b200sched/consolidation.py (all prefixes as one batch, filterOutSameInstanceType as a second command pass, the
    binary search replayed) driven through tests/fake_device.py -- the device algorithm compiled for the host -- must
    end where the reference's sequential search over oracle simulations ends, with the same command for every prefix.
    The same check against the real kernels is tests/test_gpu.py::test_multi_node_consolidation_search_as_one_batch."""
    import consolidation_ref as ref
    import fake_device
    from b200sched import consolidation as cons
    model, node_of_pod, util = synth.c4(1500, 120, solve=ob.solve)
    order = np.argsort(util, kind="stable")
    ctx = fake_device.FakeContext()
    replaced = 0
    for take, off in ((24, 0), (30, 40), (2, 0), (1, 0)):
        sub, nop = ref.sub_cluster(model, node_of_pod, order[off:off + take])
        cand = list(range(take))
        ctv = parity.capacity_type_values(sub)
        want_mid, _ = ref.first_n_consolidation_option(sub, nop, cand, ctv)
        got = cons.first_n_consolidation_option(ctx, 0, sub, nop, cand, ctv)
        assert (got.mid if got is not None else None) == want_mid
        # the same search with every prefix in ONE model (b200s_solve_batch_shared's input): same end, same command
        one = cons.first_n_consolidation_option(ctx, 0, sub, nop, cand, ctv, shared=True)
        assert (one.mid if one is not None else None) == want_mid
        if got is not None:
            assert (one.decision, one.reason, one.valid, one.n_options, one.mask.tolist(), one.spot_required) == (
                got.decision, got.reason, got.valid, got.n_options, got.mask.tolist(), got.spot_required)
            assert (one.result.n_claims, one.result.n_unscheduled) == (got.result.n_claims, got.result.n_unscheduled)
        if take < 2:
            continue
        lo_, hi_ = cons.search_bounds(take, take)
        facts = cons.node_facts(sub, cand)
        separate = cons.multi_node_probes(ctx, 0, sub, nop, cand, range(lo_, hi_ + 1), ctv)
        for p, q in zip(separate, cons.multi_node_probes(ctx, 0, sub, nop, cand, range(lo_, hi_ + 1), ctv, shared=True)):
            assert (p.mid, p.decision, p.reason, p.valid, p.n_options, p.mask.tolist()) == (
                q.mid, q.decision, q.reason, q.valid, q.n_options, q.mask.tolist()), (take, off, p.mid)
        for p in separate:
            _, cmd = ref.probe(sub, nop, cand, facts, p.mid, ctv, False)
            assert (p.decision, p.reason, p.valid, p.n_options, p.mask.tolist()) == (
                cmd["decision"], cmd["reason"], bool(cmd["valid"]), cmd["n_options"], cmd["mask"].tolist()), (take, off, p.mid)
            replaced += p.decision == capi.CMD_REPLACE
    assert replaced >= 3


def test_to_nodeclaim_lists_on_fuzzed_models():
    """The oracle's ToNodeClaim lists on 100 random models (NotIn / Exists / DoesNotExist zone and capacity-type
    requirements on pools and pods, unavailable offerings, existing nodes) against the numpy price key."""
    import fuzz_models
    checked = 0
    for seed in range(100):
        m = fuzz_models.random_model(1000 + seed)
        out, order, count, tie = ob.solve_ordered(m, 9)
        if out.n_claims == 0:
            continue
        lo, _ = _price_tables(m)
        kz, kc = m.key_names.index(M.KEY_ZONE), m.key_names.index(M.KEY_CAPACITY_TYPE)
        zones = {v for v, k in enumerate(m.arrays["val_key"]) if int(k) == kz}
        cts = {v for v, k in enumerate(m.arrays["val_key"]) if int(k) == kc}
        for c in range(out.n_claims):
            za, ca = _admitted(m, out, c, kz, zones), _admitted(m, out, c, kc, cts)
            options = parity.mask_bits(out.claim_mask[c])
            price = {s: min([lo.get((z, ct, s), np.inf) for z in za for ct in ca] + [np.inf]) for s in options}
            ranked = sorted(options, key=lambda s: (price[s], s))
            cut = ranked[:9]
            has_tie = any(price[a] == price[b] for a, b in zip(cut, ranked[1:10]))
            assert bool(tie[c]) == has_tie, (seed, c)
            if not has_tie:
                assert order[c, :count[c]].tolist() == cut, (seed, c)
                assert int(out.claim_sku[c]) == cut[0]
            checked += 1
    assert checked > 100

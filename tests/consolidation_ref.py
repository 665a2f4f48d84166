"""TEST INFRASTRUCTURE: firstNConsolidationOption (K/controllers/disruption/multinodeconsolidation.go:116-170) the way
the reference runs it -- a sequential binary search, one oracle simulation per probe -- to check the batched device
driver (b200sched/consolidation.py) against."""
import numpy as np

import oracle_binding as ob
from b200sched import consolidation as cons
from b200sched import model as M


def probe(model, node_of_pod, candidates, facts, mid, ct_values, spot_to_spot):
    _, v_spot, v_od, v_res = ct_values
    sim = cons.prefix_simulation(model, node_of_pod, candidates, mid)
    price, _ = cons.prefix_price(facts, mid)
    return ob.consolidate(sim, price, all(f[2] for f in facts[:mid + 1]), mid + 1, v_spot, v_od,
                          v_res, spot_to_spot, same_type_rows=[(s, p) for s, p, _, _ in facts[:mid + 1]])


def first_n_consolidation_option(model, node_of_pod, candidates, ct_values, max_parallel=cons.MAX_PARALLEL,
                                 spot_to_spot=False):
    """returns (mid of lastSavedCommand or None, {mid: (SolveOut, command)} for the probes the search made)"""
    if len(candidates) < 2:
        return None, {}
    facts = cons.node_facts(model, candidates)
    lo, hi = cons.search_bounds(len(candidates), min(max_parallel, len(candidates)))
    seen = {}

    def valid_of(mid):
        seen[mid] = probe(model, node_of_pod, candidates, facts, mid, ct_values, spot_to_spot)
        return bool(seen[mid][1]["valid"])

    return cons.replay_search(lo, hi, valid_of), seen


def sub_cluster(model, node_of_pod, nodes):
    """the cluster made of `nodes` only (and the pods on them): (model, node_of_pod) re-indexed"""
    nodes = np.asarray(nodes, np.int64)
    pods = np.nonzero(np.isin(node_of_pod, nodes))[0]
    new_index = {int(e): i for i, e in enumerate(nodes)}
    return M.subset(model, pods, nodes), np.array([new_index[int(e)] for e in node_of_pod[pods]], np.int64)

"""The host-side flattening: go/pkg/b200sched/flatten.go and go/pkg/providers/instancetype cannot be
compiled in the authoring image, so their rules are restated in Python (b200sched/flatten.py, azure_catalog.py) against
the same reference lines and checked here — hand-worked NewPodRequirements cases, request vectors against the library's
own helper, and a whole batch of Kubernetes-shaped pods over the Azure GPU catalog through encode + the oracle + the
device algorithm's host twin."""
import numpy as np
import pytest

import oracle_binding as ob
import parity
from b200sched import azure_catalog as az
from b200sched import capi, flatten
from b200sched import model as M


def _pod(i, spec):
    return {"metadata": {"uid": f"{i:08x}-0000-4000-8000-{i * 7919:012x}", "creationTimestamp": 1_700_000_000 + i % 3}, "spec": spec}


def test_new_pod_requirements_hand_worked():
    """requirements.go:74-110: nodeSelector, then the heaviest preference, then the FIRST required term"""
    zone = "topology.kubernetes.io/zone"
    spec = {"nodeSelector": {zone: "eastus-1", "team": "a"},
            "affinity": {"nodeAffinity": {
                "preferredDuringSchedulingIgnoredDuringExecution": [
                    {"weight": 10, "preference": {"matchExpressions": [{"key": "disk", "operator": "In", "values": ["ssd"]}]}},
                    {"weight": 50, "preference": {"matchExpressions": [{"key": "gen", "operator": "Gt", "values": ["3"]}]}},
                    {"weight": 50, "preference": {"matchExpressions": [{"key": "never", "operator": "Exists"}]}}],
                "requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
                    {"matchExpressions": [{"key": zone, "operator": "NotIn", "values": ["eastus-3"]},
                                          {"key": "gpu", "operator": "DoesNotExist"}]},
                    {"matchExpressions": [{"key": "ignored", "operator": "Exists"}]}]}}}}
    rows = flatten.pod_requirement_rows(spec)
    assert rows == [("team", M.OP_IN, ["a"]), (zone, M.OP_IN, ["eastus-1"]),   # NewLabelRequirements, key order
                    ("gen", M.OP_GT, 3),                                         # weight 50, the first of the two
                    (zone, M.OP_NOTIN, ["eastus-3"]), ("gpu", M.OP_DOESNOTEXIST, None)]
    assert flatten.relaxable(spec)
    assert not flatten.relaxable({"nodeSelector": {"a": "b"}})
    assert flatten.pod_requirement_rows({}) == []
    # one required term, no preference: nothing for Preferences.Relax to remove
    one = {"affinity": {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
        {"matchExpressions": [{"key": "a", "operator": "Lt", "values": ["9"]}]}]}}}}
    assert flatten.pod_requirement_rows(one) == [("a", M.OP_LT, 9)] and not flatten.relaxable(one)
    with pytest.raises(flatten.Fallback):
        flatten.pod_requirement_rows({"affinity": {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {
            "nodeSelectorTerms": [{"matchExpressions": [{"key": "a", "operator": "Gt", "values": ["x"]}]}]}}}})


def test_what_sends_a_batch_to_the_cpu_path():
    for spec in ({"affinity": {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [{}]}}},
                 {"topologySpreadConstraints": [{"maxSkew": 1}]},
                 {"volumes": [{"persistentVolumeClaim": {"claimName": "x"}}]},
                 {"containers": [{"ports": [{"hostPort": 80}]}]},
                 {"resourceClaims": [{"name": "gpu"}]}):
        with pytest.raises(flatten.Fallback):
            flatten.plain_pod(spec)
    flatten.plain_pod({"containers": [{"ports": [{"containerPort": 80}]}], "volumes": [{"emptyDir": {}}]})
    with pytest.raises(flatten.Fallback):
        flatten.uid_words("not-a-uuid")
    assert flatten.uid_words("00000000-0000-0000-0000-000000000001") == (0, 1)
    a, b = "0f000000-0000-4000-8000-000000000000", "10000000-0000-4000-8000-000000000000"
    assert (a < b) == (flatten.uid_words(a) < flatten.uid_words(b))  # queue.go:105-107 compares the strings


def test_request_vectors_match_the_library_helper():
    """flatten.pod_requests (the Go shim calls resources.RequestsForPods) against b200s_pod_requests on the same pods"""
    import test_host_helpers as th
    res = ["cpu", "memory", "hugepages-2Mi", "pods"]
    spec = {"containers": [{"resources": {"requests": {"cpu": "500m", "memory": "100"}}},
                           {"resources": {"requests": {"cpu": "250m"}, "limits": {"cpu": "999m", "memory": "64"}}}],
            "initContainers": [{"resources": {"requests": {"cpu": "100m"}}, "restartPolicy": "Always"},
                               {"resources": {"requests": {"cpu": "2"}}}],
            "overhead": {"cpu": "10m"}}
    want = th.pod_requests([{"req": {0: 500, 1: 100}}, {"req": {0: 250}, "lim": {0: 999, 1: 64}}],
                           [{"req": {0: 100}, "always": True}, {"req": {0: 2000}}], [10, 0, 0, 0])
    assert flatten.pod_requests(spec, res) == want
    assert flatten.quantity("memory", "2Gi") == 2 << 30 and flatten.quantity("cpu", "1.5") == 1500
    with pytest.raises(flatten.Fallback):
        flatten.quantity("cpu", "0.0001")


def _azure_batch(n_pods=400):
    b = M.ModelBuilder(az.RES)
    for k in (M.KEY_INSTANCE_TYPE, M.KEY_ZONE, M.KEY_CAPACITY_TYPE, M.KEY_ARCH, M.KEY_OS, M.KEY_NODEPOOL):
        b.key(k)
    names = az.add_catalog(b, "eastus")
    b.add_pool("gpu", [(M.KEY_NODEPOOL, M.OP_IN, ["gpu"]), (M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand"])],
               taints=[("sku", "gpu", M.EFFECT_NOSCHEDULE)], weight=10)
    b.add_pool("spot", [(M.KEY_NODEPOOL, M.OP_IN, ["spot"]), (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot"]),
                        (az.KEY_FAMILY, M.OP_IN, ["NC"])], taints=[("sku", "gpu", M.EFFECT_NOSCHEDULE)])
    rng = np.random.default_rng(5)
    pods = []
    for i in range(n_pods):
        gpus = int(rng.choice([1, 1, 2, 4]))
        spec = {"containers": [{"resources": {"requests": {"cpu": f"{int(rng.choice([500, 1000, 2000, 4000]))}m",
                                                           "memory": f"{int(rng.choice([2, 8, 16, 48]))}Gi",
                                                           "nvidia.com/gpu": gpus}}}],
                "tolerations": [{"key": "sku", "operator": "Equal", "value": "gpu", "effect": "NoSchedule"}]}
        u = rng.random()
        if u < 0.25:
            spec["nodeSelector"] = {az.KEY_GPU_NAME: str(rng.choice(["V100", "T4", "K80"]))}
        elif u < 0.4:
            spec["affinity"] = {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
                {"matchExpressions": [{"key": az.KEY_GPU_COUNT, "operator": "Gt", "values": ["1"]},
                                      {"key": M.KEY_ZONE, "operator": "NotIn", "values": ["eastus-2"]}]}]}}}
        elif u < 0.45:
            spec["tolerations"] = []          # does not tolerate the pools' taint: stays in PodErrors
        pods.append(_pod(i, spec))
    flags = flatten.add_pods(b, pods)
    return b.build(), names, pods, flags


def test_azure_catalog_serves_every_sku_of_the_reference_examples():
    """N3 / a18: GetInstanceTypes has something to return — the 23 distinct GPU SKU names of the reference's
    examples/v1-nodeclaim-gpu.yaml:22-48, each with every well-known label, 6 offerings, capacity > allocatable"""
    from b200sched import synth
    m, names, _, _ = _azure_batch(1)
    assert names == synth.REAL_GPU_SKUS and len(names) == 23
    a = m.arrays
    assert (a["sku_alloc"][:, 0] < a["sku_capacity"][:, 0]).all() and (a["sku_alloc"][:, 1] < a["sku_capacity"][:, 1]).all()
    assert (np.diff(a["sku_off_off"]) == 6).all()
    enc = capi.Encoded(m)                       # representable: no key has more than 64 values, prices are distinct
    assert enc.S == 23
    assert az.cheapest(["Standard_NC24s_v3", "Standard_NC6s_v3", "Standard_NC12s_v3"]) == "Standard_NC6s_v3"
    rc, chosen, disk, gpu = capi.create_select("ws0123456789", ["Standard_NC24s_v3", "Standard_NC6s_v3"], 120 << 30,
                                               [12.24, 3.06])
    assert (rc, chosen, gpu) == (0, 1, 1)       # the price-ordered Create (go/pkg/providers/instance/select.go)


def test_kubernetes_shaped_batch_over_the_azure_catalog_end_to_end():
    """flatten -> b200s_encode -> (oracle | host twin of the device algorithm): identical NodeClaims, and the selected SKU
    of every NodeClaim is the cheapest instance type that holds its pods under the pool's capacity type"""
    m, names, pods, flags = _azure_batch(400)
    enc = capi.Encoded(m)
    o = ob.solve(m)
    raw = parity.emul_solve(enc)
    parity.assert_same(parity.canon_device(enc, raw), parity.canon_oracle(m, o), "azure batch")
    assert o.n_claims > 5 and 0 < o.n_unscheduled < 60 and o.counters["price_ties"] == 0
    assert not any(flags)
    t = {s["name"]: s for s in az.load()["skus"]}
    alloc = m.arrays["sku_alloc"]
    for c in range(o.n_claims):
        fits = [s for s in range(len(names)) if (alloc[s] >= o.claim_requests[c]).all()]
        assert o.claim_sku[c] in fits
        pool_ct = "on-demand" if m.pool_names[int(o.claim_template[c])] == "gpu" else "spot"
        opts = parity.mask_bits(o.claim_mask[c])
        assert names[o.claim_sku[c]] == az.cheapest([names[s] for s in opts], pool_ct)

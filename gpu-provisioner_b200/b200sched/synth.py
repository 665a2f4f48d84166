"""Deterministic synthetic inputs for the BASELINE.json configs.

splitmix64 streams, fixed seeds (catalog 0xB2000001, pods 0xB2000002).  Everything the generator
leaves out is listed in SURVEY.md Appendix A.12 (no PVCs, host ports, pod (anti-)affinity, topology
spread, DRA, daemonsets, minValues, reserved offerings, PreferNoSchedule taints) so the scheduler
branches that are not kernelised stay inert.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import model as M

SEED_CATALOG = 0xB2000001
SEED_PODS = 0xB2000002
SEED_NODES = 0xB2000003
_GOLD = np.uint64(0x9E3779B97F4A7C15)

KEY_FAMILY = "karpenter.azure.com/sku-family"
KEY_SKU_CPU = "karpenter.azure.com/sku-cpu"
KEY_SKU_GPU = "karpenter.azure.com/sku-gpu-count"
KEY_TEAM = "example.com/team"
# provider labels are registered as well-known by the cloud provider at start-up
PROVIDER_WELLKNOWN = (KEY_FAMILY, KEY_SKU_CPU, KEY_SKU_GPU)

ZONES = ["eastus-1", "eastus-2", "eastus-3"]
CAPACITY_TYPES = ["on-demand", "spot"]
GiB = 1 << 30
MiB = 1 << 20

# the 26 GPU SKU names of the reference's examples/v1-nodeclaim-gpu.yaml:22-48 (deduplicated, in file order)
REAL_GPU_SKUS = [
    "Standard_NC6", "Standard_NC12", "Standard_NC24", "Standard_NC24r", "Standard_NC6s_v2", "Standard_NC12s_v2",
    "Standard_NC24s_v2", "Standard_NC24rs_v2", "Standard_NC6s_v3", "Standard_NC12s_v3", "Standard_NC24s_v3",
    "Standard_NC24rs_v3", "Standard_NC16as_T4_v3", "Standard_NC4as_T4_v3", "Standard_NC64as_T4_v3",
    "Standard_NC8as_T4_v3", "Standard_ND6s", "Standard_ND12s", "Standard_ND24s", "Standard_ND24rs", "Standard_NV6",
    "Standard_NV12", "Standard_NV24",
]


def sm64(seed: int, stream: int, n: int) -> np.ndarray:
    """n splitmix64 outputs of stream `stream` under `seed`."""
    with np.errstate(over="ignore"):
        s = np.uint64((seed * 0x100000001B3 + stream * 0xD6E8FEB86659FD93) & ((1 << 64) - 1))
        z = s + (np.arange(1, n + 1, dtype=np.uint64) * _GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, stream: int, n: int) -> np.ndarray:
    return (sm64(seed, stream, n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def choice(seed: int, stream: int, n: int, weights: Sequence[float]) -> np.ndarray:
    w = np.asarray(weights, dtype=np.float64)
    cdf = np.cumsum(w / w.sum())
    cdf[-1] = 1.0
    return np.searchsorted(cdf, uniform(seed, stream, n), side="right").astype(np.int64)


RES = ["cpu", "memory", "ephemeral-storage", "nvidia.com/gpu", "pods"]
VCPUS = [2, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128]
MEM_RATIO = [2, 4, 8, 16]
EPH = [64, 128, 256, 512, 1024, 2048]
PODS_CAP = [30, 110, 250]
FAMILIES_CPU = ["D", "E", "F"]
FAMILIES_GPU = ["NC", "ND", "NV"]


def add_catalog(b: M.ModelBuilder, n_skus: int, seed: int = SEED_CATALOG, with_instance_type_key: bool = True) -> None:
    """S Azure-shaped SKU rows: capacity, overhead -> allocatable, labels, 3 zones x 2 capacity types."""
    u = lambda stream: uniform(seed, stream, n_skus)  # noqa: E731
    vc = choice(seed, 1, n_skus, [8, 10, 4, 10, 4, 8, 3, 6, 3, 3, 2, 1])
    mr = choice(seed, 2, n_skus, [2, 4, 3, 1])
    ep = choice(seed, 3, n_skus, [2, 3, 3, 2, 1, 1])
    pc = choice(seed, 4, n_skus, [2, 5, 2])
    isgpu = u(5) < 0.35
    gc = choice(seed, 6, n_skus, [4, 3, 2, 2])  # index into [1,2,4,8]
    arm = u(7) < 0.10
    fam_c = choice(seed, 8, n_skus, [3, 2, 1])
    fam_g = choice(seed, 9, n_skus, [3, 2, 1])
    ver = choice(seed, 10, n_skus, [1, 2, 3, 2])
    avail = uniform(seed, 11, n_skus * 6).reshape(n_skus, 6) < 0.9
    names = set()
    for s in range(n_skus):
        vcpu = VCPUS[vc[s]]
        mem = vcpu * MEM_RATIO[mr[s]] * GiB
        gpu = [1, 2, 4, 8][gc[s]] if (isgpu[s] or s < len(REAL_GPU_SKUS)) else 0
        fam = FAMILIES_GPU[fam_g[s]] if gpu else FAMILIES_CPU[fam_c[s]]
        arch = "arm64" if (arm[s] and not gpu) else "amd64"
        if s < len(REAL_GPU_SKUS):
            name = REAL_GPU_SKUS[s]
            fam = name.split("_")[1][:2]
        else:
            name = f"Standard_{fam}{vcpu}{'p' if arch == 'arm64' else ''}s_v{ver[s] + 2}"
            i = 0
            while name in names:
                i += 1
                name = f"Standard_{fam}{vcpu}{'p' if arch == 'arm64' else ''}s{i}_v{ver[s] + 2}"
        names.add(name)
        cap = [vcpu * 1000, mem, EPH[ep[s]] * GiB, gpu, PODS_CAP[pc[s]]]
        # kubeReserved + evictionThreshold (types.go:181-183): cpu 100m + 10m/core over 2, memory 2 GiB-ish table + 100Mi
        kr_cpu = 100 + 10 * max(0, vcpu - 2)
        kr_mem = (2 * GiB if mem <= 64 * GiB else 4 * GiB if mem <= 512 * GiB else 8 * GiB) + 100 * MiB
        kr_eph = 10 * GiB
        alloc = [cap[0] - kr_cpu, cap[1] - kr_mem, cap[2] - kr_eph, cap[3], cap[4]]
        base = vcpu * 0.04 + (mem / GiB) * 0.005 + gpu * 0.9
        offerings = []
        zones_av = []
        cts_av = []
        for zi, z in enumerate(ZONES):
            for ci, ct in enumerate(CAPACITY_TYPES):
                o = zi * 2 + ci
                price = base * [1.0, 1.02, 1.05][zi] * [1.0, 0.33][ci] + s * 1e-7 + o * 1e-9
                offerings.append(([(M.KEY_ZONE, M.OP_IN, [z]), (M.KEY_CAPACITY_TYPE, M.OP_IN, [ct])], price,
                                  bool(avail[s, o])))
                zones_av.append(z)
                cts_av.append(ct)
        reqs = [
            (M.KEY_ARCH, M.OP_IN, [arch]), (M.KEY_OS, M.OP_IN, ["linux"]),
            (M.KEY_ZONE, M.OP_IN, ZONES), (M.KEY_CAPACITY_TYPE, M.OP_IN, CAPACITY_TYPES),
            (KEY_FAMILY, M.OP_IN, [fam]), (KEY_SKU_CPU, M.OP_IN, [str(vcpu)]), (KEY_SKU_GPU, M.OP_IN, [str(gpu)]),
        ]
        if with_instance_type_key:
            reqs.insert(0, (M.KEY_INSTANCE_TYPE, M.OP_IN, [name]))
        b.add_sku(name, reqs, alloc, cap, offerings)


TAINT_UNIVERSE = [(f"example.com/taint-{i}", ["a", "b"][i % 2], [M.EFFECT_NOSCHEDULE, M.EFFECT_NOEXECUTE][i // 8 % 2])
                  for i in range(16)]
TEAMS = [f"team-{i}" for i in range(8)]


def add_pools(b: M.ModelBuilder, kind: str) -> None:
    if kind == "c2":
        b.add_pool("default", [(M.KEY_NODEPOOL, M.OP_IN, ["default"]), (M.KEY_OS, M.OP_IN, ["linux"])])
        return
    if kind == "c3":
        T = TAINT_UNIVERSE
        b.add_pool("gpu-reserved", [(M.KEY_NODEPOOL, M.OP_IN, ["gpu-reserved"]), (KEY_SKU_GPU, M.OP_GT, 0),
                                    (KEY_TEAM, M.OP_IN, TEAMS[:4])], taints=[T[0]], weight=100)
        b.add_pool("batch", [(M.KEY_NODEPOOL, M.OP_IN, ["batch"]), (M.KEY_CAPACITY_TYPE, M.OP_IN, ["spot"]),
                             (KEY_TEAM, M.OP_IN, TEAMS), (KEY_FAMILY, M.OP_NOTIN, ["NV"])], taints=[T[1], T[2]],
                   weight=50)
        b.add_pool("arm", [(M.KEY_NODEPOOL, M.OP_IN, ["arm"]), (M.KEY_ARCH, M.OP_IN, ["arm64"])], taints=[T[3]],
                   weight=10)  # does not define the team key: pods selecting a team are incompatible here
        b.add_pool("general", [(M.KEY_NODEPOOL, M.OP_IN, ["general"]), (KEY_TEAM, M.OP_IN, TEAMS),
                               (M.KEY_ZONE, M.OP_IN, ZONES[:2] + [ZONES[2]])], weight=10)
        return
    raise ValueError(kind)


CPU_M = [100, 250, 500, 1000, 2000, 4000, 8000, 16000]
CPU_W = [20, 18, 16, 14, 10, 6, 3, 1]
MEM_B = [128 * MiB, 256 * MiB, 512 * MiB, GiB, 2 * GiB, 4 * GiB, 8 * GiB, 16 * GiB, 32 * GiB, 64 * GiB]
MEM_W = [10, 12, 14, 14, 12, 10, 8, 5, 3, 1]
GPU_N = [0, 1, 2, 4, 8]
GPU_W = [80, 10, 5, 3, 2]
EPH_B = [0, GiB, 10 * GiB, 100 * GiB]
EPH_W = [50, 30, 15, 5]


def bulk_pods(b: M.ModelBuilder, n_pods: int, kind: str, seed: int = SEED_PODS) -> Dict[str, np.ndarray]:
    """Vectorised pod generation.  Returns the arrays ModelBuilder.build(bulk_pods=...) splices in."""
    P = n_pods
    R = len(b.res_names)
    req = np.zeros((P, R), dtype=np.int64)
    req[:, 0] = np.array(CPU_M, dtype=np.int64)[choice(seed, 1, P, CPU_W)]
    req[:, 1] = np.array(MEM_B, dtype=np.int64)[choice(seed, 2, P, MEM_W)]
    req[:, 2] = np.array(EPH_B, dtype=np.int64)[choice(seed, 3, P, EPH_W)]
    req[:, 3] = np.array(GPU_N, dtype=np.int64)[choice(seed, 4, P, GPU_W)]
    req[:, 4] = 1  # RequestsForPods adds pods: 1 (resources.go:36)
    ts = 1_700_000_000 + (sm64(seed, 5, P) % np.uint64(60)).astype(np.int64)
    uid_hi = sm64(seed, 6, P)
    uid_lo = sm64(seed, 7, P)
    # UUIDv4 version/variant bits, as a real kube-apiserver would mint them
    uid_hi = (uid_hi & np.uint64(0xFFFFFFFFFFFF0FFF)) | np.uint64(0x4000)
    uid_lo = (uid_lo & np.uint64(0x3FFFFFFFFFFFFFFF)) | np.uint64(0x8000000000000000)

    # requirement columns: (has[P] bool, key name, op, values: either list-of-list-of-str per choice + idx[P], or bound[P])
    cols: List[dict] = []
    tol_cols: List[dict] = []
    if kind == "c3":
        zsel = uniform(seed, 10, P) < 0.30
        zidx = choice(seed, 11, P, [1, 1, 1])
        cols.append(dict(has=zsel, key=M.KEY_ZONE, op=M.OP_IN, options=[[z] for z in ZONES], idx=zidx))
        asel = uniform(seed, 12, P) < 0.10
        aidx = choice(seed, 13, P, [3, 1])
        cols.append(dict(has=asel, key=M.KEY_ARCH, op=M.OP_IN, options=[["amd64"], ["arm64"]], idx=aidx))
        tsel = uniform(seed, 14, P) < 0.40
        tidx = choice(seed, 15, P, [1] * 8)
        cols.append(dict(has=tsel, key=KEY_TEAM, op=M.OP_IN, options=[[t] for t in TEAMS], idx=tidx))
        aff = uniform(seed, 16, P)
        # 5 % node affinity: NotIn on the family, Gt / Lt on the gpu count
        cols.append(dict(has=aff < 0.02, key=KEY_FAMILY, op=M.OP_NOTIN, options=[["NV"], ["NC", "NV"], ["F"]],
                         idx=choice(seed, 17, P, [1, 1, 1])))
        cols.append(dict(has=(aff >= 0.02) & (aff < 0.035), key=KEY_SKU_GPU, op=M.OP_GT,
                         bound=choice(seed, 18, P, [2, 1]).astype(np.int64)))  # > 0 or > 1
        cols.append(dict(has=(aff >= 0.035) & (aff < 0.05), key=KEY_SKU_GPU, op=M.OP_LT,
                         bound=np.array([4, 8], dtype=np.int64)[choice(seed, 19, P, [1, 1])]))
        cols.append(dict(has=(aff >= 0.05) & (aff < 0.06), key=M.KEY_CAPACITY_TYPE, op=M.OP_IN,
                         options=[["on-demand"], ["spot"]], idx=choice(seed, 20, P, [1, 1])))
        # tolerations: up to three slots
        for slot in range(3):
            has = uniform(seed, 30 + slot, P) < [0.6, 0.4, 0.2][slot]
            which = choice(seed, 40 + slot, P, [4, 3, 3, 2] + [1] * 12)
            form = choice(seed, 50 + slot, P, [6, 3, 1])  # exact / key Exists / wildcard Exists
            tol_cols.append(dict(has=has, which=which, form=form))

    # ---- assemble requirement CSR ----
    nreq = np.zeros(P, dtype=np.int64)
    for c in cols:
        nreq += c["has"]
    off = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(nreq, out=off[1:])
    total = int(off[-1])
    reqs = np.zeros(total, dtype=M.REQ_DTYPE)
    reqs["min_values"] = -1
    vals: List[np.ndarray] = []
    vbase = 0
    cursor = off[:-1].copy()
    for c in cols:
        has = c["has"]
        rows = np.nonzero(has)[0]
        pos = cursor[rows]
        cursor[rows] += 1
        k = b.key(c["key"])
        reqs["key"][pos] = k
        reqs["op"][pos] = c["op"]
        if c["op"] in (M.OP_IN, M.OP_NOTIN):
            opt_ids = [[b.value(k, v) for v in opt] for opt in c["options"]]
            idx = c["idx"][rows]
            cnt = np.array([len(o) for o in opt_ids], dtype=np.int64)[idx]
            voff = np.zeros(len(rows) + 1, dtype=np.int64)
            np.cumsum(cnt, out=voff[1:])
            flat = np.zeros(int(voff[-1]), dtype=np.int32)
            for oi, ids in enumerate(opt_ids):
                sel = np.nonzero(idx == oi)[0]
                for j, vid in enumerate(ids):
                    flat[voff[sel] + j] = vid
            reqs["val_off"][pos] = vbase + voff[:-1]
            reqs["val_cnt"][pos] = cnt
            vals.append(flat)
            vbase += len(flat)
        else:
            reqs["bound"][pos] = c["bound"][rows]
    req_values = np.concatenate(vals).astype(np.int32) if vals else np.zeros(0, dtype=np.int32)

    # ---- tolerations CSR ----
    ntol = np.zeros(P, dtype=np.int64)
    for c in tol_cols:
        ntol += c["has"]
    toff = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(ntol, out=toff[1:])
    tols = np.zeros(int(toff[-1]), dtype=M.TOL_DTYPE)
    cursor = toff[:-1].copy()
    tk = np.array([b.tstr(t[0]) for t in TAINT_UNIVERSE], dtype=np.int32)
    tv = np.array([b.tstr(t[1]) for t in TAINT_UNIVERSE], dtype=np.int32)
    te = np.array([t[2] for t in TAINT_UNIVERSE], dtype=np.int32)
    for c in tol_cols:
        rows = np.nonzero(c["has"])[0]
        pos = cursor[rows]
        cursor[rows] += 1
        w = c["which"][rows]
        f = c["form"][rows]
        tols["key"][pos] = np.where(f == 2, 0, tk[w])
        tols["value"][pos] = np.where(f == 0, tv[w], 0)
        tols["op"][pos] = np.where(f == 0, M.TOLOP_EQUAL, M.TOLOP_EXISTS)
        tols["effect"][pos] = np.where(f == 2, M.EFFECT_NONE, te[w])
    return dict(reqs=reqs, req_values=req_values, pod_req_off=off.astype(np.int32), tolerations=tols,
                pod_tol_off=toff.astype(np.int32), pod_requests=req, pod_ts=ts, pod_uid_hi=uid_hi, pod_uid_lo=uid_lo)


def build_model(b: M.ModelBuilder, pods: Optional[Dict[str, np.ndarray]] = None) -> M.Model:
    m = b.build()
    if pods is None:
        return m
    a = m.arrays
    # splice the bulk pod section in: requirement pool order is sku, offering, pool, pod, node
    p0 = int(a["pod_req_off"][0])
    assert int(a["pod_req_off"][-1]) == p0, "bulk pods cannot be mixed with add_pod"
    r = pods["reqs"].copy()
    r["val_off"] += len(a["req_values"])
    a["req_values"] = np.ascontiguousarray(np.concatenate([a["req_values"], pods["req_values"]]).astype(np.int32))
    a["reqs"] = np.ascontiguousarray(np.concatenate([a["reqs"][:p0], r, a["reqs"][p0:]]))
    a["pod_req_off"] = np.ascontiguousarray((pods["pod_req_off"].astype(np.int64) + p0).astype(np.int32))
    a["node_req_off"] = np.ascontiguousarray((a["node_req_off"].astype(np.int64) + len(r)).astype(np.int32))
    a["tolerations"] = np.ascontiguousarray(pods["tolerations"])
    for k in ("pod_tol_off", "pod_requests", "pod_ts", "pod_uid_hi", "pod_uid_lo"):
        a[k] = np.ascontiguousarray(pods[k])
    # value table may have grown while pods were generated
    a["val_key"] = np.ascontiguousarray(np.array(b.val_key, dtype=np.int32))
    isint, ints = [], []
    for s in b.val_names:
        ok, v = M._parse_int(s)
        isint.append(1 if ok else 0)
        ints.append(v)
    a["val_isint"] = np.ascontiguousarray(np.array(isint, dtype=np.uint8))
    a["val_int"] = np.ascontiguousarray(np.array(ints, dtype=np.int64))
    m.key_names = list(b.key_names)
    m.val_names = list(b.val_names)
    return m.finalize()


def new_builder() -> M.ModelBuilder:
    b = M.ModelBuilder(RES)
    b.wellknown_extra.update(PROVIDER_WELLKNOWN)
    # fix the key numbering so that it does not depend on generation order
    for k in (M.KEY_INSTANCE_TYPE, M.KEY_ARCH, M.KEY_OS, M.KEY_ZONE, M.KEY_CAPACITY_TYPE, KEY_FAMILY, KEY_SKU_CPU,
              KEY_SKU_GPU, M.KEY_NODEPOOL, KEY_TEAM):
        b.key(k)
    return b


def config(kind: str, n_pods: int, n_skus: int, seed_pods: int = SEED_PODS, seed_catalog: int = SEED_CATALOG,
           with_instance_type_key: bool = True) -> M.Model:
    """kind 'c2': one untainted pool, plain pods.  kind 'c3': four weighted, tainted pools; pods with
    tolerations, zone / arch / team selectors and NotIn / Gt / Lt node affinities."""
    b = new_builder()
    add_catalog(b, n_skus, seed_catalog, with_instance_type_key)
    add_pools(b, kind)
    pods = bulk_pods(b, n_pods, kind, seed_pods)
    return build_model(b, pods)


def c2(n_pods: int = 10_000, n_skus: int = 500) -> M.Model:
    return config("c2", n_pods, n_skus)


def c3(n_pods: int = 100_000, n_skus: int = 1_000) -> M.Model:
    return config("c3", n_pods, n_skus)


def c5(n_pods: int = 1_000_000, n_skus: int = 2_000) -> M.Model:
    return config("c3", n_pods, n_skus)


def c4(n_pods: int = 50_000, n_skus: int = 500, cap: float = 0.6, solve=None):
    """configs[3], consolidation re-pack: `n_pods` RUNNING pods on nodes produced by a deliberately loose packing
    (a Solve against instance types shrunk to `cap` of their allocatable), to be re-packed onto fewer NodeClaims.

    `solve(model)` must return an object with claim_sku, claim_requests, pod_place_index (the CPU oracle's solve in
    tests, the device path in the bench).  Returns (model with the nodes as existing nodes and ALL pods, node_of_pod,
    utilisation per node).  A single-node consolidation simulation for candidate i is
    model.subset(pods on node i, all nodes but i)."""
    b = new_builder()
    add_catalog(b, n_skus)
    add_pools(b, "c2")
    pods = bulk_pods(b, n_pods, "c2", SEED_PODS + 4)
    loose = build_model(b, pods)
    real_alloc = loose.arrays["sku_alloc"].copy()
    shrunk = (real_alloc.astype(np.float64) * cap).astype(np.int64)
    shrunk[:, 4] = real_alloc[:, 4]  # the pods-per-node capacity is not shrunk
    loose.arrays["sku_alloc"] = np.ascontiguousarray(shrunk)
    loose.finalize()
    r = solve(loose)
    placed = r.pod_place_index >= 0
    n_nodes = int(r.claim_sku.shape[0])
    # second model: same catalog / pools / pods, the claims become existing nodes on their selected SKU
    b2 = new_builder()
    add_catalog(b2, n_skus)
    add_pools(b2, "c2")
    pods2 = bulk_pods(b2, n_pods, "c2", SEED_PODS + 4)
    zone_u = sm64(SEED_NODES, 1, n_nodes) % np.uint64(len(ZONES))
    ct_u = sm64(SEED_NODES, 2, n_nodes) % np.uint64(len(CAPACITY_TYPES))
    util = np.zeros(n_nodes)
    for i in range(n_nodes):
        s = int(r.claim_sku[i])
        sku = b2.skus[s]
        labels = {M.KEY_INSTANCE_TYPE: sku["name"], M.KEY_NODEPOOL: "default", M.KEY_OS: "linux",
                  M.KEY_ZONE: ZONES[int(zone_u[i])], M.KEY_CAPACITY_TYPE: CAPACITY_TYPES[int(ct_u[i])]}
        avail = real_alloc[s] - r.claim_requests[i]
        util[i] = float(r.claim_requests[i][0]) / max(float(real_alloc[s][0]), 1.0)
        b2.add_node(f"node-{i:06d}", labels, [int(x) for x in avail])
    model = build_model(b2, pods2)
    node_of_pod = np.where(placed, r.pod_place_index, -1).astype(np.int64)
    return model, node_of_pod, util


def c4_simulations(model: M.Model, node_of_pod: np.ndarray, util: np.ndarray, n_candidates: int, spread: bool = False):
    """Single-node consolidation candidates, lowest utilisation first (singlenodeconsolidation.go:68-101): one
    simulation per candidate = its pods against every other node.  `spread` takes the candidates evenly across the
    whole utilisation order instead of its head, so that a batch holds nodes that can be deleted, nodes that need a
    replacement and nodes that are best left alone."""
    E = int(model.c.n_nodes)
    order = np.argsort(util, kind="stable")
    order = order[np.linspace(0, E - 1, n_candidates).astype(np.int64)] if spread else order[:n_candidates]
    all_nodes = np.arange(E)
    sims = []
    for i in order:
        pods = np.nonzero(node_of_pod == i)[0]
        sims.append((int(i), M.subset(model, pods, all_nodes[all_nodes != i])))
    return sims


def c4_shared(model: M.Model, node_of_pod: np.ndarray, candidates: Sequence[int]):
    """The same simulations as ONE model for b200s_solve_batch_shared: every node of the cluster, the candidates' pods
    pending one candidate after the other.  Returns (model, pod_off, candidates): problem b = pods
    pod_off[b]..pod_off[b+1] against all nodes but candidates[b] (model node indices; the node ROW of a model node is
    found through the encoded nodes' node_index)."""
    pods, off = [], [0]
    for i in candidates:
        pods.append(np.nonzero(node_of_pod == i)[0])
        off.append(off[-1] + len(pods[-1]))
    allp = np.concatenate(pods) if pods else np.zeros(0, np.int64)
    return M.subset(model, allp, np.arange(int(model.c.n_nodes))), np.asarray(off, np.int32), [int(i) for i in candidates]


def candidate_prices(model: M.Model, nodes: Sequence[int]):
    """getCandidatePrices for single-node candidates (K/controllers/disruption/consolidation.go:310-328): the cheapest
    offering of the node's instance type that its zone / capacity-type labels admit (availability is NOT looked at
    there), and whether the node is a spot node.  Returns (prices, is_spot) per node."""
    from .consolidation import node_facts
    facts = node_facts(model, nodes)
    for e, (_, p, _, reserved) in zip(nodes, facts):
        if p < 0 and not reserved:  # a reserved-capacity node without an offering is free (consolidation.go:317-320)
            raise ValueError(f"node {e}: unable to determine offering")
    return [max(p, 0.0) for _, p, _, _ in facts], [sp for _, _, sp, _ in facts]

"""Random small scheduling models that poke at the edges of the requirement algebra, taints, offerings,
existing nodes and the queue's retry rule.  Deterministic per seed."""
import numpy as np

from b200sched import model as M

RES = ["cpu", "memory", "nvidia.com/gpu", "pods"]
CUSTOM = ["example.com/a", "example.com/b", "example.com/num"]
WK = [M.KEY_ARCH, M.KEY_OS]
VALS = {"example.com/a": ["x", "y", "z"], "example.com/b": ["p", "q"], "example.com/num": ["1", "2", "4", "8", "n/a"],
        M.KEY_ARCH: ["amd64", "arm64"], M.KEY_OS: ["linux", "windows"]}
ZONES = ["z1", "z2", "z3"]
CTS = ["on-demand", "spot"]
TAINTS = [("t1", "v", M.EFFECT_NOSCHEDULE), ("t2", "w", M.EFFECT_NOEXECUTE), ("t1", "other", M.EFFECT_NOSCHEDULE)]


def _rand_req(rng, key, allow_bounds=True):
    vals = VALS[key]
    ops = [M.OP_IN, M.OP_IN, M.OP_NOTIN, M.OP_EXISTS, M.OP_DOESNOTEXIST]
    if allow_bounds and key == "example.com/num":
        ops += [M.OP_GT, M.OP_LT, M.OP_GT, M.OP_LT]
    op = ops[rng.integers(len(ops))]
    if op in (M.OP_IN, M.OP_NOTIN):
        n = rng.integers(1, len(vals) + 1)
        return (key, op, list(rng.choice(vals, size=n, replace=False)))
    if op in (M.OP_GT, M.OP_LT):
        return (key, op, int(rng.integers(0, 9)))
    return (key, op, None)


def random_model(seed: int, with_nodes: bool = True, max_pods: int = 40, keyless: bool = False,
                 limits: bool = False) -> M.Model:
    rng = np.random.default_rng(seed)
    b = M.ModelBuilder(RES)
    for k in [M.KEY_INSTANCE_TYPE, M.KEY_ZONE, M.KEY_CAPACITY_TYPE] + WK + CUSTOM + [M.KEY_NODEPOOL]:
        b.key(k)
    S = int(rng.integers(1, 70))
    price_seq = rng.permutation(S * 6) + 1
    for s in range(S):
        reqs = [(M.KEY_INSTANCE_TYPE, M.OP_IN, [f"sku-{s}"])]
        for key in WK + CUSTOM:
            u = rng.random()
            if u < 0.15:
                continue                                   # SKU does not define the key
            if u < 0.22:
                reqs.append((key, M.OP_DOESNOTEXIST, None))
                continue
            n = 1 if rng.random() < 0.7 else int(rng.integers(1, len(VALS[key]) + 1))
            reqs.append((key, M.OP_IN, list(rng.choice(VALS[key], size=n, replace=False))))
        zs = [z for z in ZONES if rng.random() < 0.8] or [ZONES[0]]
        cs = [c for c in CTS if rng.random() < 0.85] or [CTS[0]]
        reqs.append((M.KEY_ZONE, M.OP_IN, zs))
        reqs.append((M.KEY_CAPACITY_TYPE, M.OP_IN, cs))
        offs = []
        i = 0
        for z in zs:
            for c in cs:
                offs.append(([(M.KEY_ZONE, M.OP_IN, [z]), (M.KEY_CAPACITY_TYPE, M.OP_IN, [c])],
                             float(price_seq[s * 6 + i]) * 0.01, bool(rng.random() < 0.85)))
                i += 1
        cpu = int(rng.choice([2000, 4000, 8000, 16000]))
        cap = [cpu, int(rng.choice([4, 8, 16, 64])) << 30, int(rng.choice([0, 0, 1, 4])), int(rng.choice([4, 10, 30]))]
        alloc = [cap[0] - 100, cap[1] - (1 << 30), cap[2], cap[3]]
        if rng.random() < 0.03:
            alloc[1] = -5                                   # a negative allocatable entry never fits (resources.go:234-238)
        b.add_sku(f"sku-{s}", reqs, alloc, cap, offs)
    npools = int(rng.integers(1, 4))
    for i in range(npools):
        reqs = [(M.KEY_NODEPOOL, M.OP_IN, [f"pool-{i}"])]
        for key in WK + CUSTOM + [M.KEY_ZONE, M.KEY_CAPACITY_TYPE]:
            if rng.random() < 0.35:
                if key in (M.KEY_ZONE, M.KEY_CAPACITY_TYPE):
                    pool = ZONES if key == M.KEY_ZONE else CTS
                    n = int(rng.integers(1, len(pool) + 1))
                    op = M.OP_IN if rng.random() < 0.8 else M.OP_NOTIN
                    reqs.append((key, op, list(rng.choice(pool, size=n, replace=False))))
                else:
                    reqs.append(_rand_req(rng, key))
                    if rng.random() < 0.2:
                        reqs.append(_rand_req(rng, key))   # two requirements on one key intersect (requirements.go:127-134)
        if rng.random() < 0.2:
            names = [f"sku-{s}" for s in rng.choice(S, size=min(S, int(rng.integers(1, 8))), replace=False)]
            reqs.append((M.KEY_INSTANCE_TYPE, M.OP_IN if rng.random() < 0.6 else M.OP_NOTIN, names + ["not-in-catalog"]))
        taints = [TAINTS[j] for j in range(len(TAINTS)) if rng.random() < 0.25]
        daemon = [int(rng.choice([0, 0, 100])), int(rng.choice([0, 0, 1 << 28])), 0, int(rng.choice([0, 1]))]
        lim = None
        if limits and rng.random() < 0.6:  # NodePool limits (scheduler.go:831-867): remaining cpu / memory / gpu of the pool
            lim = {}
            if rng.random() < 0.8:
                lim["cpu"] = int(rng.choice([1000, 4000, 16000, 40000, 100000]))
            if rng.random() < 0.5:
                lim["memory"] = int(rng.choice([2, 16, 64, 200])) << 30
            if rng.random() < 0.2:
                lim["nvidia.com/gpu"] = int(rng.choice([0, 1, 4]))
        b.add_pool(f"pool-{i}", reqs, taints=taints, weight=int(rng.integers(0, 3)), daemon=daemon, limits=lim)
    if with_nodes:
        for e in range(int(rng.integers(0, 6))):
            labels = {M.KEY_INSTANCE_TYPE: f"sku-{int(rng.integers(S))}", M.KEY_ZONE: str(rng.choice(ZONES)),
                      M.KEY_CAPACITY_TYPE: str(rng.choice(CTS)), "unrelated/label": f"v{e}"}
            for key in WK + CUSTOM:
                if rng.random() < 0.6:
                    labels[key] = str(rng.choice(VALS[key]))
            avail = [int(rng.choice([500, 2000, 8000])), int(rng.choice([1, 4, 16])) << 30, int(rng.choice([0, 1])),
                     int(rng.choice([1, 3, 10]))]
            if rng.random() < 0.1:
                avail[0] = -100
            b.add_node(f"node-{rng.integers(1000):04d}-{e}", labels, avail,
                       taints=[TAINTS[j] for j in range(len(TAINTS)) if rng.random() < 0.2],
                       initialized=bool(rng.random() < 0.7))
    P = int(rng.integers(1, max_pods + 1))
    for p in range(P):
        reqs = []
        for key in WK + CUSTOM + [M.KEY_ZONE, M.KEY_CAPACITY_TYPE]:
            if rng.random() < 0.22 and not keyless:  # keyless: pods without nodeSelector / affinity (the common case)
                if key in (M.KEY_ZONE, M.KEY_CAPACITY_TYPE):
                    pool = ZONES if key == M.KEY_ZONE else CTS
                    reqs.append((key, M.OP_IN if rng.random() < 0.7 else M.OP_NOTIN,
                                 list(rng.choice(pool, size=int(rng.integers(1, len(pool) + 1)), replace=False))))
                else:
                    reqs.append(_rand_req(rng, key))
        tols = []
        for (k, v, e) in TAINTS:
            u = rng.random()
            if u < 0.3:
                tols.append((k, M.TOLOP_EQUAL, v, e))
            elif u < 0.45:
                tols.append((k, M.TOLOP_EXISTS, "", M.EFFECT_NONE))
        if rng.random() < 0.1:
            tols.append(("", M.TOLOP_EXISTS, "", M.EFFECT_NONE))
        cpu = int(rng.choice([100, 500, 1000, 1000, 4000, 20000]))
        mem = int(rng.choice([1 << 28, 1 << 30, 1 << 30, 1 << 32, 1 << 36]))
        b.add_pod(reqs, [cpu, mem, int(rng.choice([0, 0, 0, 1, 8])), 1], tolerations=tols,
                  ts=int(rng.integers(0, 3)), uid=int(rng.integers(1, 1 << 62)) << 64 | int(rng.integers(1, 1 << 62)))
    m = b.build()
    return m

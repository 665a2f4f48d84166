"""Python twin of go/pkg/providers/instancetype: the static Azure GPU SKU catalog behind CloudProvider.GetInstanceTypes
(the reference returns an empty slice, pkg/cloudprovider/cloudprovider.go:99-101), read from the same JSON
(catalog/azure_gpu_skus.json == go/pkg/providers/instancetype/azure_gpu_skus.json) and turned into ModelBuilder rows."""
from __future__ import annotations

import json
import os
from typing import List

from . import model as M

HERE = os.path.dirname(os.path.abspath(__file__))
TABLE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "catalog", "azure_gpu_skus.json")
KEY_FAMILY = "karpenter.azure.com/sku-family"
KEY_CPU = "karpenter.azure.com/sku-cpu"
KEY_GPU_COUNT = "karpenter.azure.com/sku-gpu-count"
KEY_GPU_NAME = "karpenter.azure.com/sku-gpu-name"
PROVIDER_WELLKNOWN = (KEY_FAMILY, KEY_CPU, KEY_GPU_COUNT, KEY_GPU_NAME)
SPOT_DISCOUNT = 0.3
RES = ["cpu", "memory", "ephemeral-storage", "nvidia.com/gpu", "pods"]


def load() -> dict:
    return json.load(open(TABLE))


def family(name: str) -> str:
    s = name[len("Standard_"):]
    i = 0
    while i < len(s) and not s[i].isdigit():
        i += 1
    return s[:i]


def add_catalog(b: M.ModelBuilder, region: str = "eastus") -> List[str]:
    """one row per SKU: requirements for every well-known label, 3 zones x {on-demand, spot} offerings, capacity, and
    Allocatable() = capacity - kubeReserved - evictionThreshold (types.go:181-219).  Returns the SKU names."""
    t = load()
    assert b.res_names == RES
    ov = t["overhead"]
    b.wellknown_extra.update(PROVIDER_WELLKNOWN)
    zones = [f"{region}-{z}" for z in t["zones"]]
    for s in t["skus"]:
        reqs = [(M.KEY_INSTANCE_TYPE, M.OP_IN, [s["name"]]), (M.KEY_ARCH, M.OP_IN, ["amd64"]), (M.KEY_OS, M.OP_IN, ["linux"]),
                (M.KEY_ZONE, M.OP_IN, zones), ("topology.kubernetes.io/region", M.OP_IN, [region]),
                (M.KEY_CAPACITY_TYPE, M.OP_IN, ["on-demand", "spot"]), ("node.kubernetes.io/windows-build", M.OP_DOESNOTEXIST, None),
                (KEY_FAMILY, M.OP_IN, [family(s["name"])]), (KEY_CPU, M.OP_IN, [str(s["vcpu"])]),
                (KEY_GPU_COUNT, M.OP_IN, [str(s["gpu"])]), (KEY_GPU_NAME, M.OP_IN, [s["gpu_name"]])]
        offs = []
        for zi, z in enumerate(zones):
            for ct in ("on-demand", "spot"):
                price = s["price_on_demand"] * (SPOT_DISCOUNT if ct == "spot" else 1.0) + zi * 1e-6
                offs.append(([(M.KEY_ZONE, M.OP_IN, [z]), (M.KEY_CAPACITY_TYPE, M.OP_IN, [ct])], price, True))
        cap = [s["vcpu"] * 1000, s["memory_gib"] << 30, ov["ephemeral_gib"] << 30, s["gpu"], ov["pods"]]
        alloc = [cap[0] - ov["kube_reserved_cpu_milli"], cap[1] - ((ov["kube_reserved_memory_mib"] + ov["eviction_memory_mib"]) << 20),
                 cap[2], cap[3], cap[4]]
        b.add_sku(s["name"], reqs, alloc, cap, offs)
    return [s["name"] for s in t["skus"]]


def cheapest(names: List[str], capacity_type: str = "on-demand") -> str:
    """the price-ordered Create pick (go/pkg/providers/instance/select.go) over the table's own prices"""
    t = {s["name"]: s["price_on_demand"] for s in load()["skus"]}
    mul = SPOT_DISCOUNT if capacity_type == "spot" else 1.0
    return min(names, key=lambda n: (t[n] * mul, names.index(n)))

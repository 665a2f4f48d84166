"""Python twin of go/pkg/b200sched/flatten.go: Kubernetes-shaped pod specs (plain dicts, the JSON wire form) -> the rows
of the Level-1 model (include/b200sched.h).  The Go file cannot be compiled in the authoring image; every rule it
restates is written here a second time, against the same reference lines, and exercised by tests/test_flatten.py.

  pod_requirement_rows   NewPodRequirements     K/scheduling/requirements.go:74-110
  relaxable              Preferences.Relax       K/controllers/provisioning/scheduling/preferences.go:38-57 (would it act?)
  plain_pod              SURVEY.md Appendix A.12 (the scheduler branches that are not kernelised must be inert)
  uid_words              queue.go:105-107        (UID string order == 128-bit integer order for canonical UUIDs)
  toleration_rows        core/v1 Toleration      -> (key, op, value, effect) of b200s_toleration_t
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from . import model as M

ReqRow = Tuple[str, int, object]

_OPS = {"In": M.OP_IN, "NotIn": M.OP_NOTIN, "Exists": M.OP_EXISTS, "DoesNotExist": M.OP_DOESNOTEXIST, "Gt": M.OP_GT,
        "Lt": M.OP_LT}
_EFFECTS = {"": M.EFFECT_NONE, "NoSchedule": M.EFFECT_NOSCHEDULE, "PreferNoSchedule": M.EFFECT_PREFERNOSCHEDULE,
            "NoExecute": M.EFFECT_NOEXECUTE}


class Fallback(ValueError):
    """the batch stays on the vendored CPU scheduler (go: ErrFallback)"""


def _expr_row(e: dict) -> ReqRow:
    op = e.get("operator")
    if op not in _OPS:
        raise Fallback(f"operator {op!r}")
    code = _OPS[op]
    if code in (M.OP_GT, M.OP_LT):
        vals = e.get("values") or []
        if len(vals) != 1:
            raise Fallback(f"{op} needs exactly one value")
        try:
            return (e["key"], code, int(vals[0]))
        except ValueError:
            raise Fallback(f"{op} bound {vals[0]!r}")
    if code in (M.OP_IN, M.OP_NOTIN):
        return (e["key"], code, list(e.get("values") or []))
    return (e["key"], code, None)


def _label_rows(labels: Optional[Dict[str, str]]) -> List[ReqRow]:
    """NewLabelRequirements (requirements.go:58-64): key In [value] per label, in key order"""
    return [(k, M.OP_IN, [v]) for k, v in sorted((labels or {}).items())]


def pod_requirement_rows(spec: dict) -> List[ReqRow]:
    """nodeSelector AND the heaviest preferred node-affinity term AND the first required term.  Rows on one key are
    intersected by the encoder exactly as Requirements.Add does (requirements.go:127-134)."""
    rows = _label_rows(spec.get("nodeSelector"))
    na = ((spec.get("affinity") or {}).get("nodeAffinity")) or None
    if not na:
        return rows
    pref = na.get("preferredDuringSchedulingIgnoredDuringExecution") or []
    if pref:
        # sort.Slice(weight descending) is unstable; up to 12 elements it is an insertion sort, i.e. stable: the first
        # term of maximal weight.  A tie at the top of a longer list stays on the CPU path.
        best = 0
        for i in range(1, len(pref)):
            if pref[i].get("weight", 0) > pref[best].get("weight", 0):
                best = i
        if len(pref) > 12 and any(i != best and p.get("weight", 0) == pref[best].get("weight", 0) for i, p in enumerate(pref)):
            raise Fallback("tie among more than 12 preferred node-affinity terms")
        if pref[best].get("preference", {}).get("matchFields"):
            raise Fallback("matchFields")
        rows += [_expr_row(e) for e in pref[best].get("preference", {}).get("matchExpressions") or []]
    req = na.get("requiredDuringSchedulingIgnoredDuringExecution")
    if req and req.get("nodeSelectorTerms"):
        rows += [_expr_row(e) for e in req["nodeSelectorTerms"][0].get("matchExpressions") or []]
    return rows


def relaxable(spec: dict) -> bool:
    na = ((spec.get("affinity") or {}).get("nodeAffinity")) or None
    if not na:
        return False
    if na.get("preferredDuringSchedulingIgnoredDuringExecution"):
        return True
    req = na.get("requiredDuringSchedulingIgnoredDuringExecution")
    return bool(req and len(req.get("nodeSelectorTerms") or []) > 1)


def plain_pod(spec: dict) -> None:
    """raises Fallback when a scheduler branch outside the encoding would act on this pod"""
    aff = spec.get("affinity") or {}
    if aff.get("podAffinity") or aff.get("podAntiAffinity"):
        raise Fallback("pod (anti-)affinity")
    if spec.get("topologySpreadConstraints"):
        raise Fallback("topology spread")
    if spec.get("resourceClaims"):
        raise Fallback("resource claims")
    for v in spec.get("volumes") or []:
        if v.get("persistentVolumeClaim") or v.get("ephemeral"):
            raise Fallback("persistent volumes")
    for c in (spec.get("containers") or []) + (spec.get("initContainers") or []):
        for p in c.get("ports") or []:
            if p.get("hostPort"):
                raise Fallback("host ports")


def uid_words(uid: str) -> Tuple[int, int]:
    h = uid.replace("-", "")
    if len(h) != 32 or any(c not in "0123456789abcdef" for c in h):
        raise Fallback(f"pod UID {uid!r} is not a canonical UUID")
    return int(h[:16], 16), int(h[16:], 16)


def toleration_rows(tolerations: Sequence[dict]) -> List[Tuple[str, int, str, int]]:
    out = []
    for t in tolerations or []:
        op = {"": M.TOLOP_EMPTY, "Equal": M.TOLOP_EQUAL, "Exists": M.TOLOP_EXISTS}.get(t.get("operator", ""), M.TOLOP_OTHER)
        out.append((t.get("key", ""), op, t.get("value", ""), _EFFECTS.get(t.get("effect", ""), M.EFFECT_NONE)))
    return out


_SUFFIX = {"Ki": 1 << 10, "Mi": 1 << 20, "Gi": 1 << 30, "Ti": 1 << 40, "k": 10 ** 3, "M": 10 ** 6, "G": 10 ** 9, "T": 10 ** 12}


def quantity(name: str, q) -> int:
    """exact fixed-scale integer of a resource.Quantity string: cpu in milli-units, everything else in base units"""
    if isinstance(q, int):
        return q * 1000 if name == "cpu" else q
    s = str(q)
    if name == "cpu":
        if s.endswith("m"):
            return int(s[:-1])
        v = float(s) * 1000
        if v != int(v):
            raise Fallback(f"cpu quantity {s} is finer than 1m")
        return int(v)
    for suf, mul in _SUFFIX.items():
        if s.endswith(suf):
            return int(s[:-len(suf)]) * mul
    if s.endswith("m"):
        raise Fallback(f"quantity {s} of {name} is not an integer")
    return int(s)


def pod_requests(spec: dict, res_names: Sequence[str]) -> List[int]:
    """RequestsForPods (K/utils/resources/resources.go:30-38,114-144): sum of containers (a limit stands in for a missing
    request), max with each init container on top of the sidecars before it, + overhead, + pods: 1"""
    def one(c, r):
        res = c.get("resources") or {}
        if r in (res.get("requests") or {}):
            return quantity(r, res["requests"][r])
        if r in (res.get("limits") or {}):
            return quantity(r, res["limits"][r])
        return 0
    out = []
    for r in res_names:
        if r == "pods":
            out.append(1)
            continue
        total = sum(one(c, r) for c in spec.get("containers") or [])
        side, init_max = 0, 0
        for c in spec.get("initContainers") or []:
            if c.get("restartPolicy") == "Always":
                total += one(c, r)
                side += one(c, r)
                init_max = max(init_max, side)
            else:
                init_max = max(init_max, one(c, r) + side)
        total = max(total, init_max)
        total += quantity(r, (spec.get("overhead") or {}).get(r, 0)) if r in (spec.get("overhead") or {}) else 0
        out.append(total)
    return out


def add_pods(b: M.ModelBuilder, pods: Sequence[dict]) -> List[bool]:
    """pods = [{"metadata": {"uid", "creationTimestamp" (seconds)}, "spec": {...}}]; returns the relaxable flags"""
    flags = []
    for p in pods:
        spec = p.get("spec") or {}
        plain_pod(spec)
        hi, lo = uid_words(p["metadata"]["uid"])
        b.add_pod(pod_requirement_rows(spec), pod_requests(spec, b.res_names), toleration_rows(spec.get("tolerations")),
                  ts=int(p["metadata"].get("creationTimestamp", 0)), uid=(hi << 64) | lo)
        flags.append(relaxable(spec))
    return flags

"""The JSON lines bench.py printed at the end of round 2 (committed under profiles/) against the keys the driver's
contract names: a regression of the line's shape shows up here, without a GPU."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config"}


def test_single_gpu_line_has_the_contract_keys():
    d = _line("r02_bench_c2_final.json")
    assert BASE <= set(d) and {"roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert set(d["config"]) >= {"workload"} and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("reference", "port")
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # host buffers and copies inside the timed region cannot be faster than resident inputs
    assert d["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # round 2's headline: one config[1] step at or under 7 ms, at least 1e9 evaluations/s, the round-1 digest
    assert d["ms_per_step"] <= 7.0 and d["value"] >= 1e9 and "1423b60e3481b574" in d["parity"]
    assert d["steps"] * 0 == 0 and d["warmup"] >= 3


def test_reference_arm_line():
    d = _line("r02_bench_c2_reference_arm.json")
    assert d["impl"] == "reference" and BASE <= set(d)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert {"kind", "cores", "sample", "value"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["value"] == d["value"]
    ours = _line("r02_bench_c2_final.json")
    assert d["config"] == ours["config"] and d["metric"] == ours["metric"] and d["unit"] == ours["unit"]


def test_multi_gpu_lines():
    for name, n in (("r02_bench_c5_n2_final.json", 2), ("r02_bench_c5_n4_peer_store.json", 4), ("r02_bench_c5_n8.json", 8)):
        d = _line(name)
        assert BASE <= set(d) and d["n_gpus"] == n and d["scaling"] == "strong"
        assert {"roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "single_gpu_same_run"} <= set(d)
        assert d["allgather_bytes_per_pod"] == 8
    d = _line("r02_bench_c5_n2_final.json")
    assert d["global_pack"]["digest"] == "e07e0c4838d5194d"
    p, c = d["exchange_peer_store"], d["exchange_nccl"]
    assert p["ms_per_step"] < c["ms_per_step"] and abs(d["ms_per_step"] - p["ms_per_step"]) < 1e-9

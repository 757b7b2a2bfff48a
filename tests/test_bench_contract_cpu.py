"""The committed bench lines (profiles/r2_bench_*.json, produced by bench.py on B200s) carry every key the measurement
contract asks for, and their numbers are self-consistent.  Guards the JSON schema of bench.py against regressions."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    with open(path) as f:
        return json.loads([l for l in f if l.startswith("{")][0])


BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}


@pytest.mark.parametrize("name,n", [("r2_bench_n1.json", 1), ("r2_bench_n2.json", 2), ("r2_bench_n4.json", 4), ("r2_bench_n8.json", 8)])
def test_cfg3_lines_follow_the_contract(name, n):
    d = _line(name)
    assert BASE_KEYS <= set(d) and d["n_gpus"] == n and d["metric"] == "hybrid_queries_per_sec_top100" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"]["queries"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]       # whole-job queries / s
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 < e["d2h_bytes_per_step"]
    assert e["value"] < d["value"] * 1.02                              # host copies inside the timed region cannot make it faster
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c) and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    c4 = d["cfg4"]
    assert c4["scaling"] == "weak" and f"over {n} GPU(s)" in c4["workload"] and c4["value"] > 40_000
    if n == 1:
        b = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(b) and b["kind"] in ("port", "reference") and b["cores"] >= 1
        assert d["parity"]["green"] is True and d["parity"]["fused_ordered_id_score_pairs_exact"] == d["parity"]["queries"]
        assert d["scaling"] == "strong" and d["roofline"]["bound"] == "tensor" and 0.7 <= d["roofline"]["frac"] <= 1.0
    else:
        hashes = {l["result_hash"] for l in d["layouts"]}
        assert len(hashes) == 1 and all(l["parity"] is True for l in d["layouts"])           # every layout, bit-identical results
        assert {l["row_shards"] for l in d["layouts"]} >= {1, n}
        assert d["value"] == max(l["value"] for l in d["layouts"])                           # the headline is the fastest layout
        assert d["replicas_weak"]["scaling"] == "weak" and d["replicas_weak"]["global_batch"] == n * d["config"]["queries"]


def test_scaling_of_the_committed_lines():
    v = {n: _line(f"r2_bench_n{n}.json")["value"] for n in (1, 2, 4, 8)}
    assert v[1] < v[2] < v[4] < v[8]
    c4 = {n: _line(f"r2_bench_n{n}.json")["cfg4"]["value"] for n in (1, 8)}
    assert c4[8] / c4[1] > 0.9                                         # 100 M notes over 8 GPUs at the speed of one 12.5 M-note shard


def test_reference_arm_and_cfg5_lines():
    r = _line("r2_bench_n1_reference_arm.json")
    assert r["impl"] == "reference" and r["metric"] == "hybrid_queries_per_sec_top100" and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    b = _line("r2_bench_n1.json")
    assert {k: v for k, v in r["config"].items()} == {k: v for k, v in b["config"].items()}   # both arms: the same config
    c = _line("r2_bench_cfg5.json")
    assert c["config"]["notes"] == 10_000_000 and c["config"]["nlist"] == 4096 and c["config"]["nprobe"] == 32
    assert c["parity"]["green"] is True and c["recall_at_100_vs_flat"]["value"] >= 0.95 and c["e2e"]["h2d_bytes_per_step"] > 0

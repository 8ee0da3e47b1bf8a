"""The bench.py output contract: the reference arm (CPU, runs anywhere) prints exactly one JSON line with
the required keys, and the committed B200 lines in profiles/ carry every key the contract names."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0",
                        "--ref-pairs", "100000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["warmup"] >= 3 and j["steps"] == 2
    assert j["metric"] == "read_pairs_per_s_anchored" and j["unit"] == "pairs/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["vs_baseline"] is None and j["dtype"] == "u8" and "workload" in j["config"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_b200_lines_follow_the_contract():
    prof = os.path.join(ROOT, "profiles")
    for name, n in (("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n4.json", 4), ("r01_bench_n8.json", 8)):
        j = json.load(open(os.path.join(prof, name)))
        assert BASE_KEYS | {"roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"} <= set(j), name
        assert j["n_gpus"] == n and j["scaling"] == "weak" and j["data"] == "synthetic" and j["warmup"] >= 3
        rf = j["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "hbm" and rf["unit"] == "GB/s"
        assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["frac"] >= 0.60
        assert abs(j["value"] - n * j["config"]["pairs_per_gpu"] / (j["ms_per_step"] * 1e-3)) < 1e-3 * j["value"]
        e = j["e2e"]
        assert e["h2d_bytes_per_step"] == 800_000_000 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
        assert j["gpu_launches"] == 6 * j["steps"]
        assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            c = j["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
        else:
            assert "validated" in j["config"]["exchange"]

"""The bench.py output contract.  CPU: the reference arm prints exactly one JSON line with the required keys and
loads no product library.  GPU (-m gpu): a short LIVE run of the GPU arm carries every key the contract names,
its roofline is consistent with its own timings, and the records it delivered equal the oracle's."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_one_json_line_and_loads_no_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '0', '--ref-pairs', '100000'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read();"
            "assert 'anchored_fusion_b200' not in sys.modules and 'libafb200' not in maps, 'the CPU arm loaded the product';"
            "assert 'libaf_oracle' in maps and 'libaf_synth' in maps" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["warmup"] >= 3 and j["steps"] == 2
    assert j["metric"] == "read_pairs_per_s_anchored" and j["unit"] == "pairs/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["vs_baseline"] is None and j["dtype"] == "u8" and "workload" in j["config"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_describe_the_same_config():
    """`config` names the workload only: the two arms build it from the same function with the same flags."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(workload="config1", pairs=10_000_000, read_len=150, anchor_len=6783, sub_ppm=10_000, fusion_ppm=0,
                           total_pairs=100_000_000, cells=4000, pairs_per_cell=5000)
    c = bench.config_dict(a, 1)
    assert set(c) == {"workload", "pairs_per_gpu", "read_len", "anchor_len", "sharding", "l2_policy"} and "configs[1]" in c["workload"]
    a.workload = "config3"
    assert bench.config_dict(a, 8)["pairs_per_gpu"] == 12_500_000 and "configs[2]" in bench.config_dict(a, 8)["workload"]


@pytest.mark.gpu
def test_gpu_arm_live_line_follows_the_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3", "--cpu-pairs", "2000000",
                        "--fastq-pairs", "200000", "--genome-bases", "200000000", "--genome-reads", "300"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity", "run"} <= set(j)
    assert j["n_gpus"] == 1 and j["scaling"] == "weak" and j["data"] == "synthetic" and j["warmup"] >= 3 and j["steps"] == 20
    assert set(j["config"]) == {"workload", "pairs_per_gpu", "read_len", "anchor_len", "sharding", "l2_policy"}
    rf = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "hbm" and rf["unit"] == "GB/s"
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert abs(rf["achieved"] - 76 * 10_000_000 / (rf["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * rf["achieved"]
    assert rf["ms_per_launch"] < j["ms_per_step"] and 0.5 < rf["frac"] < 1.0          # the scan is below the step, the step below peak
    assert abs(j["value"] - j["config"]["pairs_per_gpu"] / (j["ms_per_step"] * 1e-3)) < 1e-3 * j["value"]
    e = j["e2e"]
    assert e["h2d_bytes_per_step"] == 760_000_000 and e["format"] == "wire" and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < j["value"]
    assert e["h2d_only_ceiling"]["pairs_per_s"] >= 0.95 * e["value"]                   # nothing beats the copy alone
    assert j["gpu_launches"] == 6 * j["steps"]
    assert j["parity"]["equal"] is True and j["parity"]["pairs"] == 1_000_000 and j["parity"]["records"] > 1000
    c = j["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    gp = j["genome_pass"]
    assert gp["reads_back_where_drawn"] == gp["reads"] and gp["parity"]["equal"] is True and gp["value"] > 0 and 0 < gp["scan_frac_of_hbm_peak"] < 1
    fq = j["fastq_gz"]
    assert fq["value"] > 0 and fq["threads"] >= 1 and fq["single_member_gzip"]["value"] > 0 and fq["plain_text"]["value"] > 0
    assert fq["anchored_reads"] == fq["single_member_gzip"]["anchored_reads"] == fq["plain_text"]["anchored_reads"] > 0

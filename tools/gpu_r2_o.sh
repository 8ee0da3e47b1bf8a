#!/bin/bash
# round 2, GPU call O (1 GPU): whole -m gpu suite and both bench arms at HEAD (default flags, as the driver runs them)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest_gpu.log
tail -6 gpurun_out/r02o_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02o_smoke.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02o_bench_n1_reference.json 2> gpurun_out/r02o_bench_n1_reference.err ) 2> gpurun_out/r02o_time_ref.txt; echo "ref rc=$?"
( time timeout 900 python bench.py > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err ) 2> gpurun_out/r02o_time_bench.txt; echo "bench rc=$?"
cat gpurun_out/r02o_time_ref.txt gpurun_out/r02o_time_bench.txt | grep real
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02o_bench_n1.json").read().strip().splitlines()[-1])
r = j["roofline"]; e = j["e2e"]; f = j["fastq_gz"]; g = j["genome_pass"]
print("value %.4g ms/step %.4f frac %.4f" % (j["value"], j["ms_per_step"], r["frac"]), "e2e %.4g" % e["value"], "fastq bgzf %.3g gz %.3g plain %.3g" % (f["value"], f["single_member_gzip"]["value"], f["plain_text"]["value"]),
      "genome", {k: g[k] for k in ("value", "ms_per_call", "passes", "scan_frac_of_hbm_peak", "reads_back_where_drawn")}, g["parity"]["equal"], "parity", j["parity"]["equal"], "cpu", j["cpu_baseline"]["value"], "launches", j["gpu_launches"])
PY
tail -n 3 gpurun_out/r02o_bench_n1.err

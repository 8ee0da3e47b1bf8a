#!/bin/bash
# round 2, GPU call S: sample grid at P0 = 4, tail samples behind shared uniform branches: parity tests, bench, sweep, launch list + ncu --set full of the seed scan
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_northstar.py -m gpu -x -q > gpurun_out/r02s_pytest_parity.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02s_pytest_parity.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?"
timeout 600 python tools/scan_sweep.py --configs 12:3:768,12:3:640,12:3:512 > gpurun_out/r02s_scan_sweep.jsonl 2> gpurun_out/r02s_scan_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02s_scan_sweep.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02s.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 > gpurun_out/r02s_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed_scan -c 1 -s 3 -o gpurun_out/seed_scan_r02s python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 > gpurun_out/r02s_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
tail -3 gpurun_out/r02s_pytest_parity.log
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02s_bench_n1.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value %.4g ms/step %.4f frac %.4f scan ms %.4f stages %s e2e %.4g parity %s per_step %s" % (j["value"], j["ms_per_step"], r["frac"], r["ms_per_launch"], r["stage_ms_per_step"], j["e2e"]["value"], j["parity"], j["per_step"]))
PY

#!/bin/bash
# round 2, GPU call Y (1 GPU): state at the end of the round -- whole -m gpu suite, smoke, both bench arms with default flags, launch list + ncu --set full of the seed scan, long-read bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02y_pytest_gpu.log
tail -5 gpurun_out/r02y_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02y_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_n1_reference.json 2> gpurun_out/r02y_bench_n1_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --read-len 300 --pairs 5000000 --steps 20 --warmup 3 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02y_bench_2x300.json 2> gpurun_out/r02y_bench_2x300.err; echo "bench 2x300 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02y.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02y_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed_scan -c 1 -s 3 -o gpurun_out/seed_scan_r02y python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02y_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
python - <<'PY'
import json
for f in ["n1", "n1_reference", "2x300"]:
    try:
        j = json.loads(open("gpurun_out/r02y_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "frac", r.get("frac"), "stages", r.get("stage_ms_per_step"), "e2e", (j.get("e2e") or {}).get("value"),
              "fastq", (j.get("fastq_gz") or {}).get("value"), "parity", (j.get("parity") or {}).get("equal"), "per_step", j.get("per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# round 2 (1 GPU): verify stage with its sample loop unrolled: parity tests, bench, 2x300
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_northstar.py -m gpu -x -q > gpurun_out/r02vu_pytest_parity.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02vu_pytest_parity.log
tail -3 gpurun_out/r02vu_pytest_parity.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02vu_bench_n1.json 2> gpurun_out/r02vu_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --read-len 300 --pairs 5000000 --steps 20 --warmup 3 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02vu_bench_2x300.json 2> gpurun_out/r02vu_bench_2x300.err; echo "bench 2x300 rc=$?"
python - <<'PY'
import json
for f in ["n1", "2x300"]:
    j = json.loads(open("gpurun_out/r02vu_bench_%s.json" % f).read().strip().splitlines()[-1])
    r = j.get("roofline") or {}
    print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "frac", r.get("frac"), "serial", r.get("serial_ms_per_step"), "stages", r.get("stage_ms_per_step"), "parity", (j.get("parity") or {}).get("equal"))
PY

#!/bin/bash
# round 2, GPU call Q: offset sample grid (17 instead of 18 samples per 150-base read, no probes for samples the batch lacks): parity suite, bench, scan sweep
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02q_pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02q_pytest_gpu.log
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02q_bench_n1.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value %.4g ms/step %.4f frac %.4f scan ms %.4f stages %s e2e %.4g parity %s per_step %s" % (j["value"], j["ms_per_step"], r["frac"], r["ms_per_launch"], r["stage_ms_per_step"], j["e2e"]["value"], j["parity"], j["per_step"]))
PY

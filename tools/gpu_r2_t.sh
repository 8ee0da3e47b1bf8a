#!/bin/bash
# round 2, GPU call T: reads up to 512 bases (long-read scan / verify instances, 512-bit N masks): whole -m gpu suite, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02t_pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/r02t_bench_n1.json 2> gpurun_out/r02t_bench_n1.err; echo "bench rc=$?"
tail -12 gpurun_out/r02t_pytest_gpu.log
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02t_bench_n1.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value %.4g ms/step %.4f frac %.4f scan ms %.4f stages %s e2e %.4g parity %s" % (j["value"], j["ms_per_step"], r["frac"], r["ms_per_launch"], r["stage_ms_per_step"], j["e2e"]["value"], j["parity"]["equal"]))
PY

#!/bin/bash
# round 2, GPU call V (1 GPU): 2x300 CLI test, stream-count sweep of the throughput region
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_stage.py -m gpu -x -q -k "2x300" > gpurun_out/r02v_pytest_2x300.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02v_pytest_2x300.log
tail -3 gpurun_out/r02v_pytest_2x300.log
for s in 3 4 5 6 8 3 4; do
  timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 --parity-pairs 0 --slots $s > gpurun_out/r02v_bench_s${s}.json 2> gpurun_out/r02v_bench_s${s}.err; echo "bench slots $s rc=$?"
  python - <<PY
import json
j = json.loads(open("gpurun_out/r02v_bench_s${s}.json").read().strip().splitlines()[-1])
print("slots ${s}: ms/step %.4f scan %.4f" % (j["ms_per_step"], j["roofline"]["ms_per_launch"]))
PY
done

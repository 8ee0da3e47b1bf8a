#!/bin/bash
# round 2, GPU call W (1 GPU): 2x300 CLI test; experiment: L2 fetch granularity 32 B for the verify stage's quad gathers (AF_L2_FETCH, AF_GATHER_PLAIN)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_stage.py -m gpu -x -q -k "2x300" > gpurun_out/r02w_pytest_2x300.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02w_pytest_2x300.log
tail -3 gpurun_out/r02w_pytest_2x300.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 --parity-pairs 100000 > gpurun_out/r02w_bench_$tag.json 2> gpurun_out/r02w_bench_$tag.err; echo "bench $tag rc=$?"
  grep "cudaLimit" gpurun_out/r02w_bench_$tag.err | head -1
  python - <<PY
import json
j = json.loads(open("gpurun_out/r02w_bench_$tag.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("$tag: ms/step %.4f scan %.4f verify %.4f serial %.4f parity %s" % (j["ms_per_step"], r["ms_per_launch"], r["stage_ms_per_step"]["verify"], r["serial_ms_per_step"], j["parity"]["equal"]))
PY
}
run base AF_X=0
run l2_32 AF_L2_FETCH=32
run l2_32_plain AF_L2_FETCH=32 AF_GATHER_PLAIN=1
run plain AF_GATHER_PLAIN=1
run l2_128 AF_L2_FETCH=128

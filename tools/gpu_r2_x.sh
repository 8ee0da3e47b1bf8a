#!/bin/bash
# round 2, GPU call X (1 GPU): SM partition experiment -- the scan's persistent grid leaves AF_SMALL_SMS SMs to the previous batch's verify stage
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() {
  tag=$1; sms=$2; thr=$3; slots=$4
  AF_SMALL_SMS=$sms timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 --parity-pairs 100000 --scan-threads $thr --slots $slots > gpurun_out/r02x_bench_$tag.json 2> gpurun_out/r02x_bench_$tag.err; rc=$?
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r02x_bench_$tag.json").read().strip().splitlines()[-1])
    r = j["roofline"]
    print("small_sms=$sms threads=$thr slots=$slots: ms/step %.4f scan %.4f verify %.4f extend %.4f serial %.4f parity %s" % (j["ms_per_step"], r["ms_per_launch"], r["stage_ms_per_step"]["verify"], r["stage_ms_per_step"]["extend"], r["serial_ms_per_step"], j["parity"]["equal"]))
except Exception as e:
    print("$tag rc=$rc ERR", e)
PY
}
run a 0 768 3
run b 10 512 2
run c 10 512 3
run d 10 768 2
run e 16 512 2
run f 8 512 2
run g 12 512 3
run h 20 512 2
run i 6 512 2

#!/bin/bash
# round 2 (4 GPUs): bench at N=4 at HEAD, to complete the 1 / 2 / 4 / 8 set
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 4 --steps 100 --warmup 10 --fastq-pairs 0 --genome-bases 0 --no-cpu > gpurun_out/r02end_bench_n4.json 2> gpurun_out/r02end_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02end_bench_n4.json").read().strip().splitlines()[-1])
e = j.get("e2e") or {}
print("n4 value %.4g ms/step %.4f frac %s e2e %s ceil %s parity %s" % (j["value"], j["ms_per_step"], (j.get("roofline") or {}).get("frac"), e.get("value"), (e.get("h2d_only_ceiling") or {}).get("pairs_per_s"), (j.get("parity") or {}).get("equal")))
PY

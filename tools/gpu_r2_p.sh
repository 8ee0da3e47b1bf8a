#!/bin/bash
# round 2, GPU call P (8 GPUs): bench at N=8 with the wire-format e2e, configs[4] with the reader threads shared among the ranks
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err; echo "bench n8 rc=$?"
timeout 900 $TR --master-port 29533 bench.py --gpus 8 --workload singlecell --cells 4000 --pairs-per-cell 5000 > gpurun_out/r02p_bench_singlecell_n8.json 2> gpurun_out/r02p_bench_singlecell_n8.err; echo "singlecell n8 rc=$?"
python - <<'PY'
import json
for f in ["n8", "singlecell_n8"]:
    try:
        j = json.loads(open("gpurun_out/r02p_bench_%s.json" % f).read().strip().splitlines()[-1])
        e = j.get("e2e") or {}
        print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "e2e", e.get("value"), "fmt", e.get("format"), "ceil", (e.get("h2d_only_ceiling") or {}).get("pairs_per_s"), "parity", (j.get("parity") or {}).get("equal"), "sc", j.get("singlecell"))
    except Exception as ex:
        print(f, "ERR", ex)
PY

#!/bin/bash
# round 2, GPU call ZZ (8 GPUs): bench at N=8 at the end-of-round state -- configs[1] weak scaling (p2p exchange, parity vs the oracle), configs[2] (100 M pairs read-sharded), configs[4] (20 M pairs in 4000 cells)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/r02zz_gpus.txt; nproc >> gpurun_out/r02zz_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02zz_bench_n8.json 2> gpurun_out/r02zz_bench_n8.err; echo "bench n8 rc=$?"
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --workload config3 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02zz_bench_config3_n8.json 2> gpurun_out/r02zz_bench_config3_n8.err; echo "config3 n8 rc=$?"
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --workload singlecell --cells 4000 --pairs-per-cell 5000 > gpurun_out/r02zz_bench_singlecell_n8.json 2> gpurun_out/r02zz_bench_singlecell_n8.err; echo "singlecell n8 rc=$?"
python - <<'PY'
import json
for f in ["n8", "config3_n8", "singlecell_n8"]:
    try:
        j = json.loads(open("gpurun_out/r02zz_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        e = j.get("e2e") or {}
        print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "scaling", j.get("scaling"), "frac", r.get("frac"), "e2e", e.get("value"), "ceil", (e.get("h2d_only_ceiling") or {}).get("pairs_per_s"),
              "parity", (j.get("parity") or {}).get("equal"), "sc", j.get("singlecell"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
for f in gpurun_out/r02zz_bench_*.err; do tail -3 $f; done

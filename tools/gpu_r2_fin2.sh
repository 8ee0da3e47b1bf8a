#!/bin/bash
# round 2 (1 GPU): long reads through k_verify_smem where its shared memory fits (W <= 28): parity tests, 2x300 line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire.py -m gpu -x -q > gpurun_out/r02fin2_pytest_parity.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02fin2_pytest_parity.log
tail -3 gpurun_out/r02fin2_pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_stage.py -m gpu -x -q -k "2x300" >> gpurun_out/r02fin2_pytest_parity.log 2>&1; echo "pytest stage rc=$?"
for L in 300 400; do
timeout 600 python bench.py --read-len $L --pairs 5000000 --steps 20 --warmup 3 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02fin2_bench_2x$L.json 2> gpurun_out/r02fin2_bench_2x$L.err; echo "bench 2x$L rc=$?"
done
python - <<'PY'
import json
for f in ["2x300", "2x400"]:
    try:
        j = json.loads(open("gpurun_out/r02fin2_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "frac", r.get("frac"), "stages", r.get("stage_ms_per_step"), "parity", (j.get("parity") or {}).get("equal"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# round 2, HEAD of the round (1 GPU): whole -m gpu suite, smoke, both bench arms, 2x300 / 2x250 / 2x101 lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/r02end_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02end_pytest_gpu.log
tail -4 gpurun_out/r02end_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02end_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02end_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02end_bench_n1_reference.json 2> gpurun_out/r02end_bench_n1_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02end_bench_n1.json 2> gpurun_out/r02end_bench_n1.err; echo "bench rc=$?"
for L in 300 250 101; do
timeout 600 python bench.py --read-len $L --pairs 5000000 --steps 20 --warmup 3 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02end_bench_2x$L.json 2> gpurun_out/r02end_bench_2x$L.err; echo "bench 2x$L rc=$?"
done
python - <<'PY'
import json
for f in ["n1", "2x300", "2x250", "2x101"]:
    try:
        j = json.loads(open("gpurun_out/r02end_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "frac", r.get("frac"), "stages", r.get("stage_ms_per_step"), "e2e", (j.get("e2e") or {}).get("value"),
              "fastq", (j.get("fastq_gz") or {}).get("value"), "parity", (j.get("parity") or {}).get("equal"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
# round 2, GPU call U: CLI test with 2x300 reads; does a smaller scan CTA let the small kernels of the previous batch share the SMs? (--scan-threads, --slots)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_stage.py -m gpu -x -q -k "2x300" > gpurun_out/r02u_pytest_2x300.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02u_pytest_2x300.log
tail -5 gpurun_out/r02u_pytest_2x300.log
for cfg in 768:3 640:3 512:3 512:4 384:3; do
  t=${cfg%%:*}; s=${cfg##*:}
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --fastq-pairs 0 --genome-bases 0 --parity-pairs 0 --scan-threads $t --slots $s > gpurun_out/r02u_bench_t${t}_s${s}.json 2> gpurun_out/r02u_bench_t${t}_s${s}.err; echo "bench $cfg rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02u_bench_t*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        r = j["roofline"]
        print(f, "ms/step %.4f scan %.4f serial %.4f" % (j["ms_per_step"], r["ms_per_launch"], r["serial_ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY

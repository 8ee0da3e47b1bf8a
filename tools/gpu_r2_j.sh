#!/bin/bash
# round 2, GPU call J (1 GPU): whole -m gpu suite, smoke, both bench arms, launch list + ncu of the seed scan at HEAD, k' / thread sweep, long anchors, FASTQ leg on 4 M pairs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest_gpu.log
tail -8 gpurun_out/r02j_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02j_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_n1_reference.json 2> gpurun_out/r02j_bench_n1_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --fastq-pairs 4000000 > gpurun_out/r02j_bench_fastq4m.json 2> gpurun_out/r02j_bench_fastq4m.err; echo "bench fastq4m rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 > gpurun_out/r02j_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed_scan -c 1 -s 3 -o gpurun_out/r02j_seed_scan python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 > gpurun_out/r02j_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
timeout 600 python tools/scan_sweep.py --configs 12:3:768,12:3:640,12:3:512,13:3:768,13:3:640 > gpurun_out/r02j_scan_sweep.jsonl 2> gpurun_out/r02j_scan_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02j_scan_sweep.jsonl
for a in 10000 20000 40000; do
  timeout 600 python bench.py --anchor-len $a --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02j_bench_anchor$a.json 2> gpurun_out/r02j_bench_anchor$a.err; echo "anchor $a rc=$?"
done
python - <<'PY'
import json
for f in ["n1", "n1_reference", "fastq4m", "anchor10000", "anchor20000", "anchor40000"]:
    try:
        j = json.loads(open("gpurun_out/r02j_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        print(f, "value %.4g" % j["value"], "ms/step %.4f" % j["ms_per_step"], "frac", r.get("frac"), "stages", r.get("stage_ms_per_step"), "e2e", (j.get("e2e") or {}).get("value"),
              "fastq", {k: (j.get("fastq_gz") or {}).get(k) for k in ("value", "pairs")}, "gz", ((j.get("fastq_gz") or {}).get("single_member_gzip") or {}).get("value"), "parity", j.get("parity"), "per_step", j.get("per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY

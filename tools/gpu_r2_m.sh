#!/bin/bash
# round 2, GPU call M: Bloom filter for long anchors -- parity suite, then bench per anchor length (and the headline again: the plain scan must not move)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -12 gpurun_out/r02m_pytest.log
for a in 6783 20000 40000 100000; do
  timeout 600 python bench.py --anchor-len $a --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02m_bench_anchor$a.json 2> gpurun_out/r02m_bench_anchor$a.err; echo "anchor $a rc=$?"
done
AF_NO_BLOOM=1 timeout 600 python bench.py --anchor-len 100000 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02m_bench_anchor100000_nobloom.json 2> gpurun_out/r02m_bench_anchor100000_nobloom.err; echo "anchor 100000 nobloom rc=$?"
python - <<'PY'
import json
for f in ["anchor6783", "anchor20000", "anchor40000", "anchor100000", "anchor100000_nobloom"]:
    try:
        j = json.loads(open("gpurun_out/r02m_bench_%s.json" % f).read().strip().splitlines()[-1])
        r = j.get("roofline") or {}
        print(f, "ms/step %.4f" % j["ms_per_step"], "frac %.3f" % r.get("frac"), "stages", {k: round(v, 4) for k, v in r.get("stage_ms_per_step").items()}, "parity", (j.get("parity") or {}).get("equal"), "per_step", j.get("per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY

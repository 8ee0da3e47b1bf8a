#!/bin/bash
# round 2, GPU call L: wire format (76 B per pair over PCIe) -- tests, then bench with e2e in both formats; ingest after the CRC change
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_wire.py tests/test_bench_contract.py tests/test_gpu_genome.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -15 gpurun_out/r02l_pytest.log
timeout 900 python bench.py --no-cpu --fastq-pairs 4000000 > gpurun_out/r02l_bench_n1_wire.json 2> gpurun_out/r02l_bench_n1_wire.err; echo "bench wire rc=$?"
timeout 900 python bench.py --no-cpu --fastq-pairs 0 --e2e-format tiles > gpurun_out/r02l_bench_n1_tiles.json 2> gpurun_out/r02l_bench_n1_tiles.err; echo "bench tiles rc=$?"
timeout 600 python tools/ingest_bench.py --pairs 4000000 --threads 1,4,16 --out gpurun_out/r02l_ingest.json > /dev/null 2> gpurun_out/r02l_ingest.err; echo "ingest rc=$?"; tail -12 gpurun_out/r02l_ingest.err
python - <<'PY'
import json
for f in ["n1_wire", "n1_tiles"]:
    try:
        j = json.loads(open("gpurun_out/r02l_bench_%s.json" % f).read().strip().splitlines()[-1])
        e = j.get("e2e") or {}
        print(f, "value %.4g" % j["value"], "e2e %.4g" % e.get("value"), "ms %.3f" % e.get("ms_per_step"), "h2d", e.get("h2d_bytes_per_step"), "ceil", e.get("h2d_only_ceiling"), "launches", j.get("gpu_launches"),
              "fastq", {k: (j.get("fastq_gz") or {}).get(k) for k in ("value", "pairs")}, "gz", ((j.get("fastq_gz") or {}).get("single_member_gzip") or {}).get("value"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -n 3 gpurun_out/r02l_bench_n1_wire.err

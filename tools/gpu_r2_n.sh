#!/bin/bash
# round 2, GPU call N (host-side measurement on the 16-core box): single-member gzip decoded by several workers vs the serial path
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python tools/ingest_bench.py --pairs 4000000 --threads 8,16 --real-gzip 1 --out gpurun_out/r02n_ingest.json > /dev/null 2> gpurun_out/r02n_ingest.err; echo "ingest rc=$?"; cat gpurun_out/r02n_ingest.err | tail -20
AF_GZIP_SERIAL=1 timeout 600 python tools/ingest_bench.py --pairs 4000000 --threads 16 > /dev/null 2> gpurun_out/r02n_ingest_serial.err; grep gzip gpurun_out/r02n_ingest_serial.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1])
f = j["fastq_gz"]
print("bgzf", f["value"], "gzip", f["single_member_gzip"]["value"], "plain", f["plain_text"]["value"], "e2e", j["e2e"]["value"])
PY

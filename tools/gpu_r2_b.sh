#!/bin/bash
# round 2, GPU call B: new stage / ingest / bench paths
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_stage.py tests/test_bench_contract.py tests/test_gpu_parity.py tests/test_gpu_exchange.py -m gpu -x -q -s > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
tail -12 gpurun_out/r02b_pytest_gpu.log
timeout 600 python tools/ingest_bench.py --pairs 2000000 --threads 1,2,4,8,16 --out gpurun_out/r02b_ingest.json > /dev/null 2> gpurun_out/r02b_ingest.err; echo "ingest rc=$?"; cat gpurun_out/r02b_ingest.err | tail -16
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --workload singlecell --cells 4000 --pairs-per-cell 5000 > gpurun_out/r02b_bench_singlecell_n1.json 2> gpurun_out/r02b_bench_singlecell_n1.err; echo "singlecell rc=$?"
timeout 600 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02b_bench_config3_n1.json 2> gpurun_out/r02b_bench_config3_n1.err; echo "config3 rc=$?"
cat gpurun_out/r02b_bench_n1.json | head -c 6000; echo
cat gpurun_out/r02b_bench_singlecell_n1.json | head -c 3000; echo
cat gpurun_out/r02b_bench_config3_n1.json | head -c 1500; echo
tail -3 gpurun_out/r02b_bench_n1.err gpurun_out/r02b_bench_singlecell_n1.err gpurun_out/r02b_bench_config3_n1.err

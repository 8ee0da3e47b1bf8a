#!/bin/bash
# round 2, GPU call H (2 GPUs): the two-rank tests that a 1-GPU box skips, bench at N=2 (p2p exchange, parity vs oracle), config3 and single-cell at N=2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 900 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_stage.py -m gpu -q -rs > gpurun_out/r02h_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest_2gpu.log
tail -6 gpurun_out/r02h_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r02h_bench_n2_reference.json 2> gpurun_out/r02h_bench_n2_reference.err; echo "ref n2 rc=$?"
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload config3 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02h_bench_config3_n2.json 2> gpurun_out/r02h_bench_config3_n2.err; echo "config3 n2 rc=$?"
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --workload singlecell --cells 4000 --pairs-per-cell 5000 > gpurun_out/r02h_bench_singlecell_n2.json 2> gpurun_out/r02h_bench_singlecell_n2.err; echo "singlecell n2 rc=$?"
for f in n2 n2_reference config3_n2 singlecell_n2; do echo "== $f"; head -c 2500 gpurun_out/r02h_bench_$f.json; echo; tail -3 gpurun_out/r02h_bench_$f.err; done

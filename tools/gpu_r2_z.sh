#!/bin/bash
# round 2, GPU call Z (2 GPUs): the two-rank tests that a 1-GPU box skips and bench at N=2 (p2p exchange, parity vs the oracle), at the end-of-round state
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/r02z_gpus.txt
timeout 900 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_stage.py -m gpu -q -rs > gpurun_out/r02z_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest_2gpu.log
tail -6 gpurun_out/r02z_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02z_bench_n2.json 2> gpurun_out/r02z_bench_n2.err; echo "bench n2 rc=$?"
head -c 3000 gpurun_out/r02z_bench_n2.json; echo; tail -3 gpurun_out/r02z_bench_n2.err

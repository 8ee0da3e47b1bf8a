#!/bin/bash
# round 2, GPU call I: genome pass of the contiguity filter -- parity tests, then a first timing on a 3.1 Gbp synthetic genome
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_genome.py -m gpu -x -q > gpurun_out/r02i_pytest_genome.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest_genome.log
tail -30 gpurun_out/r02i_pytest_genome.log
timeout 600 python tools/genome_bench.py --bases 3100000000 --reads 1000 > gpurun_out/r02i_genome_bench.json 2> gpurun_out/r02i_genome_bench.err; echo "genome bench rc=$?"
cat gpurun_out/r02i_genome_bench.json; tail -5 gpurun_out/r02i_genome_bench.err

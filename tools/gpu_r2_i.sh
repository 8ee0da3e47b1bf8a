#!/bin/bash
# round 2, GPU call I: genome pass of the contiguity filter -- parity tests, timing on a 3.1 Gbp synthetic genome, ncu of its scan kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_genome.py -m gpu -x -q > gpurun_out/r02i_pytest_genome.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest_genome.log
tail -30 gpurun_out/r02i_pytest_genome.log
for n in 1000 10000; do
  timeout 600 python tools/genome_bench.py --bases 3100000000 --reads $n > gpurun_out/r02i_genome_bench_$n.json 2> gpurun_out/r02i_genome_bench_$n.err; echo "genome bench $n rc=$?"
  cat gpurun_out/r02i_genome_bench_$n.json; tail -5 gpurun_out/r02i_genome_bench_$n.err
done
timeout 900 python tools/genome_bench.py --bases 100000000 --reads 1000 --oracle-bases 100000000 > gpurun_out/r02i_genome_bench_oracle.json 2> gpurun_out/r02i_genome_bench_oracle.err; echo "oracle bench rc=$?"
cat gpurun_out/r02i_genome_bench_oracle.json; tail -5 gpurun_out/r02i_genome_bench_oracle.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_genome_scan -c 1 -s 2 -o gpurun_out/r02i_genome_scan python tools/genome_bench.py --bases 3100000000 --reads 300 --repeat 1 > gpurun_out/r02i_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02i_genome_launches.csv python tools/genome_bench.py --bases 3100000000 --reads 300 --repeat 1 > gpurun_out/r02i_ncu_launches.log 2>&1; echo "ncu launches rc=$?"

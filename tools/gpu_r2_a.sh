#!/bin/bash
# round 2, GPU call A: tool probe, the whole -m gpu suite, bench (new default route vs the round-1 route), ncu
mkdir -p gpurun_out
{ echo "== command -v bwa samtools blat bedtools wgsim minimap2 bgzip pigz =="; for t in bwa samtools blat bedtools wgsim minimap2 bgzip pigz; do printf "%s: " $t; command -v $t || echo absent; done;
  echo "== nproc =="; nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread"; free -g | head -2; python -c "import pysam" 2>&1 | tail -1; } > gpurun_out/r02_probe_tools.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 --scan-mode 7 --no-cpu --no-e2e > gpurun_out/r02a_bench_n1_mode7.json 2> gpurun_out/r02a_bench_n1_mode7.err; echo "bench7 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 > gpurun_out/r02a_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_seed_scan -c 1 -o gpurun_out/r02a_seed_scan python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 > gpurun_out/r02a_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
cat gpurun_out/r02a_bench_n1.json | head -c 3000
echo
cat gpurun_out/r02a_bench_n1_mode7.json | head -c 2500

#!/bin/bash
# round 2, GPU call C: whole -m gpu suite with the new reader / stage / scan tail pool, scan pool A/B, ncu of the small kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -6 gpurun_out/r02c_pytest_gpu.log
for m in 20 23 24 25 26; do
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --parity-pairs 0 --scan-mode $m > gpurun_out/r02c_bench_pool$m.json 2> gpurun_out/r02c_bench_pool$m.err
  python -c "
import json; j=json.load(open('gpurun_out/r02c_bench_pool$m.json')); r=j['roofline']
print('mode $m', 'ms/step %.4f'%j['ms_per_step'], 'scan %.4f'%r['ms_per_launch'], 'frac %.4f'%r['frac'], 'serial %.4f'%r['serial_ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum --clock-control none -k regex:k_seed_scan -c 4 --csv --log-file gpurun_out/r02c_scan_cycles_pool.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 > /dev/null 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum --clock-control none -k regex:k_seed_scan -c 4 --csv --log-file gpurun_out/r02c_scan_cycles_static.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 --scan-mode 20 > /dev/null 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_verify_smem|k_extend" -c 2 -o gpurun_out/r02c_verify_extend python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 > gpurun_out/r02c_ncu_ve.log 2>&1; echo "ncu3 rc=$?"
cat gpurun_out/r02c_scan_cycles_pool.csv | tail -30
cat gpurun_out/r02c_scan_cycles_static.csv | tail -30

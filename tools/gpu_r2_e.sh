#!/bin/bash
# round 2, GPU call E: candidate-stream path (k_seed_scan<EMIT> + k_tail) -- parity suite, then A/B bench against the six-kernel path
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
AF_STREAM=1 timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02e_smoke.log
AF_STREAM=1 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exchange.py -m gpu -x -q > gpurun_out/r02e_pytest_parity.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02e_pytest_parity.log
for m in 12 13; do
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu --no-e2e --parity-pairs 1000000 --scan-mode $m > gpurun_out/r02e_bench_mode$m.json 2> gpurun_out/r02e_bench_mode$m.err
  python -c "
import json; j=json.load(open('gpurun_out/r02e_bench_mode$m.json')); r=j['roofline']
print('mode $m', 'ms/step %.4f'%j['ms_per_step'], 'scan %.4f'%r['ms_per_launch'], 'frac %.4f'%r['frac'], 'serial %.4f'%r['serial_ms_per_step'], r['stage_ms_per_step'], j.get('parity'), j.get('per_step'))"
done
AF_STREAM=1 python tools/tail_timing.py 2>&1 | tail -8

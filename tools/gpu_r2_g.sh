#!/bin/bash
# round 2, GPU call G: ncu --set full of k_tail (candidate-stream path) with source counters
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tail -c 1 -s 4 -o gpurun_out/r02g_tail python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 --scan-mode 12 > gpurun_out/r02g_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02g_tail.ncu-rep

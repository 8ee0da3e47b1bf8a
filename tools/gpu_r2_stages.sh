#!/bin/bash
# round 2 (1 GPU): ncu --set full of the five small stages at HEAD (one launch each)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_(flag_scatter|verify_smem|sel_scatter|extend|hit_scatter)" -c 5 -s 15 -o gpurun_out/stages_r02end python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --slots 1 --parity-pairs 0 --fastq-pairs 0 --genome-bases 0 > gpurun_out/r02end_ncu_stages.log 2>&1; echo "ncu stages rc=$?"
tail -3 gpurun_out/r02end_ncu_stages.log

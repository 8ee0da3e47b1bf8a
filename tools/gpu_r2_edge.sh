#!/bin/bash
# round 2 (1 GPU): edge batch sizes and canary zones with the long-read kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_batch or inside_their_buffers" > gpurun_out/r02edge_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02edge_pytest.log
tail -15 gpurun_out/r02edge_pytest.log

#!/bin/bash
# scratch: short validation of the current working tree on one GPU
mkdir -p gpurun_out
echo "=== gpu tests"; timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/test_gpu.log
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 400 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200

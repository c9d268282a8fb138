#!/bin/bash
# scratch script for the A/B experiment of the day (last use: the driver's GPU test command + smoke on the final tree)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=12 2>&1 | tail -22 | tee gpurun_out/test_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log

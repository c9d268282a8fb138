#!/bin/bash
mkdir -p gpurun_out
F5_GEMM_TRACE=1 timeout 200 python tools/gemm_trace.py 2>&1 | tail -8 | tee gpurun_out/gemm_trace.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches2.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
tail -1 gpurun_out/ncu_launch_run.log | cut -c1-200

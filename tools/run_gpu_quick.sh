#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
T=$PWD/f5_tts_b200/libf5tts_b200_trace.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -3 | tee gpurun_out/test_attn.log
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
F5_LIB=$T F5_ATTN_TRACE=1 timeout 300 python tools/attn_trace.py 2>&1 | tail -6 | tee gpurun_out/attn_trace.log
timeout 900 python -m pytest tests/test_gpu_sample.py -x -q 2>&1 | tail -3 | tee gpurun_out/test_sample.log
for W in cfg2 cfg3 cfg5; do STEP_WORKLOAD=$W timeout 600 python tools/step_time.py 2>&1 | tail -1 | tee -a gpurun_out/step_time_persist.log; done

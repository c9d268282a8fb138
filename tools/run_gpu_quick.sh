#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv or linear" 2>&1 | tail -3 | tee gpurun_out/test_gemm.log
for V in "" _eps0; do
  echo "lib$V" | tee -a gpurun_out/gemm_eps.log
  F5_LIB=$PWD/f5_tts_b200/libf5tts_b200$V.so SWEEP_M=1876,15008 timeout 600 python tools/gemm_sweep.py 2>&1 | cut -c1-260 | tee -a gpurun_out/gemm_eps.log
done
for V in "" _eps0 "" _eps0; do
  F5_LIB=$PWD/f5_tts_b200/libf5tts_b200$V.so timeout 600 python tools/step_time.py 2>&1 | tail -1 | sed "s/^/lib$V /" | tee -a gpurun_out/step_time_eps.log
done

#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
for V in "" _nopf "" _nopf; do
  for W in cfg2; do F5_LIB=$PWD/f5_tts_b200/libf5tts_b200$V.so STEP_WORKLOAD=$W timeout 600 python tools/step_time.py 2>&1 | tail -1 | sed "s/^/lib$V /" | tee -a gpurun_out/step_time_pf.log; done
done
for V in "" _nopf; do
  for W in cfg3 cfg5; do F5_LIB=$PWD/f5_tts_b200/libf5tts_b200$V.so STEP_WORKLOAD=$W timeout 600 python tools/step_time.py 2>&1 | tail -1 | sed "s/^/lib$V /" | tee -a gpurun_out/step_time_pf.log; done
done
timeout 600 python -m pytest tests/test_gpu_sample.py -x -q 2>&1 | tail -2

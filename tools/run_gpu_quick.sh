#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short 2>&1 | tail -6 | tee gpurun_out/quick_tests.log
timeout 600 python -m pytest tests/test_gpu_sample.py -q -m gpu --tb=short 2>&1 | tail -8 | tee gpurun_out/quick_tests_sample.log
for pdl in 1 0; do
  echo "=== F5_PDL=$pdl"
  F5_PDL=$pdl timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_pdl$pdl.json | cut -c1-330
  tail -3 gpurun_out/bench.err
done

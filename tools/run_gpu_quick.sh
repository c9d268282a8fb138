#!/bin/bash
# scratch: A/B experiments of the current working tree on one GPU
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short 2>&1 | tail -8 | tee gpurun_out/quick_tests.log
echo "=== gemm sweep"; SWEEP_M=1876,15008 timeout 400 python tools/gemm_sweep.py 2>&1 | tail -20 | tee gpurun_out/gemm_sweep.log
echo "=== bench"; timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_quick.json | cut -c1-230
echo "=== bench cfg4"; timeout 500 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_cfg4.json | cut -c1-230
echo "=== sample tests"; timeout 600 python -m pytest tests/test_gpu_sample.py -q -m gpu --tb=short 2>&1 | tail -5 | tee gpurun_out/quick_tests_sample.log

#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -3 | tee gpurun_out/test_attn.log
timeout 900 python -m pytest tests/test_gpu_infer.py -x -q 2>&1 | tail -5 | tee gpurun_out/test_infer.log
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300

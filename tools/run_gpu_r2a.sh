#!/bin/bash
# Round 2, visit A: new attention kernel (P in TMEM, FMA-pipe exponentials) parity + timing, fused-norm GEMMs, tile skipping,
# exact_varlen, new parity tests, bench line.  Every stage runs in its own process (a device trap poisons a CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
echo "=== attention parity (kernel tests)"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" -s 2>&1 | grep -v "^$" | tail -15 | tee gpurun_out/test_attn.log
echo "=== attention timing: production build"
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
for P in 0 2 4 5; do
  echo "=== attention timing: trace build POLY=$P"
  F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_trace.so F5_ATTN_POLY=$P timeout 300 python tools/attn_bench.py 2>&1 | head -2 | tee gpurun_out/attn_bench_poly$P.log
done
echo "=== attention timing: trace build POLY=3 turnstile off"
F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_trace.so F5_ATTN_TURNSTILE=0 timeout 300 python tools/attn_bench.py 2>&1 | head -2 | tee gpurun_out/attn_bench_ts0.log
echo "=== GEMM tests (incl. fused row norm)"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not attention" -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/test_gpu_kernels.log
echo "=== sampler tests"
timeout 1200 python -m pytest tests/test_gpu_sample.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/test_gpu_sample.log
echo "=== API tests"
timeout 900 python -m pytest tests/test_gpu_infer.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -20 | tee gpurun_out/test_gpu_infer.log
echo "=== full-size parity tests"
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/test_gpu_fullsize.log
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600
tail -5 gpurun_out/bench.err
echo "=== gemm sweep (M = 1876)"
SWEEP_M=1876 timeout 600 python tools/gemm_sweep.py 2>&1 | tail -8 | tee gpurun_out/gemm_sweep.log

#!/bin/bash
# Round 2, visit G: attention with the reference folded into Q K^T — parity, timing, phase trace; host profile of the e2e call.
mkdir -p gpurun_out
T=$PWD/f5_tts_b200/libf5tts_b200_trace.so
echo "=== attention tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -15 | tee gpurun_out/test_attn.log
echo "=== attention timing (production)"
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
echo "=== attention phase trace"
F5_LIB=$T F5_ATTN_TRACE=1 timeout 300 python tools/attn_trace.py 2>&1 | tail -6 | tee gpurun_out/attn_trace.log
echo "=== turnstile off / poly sweep (trace build)"
F5_LIB=$T F5_ATTN_TURNSTILE=0 timeout 300 python tools/attn_bench.py 2>&1 | head -3 | tee gpurun_out/attn_bench_ts0.log
for P in 0 2 4; do
  F5_LIB=$T F5_ATTN_POLY=$P timeout 300 python tools/attn_bench.py 2>&1 | head -3 | tee gpurun_out/attn_bench_poly$P.log
done
echo "=== step time"
timeout 300 python tools/step_time.py 2>&1 | tail -1 | tee gpurun_out/step_time.log
echo "=== sampler parity (quick)"
timeout 900 python -m pytest tests/test_gpu_sample.py -x -q 2>&1 | tail -4 | tee gpurun_out/test_sample.log
echo "=== e2e host profile"
timeout 300 python tools/e2e_profile.py 2>&1 | head -70 | tee gpurun_out/e2e_profile.log

#!/bin/bash
# Round 2, visit F: in-situ decomposition of the cfg2 step (instrumented build: one kernel class removed per run) and the
# per-phase cycle split of the attention softmax warpgroups.
mkdir -p gpurun_out
T=$PWD/f5_tts_b200/libf5tts_b200_trace.so
echo "=== attention phase trace"
F5_LIB=$T F5_ATTN_TRACE=1 timeout 300 python tools/attn_trace.py 2>&1 | tail -4 | tee gpurun_out/attn_trace.log
echo "=== attention timing (production)"
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
echo "=== step decomposition"
for S in none norm attn qkv out ff1 ff2 conv; do
  F5_LIB=$T F5_DIAG_SKIP=$S timeout 300 python tools/step_time.py 2>&1 | tail -1 | sed "s/^/skip=$S /" | tee -a gpurun_out/step_decomp.log
done
echo "=== pipe microbenchmarks"
timeout 120 tools/microbench/pipes 2>&1 | tee gpurun_out/pipes.log
timeout 120 tools/microbench/mufu 2>&1 | tail -30 | tee gpurun_out/mufu.log
timeout 120 tools/microbench/softmax_mix 2>&1 | tee gpurun_out/softmax_mix.log

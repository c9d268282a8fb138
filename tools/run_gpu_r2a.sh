#!/bin/bash
# Round 2, visit A: new attention kernel (P in TMEM, FMA-pipe exponentials) parity + timing, fused-norm GEMMs, tile skipping,
# exact_varlen, new parity tests, bench line.  Every stage runs in its own process (a device trap poisons a CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
echo "=== TS-form MMA layout probe"
timeout 60 tools/microbench/ts_mma 2>&1 | tail -12 | tee gpurun_out/ts_mma.log
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
echo "=== gemm sweep, experiment build: reduce-add epilogue on two column groups"
F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_eg2.so SWEEP_M=1876,15008 timeout 600 python tools/gemm_sweep.py 2>&1 | grep -E "out|FF2" | tail -8 | tee gpurun_out/gemm_sweep_eg2.log
SWEEP_M=15008 timeout 600 python tools/gemm_sweep.py 2>&1 | tail -4 | tee gpurun_out/gemm_sweep_15008.log
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_run.log 2>&1
tail -1 gpurun_out/ncu_launch_run.log | cut -c1-200
echo "=== ncu full: gemm (one block's four GEMMs)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 5 -o gpurun_out/prof_gemm -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_gemm_run.log 2>&1
tail -2 gpurun_out/ncu_gemm_run.log | cut -c1-200
echo "=== ncu full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd" -s 12 -c 2 -o gpurun_out/prof_attn -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_attn_run.log 2>&1
tail -2 gpurun_out/ncu_attn_run.log | cut -c1-200
echo "=== ncu full: bandwidth kernels (mel STFT, ISTFT, dwconv+LN, GRN, CFG+Euler, row norm)"
timeout 900 ncu --set full --clock-control none -k "regex:mel_stft|istft_frames|istft_ola|dwconv7_ln|grn_sumsq|grn_apply|cfg_euler|row_norm|ln_affine" -s 30 -c 30 -o gpurun_out/prof_bw -f \
  python tools/ncu_bw.py > gpurun_out/ncu_bw_run.log 2>&1
tail -2 gpurun_out/ncu_bw_run.log | cut -c1-200
ls -la gpurun_out | tail -30

#!/bin/bash
# Round 2, visit C: NORMA single-publish fix, attention single code path, fixed tests, phase trace, epilogue experiment.
mkdir -p gpurun_out
echo "=== kernel tests: attention + fused norm"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention or fused" 2>&1 | tail -3 | tee gpurun_out/test_kernels_c.log
echo "=== attention timing"
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
echo "=== attention phase trace (trace build)"
F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_trace.so F5_ATTN_TRACE=1 timeout 300 python tools/attn_trace.py 2>&1 | tail -6 | tee gpurun_out/attn_trace.log
echo "=== sampler tests"
timeout 1200 python -m pytest tests/test_gpu_sample.py -q -m gpu -s 2>&1 | grep -E "^\[|passed|failed|FAILED|assert" | cut -c1-220 | tail -40 | tee gpurun_out/test_gpu_sample.log
echo "=== API tests"
timeout 900 python -m pytest tests/test_gpu_infer.py -q -m gpu -s 2>&1 | grep -E "^\[|passed|failed|FAILED|assert|Error" | cut -c1-220 | tail -20 | tee gpurun_out/test_gpu_infer.log
echo "=== cfg2 full-size parity"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s -k cfg2 2>&1 | grep -E "step|final|passed|failed|FAILED|assert" | cut -c1-220 | tail -12 | tee gpurun_out/test_gpu_cfg2.log
echo "=== bench"
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
echo "=== gemm sweep, experiment build: reduce-add epilogue on two column groups"
F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_eg2.so SWEEP_M=1876,15008 timeout 600 python tools/gemm_sweep.py 2>&1 | grep -E "out|FF2" | tail -8 | cut -c1-260 | tee gpurun_out/gemm_sweep_eg2.log
SWEEP_M=15008 timeout 600 python tools/gemm_sweep.py 2>&1 | tail -4 | cut -c1-260 | tee gpurun_out/gemm_sweep_15008.log
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_run.log 2>&1
echo "=== ncu full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd" -s 12 -c 1 -o gpurun_out/prof_attn -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_attn_run.log 2>&1
echo "=== ncu full: mel"
timeout 600 ncu --set full --clock-control none -k "regex:mel_stft" -c 2 -o gpurun_out/prof_mel -f python tools/ncu_bw.py > gpurun_out/ncu_mel_run.log 2>&1
ls gpurun_out | wc -l

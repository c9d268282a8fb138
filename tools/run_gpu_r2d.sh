#!/bin/bash
# Round 2, visit D: separate norm kernels again (NORMA measured slower), linked FF1->FF2, attention with streamed P.
mkdir -p gpurun_out
echo "=== pipe microbenchmark"
timeout 120 tools/microbench/pipes 2>&1 | tee gpurun_out/pipes.log
echo "=== kernel tests: attention + linked GEMMs + pair tiles"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention or linked or cta_pair" 2>&1 | tail -4 | tee gpurun_out/test_kernels_d.log
echo "=== attention timing"
timeout 300 python tools/attn_bench.py 2>&1 | tail -8 | tee gpurun_out/attn_bench_prod.log
for P in 0 2 4; do
  F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_trace.so F5_ATTN_POLY=$P timeout 300 python tools/attn_bench.py 2>&1 | head -2 | tee gpurun_out/attn_bench_poly$P.log
done
echo "=== attention phase trace (trace build)"
F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_trace.so F5_ATTN_TRACE=1 timeout 300 python tools/attn_trace.py 2>&1 | tail -4 | tee gpurun_out/attn_trace.log
echo "=== FF1 -> FF2 link, isolated"
timeout 300 python tools/link_bench.py 2>&1 | tail -4 | tee gpurun_out/link_bench.log
echo "=== sampler tests"
timeout 1200 python -m pytest tests/test_gpu_sample.py -q -m gpu -s 2>&1 | grep -E "^\[|passed|failed|FAILED|assert" | cut -c1-220 | tail -30 | tee gpurun_out/test_gpu_sample.log
echo "=== bench A/B: linked (production) vs unlinked build"
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_link.err > gpurun_out/bench_link.json; cut -c1-260 gpurun_out/bench_link.json
F5_LIB=$PWD/f5_tts_b200/libf5tts_b200_nolink.so timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_nolink.err > gpurun_out/bench_nolink.json; cut -c1-260 gpurun_out/bench_nolink.json
echo "=== full bench"
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
echo "=== API tests"
timeout 900 python -m pytest tests/test_gpu_infer.py -q -m gpu -s 2>&1 | grep -E "^\[|passed|failed|FAILED|assert|Error" | cut -c1-220 | tail -10 | tee gpurun_out/test_gpu_infer.log
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_run.log 2>&1
echo "=== ncu full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd" -s 12 -c 1 -o gpurun_out/prof_attn -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_attn_run.log 2>&1
ls gpurun_out | wc -l

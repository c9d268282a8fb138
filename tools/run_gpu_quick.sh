#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -k "attention" 2>&1 | tail -3
timeout 200 python tools/attn_bench.py 2>&1 | tail -4 | tee gpurun_out/attn_bench.log
F5_ATTN_TRACE=1 timeout 200 python tools/attn_trace.py 2>&1 | head -2 | tee gpurun_out/attn_trace.log
timeout 900 python -m pytest tests/test_gpu_sample.py -q -m gpu --tb=short -s 2>&1 | tail -22 | tee gpurun_out/quick_tests_sample.log
for wl in cfg2 cfg4; do
  echo "=== $wl"
  timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload $wl 2>gpurun_out/bench.err | tee gpurun_out/bench_$wl.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],1), 'rtf', round(d['rtf'],5), 'step TF/s', d['roofline']['step_tensor']['achieved'], 'frac', d['roofline']['step_tensor']['frac'], 'launches', d['gpu_launches'])"
  tail -2 gpurun_out/bench.err
done

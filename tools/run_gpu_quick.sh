#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_bench.py 2>&1 | tail -4 | tee gpurun_out/attn_bench.log
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short 2>&1 | tail -4
for wl in cfg2 cfg4; do
  echo "=== $wl"
  timeout 500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --workload $wl 2>gpurun_out/bench.err | tee gpurun_out/bench_$wl.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],1), 'rtf', round(d['rtf'],5), 'step TF/s', d['roofline']['step_tensor']['achieved']); [print('   ', k) for k in d['roofline']['kernels']]"
  tail -2 gpurun_out/bench.err
done

#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
timeout 120 tools/microbench/mma_rate 2>&1 | tee gpurun_out/mma_rate.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2> gpurun_out/bench_q.err > gpurun_out/bench_q.json; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'parity', d['parity']['rel_l2'])"

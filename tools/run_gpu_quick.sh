#!/bin/bash
# scratch: A/B experiments of the current working tree on one GPU
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/quick_tests.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_quick.json | cut -c1-230
for w in cfg3 cfg4 cfg5; do echo "=== bench $w"; timeout 500 python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_$w.json | cut -c1-230; done

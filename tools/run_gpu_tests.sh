#!/bin/bash
# Runs each GPU test module in its own process (a device trap poisons the CUDA context) with a hard timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
for f in "$@"; do
  b=$(basename $f .py)
  echo "=== $f"
  timeout 900 python -m pytest $f -q -m gpu -s --tb=short 2>&1 | tail -${TAILN:-120} | tee gpurun_out/$b.log
done

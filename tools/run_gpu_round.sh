#!/bin/bash
# One GPU visit: parity tests, smoke, bench line (+ reference arm), ncu launch list, ncu full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  echo "=== pytest -m gpu (the driver's command, one process)"
  timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -${TAILN:-12} | tee gpurun_out/test_gpu.log
fi
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 900 python bench.py ${BENCH_ARGS} 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_reference.json | cut -c1-400
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 520 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
tail -1 gpurun_out/ncu_launch_run.log | cut -c1-200
echo "=== ncu full: gemm (one block's four GEMMs)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 5 -o gpurun_out/prof_gemm -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gemm_run.log 2>&1
tail -2 gpurun_out/ncu_gemm_run.log | cut -c1-200
echo "=== ncu full: attention + row_norm"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd|row_norm" -s 12 -c 3 -o gpurun_out/prof_attn -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_attn_run.log 2>&1
tail -2 gpurun_out/ncu_attn_run.log | cut -c1-200
fi
ls -la gpurun_out

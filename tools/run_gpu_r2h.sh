#!/bin/bash
# Round 2, final visit: all GPU tests (the driver's command), smoke, full bench line, reference arm, ncu captures of the
# final kernels (launch list, GEMM / attention / bandwidth kernels --set full).
mkdir -p gpurun_out
echo "=== pytest -m gpu (driver's command)"
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/test_gpu_all.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== full bench (driver's flags)"
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench.err > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err | cut -c1-300
echo "=== reference arm (short)"
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2> gpurun_out/bench_ref.err > gpurun_out/bench_reference.json; cut -c1-300 gpurun_out/bench_reference.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 520 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_run.log 2>&1
echo "=== ncu full: attention, gemm, bandwidth kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd" -s 12 -c 1 -o gpurun_out/prof_attn -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_attn_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 60 -c 5 -o gpurun_out/prof_gemm -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_gemm_run.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:mel_stft|istft_frames|istft_ola|dwconv7_ln|grn_sumsq|grn_apply|cfg_euler|row_norm|ln_affine" -s 30 -c 30 -o gpurun_out/prof_bw -f \
  python tools/ncu_bw.py > gpurun_out/ncu_bw_run.log 2>&1
ls gpurun_out | wc -l

#!/bin/bash
# scratch script for the A/B experiment of the day
mkdir -p gpurun_out
for V in "" _r56 _r48 _r40; do
  echo "lib$V" | tee -a gpurun_out/attn_regs.log
  F5_LIB=$PWD/f5_tts_b200/libf5tts_b200$V.so timeout 300 python tools/attn_bench.py 2>&1 | head -4 | tee -a gpurun_out/attn_regs.log
done
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/norm_large.log
import sys, torch
sys.path.insert(0, ".")
from bench import _graph_time_us
from f5_tts_b200 import ops
g = torch.Generator().manual_seed(0)
for rows in (1876, 15008, 30000):
    xs = [torch.randn(rows, 1024, generator=g).cuda() for _ in range(3)]
    a, b = torch.randn(1024).cuda(), torch.randn(1024).cuda()
    us = _graph_time_us(lambda: [ops.row_norm(xs[i % 3], 0, a, b) for i in range(12)], 12)
    byts = rows * 1024 * 6
    print(f"row_norm rows={rows}: {us:8.2f} us  {byts / us / 1e3:8.1f} GB/s (read fp32 + write fp16)")
PY

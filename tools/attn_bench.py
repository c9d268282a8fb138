"""GPU tool: time the attention kernel at the workload shapes.

A/B the experimental split-KV kernel with  F5_ATTN_VARIANT=6 python tools/attn_bench.py  (default: production kernel);
parity first:  F5_ATTN_VARIANT=6 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention
"""
import sys

import torch

sys.path.insert(0, ".")
from f5_tts_b200 import ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
for Be, seq, H in ((2, 938, 16), (16, 938, 16), (16, 1875, 16), (2, 1875, 16)):
    qkvs = [torch.randn(Be * seq, 3 * H * 64, generator=g).half().to(DEV) for _ in range(3)]
    for q in qkvs:
        ops.attention(q, Be, seq, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n):
        ops.attention(qkvs[i % 3], Be, seq, H)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    fl = 4.0 * Be * H * seq * seq * 64
    print(f"attention Be={Be} seq={seq} H={H}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)

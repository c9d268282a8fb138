"""GPU tool: time the attention kernel at the workload shapes (CUDA-graph timed, like bench.py's roofline table) beside
F.scaled_dot_product_attention on the same strided q/k/v views."""
import sys

import torch
import torch.nn.functional as TF

sys.path.insert(0, ".")
from bench import _graph_time_us  # noqa: E402
from f5_tts_b200 import ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
for Be, seq, H in ((2, 938, 16), (16, 938, 16), (16, 1875, 16), (2, 1875, 16), (2, 469, 16), (16, 939, 16)):
    qkvs = [torch.randn(Be * seq, 3 * H * 64, generator=g).half().to(DEV) for _ in range(3)]
    us = _graph_time_us(lambda: [ops.attention(qkvs[i % 3], Be, seq, H) for i in range(12)], 12)
    q4 = [x.view(Be, seq, 3, H, 64).permute(2, 0, 3, 1, 4) for x in qkvs]
    lib = _graph_time_us(lambda: [TF.scaled_dot_product_attention(q4[i % 3][0], q4[i % 3][1], q4[i % 3][2]) for i in range(12)], 12)
    fl = 4.0 * Be * H * seq * seq * 64
    print(f"attention Be={Be} seq={seq} H={H}: {us:8.2f} us  {fl / us / 1e6:7.1f} TFLOP/s   | SDPA {lib:8.2f} us", flush=True)

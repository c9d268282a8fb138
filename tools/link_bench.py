"""GPU tool: FF1 -> FF2 back to back, linked (block-level hand-off) vs unlinked, CUDA-graph timed over rotating weights."""
import sys

import torch

sys.path.insert(0, ".")
from bench import _graph_time_us  # noqa: E402
from f5_tts_b200 import ops  # noqa: E402
from f5_tts_b200.ops import ACT_GELU_TANH, ACT_NONE, EPI_F16, EPI_RESID  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
for M, D, F in ((1876, 1024, 2048), (1876, 1024, 4096), (15008, 1024, 2048)):
    n_w = 24
    a = torch.randn(M, D, generator=g).half().to(DEV)
    w1 = [(torch.randn(F, D, generator=g) / 32).half().to(DEV) for _ in range(n_w)]
    w2 = [(torch.randn(D, F, generator=g) / 45).half().to(DEV) for _ in range(n_w)]
    b1, b2 = torch.randn(F, generator=g).to(DEV), torch.randn(D, generator=g).to(DEV)
    gate = torch.randn(D, generator=g).to(DEV)
    x = torch.zeros(M, D, device=DEV)
    target = ops.gemm_link_target(M, F, D, EPI_F16, ACT_GELU_TANH)
    ctrs = [torch.zeros((M + 127) // 128, dtype=torch.int32, device=DEV) for _ in range(n_w)]

    def plain():
        for i in range(n_w):
            h = ops.linear(a, w1[i], b1, epi=EPI_F16, act=ACT_GELU_TANH, static_w=True)
            ops.linear(h, w2[i], b2, epi=EPI_RESID, resid=x, gate=gate, static_w=True)

    def linked():
        for i in range(n_w):
            ctrs[i].zero_()
        for i in range(n_w):
            h = ops.linear(a, w1[i], b1, epi=EPI_F16, act=ACT_GELU_TANH, static_w=True, done_counters=ctrs[i])
            ops.linear(h, w2[i], b2, epi=EPI_RESID, resid=x, gate=gate, static_w=True, ready=(ctrs[i], target))

    up = _graph_time_us(plain, n_w)
    ul = _graph_time_us(linked, n_w)
    print(f"M={M} F={F}: FF1+FF2 unlinked {up:7.2f} us | linked {ul:7.2f} us per pair", flush=True)

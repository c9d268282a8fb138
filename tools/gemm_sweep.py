"""GPU tool: time the persistent tcgen05 GEMM per (shape, epilogue, tile) from a CUDA graph (rotating weights > L2)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from bench import _graph_time_us  # noqa: E402
from f5_tts_b200 import ops  # noqa: E402
from f5_tts_b200.ops import *  # noqa: E402,F403

DEV = "cuda:0"
torch.cuda.init()
g = torch.Generator().manual_seed(0)


def bench(M, N, K, epi, act, bn, nw=24, pair=0):
    a = [torch.randn(M, K, generator=g).half().to(DEV) for _ in range(2)]
    w = [(torch.randn(N, K, generator=g) / 32).half().to(DEV) for _ in range(nw)]
    b = torch.randn(N, generator=g).to(DEV)
    kw = dict(epi=epi, act=act, bn=bn, pair=pair, static_w=True)
    if epi == EPI_RESID:
        kw["resid"] = torch.zeros(M, N, device=DEV)
        kw["gate"] = torch.randn(N, generator=g).to(DEV)
    if epi == EPI_QKV_ROPE:
        seq = M // 2
        kw.update(seq=seq, rope=ops.rope_tables(seq, DEV), inner=N // 3, pe_heads=1)
    us = _graph_time_us(lambda: [ops.linear(a[i % 2], w[i], b, **kw) for i in range(nw)], nw, rounds=4)
    return us, 2.0 * M * N * K / (us * 1e-6) / 1e12


Ms = [int(x) for x in os.environ.get("SWEEP_M", "1876,3752,7504,15008").split(",")]
for M in Ms:
    for (N, K, epi, act, tag) in ((3072, 1024, EPI_QKV_ROPE, ACT_NONE, "QKV"), (1024, 1024, EPI_RESID, ACT_NONE, "out"),
                                  (2048, 1024, EPI_F16, ACT_GELU_TANH, "FF1"), (1024, 2048, EPI_RESID, ACT_NONE, "FF2"),
                                  (4096, 1024, EPI_F16, ACT_GELU_TANH, "FF1-e2"), (1024, 4096, EPI_RESID, ACT_NONE, "FF2-e2")):
        if M > 2000 and "e2" in tag:
            continue
        nw = -(-200_000_000 // (N * K * 2))  # distinct weights > L2 (126 MB), as in the real step
        nw = nw if M < 8000 else max(6, nw // 4)
        row = []
        for bn, pair in ((128, 0), (192, 0), (256, 0), (128, 1), (192, 1), (256, 1)):
            us, tf = bench(M, N, K, epi, act, bn, nw=nw, pair=pair)
            row.append(f"{'P' if pair else 'bn'}{bn}: {us:6.1f}us {tf:5.0f}TF")
        print(f"M={M:6d} {tag:7s} N={N:5d} K={K:5d} | " + " | ".join(row) + f" | auto={ops.gemm_tile(M, N, K, epi, act)}",
              flush=True)

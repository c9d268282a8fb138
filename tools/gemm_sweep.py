"""GPU tool: time the persistent tcgen05 GEMM per (shape, epilogue, BN) with CUDA events (rotating weights > L2)."""
import sys

import torch

sys.path.insert(0, ".")
from f5_tts_b200 import ops  # noqa: E402
from f5_tts_b200.ops import *  # noqa: E402,F403

DEV = "cuda:0"
torch.cuda.init()
g = torch.Generator().manual_seed(0)


def bench(M, N, K, epi, act, bn, nw=24, rounds=4, pair=0):
    a = [torch.randn(M, K, generator=g).half().to(DEV) for _ in range(2)]
    w = [(torch.randn(N, K, generator=g) / 32).half().to(DEV) for _ in range(nw)]
    b = torch.randn(N, generator=g).to(DEV)
    kw = dict(epi=epi, act=act, bn=bn, pair=pair)
    if epi == EPI_RESID:
        kw["resid"] = torch.zeros(M, N, device=DEV)
        kw["gate"] = torch.randn(N, generator=g).to(DEV)
    if epi == EPI_QKV_ROPE:
        seq = M // 2
        kw.update(seq=seq, rope=ops.rope_tables(seq, DEV), inner=N // 3, pe_heads=1)
    for i in range(nw):
        ops.linear(a[i % 2], w[i], b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        for i in range(nw):
            ops.linear(a[i % 2], w[i], b, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (rounds * nw)
    tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
    return us, tf


names = {EPI_F16: "f16", EPI_F32: "f32", EPI_RESID: "resid", EPI_QKV_ROPE: "qkv"}
import os
for M in ((1876, 30000) if os.environ.get('F5_GEMM_DBG') else (1876, 15008, 30000)):
    for (N, K, epi, act, tag) in ((3072, 1024, EPI_QKV_ROPE, ACT_NONE, "QKV"), (1024, 1024, EPI_RESID, ACT_NONE, "out"),
                                  (2048, 1024, EPI_F16, ACT_GELU_TANH, "FF1"), (1024, 2048, EPI_RESID, ACT_NONE, "FF2"),
                                  (4096, 1024, EPI_F16, ACT_GELU_TANH, "FF1-e2"), (1024, 4096, EPI_RESID, ACT_NONE, "FF2-e2")):
        if M > 2000 and "e2" in tag:
            continue
        row = []
        for bn in (64, 128, 256):
            if epi == EPI_QKV_ROPE and bn == 64:
                row.append("   --   ")
                continue
            nw = 24 if M < 2000 else 6
            us, tf = bench(M, N, K, epi, act, bn, nw=nw, rounds=3)
            row.append(f"bn{bn}: {us:7.1f}us {tf:6.0f}TF")
        for bn in (128, 256):
            us, tf = bench(M, N, K, epi, act, bn, nw=24 if M < 2000 else 6, rounds=3, pair=1)
            row.append(f"PAIR{bn}: {us:7.1f}us {tf:6.0f}TF")
        print(f"M={M:6d} {tag:7s} N={N:5d} K={K:5d} | " + " | ".join(row), flush=True)

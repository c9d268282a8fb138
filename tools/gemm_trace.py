"""GPU diagnostic: per-CTA phase timestamps of the persistent GEMM (run with F5_GEMM_TRACE=1)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from f5_tts_b200 import _lib, ops  # noqa: E402
from f5_tts_b200.ops import *  # noqa: E402,F403

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
L = _lib.lib()
L.f5_debug_gemm_trace.argtypes = [C.c_void_p, C.c_int]
for (M, N, K, epi, act, tag) in ((1876, 3072, 1024, EPI_QKV_ROPE, ACT_NONE, "QKV"), (1876, 1024, 1024, EPI_RESID, ACT_NONE, "out"),
                                 (1876, 2048, 1024, EPI_F16, ACT_GELU_TANH, "FF1"), (1876, 1024, 2048, EPI_RESID, ACT_NONE, "FF2")):
    a = torch.randn(M, K, generator=g).half().to(DEV)
    ws = [(torch.randn(N, K, generator=g) / 32).half().to(DEV) for _ in range(40)]  # > L2: weights come from HBM
    b = torch.randn(N, generator=g).to(DEV)
    kw = dict(epi=epi, act=act, bn=0, static_w=True)  # tile shape left to the planner, as in the engine
    if epi == EPI_RESID:
        kw["resid"] = torch.zeros(M, N, device=DEV)
        kw["gate"] = torch.randn(N, generator=g).to(DEV)
    if epi == EPI_QKV_ROPE:
        kw.update(seq=M // 2, rope=ops.rope_tables(M // 2, DEV), inner=N // 3, pe_heads=1)
    for i in range(40):
        ops.linear(a, ws[i], b, **kw)
    torch.cuda.synchronize()
    n = 148
    buf = np.zeros((n, 16), dtype=np.int64)
    rc = L.f5_debug_gemm_trace(buf.ctypes.data, n)
    t = buf.astype(np.float64)
    t = t[t[:, 1] > 0]
    clk = 1.9  # cycles per ns (approx.)
    rel = lambda k: (t[:, k] - t[:, 1]) / clk / 1000.0  # us since CTA start
    g0 = t[:, 0].min()
    print(f"{tag} tile {ops.gemm_tile(M, N, K, epi, act)} ({len(t)} CTAs): start spread {(t[:,0].max()-g0)/1000:.2f} us | setup {np.median(rel(2)):.2f} | "
          f"first operands {np.median(rel(3)):.2f} | MMAs issued {np.median(rel(4)):.2f} (max {rel(4).max():.2f}) | first acc {np.median(rel(5)):.2f} | "
          f"last tile loop done {np.median(rel(12)):.2f} | epilogue done {np.median(rel(6)):.2f} (max {rel(6).max():.2f}) | exit {np.median(rel(7)):.2f} (max {rel(7).max():.2f}) us"
          f" | epilogue warps waited for accumulators {np.median(t[:,11])/clk/1000:.2f} us", flush=True)

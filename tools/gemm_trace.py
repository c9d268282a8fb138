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
for (M, N, K, epi, act, bn, tag) in ((1876, 3072, 1024, EPI_QKV_ROPE, ACT_NONE, 128, "QKV"), (1876, 1024, 1024, EPI_RESID, ACT_NONE, 128, "out"),
                                     (1876, 2048, 1024, EPI_F16, ACT_GELU_TANH, 128, "FF1"), (1876, 1024, 2048, EPI_RESID, ACT_NONE, 128, "FF2"),
                                     (1876, 1024, 1024, EPI_RESID, ACT_NONE, 128, "out+norm"), (1876, 1024, 2048, EPI_RESID, ACT_NONE, 128, "FF2+norm")):
    a = torch.randn(M, K, generator=g).half().to(DEV)
    w = (torch.randn(N, K, generator=g) / 32).half().to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    kw = dict(epi=epi, act=act, bn=bn)
    if epi == EPI_RESID:
        kw["resid"] = torch.zeros(M, N, device=DEV)
    if "norm" in tag:
        kw["norm"] = (0, torch.randn(N, generator=g).to(DEV), torch.randn(N, generator=g).to(DEV),
                      torch.zeros(256, dtype=torch.int32, device=DEV))
    if epi == EPI_QKV_ROPE:
        kw.update(seq=M // 2, rope=ops.rope_tables(M // 2, DEV), inner=N // 3, pe_heads=1)
    for _ in range(3):
        ops.linear(a, w, b, **kw)
    torch.cuda.synchronize()
    ops.linear(a, w, b, **kw)
    n = 148
    buf = np.zeros((n, 16), dtype=np.int64)
    rc = L.f5_debug_gemm_trace(buf.ctypes.data, n)
    t = buf.astype(np.float64)
    g0 = t[:, 0].min()
    clk = 1.9  # cycles per ns (approx.)
    rel = lambda k: (t[:, k] - t[:, 1]) / clk / 1000.0  # us since CTA start
    print(f"{tag}: CTA start spread {(t[:,0].max()-g0)/1000:.2f} us | setup {np.median(rel(2)):.2f} | first operands {np.median(rel(3)):.2f} | "
          f"MMAs issued {np.median(rel(4)):.2f} (max {rel(4).max():.2f}) | first acc {np.median(rel(5)):.2f} | epilogue done {np.median(rel(6)):.2f} (max {rel(6).max():.2f}) | exit {np.median(rel(7)):.2f} (max {rel(7).max():.2f}) us",
          flush=True)
    print(f"     wait-acc {np.median(t[:,11])/clk/1000:.2f} us | final drain {np.median(t[:,6]-t[:,12])/clk/1000:.2f} us", flush=True)
    if "norm" in tag:
        act = t[:, 8] > 0
        print(f"     fused norm tail (us since CTA start, median/max over {int(act.sum())} CTAs): tile loop done {np.median(rel(12)[act]):.2f}/{rel(12)[act].max():.2f} | "
              f"stores performed+fenced {np.median(rel(8)[act]):.2f}/{rel(8)[act].max():.2f} | arrivals posted {np.median(rel(9)[act]):.2f}/{rel(9)[act].max():.2f} | "
              f"block complete {np.median(rel(10)[act]):.2f}/{rel(10)[act].max():.2f} | tail done {np.median(rel(6)[act]):.2f}/{rel(6)[act].max():.2f}", flush=True)

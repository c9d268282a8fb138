"""GPU diagnostic: where the softmax warpgroups of the attention kernel spend their cycles (F5_ATTN_TRACE=1)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from f5_tts_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
L = _lib.lib()
L.f5_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
for Be, seq, H in ((2, 938, 16), (16, 938, 16)):
    qkv = torch.randn(Be * seq, 3 * H * 64, generator=g).half().to(DEV)
    for _ in range(3):
        ops.attention(qkv, Be, seq, H)
    torch.cuda.synchronize()
    n = ((seq + 255) // 256) * H * Be
    buf = np.zeros((n, 16), dtype=np.int64)
    L.f5_debug_attn_trace(buf.ctypes.data, n)
    t = buf.astype(np.float64) / 1.9 / 1000.0
    for w in (0, 1):
        o = w * 8
        print(f"Be={Be} seq={seq} WG{w} (median over {n} CTAs, us): wait S {np.median(t[:,o+0]):.2f} | turnstile {np.median(t[:,o+1]):.2f} | "
              f"tmem ld {np.median(t[:,o+7]):.2f} | max {np.median(t[:,o+6]):.2f} | exp {np.median(t[:,o+2]):.2f} | wait O {np.median(t[:,o+3]):.2f} | write P+rescale+arrive {np.median(t[:,o+4]):.2f} | "
              f"total loop {np.median(t[:,o+5]):.2f} | kv tiles {(seq + 127) // 128}", flush=True)

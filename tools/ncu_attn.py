"""Tiny driver for ncu: a few attention launches at the cfg2 shape (Be=2, seq=938, H=16)."""
import sys
import torch
sys.path.insert(0, ".")
from f5_tts_b200 import ops
g = torch.Generator().manual_seed(0)
qkv = torch.randn(2 * 938, 3072, generator=g).half().to("cuda:0")
for _ in range(4):
    ops.attention(qkv, 2, 938, 16)
torch.cuda.synchronize()

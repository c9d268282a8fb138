"""GPU diagnostic: repeat each kernel at the engine's shapes and report run-to-run bit differences."""
import math
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from f5_tts_b200 import ops  # noqa: E402
from f5_tts_b200.ops import *  # noqa: E402,F403

DEV = "cuda:0"
torch.cuda.init()


def gen(shape, seed, scale=1.0, dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


def repeat(name, fn, n=8):
    outs = []
    t0 = time.time()
    for _ in range(n):
        outs.append(fn().clone())
    torch.cuda.synchronize()
    bad = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
    worst = max(float((outs[0].float() - o.float()).abs().max()) for o in outs[1:])
    nd = max(int((outs[0] != o).sum()) for o in outs[1:])
    print(f"[det] {name}: {bad}/{n - 1} runs differ, max|d| {worst:.3e}, max #elements differing {nd}  ({time.time() - t0:.2f}s)",
          flush=True)
    return outs[0]


M, D, H, seq, Be = 1876, 1024, 16, 938, 2
a = gen((M, D), 1)
for bn in (64, 128):
    w = gen((D, D), 2, 1 / 32)
    b = gen((D,), 3, 1, torch.float32)
    repeat(f"gemm f32 bn{bn}", lambda: ops.linear(a, w, b, epi=EPI_F32, bn=bn))
    x0 = gen((M, D), 4, 1, torch.float32)
    gate = gen((D,), 5, 1, torch.float32)

    def resid():
        x = x0.clone()
        ops.linear(a, w, b, epi=EPI_RESID, bn=bn, resid=x, gate=gate)
        return x
    repeat(f"gemm resid bn{bn}", resid)
w2 = gen((2048, D), 6, 1 / 32)
repeat("gemm f16 gelu bn128", lambda: ops.linear(a, w2, None, epi=EPI_F16, act=ACT_GELU_TANH, bn=128))
wq = gen((3072, D), 7, 1 / 32)
cs, sn = ops.rope_tables(seq, DEV)
bq = gen((3072,), 8, 0.1, torch.float32)
qkv = repeat("gemm qkv rope bn128", lambda: ops.linear(a, wq, bq, epi=EPI_QKV_ROPE, bn=128, seq=seq, rope=(cs, sn), inner=1024, pe_heads=1))
att = repeat("attention 2x938x16", lambda: ops.attention(qkv, Be, seq, H), n=12)
q, k, v = qkv.float().view(Be, seq, 3, H, 64).permute(2, 0, 3, 1, 4)
ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Be * seq, H * 64)
print("attention rel-L2 vs torch", float((att.float() - ref).norm() / ref.norm()), flush=True)
qkv_r = gen((Be * seq, 3072), 9, 1.0)
att2 = repeat("attention random qkv", lambda: ops.attention(qkv_r, Be, seq, H), n=12)
q, k, v = qkv_r.float().view(Be, seq, 3, H, 64).permute(2, 0, 3, 1, 4)
ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Be * seq, H * 64)
print("attention(random) rel-L2 vs torch", float((att2.float() - ref).norm() / ref.norm()), flush=True)
kvl = torch.tensor([938, 500], dtype=torch.int32, device=DEV)
repeat("attention kv_len", lambda: ops.attention(qkv_r, Be, seq, H, kvl), n=8)
# grouped conv
x = gen((Be, seq, D), 10)
wc = gen((D, 64, 31), 11, 1 / math.sqrt(64 * 31))
bc = gen((D,), 12, 0.1, torch.float32)
wp = wc.permute(2, 0, 1).contiguous()
print("conv: launching", flush=True)
c1 = repeat("conv31 f16", lambda: ops.grouped_conv31(x, wp, bc))
refc = F.mish(F.conv1d(x.float().cpu().transpose(1, 2), wc.float().cpu(), bc.cpu(), padding=15, groups=16).transpose(1, 2))
print("conv rel-L2 vs torch(cpu)", float((c1.float().cpu() - refc).norm() / refc.norm()), flush=True)
r0 = gen((Be, seq, D), 13, 1, torch.float32)


def convres():
    r = r0.clone()
    ops.grouped_conv31(x, wp, bc, resid=r)
    return r
repeat("conv31 resid", convres)
xx = gen((M, D), 14, 2, torch.float32)
aa, bb = gen((D,), 15, 0.3, torch.float32), gen((D,), 16, 0.3, torch.float32)
repeat("row_norm", lambda: ops.row_norm(xx, 0, aa, bb))
print("done", flush=True)

"""Re-pack a loaded ``state_dict`` (released checkpoint layout) into the kernels' operand layouts.

 * GEMM weights -> fp16, row-major [N, K] (K-major operands for tcgen05); to_q|to_k|to_v stacked into one
   [3*inner, D] matrix so QKV is a single GEMM; all AdaLN linears of a DiT stacked into one
   [depth*6D + 2D, D] matrix (rows per block in the reference's chunk order shift_msa, scale_msa, gate_msa,
   shift_mlp, scale_mlp, gate_mlp — model/modules.py:323; final: scale, shift — modules.py:344);
 * input_embed.proj K padded with zero columns to a multiple of 64;
 * grouped Conv1d(k=31, g=16) weight [D, 64, 31] -> [31][D][64] so that tap t / group g is a K-major 64x64 tile;
 * biases, norm gains, GRN parameters, embedding table -> fp32.
The packed tensors are owned by Python (kept alive in the returned dict); the C engine only stores pointers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _h(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float16).contiguous()


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def pack_backbone(m) -> dict:
    sd = {k: v for k, v in m.state_dict().items()}
    dev = next(m.parameters()).device
    if dev.type != "cuda":
        raise _lib.F5LibraryError("the backbone must live on a CUDA (B200) device before sampling; no CPU path exists")
    keep: list[torch.Tensor] = []

    def H(t):
        t = _h(t)
        keep.append(t)
        return t.data_ptr()

    def Fp(t):
        t = _f(t)
        keep.append(t)
        return t.data_ptr()

    D, depth, inner = m.dim, m.depth, m.heads * m.dim_head
    dit = m.KIND == "DiT"
    W = _lib.Weights()
    W.time_w0, W.time_b0 = H(sd["time_embed.time_mlp.0.weight"]), Fp(sd["time_embed.time_mlp.0.bias"])
    W.time_w1, W.time_b1 = H(sd["time_embed.time_mlp.2.weight"]), Fp(sd["time_embed.time_mlp.2.bias"])
    W.text_table = Fp(sd["text_embed.text_embed.weight"])
    for i in range(m.conv_layers):
        p = f"text_embed.text_blocks.{i}."
        tb = W.text_blocks[i]
        tb.dw_w, tb.dw_b = Fp(sd[p + "dwconv.weight"].reshape(m.text_dim, 7)), Fp(sd[p + "dwconv.bias"])
        tb.ln_w, tb.ln_b = Fp(sd[p + "norm.weight"]), Fp(sd[p + "norm.bias"])
        tb.pw1_w, tb.pw1_b = H(sd[p + "pwconv1.weight"]), Fp(sd[p + "pwconv1.bias"])
        tb.grn_gamma, tb.grn_beta = Fp(sd[p + "grn.gamma"].reshape(-1)), Fp(sd[p + "grn.beta"].reshape(-1))
        tb.pw2_w, tb.pw2_b = H(sd[p + "pwconv2.weight"]), Fp(sd[p + "pwconv2.bias"])
    pw = sd["input_embed.proj.weight"]
    kin = pw.shape[1]
    kpad = (kin + 63) // 64 * 64
    pwp = torch.zeros((D, kpad), dtype=torch.float16, device=dev)
    pwp[:, :kin] = pw.to(torch.float16)
    keep.append(pwp)
    W.proj_w, W.proj_b, W.proj_kpad = pwp.data_ptr(), Fp(sd["input_embed.proj.bias"]), kpad
    for j, idx in enumerate((0, 2)):
        cw = sd[f"input_embed.conv_pos_embed.conv1d.{idx}.weight"]  # [D_out, 64, 31]
        W.conv_w[j] = H(cw.permute(2, 0, 1))                          # [31, D_out, 64]
        W.conv_b[j] = Fp(sd[f"input_embed.conv_pos_embed.conv1d.{idx}.bias"])
    layers = (_lib.LayerWeights * depth)()
    for i in range(depth):
        lw = layers[i]
        if dit:
            p, a, f = f"transformer_blocks.{i}.", f"transformer_blocks.{i}.attn.", f"transformer_blocks.{i}.ff."
        else:
            p, a, f = f"layers.{i}.", f"layers.{i}.2.", f"layers.{i}.4."
        lw.w_qkv = H(torch.cat([sd[a + "to_q.weight"], sd[a + "to_k.weight"], sd[a + "to_v.weight"]], dim=0))
        lw.b_qkv = Fp(torch.cat([sd[a + "to_q.bias"], sd[a + "to_k.bias"], sd[a + "to_v.bias"]], dim=0))
        lw.w_out, lw.b_out = H(sd[a + "to_out.0.weight"]), Fp(sd[a + "to_out.0.bias"])
        lw.w_ff1, lw.b_ff1 = H(sd[f + "ff.0.0.weight"]), Fp(sd[f + "ff.0.0.bias"])
        lw.w_ff2, lw.b_ff2 = H(sd[f + "ff.2.weight"]), Fp(sd[f + "ff.2.bias"])
        if not dit:
            lw.w_skip = H(sd[p + "0.weight"]) if (p + "0.weight") in sd else None
            lw.g_attn, lw.g_ff = Fp(sd[p + "1.g"]), Fp(sd[p + "3.g"])
    W.layers = C.cast(layers, C.POINTER(_lib.LayerWeights))
    if dit:
        mw = [sd[f"transformer_blocks.{i}.attn_norm.linear.weight"] for i in range(depth)] + [sd["norm_out.linear.weight"]]
        mb = [sd[f"transformer_blocks.{i}.attn_norm.linear.bias"] for i in range(depth)] + [sd["norm_out.linear.bias"]]
        W.mod_w, W.mod_b = H(torch.cat(mw, dim=0)), Fp(torch.cat(mb, dim=0))
    else:
        W.g_out = Fp(sd["norm_out.g"])
    W.out_w, W.out_b = H(sd["proj_out.weight"]), Fp(sd["proj_out.bias"])

    A = _lib.Arch()
    A.backbone = 0 if dit else 1
    A.dim, A.depth, A.heads, A.dim_head, A.ff_inner = D, depth, m.heads, m.dim_head, m.ff_inner
    A.mel_dim, A.text_dim, A.text_num_embeds, A.conv_layers = m.mel_dim, m.text_dim, m.text_num_embeds, m.conv_layers
    A.text_mask_padding = 1 if m.text_mask_padding else 0
    A.pe_attn_head = -1 if m.pe_attn_head is None else int(m.pe_attn_head)
    A.attn_mask_enabled = 1 if m.attn_mask_enabled else 0
    handle = C.c_void_p()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().f5_engine_create(C.byref(A), C.byref(W), C.byref(handle)), "f5_engine_create")
    return {"handle": handle, "keep": keep, "layers": layers, "weights": W, "arch": A}

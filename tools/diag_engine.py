"""GPU diagnostic: run-to-run reproducibility of the whole engine (1 step) for DiT and UNetT."""
import sys

import torch

sys.path.insert(0, ".")
import f5_tts_b200 as F5  # noqa: E402
from oracle import f5_oracle as O  # noqa: E402

DEV = "cuda:0"
for cfg in (O.f5tts_base(), O.e2tts_base()):
    cls = F5.DiT if cfg.backbone == "DiT" else F5.UNetT
    m = cls(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_dim=cfg.text_dim,
            text_mask_padding=cfg.text_mask_padding, conv_layers=cfg.conv_layers, pe_attn_head=cfg.pe_attn_head,
            text_num_embeds=2545, mel_dim=100)
    model = F5.CFM(transformer=m)
    model.load_state_dict(O.synthetic_state_dict(cfg), strict=True)
    model = model.to(DEV)
    g = torch.Generator().manual_seed(1)
    for N in (200, 938):
        x = torch.randn(1, N, 100, generator=g).to(DEV)
        cond = torch.randn(1, N, 100, generator=g).to(DEV)
        text = torch.randint(0, 2545, (1, 40), generator=g).to(DEV)
        outs = [model.transformer(x, cond, text, torch.tensor(0.3), cfg_infer=True).clone() for _ in range(6)]
        d = [float((outs[0] - o).abs().max()) for o in outs[1:]]
        nd = [int((outs[0] != o).sum()) for o in outs[1:]]
        print(f"[engine {cfg.backbone} N={N}] max|d| per repeat {d}  #diff {nd}  |v| max {float(outs[0].abs().max()):.3f}", flush=True)
    del model, m
    torch.cuda.empty_cache()

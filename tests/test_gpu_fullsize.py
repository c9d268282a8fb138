"""GPU parity of CFM.sample at the REAL sizes of the BASELINE.json configurations.

  cfg2  B=1, 938 frames, NFE 32 in full — against tests/golden/cfg2_full_nfe32.npz, produced by the UNMODIFIED reference
        in fp32 (oracle/make_golden_baseline.py); drift reported at steps 1 / 8 / 16 / 32 next to the reference's own
        fp16-vs-fp32 drift on the same inputs (stored in the fixture; SURVEY.md §9.4 measured 1.5e-3 at N = 400).
  cfg3  B=8 variable length 469..1875 frames (padded), NFE 8 of 32, faithful mode (padded keys attended,
        attn_mask_enabled=False) and masked mode (attn_mask_enabled=True)
  cfg4  B=8 x 938 frames (one GPU's shard of the 64-utterance config), the full EPSS-16 grid
  cfg5  E2-TTS UNetT, B=8 x 938 frames, NFE 8 of 32 — the batched UNetT path (time token, row_len + 1)
cfg3-5 compare with the CPU oracle computed live on the box's host cores (the oracle is pinned bit-exactly to the
reference by tests/test_oracle_vs_golden.py; their outputs are too large to commit).  Identical injected y0 everywhere.

Tolerance: rel-L2 of the generated region <= 5e-3 at every checked step.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.fullsize]

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import f5_tts_b200 as F5  # noqa: E402
import synthdata as SD  # noqa: E402
from oracle import f5_oracle as O  # noqa: E402

DEV = "cuda:0"
TOL = 5e-3
_cache = {}


def build(cfg, wseed=1234):
    key = (repr(cfg), wseed)
    if key not in _cache:
        _cache.clear()  # one 1.3 GB model at a time
        cls = F5.DiT if cfg.backbone == "DiT" else F5.UNetT
        model = F5.CFM(transformer=cls(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head,
                                       ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim, text_num_embeds=cfg.text_num_embeds,
                                       text_dim=cfg.text_dim, text_mask_padding=cfg.text_mask_padding,
                                       conv_layers=cfg.conv_layers, pe_attn_head=cfg.pe_attn_head,
                                       attn_mask_enabled=cfg.attn_mask_enabled))
        sd = SD.synthetic_state_dict(cfg, seed=wseed)
        model.load_state_dict(sd, strict=True)
        _cache[key] = (model.to(DEV), sd)
    return _cache[key]


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def host_threads():
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(round(int(quota) / int(period)))))
    except Exception:  # noqa: BLE001
        pass
    return n


def test_cfg2_full_nfe32_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "cfg2_full_nfe32.npz"))
    w = SD.WORKLOADS["cfg2"]
    cfg = SD.f5tts_base()
    model, _ = build(cfg, int(z["wseed"]))
    wav, text, duration, _ = SD.synth_inputs(w)
    out, traj = model.sample(wav.to(DEV), text.to(DEV), int(duration[0]), steps=int(z["steps"]),
                             cfg_strength=float(z["cfg_strength"]), sway_sampling_coef=float(z["sway"]), seed=int(z["seed"]),
                             y0=torch.from_numpy(z["y0"]).to(DEV))
    n_ref = int(z["n_ref"])
    gen = slice(n_ref, None)
    drift = {int(k): rel(traj[int(k)][:, gen], torch.from_numpy(z[f"traj_{int(k)}"])[:, gen]) for k in z["kept"]}
    ref16 = dict(zip([int(k) for k in z["kept"]], z["ref_fp16_drift"].tolist()))
    print("[cfg2 full NFE 32] rel-L2 of the generated region vs the fp32 reference, per step:")
    for k in sorted(drift):
        print(f"   step {k:2d}: B200 path {drift[k]:.3e}   reference's own fp16 path {ref16[k]:.3e}")
    final = rel(out, torch.from_numpy(z["out"]))
    print(f"   final mel (all rows) {final:.3e}   gate {TOL:.0e}")
    # prompt rows are the prompt mel itself (cfm.py:221-223): CUDA mel kernel vs torchaudio on the CPU, fp32 rounding only
    assert torch.allclose(out[:, :n_ref].float().cpu(), torch.from_numpy(z["out"])[:, :n_ref], atol=2e-4), "prompt rows"
    assert all(v <= TOL for v in drift.values()) and final <= TOL
    # fp32 state / residual / statistics: the B200 path must not drift more than the reference's fp16 path does
    assert drift[32] <= ref16[32]


def _oracle_vs_gpu(name, cfg, w, steps, wseed=1234):
    torch.set_num_threads(host_threads())
    model, sd = build(cfg, wseed)
    wav, text, duration, lens = SD.synth_inputs(w)
    cond = O.mel_spectrogram(wav).permute(0, 2, 1).contiguous()  # [B, n_ref, 100]; prompt lengths via `lens`
    kw = dict(lens=lens, steps=steps, cfg_strength=SD.CFG_STRENGTH, sway_sampling_coef=SD.SWAY, seed=0)
    ref = O.sample(sd, cfg, cond, text, duration, **kw)
    dkw = dict(kw, lens=lens.to(DEV))
    out, traj = model.sample(cond.to(DEV), text.to(DEV), duration.to(DEV), **dkw, y0=ref.y0.to(DEV))
    worst = 0.0
    for b in range(w["B"]):  # valid generated rows of each utterance
        sl = slice(int(lens[b]), int(duration[b]))
        r1 = rel(traj[1][b, sl], ref.trajectory[1][b, sl])
        rN = rel(out[b, sl], ref.out[b, sl])
        worst = max(worst, r1, rN)
        print(f"[{name}] utt {b} frames {int(duration[b])}: step-1 {r1:.3e}  final({steps} steps) {rN:.3e}")
    assert worst <= TOL
    return worst


@pytest.mark.parametrize("attn_mask", [False, True])
def test_cfg3_varlen_b8_nfe8(attn_mask):
    cfg = SD.f5tts_base()
    cfg.attn_mask_enabled = attn_mask
    _oracle_vs_gpu(f"cfg3 {'masked' if attn_mask else 'faithful'}", cfg, SD.WORKLOADS["cfg3"], steps=8)


def test_cfg4_b8_epss16():
    _oracle_vs_gpu("cfg4 (one GPU's shard, EPSS-16)", SD.f5tts_base(), SD.WORKLOADS["cfg4"], steps=16)


def test_cfg5_unett_b8_nfe8():
    _oracle_vs_gpu("cfg5 UNetT", SD.e2tts_base(), SD.WORKLOADS["cfg5"], steps=8, wseed=1234)

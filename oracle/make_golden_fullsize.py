"""Golden vectors for the full-size GPU parity tests of BASELINE.json configs 3, 4 and 5 (tests/test_gpu_fullsize.py).

TEST INFRASTRUCTURE.  Runs the CPU oracle (oracle/f5_oracle.py — pinned bit-exactly against the unmodified reference by
tests/test_oracle_vs_golden.py and oracle/make_golden.py) on exactly the inputs the tests use and stores every `STRIDE`-th
generated row of each utterance after solver step 1 and after the last step, in fp32.  The live oracle computation took
~6 of the GPU suite's 12 minutes on the GPU box's host and scales with that host's cores; the driver allows the suite
20 minutes, so the expensive side is computed once, here.

    python oracle/make_golden_fullsize.py            # ~15 min on 8 cores; writes tests/golden/fullsize_*.npz

The initial noise is not stored: cfm.py:196-201 draws it per utterance from `torch.manual_seed(seed)` on the CPU
generator, which the test repeats (`draw_y0`); a checksum of it is stored and verified.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthdata as SD  # noqa: E402
from oracle import f5_oracle as O  # noqa: E402

STRIDE = 3
CASES = {
    # name: (arch factory, attn_mask_enabled, workload, solver steps)
    "cfg3_faithful": ("f5tts_base", False, "cfg3", 8),
    "cfg3_masked": ("f5tts_base", True, "cfg3", 8),
    "cfg4_epss16": ("f5tts_base", False, "cfg4", 16),
    "cfg5_unett": ("e2tts_base", False, "cfg5", 8),
}


def draw_y0(duration, mel_dim=100, seed=0):
    """cfm.py:196-201 on the CPU generator: one `manual_seed` + `randn(dur, mel)` per utterance, zero-padded."""
    rows = []
    for dur in duration.tolist():
        torch.manual_seed(seed)
        rows.append(torch.randn(int(dur), mel_dim, dtype=torch.float32))
    return torch.nn.utils.rnn.pad_sequence(rows, padding_value=0, batch_first=True)


def kept_rows(t, lens, duration):
    """every STRIDE-th valid generated row of each utterance, concatenated: [sum_b ceil((dur_b - len_b) / STRIDE), mel]"""
    return torch.cat([t[b, int(lens[b]): int(duration[b]): STRIDE] for b in range(t.shape[0])], 0)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (arch, attn_mask, wl, steps) in CASES.items():
        cfg = getattr(SD, arch)()
        cfg.attn_mask_enabled = attn_mask
        w = SD.WORKLOADS[wl]
        sd = SD.synthetic_state_dict(cfg, seed=1234)
        wav, text, duration, lens = SD.synth_inputs(w)
        cond = O.mel_spectrogram(wav).permute(0, 2, 1).contiguous()
        y0 = draw_y0(duration, cfg.mel_dim, seed=0)
        t0 = time.time()
        ref = O.sample(sd, cfg, cond, text, duration, lens=lens, steps=steps, cfg_strength=SD.CFG_STRENGTH,
                       sway_sampling_coef=SD.SWAY, seed=0, y0=y0)
        chk = O.sample  # noqa: F841  (keep the symbol referenced for readers following the call above)
        np.savez(os.path.join(out_dir, f"fullsize_{name}.npz"), stride=STRIDE, steps=steps, wseed=1234, seed=0,
                 arch=arch, attn_mask_enabled=attn_mask, workload=wl,
                 y0_checksum=np.array([float(y0.double().sum()), float(y0.double().abs().sum())]),
                 step1=kept_rows(ref.trajectory[1], lens, duration).numpy(),
                 final=kept_rows(ref.out, lens, duration).numpy())
        print(f"{name}: {steps} steps in {time.time() - t0:.0f} s, {os.path.getsize(os.path.join(out_dir, f'fullsize_{name}.npz')) / 1e6:.1f} MB",
              flush=True)


if __name__ == "__main__":
    main()

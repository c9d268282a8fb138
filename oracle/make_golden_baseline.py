"""Golden vectors of the BASELINE.json configurations at their REAL size and NFE (TEST INFRASTRUCTURE).

Run in the build container, where /root/reference exists:

    python -m oracle.make_golden_baseline [cfg2] [cfg2_fp16]

cfg2 = BASELINE.json configs[1] = the bench.py default workload, bit for bit the same synthetic inputs
(weights seed 1234, inputs `bench.synth_inputs(seed=7)`: 282 prompt frames of 0.1*randn audio, 150 text ids, 938
frames, NFE 32, cfg 2.0, sway -1.0, seed 0).  The UNMODIFIED reference (`/root/reference/src/f5_tts/model/cfm.py`
`CFM.sample`, imported through oracle/ref_shims.py) runs it in fp32 on the CPU; the fixture keeps y0 and the
trajectory at steps 1, 8, 16 and 32 so that the GPU tests and `bench.py` (`parity` key) can report the drift per step
next to the reference's own fp16-vs-fp32 drift.  `cfg2_fp16` additionally runs the reference module in fp16 on the CPU
(the dtype the reference uses on GPU, utils_infer.py:190-199) with the same injected y0 and stores ITS drift at the same
steps — the yardstick SURVEY.md §9.4 measured at N = 400.

The other BASELINE configs (cfg3 B=8 var-len, cfg4 B=8 EPSS-16, cfg5 UNetT B=8) are checked against the CPU oracle
computed live on the GPU box (tests/test_gpu_fullsize.py): their outputs are 3-6 MB each, too large to commit, and
the oracle is pinned bit-exactly to the reference by tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import f5_oracle as O  # noqa: E402
from oracle.make_golden import GOLD, build_reference, rel_l2  # noqa: E402

STEPS_KEPT = (1, 8, 16, 32)


def cfg2_inputs():
    g = torch.Generator().manual_seed(7)  # == bench.synth_inputs(WORKLOADS["cfg2"])
    wav = 0.1 * torch.randn(1, 282 * 256, generator=g)
    text = torch.randint(0, 2545, (1, 150), generator=g)
    return wav, text


def cfg2(fp16_too: bool):
    cfg = O.f5tts_base()
    sd = O.synthetic_state_dict(cfg, seed=1234)
    model = build_reference(cfg, sd)
    wav, text = cfg2_inputs()
    t0 = time.time()
    with torch.no_grad():
        out, traj = model.sample(cond=wav, text=text, duration=938, steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0,
                                 seed=0)
    print(f"[cfg2] reference fp32: out {tuple(out.shape)} in {time.time() - t0:.1f} s")
    save = dict(y0=traj[0].numpy(), out=out.numpy(), steps=32, cfg_strength=2.0, sway=-1.0, seed=0, wseed=1234,
                n_ref=282, kept=np.asarray(STEPS_KEPT))
    for k in STEPS_KEPT:
        save[f"traj_{k}"] = traj[k].numpy()
    res = O.sample(sd, cfg, wav, text, 938, steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=0)
    print(f"[cfg2] oracle-vs-reference final rel-L2 {rel_l2(res.out, out):.3e}")
    if fp16_too:
        # the reference's own reduced-precision path (fp16 parameters + activations, as load_checkpoint casts them on
        # GPU) on the same y0: inject through the RNG hook — sample() draws randn(938, 100) after manual_seed(seed).
        half = build_reference(cfg, sd).half()
        y0 = traj[0]
        orig = torch.randn

        def fake_randn(*a, **k):
            return y0[0].to(k.get("dtype", torch.float32))

        torch.randn = fake_randn
        try:
            t0 = time.time()
            with torch.no_grad():
                out16, traj16 = half.sample(cond=wav, text=text, duration=938, steps=32, cfg_strength=2.0,
                                            sway_sampling_coef=-1.0, seed=0)
        finally:
            torch.randn = orig
        gen = slice(282, 938)
        drift = [rel_l2(traj16[k].float()[:, gen], traj[k][:, gen]) for k in STEPS_KEPT]
        print(f"[cfg2] reference fp16 (CPU) in {time.time() - t0:.1f} s: drift at steps {STEPS_KEPT} = "
              + ", ".join(f"{d:.2e}" for d in drift))
        save["ref_fp16_drift"] = np.asarray(drift)
    np.savez_compressed(os.path.join(GOLD, "cfg2_full_nfe32.npz"), **save)


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    what = sys.argv[1:] or ["cfg2"]
    cfg2(fp16_too="cfg2_fp16" in what)


if __name__ == "__main__":
    main()

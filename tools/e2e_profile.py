"""GPU diagnostic: where the end-to-end call (infer_process: host audio -> host waveform) spends host time on top of the
device time of the sampler — cProfile of a few warm calls, sorted by cumulative time."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from f5_tts_b200 import infer as INF  # noqa: E402

dev = "cuda:0"
w = bench.WORKLOADS["cfg2"]
model, voc, _ = bench.build_gpu_model(w["arch"], dev)
wav, text, duration, lens = bench.synth_inputs(w)
n_ref_txt = 45
ref_text = ("some call me nature others call me mother na" + ".")[:n_ref_txt - 1] + "."
gen_text = ("i have been a silent spectator watching species evolve and empires rise and fall but always remember "
            "i am mighty")[: w["nt"] - n_ref_txt - 1]
fix_dur = (w["frames"][0] + 0.25) * 256 / 24000.0
audio_h = wav.clone().pin_memory()


def e2e():
    return INF.infer_process((audio_h, 24000), ref_text, gen_text, model, voc, nfe_step=w["nfe"],
                             cfg_strength=bench.CFG_STRENGTH, sway_sampling_coef=bench.SWAY, fix_duration=fix_dur,
                             device=dev, show_info=lambda *_: None)


wd, td, dd, ld = (t.to(dev) for t in (wav, text, duration, lens))


def dev_only():
    out = bench.hot_path(model, voc, wd, td, dd, ld, w["nfe"], w["frames"][0])
    torch.cuda.synchronize()
    return out


for fn, name in ((dev_only, "device-resident hot_path + sync"), (e2e, "infer_process")):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call (wall)", flush=True)

pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    e2e()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)

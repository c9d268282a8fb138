"""Tiny driver for the ncu capture of the bandwidth-bound kernels: the cfg2 hot path with 2 NFE steps (mel STFT, text
blocks with dwconv+LN / GRN, CFG+Euler, final row norm, Vocos decode with dwconv+LN and the two ISTFT kernels)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

dev = torch.device("cuda", 0)
model, voc, cfg = bench.build_gpu_model("f5tts_base", dev)
w = bench.WORKLOADS["cfg2"]
wav, text, duration, lens = (t.to(dev) for t in bench.synth_inputs(w))
for _ in range(2):
    bench.hot_path(model, voc, wav, text, duration, lens, 2, w["frames"][0])
torch.cuda.synchronize()
print("done")

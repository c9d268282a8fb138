"""GPU diagnostic: device time of one CFM.sample call (default cfg2, NFE 32) — used with the instrumented build and
F5_DIAG_SKIP=<kernel class> to read the in-situ cost of that class as the difference to the full step."""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

dev = "cuda:0"
name = os.environ.get("STEP_WORKLOAD", "cfg2")
w = bench.WORKLOADS[name]
model, voc, _ = bench.build_gpu_model(w["arch"], dev)
wav, text, duration, lens = (t.to(dev) for t in bench.synth_inputs(w))
fn = lambda: bench.hot_path(model, voc, wav, text, duration, lens, w["nfe"])  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{name} ms_per_call {ms:.3f}  per_NFE_us {ms * 1e3 / w['nfe']:.1f}  skip={os.environ.get('F5_DIAG_SKIP', '')}")

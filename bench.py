#!/usr/bin/env python
"""bench.py — RTF / mel-frames-per-second of the F5-TTS ODE-sampling hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg3|cfg4|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic utterances:
    mel front-end (STFT kernel) -> CFM.sample (NFE x backbone + CFG + Euler) -> Vocos decode (ISTFT kernel).
Default workload = BASELINE.json configs[1] ("cfg2"): F5-TTS Base, batch 1, 10 s total (938 frames, 282 prompt frames),
NFE 32, cfg 2.0, sway -1.0, random-init weights in the released checkpoint layout, fp16 tensor-core
operands with fp32 accumulation / residual / ODE state.  With --gpus N every rank runs the same workload on its own
GPU (weak scaling, utterances are independent) and all-gathers the finished mel + audio each step (NCCL).

`value`  = generated mel frames per second, whole job, inputs already resident in HBM.
`e2e`    = the same metric through the public API (CFM.sample + Vocos.decode) with PINNED HOST buffers: H2D of the
           reference audio + token ids and D2H of the waveform inside the timed region.
`roofline` = dominant kernel (tcgen05 GEMM, FF1 shape of the workload) timed in isolation with CUDA events over a
           weight set larger than L2; `step_tensor` = whole-step algorithmic FLOP/s against the sustained peak.
`cpu_baseline` / --impl reference = the CPU oracle port of the reference (oracle/f5_oracle.py, fp32, all host cores)
           on a bounded sample (few NFE) of the same workload, scaled linearly to the full NFE.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: backbone, B (per GPU), frames, prompt frames, text tokens, NFE
    "cfg2": dict(arch="f5tts_base", B=1, frames=[938], ref=[282], nt=150, nfe=32),
    "cfg3": dict(arch="f5tts_base", B=8, frames=[469, 670, 871, 1072, 1272, 1473, 1674, 1875],
                 ref=[141, 201, 261, 322, 382, 442, 502, 562], nt=300, nfe=32),
    "cfg4": dict(arch="f5tts_base", B=8, frames=[938] * 8, ref=[282] * 8, nt=150, nfe=16),
    "cfg5": dict(arch="e2tts_base", B=8, frames=[938] * 8, ref=[282] * 8, nt=150, nfe=32),
}
CFG_STRENGTH, SWAY = 2.0, -1.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops", 1590.0),
                    tf_sus=d.get("bf16_tflops_sustained", 1400.0), src="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


def synth_inputs(w, seed=7):
    g = torch.Generator().manual_seed(seed)
    B = w["B"]
    n_ref = max(w["ref"])
    wav = 0.1 * torch.randn(B, n_ref * 256, generator=g)
    text = torch.randint(0, 2545, (B, w["nt"]), generator=g)
    duration = torch.tensor(w["frames"], dtype=torch.long)
    lens = torch.tensor(w["ref"], dtype=torch.long)
    return wav, text, duration, lens


class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows]
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = float(rows[0][2])
            out["samples"] = len(rows)
            out["power_w_max"] = max(float(r[3]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any("Active" in r[5 + i] and "Not" not in r[5 + i] for r in rows):
                    out["reasons"].append(n)
        except Exception as e:  # noqa: BLE001
            out["error"] = str(e)
        finally:
            if self.path and os.path.exists(self.path):
                os.unlink(self.path)
        return out


def build_gpu_model(arch_name, dev):
    import f5_tts_b200 as F5
    from f5_tts_b200.vocoder import Vocos
    from oracle import f5_oracle as O  # synthetic weights only (shared with the CPU arm); not on the timed path

    cfg = getattr(O, arch_name)()
    cls = F5.DiT if cfg.backbone == "DiT" else F5.UNetT
    model = F5.CFM(transformer=cls(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult,
                                   text_dim=cfg.text_dim, text_mask_padding=cfg.text_mask_padding,
                                   conv_layers=cfg.conv_layers, pe_attn_head=cfg.pe_attn_head,
                                   text_num_embeds=cfg.text_num_embeds, mel_dim=100))
    sd = O.synthetic_state_dict(cfg, seed=1234)
    model.load_state_dict(sd, strict=True)
    voc = Vocos()
    voc.load_state_dict(O.synthetic_vocos_state_dict(), strict=False)
    return model.to(dev), voc.to(dev), cfg


def hot_path(model, voc, wav, text, duration, lens, nfe, frames0=None):
    """mel front-end + CFM.sample + vocoder; returns (mel [B,N,100], audio [B, nw])."""
    B = wav.shape[0]
    if B == 1:
        out, _ = model.sample(wav, text, frames0 if frames0 is not None else int(duration[0]), steps=nfe, cfg_strength=CFG_STRENGTH,
                              sway_sampling_coef=SWAY, seed=0)
        ref = wav.shape[-1] // 256
    else:
        cond = model.mel_spec(wav, frames_last=False)
        out, _ = model.sample(cond, text, duration, lens=lens, steps=nfe, cfg_strength=CFG_STRENGTH,
                              sway_sampling_coef=SWAY, seed=0)
        ref = int(lens.min())
    audio = voc.decode(out[:, ref:, :].permute(0, 2, 1).float())
    return out, audio


def _graph_time_us(fn, n_launch, rounds=5):
    """Average device time per launch of `fn` (which enqueues n_launch kernels): captured into a CUDA graph so that
    host-side launch cost (ctypes + tensor-map encoding) is not in the measurement; CUDA events on the launching stream."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rounds):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (rounds * n_launch)


def isolated_gemm_roofline(M, seq, peaks, dev):
    """Dominant kernel alone: FF1 GEMM (M x 2048 x 1024, bias + GELU-tanh epilogue), 48 distinct weight matrices
    (192 MB > L2) launched back to back from a CUDA graph, CUDA events on the launching stream.  Also times the other
    GEMM shapes of a block and the attention kernel the same way (`kernels`)."""
    from f5_tts_b200 import ops

    nw = 48
    g = torch.Generator().manual_seed(3)
    bn, pair = 0, 0  # the planner picks the tile shape, exactly as inside the engine

    def run(N, K, epi, act, tag, **kw):
        n_w = max(24, -(-192_000_000 // (N * K * 2)))  # distinct weights > L2: every launch streams W from HBM
        a = [torch.randn(M, K, generator=g).half().to(dev) for _ in range(2)]
        w = [(torch.randn(N, K, generator=g) / 32).half().to(dev) for _ in range(n_w)]
        b = torch.randn(N, generator=g).to(dev)
        if epi == ops.EPI_RESID:
            kw["resid"] = torch.zeros(M, N, device=dev)
            kw["gate"] = torch.randn(N, generator=g).to(dev)
        us = _graph_time_us(lambda: [ops.linear(a[i % 2], w[i], b, epi=epi, act=act, bn=bn, pair=pair, static_w=True, **kw)
                                     for i in range(n_w)], n_w)
        fl = 2.0 * M * N * K
        tile = ops.gemm_tile(M, N, K, epi, act)
        return dict(kernel=tag, shape=[M, N, K], tile=f"{'256' if tile[1] else '128'}x{tile[0]}{' cta_group::2' if tile[1] else ''}",
                    us_per_launch=round(us, 2), tflops=round(fl / us / 1e6, 1), frac=round(fl / us / 1e6 / peaks["tf"], 4))

    rows = [run(2048, 1024, ops.EPI_F16, ops.ACT_GELU_TANH, "FF1 (bias+GELU-tanh, fp16 out)"),
            run(3072, 1024, ops.EPI_QKV_ROPE, ops.ACT_NONE, "QKV (bias+RoPE)", seq=seq,
                rope=ops.rope_tables(seq, dev), inner=1024, pe_heads=1),
            run(1024, 1024, ops.EPI_RESID, ops.ACT_NONE, "out-proj (gate, TMA reduce-add)"),
            run(1024, 2048, ops.EPI_RESID, ops.ACT_NONE, "FF2 (gate, TMA reduce-add)")]
    qkv = [torch.randn(M, 3072, generator=g).half().to(dev) for _ in range(3)]
    us = _graph_time_us(lambda: [ops.attention(qkv[i % 3], M // seq, seq, 16) for i in range(12)], 12)
    afl = 4.0 * (M // seq) * 16 * seq * seq * 64
    rows.append(dict(kernel="attention (dh 64, non-causal)", shape=[M // seq, seq, 16], us_per_launch=round(us, 2),
                     tflops=round(afl / us / 1e6, 1), frac=round(afl / us / 1e6 / peaks["tf"], 4)))
    ff1 = rows[0]
    return dict(bound="tensor",
                kernel=f"gemm_tcgen05_kernel<tile {ff1['tile']},EPI_F16,GELU_TANH> (FF1)",
                shape=ff1["shape"], us_per_launch=ff1["us_per_launch"], achieved=ff1["tflops"], peak=peaks["tf"],
                unit="TFLOP/s", frac=ff1["frac"], peak_source=peaks["src"] + " bf16_tflops (burst)",
                flops_per_launch=2.0 * M * 2048 * 1024, traffic=load_traffic(), kernels=rows,
                method=">= 192 MB of distinct weights per shape (> L2) launched back to back from a CUDA graph, CUDA events")


def load_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("gemm_ff1_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            return None
    return None


def host_cores():
    """(usable cores, torch threads): honours the cgroup CPU quota (the GPU box shows 128 CPUs but grants 16 cores;
    oversubscribing it made the fp32 oracle 50x slower) and the affinity mask."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(round(int(quota) / int(period)))))
    except Exception:  # noqa: BLE001
        pass
    return n, n


class CpuOracle:
    """CPU oracle (port of the reference, fp32) on all host cores.  Built once (1.3 GB of synthetic weights, one
    untimed warm-up); `measure(budget_s)` times NFE 1 and NFE b, with b chosen so that the sample costs about `budget_s`
    seconds of CPU work, and scales linearly to the workload NFE."""

    def __init__(self, w):
        from oracle import f5_oracle as O

        self.O, self.w = O, w
        self.cores, self.threads = host_cores()
        torch.set_num_threads(self.threads)
        self.cfg = getattr(O, w["arch"])()
        self.sd = O.synthetic_state_dict(self.cfg, seed=1234)
        self.vsd = O.synthetic_vocos_state_dict()
        self.wav, self.text, self.duration, self.lens = synth_inputs(w)
        self.run(1)  # untimed warm-up: thread pool, oneDNN primitive caches, first touch of the weights

    def run(self, n):
        O, w = self.O, self.w
        t0 = time.perf_counter()
        if w["B"] == 1:
            res = O.sample(self.sd, self.cfg, self.wav, self.text, int(self.duration[0]), steps=n,
                           cfg_strength=CFG_STRENGTH, sway_sampling_coef=SWAY, seed=0)
            ref = self.wav.shape[-1] // 256
        else:
            cond = O.mel_spectrogram(self.wav).permute(0, 2, 1)
            res = O.sample(self.sd, self.cfg, cond, self.text, self.duration, lens=self.lens, steps=n,
                           cfg_strength=CFG_STRENGTH, sway_sampling_coef=SWAY, seed=0)
            ref = int(self.lens.min())
        t1 = time.perf_counter()
        O.vocos_decode(self.vsd, res.out[:, ref:, :].permute(0, 2, 1))
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    def measure(self, budget_s=12.0):
        w = self.w
        ta, tv = self.run(1)
        b = max(3, min(w["nfe"], 1 + int(budget_s / max(ta, 1e-3))))
        tb, _ = self.run(b)
        per = (tb - ta) / (b - 1)
        fixed = max(ta - per, 0.0)
        total = fixed + w["nfe"] * per + tv
        gen = sum(f - r for f, r in zip(w["frames"], w["ref"]))
        return dict(value=gen / total, unit="mel_frames/s", cores=self.cores, threads=self.threads, kind="port",
                    sample=f"oracle fp32 CPU: NFE 1 and {b} of {w['nfe']} timed ({ta + tb + tv:.1f} s), scaled linearly "
                           f"(per-NFE {per:.2f} s, fixed {fixed:.2f} s, vocoder {tv:.3f} s)",
                    rtf=total / (gen * 256 / 24000.0), seconds_full_extrapolated=total)


def cpu_leg(w, budget_s=12.0):
    return CpuOracle(w).measure(budget_s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    gen_frames = sum(f - r for f, r in zip(w["frames"], w["ref"]))
    config = dict(workload=f"{args.workload}: {w['arch']} B={w['B']}/GPU frames={w['frames'] if w['B'] > 1 and len(set(w['frames'])) > 1 else w['frames'][0]} "
                           f"prompt={w['ref'][0] if len(set(w['ref'])) == 1 else w['ref']} NFE={w['nfe']} cfg={CFG_STRENGTH} sway={SWAY}",
                  global_batch=w["B"] * max(world, 1), parallelism=f"dp{world} (utterance sharding, all-gather of mel+audio per step)",
                  l2="no explicit flush: each step streams 0.67 GB of fp16 weights (> 126 MB L2)")

    if args.impl == "reference":
        if rank != 0:
            return
        vals, last = [], None
        oracle = CpuOracle(w)
        n_runs = args.warmup + args.steps
        budget = min(15.0, max(2.0, 150.0 / n_runs))  # the whole reference run stays within a few minutes
        for i in range(n_runs):
            last = oracle.measure(budget)
            if i >= args.warmup:
                vals.append(last["value"])
        v = statistics.mean(vals)
        sec = gen_frames * 256 / 24000.0
        line = dict(metric="mel_frames_per_sec", value=v, unit="mel_frames/s", n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=1e3 * gen_frames / v, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="fp32", data="synthetic", config=config, impl="reference",
                    rtf=(gen_frames / v) / sec,
                    cpu_baseline=dict(value=v, unit="mel_frames/s", cores=last["cores"], kind="port", sample=last["sample"]),
                    e2e=dict(value=v, unit="mel_frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA (B200) device; there is no CPU fallback"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from f5_tts_b200 import _lib

    peaks = load_peaks()
    model, voc, cfg = build_gpu_model(w["arch"], dev)
    wav, text, duration, lens = synth_inputs(w)
    wav_d, text_d, dur_d, lens_d = wav.to(dev), text.to(dev), duration.to(dev), lens.to(dev)
    nfe = w["nfe"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(mel, audio):
        if dist is None:
            return
        m = torch.empty((world * mel.shape[0],) + tuple(mel.shape[1:]), device=dev, dtype=mel.dtype)
        a = torch.empty((world * audio.shape[0],) + tuple(audio.shape[1:]), device=dev, dtype=audio.dtype)
        dist.all_gather_into_tensor(m, mel.contiguous())
        dist.all_gather_into_tensor(a, audio.contiguous())

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        gather(*hot_path(model, voc, wav_d, text_d, dur_d, lens_d, nfe, w['frames'][0]))
    barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        barrier()
        e0.record()
        for _ in range(args.steps):
            gather(*hot_path(model, voc, wav_d, text_d, dur_d, lens_d, nfe, w['frames'][0]))
        e1.record()
        barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    clocks = cs.summary()

    # ---- end to end through the public API with pinned host buffers -----------------------------------------
    wav_h, text_h = wav.pin_memory(), text.pin_memory()
    n_audio = 256 * (max(w["frames"]) - min(w["ref"]) - 1)
    out_h = torch.empty((w["B"], n_audio), dtype=torch.float32).pin_memory()

    def e2e_step():
        wd, td = wav_h.to(dev, non_blocking=True), text_h.to(dev, non_blocking=True)
        _, audio = hot_path(model, voc, wd, td, dur_d, lens_d, nfe, w['frames'][0])
        out_h.copy_(audio, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller holds the waveform on the host here

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        e2e_step()
    g1.record()
    barrier()
    ms_e2e = max(g0.elapsed_time(g1), (time.perf_counter() - t0) * 1e3) / args.steps

    # ---- max over ranks --------------------------------------------------------------------------------------
    if dist is not None:
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
    total_frames = gen_frames * world
    value = total_frames / (ms * 1e-3)
    e2e_value = total_frames / (ms_e2e * 1e-3)
    audio_sec = gen_frames * 256 / 24000.0

    line = None
    if rank == 0:
        flops = model.transformer.sample_flops(w["B"], max(w["frames"]), nfe, CFG_STRENGTH)
        roof = isolated_gemm_roofline(2 * w["B"] * max(w["frames"]), max(w["frames"]), peaks, dev)
        step_tf = flops / (ms * 1e-3) / 1e12
        roof["step_tensor"] = dict(flops_per_step=flops, achieved=round(step_tf, 1), peak=peaks["tf_sus"],
                                   unit="TFLOP/s", frac=round(step_tf / peaks["tf_sus"], 4),
                                   note="whole hot-path step (all kernels, launch gaps included) vs sustained bf16 peak")
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_leg(w)  # timed at N = 1 only
        line = dict(metric="mel_frames_per_sec", value=value, unit="mel_frames/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="fp16 operands / fp32 accumulate+state", data="synthetic (random-init weights, released checkpoint layout)",
                    config=config, rtf=(ms * 1e-3) / audio_sec, rtf_e2e=(ms_e2e * 1e-3) / audio_sec,
                    e2e=dict(value=e2e_value, unit="mel_frames/s", ms_per_step=ms_e2e,
                             h2d_bytes_per_step=wav.numel() * 4 + text.numel() * 8, d2h_bytes_per_step=out_h.numel() * 4),
                    gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu, clocks=clocks, impl="b200")
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

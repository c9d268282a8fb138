"""Request-level serving front end on top of the B200 sampler — mirror of the reference's Triton Python backend
(``runtime/triton_trtllm/model_repo_f5_tts/f5_tts/1/model.py`` + ``config.pbtxt``), SURVEY.md §8(f)-3.

I/O contract (config.pbtxt:43-72), one request =

    reference_wav      float32 [1, n]   reference audio at `reference_sample_rate`
    reference_wav_len  int32   [1]      valid samples of reference_wav              (optional: default n)
    reference_text     str              transcript of the reference audio
    target_text        str              text to synthesise
 -> waveform           float32 [m]      generated audio (reference part removed), 24 kHz

`F5TTSRequestProcessor.execute(requests)` is ``TritonPythonModel.execute`` (model.py:176-269): per request the
reference RMS gain (model.py:208-211) and the duration estimate ``int(n_ref_frames * (1 + len(target) / len(reference)))``
in UTF-8 bytes (model.py:223-227); texts are ``reference_text + target_text`` through ``convert_char_to_pinyin`` and the
vocabulary (model.py:238-239); the whole batch goes through ONE sampler call with 32 NFE, cfg 2.0, sway -1
(f5_tts_trtllm.py:239,310); each generated mel is cut to ``[ref_len, estimated_len)`` and vocoded (model.py:256-263).
Where the reference hands the batch to a TensorRT-LLM engine with per-sample length masks, this module calls
``CFM.sample(..., exact_varlen=True)``: every request is computed exactly as if it were alone in the batch, so the
answer to a request does not depend on what it was batched with (tested bit-for-bit on the GPU).

`DynamicBatcher` is the ``dynamic_batching { max_queue_delay_microseconds: 1000 }`` / ``max_batch_size: 4`` policy of
config.pbtxt:17-20 as a worker thread: callers `submit()` single requests and receive futures.

No tritonserver / pb_utils in this image: requests are plain dicts of numpy arrays / strings (the decoded form of the
pb_utils tensors), responses are numpy arrays.  INTEGRATION.md shows the 10-line ``TritonPythonModel`` adapter.
"""
from __future__ import annotations

import concurrent.futures
import queue
import threading
import time

import numpy as np
import torch

from . import infer as _infer

MAX_BATCH_SIZE = 4            # config.pbtxt:17
MAX_QUEUE_DELAY_US = 1000     # config.pbtxt:18-20
MAX_MEL_LEN = 4096            # model.py:115
NFE_STEPS = 32                # f5_tts_trtllm.py:239
CFG_STRENGTH = 2.0            # f5_tts_trtllm.py:310
SWAY_COEF = -1.0              # f5_tts_trtllm.py:248 (the time grid there is the sway-sampled one)


def _as_text(v) -> str:
    """reference_text / target_text arrive as TYPE_STRING tensors of shape [1, 1] holding bytes (model.py:192-197)."""
    if isinstance(v, np.ndarray):
        v = v.reshape(-1)[0]
    if isinstance(v, bytes):
        v = v.decode("utf-8")
    return str(v)


class F5TTSRequestProcessor:
    """`model_obj` is a loaded `f5_tts_b200.CFM` (infer.load_model), `vocoder` a loaded `vocoder.Vocos`."""

    def __init__(self, model_obj, vocoder, *, reference_sample_rate: int = 24000, device=None,
                 max_batch_size: int = MAX_BATCH_SIZE, nfe_step: int = NFE_STEPS, cfg_strength: float = CFG_STRENGTH,
                 sway_sampling_coef: float = SWAY_COEF, seed: int | None = None):
        self.model, self.vocoder = model_obj, vocoder
        self.device = device or _infer.device
        self.target_audio_sample_rate = _infer.target_sample_rate
        self.reference_sample_rate = int(reference_sample_rate)
        self.target_rms = _infer.target_rms            # model.py:111
        self.hop_length = _infer.hop_length
        self.max_mel_len = MAX_MEL_LEN
        self.max_batch_size = int(max_batch_size)
        self.nfe_step, self.cfg_strength, self.sway = int(nfe_step), float(cfg_strength), sway_sampling_coef
        self.seed = seed
        self._resampler = None
        if self.reference_sample_rate != self.target_audio_sample_rate:
            import torchaudio

            self._resampler = torchaudio.transforms.Resample(self.reference_sample_rate, self.target_audio_sample_rate)

    # -- one request -> (wave tensor [1, n] on the host, rms, texts) ------------------------------------------------
    def _decode(self, request: dict):
        wav = torch.as_tensor(np.asarray(request["reference_wav"], dtype=np.float32))
        if wav.ndim == 1:
            wav = wav[None, :]
        if wav.shape[0] != 1:
            raise ValueError("Only support batch size 1 for now.")  # model.py:205
        n = request.get("reference_wav_len")
        if n is not None:
            wav = wav[:, : int(np.asarray(n).reshape(-1)[0])]
        rms = torch.sqrt(torch.mean(torch.square(wav)))
        if rms < self.target_rms:
            wav = wav * self.target_rms / rms
        if self._resampler is not None:
            wav = self._resampler(wav)
        return wav, float(rms), _as_text(request["reference_text"]), _as_text(request["target_text"])

    @torch.inference_mode()
    def execute(self, requests: list) -> list:
        """model.py:176-269 for up to `max_batch_size` requests; longer lists are processed in slices of that size."""
        out: list = []
        for i in range(0, len(requests), self.max_batch_size):
            out.extend(self._execute_batch(requests[i: i + self.max_batch_size]))
        return out

    def _execute_batch(self, requests: list) -> list:
        if not requests:
            return []
        decoded = [self._decode(r) for r in requests]
        # mel front end per request (lengths differ), padded into one [B, n_max, 100] batch (model.py:229-236)
        mels = [self.model.mel_spec(w.to(self.device), frames_last=False)[0] for w, _, _, _ in decoded]  # [n_i, 100]
        ref_len = [int(m.shape[0]) for m in mels]
        est = [int(n * (1 + len(tt.encode("utf-8")) / len(rt.encode("utf-8")))) for n, (_, _, rt, tt) in
               zip(ref_len, decoded)]
        est = [min(e, self.max_mel_len) for e in est]
        cond = torch.zeros((len(mels), max(ref_len), mels[0].shape[-1]), dtype=torch.float32, device=self.device)
        for i, m in enumerate(mels):
            cond[i, : m.shape[0]] = m
        tokens = _infer.convert_char_to_pinyin([rt + tt for _, _, rt, tt in decoded], polyphone=True)
        lens = torch.tensor(ref_len, dtype=torch.long, device=self.device)
        if len(requests) == 1:
            generated, _ = self.model.sample(cond=cond, text=tokens, duration=est[0], lens=lens, steps=self.nfe_step,
                                             cfg_strength=self.cfg_strength, sway_sampling_coef=self.sway, seed=self.seed)
        else:
            dur = torch.tensor(est, dtype=torch.long, device=self.device)
            generated, _ = self.model.sample(cond=cond, text=tokens, duration=dur, lens=lens, steps=self.nfe_step,
                                             cfg_strength=self.cfg_strength, sway_sampling_coef=self.sway, seed=self.seed,
                                             exact_varlen=True)
        # the sampler lengthens a request whose text has more tokens than frames (cfm.py:134-137)
        n_tok = [len(t) for t in tokens]
        est = [max(max(nt, rl) + 1, e) for nt, rl, e in zip(n_tok, ref_len, est)]
        waves = []
        for i, (_, rms, _, _) in enumerate(decoded):
            mel = generated[i: i + 1, ref_len[i]: est[i], :].to(torch.float32).permute(0, 2, 1).contiguous()
            audio = self.vocoder.decode(mel)
            if rms < self.target_rms:
                audio = audio * rms / self.target_rms
            waves.append(audio.squeeze(0))
        return [w.cpu().numpy() for w in waves]


class DynamicBatcher:
    """config.pbtxt:17-20: requests that arrive within `max_queue_delay_us` of the first waiting one (at most
    `max_batch_size`) are executed together.  `submit()` returns a `concurrent.futures.Future` of the waveform."""

    def __init__(self, processor: F5TTSRequestProcessor, max_batch_size: int | None = None,
                 max_queue_delay_us: int = MAX_QUEUE_DELAY_US):
        self.processor = processor
        self.max_batch_size = int(max_batch_size or processor.max_batch_size)
        self.delay = max_queue_delay_us * 1e-6
        self.batches_run: list = []  # sizes of the executed batches (observability / tests)
        self._q: queue.Queue = queue.Queue()
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def submit(self, request: dict) -> concurrent.futures.Future:
        if self._stop.is_set():
            raise RuntimeError("batcher is closed")
        fut: concurrent.futures.Future = concurrent.futures.Future()
        self._q.put((request, fut))
        return fut

    def _loop(self):
        while not self._stop.is_set():
            try:
                first = self._q.get(timeout=0.05)
            except queue.Empty:
                continue
            batch, deadline = [first], time.monotonic() + self.delay
            while len(batch) < self.max_batch_size:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                try:
                    batch.append(self._q.get(timeout=left))
                except queue.Empty:
                    break
            self.batches_run.append(len(batch))
            try:
                waves = self.processor.execute([r for r, _ in batch])
                for (_, fut), w in zip(batch, waves):
                    fut.set_result(w)
            except Exception as e:  # noqa: BLE001 — the error belongs to the callers, the worker keeps serving
                for _, fut in batch:
                    if not fut.done():
                        fut.set_exception(e)

    def close(self):
        self._stop.set()
        self._thread.join()
        while not self._q.empty():
            _, fut = self._q.get_nowait()
            fut.set_exception(RuntimeError("batcher closed"))

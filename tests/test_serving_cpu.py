"""Host logic of the request-level front end (f5_tts_b200/serving.py) against the reference's Triton Python backend
(runtime/triton_trtllm/model_repo_f5_tts/f5_tts/1/model.py:176-269, config.pbtxt) — recording fakes stand in for the
sampler and the vocoder, so this runs without a GPU."""
import time

import numpy as np
import torch

from f5_tts_b200 import infer, serving


class _FakeMel:
    def __call__(self, wav, frames_last=False):  # [1, n] -> [1, 1 + n // 256, 100], value = mean of the wave
        t = 1 + wav.shape[-1] // 256
        return torch.full((1, t, 100), float(wav.mean()))


class _FakeModel:
    def __init__(self):
        self.calls = []
        self.mel_spec = _FakeMel()

    def sample(self, cond, text, duration, lens=None, steps=32, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=None,
               exact_varlen=False):
        durs = [duration] if isinstance(duration, int) else [int(d) for d in duration]
        self.calls.append(dict(cond=cond.clone(), text=text, duration=durs, lens=lens.tolist(), steps=steps,
                               cfg=cfg_strength, sway=sway_sampling_coef, exact=exact_varlen))
        n = max(max(durs), max(max(len(t) for t in text), max(lens.tolist())) + 1)
        return torch.arange(n, dtype=torch.float32).view(1, -1, 1).repeat(len(durs), 1, 100), None


class _FakeVocoder:
    def decode(self, mel):  # [1, 100, n] -> [1, 256 (n - 1)], constant 0.5 so gains are visible
        return torch.full((1, 256 * (mel.shape[-1] - 1)), 0.5)


def _request(seconds, level, ref_text, target_text, pad=0):
    n = int(seconds * 24000)
    wav = np.full((1, n + pad), level, dtype=np.float32)
    return {"reference_wav": wav, "reference_wav_len": np.array([n], dtype=np.int32),
            "reference_text": np.array([[ref_text.encode()]], dtype=object), "target_text": target_text}


def test_execute_follows_the_triton_backend():
    model, voc = _FakeModel(), _FakeVocoder()
    proc = serving.F5TTSRequestProcessor(model, voc, device="cpu")
    reqs = [_request(1.0, 0.02, "short reference.", "something to say, twice as long as that.", pad=5000),
            _request(2.0, 0.3, "a longer reference text here.", "brief.")]
    waves = proc.execute(reqs)
    assert len(model.calls) == 1 and len(waves) == 2                 # one batched sampler call (model.py:244)
    c = model.calls[0]
    assert (c["steps"], c["cfg"], c["sway"], c["exact"]) == (32, 2.0, -1.0, True)  # f5_tts_trtllm.py:239,310
    ref_len = [1 + 24000 // 256, 1 + 48000 // 256]                   # reference_wav_len cuts the padding (model.py:206)
    assert c["lens"] == ref_len and c["cond"].shape == (2, ref_len[1], 100)
    # RMS below 0.1 is raised to 0.1 before the mel, louder references are left alone (model.py:208-211)
    assert torch.allclose(c["cond"][0, : ref_len[0]], torch.full((ref_len[0], 100), 0.1), atol=1e-6)
    assert torch.allclose(c["cond"][0, ref_len[0]:], torch.zeros(ref_len[1] - ref_len[0], 100))
    assert torch.allclose(c["cond"][1], torch.full((ref_len[1], 100), 0.3), atol=1e-6)
    # duration estimate in UTF-8 bytes (model.py:223-227); text = reference + target, no separator (model.py:199)
    texts = [("short reference.", "something to say, twice as long as that."), ("a longer reference text here.", "brief.")]
    est = [int(n * (1 + len(t.encode()) / len(r.encode()))) for n, (r, t) in zip(ref_len, texts)]
    assert c["duration"] == est
    assert c["text"] == infer.convert_char_to_pinyin([r + t for r, t in texts])
    # generated part only, vocoded, input gain undone (model.py:259-262)
    assert [len(w) for w in waves] == [256 * (e - n - 1) for e, n in zip(est, ref_len)]
    assert np.allclose(waves[0], 0.5 * 0.02 / 0.1, atol=1e-6) and np.allclose(waves[1], 0.5)


def test_single_request_and_slicing_by_max_batch_size():
    model, voc = _FakeModel(), _FakeVocoder()
    proc = serving.F5TTSRequestProcessor(model, voc, device="cpu", max_batch_size=2)
    reqs = [_request(1.0, 0.2, "reference text.", f"target number {i}.") for i in range(5)]
    waves = proc.execute(reqs)
    assert len(waves) == 5 and [len(c["duration"]) for c in model.calls] == [2, 2, 1]
    assert model.calls[2]["exact"] is False                         # a single request is the plain B = 1 call
    try:
        proc.execute([{"reference_wav": np.zeros((2, 100), np.float32), "reference_text": "a.", "target_text": "b."}])
        raise AssertionError("two reference waves in one request must be refused (model.py:205)")
    except ValueError:
        pass


def test_dynamic_batcher_groups_requests_that_arrive_together():
    model, voc = _FakeModel(), _FakeVocoder()
    proc = serving.F5TTSRequestProcessor(model, voc, device="cpu")
    batcher = serving.DynamicBatcher(proc, max_batch_size=4, max_queue_delay_us=200_000)
    try:
        futs = [batcher.submit(_request(1.0, 0.2, "reference text.", f"target {i}.")) for i in range(6)]
        waves = [f.result(timeout=30) for f in futs]
        assert all(isinstance(w, np.ndarray) and w.ndim == 1 for w in waves)
        assert batcher.batches_run[0] == 4 and sum(batcher.batches_run) == 6   # max_batch_size, then the rest
        time.sleep(0.3)
        lone = batcher.submit(_request(1.0, 0.2, "reference text.", "alone."))
        assert lone.result(timeout=30).ndim == 1 and batcher.batches_run[-1] == 1
        bad = batcher.submit({"reference_wav": np.zeros((2, 10), np.float32), "reference_text": "a.", "target_text": "b."})
        try:
            bad.result(timeout=30)
            raise AssertionError("the request's error must reach its future")
        except ValueError:
            pass
        assert batcher.submit(_request(1.0, 0.2, "reference text.", "still serving.")).result(timeout=30).ndim == 1
    finally:
        batcher.close()

"""GPU test through the TOP-LEVEL API: F5TTS(...).infer(ref_file, ref_text, gen_text) -> (wav, sr, spec), i.e.
api.py:98-149 -> infer_process (utils_infer.py:384-434) -> infer_batch_process / _infer_basic (440-593) -> CFM.sample ->
Vocos.decode, with the reference's own shipped fixture (infer/examples/basic/basic_ref_en.wav + infer/examples/vocab.txt,
copied to tests/golden/) and a synthetic checkpoint written in the released layouts (EMA .safetensors; vocos
config.yaml + pytorch_model.bin).  The expected result is rebuilt here from the reference's formulas (RMS gain, pinyin
tokens, duration heuristic, prompt slicing) on top of the CPU oracle, with the SAME initial noise: `infer` seeds the
global generators (api.py:117-121) and `sample` draws randn on the device (cfm.py:196-201), so the test re-draws it.

Tolerances: spectrogram rel-L2 <= 5e-3 (+ one fp16 rounding: the reference keeps the model in fp16 on CUDA,
utils_infer.py:190-199), waveform rel-L2 <= 2e-2 ("parity unpinned" for Vocos: the oracle restates the vocos package).
"""
import os

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import synthdata as SD  # noqa: E402
from f5_tts_b200 import api, infer  # noqa: E402
from f5_tts_b200.model import list_str_to_idx  # noqa: E402
from oracle import f5_oracle as O  # noqa: E402

DEV = "cuda:0"
REF_TEXT = "Some call me nature, others call me mother nature."  # infer/examples/basic/basic.toml
GEN_SHORT = "I don't really care what you call me."
GEN_LONG = ("I don't really care what you call me. I've been a silent spectator, watching species evolve, empires rise "
            "and fall. But always remember, I am mighty and enduring.")
NFE = 8


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def assets(tmp_path_factory, golden_dir):
    """Checkpoint + vocoder folder in the released on-disk layouts, and the loaded F5TTS object."""
    from safetensors.torch import save_file

    d = tmp_path_factory.mktemp("f5assets")
    cfg = SD.f5tts_base()
    sd = SD.synthetic_state_dict(cfg, seed=1234)
    ema = {"ema_model." + k: v for k, v in sd.items()}
    ema["initted"], ema["step"] = torch.tensor(True), torch.tensor(1)
    ckpt = str(d / "model_1.safetensors")
    save_file(ema, ckpt)
    vcfg = {"feature_extractor": {"class_path": "vocos.feature_extractors.MelSpectrogramFeatures",
                                  "init_args": {"sample_rate": 24000, "n_fft": 1024, "hop_length": 256, "n_mels": 100,
                                                "padding": "center"}},
            "backbone": {"class_path": "vocos.models.VocosBackbone",
                         "init_args": {"input_channels": 100, "dim": 512, "intermediate_dim": 1536, "num_layers": 8}},
            "head": {"class_path": "vocos.heads.ISTFTHead",
                     "init_args": {"dim": 512, "n_fft": 1024, "hop_length": 256, "padding": "center"}}}
    vdir = d / "vocos"
    vdir.mkdir()
    (vdir / "config.yaml").write_text(yaml.safe_dump(vcfg))
    vsd = SD.synthetic_vocos_state_dict()
    full = dict(vsd)
    full["feature_extractor.mel_spec.spectrogram.window"] = torch.hann_window(1024)
    full["feature_extractor.mel_spec.mel_scale.fb"] = O.mel_filterbank()
    torch.save(full, str(vdir / "pytorch_model.bin"))
    tts = api.F5TTS(model="F5TTS_Base", ckpt_file=ckpt, vocab_file=os.path.join(golden_dir, "vocab.txt"),
                    vocoder_local_path=str(vdir), device=DEV)
    return dict(tts=tts, sd=sd, vsd=vsd, cfg=cfg, ckpt=ckpt, ref=os.path.join(golden_dir, "basic_ref_en.wav"))


def expected_chunk(a, gen_text, y0):
    """`_infer_basic` (utils_infer.py:477-520) restated on the CPU oracle for one text chunk."""
    audio, sr = infer._load_wav(a["ref"])
    assert sr == 24000 and audio.shape[0] == 1
    rms = float(torch.sqrt(torch.mean(torch.square(audio))))
    if rms < 0.1:
        audio = audio * 0.1 / rms
    # preprocess_ref_audio_text (utils_infer.py:369-376) appends " " after the final "."; infer_batch_process appends
    # another one because the last character is a single byte (utils_infer.py:474-475)
    ref_text = REF_TEXT + "  "
    tokens = infer.convert_char_to_pinyin([ref_text + gen_text])
    ids = list_str_to_idx(tokens, a["tts"].ema_model.vocab_char_map)
    ref_len = audio.shape[-1] // 256
    duration = ref_len + int(ref_len / len(ref_text.encode("utf-8")) * len(gen_text.encode("utf-8")) / 1.0)
    cond = O.mel_spectrogram(audio).permute(0, 2, 1).half().float()  # fp16 model dtype on CUDA (cfm.py:112)
    res = O.sample(a["sd"], a["cfg"], cond, ids, duration, steps=NFE, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
    gen = res.out.half().float()[:, ref_len:, :].permute(0, 2, 1)
    wav = O.vocos_decode(a["vsd"], gen)
    if rms < 0.1:
        wav = wav * rms / 0.1
    return wav.squeeze().numpy(), gen[0].numpy(), duration, ref_len


def test_f5tts_infer_single_chunk(assets):
    a = assets
    seed = 1234
    wav, sr, spec = a["tts"].infer(a["ref"], REF_TEXT, GEN_SHORT, nfe_step=NFE, seed=seed, show_info=lambda *_: None)
    assert sr == 24000 and a["tts"].seed == seed
    # the noise `sample` drew: first device randn after seed_everything(seed) (api.py:117-121, cfm.py:196-201)
    ref_len = 127987 // 256
    duration = ref_len + int(ref_len / len((REF_TEXT + "  ").encode()) * len(GEN_SHORT.encode()))
    api.seed_everything(seed)
    y0 = torch.randn(duration, 100, device=DEV, dtype=torch.float16).float().cpu()[None]
    w_ref, s_ref, dur, rl = expected_chunk(a, GEN_SHORT, y0)
    assert dur == duration and rl == ref_len
    assert spec.shape == s_ref.shape == (100, duration - ref_len) and wav.shape == w_ref.shape == (256 * (duration - ref_len - 1),)
    rs, rw = rel(spec, s_ref), rel(wav, w_ref)
    print(f"[F5TTS.infer] {duration} frames ({ref_len} prompt), NFE {NFE}: spectrogram rel-L2 {rs:.3e}, waveform rel-L2 {rw:.3e}")
    assert rs <= 5e-3 and rw <= 2e-2


def test_f5tts_infer_multi_chunk_shapes_and_files(assets, tmp_path):
    """Long text -> several chunks (utils_infer.py:400-406), cross-faded (549-585); file outputs of api.py:139-147."""
    a = assets
    audio, sr = infer._load_wav(a["ref"])
    secs = audio.shape[-1] / sr
    ref_text = REF_TEXT + " "
    max_chars = int(len(ref_text.encode()) / secs * (22 - secs) * 1.0)
    chunks = infer.chunk_text(GEN_LONG, max_chars=max_chars)
    assert len(chunks) >= 2
    fw, fs = str(tmp_path / "o.wav"), str(tmp_path / "o.npy")
    wav, sr, spec = a["tts"].infer(a["ref"], REF_TEXT, GEN_LONG, nfe_step=4, seed=7, file_wave=fw, file_spec=fs,
                                   show_info=lambda *_: None)
    ref_len = audio.shape[-1] // 256
    rt2 = ref_text + " "  # infer_batch_process appends a second space (utils_infer.py:474-475)
    durs = [ref_len + int(ref_len / len(rt2.encode()) * len(c.encode()) / (0.3 if len(c.encode()) < 10 else 1.0))
            for c in chunks]
    fade = int(0.15 * 24000)
    assert spec.shape == (100, sum(d - ref_len for d in durs))
    assert len(wav) == sum(256 * (d - ref_len - 1) for d in durs) - fade * (len(chunks) - 1)
    assert np.isfinite(wav).all() and float(np.abs(wav).max()) > 0
    back, sr2 = infer._load_wav(fw)
    assert sr2 == 24000 and back.shape[-1] == len(wav)
    assert np.load(fs).shape == spec.shape


def test_packed_weight_cache_roundtrip(assets, tmp_path, golden_dir):
    """the kernel-layout operands are cached on disk next to the checkpoint identity; the second load
    reads them straight onto the GPU (no re-packing) and samples bit-identically."""
    import glob

    from f5_tts_b200 import weights as Wt

    a = assets
    ckpt = a["ckpt"]
    cls, arch = api.MODEL_ARCH["F5TTS_Base"]
    vocab = os.path.join(golden_dir, "vocab.txt")
    cache = str(tmp_path / "pack")
    m1 = infer.load_model(cls, arch, ckpt, vocab_file=vocab, device=DEV, packed_cache_dir=cache)
    files = glob.glob(os.path.join(cache, "f5pack_*.safetensors"))
    assert len(files) == 1 and os.path.getsize(files[0]) > 600e6  # 0.67 GB of fp16 operands + fp32 table / biases
    calls = {"n": 0}
    orig = Wt.packed_tensors

    def counting(m):
        calls["n"] += 1
        return orig(m)

    Wt.packed_tensors = counting
    try:
        m2 = infer.load_model(cls, arch, ckpt, vocab_file=vocab, device=DEV, packed_cache_dir=cache)
        g = torch.Generator().manual_seed(3)
        cond = torch.randn(1, 40, 100, generator=g).to(DEV)
        text = torch.randint(0, 2545, (1, 30), generator=g).to(DEV)
        y0 = torch.randn(1, 150, 100, generator=g).to(DEV)
        kw = dict(steps=2, cfg_strength=2.0, sway_sampling_coef=-1.0)
        o2, _ = m2.sample(cond.half(), text, 150, **kw, y0=y0)
        assert calls["n"] == 0, "the second load must not re-pack"
    finally:
        Wt.packed_tensors = orig
    o1, _ = m1.sample(cond.half(), text, 150, **kw, y0=y0)
    assert torch.equal(o1, o2)


def test_serving_requests_batched_equal_single(assets):
    """serving.F5TTSRequestProcessor (Triton Python backend mirror, model.py:176-269): a batch of requests with different
    reference lengths and texts goes through ONE sampler call (exact_varlen) and every answer is bit-identical to the
    same request served alone; the single-request answer equals the reference's formulas on top of `CFM.sample`."""
    from f5_tts_b200 import serving

    a = assets
    audio, sr = infer._load_wav(a["ref"])
    wav = audio.numpy()
    texts = ["I don't really care what you call me.", "Hello there.", "But always remember, I am mighty and enduring."]
    reqs = [
        {"reference_wav": wav, "reference_wav_len": np.array([wav.shape[1]], np.int32), "reference_text": REF_TEXT,
         "target_text": texts[0]},
        {"reference_wav": wav, "reference_wav_len": np.array([60000], np.int32),  # a shorter prompt out of the same file
         "reference_text": "Some call me nature,", "target_text": texts[1]},
        {"reference_wav": 0.2 * wav[:, :90000], "reference_text": np.array([[b"Some call me nature, others call me"]], dtype=object),
         "target_text": texts[2]},
    ]
    proc = serving.F5TTSRequestProcessor(a["tts"].ema_model, a["tts"].vocoder, device=DEV, nfe_step=NFE, seed=11)
    batched = proc.execute(reqs)
    singles = [proc.execute([r])[0] for r in reqs]
    for i, (b, s) in enumerate(zip(batched, singles)):
        assert b.shape == s.shape and np.isfinite(b).all() and float(np.abs(b).max()) > 0
        assert np.array_equal(b, s), f"request {i}: batched answer differs from the single-request answer"
    # request 0 alone == the reference's formulas around a plain sampler call with the same seed
    ref_len = 1 + wav.shape[1] // 256
    est = int(ref_len * (1 + len(texts[0].encode()) / len(REF_TEXT.encode())))
    model = a["tts"].ema_model
    rms = float(torch.sqrt(torch.mean(torch.square(audio))))
    w0 = audio * (0.1 / rms) if rms < 0.1 else audio
    cond = model.mel_spec(w0.to(DEV), frames_last=False)
    tok = infer.convert_char_to_pinyin([REF_TEXT + texts[0]])
    out, _ = model.sample(cond=cond, text=tok, duration=est, steps=NFE, cfg_strength=2.0, sway_sampling_coef=-1.0, seed=11)
    mel = out[:, ref_len:est, :].float().permute(0, 2, 1).contiguous()
    w_ref = a["tts"].vocoder.decode(mel).squeeze(0)
    if rms < 0.1:
        w_ref = w_ref * rms / 0.1
    assert len(singles[0]) == 256 * (est - ref_len - 1)
    assert np.array_equal(singles[0], w_ref.cpu().numpy())

"""CPU: the streaming TCP front door (f5_tts_b200/socket_server.py, mirror of the reference's socket_server.py:72-189) over
a real localhost socket with a recording fake sampler / vocoder: wire format (little-endian float32 pieces of at most 2048
samples, END marker), first-package chunk splitting, several requests per connection, state reset on disconnect."""
import socket
import threading

import numpy as np
import torch

from f5_tts_b200 import infer, socket_server as SS


class _Model:
    def __init__(self):
        self.texts = []

    def sample(self, cond, text, duration, steps, cfg_strength, sway_sampling_coef, lens=None, exact_varlen=False):
        self.texts.append("".join(text[0]))
        d = duration if isinstance(duration, int) else int(duration[0])
        return torch.zeros(1, d, 100), None


class _Vocoder:
    def decode(self, mel):
        n = 256 * (mel.shape[-1] - 1)
        return (torch.arange(n, dtype=torch.float32) % 100 / 100.0).view(1, n)


def _processor():
    sr = infer.target_sample_rate
    ref = torch.full((1, 2 * sr), 0.2)
    return SS.TTSStreamingProcessor("F5TTS_Base", "", "", (ref, sr), "a reference sentence, nothing more.", device="cpu",
                                    model_obj=_Model(), vocoder=_Vocoder())


def test_protocol_roundtrip_and_first_package_split():
    proc = _processor()
    assert proc.model.texts and "Warm" in proc.model.texts[0]  # _warm_up ran one streaming request
    assert proc.min_chars < proc.few_chars < proc.max_chars
    ready, stop = threading.Event(), threading.Event()
    th = threading.Thread(target=SS.start_server, args=("127.0.0.1", 0, proc, ready, stop), daemon=True)
    th.start()
    assert ready.wait(5)
    text = ("This is the first sentence of the request, it is fairly long. Then comes a second one, also long enough. "
            "And a third sentence closes the paragraph, so that several chunks exist.")
    with socket.create_connection(("127.0.0.1", ready.port), timeout=20) as c:
        n_before = len(proc.model.texts)
        c.sendall(text.encode("utf-8"))
        audio = SS.receive_stream(c)
        first = proc.model.texts[n_before:]
        # first package: the leading chunk was cut down to min_chars so the first audio leaves early
        plain = infer.chunk_text(text, max_chars=proc.max_chars)
        assert len(first) > len(plain) and proc.first_package is False
        ref_len = (2 * infer.target_sample_rate) // infer.hop_length
        assert audio.dtype == np.float32 and len(audio) > 0 and len(audio) % 256 == 0
        assert np.all((audio >= 0) & (audio < 1.0))
        # a second request on the same connection is not re-split
        n_before = len(proc.model.texts)
        c.sendall(text.encode("utf-8"))
        audio2 = SS.receive_stream(c)
        assert len(proc.model.texts) - n_before == len(plain) and len(audio2) > 0
    # closing the connection re-arms the first-package behaviour
    for _ in range(50):
        if proc.first_package:
            break
        threading.Event().wait(0.05)
    assert proc.first_package is True
    stop.set()
    th.join(2)
    assert ref_len > 0

"""Vocos-compatible vocoder on the B200 path (mirror of ``vocos.Vocos`` as the reference constructs and calls it:
infer/utils_infer.py:118-129 ``Vocos.from_hparams(config)`` + ``load_state_dict`` and ``vocoder.decode(mel)`` at
utils_infer.py:511).  Parameter names follow the ``charactr/vocos-mel-24khz`` ``pytorch_model.bin`` layout
(SURVEY.md §8b): feature_extractor.mel_spec.*, backbone.embed/norm/convnext.N/final_layer_norm, head.out,
head.istft.window.  Only ``decode`` runs on the GPU kernels; there is no PyTorch compute path.
"""
from __future__ import annotations

import ctypes as C
import threading

import torch
from torch import nn

from . import _lib
from .model import _attach


class Vocos(nn.Module):
    def __init__(self, input_channels=100, dim=512, intermediate_dim=1536, num_layers=8, n_fft=1024, hop_length=256,
                 sample_rate=24000, padding="center"):
        super().__init__()
        if (input_channels, dim, intermediate_dim, n_fft, hop_length) != (100, 512, 1536, 1024, 256) or num_layers > 8:
            raise NotImplementedError("kernels are built for the vocos-mel-24khz shape (100/512/1536, n_fft 1024, hop 256)")
        if padding != "center":
            raise NotImplementedError("ISTFT padding='center' only (the shipped vocos-mel-24khz config)")
        self.dim, self.inter, self.layers, self.n_mels = dim, intermediate_dim, num_layers, input_channels
        g = torch.Generator().manual_seed(0)
        r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
        _attach(self, "feature_extractor.mel_spec.spectrogram.window", torch.hann_window(n_fft), buffer=True)
        _attach(self, "feature_extractor.mel_spec.mel_scale.fb", torch.zeros(n_fft // 2 + 1, input_channels), buffer=True)
        _attach(self, "backbone.embed.weight", r(dim, input_channels, 7) / (7 * input_channels) ** 0.5)
        _attach(self, "backbone.embed.bias", torch.zeros(dim))
        _attach(self, "backbone.norm.weight", torch.ones(dim))
        _attach(self, "backbone.norm.bias", torch.zeros(dim))
        for i in range(num_layers):
            p = f"backbone.convnext.{i}."
            _attach(self, p + "dwconv.weight", r(dim, 1, 7) / 7 ** 0.5)
            _attach(self, p + "dwconv.bias", torch.zeros(dim))
            _attach(self, p + "norm.weight", torch.ones(dim))
            _attach(self, p + "norm.bias", torch.zeros(dim))
            _attach(self, p + "pwconv1.weight", r(intermediate_dim, dim) / dim ** 0.5)
            _attach(self, p + "pwconv1.bias", torch.zeros(intermediate_dim))
            _attach(self, p + "pwconv2.weight", r(dim, intermediate_dim) / intermediate_dim ** 0.5)
            _attach(self, p + "pwconv2.bias", torch.zeros(dim))
            _attach(self, p + "gamma", torch.full((dim,), 1.0 / num_layers))
        _attach(self, "backbone.final_layer_norm.weight", torch.ones(dim))
        _attach(self, "backbone.final_layer_norm.bias", torch.zeros(dim))
        _attach(self, "head.out.weight", r(n_fft + 2, dim) / dim ** 0.5)
        _attach(self, "head.out.bias", torch.zeros(n_fft + 2))
        _attach(self, "head.istft.window", torch.hann_window(n_fft), buffer=True)
        self._packed = None
        self._lock = threading.Lock()
        self._tls = threading.local()

    @classmethod
    def from_hparams(cls, config_path: str) -> "Vocos":
        """Reads the vocos ``config.yaml`` (backbone / head init_args) like vocos.Vocos.from_hparams."""
        import yaml

        with open(config_path, "r") as f:
            cfg = yaml.safe_load(f)
        b = cfg["backbone"]["init_args"]
        h = cfg["head"]["init_args"]
        return cls(input_channels=b["input_channels"], dim=b["dim"], intermediate_dim=b["intermediate_dim"],
                   num_layers=b["num_layers"], n_fft=h["n_fft"], hop_length=h["hop_length"],
                   padding=h.get("padding", "same"))

    def _pack(self):
        mods = self.__dict__.get("_param_modules")
        if mods is None:  # the module tree is fixed after __init__; nn.Module.parameters() re-walks it on every call
            mods = [m for m in self.modules() if m._parameters]
            self.__dict__["_param_modules"] = mods
        fp = tuple((p.data_ptr(), p._version) for m in mods for p in m._parameters.values() if p is not None)
        with self._lock:
            if self._packed is not None and self._packed["fp"] == fp:
                return self._packed
            sd = self.state_dict()
            dev = sd["head.out.weight"].device
            if dev.type != "cuda":
                raise _lib.F5LibraryError("the vocoder must live on a CUDA (B200) device; no CPU path exists")
            keep = []

            def H(t):
                t = t.detach().to(torch.float16).contiguous()
                keep.append(t)
                return t.data_ptr()

            def Fp(t):
                t = t.detach().to(torch.float32).contiguous()
                keep.append(t)
                return t.data_ptr()

            W = _lib.VocosWeights()
            ew = sd["backbone.embed.weight"]  # [512, 100, 7] -> im2col [512, tap*100 + c], K padded to 704
            ewp = torch.zeros((self.dim, 704), dtype=torch.float16, device=dev)
            ewp[:, :700] = ew.permute(0, 2, 1).reshape(self.dim, 700).to(torch.float16)
            keep.append(ewp)
            W.embed_w, W.embed_b = ewp.data_ptr(), Fp(sd["backbone.embed.bias"])
            W.norm_w, W.norm_b = Fp(sd["backbone.norm.weight"]), Fp(sd["backbone.norm.bias"])
            for i in range(self.layers):
                p = f"backbone.convnext.{i}."
                W.dw_w[i], W.dw_b[i] = Fp(sd[p + "dwconv.weight"].reshape(self.dim, 7)), Fp(sd[p + "dwconv.bias"])
                W.ln_w[i], W.ln_b[i] = Fp(sd[p + "norm.weight"]), Fp(sd[p + "norm.bias"])
                W.pw1_w[i], W.pw1_b[i] = H(sd[p + "pwconv1.weight"]), Fp(sd[p + "pwconv1.bias"])
                W.pw2_w[i], W.pw2_b[i] = H(sd[p + "pwconv2.weight"]), Fp(sd[p + "pwconv2.bias"])
                W.gamma[i] = Fp(sd[p + "gamma"])
            W.final_w, W.final_b = Fp(sd["backbone.final_layer_norm.weight"]), Fp(sd["backbone.final_layer_norm.bias"])
            W.head_w, W.head_b = H(sd["head.out.weight"]), Fp(sd["head.out.bias"])
            W.dim, W.inter, W.layers, W.n_mels = self.dim, self.inter, self.layers, self.n_mels
            self._packed = {"fp": fp, "W": W, "keep": keep}
            return self._packed

    @torch.no_grad()
    def decode(self, features_input: torch.Tensor, **kwargs) -> torch.Tensor:
        """mel float[b, 100, n] -> waveform float[b, 256 * (n - 1)]"""
        mel = features_input
        if not mel.is_cuda:
            raise _lib.F5LibraryError("Vocos.decode runs on the B200 only: move the mel to a CUDA device")
        pk = self._pack()
        mel = mel.float().contiguous()
        B, Cc, T = mel.shape
        assert Cc == self.n_mels
        L = _lib.lib()
        need = L.f5_vocos_workspace_bytes(B, T)
        ws = getattr(self._tls, "ws", None)
        if ws is None or ws.numel() < need or ws.device != mel.device:
            ws = torch.empty(int(need * 1.05) + 4096, dtype=torch.uint8, device=mel.device)
            self._tls.ws = ws
        wav = torch.empty((B, 256 * (T - 1)), dtype=torch.float32, device=mel.device)
        with torch.cuda.device(mel.device):
            _lib.check(L.f5_vocos_decode(C.byref(pk["W"]), mel.data_ptr(), B, T, ws.data_ptr(), ws.numel(),
                                         wav.data_ptr(), torch.cuda.current_stream(mel.device).cuda_stream),
                       "f5_vocos_decode")
        return wav

    def forward(self, mel: torch.Tensor) -> torch.Tensor:  # CFM.sample(vocoder=...) calls the vocoder directly
        return self.decode(mel)

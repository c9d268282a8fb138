"""TEST INFRASTRUCTURE ONLY — never imported by the product package.

Lets the *unmodified* reference (`/root/reference/src/f5_tts`) import in this
container, where several of its third-party dependencies are not installed
(SURVEY.md §8c).  Two kinds of stand-ins are registered in ``sys.modules``:

* empty stubs for packages that are only touched at import time on the
  ODE-sampling path (librosa, rjieba, pypinyin, accelerate, ema_pytorch, ...);
* faithful restatements of the four third-party functions that ARE on the hot
  path: ``torchdiffeq.odeint(method="euler")`` (cfm.py:20,218),
  ``x_transformers.RotaryEmbedding / apply_rotary_pos_emb`` (dit.py:18,207,352;
  modules.py:22,499-509) and ``x_transformers.RMSNorm`` (unett.py:19,154).
  Their published algorithms are restated here from package knowledge
  (x_transformers>=1.31.14, torchdiffeq unpinned; pyproject.toml:37,44) and are
  cross-checked against the reference's own in-tree restatements
  (runtime/triton_trtllm/model_repo_f5_tts/f5_tts/1/f5_tts_trtllm.py:230-261,
  360-369; patch/f5tts/modules.py:210-276).  The reference holds no tests for
  them, so these four are "parity unpinned" (DESIGN.md §oracle).

This module is used by ``oracle/make_golden.py`` (run HERE, where
/root/reference exists) and by the CPU tests that re-check the oracle against
the live reference when it is present.  Nothing in `-m gpu` tests, smoke() or
bench.py touches it.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

REFERENCE_SRC = os.environ.get("F5_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "f5_tts", "model"))


# ---------------------------------------------------------------------------
# torchdiffeq.odeint, fixed-grid explicit Euler on the caller's grid
# ---------------------------------------------------------------------------
def _odeint(func, y0, t, *, method="euler", **_unused):
    if method != "euler":
        raise NotImplementedError("shim restates only method='euler' (the shipped default, cfm.py:41)")
    ys = [y0]
    y = y0
    for k in range(t.shape[0] - 1):
        t0, t1 = t[k], t[k + 1]
        dt = t1 - t0
        y = y + dt * func(t0, y)
        ys.append(y)
    return torch.stack(ys, dim=0)


# ---------------------------------------------------------------------------
# x_transformers rotary helpers + RMSNorm
# ---------------------------------------------------------------------------
class _RotaryEmbedding(nn.Module):
    def __init__(self, dim, use_xpos=False, scale_base=512, interpolation_factor=1.0, base=10000,
                 base_rescale_factor=1.0):
        super().__init__()
        base = base * base_rescale_factor ** (dim / (dim - 2))
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)
        self.interpolation_factor = interpolation_factor
        assert not use_xpos, "xpos is never enabled by the reference"
        self.register_buffer("scale", None)

    def forward_from_seq_len(self, seq_len):
        return self.forward(torch.arange(seq_len, device=self.inv_freq.device))

    def forward(self, t):
        if t.ndim == 1:
            t = t[None, :]
        freqs = torch.einsum("bi,j->bij", t.to(self.inv_freq.dtype), self.inv_freq.float())
        freqs = freqs / self.interpolation_factor
        freqs = torch.stack((freqs, freqs), dim=-1).flatten(-2)  # interleaved duplicate
        return freqs, 1.0


def _rotate_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def _apply_rotary_pos_emb(t, freqs, scale=1):
    rot_dim, seq_len, orig_dtype = freqs.shape[-1], t.shape[-2], t.dtype
    freqs = freqs[:, -seq_len:, :]
    if t.ndim == 4 and freqs.ndim == 3:
        freqs = freqs[:, None]
    t_rot, t_pass = t[..., :rot_dim].float(), t[..., rot_dim:]
    freqs = freqs.float()
    t_rot = (t_rot * freqs.cos() * scale) + (_rotate_half(t_rot) * freqs.sin() * scale)
    return torch.cat((t_rot.to(orig_dtype), t_pass), dim=-1).to(orig_dtype)


class _XRMSNorm(nn.Module):
    def __init__(self, dim, unit_offset=False):
        super().__init__()
        self.unit_offset = unit_offset
        self.scale = dim ** 0.5
        self.g = nn.Parameter(torch.zeros(dim))
        nn.init.constant_(self.g, 1.0 - float(unit_offset))

    def forward(self, x):
        gamma = self.g + float(self.unit_offset)
        return F.normalize(x, dim=-1) * self.scale * gamma


# ---------------------------------------------------------------------------
# registration
# ---------------------------------------------------------------------------
def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []  # behave like a package so "import a.b" works
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


class _Anything:
    """Placeholder for symbols that are imported but never called on this path."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party symbol called on the oracle path")

    def __getattr__(self, item):
        return _Anything()


def install() -> None:
    """Register the stand-ins and put the reference on sys.path (idempotent)."""
    if "torchdiffeq" not in sys.modules:
        _stub("torchdiffeq", odeint=_odeint)
    if "x_transformers" not in sys.modules:
        xt = _stub("x_transformers", RMSNorm=_XRMSNorm)
        xtx = _stub("x_transformers.x_transformers", RotaryEmbedding=_RotaryEmbedding,
                    apply_rotary_pos_emb=_apply_rotary_pos_emb, RMSNorm=_XRMSNorm)
        xt.x_transformers = xtx
    if "librosa" not in sys.modules:
        _stub("librosa")
        _stub("librosa.filters", mel=_Anything())
    for name, attrs in (
        ("rjieba", {}),
        ("pypinyin", {"Style": _Anything(), "lazy_pinyin": _Anything()}),
        ("ema_pytorch", {"EMA": _Anything}),
        ("accelerate", {"Accelerator": _Anything}),
        ("accelerate.utils", {"DistributedDataParallelKwargs": _Anything}),
    ):
        if name not in sys.modules:
            _stub(name, **attrs)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)


def import_reference():
    """Returns (cfm_module, dit_module, unett_module, modules_module, utils_module)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SRC}")
    install()
    import importlib

    # f5_tts/model/__init__.py pulls in the trainer; import the leaf modules it needs.
    cfm = importlib.import_module("f5_tts.model.cfm")
    dit = importlib.import_module("f5_tts.model.backbones.dit")
    unett = importlib.import_module("f5_tts.model.backbones.unett")
    modules = importlib.import_module("f5_tts.model.modules")
    utils = importlib.import_module("f5_tts.model.utils")
    return cfm, dit, unett, modules, utils

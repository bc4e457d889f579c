"""Drop-ins for the reference's HiFi-GAN/NSF generator and its vocoder wrapper.

Mirrors (paths relative to /root/reference/train_bisinger/):
  * ``HifiGanGenerator(h)`` / ``.forward(x[B,80,T], f0[B,T]|None) -> [B,1,T*hop]`` / ``.remove_weight_norm()``
                                                   modules/hifigan/hifigan.py:104-182
  * ``HifiGAN.spec2wav(mel[T,80], f0=...) -> np.ndarray``   vocoders/hifigan.py:36-69 (registry: vocoders/base_vocoder.py:6-20)
Parameter names/shapes are the reference's (``conv_pre``, ``ups.i``, ``noise_convs.i``, ``resblocks.r.convs1/2.m``,
``conv_post``, ``m_source.l_linear``; weight-normed ``weight_g``/``weight_v`` until ``remove_weight_norm()``), so a
reference checkpoint loads with ``strict=True``.  The forward pass is ``bsg_hifigan_forward`` (CUDA); no fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
from torch import nn
from torch.nn.utils import remove_weight_norm, weight_norm

from . import _lib


def _get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


class _ResBlock1Params(nn.Module):
    """Parameter container with the names of ResBlock1 (modules/hifigan/hifigan.py:30-52)."""

    def __init__(self, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=_get_padding(kernel_size, d)))
            for d in dilation])
        self.convs2 = nn.ModuleList([
            weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=_get_padding(kernel_size, 1)))
            for _ in dilation])

    def remove_weight_norm(self):
        for l in list(self.convs1) + list(self.convs2):
            remove_weight_norm(l)


class _SourceParams(nn.Module):
    """SourceModuleHnNSF's only parameters: l_linear (models/source.py:382-383)."""

    def __init__(self, harmonic_num):
        super().__init__()
        self.l_linear = nn.Linear(harmonic_num + 1, 1)


def _conv_weight(m: nn.Module) -> torch.Tensor:
    """Effective weight of a (possibly still weight-normed) conv: w = g * v / ||v|| over all dims but 0
    (torch.nn.utils.weight_norm, dim=0)."""
    if hasattr(m, "weight_g"):
        v, g = m.weight_v.detach(), m.weight_g.detach()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
        return v * (g / norm)
    return m.weight.detach()


class B200HifiGanGenerator(nn.Module):
    def __init__(self, h, c_out=1):
        super().__init__()
        if h.get("resblock", "1") != "1":
            raise NotImplementedError("only ResBlock1 generators are built (BiSinger's vocoder)")
        if c_out != 1:
            raise NotImplementedError("c_out must be 1")
        self.h = h
        self.num_kernels = len(h["resblock_kernel_sizes"])
        self.num_upsamples = len(h["upsample_rates"])
        self.use_pitch_embed = bool(h.get("use_pitch_embed"))
        C0 = h["upsample_initial_channel"]
        if self.use_pitch_embed:
            self.harmonic_num = 8
            self.m_source = _SourceParams(self.harmonic_num)
            self.noise_convs = nn.ModuleList()
        self.conv_pre = weight_norm(nn.Conv1d(80, C0, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            c_cur = C0 // (2 ** (i + 1))
            self.ups.append(weight_norm(nn.ConvTranspose1d(c_cur * 2, c_cur, k, u, padding=(k - u) // 2)))
            if self.use_pitch_embed:
                if i + 1 < len(h["upsample_rates"]):
                    s = int(np.prod(h["upsample_rates"][i + 1:]))
                    self.noise_convs.append(nn.Conv1d(1, c_cur, kernel_size=s * 2, stride=s, padding=s // 2))
                else:
                    self.noise_convs.append(nn.Conv1d(1, c_cur, kernel_size=1))
        self.resblocks = nn.ModuleList()
        ch = C0
        for i in range(len(self.ups)):
            ch = C0 // (2 ** (i + 1))
            for k, d in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
                self.resblocks.append(_ResBlock1Params(ch, k, d))
        self.conv_post = weight_norm(nn.Conv1d(ch, c_out, 7, 1, padding=3))
        self.hop = int(np.prod(h["upsample_rates"]))
        self._plan = None

    def remove_weight_norm(self):
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)
        self._plan = None

    def load_folded_state_dict(self, sd, strict=True):
        """Load a state dict that was saved AFTER remove_weight_norm() (plain ``weight`` keys)."""
        if hasattr(self.conv_pre, "weight_g"):
            self.remove_weight_norm()
        out = self.load_state_dict(sd, strict=strict)
        self._plan = None
        return out

    # weight blob in the order include/bisinger_b200.h documents
    def flat_weights(self) -> torch.Tensor:
        parts = []
        f = lambda t: parts.append(t.detach().to("cpu", torch.float32).reshape(-1))
        if self.use_pitch_embed:
            f(self.m_source.l_linear.weight); f(self.m_source.l_linear.bias)
        f(_conv_weight(self.conv_pre)); f(self.conv_pre.bias)
        for u in self.ups:
            f(_conv_weight(u)); f(u.bias)
        if self.use_pitch_embed:
            for n in self.noise_convs:
                f(n.weight); f(n.bias)
        for rb in self.resblocks:
            for c in rb.convs1:
                f(_conv_weight(c)); f(c.bias)
            for c in rb.convs2:
                f(_conv_weight(c)); f(c.bias)
        f(_conv_weight(self.conv_post)); f(self.conv_post.bias)
        return torch.cat(parts).contiguous()

    def build_plan(self, device=None) -> "HifiganPlan":
        self._plan = HifiganPlan(self, device)
        return self._plan

    @property
    def plan(self) -> "HifiganPlan":
        return self._plan if self._plan is not None else self.build_plan()

    # a plan holds packed copies of the weights on the device: anything that changes the parameters drops it
    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x, f0=None, rand_ini=None, src_noise=None, seed: Optional[int] = None):
        """x [B,80,T], f0 [B,T] or None -> [B,1,T*hop] (hifigan.py:144-173).  ``rand_ini`` [B,9] / ``src_noise`` [B,L,9]
        inject the tensors the reference draws from the global RNG (source.py:54,133); otherwise both are drawn on the device
        from ``seed`` (None: a fresh seed from torch's global generator per call)."""
        return self.plan.forward(x, f0, rand_ini, src_noise, seed)[:, None, :]


class HifiganPlan:
    def __init__(self, gen: B200HifiGanGenerator, device=None):
        L = _lib.lib()
        if device is None:
            device = next(gen.parameters()).device
            if device.type != "cuda":
                device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        h = gen.h
        cfg = _lib.HifiganConfig()
        cfg.num_mels = 80
        cfg.upsample_initial_channel = h["upsample_initial_channel"]
        cfg.num_upsamples = len(h["upsample_rates"])
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            cfg.upsample_rates[i] = int(u)
            cfg.upsample_kernel_sizes[i] = int(k)
        cfg.num_kernels = len(h["resblock_kernel_sizes"])
        nd = {len(d) for d in h["resblock_dilation_sizes"]}
        if len(nd) != 1:
            raise RuntimeError("all resblock_dilation_sizes entries must have the same length")
        cfg.num_dilations = nd.pop()
        for j, k in enumerate(h["resblock_kernel_sizes"]):
            cfg.resblock_kernel_sizes[j] = int(k)
            for m, d in enumerate(h["resblock_dilation_sizes"][j]):
                cfg.resblock_dilation_sizes[j][m] = int(d)
        cfg.use_pitch_embed = 1 if gen.use_pitch_embed else 0
        cfg.audio_sample_rate = int(h.get("audio_sample_rate", 24000))
        cfg.harmonic_num = 8
        cfg.precision = _lib.BSG_PRECISION_BF16
        w = gen.flat_weights()
        hnd = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(L.bsg_hifigan_plan_create(C.byref(cfg), _lib.fptr(w), w.numel(), idx, C.byref(hnd)))
        self._h = hnd
        self.hop = gen.hop

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().bsg_hifigan_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def _prep(self, t):
        return None if t is None else t.to(self.device, torch.float32).contiguous()

    def forward(self, mel, f0=None, rand_ini=None, src_noise=None, seed: Optional[int] = None):
        mel, f0, rand_ini, src_noise = self._prep(mel), self._prep(f0), self._prep(rand_ini), self._prep(src_noise)
        B, M, T = mel.shape
        wav = torch.empty((B, T * self.hop), device=self.device, dtype=torch.float32)
        _lib.check(_lib.lib().bsg_hifigan_forward(self._h, _lib.dev_ptr(mel), _lib.dev_ptr(f0), _lib.dev_ptr(rand_ini),
                                                  _lib.dev_ptr(src_noise), C.c_ulonglong(_lib.resolve_seed(seed)), B, T,
                                                  _lib.dev_ptr(wav), _lib.current_stream_ptr(self.device)))
        return wav

    def source(self, f0, rand_ini=None, src_noise=None, seed: Optional[int] = None):
        f0, rand_ini, src_noise = self._prep(f0), self._prep(rand_ini), self._prep(src_noise)
        B, T = f0.shape
        har = torch.empty((B, T * self.hop), device=self.device, dtype=torch.float32)
        _lib.check(_lib.lib().bsg_hifigan_source(self._h, _lib.dev_ptr(f0), _lib.dev_ptr(rand_ini), _lib.dev_ptr(src_noise),
                                                 C.c_ulonglong(_lib.resolve_seed(seed)), B, T, _lib.dev_ptr(har),
                                                 _lib.current_stream_ptr(self.device)))
        return har


def _ref_hparams() -> dict:
    """The reference's global ``utils.hparams.hparams`` when the reference tree is importable, else an empty dict."""
    try:
        from utils.hparams import hparams as ref_hp  # type: ignore
        return ref_hp
    except Exception:
        return {}


def _base_vocoder_cls():
    """``vocoders.base_vocoder.BaseVocoder`` and ``register_vocoder`` of the reference when it is importable (so that the drop-in IS a
    BaseVocoder and sits in the VOCODERS registry, vocoders/base_vocoder.py:3-20); a stand-in with the same two methods otherwise."""
    try:
        from vocoders.base_vocoder import BaseVocoder, register_vocoder  # type: ignore
        return BaseVocoder, register_vocoder
    except Exception:
        class BaseVocoder:  # vocoders/base_vocoder.py:23-40
            def spec2wav(self, mel):
                raise NotImplementedError

            @staticmethod
            def wav2spec(wav_fn):
                raise NotImplementedError

        return BaseVocoder, (lambda cls: cls)


_BaseVocoder, _register_vocoder = _base_vocoder_cls()


def load_vocoder_config(config_path: str) -> dict:
    """``config.yaml`` (with the ``base_config`` inheritance chain of utils/hparams.py:48-66, deep first, later files override) or
    ``config.json`` (vocoders/hifigan.py:19-25)."""
    if config_path.endswith(".json"):
        import json
        with open(config_path) as f:
            return json.load(f)
    import os
    import yaml
    loaded = set()

    def override(old, new):
        for k, v in new.items():
            if isinstance(v, dict) and k in old:
                override(old[k], v)
            else:
                old[k] = v

    def load(fn):
        with open(fn) as f:
            hp = yaml.safe_load(f) or {}
        loaded.add(fn)
        if "base_config" not in hp:
            return hp
        ret = {}
        bases = hp["base_config"] if isinstance(hp["base_config"], list) else [hp["base_config"]]
        for c in bases:
            if c in loaded:
                continue
            if c.startswith("."):
                c = os.path.normpath(f"{os.path.dirname(fn)}/{c}")
            override(ret, load(c))
        override(ret, hp)
        return ret

    return load(config_path)


def find_vocoder_checkpoint(base_dir: str):
    """vocoders/hifigan.py:40-52: ``config.yaml`` + the ``model_ckpt_steps_*.ckpt`` with the highest step count, else
    ``config.json`` + ``generator_v1``.  Returns (config_path, checkpoint_path)."""
    import glob
    import os
    import re as _re
    cfg = os.path.join(base_dir, "config.yaml")
    if os.path.exists(cfg):
        found = glob.glob(os.path.join(base_dir, "model_ckpt_steps_*.ckpt"))
        if not found:
            raise FileNotFoundError(f"no model_ckpt_steps_*.ckpt in {base_dir}")
        return cfg, max(found, key=lambda p: int(_re.findall(r"model_ckpt_steps_(\d+)\.ckpt$", p)[0]))
    cfg = os.path.join(base_dir, "config.json")
    if os.path.exists(cfg):
        return cfg, os.path.join(base_dir, "generator_v1")
    raise FileNotFoundError(f"neither config.yaml nor config.json in vocoder_ckpt = {base_dir}")


def denoise(wav, v=0.0, fft_size=512, hop_size=128, win_size=512):
    """The optional spectral-subtraction post-filter (vocoders/vocoder_utils.py:7-15; hparams['vocoder_denoise_c'] > 0):
    STFT (hann window, centred, constant = zero padding) -> magnitudes reduced by ``v`` and clipped at 0, phases kept -> inverse
    STFT.  The reference calls librosa.stft / librosa.istft; the same transform pair is evaluated here with torch.stft / torch.istft on
    the host (periodic hann of ``win_size`` zero-padded to ``fft_size``, squared-window overlap-add normalisation, the centre padding
    trimmed, output length hop * (frames - 1)).  librosa is not installed in the build container, so this one function is pinned
    by its algebra (identity at v = 0, tests/test_vocoder_registry.py), not against librosa's output."""
    x = torch.as_tensor(np.asarray(wav), dtype=torch.float32)
    win = torch.hann_window(win_size, periodic=True)
    spec = torch.stft(x, n_fft=fft_size, hop_length=hop_size, win_length=win_size, window=win, center=True, pad_mode="constant",
                      return_complex=True)
    mag = torch.clamp(spec.abs() - float(v), min=0.0)
    out = torch.istft(torch.polar(mag, torch.angle(spec)), n_fft=fft_size, hop_length=hop_size, win_length=win_size, window=win,
                      center=True)
    return out.numpy()


@_register_vocoder
class B200HifiGAN(_BaseVocoder):
    """Vocoder-registry drop-in for ``vocoders.hifigan.HifiGAN`` (vocoders/hifigan.py:36-69).

    ``get_vocoder_cls(hparams)()`` (tasks/tts/tts.py:109, usr/diffsinger_task.py:36) instantiates the class WITHOUT arguments: the
    constructor then reads ``hparams['vocoder_ckpt']`` exactly like the reference -- ``config.yaml`` + newest
    ``model_ckpt_steps_*.ckpt`` (``state_dict.model_gen``) or ``config.json`` + ``generator_v1`` (``generator``) -- loads the state dict
    with ``strict=True``, folds weight-norm and builds the device plan.  Select it with the config line
    ``vocoder: bisinger_b200.vocoder.B200HifiGAN`` (``get_vocoder_cls`` imports dotted paths, vocoders/base_vocoder.py:12-20); when the
    reference's ``vocoders`` package is importable the class also IS a ``BaseVocoder`` registered as ``B200HifiGAN`` / ``b200hifigan``.
    ``spec2wav(mel[T,80], f0=...)`` returns the flat numpy waveform, honours ``hparams['use_nsf']`` and ``hparams['vocoder_denoise_c']``.
    Also offers ``spec2wav_batch`` returning ``[B, L]`` (the reference's ``run_vocoder`` flattens across the batch,
    inference/m4singer/base_svs_infer.py:142-151 -- a latent bug for B > 1)."""

    def __init__(self, generator: Optional[B200HifiGanGenerator] = None, config: Optional[dict] = None, use_nsf: Optional[bool] = None,
                 hparams: Optional[dict] = None):
        self.hparams = hp = hparams if hparams is not None else _ref_hparams()
        if generator is None:
            base_dir = hp.get("vocoder_ckpt")
            if not base_dir:
                raise RuntimeError("B200HifiGAN(): hparams['vocoder_ckpt'] is not set (the zero-argument constructor reads the "
                                   "reference's global hparams, vocoders/hifigan.py:38; or pass hparams=...)")
            config_path, ckpt = find_vocoder_checkpoint(base_dir)
            generator, config = self._load_model(config_path, ckpt)
            print("| load B200HifiGAN: ", ckpt)
        self.model = generator
        self.config = config or generator.h
        self.use_nsf = bool(hp.get("use_nsf")) if use_nsf is None else bool(use_nsf)
        self.device = generator.plan.device

    @staticmethod
    def _load_model(config_path: str, checkpoint_path: str):
        """load_model (vocoders/hifigan.py:17-33)."""
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        config = load_vocoder_config(config_path)
        state = ckpt["state_dict"]["model_gen"] if config_path.endswith(".yaml") else ckpt["generator"]
        gen = B200HifiGanGenerator(config)
        gen.load_state_dict(state, strict=True)
        gen.remove_weight_norm()
        return gen.eval().cuda(), config

    @classmethod
    def from_checkpoint(cls, config, checkpoint_path: str, use_nsf: bool = True, hparams: Optional[dict] = None):
        """Explicit form: ``config`` is a dict or a path to config.yaml / config.json; ``state_dict.model_gen`` or ``generator``."""
        if isinstance(config, str):
            config = load_vocoder_config(config)
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        state = ckpt["state_dict"]["model_gen"] if "state_dict" in ckpt else ckpt["generator"]
        gen = B200HifiGanGenerator(config)
        gen.load_state_dict(state, strict=True)
        gen.remove_weight_norm()
        return cls(gen.eval().cuda(), config, use_nsf, hparams=hparams if hparams is not None else {})

    def spec2wav(self, mel, **kwargs):
        """mel [T,80] (numpy) -> wav [T*hop] (numpy); f0=[T] Hz is used when hparams['use_nsf'] (vocoders/hifigan.py:55-69)."""
        c = torch.as_tensor(np.asarray(mel), dtype=torch.float32).unsqueeze(0).transpose(2, 1)
        f0 = kwargs.get("f0")
        if f0 is not None and self.use_nsf:
            f0 = torch.as_tensor(np.asarray(f0), dtype=torch.float32)[None, :]
        else:
            f0 = None
        y = self.model(c, f0, seed=kwargs.get("seed")).view(-1)
        wav_out = y.cpu().numpy()
        hp = self.hparams
        if hp.get("vocoder_denoise_c", 0.0) > 0:
            wav_out = denoise(wav_out, v=hp["vocoder_denoise_c"], fft_size=hp["fft_size"], hop_size=hp["hop_size"],
                              win_size=hp["win_size"])
        return wav_out

    def spec2wav_batch(self, mel_btm, f0_bt=None, seed: Optional[int] = None):
        y = self.model(torch.as_tensor(mel_btm).transpose(2, 1), f0_bt if self.use_nsf else None, seed=seed)
        return y[:, 0]

    @staticmethod
    def wav2spec(wav_fn):
        raise NotImplementedError("wav2spec is data preparation (vocoders/pwg.py:95-139) and out of scope for the hot path")

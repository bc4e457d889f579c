"""Drop-in for the reference's PitchExtractor (mel -> f0 between the sampler and the vocoder).

Mirrors ``PitchExtractor(n_mel_bins=80, conv_layers=2)`` / ``.forward(mel_input[B,T,80]) -> {'pitch_pred' [B,T,2],
'f0_denorm_pred' [B,T]}`` (modules/fastspeech/pe.py:120-150, paths relative to /root/reference/train_bisinger/).  The
parameter and buffer names are the reference's (``mel_prenet.layers.i.{0,2}``, ``mel_prenet.out_proj``,
``mel_encoder.{in_proj,conv.j.conv.conv,conv.j.norm,out_proj}``, ``pitch_predictor.{pos_embed_alpha,conv.i.{1,3},linear,
embed_positions._float_tensor}``), so the ``checkpoints/m4singer_pe`` checkpoint loads with ``strict=True``
(inference/m4singer/bisinger/a-lang-esm-style-ori-shift.py:600-603).  Inference (eval mode) only; the forward pass is
``bsg_pe_forward`` (CUDA), there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch import nn

from . import _lib

# what PitchExtractor reads from the global hparams (configs/tts/fs2.yaml:13-14,23,33; configs/tts/base.yaml:64; usr/configs/base.yaml:2)
DEFAULT_HPARAMS = dict(predictor_hidden=-1, ffn_padding="SAME", predictor_kernel=5, pitch_type="frame", use_uv=True,
                       pitch_norm="log", f0_mean=0.0, f0_std=1.0)


class _ConvNormParams(nn.Module):          # common_layers.py:43-68: ConvNorm.conv
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, padding=(k - 1) // 2)


class _ConvBlockParams(nn.Module):         # pe.py:45-60, norm='gn'
    def __init__(self, c, k):
        super().__init__()
        self.conv = _ConvNormParams(c, c, k)
        self.norm = nn.GroupNorm(c // 16, c)


class _PrenetParams(nn.Module):            # pe.py:8-22
    def __init__(self, in_dim, out_dim, kernel=5, n_layers=3):
        super().__init__()
        layers = []
        for _ in range(n_layers):
            layers.append(nn.Sequential(nn.Conv1d(in_dim, out_dim, kernel, padding=kernel // 2), nn.ReLU(), nn.BatchNorm1d(out_dim)))
            in_dim = out_dim
        self.layers = nn.ModuleList(layers)
        self.out_proj = nn.Linear(out_dim, out_dim)


class _ConvStacksParams(nn.Module):        # pe.py:81-98
    def __init__(self, c, n_layers, kernel_size=5):
        super().__init__()
        self.conv = nn.ModuleList([_ConvBlockParams(c, kernel_size) for _ in range(n_layers)])
        self.in_proj = nn.Linear(c, c)
        self.out_proj = nn.Linear(c, c)


class _EmbedPositionsParams(nn.Module):    # common_layers.py:121: the only registered state of SinusoidalPositionalEmbedding
    def __init__(self):
        super().__init__()
        self.register_buffer("_float_tensor", torch.zeros(1))


class _PitchPredictorParams(nn.Module):    # tts_modules.py:205-222
    def __init__(self, idim, n_chans, n_layers, kernel_size, odim=2):
        super().__init__()
        self.conv = nn.ModuleList()
        for idx in range(n_layers):
            self.conv.append(nn.Sequential(nn.Identity(), nn.Conv1d(idim if idx == 0 else n_chans, n_chans, kernel_size),
                                           nn.ReLU(), nn.LayerNorm(n_chans, eps=1e-12), nn.Identity()))
        self.linear = nn.Linear(n_chans, odim)
        self.embed_positions = _EmbedPositionsParams()
        self.pos_embed_alpha = nn.Parameter(torch.Tensor([1]))


class B200PitchExtractor(nn.Module):
    def __init__(self, n_mel_bins=80, conv_layers=2, hparams: Optional[dict] = None):
        super().__init__()
        hp = dict(DEFAULT_HPARAMS)
        if hparams:
            hp.update({k: hparams[k] for k in DEFAULT_HPARAMS if k in hparams})
        self.hp = hp
        self.n_mel_bins = n_mel_bins
        self.hidden_size = 256
        self.predictor_hidden = hp["predictor_hidden"] if hp["predictor_hidden"] > 0 else self.hidden_size
        self.conv_layers = conv_layers
        self.mel_prenet = _PrenetParams(n_mel_bins, self.hidden_size)
        if conv_layers > 0:
            self.mel_encoder = _ConvStacksParams(self.hidden_size, conv_layers)
        self.pitch_predictor = _PitchPredictorParams(self.hidden_size, self.predictor_hidden, 5, hp["predictor_kernel"])
        self._plan = None

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("B200PitchExtractor is inference-only (BatchNorm / Dropout in eval mode)")
        return super().train(False)

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    @classmethod
    def from_checkpoint(cls, ckpt_base_dir: str, hparams: Optional[dict] = None, n_mel_bins=80, conv_layers=2):
        """``PitchExtractor().to(device); load_ckpt(pe, hparams['pe_ckpt'], 'model', strict=True); pe.eval()``
        (inference/m4singer/bisinger/a-*.py:600-603) in one call."""
        from .io import load_pitch_extractor_checkpoint
        pe = cls(n_mel_bins, conv_layers, hparams).eval()
        load_pitch_extractor_checkpoint(pe, ckpt_base_dir)
        return pe

    # weight blob in the order include/bisinger_b200.h documents
    def flat_weights(self) -> torch.Tensor:
        parts = []
        f = lambda t: parts.append(t.detach().to("cpu", torch.float32).reshape(-1))
        for l in self.mel_prenet.layers:
            conv, bn = l[0], l[2]
            f(conv.weight); f(conv.bias)
            scale = bn.weight.detach().float().cpu() / torch.sqrt(bn.running_var.detach().float().cpu() + bn.eps)
            f(scale); f(bn.bias.detach().float().cpu() - bn.running_mean.detach().float().cpu() * scale)
        f(self.mel_prenet.out_proj.weight); f(self.mel_prenet.out_proj.bias)
        if self.conv_layers > 0:
            enc = self.mel_encoder
            f(enc.in_proj.weight); f(enc.in_proj.bias)
            for b in enc.conv:
                f(b.conv.conv.weight); f(b.conv.conv.bias); f(b.norm.weight); f(b.norm.bias)
            f(enc.out_proj.weight); f(enc.out_proj.bias)
        pp = self.pitch_predictor
        f(pp.pos_embed_alpha)
        half = self.hidden_size // 2
        f(torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1))))   # common_layers.py:130-132
        for l in pp.conv:
            f(l[1].weight); f(l[1].bias); f(l[3].weight); f(l[3].bias)
        f(pp.linear.weight); f(pp.linear.bias)
        return torch.cat(parts).contiguous()

    def build_plan(self, device=None) -> "PitchExtractorPlan":
        self._plan = PitchExtractorPlan(self, device)
        return self._plan

    @property
    def plan(self) -> "PitchExtractorPlan":
        return self._plan if self._plan is not None else self.build_plan()

    @torch.no_grad()
    def forward(self, mel_input=None):
        """mel_input [B,T,80] -> {'pitch_pred': [B,T,2], 'f0_denorm_pred': [B,T]} (pe.py:138-150)."""
        pred, f0 = self.plan.forward(mel_input)
        return {"pitch_pred": pred, "f0_denorm_pred": f0}


class PitchExtractorPlan:
    def __init__(self, pe: B200PitchExtractor, device=None):
        L = _lib.lib()
        if device is None:
            device = next(pe.parameters()).device
            if device.type != "cuda":
                device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        hp = pe.hp
        cfg = _lib.PeConfig()
        cfg.n_mel_bins = pe.n_mel_bins
        cfg.hidden_size = pe.hidden_size
        cfg.prenet_layers = len(pe.mel_prenet.layers)
        cfg.conv_layers = pe.conv_layers
        cfg.predictor_layers = len(pe.pitch_predictor.conv)
        cfg.kernel_size = pe.mel_prenet.layers[0][0].kernel_size[0]
        cfg.predictor_kernel = hp["predictor_kernel"]
        cfg.predictor_hidden = pe.predictor_hidden
        cfg.gn_group_size = 16
        cfg.left_padding = 0 if hp["ffn_padding"] == "SAME" else 1
        cfg.pitch_norm = {"log": 0, "standard": 1}.get(hp["pitch_norm"], 2)
        cfg.use_uv = 1 if (hp["pitch_type"] == "frame" and hp["use_uv"]) else 0
        cfg.f0_mean = float(hp.get("f0_mean", 0.0))
        cfg.f0_std = float(hp.get("f0_std", 1.0))
        w = pe.flat_weights()
        hnd = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(L.bsg_pe_plan_create(C.byref(cfg), _lib.fptr(w), w.numel(), idx, C.byref(hnd)))
        self._h = hnd

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().bsg_pe_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward(self, mel):
        mel = mel.to(self.device, torch.float32).contiguous()
        B, T, M = mel.shape
        pred = torch.empty((B, T, 2), device=self.device, dtype=torch.float32)
        f0 = torch.empty((B, T), device=self.device, dtype=torch.float32)
        _lib.check(_lib.lib().bsg_pe_forward(self._h, _lib.dev_ptr(mel), B, T, _lib.dev_ptr(pred), _lib.dev_ptr(f0),
                                             _lib.current_stream_ptr(self.device)))
        return pred, f0

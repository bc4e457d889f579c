"""Drop-in for the reference's mel-rate FFT decoder and its handoff to the sampler.

Mirrors (paths relative to /root/reference/train_bisinger/):
  * ``FastspeechDecoder(hidden_size=None, num_layers=None, kernel_size=None, num_heads=None)`` = ``FFTBlocks`` with positional
    embedding, ``forward(x[B,T,C], padding_mask=None) -> [B,T,C]``        modules/fastspeech/tts_modules.py:253-310,340-347
  * ``FastSpeech2.run_decoder``: ``decoder -> mel_out -> * tgt_nonpadding``  modules/fastspeech/fs2.py:236-240
Parameter / buffer names are the reference's (``pos_embed_alpha``, ``embed_positions._float_tensor``,
``layers.i.op.{layer_norm1,self_attn.in_proj_weight,self_attn.out_proj.weight,layer_norm2,ffn.ffn_1,ffn.ffn_2}``, ``layer_norm``), so
the ``decoder.*`` part of a FastSpeech2 / FastSpeech2MIDI checkpoint loads with ``strict=True``.  The forward pass is
``bsg_fft_forward`` (CUDA: bf16x3 tcgen05 GEMMs + a tcgen05 attention kernel); eval mode only, no fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch import nn

from . import _lib
from .diffusion import _hp

FFT_DEFAULTS = dict(hidden_size=256, dec_layers=4, num_heads=2, dec_ffn_kernel_size=9, ffn_act="gelu", ffn_padding="SAME")


class _Attn(nn.Module):
    """Parameter container of MultiheadAttention(self_attention=True, bias=False) (modules/commons/common_layers.py:199-247)."""

    def __init__(self, c):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * c, c))
        self.out_proj = nn.Linear(c, c, bias=False)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)


class _Ffn(nn.Module):
    """TransformerFFNLayer's parameters (common_layers.py:598-621)."""

    def __init__(self, c, k):
        super().__init__()
        self.ffn_1 = nn.Conv1d(c, 4 * c, k, padding=k // 2)
        self.ffn_2 = nn.Linear(4 * c, c)


class _EncSALayer(nn.Module):
    def __init__(self, c, k):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(c)
        self.self_attn = _Attn(c)
        self.layer_norm2 = nn.LayerNorm(c)
        self.ffn = _Ffn(c, k)


class _Layer(nn.Module):
    """TransformerEncoderLayer: the block lives under ``.op`` (tts_modules.py:18-33)."""

    def __init__(self, c, k):
        super().__init__()
        self.op = _EncSALayer(c, k)


class _Positions(nn.Module):
    """SinusoidalPositionalEmbedding's only state-dict entry (common_layers.py:117-121)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_float_tensor", torch.zeros(1))


class B200FastspeechDecoder(nn.Module):
    def __init__(self, hidden_size=None, num_layers=None, kernel_size=None, num_heads=None, hparams: Optional[dict] = None):
        super().__init__()
        hp = {**FFT_DEFAULTS, **_hp(hparams)}
        self.hidden_size = hp["hidden_size"] if hidden_size is None else hidden_size
        self.num_layers = hp["dec_layers"] if num_layers is None else num_layers
        self.kernel_size = hp["dec_ffn_kernel_size"] if kernel_size is None else kernel_size
        self.num_heads = hp["num_heads"] if num_heads is None else num_heads
        self.ffn_act = hp.get("ffn_act", "gelu")
        if hp.get("ffn_padding", "SAME") != "SAME":
            raise NotImplementedError("only ffn_padding == 'SAME' is built (BiSinger's configuration)")
        if self.ffn_act not in ("gelu", "relu"):
            raise NotImplementedError("ffn_act must be gelu or relu")
        self.use_pos_embed = True
        self.pos_embed_alpha = nn.Parameter(torch.Tensor([1]))
        self.embed_positions = _Positions()
        self.layers = nn.ModuleList([_Layer(self.hidden_size, self.kernel_size) for _ in range(self.num_layers)])
        self.layer_norm = nn.LayerNorm(self.hidden_size)
        self._plan = None
        self._plan_mel = None

    # a plan holds packed copies of the weights on the device: anything that changes the parameters drops it
    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._plan = None
        return super()._apply(fn, *a, **k)

    # weight blob in the order include/bisinger_b200.h documents
    def flat_weights(self, mel_out: Optional[nn.Linear] = None) -> torch.Tensor:
        parts = []
        f = lambda t: parts.append(t.detach().to("cpu", torch.float32).reshape(-1))
        half = self.hidden_size // 2
        f(self.pos_embed_alpha)
        f(torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1))))     # common_layers.py:130-132
        for l in self.layers:
            op = l.op
            f(op.layer_norm1.weight); f(op.layer_norm1.bias)
            f(op.self_attn.in_proj_weight); f(op.self_attn.out_proj.weight)
            f(op.layer_norm2.weight); f(op.layer_norm2.bias)
            f(op.ffn.ffn_1.weight); f(op.ffn.ffn_1.bias); f(op.ffn.ffn_2.weight); f(op.ffn.ffn_2.bias)
        f(self.layer_norm.weight); f(self.layer_norm.bias)
        if mel_out is not None:
            f(mel_out.weight); f(mel_out.bias)
        return torch.cat(parts).contiguous()

    def build_plan(self, device=None, mel_out: Optional[nn.Linear] = None) -> "FftDecoderPlan":
        self._plan = FftDecoderPlan(self, device, mel_out)
        self._plan_mel = mel_out
        return self._plan

    def _plan_for(self, mel_out):
        if self._plan is None or (mel_out is not None and self._plan_mel is not mel_out):
            self.build_plan(mel_out=mel_out if mel_out is not None else self._plan_mel)
        return self._plan

    @torch.no_grad()
    def forward(self, x, padding_mask=None, attn_mask=None, return_hiddens=False):
        """x [B,T,C] -> [B,T,C] (tts_modules.py:286-310).  The padding mask is derived from all-zero frames as the reference does when
        none is passed (run_decoder passes none); an explicit mask, an attention mask or return_hiddens are not part of this path."""
        if padding_mask is not None or attn_mask is not None or return_hiddens:
            raise NotImplementedError("B200FastspeechDecoder runs run_decoder's call: forward(x) with the mask derived from x")
        return self._plan_for(None).forward(x)[0]

    @torch.no_grad()
    def run_decoder(self, decoder_inp, tgt_nonpadding, mel_out: nn.Linear):
        """FastSpeech2.run_decoder (fs2.py:236-240) in one device call: decoder -> mel_out -> * tgt_nonpadding ([B,T,1] or [B,T])."""
        tn = tgt_nonpadding
        if tn is not None and tn.dim() == 3:
            tn = tn[:, :, 0]
        return self._plan_for(mel_out).forward(decoder_inp, tn, want_hidden=False, want_mel=True)[1]


class FftDecoderPlan:
    """Owner of one ``bsg_fft_plan`` handle."""

    def __init__(self, dec: B200FastspeechDecoder, device=None, mel_out: Optional[nn.Linear] = None):
        L = _lib.lib()
        if device is None:
            device = next(dec.parameters()).device
            if device.type != "cuda":
                device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.out_dims = 0 if mel_out is None else int(mel_out.out_features)
        cfg = _lib.FftConfig(dec.hidden_size, dec.num_layers, dec.num_heads, dec.kernel_size, 0 if dec.ffn_act == "gelu" else 1,
                             1 if dec.use_pos_embed else 0, self.out_dims)
        w = dec.flat_weights(mel_out)
        hnd = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(L.bsg_fft_plan_create(C.byref(cfg), _lib.fptr(w), w.numel(), idx, C.byref(hnd)))
        self._h = hnd
        self.hidden = dec.hidden_size

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().bsg_fft_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def forward(self, x, tgt_nonpad=None, want_hidden=True, want_mel=None):
        x = x.to(self.device, torch.float32).contiguous()
        B, T, Cc = x.shape
        if Cc != self.hidden:
            raise RuntimeError(f"input has {Cc} channels, plan expects {self.hidden}")
        want_mel = self.out_dims > 0 if want_mel is None else want_mel
        hid = torch.empty((B, T, Cc), device=self.device, dtype=torch.float32) if want_hidden else None
        mel = torch.empty((B, T, self.out_dims), device=self.device, dtype=torch.float32) if want_mel else None
        tn = None if tgt_nonpad is None else tgt_nonpad.to(self.device, torch.float32).contiguous()
        _lib.check(_lib.lib().bsg_fft_forward(self._h, _lib.dev_ptr(x), _lib.dev_ptr(tn), B, T, _lib.dev_ptr(hid), _lib.dev_ptr(mel),
                                              _lib.current_stream_ptr(self.device)))
        return hid, mel

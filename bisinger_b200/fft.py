"""Drop-ins for the reference's FFT block stacks on either side of the sampler's inputs: the mel-rate decoder
with its handoff to the sampler, and the phoneme-rate encoder.

Mirrors (paths relative to /root/reference/train_bisinger/):
  * ``FastspeechDecoder(hidden_size=None, num_layers=None, kernel_size=None, num_heads=None)`` = ``FFTBlocks`` with positional
    embedding, ``forward(x[B,T,C], padding_mask=None) -> [B,T,C]``        modules/fastspeech/tts_modules.py:253-310,340-347
  * ``FastSpeech2.run_decoder``: ``decoder -> mel_out -> * tgt_nonpadding``  modules/fastspeech/fs2.py:236-240
  * ``FastspeechEncoder(embed_tokens, hidden_size, num_layers, kernel_size, num_heads)`` = ``FFTBlocks(use_pos_embed=False)`` behind a
    token embedding and its own positional table, ``forward(txt_tokens[B,T]) -> [B,T,C]``        tts_modules.py:310-346
  * ``FastspeechMIDIEncoder.forward``: the same block stack behind the MIDI / slur / ESM embeddings, which stay the reference's
    (SURVEY.md section 2 row 8): ``forward_blocks(x, padding_mask)`` is the call it makes        modules/diffsinger_midi/fs2.py:44-65
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

FFT_DEFAULTS = dict(hidden_size=256, dec_layers=4, enc_layers=4, num_heads=2, dec_ffn_kernel_size=9, enc_ffn_kernel_size=9, ffn_act="gelu",
                    ffn_padding="SAME", use_pos_embed=True)


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


class B200FFTBlocks(nn.Module):
    """FFTBlocks (tts_modules.py:253-310) with the reference's parameter names; the forward pass is the device plan."""

    def __init__(self, hidden_size, num_layers, ffn_kernel_size=9, num_heads=2, use_pos_embed=True, ffn_act="gelu", ffn_padding="SAME"):
        super().__init__()
        self.hidden_size, self.num_layers, self.kernel_size, self.num_heads = hidden_size, num_layers, ffn_kernel_size, num_heads
        self.ffn_act = ffn_act
        if ffn_padding != "SAME":
            raise NotImplementedError("only ffn_padding == 'SAME' is built (BiSinger's configuration)")
        if self.ffn_act not in ("gelu", "relu"):
            raise NotImplementedError("ffn_act must be gelu or relu")
        self.use_pos_embed = use_pos_embed
        if use_pos_embed:
            self.pos_embed_alpha = nn.Parameter(torch.Tensor([1]))
            self.embed_positions = _Positions()
        self.layers = nn.ModuleList([_Layer(hidden_size, ffn_kernel_size) for _ in range(num_layers)])
        self.layer_norm = nn.LayerNorm(hidden_size)
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
        f(self.pos_embed_alpha if self.use_pos_embed else torch.ones(1))
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

    @classmethod
    def from_reference(cls, ref: nn.Module) -> "B200FFTBlocks":
        """Device twin of a live reference ``FFTBlocks`` (or FastspeechEncoder / FastspeechMIDIEncoder / FastspeechDecoder: they derive
        from it): geometry read from the module, ``layers.*`` / ``layer_norm.*`` / ``pos_embed_alpha`` copied by name.  The reference
        module keeps its embeddings; route its block stack here with ``twin.forward_blocks(x, padding_mask)``."""
        op = ref.layers[0].op
        if not isinstance(op.ffn.ffn_1, nn.Conv1d):
            raise NotImplementedError("ffn_padding == 'LEFT' is not BiSinger's configuration and is not built")
        k = op.ffn.ffn_1.kernel_size[0]
        use_pos = bool(getattr(ref, "use_pos_embed", False)) and isinstance(getattr(ref, "pos_embed_alpha", None), torch.Tensor)
        if getattr(ref, "layer_norm", None) is None or not isinstance(ref.layer_norm, nn.LayerNorm):
            raise NotImplementedError("FFTBlocks(use_last_norm=False) / norm='bn' are not BiSinger's configuration and are not built")
        self = cls(ref.hidden_size, len(ref.layers), k, op.num_heads, use_pos_embed=use_pos,
                   ffn_act=getattr(op.ffn, "act", "gelu"))
        own = set(self.state_dict().keys())
        self.load_state_dict({n: v for n, v in ref.state_dict().items() if n in own}, strict=True)
        dev = next(ref.parameters()).device
        return self.to(dev).eval() if dev.type == "cuda" else self.eval()

    @torch.no_grad()
    def forward_blocks(self, x, padding_mask=None):
        """FFTBlocks.forward(x, padding_mask) (tts_modules.py:286-310): x [B,T,C] -> [B,T,C]; padding_mask [B,T] bool (True = padding)
        or None (all-zero frames are padding, :291).  attn_mask / return_hiddens are not part of this path."""
        return self._plan_for(None).forward(x, padding_mask=padding_mask)[0]

    def forward(self, x, padding_mask=None, attn_mask=None, return_hiddens=False):
        if attn_mask is not None or return_hiddens:
            raise NotImplementedError("the device path runs FFTBlocks.forward(x, padding_mask); attn_mask / return_hiddens are not built")
        return self.forward_blocks(x, padding_mask)


class B200FastspeechDecoder(B200FFTBlocks):
    def __init__(self, hidden_size=None, num_layers=None, kernel_size=None, num_heads=None, hparams: Optional[dict] = None):
        hp = {**FFT_DEFAULTS, **_hp(hparams)}
        super().__init__(hp["hidden_size"] if hidden_size is None else hidden_size,
                         hp["dec_layers"] if num_layers is None else num_layers,
                         hp["dec_ffn_kernel_size"] if kernel_size is None else kernel_size,
                         hp["num_heads"] if num_heads is None else num_heads,
                         use_pos_embed=True, ffn_act=hp.get("ffn_act", "gelu"), ffn_padding=hp.get("ffn_padding", "SAME"))

    @torch.no_grad()
    def run_decoder(self, decoder_inp, tgt_nonpadding, mel_out: nn.Linear):
        """FastSpeech2.run_decoder (fs2.py:236-240) in one device call: decoder -> mel_out -> * tgt_nonpadding ([B,T,1] or [B,T])."""
        tn = tgt_nonpadding
        if tn is not None and tn.dim() == 3:
            tn = tn[:, :, 0]
        return self._plan_for(mel_out).forward(decoder_inp, tn, want_hidden=False, want_mel=True)[1]


class B200FastspeechEncoder(B200FFTBlocks):
    """FastspeechEncoder (tts_modules.py:310-346).  State-dict names: ``embed_tokens.weight``, ``embed_positions._float_tensor``,
    ``layers.*``, ``layer_norm.*`` (no ``pos_embed_alpha``: the block stack is built with use_pos_embed=False, :316-317).  The token
    embedding and the sinusoidal table are two gathers at phoneme rate and run as torch ops on the plan's device; the four FFT blocks
    run on the device plan.  A FastspeechMIDIEncoder keeps its own ``forward_embedding`` (MIDI / slur / ESM terms) and calls
    ``forward_blocks(x, txt_tokens.eq(0))``."""

    def __init__(self, embed_tokens: nn.Embedding, hidden_size=None, num_layers=None, kernel_size=None, num_heads=2,
                 hparams: Optional[dict] = None):
        hp = {**FFT_DEFAULTS, **_hp(hparams)}
        super().__init__(hp["hidden_size"] if hidden_size is None else hidden_size,
                         hp["dec_layers"] if num_layers is None else num_layers,          # sic: the reference's default (:314)
                         hp["enc_ffn_kernel_size"] if kernel_size is None else kernel_size, num_heads,
                         use_pos_embed=False, ffn_act=hp.get("ffn_act", "gelu"), ffn_padding=hp.get("ffn_padding", "SAME"))
        self.embed_tokens = embed_tokens
        self.embed_scale = math.sqrt(self.hidden_size)
        self.padding_idx = 0
        self.embed_positions = _Positions()
        self.hp_use_pos_embed = bool(hp.get("use_pos_embed", True))     # hparams['use_pos_embed'] (:341), not the FFTBlocks flag
        if hp.get("rel_pos"):
            raise NotImplementedError("rel_pos (RelPositionalEncoding) is not BiSinger's configuration and is not built")

    def _positions(self, txt_tokens):
        # SinusoidalPositionalEmbedding.forward(txt_tokens) (common_layers.py:146-170): positions count the non-padding tokens
        # (utils/__init__.py:146-158 make_positions), the table row of position p is [sin(p f) | cos(p f)], row 0 (padding) zero
        nonpad = txt_tokens.ne(self.padding_idx)
        pos = torch.cumsum(nonpad.int(), dim=1) * nonpad.int() + self.padding_idx
        half = self.hidden_size // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float, device=txt_tokens.device) * -(math.log(10000) / (half - 1)))
        ang = pos.float()[:, :, None] * freq[None, None, :]
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1) * nonpad[:, :, None]

    @torch.no_grad()
    def forward_embedding(self, txt_tokens):
        x = self.embed_scale * self.embed_tokens(txt_tokens)
        if self.hp_use_pos_embed:
            x = x + self._positions(txt_tokens)
        return x

    @torch.no_grad()
    def forward(self, txt_tokens):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("bisinger_b200: the encoder's parameters must live on a CUDA device (no CPU path exists)")
        txt_tokens = txt_tokens.to(dev)
        return self.forward_blocks(self.forward_embedding(txt_tokens), txt_tokens.eq(self.padding_idx))


def device_blocks_encoder(base_cls):
    """Subclass factory for the reference's FFTBlocks-derived encoders (``FastspeechMIDIEncoder``, ``FastspeechEncoder``;
    modules/diffsinger_midi/fs2.py:14-65, modules/fastspeech/tts_modules.py:310-346): the returned class IS the reference class --
    constructor, parameters, state-dict keys, ``forward_embedding`` (token / MIDI / slur embeddings, ESM, positional table) and the
    training-mode forward are inherited unchanged -- and in eval mode its block stack, the
    ``super(FastspeechEncoder, self).forward(x, encoder_padding_mask)`` call, runs on the device plan.  The device twin is built from
    the module's own weights on first use and dropped whenever they change.

        FS_ENCODERS["fft"] = lambda esm, hp, emb, d: device_blocks_encoder(FastspeechMIDIEncoder)(
            esm, emb, hp["hidden_size"], hp["enc_layers"], hp["enc_ffn_kernel_size"], num_heads=hp["num_heads"])
    """

    class _DeviceBlocksEncoder(base_cls):
        def _b200_blocks(self) -> B200FFTBlocks:
            twin = self.__dict__.get("_b200")                 # kept out of the module tree: no extra state-dict keys
            if twin is None:
                twin = self.__dict__["_b200"] = B200FFTBlocks.from_reference(self)
            return twin

        def load_state_dict(self, *a, **k):
            self.__dict__.pop("_b200", None)
            return super().load_state_dict(*a, **k)

        def _apply(self, fn, *a, **k):
            self.__dict__.pop("_b200", None)
            return super()._apply(fn, *a, **k)

        def forward(self, txt_tokens, *embeddings):
            if self.training:
                return super().forward(txt_tokens, *embeddings)
            mask = txt_tokens.eq(self.padding_idx).data
            x = self.forward_embedding(txt_tokens, *embeddings)
            return self._b200_blocks().forward_blocks(x, mask)

    _DeviceBlocksEncoder.__name__ = _DeviceBlocksEncoder.__qualname__ = "B200" + base_cls.__name__
    return _DeviceBlocksEncoder


class FftDecoderPlan:
    """Owner of one ``bsg_fft_plan`` handle."""

    def __init__(self, dec: B200FFTBlocks, device=None, mel_out: Optional[nn.Linear] = None):
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

    def forward(self, x, tgt_nonpad=None, want_hidden=True, want_mel=None, padding_mask=None):
        x = x.to(self.device, torch.float32).contiguous()
        B, T, Cc = x.shape
        if Cc != self.hidden:
            raise RuntimeError(f"input has {Cc} channels, plan expects {self.hidden}")
        want_mel = self.out_dims > 0 if want_mel is None else want_mel
        hid = torch.empty((B, T, Cc), device=self.device, dtype=torch.float32) if want_hidden else None
        mel = torch.empty((B, T, self.out_dims), device=self.device, dtype=torch.float32) if want_mel else None
        tn = None if tgt_nonpad is None else tgt_nonpad.to(self.device, torch.float32).contiguous()
        if padding_mask is None:
            _lib.check(_lib.lib().bsg_fft_forward(self._h, _lib.dev_ptr(x), _lib.dev_ptr(tn), B, T, _lib.dev_ptr(hid), _lib.dev_ptr(mel),
                                                  _lib.current_stream_ptr(self.device)))
        else:
            if tuple(padding_mask.shape) != (B, T):
                raise RuntimeError(f"padding_mask must be [B={B},T={T}], got {tuple(padding_mask.shape)}")
            pm = padding_mask.to(self.device).ne(0).to(torch.uint8).contiguous()
            _lib.check(_lib.lib().bsg_fft_forward_masked(self._h, _lib.dev_ptr(x), _lib.dev_ptr(pm), _lib.dev_ptr(tn), B, T, _lib.dev_ptr(hid),
                                                         _lib.dev_ptr(mel), _lib.current_stream_ptr(self.device)))
        return hid, mel

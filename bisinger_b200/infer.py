"""The tail of the reference's ``forward_model`` on the device: mel_out -> PitchExtractor -> f0 -> vocoder -> wav.

Mirrors inference/m4singer/bisinger/a-lang-esm-style-ori-shift.py:626-633 (paths relative to /root/reference/train_bisinger/):

    mel_out = output["mel_out"]                                   # [B, T, 80]
    if hparams.get("pe_enable"): f0_pred = self.pe(mel_out)["f0_denorm_pred"]
    else:                        f0_pred = output["f0_denorm"]
    wav_out = self.run_vocoder(mel_out, f0=f0_pred)               # base_svs_infer.py:142-151

with every tensor staying in HBM between the three stages (the reference's ``run_vocoder`` also stays on the device; its
``spec2wav`` path goes through numpy per utterance, vocoders/hifigan.py:55-69).  All compute is in libbisinger_b200.so.
"""
from __future__ import annotations

from typing import Optional

import torch


@torch.no_grad()
def mel_to_wav(mel_out: torch.Tensor, generator, pe=None, f0: Optional[torch.Tensor] = None, seed: Optional[int] = None, rand_ini=None,
               src_noise=None) -> torch.Tensor:
    """mel_out [B,T,80] (device) -> wav [B, T*hop].  ``pe``: a B200PitchExtractor (hparams['pe_enable']) or None, in which case
    ``f0`` [B,T] is used as given (``output['f0_denorm']``); ``generator``: a B200HifiGanGenerator.  ``rand_ini`` / ``src_noise``
    inject the NSF source's random tensors (parity tests); otherwise they are drawn on the device from ``seed``."""
    if pe is not None:
        f0 = pe(mel_out)["f0_denorm_pred"]
    c = mel_out.transpose(2, 1).contiguous()                      # base_svs_infer.py:143
    return generator(c, f0, rand_ini, src_noise, seed=seed)[:, 0]                      # .view(-1) per utterance at :150


@torch.no_grad()
def synthesize(diffusion, generator, pe, txt_tokens, seed: Optional[int] = None, **model_kwargs) -> torch.Tensor:
    """``forward_model`` (a-lang-esm-style-ori-shift.py:606-633) with the three drop-ins: ``diffusion`` is a
    B200GaussianDiffusion holding the reference's FastSpeech2 conditioner (``fs2``).  ``seed`` keys the sampler's and the NSF
    source's device-side noise (None: fresh per call, drawn from torch's global generator like the reference's noise)."""
    from . import _lib
    s = _lib.resolve_seed(seed)
    out = diffusion(txt_tokens, infer=True, seed=s, **model_kwargs)
    return mel_to_wav(out["mel_out"], generator, pe, out.get("f0_denorm"), s + 1)

"""I/O edges of the hot path: reference checkpoint layouts -> the drop-in modules, and the waveform
writer.  Host-side Python, no device work.

Reference (relative to train_bisinger/):
  utils/__init__.py:179-210   load_ckpt: newest ``model_ckpt_steps_*.ckpt``, ``state_dict`` keys under the prefix ``model.``
  usr/diff/shallow_diffusion_tts.py:75-79,103-126   ``denoise_fn.*`` / ``fs2.*`` sub-modules and the schedule buffers
  utils/audio.py:13-18        save_wav: optional peak normalisation, * 32767, int16 PCM
"""
from __future__ import annotations

import glob
import os
import re
import struct
from typing import Dict, Optional, Tuple

import numpy as np
import torch


def find_checkpoint(ckpt_base_dir: str) -> str:
    """utils/__init__.py:180-188: a file is taken as is; in a directory the checkpoint with the highest step count wins."""
    if os.path.isfile(ckpt_base_dir):
        return ckpt_base_dir
    found = glob.glob(os.path.join(ckpt_base_dir, "model_ckpt_steps_*.ckpt"))
    if not found:
        raise FileNotFoundError(f"| ckpt not found in {ckpt_base_dir}.")
    return max(found, key=lambda p: int(re.findall(r"model_ckpt_steps_(\d+)\.ckpt$", p)[0]))


def strip_prefix(state_dict: Dict[str, torch.Tensor], prefix: str = "model") -> Dict[str, torch.Tensor]:
    """utils/__init__.py:190-191: keep the keys under ``prefix.`` and drop the prefix."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in state_dict.items() if k.startswith(prefix + ".")}


def split_diffusion_state(state_dict: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """A GaussianDiffusion state dict (already stripped of ``model.``) -> (denoiser, conditioner, buffers): ``denoise_fn.*`` feeds
    B200DiffNet.load_state_dict(strict=True), ``fs2.*`` stays with the reference FastSpeech2MIDI, the rest are the schedule
    buffers and spec_min/max that B200GaussianDiffusion registers under the same names (shallow_diffusion_tts.py:103-126)."""
    den = {k[len("denoise_fn."):]: v for k, v in state_dict.items() if k.startswith("denoise_fn.")}
    fs2 = {k[len("fs2."):]: v for k, v in state_dict.items() if k.startswith("fs2.")}
    rest = {k: v for k, v in state_dict.items() if not k.startswith(("denoise_fn.", "fs2."))}
    return den, fs2, rest


def load_diffusion_checkpoint(model, ckpt_base_dir: str, prefix_in_ckpt: str = "model", strict: bool = True):
    """load_ckpt (utils/__init__.py:179-210) for a B200GaussianDiffusion: denoiser weights and schedule buffers from the task
    checkpoint; the conditioner part is returned for the caller's reference ``fs2`` module.  Rebuilds the device plan."""
    path = find_checkpoint(ckpt_base_dir)
    sd = strip_prefix(torch.load(path, map_location="cpu")["state_dict"], prefix_in_ckpt)
    den, fs2, rest = split_diffusion_state(sd)
    model.denoise_fn.load_state_dict(den, strict=strict)
    own = model.state_dict()
    for k, v in rest.items():
        if k in own:
            if own[k].shape != v.shape:
                raise RuntimeError(f"checkpoint buffer {k} has shape {tuple(v.shape)}, model expects {tuple(own[k].shape)}")
            own[k].copy_(v)
        elif strict:
            raise RuntimeError(f"unexpected key {k} in checkpoint")
    if getattr(model, "fs2", None) is not None and fs2:
        model.fs2.load_state_dict(fs2, strict=strict)
    if hasattr(model, "build_plan"):
        model.build_plan()
    return path, fs2


def load_pitch_extractor_checkpoint(pe, ckpt_base_dir: str, prefix_in_ckpt: str = "model", strict: bool = True) -> str:
    """``utils.load_ckpt(self.pe, hparams['pe_ckpt'], 'model', strict=True)`` (inference/m4singer/bisinger/a-*.py:600-603,
    usr/diffsinger_task.py:37-40) for a B200PitchExtractor: the newest ``model_ckpt_steps_*.ckpt`` under ``checkpoints/m4singer_pe``, keys
    under ``model.``; the parameter / buffer names are the reference's, so the load is strict.  The device plan is rebuilt lazily."""
    path = find_checkpoint(ckpt_base_dir)
    sd = strip_prefix(torch.load(path, map_location="cpu")["state_dict"], prefix_in_ckpt)
    pe.load_state_dict(sd, strict=strict)
    return path


def wav_to_int16(wav, norm: bool = False) -> np.ndarray:
    """utils/audio.py:13-17: ``wav / max|wav|`` if norm, ``* 32767``, C-style truncation to int16."""
    wav = np.asarray(wav, dtype=np.float32).reshape(-1).copy()
    if norm:
        wav = wav / np.abs(wav).max()
    wav *= 32767
    return wav.astype(np.int16)


def save_wav(wav, path: str, sr: int, norm: bool = False) -> None:
    """utils/audio.py:13-18 without scipy: mono 16-bit PCM RIFF/WAVE, the container scipy.io.wavfile.write produces."""
    pcm = wav_to_int16(wav, norm)
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, int(sr), int(sr) * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)

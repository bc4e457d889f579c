"""CPU tests of the I/O edges (bisinger_b200/io.py): checkpoint layouts and the wav writer."""
import os

import numpy as np
import torch

import synth
from bisinger_b200 import io as bio


def test_find_checkpoint_picks_highest_step(tmp_path):
    for s in (1000, 160000, 20000):
        torch.save({"state_dict": {}}, tmp_path / f"model_ckpt_steps_{s}.ckpt")
    assert os.path.basename(bio.find_checkpoint(str(tmp_path))) == "model_ckpt_steps_160000.ckpt"
    assert bio.find_checkpoint(str(tmp_path / "model_ckpt_steps_1000.ckpt")).endswith("model_ckpt_steps_1000.ckpt")


def test_task_checkpoint_layout_round_trip():
    """A task checkpoint carries ``model.denoise_fn.*``, ``model.fs2.*`` and the schedule buffers."""
    den = synth.diffnet_state(1234)
    full = {"model.denoise_fn." + k: v for k, v in den.items()}
    full["model.fs2.encoder.embed_tokens.weight"] = torch.zeros(10, 256)
    full["model.betas"] = torch.linspace(1e-4, 0.06, 100)
    full["model.spec_min"] = torch.zeros(1, 1, 80)
    full["other_module.x"] = torch.zeros(1)
    sd = bio.strip_prefix(full, "model")
    assert "other_module.x" not in sd and "betas" in sd
    d, f, rest = bio.split_diffusion_state(sd)
    assert set(d) == set(den) and all(torch.equal(d[k], den[k]) for k in den)
    assert set(f) == {"encoder.embed_tokens.weight"}
    assert set(rest) == {"betas", "spec_min"}
    from bisinger_b200 import B200DiffNet
    B200DiffNet(80).load_state_dict(d, strict=True)   # same names / shapes as the reference DiffNet


def test_save_wav_matches_reference_semantics(tmp_path):
    rng = np.random.default_rng(0)
    wav = (rng.standard_normal(2400) * 0.3).astype(np.float32)
    pcm = bio.wav_to_int16(wav)
    assert pcm.dtype == np.int16 and np.array_equal(pcm, (wav * 32767).astype(np.int16))
    pn = bio.wav_to_int16(wav, norm=True)
    assert abs(int(np.abs(pn).max()) - 32767) <= 1
    assert np.allclose(wav, (rng.standard_normal(0).sum() + wav))   # the input array is not modified in place
    path = tmp_path / "a.wav"
    bio.save_wav(wav, str(path), 24000)
    try:
        from scipy.io import wavfile
    except Exception:
        return
    sr, back = wavfile.read(str(path))
    assert sr == 24000 and np.array_equal(back, pcm)


def test_pitch_extractor_checkpoint_layout(tmp_path):
    """checkpoints/m4singer_pe layout: ``state_dict`` keys under ``model.`` (tasks/tts/pe.py:109 trains ``self.model = PitchExtractor``); the
    drop-in loads it strictly and folds BatchNorm from the loaded running statistics."""
    from bisinger_b200.pitch import B200PitchExtractor
    sd = synth.pe_state(777, 2)
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}, "global_step": 60000}, tmp_path / "model_ckpt_steps_60000.ckpt")
    torch.save({"state_dict": {"model." + k: torch.zeros_like(v) for k, v in sd.items()}}, tmp_path / "model_ckpt_steps_100.ckpt")
    pe = B200PitchExtractor.from_checkpoint(str(tmp_path))
    assert not pe.training
    got = pe.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    w = pe.flat_weights()
    bn = sd["mel_prenet.layers.0.2.weight"] / torch.sqrt(sd["mel_prenet.layers.0.2.running_var"] + 1e-5)
    off = 256 * 80 * 5 + 256
    assert torch.allclose(w[off:off + 256], bn)                       # folded scale follows conv weight + bias in the blob

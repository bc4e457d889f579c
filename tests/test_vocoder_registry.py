"""The vocoder boundary as the reference uses it: ``get_vocoder_cls(hparams)()`` -- a ZERO-argument
constructor that reads ``hparams['vocoder_ckpt']`` -- then ``spec2wav(mel[T,80], f0=...)``.

Reference: vocoders/base_vocoder.py:6-20 (registry, dotted paths), vocoders/hifigan.py:17-69 (load_model, HifiGAN.__init__, spec2wav),
tasks/tts/tts.py:109 and usr/diffsinger_task.py:36 (the call sites), vocoders/vocoder_utils.py:7-15 (denoise).
The reference tree does not travel to the GPU box, so its 8-line ``get_vocoder_cls`` is restated here and the global
``utils.hparams.hparams`` dict it reads is provided as a stub module -- the drop-in imports exactly that name."""
import importlib
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

import svs_oracle as O
import synth


def get_vocoder_cls(hparams, registry=None):
    """vocoders/base_vocoder.py:12-20."""
    registry = registry or {}
    if hparams["vocoder"] in registry:
        return registry[hparams["vocoder"]]
    vocoder_cls = hparams["vocoder"]
    pkg = ".".join(vocoder_cls.split(".")[:-1])
    cls_name = vocoder_cls.split(".")[-1]
    return getattr(importlib.import_module(pkg), cls_name)


def _weight_normed_state(seed):
    """The synthetic generator as a checkpoint stores it: weight_g / weight_v pairs (244 keys); g = 1.7 ||v|| and
    v = w / 1.7, so the folded weight is the synthetic one."""
    from bisinger_b200.vocoder import B200HifiGanGenerator
    folded = synth.hifigan_state(seed)
    gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
    sd = {}
    for k, ref in gen.state_dict().items():
        if k.endswith("weight_v"):
            sd[k] = folded[k[:-2]] / 1.7
        elif k.endswith("weight_g"):
            w = folded[k[:-2]]
            sd[k] = w.reshape(w.shape[0], -1).norm(dim=1).reshape(ref.shape)
        else:
            sd[k] = folded[k]
    assert len(sd) == 244
    return folded, sd


def _write_yaml_ckpt(d, sd_wn):
    import yaml
    os.makedirs(d, exist_ok=True)
    base = dict(synth.HIFIGAN_CONFIG)
    rates = base.pop("upsample_rates")
    with open(os.path.join(d, "base.yaml"), "w") as f:          # the inheritance chain of utils/hparams.py:48-66
        yaml.safe_dump(dict(base, upsample_rates=[4, 4, 4, 2]), f)
    with open(os.path.join(d, "config.yaml"), "w") as f:
        yaml.safe_dump(dict(base_config="./base.yaml", upsample_rates=rates), f)
    torch.save({"state_dict": {"model_gen": sd_wn, "model_disc": {}}}, os.path.join(d, "model_ckpt_steps_1000.ckpt"))
    torch.save({"state_dict": {"model_gen": sd_wn, "model_disc": {}}}, os.path.join(d, "model_ckpt_steps_250000.ckpt"))


def _write_json_ckpt(d, sd_wn):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(synth.HIFIGAN_CONFIG, f)
    torch.save({"generator": sd_wn}, os.path.join(d, "generator_v1"))


@pytest.fixture
def ref_hparams(monkeypatch):
    """A stand-in for the reference's ``utils.hparams`` module holding the global ``hparams`` dict (utils/hparams.py:7)."""
    hp = {}
    if "utils" not in sys.modules:
        pkg = types.ModuleType("utils")
        pkg.__path__ = []
        monkeypatch.setitem(sys.modules, "utils", pkg)
    mod = types.ModuleType("utils.hparams")
    mod.hparams = hp
    monkeypatch.setitem(sys.modules, "utils.hparams", mod)
    return hp


def test_checkpoint_discovery_and_config_chain(tmp_path):
    from bisinger_b200 import vocoder as V
    folded, sd_wn = _weight_normed_state(4321)
    _write_yaml_ckpt(str(tmp_path / "y"), sd_wn)
    _write_json_ckpt(str(tmp_path / "j"), sd_wn)
    cfg, ckpt = V.find_vocoder_checkpoint(str(tmp_path / "y"))
    assert cfg.endswith("config.yaml") and ckpt.endswith("model_ckpt_steps_250000.ckpt")     # highest step, hifigan.py:43-44
    conf = V.load_vocoder_config(cfg)
    assert conf["upsample_rates"] == [8, 4, 2, 2] and conf["upsample_initial_channel"] == 512  # child overrides base
    cfg, ckpt = V.find_vocoder_checkpoint(str(tmp_path / "j"))
    assert cfg.endswith("config.json") and ckpt.endswith("generator_v1")
    assert V.load_vocoder_config(cfg)["resblock_kernel_sizes"] == [3, 7, 11]
    with pytest.raises(FileNotFoundError):
        V.find_vocoder_checkpoint(str(tmp_path / "missing"))
    # the weight-normed checkpoint loads strictly and folds to the synthetic weights
    gen = V.B200HifiGanGenerator(conf)
    gen.load_state_dict(sd_wn, strict=True)
    flat_wn = gen.flat_weights()
    gen.remove_weight_norm()
    assert torch.allclose(flat_wn, gen.flat_weights(), atol=1e-6)
    ref = V.B200HifiGanGenerator(conf)
    ref.load_folded_state_dict(folded, strict=True)
    assert torch.allclose(ref.flat_weights(), gen.flat_weights(), atol=2e-6)


def test_registry_class_surface(ref_hparams):
    from bisinger_b200 import vocoder as V
    cls = get_vocoder_cls({"vocoder": "bisinger_b200.vocoder.B200HifiGAN"})
    assert cls is V.B200HifiGAN
    assert issubclass(cls, V._BaseVocoder) and callable(getattr(cls, "spec2wav")) and callable(getattr(cls, "wav2spec"))
    with pytest.raises(RuntimeError, match="vocoder_ckpt"):          # zero-arg ctor without hparams['vocoder_ckpt']
        cls()


def test_zero_arg_ctor_needs_a_gpu_not_a_fallback(tmp_path, ref_hparams):
    """On a machine without a GPU the zero-argument constructor gets as far as the checkpoint (strict load, weight-norm folding)
    and then fails loudly when the device plan is built: there is no CPU path."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    folded, sd_wn = _weight_normed_state(4321)
    _write_yaml_ckpt(str(tmp_path / "voc"), sd_wn)
    ref_hparams.update(vocoder="bisinger_b200.vocoder.B200HifiGAN", vocoder_ckpt=str(tmp_path / "voc"), use_nsf=True)
    with pytest.raises((RuntimeError, AssertionError)):
        get_vocoder_cls(ref_hparams)()


def test_denoise_post_filter_algebra():
    """vocoders/vocoder_utils.py:7-15.  v = 0 is the identity on the interior (STFT -> ISTFT with a hann window at 75 % overlap);
    a threshold above every magnitude gives silence; the output length is hop * frames-1 like librosa.istft."""
    from bisinger_b200.vocoder import denoise
    rng = np.random.default_rng(0)
    wav = (0.3 * rng.standard_normal(24000)).astype(np.float32)
    out = denoise(wav, v=0.0, fft_size=512, hop_size=128, win_size=512)
    n = (1 + len(wav) // 128 - 1) * 128
    assert out.shape == (n,)
    assert np.abs(out[512:n - 512] - wav[512:n - 512]).max() < 1e-5
    assert np.abs(denoise(wav, v=1e6, fft_size=512, hop_size=128, win_size=512)).max() == 0.0
    mid = denoise(wav, v=0.5, fft_size=512, hop_size=128, win_size=512)
    assert 0.0 < np.abs(mid).mean() < np.abs(out).mean()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["yaml", "json"])
def test_get_vocoder_cls_zero_arg_spec2wav_vs_oracle(tmp_path, ref_hparams, layout):
    """The reference's own sequence: hparams -> get_vocoder_cls(hparams)() -> spec2wav(mel[T,80], f0=f0[T]) -> numpy wav, from a
    weight-normed checkpoint directory, against the fp32 oracle (SNR >= 40 dB).  The NSF source noise is drawn on the device, so the
    comparison uses the unvoiced-free deterministic part: noise_std is 0.003 against a sine amplitude of 0.1."""
    assert torch.cuda.is_available()
    folded, sd_wn = _weight_normed_state(4321)
    d = str(tmp_path / "m4singer_hifigan")
    (_write_yaml_ckpt if layout == "yaml" else _write_json_ckpt)(d, sd_wn)
    ref_hparams.update(vocoder="bisinger_b200.vocoder.B200HifiGAN", vocoder_ckpt=d, use_nsf=True, vocoder_denoise_c=0.0,
                       profile_infer=False)
    vocoder = get_vocoder_cls(ref_hparams)()                                     # tasks/tts/tts.py:109
    T = 300
    vin = synth.vocoder_inputs(8100, 1, T)
    mel = vin["mel"][0].t().contiguous().numpy()                                 # [T, 80] as the task hands it over
    # without f0 (hparams use_nsf False / f0 None => self.model(c)): fully deterministic
    wav0 = vocoder.spec2wav(mel)
    assert isinstance(wav0, np.ndarray) and wav0.shape == (T * 128,) and wav0.dtype == np.float32
    with torch.no_grad():
        ref0 = O.hifigan_forward(folded, synth.HIFIGAN_CONFIG, vin["mel"], None, None, None)
    assert O.snr_db(ref0.reshape(-1), torch.from_numpy(wav0)) >= 40.0
    # with f0: phases and source noise are random (device Philox here, global RNG in the reference) -> same statistics, fresh per call
    f0 = vin["f0"][0].numpy()
    w1 = vocoder.spec2wav(mel, f0=f0)
    w2 = vocoder.spec2wav(mel, f0=f0)
    assert w1.shape == (T * 128,) and np.isfinite(w1).all() and np.abs(w1).max() <= 1.0
    assert not np.array_equal(w1, w2)                                            # fresh noise per call, like the reference
    torch.manual_seed(5); w3 = vocoder.spec2wav(mel, f0=f0)
    torch.manual_seed(5); w4 = vocoder.spec2wav(mel, f0=f0)
    assert np.array_equal(w3, w4)                                                # reproducible under torch.manual_seed
    assert not np.array_equal(w1, wav0)
    # use_nsf off: f0 is ignored (vocoders/hifigan.py:60-64)
    ref_hparams["use_nsf"] = False
    v2 = get_vocoder_cls(ref_hparams)()
    assert np.array_equal(v2.spec2wav(mel, f0=f0), v2.spec2wav(mel))
    # the denoise post-filter is applied when vocoder_denoise_c > 0
    ref_hparams.update(vocoder_denoise_c=0.02, fft_size=512, hop_size=128, win_size=512)
    v3 = get_vocoder_cls(ref_hparams)()
    wd = v3.spec2wav(mel)
    assert wd.shape == (T * 128,) and not np.array_equal(wd, wav0)


@pytest.mark.gpu
def test_null_rand_ini_draws_device_phases():
    """with rand_ini == NULL (production) the harmonics h >= 1 get a random initial phase from the Philox key
    (source.py:54-57) instead of all starting at phase 0."""
    from bisinger_b200.vocoder import B200HifiGanGenerator
    dev = torch.device("cuda", 0)
    gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
    gen.load_folded_state_dict(synth.hifigan_state(4321), strict=True)
    gen.build_plan(dev)
    vin = synth.vocoder_inputs(8200, 2, 64)
    f0 = torch.full_like(vin["f0"], 220.0)
    zeros = torch.zeros_like(vin["src_noise"])
    aligned = gen.plan.source(f0, torch.zeros(2, 9), zeros)
    a = gen.plan.source(f0, None, zeros, seed=1)
    b = gen.plan.source(f0, None, zeros, seed=1)
    c = gen.plan.source(f0, None, zeros, seed=2)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, aligned)
    assert not torch.equal(a[0], a[1])          # per-utterance phases
    # the fundamental keeps phase 0: with only the fundamental's weight non-zero the source does not depend on the seed
    sd = synth.hifigan_state(4321)
    sd["m_source.l_linear.weight"] = torch.tensor([[0.3, 0, 0, 0, 0, 0, 0, 0, 0]])
    g2 = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
    g2.load_folded_state_dict(sd, strict=True)
    g2.build_plan(dev)
    assert torch.equal(g2.plan.source(f0, None, zeros, seed=1), g2.plan.source(f0, None, zeros, seed=2))

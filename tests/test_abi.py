"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/bisinger_b200.h declares; the host
wrappers mirror the reference's parameter names; no CPU fallback exists."""
import os
import re

import pytest
import torch

import synth
from bisinger_b200 import _lib
from bisinger_b200.diffusion import B200DiffNet, B200GaussianDiffusion
from bisinger_b200.vocoder import B200HifiGanGenerator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def test_header_symbols_exported():
    _ensure_built()
    hdr = open(os.path.join(ROOT, "include", "bisinger_b200.h")).read()
    declared = set(re.findall(r"\b(bsg_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.bsg_abi_version() == 1


def test_no_cpu_fallback():
    _ensure_built()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = B200DiffNet(80)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 80, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 256, 8))


def test_diffnet_state_dict_names_match_reference_layout():
    sd = synth.diffnet_state(1)
    net = B200DiffNet(80)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(net.state_dict().keys()) == list(sd.keys())          # registration order == blob order
    assert net.flat_weights().numel() == 15086416


def test_hifigan_state_dict_names():
    gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
    keys = list(gen.state_dict().keys())
    assert "conv_pre.weight_g" in keys and "ups.0.weight_v" in keys and "resblocks.11.convs2.2.weight_g" in keys
    assert len(keys) == 244                                          # SURVEY.md §9.1 (weight-normed)
    w_before = gen.flat_weights()
    gen.remove_weight_norm()
    assert len(gen.state_dict()) == 166
    assert torch.allclose(w_before, gen.flat_weights(), atol=1e-6)   # folding == remove_weight_norm
    gen.load_folded_state_dict(synth.hifigan_state(2), strict=True)
    assert gen.flat_weights().numel() == 13673867


def test_gaussian_diffusion_buffers():
    import svs_oracle as O
    net = B200DiffNet(80)
    gd = B200GaussianDiffusion(None, 80, net, timesteps=100, K_step=100, betas=O.linear_beta_schedule(100, 0.06),
                               spec_min=synth.SPEC_MIN, spec_max=synth.SPEC_MAX)
    ref = O.schedule_buffers(O.linear_beta_schedule(100, 0.06))
    for k, v in ref.items():
        assert torch.equal(getattr(gd, k), v), k
    assert gd.spec_min.shape == (1, 1, 80)
    with pytest.raises(NotImplementedError):
        gd(torch.zeros(1, 4, dtype=torch.long), infer=False)


def test_ctypes_structs_follow_the_header():
    """Field names and order of the config structs in include/bisinger_b200.h == the ctypes mirrors in _lib.py (an ABI drift between
    the header and the binding would otherwise only show up as wrong results on the GPU)."""
    hdr = open(os.path.join(ROOT, "include", "bisinger_b200.h")).read()
    structs = {name: body for body, name in re.findall(r"typedef struct \{([^}]*)\}\s*(\w+);", hdr)}
    for cname, ctype in (("bsg_diffnet_config", _lib.DiffnetConfig), ("bsg_hifigan_config", _lib.HifiganConfig), ("bsg_pe_config", _lib.PeConfig),
                         ("bsg_schedule", _lib.Schedule)):
        body = re.sub(r"/\*.*?\*/", "", structs[cname], flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(int|float|double|unsigned|size_t)\s*\*?\s*", "", decl)
            for part in decl.split(","):
                names.append(re.sub(r"\[.*", "", part.strip()))
        assert names == [f[0] for f in ctype._fields_], (cname, names)


def test_library_sass_uses_tcgen05_tma_tmem():
    """The built library really is a tcgen05 / TMA / TMEM implementation (B200_PROFILING.md's SASS mnemonics): UTCHMMA (tcgen05.mma, also the
    2-CTA form), UTCQMMA (kind::f8f6f4 correction MMAs), UTMALDG (cp.async.bulk.tensor, 2-D / 3-D, multicast), LDTM (tcgen05.ld), UTCBAR
    (tcgen05.commit) and the 256-bit global accesses of the row-per-thread epilogues -- and no legacy HMMA / wgmma path."""
    import shutil
    import subprocess
    _ensure_built()
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTMALDG.2D", "UTMALDG.3D", "UTMALDG.2D.MULTICAST.2CTA", "LDTM.x32", "UTCBAR",
                     "UTCBAR.2CTA.MULTICAST", "STG.E.ENL2.256", "LDG.E.ENL2.256"):
        assert mnemonic in sass, mnemonic
    assert " HMMA." not in sass and "HGMMA" not in sass
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout


def test_parameter_changes_invalidate_device_plans():
    """a plan holds packed copies of the weights on the device; load_state_dict / .to() / a reloaded denoiser must drop it
    (a stale plan would keep computing with the old weights without any error).  Checked on the host: the cached handles are reset."""
    from bisinger_b200.fft import B200FastspeechDecoder
    net = B200DiffNet(80)
    net._standalone_plan = object()
    v0 = getattr(net, "_version", 0)
    net.load_state_dict(synth.diffnet_state(1), strict=True)
    assert net._standalone_plan is None and net._version == v0 + 1
    gd = B200GaussianDiffusion(None, 80, net, timesteps=100, K_step=100, betas=torch.linspace(1e-4, 0.06, 100), spec_min=synth.SPEC_MIN,
                               spec_max=synth.SPEC_MAX)
    gd._plan, gd._plan_version = object(), net._version
    net.load_state_dict(synth.diffnet_state(2), strict=True)          # utils.load_ckpt(model.denoise_fn, ...) behind the sampler's back
    assert gd._plan_version != net._version                            # .plan rebuilds on next use
    gd._plan = object()
    gd.load_state_dict(gd.state_dict())
    assert gd._plan is None
    gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
    gen._plan = object()
    gen.float()                                                        # any _apply (.to / .cuda / .float)
    assert gen._plan is None
    gen._plan = object()
    gen.load_state_dict(gen.state_dict())
    assert gen._plan is None
    dec = B200FastspeechDecoder(hparams=dict(hidden_size=256, dec_layers=1, num_heads=2, dec_ffn_kernel_size=9))
    dec._plan = object()
    dec.load_state_dict(dec.state_dict())
    assert dec._plan is None

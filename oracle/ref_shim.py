"""Import shim for the UNMODIFIED reference modules (TEST INFRASTRUCTURE ONLY).

Only usable where /root/reference exists (the build container).  Nothing that runs on the
GPU box (``-m gpu`` tests, ``smoke()``, ``bench.py``) may import this file; it exists to
generate ``tests/golden/*.npz`` (``oracle/make_golden.py``) and to let ``tests/test_oracle.py``
check the restatement in ``svs_oracle.py`` against the live reference.

Shims (none touches arithmetic):
  * stub modules for librosa / pycwt (imported at module top of utils/pitch_utils.py:4,
    utils/cwt.py:1,3, never called on this path);
  * pre-seeded ``modules.parallel_wavegan.layers`` / ``.models`` packages: the reference's
    ``layers/__init__.py`` has a circular import and pulls scipy.signal.kaiser / tensorflow,
    none of which ``HifiGanGenerator`` uses (modules/hifigan/hifigan.py:5-7);
  * ``hparams`` populated before ``usr.diff.shallow_diffusion_tts`` is imported, because its
    default arguments read hparams at import time (shallow_diffusion_tts.py:44,73).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("BISINGER_REFERENCE", "/root/reference")
TB = os.path.join(REF_ROOT, "train_bisinger")


def available() -> bool:
    return os.path.isdir(TB)


_loaded = {}

# hparams the hot-path modules read at construct/import time
# (values: usr/configs/lang-esm-style-ori-shift/{diff,base}.yaml, overridden to the BASELINE
#  K=100 p_sample configuration as in usr/configs/popcs_ds_beta6.yaml:63-68)
HOTPATH_HPARAMS = dict(
    hidden_size=256, residual_layers=20, residual_channels=256, dilation_cycle_length=4,
    audio_num_mel_bins=80, keep_bins=80, timesteps=100, K_step=100, max_beta=0.06,
    schedule_type="linear", diff_loss_type="l1", use_midi=False, gaussian_start=False,
    pndm_speedup=None,
)


def load(hp_override: dict | None = None):
    """Returns a namespace with the reference classes: DiffNet, GaussianDiffusion module
    (``gd``), HifiGanGenerator, SourceModuleHnNSF and the global ``hparams`` dict."""
    if not available():
        raise RuntimeError(f"reference tree not found at {TB}")
    if "ns" in _loaded:
        ns = _loaded["ns"]
        if hp_override:
            ns.hparams.update(hp_override)
        return ns
    for name in ("librosa", "librosa.filters", "pycwt"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    sys.modules["pycwt"].wavelet = types.SimpleNamespace()
    if TB not in sys.path:
        sys.path.insert(0, TB)

    import importlib
    pkg_layers = types.ModuleType("modules.parallel_wavegan.layers")
    pkg_layers.ConvInUpsampleNetwork = object
    pkg_layers.UpsampleNetwork = object
    pkg_layers.__path__ = []
    importlib.import_module("modules")
    pwg = importlib.import_module("modules.parallel_wavegan") if os.path.exists(
        os.path.join(TB, "modules/parallel_wavegan/__init__.py")) else None
    if pwg is None:
        pwg = types.ModuleType("modules.parallel_wavegan")
        pwg.__path__ = [os.path.join(TB, "modules/parallel_wavegan")]
        sys.modules["modules.parallel_wavegan"] = pwg
    sys.modules["modules.parallel_wavegan.layers"] = pkg_layers
    pkg_models = types.ModuleType("modules.parallel_wavegan.models")
    pkg_models.__path__ = [os.path.join(TB, "modules/parallel_wavegan/models")]
    sys.modules["modules.parallel_wavegan.models"] = pkg_models

    from utils.hparams import hparams  # type: ignore
    hparams.clear()
    hparams.update(HOTPATH_HPARAMS)
    if hp_override:
        hparams.update(hp_override)

    from usr.diff.net import DiffNet  # type: ignore
    import usr.diff.shallow_diffusion_tts as gd  # type: ignore
    from modules.hifigan.hifigan import HifiGanGenerator  # type: ignore
    from modules.parallel_wavegan.models.source import SourceModuleHnNSF, SineGen  # type: ignore

    ns = types.SimpleNamespace(DiffNet=DiffNet, gd=gd, HifiGanGenerator=HifiGanGenerator,
                               SourceModuleHnNSF=SourceModuleHnNSF, SineGen=SineGen, hparams=hparams)
    _loaded["ns"] = ns
    return ns

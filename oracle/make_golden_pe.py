"""Generate tests/golden/pe_golden.npz by executing the UNMODIFIED reference PitchExtractor (TEST INFRASTRUCTURE ONLY).

Run in the build container (needs /root/reference):  python oracle/make_golden_pe.py
modules/fastspeech/pe.py:120-150 (PitchExtractor), :8-42 (Prenet), :45-117 (ConvBlock / ConvStacks),
modules/fastspeech/tts_modules.py:194-237 (PitchPredictor), utils/pitch_utils.py:63-76 (denorm_f0).
Weights and inputs come from oracle/synth.py (pe_state / pe_inputs), so the fixture holds outputs only."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import svs_oracle as O  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
PE_CASES = [dict(seed=71, B=2, T=50, pad_tail=7, conv_layers=2), dict(seed=72, B=1, T=300, pad_tail=0, conv_layers=2),
            dict(seed=73, B=3, T=33, pad_tail=5, conv_layers=0)]


def build_reference_pe(ns, conv_layers):
    ns.hparams.update(O.PE_HPARAMS)
    from modules.fastspeech.pe import PitchExtractor  # type: ignore
    pe = PitchExtractor(conv_layers=conv_layers).eval()
    pe.load_state_dict(synth.pe_state(777, conv_layers), strict=True)
    return pe


def main():
    warnings.filterwarnings("ignore")
    ns = ref_shim.load()
    out = {}
    with torch.no_grad():
        for i, c in enumerate(PE_CASES):
            pe = build_reference_pe(ns, c["conv_layers"])
            mel = synth.pe_inputs(c["seed"], c["B"], c["T"], pad_tail=c["pad_tail"])
            ret = pe(mel)
            out[f"pitch_pred.{i}"] = ret["pitch_pred"].numpy()
            out[f"f0.{i}"] = ret["f0_denorm_pred"].numpy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "pe_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
    for i in range(len(PE_CASES)):
        f0 = out[f"f0.{i}"]
        print(i, "voiced frac", float((f0 > 0).mean()), "f0 range", float(f0[f0 > 0].min()), float(f0.max()))


if __name__ == "__main__":
    main()

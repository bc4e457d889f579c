"""Generate tests/golden/fft_golden.npz and tests/golden/fft_encoder_golden.npz by executing the UNMODIFIED reference FastspeechDecoder (+ mel_out Linear and the
tgt_nonpadding mask of FastSpeech2.run_decoder) on the synthetic state (TEST INFRASTRUCTURE ONLY; build container:
python oracle/make_golden_fft.py).  modules/fastspeech/tts_modules.py:253-347, modules/fastspeech/fs2.py:236-240."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FFT_CASES = [dict(seed=31, B=2, T=50, pad_tail=9), dict(seed=32, B=1, T=131, pad_tail=0), dict(seed=33, B=3, T=160, pad_tail=40)]
ENC_CASES = [dict(seed=41, B=2, T=37, pad_tail=5), dict(seed=42, B=1, T=130, pad_tail=0), dict(seed=43, B=4, T=64, pad_tail=20)]
FFT_HP = dict(dropout=0.1, enc_ffn_kernel_size=9, dec_ffn_kernel_size=9, ffn_padding="SAME", ffn_act="gelu", num_heads=2, dec_layers=4,
              enc_layers=4, use_pos_embed=True, hidden_size=256)


def main():
    import ref_shim
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_shim.load()
    ns.hparams.update(FFT_HP)
    from modules.fastspeech.tts_modules import FastspeechDecoder  # type: ignore
    sd = synth.fft_state(555)
    dec = FastspeechDecoder().eval()
    dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("mel_out.")}, strict=True)
    mel_out = torch.nn.Linear(256, 80)
    mel_out.load_state_dict({"weight": sd["mel_out.weight"], "bias": sd["mel_out.bias"]})
    out = {}
    with torch.no_grad():
        for i, c in enumerate(FFT_CASES):
            x = synth.fft_inputs(c["seed"], c["B"], c["T"], pad_tail=c["pad_tail"])
            tgt = (x.abs().sum(-1) > 0).float()[:, :, None]
            h = dec(x)                                   # fs2.py:238
            out[f"hidden.{i}"] = h.numpy()
            out[f"mel.{i}"] = (mel_out(h) * tgt).numpy() # :239-240
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "fft_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")

    # the phoneme-rate side: FastspeechEncoder(embed_tokens, hidden, enc_layers, enc_ffn_kernel_size, num_heads) as FS_ENCODERS builds it
    # (modules/diffsinger_midi/fs2.py:70-78), forward(txt_tokens) (tts_modules.py:327-346), and its FFT-block stack alone with an explicit
    # padding mask (FFTBlocks.forward(x, padding_mask), :286-310) -- the call FastspeechMIDIEncoder.forward makes after its embeddings
    from modules.commons.common_layers import Embedding  # type: ignore
    from modules.fastspeech.tts_modules import FastspeechEncoder, FFTBlocks  # type: ignore
    esd = synth.fft_encoder_state(777)
    emb = Embedding(esd["embed_tokens.weight"].shape[0], 256, 0)
    enc = FastspeechEncoder(emb, 256, 4, 9, num_heads=2).eval()
    enc.load_state_dict(esd, strict=True)
    eout = {}
    with torch.no_grad():
        for i, c in enumerate(ENC_CASES):
            tok = synth.fft_tokens(c["seed"], c["B"], c["T"], pad_tail=c["pad_tail"])
            eout[f"enc.{i}"] = enc(tok).numpy()
            x = synth.fft_inputs(c["seed"] + 100, c["B"], c["T"])        # no all-zero frames: the mask comes from the tokens alone
            eout[f"blocks.{i}"] = FFTBlocks.forward(enc, x, tok.eq(0)).numpy()
    path = os.path.join(OUT, "fft_encoder_golden.npz")
    np.savez_compressed(path, **eout)
    print(f"wrote {len(eout)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()

"""Times the FFT decoder + mel_out handoff (bsg_fft_forward) at a BASELINE shape with CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200.fft import B200FastspeechDecoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
dev = torch.device("cuda", 0)
sd = synth.fft_state(555)
dec = B200FastspeechDecoder(hparams=dict(hidden_size=256, dec_layers=4, num_heads=2, dec_ffn_kernel_size=9)).eval()
dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("mel_out.")}, strict=True)
mel_out = torch.nn.Linear(256, 80)
mel_out.load_state_dict({"weight": sd["mel_out.weight"], "bias": sd["mel_out.bias"]})
dec, mel_out = dec.to(dev), mel_out.to(dev)
x = synth.fft_inputs(1, B, T, pad_tail=T // 10).to(dev)
tgt = (x.abs().sum(-1) > 0).float()
for _ in range(3):
    dec.run_decoder(x, tgt, mel_out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    dec.run_decoder(x, tgt, mel_out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flops = B * T * 4 * 2 * (256 * 768 + 256 * 256 + 9 * 256 * 1024 + 1024 * 256) + B * 2 * 4 * 4 * T * T * 128 + B * T * 2 * 256 * 80
print(f"FFT decoder + mel_out B={B} T={T}: {ms:.3f} ms per call ({flops / ms / 1e9:.0f} algorithmic TFLOP/s incl. attention {B * 2 * 4 * 4 * T * T * 128 / 1e9:.0f} GFLOP)")

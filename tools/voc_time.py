"""Times the vocoder forward at a BASELINE shape (default cfg3: B=32, T=1875) with CUDA events; prints ms and audio-s/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200.vocoder import B200HifiGanGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
dev = torch.device("cuda", 0)
gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
gen.load_folded_state_dict(synth.hifigan_state(4321), strict=True)
gen.build_plan(dev)
vin = synth.vocoder_inputs(3, B, T)
mel, f0 = vin["mel"].to(dev), vin["f0"].to(dev)
for _ in range(3):
    gen(mel, f0, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 10
for _ in range(n):
    gen(mel, f0, seed=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"vocoder B={B} T={T} {os.environ.get('BSG_VOC_NOISE_V2', '')}: {ms:.2f} ms per call, {B * T * 128 / 24000 / (ms / 1e3):.0f} audio-s/s")

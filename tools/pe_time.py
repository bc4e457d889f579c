"""Times the PitchExtractor plan at a BASELINE shape (default cfg3: B=32, T=1875) with CUDA events; prints per-call ms."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200 import launch_count  # noqa: E402
from bisinger_b200.pitch import B200PitchExtractor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
dev = torch.device("cuda", 0)
pe = B200PitchExtractor().eval()
pe.load_state_dict(synth.pe_state(777, 2), strict=True)
pe.build_plan(dev)
mel = synth.pe_inputs(5, B, T).to(dev)
for _ in range(3):
    pe(mel)
torch.cuda.synchronize()
n0 = launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pe(mel)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
flops = 2.0 * B * T * (256 * 80 * 5 + 256 * 256 * 5 * 2 + 256 * 256 * 2 + 256 * 256 * 5 * 2 + 256 * 256 + 256 * 256 * 5 * 5 + 512)
print(f"pe B={B} T={T}: {ms:.3f} ms per call, {(launch_count() - n0) // 10} launches, {flops / ms / 1e9:.1f} algorithmic TFLOP/s "
      f"(x3 issued, bf16x3), {B * T * 128 / 24000 / (ms / 1e3):.0f} audio-s/s")

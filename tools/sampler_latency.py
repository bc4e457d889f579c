"""Latency of the K=100 sampler (graph replay, device RNG) at small batch shapes, fused layer kernel vs the two-launch-per-layer path
(BSG_NO_FUSE=1) and its variants.  usage: python tools/sampler_latency.py [B T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200 import B200DiffNet, DiffusionPlan  # noqa: E402
from bisinger_b200.diffusion import _schedule_buffers, linear_beta_schedule  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 938
dev = torch.device("cuda", 0)
K = 100
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
plan = DiffusionPlan(net, _schedule_buffers(linear_beta_schedule(K, 0.06)), K, K, synth.SPEC_MIN, synth.SPEC_MAX, device=dev)
inp = synth.kernel_inputs(7, B, T, 1)
cond, fs2 = inp["cond"].to(dev), inp["fs2_mel"].to(dev)
for _ in range(3):
    plan.sample(cond, fs2, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n):
    plan.sample(cond, fs2, seed=i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
env = {k: v for k, v in os.environ.items() if k.startswith("BSG_")}
print(f"sampler K={K} B={B} T={T} {env}: {ms:.2f} ms per call, {B * T * 128 / 24000 / (ms / 1e3):.0f} audio-s/s")

"""What a NEW utterance length costs the sampler (workspaces and graphs are cached per exact (B, T)): wall-clock time of the first
and the second call at ten lengths never seen before (B = 1, ~5 s phrases; more than the 6 cached shapes, so the LRU evicts), graph path
(device RNG) and eager path (BSG_DIFF_GRAPH=0).  Product-side tool: no oracle.   usage: varying_length_latency.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import B200DiffNet, DiffusionPlan, synthetic as synth  # noqa: E402
from bisinger_b200.diffusion import _schedule_buffers, linear_beta_schedule  # noqa: E402

K = 100
dev = torch.device("cuda", 0)
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
plan = DiffusionPlan(net, _schedule_buffers(linear_beta_schedule(K, 0.06)), K, K, synth.SPEC_MIN, synth.SPEC_MAX, device=dev)
g = torch.Generator(device=dev).manual_seed(1)


def call(T):
    cond = torch.randn((1, T, 256), generator=g, device=dev)
    fs2 = -6.0 + 5.5 * torch.rand((1, T, 80), generator=g, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.sample(cond, fs2, seed=3)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


call(800); call(800)          # library warm-up (module load, first launches)
first, second = [], []
for T in range(900, 1000, 10):
    first.append(call(T))
    second.append(call(T))
fmt = lambda v: " ".join(f"{x:6.1f}" for x in v)
print(f"graphs {'on' if os.environ.get('BSG_DIFF_GRAPH', '1') != '0' else 'off'}: first call at a new length [ms] {fmt(first)}")
print(f"           second call at that length  [ms] {fmt(second)}")
print(f"           mean first {sum(first) / len(first):.1f} ms, mean second {sum(second) / len(second):.1f} ms")

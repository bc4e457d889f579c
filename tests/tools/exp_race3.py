"""Race hunting at the cfg4 shape (B=8, T=11250; 352 row tiles over 74 CTA pairs): N sampler runs with identical injected noise, report
where runs differ from the first one (batch row, frame range, 256-row tile index, mel bins).  MEASUREMENT INFRASTRUCTURE.
usage: python tests/tools/exp_race3.py [B T runs K]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_oracle as O  # noqa: E402
import synth  # noqa: E402
from bisinger_b200 import B200DiffNet, DiffusionPlan  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 11250
RUNS = int(sys.argv[3]) if len(sys.argv) > 3 else 6
K = int(sys.argv[4]) if len(sys.argv) > 4 else 100
dev = torch.device("cuda", 0)
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
sched = O.schedule_buffers(O.linear_beta_schedule(K, 0.06))
plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, precision="fp16x2", device=dev)
g = torch.Generator(device=dev)
g.manual_seed(20250)
cond = torch.randn((B, T, 256), generator=g, device=dev)
fs2 = -6.0 + torch.rand((B, T, 80), generator=g, device=dev) * 5.5
sn = torch.randn((B, 1, 80, T), generator=g, device=dev)
zn = torch.randn((K, B, 1, 80, T), generator=g, device=dev)
env = {k: v for k, v in os.environ.items() if k.startswith("BSG_")}
outs = [plan.sample(cond, fs2, sn, zn) for _ in range(RUNS)]
torch.cuda.synchronize()
bad = 0
for i, o in enumerate(outs[1:], 1):
    d = (o != outs[0])
    n = int(d.sum())
    if n:
        bad += 1
        idx = d.nonzero()
        b_, t_ = idx[:, 0], idx[:, 1]
        print(f"run {i}: {n} differing elements, max |diff| {float((o - outs[0]).abs().max()):.3e}; batch rows {sorted(set(b_.tolist()))}; "
              f"frames {int(t_.min())}..{int(t_.max())} (tiles {int(t_.min()) // 256}..{int(t_.max()) // 256}); "
              f"frames mod 128 in [{int((t_ % 128).min())}, {int((t_ % 128).max())}]", flush=True)
print(f"env {env}: B={B} T={T} K={K}: {bad} of {RUNS - 1} repeat runs differ from the first", flush=True)
# graph path too
a = plan.sample(cond, fs2, seed=3)
bad_g = sum(0 if torch.equal(plan.sample(cond, fs2, seed=3), a) else 1 for _ in range(RUNS - 1))
print(f"graph path: {bad_g} of {RUNS - 1} replays differ", flush=True)

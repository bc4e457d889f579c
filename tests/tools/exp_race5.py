"""Reproduction of the rare run-to-run difference of the K=100 sampler at the cfg4 shape (seen once in ~4 processes, always in batch
row 0, spreading from the first frames): like tests/test_gpu_bench_regime.py -- other shapes first, then N runs at (8, 11250) -- and
report which runs differ and from which frame on.  MEASUREMENT INFRASTRUCTURE.  usage: exp_race5.py [runs] [pre-shapes 0/1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_oracle as O  # noqa: E402
import synth  # noqa: E402
from bisinger_b200 import B200DiffNet, DiffusionPlan  # noqa: E402

RUNS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
PRE = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = 100
dev = torch.device("cuda", 0)
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
sched = O.schedule_buffers(O.linear_beta_schedule(K, 0.06))
plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, precision="fp16x2", device=dev)


def inputs(B, T, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return (torch.randn((B, T, 256), generator=g, device=dev), -6.0 + torch.rand((B, T, 80), generator=g, device=dev) * 5.5,
            torch.randn((B, 1, 80, T), generator=g, device=dev), torch.randn((K, B, 1, 80, T), generator=g, device=dev))


if PRE:
    for B, T in ((96, 256), (32, 1875)):
        a = inputs(B, T, 1)
        o = [plan.sample(*a) for _ in range(3)]
        print(f"pre-shape {B}x{T}: repeat runs equal: {all(torch.equal(x, o[0]) for x in o[1:])}", flush=True)
        del a, o
a = inputs(8, 11250, 2)
SLEEP = float(os.environ.get("RACE_SLEEP", "0"))
outs = []
for i in range(RUNS):
    if SLEEP > 0 and i % 2 == 0:
        torch.cuda.synchronize()
        import time
        time.sleep(SLEEP)          # let the GPU fall idle (clocks drop), as it does while the tests run the CPU oracle
    outs.append(plan.sample(*a))
torch.cuda.synchronize()
# majority vote: which runs are the odd ones
groups = []
for i, o in enumerate(outs):
    for grp in groups:
        if torch.equal(o, outs[grp[0]]):
            grp.append(i)
            break
    else:
        groups.append([i])
groups.sort(key=len, reverse=True)
env = {k: v for k, v in os.environ.items() if k.startswith("BSG_")}
print(f"env {env}: groups of bit-identical runs: {groups}", flush=True)
ref = outs[groups[0][0]]
for grp in groups[1:]:
    o = outs[grp[0]]
    idx = (o != ref).nonzero()
    print(f"  runs {grp}: {idx.shape[0]} elements differ, max |diff| {float((o - ref).abs().max()):.3e}, batch rows {sorted(set(idx[:, 0].tolist()))}, "
          f"frames {int(idx[:, 1].min())}..{int(idx[:, 1].max())}", flush=True)

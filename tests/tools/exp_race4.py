"""Race localisation: N single DiffNet evaluations (bsg_diffnet_forward, eager launches) on identical inputs at a bench-regime shape;
any run that differs from the first is reported with the frames / mel bins it differs in (one evaluation has a receptive field of
+-75 frames, so the differing frames point at the row tile where the glitch happened).  MEASUREMENT INFRASTRUCTURE.
usage: python tests/tools/exp_race4.py [B T runs]"""
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
RUNS = int(sys.argv[3]) if len(sys.argv) > 3 else 300
K = 100
dev = torch.device("cuda", 0)
junk = [torch.randn(256, 1024, 1024, device=dev) for _ in range(4)]   # 4 GB of non-zero memory, freed below: recycled by later allocations
del junk
torch.cuda.empty_cache()
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
sched = O.schedule_buffers(O.linear_beta_schedule(K, 0.06))
plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, precision="fp16x2", device=dev)
g = torch.Generator(device=dev)
g.manual_seed(777)
cond = torch.randn((B, T, 256), generator=g, device=dev)
x = torch.randn((B, 1, 80, T), generator=g, device=dev)
first = {}
bad = 0
SLEEP = float(os.environ.get("RACE_SLEEP", "0"))
EVERY = int(os.environ.get("RACE_EVERY", "10"))
for i in range(RUNS):
    if SLEEP > 0 and i >= 3 and i % EVERY == 0:
        import time
        torch.cuda.synchronize()
        time.sleep(SLEEP)             # idle GPU: clocks drop; the next evaluation starts cold
    t = (i * 7) % 3 * 40 + 5          # three different steps, revisited
    out = plan.denoise(x, t, cond)
    if t not in first:
        first[t] = out.clone()
        continue
    if not torch.equal(out, first[t]):
        bad += 1
        d = (out != first[t])
        idx = d.nonzero()
        fr = idx[:, 3]
        print(f"run {i} (t={t}): {idx.shape[0]} elements differ, max |diff| {float((out - first[t]).abs().max()):.3e}; batch rows "
              f"{sorted(set(idx[:, 0].tolist()))}; mel bins {len(set(idx[:, 2].tolist()))}; frames {int(fr.min())}..{int(fr.max())} (256-row tiles {int(fr.min()) // 256}..{int(fr.max()) // 256})",
              flush=True)
print(f"B={B} T={T}: {bad} of {RUNS - 3} repeated evaluations differ", flush=True)

"""Determinism of one DiffNet evaluation (debugging aid): rows that differ between repeated runs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_probe as G
B, T, K = int(os.environ.get("RB", 3)), int(os.environ.get("RT", 700)), 100
sd, sched, plan, inp, O, synth = G._diff_setup(B, T, K, "fp16x2")
x = inp["start_noise"].cuda(); cond = inp["cond"].cuda()
ref = O.diffnet_forward(sd, inp["start_noise"], torch.full((B,), 50), inp["cond"].transpose(1, 2))
outs = [plan.denoise(x, 50, cond).cpu() for _ in range(12)]
print("max|out-ref| %.3e" % (outs[0] - ref).abs().max().item())
for i, o in enumerate(outs[1:]):
    d = (o - outs[0]).abs()            # [B,1,M,T]
    rows = d.amax(dim=(1, 2))           # [B,T]
    bad = (rows > 0).nonzero().tolist()
    print("run", i + 1, "max diff %.3e" % d.max().item(), "rows differing", len(bad), bad[:12])

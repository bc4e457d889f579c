"""Precision / time of the skip-sum GEMM with ONE fp16 MMA per product (BSG_SKIP_X1=1) against the default two (hi + lo weights):
max |mel - oracle| of the full K=100 sampling for several seeds, and the GEMM's duration at cfg3.  MEASUREMENT INFRASTRUCTURE."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_oracle as O  # noqa: E402
import synth  # noqa: E402
from bisinger_b200 import B200DiffNet, DiffusionPlan  # noqa: E402

K = 100
dev = torch.device("cuda", 0)
sched = O.schedule_buffers(O.linear_beta_schedule(K, 0.06))
sd = synth.diffnet_state(1234)
net = B200DiffNet(80)
net.load_state_dict(sd, strict=True)
plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, precision="fp16x2", device=dev)
errs = []
for seed, B, T in ((7, 2, 40), (8, 1, 150), (31, 2, 300), (77, 3, 700), (91, 1, 938)):
    inp = synth.kernel_inputs(seed, B, T, K)
    with torch.no_grad():
        ref = O.diffusion_infer(sd, sched, torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX), inp["cond"], K, inp["step_noise"],
                                inp["fs2_mel"], inp["start_noise"])
    mel = plan.sample(inp["cond"].to(dev), inp["fs2_mel"].to(dev), inp["start_noise"].to(dev), inp["step_noise"].to(dev)).cpu()
    errs.append(float((mel - ref).abs().max()))
print(f"BSG_SKIP_X1={os.environ.get('BSG_SKIP_X1', '0')}: max |mel - oracle| per case " + " ".join(f"{e:.2e}" for e in errs) +
      f" | skip-sum GEMM at cfg3: {plan.time_kernel(2, 32, 1875, 50) * 1e3:.1f} us", flush=True)

"""Which 100x outlier weight rows hurt which contraction mode (range-robustness study; MEASUREMENT INFRASTRUCTURE)."""
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
inp = synth.kernel_inputs(502, 2, 200, K)


def run(sd, tag):
    with torch.no_grad():
        ref = O.diffusion_infer(sd, sched, torch.tensor(synth.SPEC_MIN), torch.tensor(synth.SPEC_MAX), inp["cond"], K, inp["step_noise"],
                                inp["fs2_mel"], inp["start_noise"])
    out = []
    for prec in ("fp16x2!", "bf16x3"):
        net = B200DiffNet(80)
        net.load_state_dict(sd, strict=True)
        plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, precision=prec, device=dev)
        mel = plan.sample(inp["cond"].to(dev), inp["fs2_mel"].to(dev), inp["start_noise"].to(dev), inp["step_noise"].to(dev)).cpu()
        out.append(f"{prec} {float((mel - ref).abs().max()):.2e}")
    print(f"{tag:40s} " + "  ".join(out), flush=True)


base = synth.diffnet_state(1234)
run(base, "no outliers")
for names, scale in ((("dilated_conv",), 100.0), (("conditioner_projection",), 100.0), (("output_projection:skip",), 100.0),
                     (("output_projection:res",), 20.0), (("dilated_conv", "conditioner_projection", "output_projection:skip"), 100.0),
                     (("output_projection:skip",), 10.0)):
    sd = {k: v.clone() for k, v in base.items()}
    g = torch.Generator().manual_seed(9)
    for l in (0, 7, 19):
        for name in names:
            nm, _, part = name.partition(":")
            rows = torch.randint(0, 256 if part else 512, (3,), generator=g)
            if part == "skip":
                rows = rows + 256
            sd[f"residual_layers.{l}.{nm}.weight"][rows] *= scale
    run(sd, f"{'+'.join(names)} x{scale:g}")

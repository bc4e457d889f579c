"""The "existing GPU path" beside the product: the reference algorithm as eager PyTorch on the same B200 (SURVEY.md section 8d asks
for it as the honest baseline).  TEST / MEASUREMENT INFRASTRUCTURE: it executes the oracle restatement (oracle/svs_oracle.py, pinned
to the executed reference) with every tensor on cuda:0 -- fp32 (cuDNN TF32 as torch defaults), TF32 everywhere, and bf16 autocast --
on one cfg3 batch (32 x 10 s, K = 100 sampler + vocoder).  Not part of bench.py; numbers go to profiles/."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import svs_oracle as O  # noqa: E402
import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
K = 100
dev = torch.device("cuda", 0)
to = lambda d: {k: v.to(dev) for k, v in d.items()}
sd, vsd = to(synth.diffnet_state(1234)), to(synth.hifigan_state(4321))
sched = to(O.schedule_buffers(O.linear_beta_schedule(K, 0.06)))
smin, smax = torch.tensor(synth.SPEC_MIN, device=dev), torch.tensor(synth.SPEC_MAX, device=dev)
inp = to(synth.kernel_inputs(1000, B, T, 1))
vin = to(synth.vocoder_inputs(1001, B, T))
step_noise = torch.randn((K, B, 1, 80, T), device=dev)
audio = B * T * 128 / 24000


def one():
    with torch.no_grad():
        mel = O.diffusion_infer(sd, sched, smin, smax, inp["cond"], K, step_noise, inp["fs2_mel"], inp["start_noise"])
        return O.hifigan_forward(vsd, synth.HIFIGAN_CONFIG, mel.transpose(1, 2), vin["f0"], vin["rand_ini"], vin["src_noise"])


def timed(name):
    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"eager torch {torch.__version__} on B200, {name}: B={B} T={T} K={K}: {dt * 1e3:.0f} ms per batch, {audio / dt:.1f} audio-s/s", flush=True)


torch.backends.cuda.matmul.allow_tf32 = False
timed("fp32 (matmul fp32, cuDNN conv TF32 = torch default)")
torch.backends.cudnn.allow_tf32 = False
timed("strict fp32 (no TF32 anywhere; the reference's numerics)")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
timed("TF32 everywhere")
with torch.autocast("cuda", dtype=torch.bfloat16):
    timed("bf16 autocast")

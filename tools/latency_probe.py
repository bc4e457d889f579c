"""Latency of the three stages at a small shape (default cfg1/cfg2: B=1, T=938 = 5 s) with CUDA events, graph replay vs plain
launches (BSG_VOC_GRAPH / BSG_PE_GRAPH are read at plan creation).  Synthetic weights only."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200 import B200DiffNet, DiffusionPlan  # noqa: E402
from bisinger_b200.diffusion import _schedule_buffers, linear_beta_schedule  # noqa: E402
from bisinger_b200.pitch import B200PitchExtractor  # noqa: E402
from bisinger_b200.vocoder import B200HifiGanGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 938
dev = torch.device("cuda", 0)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def build(graph):
    os.environ["BSG_VOC_GRAPH"] = os.environ["BSG_PE_GRAPH"] = "1" if graph else "0"
    gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
    gen.load_folded_state_dict(synth.hifigan_state(4321), strict=True)
    gen.build_plan(dev)
    pe = B200PitchExtractor().eval()
    pe.load_state_dict(synth.pe_state(777, 2), strict=True)
    pe.build_plan(dev)
    return gen, pe


vin = synth.vocoder_inputs(3, B, T)
mel_v, f0 = vin["mel"].to(dev), vin["f0"].to(dev)
mel_p = vin["mel"].transpose(1, 2).contiguous().to(dev)
audio = B * T * 128 / 24000
n = 20 if B * T < 20000 else 5
for graph in (True, False):
    gen, pe = build(graph)
    v = timeit(lambda: gen(mel_v, f0, seed=1), n)
    p = timeit(lambda: pe(mel_p), n)
    print(f"B={B} T={T} graph={int(graph)}: vocoder {v:.3f} ms ({audio / v * 1e3:.0f} audio-s/s), pitch extractor {p:.3f} ms")
    if not graph:
        v, p = vg, pg
    vg, pg = v, p
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
K = 100
sched = _schedule_buffers(linear_beta_schedule(K, 0.06))
plan = DiffusionPlan(net, sched, K, K, synth.SPEC_MIN, synth.SPEC_MAX, device=dev)
inp = synth.kernel_inputs(7, B, T, 1)
cond, fs2 = inp["cond"].to(dev), inp["fs2_mel"].to(dev)
s = timeit(lambda: plan.sample(cond, fs2, seed=1), n=5)
print(f"B={B} T={T}: sampler K={K} {s:.2f} ms ({audio / s * 1e3:.0f} audio-s/s); mel->f0->wav chain {audio / (s + v + p) * 1e3:.0f} audio-s/s, "
      f"RTF {(s + v + p) / 1e3 / audio:.5f}")

"""ncu target for the per-kernel roofline table: one DiffNet evaluation (= the kernels of one sampler step), one PitchExtractor
forward and one vocoder forward at cfg3 (B=32, T=1875), bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off).
Plain launches (BSG_VOC_GRAPH=0, BSG_PE_GRAPH=0) so that every kernel is listed.  Synthetic weights only."""
import os
import sys

os.environ.setdefault("BSG_VOC_GRAPH", "0")
os.environ.setdefault("BSG_PE_GRAPH", "0")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200 import B200DiffNet, DiffusionPlan  # noqa: E402
from bisinger_b200.diffusion import _schedule_buffers, linear_beta_schedule  # noqa: E402
from bisinger_b200.pitch import B200PitchExtractor  # noqa: E402
from bisinger_b200.vocoder import B200HifiGanGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
dev = torch.device("cuda", 0)
net = B200DiffNet(80)
net.load_state_dict(synth.diffnet_state(1234), strict=True)
K = 100
plan = DiffusionPlan(net, _schedule_buffers(linear_beta_schedule(K, 0.06)), K, K, synth.SPEC_MIN, synth.SPEC_MAX, device=dev)
gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
gen.load_folded_state_dict(synth.hifigan_state(4321), strict=True)
gen.build_plan(dev)
pe = B200PitchExtractor().eval()
pe.load_state_dict(synth.pe_state(777, 2), strict=True)
pe.build_plan(dev)
inp = synth.kernel_inputs(7, B, T, 1)
vin = synth.vocoder_inputs(3, B, T)
spec, cond = inp["start_noise"].to(dev), inp["cond"].to(dev)
mel_v, f0 = vin["mel"].to(dev), vin["f0"].to(dev)
mel_p = vin["mel"].transpose(1, 2).contiguous().to(dev)
for _ in range(2):
    plan.denoise(spec, 50, cond)
    pe(mel_p)
    gen(mel_v, f0, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.denoise(spec, 50, cond)
pe(mel_p)
gen(mel_v, f0, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

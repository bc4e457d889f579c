"""ncu target: ONE vocoder forward at a BASELINE shape (default cfg3: B=32, T=1875) between cudaProfilerStart/Stop, plain launches
(BSG_VOC_GRAPH=0) so that every kernel is listed.  Use with
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/voc.csv --metrics gpu__time_duration.sum,\
dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed python tools/profile_vocoder.py
and summarise with  python tools/ncu_table.py gpurun_out/voc.csv"""
import os
import sys

os.environ.setdefault("BSG_VOC_GRAPH", "0")
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bisinger_b200 import synthetic as synth  # noqa: E402
from bisinger_b200.vocoder import B200HifiGanGenerator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
dev = torch.device("cuda", 0)
gen = B200HifiGanGenerator(synth.HIFIGAN_CONFIG)
gen.load_folded_state_dict(synth.hifigan_state(4321), strict=True)
gen.build_plan(dev)
vin = synth.vocoder_inputs(3, B, T)
mel_v, f0 = vin["mel"].to(dev), vin["f0"].to(dev)
for _ in range(2):
    gen(mel_v, f0, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
gen(mel_v, f0, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

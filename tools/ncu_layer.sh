#!/bin/bash
# ncu --set full of the two DiffNet layer kernels (3 warm-up launches skipped, 2 captured each), source-level info on.
mkdir -p gpurun_out
for w in 1 0; do
  BSG_WHICH=$w timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 2 -f -o gpurun_out/prof_k$w \
      python tools/gpu_probe.py --run proftarget > gpurun_out/ncu_k$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

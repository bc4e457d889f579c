#!/bin/bash
# ncu --set full of the DiffNet per-layer kernels (0 gate, 1 residual, 2 skip-sum): 3 warm-up launches skipped, 2 captured each.
mkdir -p gpurun_out
for w in ${BSG_KERNELS:-0 1 2}; do
  BSG_WHICH=$w timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 2 -f -o gpurun_out/prof_k$w \
      python tests/tools/gpu_probe.py --run proftarget > gpurun_out/ncu_k$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# quick visit: bench (x3) + ncu full capture of the residual/skip GEMM
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
BSG_WHICH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o gpurun_out/prof_k1 \
    python tests/tools/gpu_probe.py --run proftarget > gpurun_out/ncu_k1.log 2>&1

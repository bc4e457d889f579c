#!/bin/bash
# One GPU-box visit: tests, smoke, bench, ncu launch list + full capture of the dominant kernels. Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench.log
timeout 300 python bench.py --steps 3 --warmup 3 --precision bf16 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16x1.log
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4500 -c 500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 4309 -c 4 -o gpurun_out/prof_layer \
      python tests/tools/gpu_probe.py --run proftarget > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi

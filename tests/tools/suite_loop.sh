#!/bin/bash
# Runs the part of the GPU suite that has shown the rare run-to-run difference N times and prints one line per run.
# usage: suite_loop.sh N "<pytest -k expression>"   (environment switches are inherited)
N=${1:-5}; K=${2:-"fft_decoder or sampler_bench_regime or graph_path"}
for i in $(seq 1 $N); do
  out=$(python -m pytest tests -x -q -m gpu -k "$K" 2>&1)
  echo "run $i: $(echo "$out" | tail -1)"
  echo "$out" | grep -E "^E +AssertionError" | head -2
done

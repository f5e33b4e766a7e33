#!/bin/bash
# Round 2, session s (N GPUs): final multi-GPU lines with default flags.
N=${DP_N:-8}
mkdir -p gpurun_out
SKIP_TEST=1 source tools/gpu_r2c.sh.lib
run c3_default "A=1" --config c3 --steps 50 --warmup 5 --no-cpu-baseline
if [ "$WITH_C4" == "1" ]; then
  run c4_default "A=1" --config c4 --steps 40 --warmup 5 --no-cpu-baseline --no-parity-check
fi

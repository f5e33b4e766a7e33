#!/bin/bash
# Round 2, session d (8 GPUs): multi-GPU parity at 8 ranks (multicast paths), then the BASELINE 8-GPU configs.
N=${DP_N:-8}
source tools/gpu_r2c.sh.lib
S=${BENCH_STEPS:-50}
run c3_items "A=1" --config c3 --steps $S --warmup 5 --no-cpu-baseline --parallel items --per-rank-kernels
run c3_rows "A=1" --config c3 --steps 30 --warmup 5 --no-cpu-baseline --parallel rows --no-parity-check --skip-e2e
run c4_rows "A=1" --config c4 --steps 30 --warmup 5 --no-cpu-baseline --parallel rows
run c5_items_b2048 "A=1" --config c5 --users 1000000 --steps 20 --warmup 5 --no-cpu-baseline --parallel items

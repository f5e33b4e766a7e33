#!/bin/bash
# Round 2, session n (2 GPUs): multi-GPU parity incl. sharded host staging; default multi-GPU bench lines; stream-overlap A/B.
N=2
mkdir -p gpurun_out
echo "== dp test"; RCD_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_d_multigpu.py -q -m gpu -x -s > gpurun_out/t_dp_n${N}b.log 2>&1; echo "rc=$?"; grep -E "DP_|passed|failed|Error|differs|host-staged" gpurun_out/t_dp_n${N}b.log | tail -12
DP_N=2 SKIP_TEST=1 source tools/gpu_r2c.sh.lib
run c3_default "A=1" --config c3 --steps 50 --warmup 5 --no-cpu-baseline
run c4_default "A=1" --config c4 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-check
echo "== single GPU: RCD_OVERLAP A/B"
source tools/gpu_r2b.sh.lib
run c3_overlap0 "RCD_OVERLAP=0 CUDA_VISIBLE_DEVICES=0" --config c3 --no-cpu-baseline --no-parity-check --skip-e2e --no-profile
run c3_overlap1 "RCD_OVERLAP=1 CUDA_VISIBLE_DEVICES=0" --config c3 --no-cpu-baseline --no-parity-check --skip-e2e --no-profile

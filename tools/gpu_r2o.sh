#!/bin/bash
# Round 2, session o (N GPUs): MF row-parallel with the NCCL exchange + deferred Adam against the fused peer-memory exchange.
N=${DP_N:-2}
mkdir -p gpurun_out
if [ "$SKIP_TEST" != "1" ]; then
  echo "== dp test"; RCD_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_d_multigpu.py -q -m gpu -x -s > gpurun_out/t_dp_n${N}c.log 2>&1; echo "rc=$?"; grep -E "DP_|passed|failed|Error|differs|deferred" gpurun_out/t_dp_n${N}c.log | tail -12
fi
SKIP_TEST=1 source tools/gpu_r2c.sh.lib
run c4_nccl_lazy "A=1" --config c4 --steps 40 --warmup 5 --no-cpu-baseline --no-parity-check --dp-exchange nccl
run c4_p2p "A=1" --config c4 --steps 40 --warmup 5 --no-cpu-baseline --no-parity-check --dp-exchange p2p --skip-e2e
if [ "$WITH_C3" == "1" ]; then
  run c3_default "A=1" --config c3 --steps 50 --warmup 5 --no-cpu-baseline
fi

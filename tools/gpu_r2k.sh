#!/bin/bash
# Round 2, session k (1 GPU): CTA-pair fused decoder after relaxing the remote accumulator-empty arrive.
mkdir -p gpurun_out
echo "== pytest pair"; RCD_GEMM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_k.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|timed out|trap|aligned" gpurun_out/pytest_k.log | tail -8
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --no-parity-check --skip-e2e"
run c3_pair6b "RCD_GEMM_PAIR=1" --config c3 $Q
run c3_single4 "RCD_GEMM_PAIR=0" --config c3 $Q
run c5_b8192_pair6b "RCD_GEMM_PAIR=1" --config c5 --users 1000000 --batch 8192 --steps 20 $Q
run c5_b8192_single4 "RCD_GEMM_PAIR=0" --config c5 --users 1000000 --batch 8192 --steps 20 $Q

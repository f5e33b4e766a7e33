#!/bin/bash
# Round 2, session h (1 GPU): CTA-pair fused decoder with a 5-stage ring.
mkdir -p gpurun_out
echo "== pytest RCD_GEMM_PAIR=1"; RCD_GEMM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py tests/test_gpu_h_native.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_h.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|timed out|trap|aligned" gpurun_out/pytest_h.log | tail -8
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --no-parity-check"
run c3_pair5 "RCD_GEMM_PAIR=1" --config c3 $Q
run c3_single "RCD_GEMM_PAIR=0" --config c3 $Q
run c5_b8192_pair5 "RCD_GEMM_PAIR=1" --config c5 --users 1000000 --batch 8192 --steps 20 $Q

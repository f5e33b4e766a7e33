#!/bin/bash
# Round 2, session f (1 GPU): CTA-pair fused decoder kernel — parity, then A/B on C3.
mkdir -p gpurun_out
echo "== pytest with RCD_GEMM_PAIR=1"; RCD_GEMM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py tests/test_gpu_g_benchshapes.py -q -m gpu --timeout 600 -x -k "decoder or nll or step or bench or curve" > gpurun_out/pytest_f.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|timed out|trap" gpurun_out/pytest_f.log | tail -20
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --no-parity-check"
run c3_pair "RCD_GEMM_PAIR=1" --config c3 $Q
run c3_nopair "RCD_GEMM_PAIR=0" --config c3 $Q --no-profile
run c5_b8192_pair "RCD_GEMM_PAIR=1" --config c5 --users 1000000 --batch 8192 --steps 20 $Q

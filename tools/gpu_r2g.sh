#!/bin/bash
# Round 2, session g (1 GPU): fused decoder kernel with 16 epilogue warps (single-CTA and CTA-pair), catch-up unroll.
mkdir -p gpurun_out
for pair in 0 1; do
  echo "== pytest RCD_GEMM_PAIR=$pair"; RCD_GEMM_PAIR=$pair timeout 900 python -m pytest tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py tests/test_gpu_i_lazy_adam.py tests/test_gpu_h_native.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_g$pair.log 2>&1; echo "rc=$?"
  grep -E "passed|failed|FAILED|Error|error|timed out|trap" gpurun_out/pytest_g$pair.log | tail -8
done
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --no-parity-check"
run c3_e16 "RCD_GEMM_PAIR=0" --config c3 $Q
run c3_e16_pair "RCD_GEMM_PAIR=1" --config c3 $Q
run c2_e16 "RCD_GEMM_PAIR=0" --config c2 --steps 100 --warmup 10 $Q
run c4_e16_pair "RCD_GEMM_PAIR=1" --config c4 --steps 50 $Q

#!/bin/bash
# Round 2, session i (1 GPU): fused decoder kernel with a 4-stage operand ring (single staging buffer), single and pair.
mkdir -p gpurun_out
ST4=$PWD/recoder_b200/csrc/librecoder_b200_st4.so
echo "== pytest 4-stage build"; RCD_LIB=$ST4 timeout 900 python -m pytest tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_i.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|timed out|trap|aligned" gpurun_out/pytest_i.log | tail -8
echo "== pytest 4-stage build, pair"; RCD_GEMM_PAIR=1 RCD_LIB=$ST4 timeout 900 python -m pytest tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_i2.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|timed out|trap|aligned" gpurun_out/pytest_i2.log | tail -8
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --no-parity-check --skip-e2e"
run c3_st3 "A=1" --config c3 $Q
run c3_st4 "RCD_LIB=$ST4" --config c3 $Q
run c3_st4_pair6 "RCD_LIB=$ST4 RCD_GEMM_PAIR=1" --config c3 $Q

#!/bin/bash
# Round 2, session m (1 GPU): deferred Adam with the next pool's rows caught up on the side stream.
mkdir -p gpurun_out
echo "== pytest lazy / native / eval"; timeout 900 python -m pytest tests/test_gpu_i_lazy_adam.py tests/test_gpu_h_native.py tests/test_gpu_f_eval.py tests/test_gpu_c_step.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_m.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|timed out|trap" gpurun_out/pytest_m.log | tail -8
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --skip-e2e"
run c3_prefetch "RCD_LAZY_PREFETCH=1" --config c3 $Q
run c3_noprefetch "RCD_LAZY_PREFETCH=0" --config c3 $Q --no-profile
run c5_b512_prefetch "RCD_LAZY_PREFETCH=1" --config c5 --users 1000000 --batch 512 --steps 40 $Q
run c4_prefetch "RCD_LAZY_PREFETCH=1" --config c4 --steps 50 $Q

#!/bin/bash
# Quick GPU check: kernel tests first (fail fast), then the step tests, then a short bench.
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_b_kernels.py -q -m gpu -x --timeout 600 > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_kernels.log
echo "== collate+step"; timeout 1200 python -m pytest tests/test_gpu_a_collate.py tests/test_gpu_c_step.py -q -m gpu -x --timeout 600 > gpurun_out/t_step.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_step.log
echo "== bench c3"; timeout 900 python bench.py --config c3 ${BENCH_ARGS} > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "rc=$?"; tail -3 gpurun_out/bench_c3.log; cat gpurun_out/bench_c3.json

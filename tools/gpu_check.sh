#!/bin/bash
# First-contact GPU run: every stage is time-boxed and logged to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gemm probe"; timeout 900 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/gemm_probe.log
for f in tests/test_gpu_a_collate.py tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py; do
  echo "== $f"
  timeout 1200 python -m pytest $f -q -m gpu -x --timeout 600 > gpurun_out/$(basename $f .py).log 2>&1
  echo "rc=$?"; tail -25 gpurun_out/$(basename $f .py).log
done

#!/bin/bash
# One GPU session: smoke, parity tests, bench, ncu launch list, ncu full capture of the top kernels.
# Every stage is time-boxed and logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
if [ "$SKIP_TESTS" != "1" ]; then
  echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
fi
echo "== bench c3"; timeout 900 python bench.py --config c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "rc=$?"; tail -3 gpurun_out/bench_c3.log; cat gpurun_out/bench_c3.json
if [ "$SKIP_NCU" != "1" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
  echo "== ncu full: adam + gemm"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adam|k_gemm_tc|k_loss_grad' -s 24 -c 10 -f -o gpurun_out/prof_top \
    python bench.py --config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
  ls -la gpurun_out
fi

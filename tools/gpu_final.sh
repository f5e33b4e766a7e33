#!/bin/bash
# Final single-GPU evidence of a round: smoke, full GPU test suite, default bench, ncu launch list + full capture.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.log; echo "rc=$?"; cut -c1-1200 gpurun_out/bench_c3.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
ARGS="--config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_adam|k_gemm_tc|k_decoder_fused|k_encoder_wgrad" -s 24 -c 12 -f -o gpurun_out/prof_top python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -8

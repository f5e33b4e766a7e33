#!/bin/bash
# ncu captures of the bench workload (short matrix, 2 timed steps): launch list + full sets of the top kernels.
mkdir -p gpurun_out
ARGS="--config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
  python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_adam|k_gemm_tc|k_decoder_fused}" -s ${NCU_SKIP:-16} -c ${NCU_COUNT:-10} -f -o gpurun_out/prof_top \
  python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out

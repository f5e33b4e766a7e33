#!/bin/bash
# Round 2, session e (1 GPU): deferred dense Adam (parity + effect per config), activations, ncu evidence.
mkdir -p gpurun_out
echo "== pytest lazy adam + activations + native"; timeout 1200 python -m pytest tests/test_gpu_i_lazy_adam.py tests/test_gpu_e_general.py tests/test_gpu_h_native.py tests/test_gpu_f_eval.py -q -m gpu --timeout 900 -x > gpurun_out/pytest_e.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_e.log | tail -20
source tools/gpu_r2b.sh.lib
Q="--no-cpu-baseline --no-parity-check"
run c3_lazy "A=1" --config c3 $Q
run c3_dense "RCD_LAZY_ADAM=0" --config c3 $Q --no-profile
run c4_lazy "A=1" --config c4 --steps 50 $Q
run c5_b512_lazy "A=1" --config c5 --users 1000000 --batch 512 --steps 40 $Q
run c5_b2048_lazy "A=1" --config c5 --users 1000000 --batch 2048 --steps 30 $Q
run c2_lazy "A=1" --config c2 --steps 100 --warmup 10 $Q --no-profile
echo "== ncu launch list c3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c3.csv \
  python bench.py --config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile --no-parity-check > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full: top kernels of one steady-state step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_adam|k_gemm_tc|k_decoder_fused|k_encoder_wgrad|k_sddmm|k_encoder_fwd|k_gather_rows|k_dz_act' -s 60 -c 24 -f -o gpurun_out/prof_top \
  python bench.py --config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile --no-parity-check > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_top.ncu-rep

#!/bin/bash
# Round 2, session j (1 GPU): ncu full capture of the fused decoder kernel, single-CTA (4-stage) and CTA-pair (6-stage).
mkdir -p gpurun_out
for pair in 0 1; do
  echo "== ncu full fused kernel pair=$pair"
  RCD_GEMM_PAIR=$pair timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_decoder_fused' -s 12 -c 6 -f -o gpurun_out/prof_fused_pair$pair \
    python bench.py --config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile --no-parity-check > gpurun_out/ncu_fused$pair.log 2>&1; echo "rc=$?"
done
ls -la gpurun_out/prof_fused_pair*.ncu-rep

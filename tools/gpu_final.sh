#!/bin/bash
# Final single-GPU validation: what the driver runs at round end (smoke, all GPU tests, reference arm, default bench).
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -5
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.log; echo "rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
echo "== bench default (driver flags)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver.json 2> gpurun_out/bench_driver.log; echo "rc=$?"
source tools/gpu_r2b.sh.lib
echo "== bench default (100 steps)"
run c3_final "A=1" --config c3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_driver.json').read().strip().splitlines()[-1])
print('driver-flags run:', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e'], 'clocks', d['clocks'])
print('roofline', d['roofline']); print('cpu', d['cpu_baseline']); print('parity', d['parity_check'])
PY

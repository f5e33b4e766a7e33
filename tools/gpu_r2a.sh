#!/bin/bash
# Round 2, session a: smoke, all GPU parity tests, one bench line per BASELINE config on one GPU.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
grep -E "loss curve|passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -30
run() { # name, args...
  name=$1; shift
  echo "== bench $name"
  timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.log; echo "rc=$?"
  tail -2 gpurun_out/bench_$name.log | cut -c1-300
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
  print({k:d[k] for k in ('value','ms_per_step','gpu_launches','items_per_batch')}, 'e2e',d['e2e']['value'], 'cpu',(d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'))
  print('roofline',d['roofline']); print('parity',d['parity_check']); print('host',d['host_ms_per_step'])
  print({k:v['ms_per_step'] for k,v in list(d['kernels'].items())[:12]})
except Exception as e: print('parse failed',e)
PY
}
run c3 --config c3
run c1 --config c1 --steps 200 --warmup 10
run c2 --config c2 --steps 100 --warmup 10
run c4 --config c4 --steps 50
run c5_b512 --config c5 --batch 512 --steps 30 --no-cpu-baseline
run c5_b2048 --config c5 --batch 2048 --steps 30
run c5_b8192 --config c5 --batch 8192 --steps 20 --no-cpu-baseline
echo "== reference arm c3"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_c3_reference_arm.json 2> gpurun_out/bench_c3_reference_arm.log; echo "rc=$?"; cut -c1-400 gpurun_out/bench_c3_reference_arm.json
echo "== ncu launch list c3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c3.csv \
  python bench.py --config c3 --users 100000 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e --no-profile --no-parity-check > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -50

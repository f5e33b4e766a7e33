#!/bin/bash
# Round 2, session b: native step executor (parity + host time), optimizer grid A/B, small configs.
mkdir -p gpurun_out
echo "== pytest native + benchshapes + kernels"; timeout 1200 python -m pytest tests/test_gpu_h_native.py tests/test_gpu_g_benchshapes.py tests/test_gpu_b_kernels.py tests/test_gpu_c_step.py -q -m gpu --timeout 900 -s -x > gpurun_out/pytest_b.log 2>&1; echo "rc=$?"
grep -E "loss curve|Recoder.train native|passed|failed|FAILED|Error|error" gpurun_out/pytest_b.log | tail -20
run() { # name, env, args...
  name=$1; shift; envs=$1; shift
  echo "== bench $name ($envs)"
  env $envs timeout 600 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.log; echo "rc=$?"
  tail -1 gpurun_out/bench_$name.log | cut -c1-200
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
  print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e',d['e2e']['value'],d['e2e']['ms_per_step'], 'cpu',(d.get('cpu_baseline') or {}).get('value'))
  print('roofline',d['roofline']); print('parity',(d['parity_check'] or {}).get('rel_err')); print('host',d['host_ms_per_step'])
  print({k:v['ms_per_step'] for k,v in list(d['kernels'].items())[:10]})
except Exception as e: print('parse failed',e)
PY
}
Q="--no-cpu-baseline --no-parity-check"
run c3_native_g4 "RCD_STREAM_CTAS_PER_SM=4" --config c3 $Q
run c3_native_g8 "RCD_STREAM_CTAS_PER_SM=8" --config c3 $Q --no-profile
run c3_native_g3 "RCD_STREAM_CTAS_PER_SM=3" --config c3 $Q --no-profile
run c3_native_g2 "RCD_STREAM_CTAS_PER_SM=2" --config c3 $Q --no-profile
run c3_python_g4 "RCD_NATIVE_STEP=0" --config c3 $Q --no-profile
run c1_native "A=1" --config c1 --steps 200 --warmup 10
run c1_python "RCD_NATIVE_STEP=0" --config c1 --steps 200 --warmup 10 $Q --no-profile
run c2_native "A=1" --config c2 --steps 100 --warmup 10
run c4_native "A=1" --config c4 --steps 50 $Q
ls gpurun_out | wc -l

#!/bin/bash
# Round 2, session c (N GPUs): multi-GPU parity test, then C3 bench in both parallel modes with per-rank breakdowns.
N=${DP_N:-2}
mkdir -p gpurun_out
nproc > gpurun_out/host_n$N.txt; free -g >> gpurun_out/host_n$N.txt; nvidia-smi topo -m >> gpurun_out/host_n$N.txt 2>&1
if [ "$SKIP_TEST" != "1" ]; then
  echo "== dp test"; RCD_TEST_WORLD=$N timeout 900 python -m pytest tests/test_gpu_d_multigpu.py -q -m gpu -x -s > gpurun_out/t_dp_n$N.log 2>&1; echo "rc=$?"; grep -E "losses|DP_|passed|failed|Error|differs" gpurun_out/t_dp_n$N.log | tail -40
fi
run() { # tag, env, args
  tag=$1; shift; envs=$1; shift
  echo "== bench N=$N $tag ($envs) $@"
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.log
  echo "rc=$?"; grep -E "Error|error|Traceback" gpurun_out/bench_n${N}_$tag.log | head -5
  python - <<PY
import json
try:
  d=[json.loads(l) for l in open("gpurun_out/bench_n${N}_$tag.json") if l.startswith("{")][-1]
  print('ms_per_step', round(d['ms_per_step'],4), 'users/s', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'h2d', d['e2e']['h2d_bytes_per_step'], 'n', d['items_per_batch'], 'launches', d['gpu_launches'])
  print('host', d.get('host_ms_per_step')); print('per_rank', d.get('per_rank_ms')); print('parity', d.get('parity_check'))
  print('roofline', d.get('roofline'))
  print({k:v['ms_per_step'] for k,v in d['kernels'].items()})
  for i,k in enumerate(d.get('per_rank_kernels') or []): print(i, {a:b for a,b in list(sorted(k.items(), key=lambda kv:-kv[1]))[:8]})
except Exception as ex:
  print('no json', ex)
PY
}
S=${BENCH_STEPS:-50}
run c3_items "A=1" --config c3 --steps $S --warmup 5 --no-cpu-baseline --parallel items --per-rank-kernels
run c3_rows "A=1" --config c3 --steps $S --warmup 5 --no-cpu-baseline --parallel rows --per-rank-kernels --no-parity-check
if [ "$EXTRA" == "1" ]; then
  run c3_items_py "RCD_NATIVE_STEP=0" --config c3 --steps $S --warmup 5 --no-cpu-baseline --parallel items --no-profile --no-parity-check
  run c4_rows "A=1" --config c4 --steps 30 --warmup 5 --no-cpu-baseline --parallel rows --no-parity-check
fi

#!/bin/bash
# 2-GPU session: DP parity test (nccl vs p2p vs single process), then the bench with the chosen exchanges.
# DP_MODES="p2p nccl", DP_ENVS="RCD_OVERLAP=1;RCD_OVERLAP=0" (semicolon-separated env settings), SKIP_TEST=1
mkdir -p gpurun_out
if [ "$SKIP_TEST" != "1" ]; then
  echo "== dp test"; RCD_TEST_WORLD=${DP_N:-2} timeout 600 python -m pytest tests/test_gpu_d_multigpu.py -q -m gpu -x -s > gpurun_out/t_dp.log 2>&1; echo "rc=$?"; grep -E "losses|DP_|passed|failed|Error|differs" gpurun_out/t_dp.log | tail -30
fi
N=${DP_N:-2}
IFS=';' read -ra ENVS <<< "${DP_ENVS:-RCD_OVERLAP=1}"
for mode in ${DP_MODES:-p2p nccl}; do
  for e in "${ENVS[@]}"; do
    tag="n${N}_${mode}_$(echo $e | tr -c 'A-Za-z0-9\n' '_')"
    echo "== bench c3 N=$N $mode $e"
    env $e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --dp-exchange $mode ${BENCH_ARGS} > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.log
    echo "rc=$?"; grep -E "Error|error" gpurun_out/bench_c3_$tag.log | head -5
    python - <<PY
import json
try:
  d=[json.loads(l) for l in open("gpurun_out/bench_c3_$tag.json") if l.startswith("{")][-1]
  print('ms_per_step', round(d['ms_per_step'],4), 'host', d.get('host_ms_per_step'), 'users/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'n', d['items_per_batch'])
  print({k:v['ms_per_step'] for k,v in d['kernels'].items()})
except Exception as ex:
  print('no json', ex)
PY
  done
done

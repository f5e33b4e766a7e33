#!/bin/bash
# 2-GPU session: DP parity test (nccl vs p2p vs single process), then the bench with both exchanges.
mkdir -p gpurun_out
echo "== dp test"; timeout 600 python -m pytest tests/test_gpu_d_multigpu.py -q -m gpu -x -s > gpurun_out/t_dp.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_dp.log
for mode in ${DP_MODES:-p2p nccl}; do
  echo "== bench c3 N=2 $mode"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --dp-exchange $mode ${BENCH_ARGS} > gpurun_out/bench_c3_n2_$mode.json 2> gpurun_out/bench_c3_n2_$mode.log
  echo "rc=$?"; tail -3 gpurun_out/bench_c3_n2_$mode.log; cat gpurun_out/bench_c3_n2_$mode.json
done

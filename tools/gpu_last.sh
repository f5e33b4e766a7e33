#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -5

"""Data-parallel parity on real GPUs (needs >= 2): NCCL all-reduce exchange and the fused peer-memory exchange
against the single-process step with the global batch.  Skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


@pytest.mark.gpu
def test_dp_exchange_matches_single_process():
  n = torch.cuda.device_count()
  if n < 2:
    pytest.skip('needs at least 2 GPUs')
  world = min(n, int(os.environ.get('RCD_TEST_WORLD', '2')))   # RCD_TEST_WORLD=4|8 on a bigger box (multicast path)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
         '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dp_worker.py')]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  sys.stdout.write(r.stdout[-4000:])
  sys.stderr.write(r.stderr[-4000:])
  assert r.returncode == 0 and 'DP_OK' in r.stdout

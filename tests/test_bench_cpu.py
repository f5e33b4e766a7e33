"""bench.py helpers that need no GPU: roofline denominators, algorithmic work per entry point, the DRAM
traffic lookup in the committed ncu summaries, the workload table against BASELINE.json's configs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_peaks_come_from_the_measured_file_or_the_stated_fallback():
  p = bench.load_peaks()
  assert p['source'] in ('measured', 'fallback')
  assert 3000 < p['hbm'] < 9000 and 500 < p['tensor_sustained'] <= p['tensor_burst'] < 2500


def test_algorithmic_work_per_entry_point():
  w = bench.WORKLOADS['c3']
  rows, n, H, I = 2048, 113000.0, w['width'], w['items']
  dense = 2.0 * rows * n * H
  for name in ('rcd_decoder_fwd_loss', 'rcd_decoder_dgrad', 'rcd_decoder_wgrad'):
    assert bench.kernel_work(name, w, rows, n, n, 2e5, (0, 0, 1)) == ('tensor', dense)
  params = 2 * I * H + I + H
  grads = 2 * n * H + n + H
  bound, work = bench.kernel_work('rcd_adam_step', w, rows, n, n, 2e5, (params, grads, 1))
  assert bound == 'hbm' and work == 24.0 * params + 4.0 * grads          # 24 B/param + 4 B per compact gradient element
  bound, work = bench.kernel_work('rcd_adam_step_p2p', w, rows, n, n, 2e5, (params, grads, 8))
  assert bound == 'nvlink' and work == 4.0 * (grads + params) * 7 / 8
  assert bench.kernel_work('rcd_gather_rows', w, rows, n, n, 2e5, (params, grads, 1)) == ('hbm', n * H * 6.0)


def test_ncu_traffic_lookup_matches_the_committed_summary():
  traffic, src = bench.ncu_traffic('rcd_adam_step')
  assert src is not None and src.startswith('r01') and src.endswith('_ncu_full_top_kernels.csv')
  assert 2.4e9 < traffic < 2.9e9            # one 200K x 512 table: 1.46 GB read + 1.17 GB written
  assert bench.ncu_traffic('rcd_collate') == (None, None)


def test_workloads_follow_baseline_json():
  with open(os.path.join(ROOT, 'BASELINE.json')) as fh:
    base = json.load(fh)
  assert len(base['configs']) == 5 and set(bench.WORKLOADS) == {'c1', 'c2', 'c3', 'c4', 'c5'}
  c3 = bench.WORKLOADS['c3']
  assert (c3['users'], c3['items'], c3['nnz'], c3['width'], c3['loss'], c3['model']) == \
         (1_000_000, 200_000, 100, 512, 'logloss', 'ae')                 # "1M x 200K synthetic, ~100 nnz/user, AE hidden=512"
  assert bench.WORKLOADS['c2']['users'] == 138_493 and bench.WORKLOADS['c2']['width'] == 200
  assert bench.WORKLOADS['c4']['model'] == 'mf' and bench.WORKLOADS['c4']['width'] == 256
  assert bench.WORKLOADS['c5']['items'] == 500_000 and bench.WORKLOADS['c5']['width'] == 1024


class _Args:
  parallel = 'items'
  dp_exchange = 'auto'


def test_l2_statement_is_computed_per_config():
  """Timing rule: no flush between steps, so `config.l2` must say whether the step's working set exceeds the L2 — true at
  the configuration the metric is quoted on (C3, every GPU count), NOT true at C1, and the line has to say so."""
  for world in (1, 2, 4, 8):
    w = bench.WORKLOADS['c3']
    assert bench.working_set_bytes(_Args, w, w['users'], 2048, world) > 4 * bench.L2_BYTES
    assert 'larger than the 126 MB L2' in bench.workload_config(_Args, w, w['users'], 2048, world)['l2']
  w = bench.WORKLOADS['c1']
  assert bench.working_set_bytes(_Args, w, w['users'], 256, 1) < bench.L2_BYTES
  assert 'FITS' in bench.workload_config(_Args, w, w['users'], 256, 1)['l2']
  cfg = bench.workload_config(_Args, bench.WORKLOADS['c3'], 1_000_000, 2048, 8)
  assert cfg['global_batch'] == 16384 and cfg['parallelism'] == 'items8' and 'model' in cfg


def test_reference_arm_prints_the_contract_line():
  """`bench.py --impl reference` (tier framing (4)): the reference's own CPU implementation on this arm's config, metric
  and unit, with `impl`, `cpu_baseline` and a zero-copy `e2e`.  Runs the unmodified reference when it is importable
  (/root/reference here, baseline/_ref on the GPU box), the oracle port otherwise."""
  import subprocess
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'c1',
                        '--steps', '2', '--warmup', '1'], capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  line = json.loads(out.stdout.strip().splitlines()[-1])
  assert line['impl'] == 'reference' and line['metric'] == 'users/sec (train step)' and line['unit'] == 'users/s'
  assert line['higher_is_better'] is True and line['n_gpus'] == 1 and line['steps'] == 2 and line['warmup'] == 1
  assert line['value'] > 0 and abs(line['value'] - 256 / (line['ms_per_step'] * 1e-3)) < 1e-6 * line['value']
  assert line['config'] == bench.workload_config(_Args, bench.WORKLOADS['c1'], 10_000, 256, 1)
  cb = line['cpu_baseline']
  assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == line['value'] and cb['sample']
  assert line['e2e'] == {'value': line['value'], 'unit': 'users/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}

#!/usr/bin/env python
"""Benchmark of the Recoder train-step hot path on B200 (BASELINE.json metric: users/sec of the train step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3]

A step = GPU collate of one batch of users + forward + loss + backward + optimizer update, through the public
`recoder_b200.model.Recoder.train()` call.  Prints ONE JSON line (rank 0):
  value     users/sec with the interaction matrix resident in HBM (whole job, all ranks)
  e2e       the same through `Recoder.train()` with the matrix in HOST memory: every step stages its rows in pinned
            memory, copies them H2D and reads the loss back D2H inside the timed region
  roofline  the dominant kernel of the step against the measured peak (MEASURED_PEAKS.json): `achieved` is timed inside
            the benchmarked multi-stream step, `single_stream` is the same kernel with nothing else on the GPU;
            `traffic` = DRAM bytes per launch from the newest committed ncu summary (profiles/)
  cpu_baseline  the reference's own CPU path on this box's host cores (rank 0, N=1 only): the UNMODIFIED reference
            (`recoder.model.Recoder.train(use_cuda=False)`, imported from baseline/_ref or /root/reference through
            oracle/ref_shims.py; kind "reference") or, when it is not importable, the oracle port (kind "port")
  parity_check  loss of the FIRST timed-run step against the CPU oracle evaluated with batch_size = global batch on the
            same users and initial parameters, and (N>1) whether every rank ended with bit-identical parameters
  kernels / host_ms_per_step / per_rank_ms  CUDA-event breakdown per entry point (separate profiling leg: `value` and
            `e2e` are timed with no per-kernel events), host enqueue vs wait time, per-rank times for N>1
N>1: `--parallel rows` splits the users of the global batch (gradient exchange per `--dp-exchange`: the fused
peer-memory reduce-scatter/Adam/all-gather kernel or one NCCL all-reduce); `--parallel items` splits the item axis
(recoder_b200/itempar.py); `auto` = items for the autoencoder configs (the faster mode at every N measured,
profiles/README.md r02c/r02d), rows for matrix factorisation.
`--impl reference` times that CPU arm alone, on the same config (global batch = per-GPU batch x N; steps of more
than --cpu-max-batch users are sampled at that many users, stated in `cpu_baseline.sample`).
Under torchrun (N>1) one process per GPU, NCCL; timing = CUDA events, max over ranks, barrier + synchronize on
both sides of the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
  # BASELINE.json configs: users, items, nnz/user, model, width, loss, per-GPU batch, activation
  'c1': dict(users=10_000, items=5_000, nnz=50, model='ae', width=128, loss='mse', batch=256,
             desc='C1 synthetic 10Kx5K, 50 nnz/user, AE[128], MSE'),
  'c2': dict(users=138_493, items=26_744, nnz=144, model='ae', width=200, loss='mse', batch=500,
             desc='C2 MovieLens-20M shaped synthetic 138Kx27K, AE[200], MSE'),
  'c3': dict(users=1_000_000, items=200_000, nnz=100, model='ae', width=512, loss='logloss', batch=2048,
             desc='C3 synthetic 1Mx200K, ~100 nnz/user, AE[512], multinomial-NLL + mini-batch negative sampling'),
  'c4': dict(users=1_000_000, items=200_000, nnz=100, model='mf', width=256, loss='mse', batch=2048,
             desc='C4 synthetic 1Mx200K, MatrixFactorization(256), MSE'),
  'c5': dict(users=5_000_000, items=500_000, nnz=100, model='ae', width=1024, loss='logloss', batch=2048,
             desc='C5 synthetic 5Mx500K, AE[1024], multinomial-NLL'),
}
LR = 1e-3


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def load_peaks():
  """Measured roofline denominators (driver-written); fallback figures from B200_PROFILING.md otherwise."""
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as fh:
      p = json.load(fh)
    return dict(hbm=float(p['hbm_gbs']), tensor_burst=float(p['bf16_tflops']),
                tensor_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])), source='measured')
  except Exception:
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source='fallback')


NCU_KERNEL_OF = {'rcd_adam_step': 'k_adam<4>', 'rcd_decoder_fwd_loss': 'k_decoder_fused', 'rcd_adam_step_p2p': 'k_adam_p2p',
                 'rcd_adam_lazy_update': 'k_adam_lazy_update<4>', 'rcd_adam_lazy_catchup': 'k_adam_lazy_catchup<4>',
                 'rcd_decoder_wgrad': 'k_gemm_tc<2>', 'rcd_decoder_dgrad': 'k_gemm_tc<2>'}


def ncu_traffic(entry_point):
  """DRAM bytes (read + write) per launch of the kernel behind `entry_point`, from the newest committed
  `ncu --set full` summary under profiles/ (same workload shape as the default bench); None when there is none."""
  import csv
  import glob
  kernel = NCU_KERNEL_OF.get(entry_point)
  if kernel is None:
    return None, None
  for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_ncu_full_top_kernels.csv')), reverse=True):
    try:
      rows = list(csv.reader(open(path)))
      hdr = rows[0]
      ni, ri, wi = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
      unit = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(rows[1][ri], 1e9)
      vals = [(float(r[ri]) + float(r[wi])) * unit for r in rows[2:] if kernel in r[ni]]
      vals = [v for v in vals if v > 0.5 * max(vals)] if vals else vals   # the big launches (tables, not biases)
      if vals:
        return sum(vals) / len(vals), os.path.basename(path)
    except Exception:
      continue
  return None, None


def make_matrix(w, users_override=None):
  from recoder_b200.synth import synthetic_csr
  U = users_override or w['users']
  t0 = time.time()
  indptr, indices, data = synthetic_csr(U, w['items'], w['nnz'], seed=1234)
  log('[bench] synthetic CSR %d x %d, nnz=%d (%.1fs)' % (U, w['items'], len(indices), time.time() - t0))
  return U, indptr, indices, data


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's step (collate + __compute_loss + backward + optimizer step)
# ---------------------------------------------------------------------------------------------------------------
def cpu_threads():
  try:
    return len(os.sched_getaffinity(0))
  except Exception:
    return os.cpu_count() or 1


def run_cpu_port(w, U, indptr, indices, data, steps, warmup, batch):
  """Times `steps` reference train steps of `batch` users each with the CPU oracle port (kind "port").
  Returns (users_per_sec, ms_per_step, cores, what)."""
  import torch
  from oracle import recoder_oracle as O
  from recoder_b200.synth import epoch_user_order
  cores = cpu_threads()
  torch.set_num_threads(cores)
  I, H = w['items'], w['width']
  if w['model'] == 'ae':
    params = O.init_ae_params(I, [H], seed=0)
    act = 'tanh'
  else:
    params = O.init_mf_params(I, U, H, seed=0)
    act = 'none'
  tr = O.OracleTrainer(w['model'], params, loss=w['loss'], optimizer='adam', lr=LR, weight_decay=0.0, activation=act)
  order = epoch_user_order(U, 1)

  def one_step(s):
    users = order[(s * batch) % max(U - batch, 1):][:batch]
    ob = O.collate(indptr, indices, data, I, users, batch, True)[0]
    tr.step(ob)

  for s in range(warmup):
    one_step(s)
  t0 = time.perf_counter()
  for s in range(steps):
    one_step(s + warmup)
  dt = time.perf_counter() - t0
  what = ('oracle/recoder_oracle.py (CPU port of recoder/data.py collate + model.py __compute_loss + autograd + '
          'torch.optim.Adam: the same torch CPU ops as the reference)')
  return steps * batch / dt, dt / steps * 1e3, cores, what


def run_cpu_reference(w, matrix, steps, warmup, batch):
  """Times the UNMODIFIED reference: `recoder.model.Recoder.train(use_cuda=False)` (recoder/model.py:256) on the same
  matrix, model and hyper-parameters — its own sampler, SciPy/NumPy collate, autograd and torch.optim.Adam (kind
  "reference").  One `train()` call of `warmup` steps, then a timed call of `steps` steps on the same instance.
  Returns (users_per_sec, ms_per_step, cores, what)."""
  import logging
  import torch
  from oracle import ref_shims
  rdata, rnn, _, rmodel = ref_shims.import_reference()
  logging.getLogger('glog').setLevel(logging.WARNING)
  cores = cpu_threads()
  torch.set_num_threads(cores)
  torch.manual_seed(0)
  H = w['width']
  if w['model'] == 'ae':
    model = rnn.DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
  else:
    model = rnn.MatrixFactorization(embedding_size=H, activation_type='none')
  trainer = rmodel.Recoder(model=model, use_cuda=False, optimizer_type='adam', loss=w['loss'])
  ds = rdata.RecommendationDataset(matrix)
  kw = dict(lr=LR, weight_decay=0, num_epochs=1, batch_size=batch, negative_sampling=True, num_data_workers=0)
  trainer.train(ds, iters_per_epoch=max(warmup, 1), **kw)
  t0 = time.perf_counter()
  trainer.train(ds, iters_per_epoch=steps, **kw)
  dt = time.perf_counter() - t0
  what = ('unmodified reference recoder.model.Recoder.train(use_cuda=False) from %s (oracle/ref_shims.py compat '
          'shims only)' % ref_shims.REFERENCE_ROOT)
  return steps * batch / dt, dt / steps * 1e3, cores, what


def cpu_arm(w, U, indptr, indices, data, matrix, steps, warmup, global_batch, max_batch, kind='auto'):
  """The reference's CPU path on this box's host cores; returns the `cpu_baseline` object.  The per-step sample is
  the config's global batch unless that exceeds `max_batch` users (the reference densifies [B, n] fp32 matrices:
  16384 x 199K x 4 B = 13 GB apiece at C3 / 8 GPUs), in which case steps of `max_batch` users are timed and said so."""
  b = min(global_batch, max_batch)
  note = '' if b == global_batch else ' (global batch %d sampled at %d users per step)' % (global_batch, b)
  out = None
  if kind in ('auto', 'reference'):
    try:
      from oracle import ref_shims
      if ref_shims.reference_available():
        ups, ms, cores, what = run_cpu_reference(w, matrix, steps, warmup, b)
        out = ('reference', ups, ms, cores, what)
    except Exception as exc:  # pragma: no cover
      log('[bench/cpu] reference arm failed (%r); falling back to the oracle port' % (exc,))
      if kind == 'reference':
        raise
  if out is None:
    ups, ms, cores, what = run_cpu_port(w, U, indptr, indices, data, steps, warmup, b)
    out = ('port', ups, ms, cores, what)
  k, ups, ms, cores, what = out
  return {'value': ups, 'unit': 'users/s', 'cores': cores, 'kind': k, 'ms_per_step': ms,
          'sample': '%d timed + %d warm-up steps of %d users each on the full matrix%s; %s; torch %d threads' %
                    (steps, warmup, b, note, what, cores)}


def reference_arm(args, w):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  world = int(os.environ.get('WORLD_SIZE', '1'))
  from recoder_b200.synth import to_scipy
  U, indptr, indices, data = make_matrix(w, args.users)
  B = args.batch or w['batch']
  matrix = to_scipy(indptr, indices, data, w['items'])
  cpu = cpu_arm(w, U, indptr, indices, data, matrix, args.steps, args.warmup, B * world, args.cpu_max_batch,
                kind=args.cpu_kind)
  line = {
    'impl': 'reference', 'metric': 'users/sec (train step)', 'value': cpu['value'], 'unit': 'users/s',
    'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cpu['ms_per_step'],
    'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
    'config': workload_config(args, w, U, B, world),
    'cpu_baseline': cpu,
    'e2e': {'value': cpu['value'], 'unit': 'users/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line), flush=True)


def workload_config(args, w, U, batch, world):
  return {'workload': w['desc'], 'users': U, 'items': w['items'], 'nnz_per_user': w['nnz'], 'model': w['model'],
          'width': w['width'], 'loss': w['loss'], 'optimizer': 'adam (dense, torch.optim.Adam semantics)',
          'batch_per_gpu': batch, 'global_batch': batch * world, 'negative_sampling': True,
          'parallelism': ('dp%d' if args.parallel == 'rows' or world == 1 else 'items%d') % world,
          'dp_exchange': args.dp_exchange,
          'l2': l2_statement(args, w, U, batch, world)}


L2_BYTES = 126 * 2 ** 20


def working_set_bytes(args, w, U, batch, world):
  """Bytes one GPU touches per step, estimated: the rows of the embedding tables the step updates with their Adam state
  (p, m, v; the batch's items — all table rows with the dense Adam, so this is the smaller figure — and a 1/world shard
  in item-parallel runs) plus the bf16 dL/dlogits matrix [rows, batch items].  Batch items are estimated as for uniform
  item popularity, I * (1 - exp(-B * nnz / I)).  No flush is issued between steps: `config.l2` states this figure
  against the 126 MB L2."""
  import math
  items_mode = w['model'] == 'ae' and args.parallel != 'rows' and world > 1
  gb = batch * world
  n = int(w['items'] * (1.0 - math.exp(-gb * w['nnz'] / w['items'])))
  rows = 2 * n if w['model'] == 'ae' else n + gb
  shards = world if items_mode else 1
  slice_rows = gb if items_mode else batch
  return rows * w['width'] * 12 // shards + slice_rows * (n // shards) * 2


def l2_statement(args, w, U, batch, world):
  ws = working_set_bytes(args, w, U, batch, world)
  if ws > L2_BYTES:
    return ('no flush: per-step working set (touched embedding rows + Adam state + logits, about %d MB per GPU) is larger than '
            'the 126 MB L2' % (ws >> 20))
  return ('no flush: per-step working set (touched embedding rows + Adam state + logits, about %d MB) FITS in the 126 MB L2 — '
          'a parity / host-path configuration, not the one the metric is quoted on (C3)' % (ws >> 20))


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
  FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
            'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.gpu = gpu_index
    self.proc = None

  def start(self):
    if self.gpu is None:
      return
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits',
                                    '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
    except Exception:
      self.proc = None

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      out, _ = self.proc.communicate(timeout=5)
    except Exception:
      self.proc.kill()
      out = ''
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for ln in out.strip().splitlines():
      parts = [p.strip() for p in ln.split(',')]
      if len(parts) < 8:
        continue
      try:
        sm.append(float(parts[1]))
        mx.append(float(parts[2]))
      except ValueError:
        continue
      for nm, val in zip(names, parts[4:8]):
        if val.lower().startswith('active'):
          reasons.add(nm)
    return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def kernel_work(name, w, rows, n, n_in, nnz_rows, tables):
  """Algorithmic work of ONE optimizer step for entry point `name`: ('hbm', bytes) or ('tensor', flops).
  Figures are the per-unit numbers of DESIGN.md §4."""
  H, I = w['width'], w['items']
  dense = 2.0 * rows * n * H
  if name in ('rcd_decoder_fwd', 'rcd_decoder_fwd_loss', 'rcd_decoder_dgrad', 'rcd_decoder_wgrad'):
    return 'tensor', dense
  if name == 'rcd_adam_step':
    params, grads = tables[0], tables[1]
    return 'hbm', 24.0 * params + 4.0 * grads
  if name == 'rcd_adam_lazy_update':
    # deferred dense Adam: the batch's rows only — read p, m, v + the gradient row, write p, m, v = 28 B/param
    grads = tables[1]
    return 'hbm', 28.0 * grads
  if name == 'rcd_adam_step_p2p':
    # per rank: NVLink ingress = the other ranks' gradient rows of the owned shard + the other ranks' pushed rows
    params, grads, world = tables if len(tables) == 3 else (tables[0], tables[1], 1)
    return 'nvlink', 4.0 * (grads + params) * (world - 1) / max(world, 1)
  # HBM-bound kernels: ALGORITHMIC DRAM bytes.  The embedding-row gathers of the sparse kernels touch
  # nnz rows but only the n DISTINCT rows of the batch have to come from DRAM (repeats are L2 hits), and the [rows, H]
  # activations they re-read (Z, dA: a few MB) live in L2 — counting those as HBM traffic is what made r01's
  # fractions exceed 1.
  if name == 'rcd_gather_rows':
    return 'hbm', n * H * (4.0 + 2.0)       # fp32 master rows in, bf16 operand out
  if name == 'rcd_ae_encoder_fwd':
    return 'hbm', n_in * H * 4.0 + rows * H * 6.0 + nnz_rows * 8.0
  if name == 'rcd_ae_encoder_wgrad':
    return 'hbm', n_in * H * 4.0 + rows * H * 4.0 + nnz_rows * 8.0
  if name == 'rcd_csc_rows_accumulate':
    return 'hbm', 2 * n * H * 4.0 + rows * H * 4.0 + nnz_rows * 12.0   # read-modify-write of dW_d rows
  if name == 'rcd_sparse_dgrad':
    return 'hbm', n * H * 4.0 + rows * H * 4.0 + nnz_rows * 8.0
  return 'hbm', None


def parity_check(w, U, indptr, indices, data, global_batch, gpu_first_loss):
  """Loss of the first step of the benchmarked run against the CPU oracle: same users (the first `global_batch`
  entries of epoch 1's order), same initial parameters (the model initialised under torch.manual_seed(0), as every
  rank's replica is), the reference's semantics with batch_size = global batch.  Forward only, over
  row chunks (oracle.loss_in_row_chunks): a 16384 x 199K batch is never densified at once."""
  import torch
  from oracle import recoder_oracle as O
  from recoder_b200.nn import DynamicAutoencoder, MatrixFactorization
  from recoder_b200.synth import epoch_user_order
  if gpu_first_loss is None:
    return {'error': 'first-step loss not recorded'}
  torch.set_num_threads(cpu_threads())     # (torchrun exports OMP_NUM_THREADS=1 to every rank)
  I, H = w['items'], w['width']
  torch.manual_seed(0)
  if w['model'] == 'ae':
    model, act = DynamicAutoencoder(hidden_layers=[H], activation_type='tanh'), 'tanh'
  else:
    model, act = MatrixFactorization(embedding_size=H, activation_type='none'), 'none'
  model.init_model(num_items=I, num_users=U)
  params = {k: v.detach().cpu() for k, v in model.named_parameters()}
  tr = O.OracleTrainer(w['model'], params, loss=w['loss'], optimizer='adam', lr=LR, weight_decay=0.0, activation=act)
  users = epoch_user_order(U, 1)[:global_batch]
  ob = O.collate(indptr, indices, data, I, users, global_batch, True)[0]
  t0 = time.perf_counter()
  want = tr.loss_in_row_chunks(ob, 2048)
  rel = abs(gpu_first_loss - want) / max(abs(want), 1e-30)
  return {'first_step_loss': gpu_first_loss, 'oracle_first_step_loss': want, 'rel_err': rel, 'tolerance': 1e-3,
          'ok': bool(rel <= 1e-3), 'oracle': 'oracle/recoder_oracle.py forward with batch_size = %d (global batch), '
          '%d items, %.1f s on the host' % (global_batch, ob.size[1], time.perf_counter() - t0)}


def b200_arm(args, w):
  import torch
  import torch.distributed as dist
  from recoder_b200 import _native
  from recoder_b200 import data as rdata
  from recoder_b200.data import RecommendationDataset
  from recoder_b200.model import Recoder
  from recoder_b200.nn import DynamicAutoencoder, MatrixFactorization
  from recoder_b200.synth import epoch_user_order, to_scipy

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus:
    log('[bench] warning: --gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1); using %d' %
        (args.gpus, world, world))
  torch.cuda.set_device(local_rank)
  if world > 1:
    # One process per GPU on a shared host: give every rank its own slice of the cores.  Unpinned, six of eight ranks
    # spent 5 ms per step in the host staging of the e2e leg against 1.3 ms on the other two (profiles/README.md r02t):
    # their main threads share cores with the other ranks' busy-polling NCCL proxy threads.
    try:
      cores = sorted(os.sched_getaffinity(0))
      per = len(cores) // world
      if per >= 2 and os.environ.get('RCD_PIN_RANKS', '1') != '0':
        os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
        os.environ.setdefault('RCD_STAGE_THREADS', str(max(1, min(4, per - 1))))
    except (AttributeError, OSError):
      pass
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
  lib = _native.load()
  peaks = load_peaks()
  K, W = args.steps, max(args.warmup, 3)
  B = args.batch or w['batch']
  U, indptr, indices, data = make_matrix(w, args.users)
  I, H = w['items'], w['width']
  matrix = to_scipy(indptr, indices, data, I)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def collect_timings(trainer):
    """{entry point: (total device ms, launches)} since the last call: CUDA-event pairs of the ctypes path
    (`_native.TIMINGS`: collate, staging, Python-path steps) plus those the native step executor recorded itself."""
    torch.cuda.synchronize()
    out = {}
    for name, evs in _native.TIMINGS.items():
      out[name] = (sum(a.elapsed_time(b) for a, b in evs), len(evs))
    _native.TIMINGS.clear()
    nat = getattr(trainer.engine, '_native', None)
    if nat is not None:
      for name, (tot, cnt) in nat.read_profile().items():
        t0, c0 = out.get(name, (0.0, 0))
        out[name] = (t0 + tot, c0 + cnt)
    return out

  def run(device_resident, sync_loss, profile, overlap=True, K=K):
    """One `Recoder.train()` call of W+K steps; returns (elapsed_ms max over ranks, stats)."""
    prev_overlap = os.environ.get('RCD_OVERLAP')
    if not overlap:
      os.environ['RCD_OVERLAP'] = '0'     # read by the engine / trainer when they are constructed below
    torch.manual_seed(0)
    if w['model'] == 'ae':
      model = DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
    else:
      model = MatrixFactorization(embedding_size=H, activation_type='none')
    trainer = Recoder(model=model, use_cuda=True, optimizer_type='adam', loss=w['loss'], dp_exchange=args.dp_exchange,
                      parallel=args.parallel)
    ds = RecommendationDataset(matrix, device_resident=device_resident)
    st = {'n': [], 'launch0': 0, 'launch1': 0, 'bytes0': None, 'bytes1': None, 'clocks': None, 'warm': {},
          'dominant': None}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank if rank == 0 else None)   # one nvidia-smi poller per job, not per rank
    if profile:
      _native.PROFILE = 'all'
      _native.TIMINGS.clear()

    def cb(step):
      if step > W:
        st['n'].append(trainer.engine.last.get('n', 0))
      if step == W:
        torch.cuda.synchronize()
        if profile:
          # per entry point: mean device time per step over the warm-up steps
          warm = {k: tot / max(W, 1) for k, (tot, cnt) in collect_timings(trainer).items()}
          st['warm'] = warm
          cand = {k: v for k, v in warm.items() if k not in ('rcd_collate', 'rcd_p2p_barrier')}
          st['dominant'] = max(cand, key=cand.get) if cand else None
          _native.PROFILE = {st['dominant']} if st['dominant'] else None
        barrier()
        sampler.start()
        st['launch0'] = lib.rcd_launch_count()
        st['bytes0'] = dict(rdata.TRANSFER_BYTES)
        trainer.engine.join()
        ev0.record()
        st['host_t0'] = time.perf_counter()
        ht = getattr(trainer, '_host_timing', None)
        if ht is not None:     # host-side accounting covers the timed region only
          for k in ht:
            ht[k] = 0
      elif step == W + K:
        st['host_enqueue_ms'] = (time.perf_counter() - st['host_t0']) * 1e3 / K   # host time to ENQUEUE one step
        trainer.engine.join()   # the optimizer / exchange kernels of the last step run on the update stream
        ev1.record()
        torch.cuda.synchronize()
        st['launch1'] = lib.rcd_launch_count()
        st['bytes1'] = dict(rdata.TRANSFER_BYTES)
        st['clocks'] = sampler.stop()
        barrier()

    # small matrices (C1: 39 full batches per pass) take several passes; every pass uses full batches only
    per_pass = U // (B * world)
    assert per_pass >= 1, 'batch larger than the matrix'
    if W + K <= per_pass:
      epochs, iters = 1, W + K
    else:
      epochs, iters = -(-(W + K) // per_pass), None
    trainer.train(ds, lr=LR, weight_decay=0, num_epochs=epochs, iters_per_epoch=iters, batch_size=B,
                  negative_sampling=True, user_order=lambda e: epoch_user_order(U, e)[:per_pass * B * world],
                  step_callback=cb, sync_loss_every_step=sync_loss)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    ht = getattr(trainer, '_host_timing', None)
    host = {k: round(v * 1e3 / max(ht['n'], 1), 4) for k, v in ht.items() if k != 'n'} if ht else {}
    st['host'] = host
    if world > 1:
      # per-rank view of the timed region (device ms/step, host ms/step enqueueing steps, enqueueing the next collate,
      # blocked on the GPU): a rank whose host column approaches the device column is what the others wait for
      mine = torch.tensor([ms / K, host.get('step', 0.0), host.get('launch', 0.0), host.get('wait', 0.0)],
                          device='cuda', dtype=torch.float64)
      every = [torch.zeros_like(mine) for _ in range(world)]
      dist.all_gather(every, mine)
      st['per_rank'] = [[round(float(x), 4) for x in t.tolist()] for t in every]
      if args.per_rank_kernels and st.get('warm'):
        # diagnostic (off by default): every rank's warm-up breakdown, to find the rank / entry point the others wait
        # for at the barriers of the multi-GPU modes
        warms = [None] * world
        dist.all_gather_object(warms, {k: round(v, 4) for k, v in st['warm'].items()})
        st['per_rank_kernels'] = warms
      t = torch.tensor([ms], device='cuda', dtype=torch.float64)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    st['dom_ms'] = None
    if profile and st['dominant']:
      tot, cnt = collect_timings(trainer).get(st['dominant'], (0.0, 0))
      if cnt:
        st['dom_ms'] = tot / cnt
        st['dom_launches_per_step'] = cnt / K
    _native.PROFILE = None
    _native.TIMINGS.clear()
    done = trainer.engine.steps_done      # (a run over several passes of a small matrix rounds up to whole passes)
    all_losses = trainer.engine.losses(done)
    st['loss'] = float(all_losses[-1])
    st['first_loss'] = float(all_losses[0]) if len(all_losses) == done else None
    st['params'] = sum(p.numel() for p in model.parameters())
    if world > 1 and not profile:
      # every rank must hold bit-identical parameters after the run (item-parallel: after gathering the shards)
      trainer.sync_parameters()
      torch.cuda.synchronize()
      sig = torch.stack([torch.stack([p.data.double().sum(), p.data.double().abs().sum(),
                                      (p.data.double().flatten()[::7]).sum()]) for p in model.parameters()])
      sigs = [torch.zeros_like(sig) for _ in range(world)]
      dist.all_gather(sigs, sig)
      st['replicas_identical'] = bool(all(torch.equal(sigs[0], t) for t in sigs[1:]))
    del trainer, model, ds
    torch.cuda.empty_cache()
    if prev_overlap is None:
      os.environ.pop('RCD_OVERLAP', None)
    else:
      os.environ['RCD_OVERLAP'] = prev_overlap
    return ms, st

  # ---- leg 1: matrix resident in HBM (value) — no per-kernel events anywhere in this run -----------------------------
  ms_dev, s_dev = run(device_resident=True, sync_loss=False, profile=False)
  # ---- leg 1p: the same run with CUDA events around every entry point during warm-up (breakdown) and around the
  # dominant one during its timed steps (roofline `achieved`: the kernel inside the multi-stream step) -----------------
  s_prof = None
  if not args.no_profile:
    _, s_prof = run(device_resident=True, sync_loss=False, profile=True, K=max(6, min(K, 20)))
  # ---- leg 1b: the dominant kernel alone — in the step the optimizer kernels share the GPU with the dgrad GEMM and the
  # encoder backward (update stream), which stretches their event-timed duration; a few more steps on ONE stream give
  # the kernel's own launch duration for the roofline (both figures are reported)
  iso = None
  if s_prof is not None and world == 1 and s_prof.get('dominant'):
    _, s_iso = run(device_resident=True, sync_loss=False, profile=True, overlap=False, K=max(6, min(K, 10)))
    if s_iso.get('dominant') == s_prof['dominant'] and s_iso.get('dom_ms'):
      iso = s_iso
  # ---- leg 2: host-resident matrix, H2D staging + loss readback every step (e2e) ------------------------------
  if args.skip_e2e:
    ms_e2e, s_e2e = ms_dev, {'bytes0': {'h2d': 0, 'd2h': 0}, 'bytes1': {'h2d': 0, 'd2h': 0}}
  else:
    ms_e2e, s_e2e = run(device_resident=False, sync_loss=True, profile=False)
  if s_prof is None:
    s_prof = {'warm': {}, 'dominant': None, 'dom_ms': None}

  users_per_step = B * world
  value = K * users_per_step / (ms_dev / 1e3)
  e2e = K * users_per_step / (ms_e2e / 1e3)
  n_avg = float(np.mean(s_dev['n'])) if s_dev['n'] else 0.0
  nnz_rows = float(B * (len(indices) / U))

  # ---- roofline of the dominant kernel ------------------------------------------------------------------------
  dom = s_prof['dominant']
  n_tab = 2 if w['model'] == 'ae' else 1
  if args.parallel == 'items' and world > 1:
    s_dev['params'] = s_dev['params'] / world    # each rank owns (and updates) 1/world of the item-indexed tensors
  if w['model'] == 'ae':
    grads = 2 * n_avg * H + n_avg + H
  else:
    grads = n_avg * H + n_avg + users_per_step * H
  items_mode = args.parallel == 'items' and world > 1
  rows_k = B * world if items_mode else B     # item-parallel: every rank runs all rows over its item shard
  nnz_k = nnz_rows if items_mode else nnz_rows  # (all rows x 1/world of the columns == one rank's rows)
  kinds = {}
  for name, ms in sorted(s_prof['warm'].items(), key=lambda kv: -kv[1]):
    bound, work = kernel_work(name, w, rows_k, n_avg, n_avg, nnz_k, (s_dev['params'], grads, world))
    entry = {'ms_per_step': round(ms, 4), 'bound': bound}
    if work:
      if bound == 'nvlink':
        entry['achieved_gbs_ingress_per_gpu'] = round(work / (ms * 1e-3) / 1e9, 1)
        entry['frac_of_770_gbs_peer_copy'] = round(entry['achieved_gbs_ingress_per_gpu'] / 770.0, 4)
      elif bound == 'tensor':
        entry['achieved_tflops'] = round(work / (ms * 1e-3) / 1e12, 2)
        entry['frac_of_measured_sustained'] = round(entry['achieved_tflops'] / peaks['tensor_sustained'], 4)
      else:
        entry['achieved_gbs'] = round(work / (ms * 1e-3) / 1e9, 1)
        entry['frac_of_measured'] = round(entry['achieved_gbs'] / peaks['hbm'], 4)
    kinds[name] = entry
  roofline = None
  if dom and s_prof['dom_ms']:
    bound, work = kernel_work(dom, w, rows_k, n_avg, n_avg, nnz_k, (s_dev['params'], grads, world))
    lps = s_prof.get('dom_launches_per_step', 1.0)
    if work:
      per_launch = work / lps
      sec = s_prof['dom_ms'] * 1e-3
      if bound == 'tensor':
        ach, peak, unit = per_launch / sec / 1e12, peaks['tensor_sustained'], 'TFLOP/s'
      elif bound == 'nvlink':
        ach, peak, unit = per_launch / sec / 1e9, 900.0, 'GB/s'
      else:
        ach, peak, unit = per_launch / sec / 1e9, peaks['hbm'], 'GB/s'
      # the committed ncu summaries are captures of the default workload (C3, 2048 users per GPU)
      traffic, traffic_src = ncu_traffic(dom) if (args.config == 'c3' and B == WORKLOADS['c3']['batch'] and
                                                  world == 1) else (None, None)
      if traffic is not None and dom == 'rcd_adam_step':
        # the ncu figure is the mean over the table launches; a step also has the (KB-sized) bias launches, and
        # `achieved` averages over all `lps` launches of a step — put both on the same per-launch footing
        traffic = traffic * n_tab / lps
      if iso is not None:
        sec_iso = iso['dom_ms'] * 1e-3
        ach_iso = per_launch / sec_iso / (1e12 if bound == 'tensor' else 1e9)
      roofline = {'kernel': dom, 'bound': bound, 'achieved': round(ach, 2),
                  'peak': peak, 'unit': unit, 'frac': round(ach / peak, 4), 'traffic': traffic,
                  'traffic_source': traffic_src, 'algorithmic_per_launch': per_launch,
                  'peak_source': ('NVLink 5 nominal per direction per GPU (B200_PROFILING.md: 770 GB/s measured peer '
                                  'copy); ingress bytes of this rank' if bound == 'nvlink' else
                                  peaks['source'] + (' (sustained)' if bound == 'tensor' else '')),
                  'launches_per_step': lps, 'avg_launch_ms': round(s_prof['dom_ms'], 4)}
      if iso is not None:
        # `achieved` above is timed inside the benchmarked (multi-stream) region; this is the same kernel with nothing
        # else on the GPU
        roofline['single_stream'] = {'achieved': round(ach_iso, 2), 'frac': round(ach_iso / peak, 4),
                                     'avg_launch_ms': round(iso['dom_ms'], 4)}

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- parity of the benchmarked run: loss of its FIRST step against the CPU oracle with batch_size = global batch
  # (the reference's step on the same users, forward only, evaluated over row chunks), replicas bit-identical ---------
  parity = None
  # the oracle forward costs 4*B*n*H flops on the host cores while every GPU of the box sits idle: beyond
  # --parity-max-tflop it is skipped (and said so) unless --parity-check full
  est_tflop = 4.0 * users_per_step * (np.mean(s_dev['n']) * (world if args.parallel == 'items' and world > 1 else 1)
                                      if s_dev['n'] else 0.0) * H / 1e12
  if not args.no_parity_check and args.parity_check != 'full' and est_tflop > args.parity_max_tflop:
    parity = {'skipped': 'oracle forward of %.1f TFLOP on the host exceeds --parity-max-tflop %.1f (run with '
                         '--parity-check full)' % (est_tflop, args.parity_max_tflop),
              'replicas_identical': s_dev.get('replicas_identical')}
  elif not args.no_parity_check:
    try:
      parity = parity_check(w, U, indptr, indices, data, users_per_step, s_dev.get('first_loss'))
      if world > 1:
        parity['replicas_identical'] = s_dev.get('replicas_identical')
    except Exception as exc:  # pragma: no cover
      parity = {'error': repr(exc)}

  # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------------
  cpu = None
  if world == 1 and not args.no_cpu_baseline:
    try:
      cpu = cpu_arm(w, U, indptr, indices, data, matrix, 2, 1, B, args.cpu_max_batch, kind=args.cpu_kind)
    except Exception as exc:  # pragma: no cover
      cpu = {'value': None, 'unit': 'users/s', 'cores': cpu_threads(), 'kind': 'port', 'sample': 'failed: %r' % exc}

  h2d = (s_e2e['bytes1']['h2d'] - s_e2e['bytes0']['h2d']) / K
  d2h = (s_e2e['bytes1']['d2h'] - s_e2e['bytes0']['d2h']) / K
  line = {
    'metric': 'users/sec (train step)', 'value': value, 'unit': 'users/s', 'n_gpus': world, 'steps': K, 'warmup': W,
    'ms_per_step': ms_dev / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': 'bf16 tensor-core operands, fp32 accumulate / master weights / optimizer', 'data': 'synthetic',
    'config': workload_config(args, w, U, B, world),
    'e2e': {'value': e2e, 'unit': 'users/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
            'ms_per_step': ms_e2e / K},
    'gpu_launches': int(s_dev['launch1'] - s_dev['launch0']),
    'clocks': s_dev['clocks'],
    'roofline': roofline,
    'cpu_baseline': cpu,
    'parity_check': parity,
    'items_per_batch': n_avg,
    'host_ms_per_step': s_dev.get('host'),   # timed region; wait = blocked on the GPU, launch / step = enqueue work
    'e2e_host_ms_per_step': s_e2e.get('host'), 'e2e_per_rank_ms': s_e2e.get('per_rank'),
    'per_rank_ms': s_dev.get('per_rank'),    # N>1: [device ms/step, host step, host launch, host wait] per rank
    'per_rank_kernels': s_dev.get('per_rank_kernels'),
    'final_loss': s_dev['loss'],
    'kernels': kinds,
  }
  print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=100)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--config', default='c3', choices=sorted(WORKLOADS))
  ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the config\'s)')
  ap.add_argument('--users', type=int, default=0, help='use a user prefix of the matrix (0 = all)')
  ap.add_argument('--dp-exchange', default='auto', choices=['auto', 'p2p', 'nccl'],
                  help='N>1: fused peer-memory reduce-scatter/Adam/all-gather kernel (p2p) or NCCL all-reduce + full Adam')
  ap.add_argument('--parallel', default='auto', choices=['auto', 'rows', 'items'],
                  help='N>1: split the users of the global batch (data parallel, gradient exchange per --dp-exchange) '
                       'or the item axis (itempar.py); auto = items for the autoencoder configs, rows for matrix factorisation')
  ap.add_argument('--per-rank-kernels', action='store_true',
                  help='N>1 diagnostic: gather every rank\'s per-entry-point warm-up timings into the JSON line')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--skip-e2e', action='store_true', help='profiling runs only: skip the host-staged leg')
  ap.add_argument('--no-profile', action='store_true', help='no CUDA-event kernel breakdown during warm-up')
  ap.add_argument('--cpu-max-batch', type=int, default=4096,
                  help='CPU arm: largest per-step user sample (the reference densifies [B, n] fp32 matrices)')
  ap.add_argument('--cpu-kind', default='auto', choices=['auto', 'reference', 'port'],
                  help='CPU arm: the unmodified reference (needs baseline/_ref or /root/reference) or the oracle port')
  ap.add_argument('--no-parity-check', action='store_true', help='skip the first-step loss check against the oracle')
  ap.add_argument('--parity-check', default='auto', choices=['auto', 'full'],
                  help='auto: skip the oracle forward when it exceeds --parity-max-tflop; full: always run it')
  ap.add_argument('--parity-max-tflop', type=float, default=12.0)
  args = ap.parse_args()
  w = WORKLOADS[args.config]
  if args.parallel == 'auto':
    # measured on C3 (profiles/README.md r02c / r02d): with the native step executor the item-parallel mode runs at
    # 1.69 M users/s on 2 GPUs and 6.71 M on 8 (one GPU: 0.85-0.88 M) against 1.35 M / 4.34 M for row-parallel + fused
    # peer-memory exchange, so it is the autoencoder default at every N; matrix factorisation has no item-parallel
    # mode (its user table is indexed by row) and uses the row-parallel exchange
    args.parallel = 'items' if w['model'] == 'ae' else 'rows'
  if args.impl == 'reference':
    reference_arm(args, w)
  else:
    b200_arm(args, w)


if __name__ == '__main__':
  main()

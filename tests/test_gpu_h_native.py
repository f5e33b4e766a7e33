"""The native step executor (`rcd_step_run`, include/recoder_b200.h "K12") against the Python launch sequence of
`TrainEngine`: the same entry points in the same order on the same streams — so parameters, optimizer state, losses and
gradients must be BIT-identical — and against the CPU oracle through the public `Recoder.train()` call."""
import time

import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200.data import RecommendationDataset, collate_pool
from recoder_b200.model import Recoder
from recoder_b200.nn import DynamicAutoencoder
from recoder_b200.synth import epoch_user_order, synthetic_csr, to_scipy
from tests.gpu_util import device_dataset, make_engine, make_model

pytestmark = pytest.mark.gpu

CASES = [
  # kind, U, I, nnz, H, B, loss, act, optimizer, negative sampling
  ('ae', 3000, 5000, 50, 128, 256, 'mse', 'tanh', 'adam', True),
  ('ae', 3000, 26744, 144, 200, 500, 'logloss', 'tanh', 'adam', True),
  ('ae', 2000, 3001, 30, 72, 333, 'logistic', 'sigmoid', 'sgd', True),
  ('ae', 1500, 2000, 40, 64, 200, 'mse', 'relu', 'rmsprop', False),
  ('ae', 1500, 2000, 40, 64, 200, 'logloss', 'tanh', 'adagrad', True),
  ('mf', 3000, 20000, 100, 256, 512, 'mse', 'none', 'adam', True),
  ('mf', 2000, 3000, 30, 40, 128, 'logloss', 'tanh', 'sgd', True),
  ('ae', 9000, 4000, 60, 96, 5000, 'logloss', 'tanh', 'adam', True),   # > 4096 rows: chunked heavy columns
]


def _run(kind, U, I, nnz, H, B, loss, act, opt, neg, native, steps=4):
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=21)
  if kind == 'ae':
    params = O.init_ae_params(I, [H], seed=4)
  else:
    params = O.init_mf_params(I, U, H, seed=4)
  model = make_model(kind, I, U, [H] if kind == 'ae' else H, act, {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, loss, 0.5 if loss == 'mse' else 0.0, opt, 1e-3, 1e-4, _native.GEMM_TCGEN05)
  eng.native_enabled = native
  ds = device_dataset(indptr, indices, data, I)
  order = np.random.default_rng(3).permutation(U)
  grads = []
  for s in range(steps):
    users = order[(s * B) % max(U - B, 1):][:B]
    pool = collate_pool(ds.device_csr(), users, neg)
    eng.train_step(pool, 0, len(users))
    grads.append({k: v.detach().clone() for k, v in eng.last.items() if torch.is_tensor(v) and v.numel()})
  val = eng.eval_loss(collate_pool(ds.device_csr(), order[:B], neg), 0, B)
  losses = eng.losses(steps).clone()
  state = {n: p.detach().clone() for n, p in model.named_parameters()}
  opt_state = {n: (st.m.clone(), None if st.v is None else st.v.clone()) for n, st in eng.opt.states.items()}
  used_native = eng._native is not None
  return losses, val, state, opt_state, grads, used_native


@pytest.mark.parametrize('kind,U,I,nnz,H,B,loss,act,opt,neg', CASES)
def test_native_step_is_bit_identical_to_python_path(kind, U, I, nnz, H, B, loss, act, opt, neg):
  a = _run(kind, U, I, nnz, H, B, loss, act, opt, neg, native=True)
  b = _run(kind, U, I, nnz, H, B, loss, act, opt, neg, native=False)
  assert a[5] and not b[5]
  assert torch.equal(a[0], b[0]), (a[0], b[0])
  assert a[1] == b[1]
  for n in a[2]:
    assert torch.equal(a[2][n], b[2][n]), n
  for n in a[3]:
    assert torch.equal(a[3][n][0], b[3][n][0]), n
    if a[3][n][1] is not None:
      assert torch.equal(a[3][n][1], b[3][n][1]), n
  for ga, gb in zip(a[4], b[4]):
    assert set(ga) == set(gb)
    for k in ga:
      assert torch.equal(ga[k], gb[k]), k


def test_native_and_python_steps_can_alternate():
  """Mixed use (e.g. an evaluation pass on the Python path between native training steps) keeps stream order."""
  U, I, H, B = 2000, 4000, 64, 256
  indptr, indices, data = synthetic_csr(U, I, 40, seed=2)
  params = O.init_ae_params(I, [H], seed=1)
  runs = []
  for pattern in ([True] * 6, [True, False, True, True, False, True]):
    model = make_model('ae', I, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
    eng = make_engine(model, 'logloss', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
    ds = device_dataset(indptr, indices, data, I)
    for s, nat in enumerate(pattern):
      eng.native_enabled = nat
      eng.train_step(collate_pool(ds.device_csr(), np.arange(s * B, (s + 1) * B), True), 0, B)
    runs.append((eng.losses(6).clone(), {n: p.detach().clone() for n, p in model.named_parameters()}))
  assert torch.equal(runs[0][0], runs[1][0])
  for n in runs[0][1]:
    assert torch.equal(runs[0][1][n], runs[1][1][n]), n


def test_recoder_train_native_matches_oracle_and_reports_host_time():
  """`Recoder.train()` on the native executor at C1's shape: the loss curve follows the oracle, and the host time per
  step (printed) is what the executor is for."""
  U, I, H, B, steps = 10_000, 5_000, 128, 256, 39
  indptr, indices, data = synthetic_csr(U, I, 50, seed=1234)
  ds = RecommendationDataset(to_scipy(indptr, indices, data, I))
  order = epoch_user_order(U, 1)
  # the trainer initialises its model under the global seed inside train(); the same seed gives the oracle's start
  torch.manual_seed(0)
  probe = DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
  probe.init_model(num_items=I, num_users=U)
  init = {k: v.detach().clone() for k, v in probe.named_parameters()}
  model2 = DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
  trainer2 = Recoder(model=model2, use_cuda=True, optimizer_type='adam', loss='mse')
  torch.manual_seed(0)
  t0 = time.perf_counter()
  trainer2.train(ds, lr=1e-3, weight_decay=0, num_epochs=1, iters_per_epoch=steps, batch_size=B,
                 negative_sampling=True, user_order=lambda e: order)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  assert trainer2.engine._native is not None
  tr = O.OracleTrainer('ae', {k: v.cpu() for k, v in init.items()}, loss='mse', optimizer='adam', lr=1e-3,
                       activation='tanh')
  want = []
  for s in range(steps):
    ob = O.collate(indptr, indices, data, I, order[s * B:(s + 1) * B], B, True)[0]
    want.append(tr.step(ob)[0])
  got = trainer2.last_epoch_losses
  rel = np.abs(got - np.asarray(want)) / np.abs(want)
  print('Recoder.train native: %d steps in %.1f ms (%.3f ms/step incl. set-up), max rel err %.2e' %
        (steps, dt * 1e3, dt * 1e3 / steps, rel.max()))
  assert rel.max() < 2e-3

"""Parity at the BENCHMARKED shapes and over un-teacher-forced training runs (SURVEY.md §8d gates ii-iii).

* one step at C3's real shape (I = 200K, H = 512, B = 2048, multinomial NLL) and at C5's (I = 500K, H = 1024, B = 2048)
  against the CPU oracle: loss and gradient L2 norms within 1e-3 (north_star), item bookkeeping bit-exact;
* C5's largest sweep point (B = 8192, n ~ 359K) through a size-independent property — the gradient of a batch is the
  sum of the gradients of its row blocks (same item set, same 1/B): four 2048-row steps must add up to the 8192-row step;
* loss curves: >= 200 optimizer steps at C1's shape and 60 at C2's, both sides free-running from the same initial
  parameters and user order (no teacher forcing), compared step by step;
* a model that represents more items than the matrix has columns (reference recoder/model.py:241 allows it), and a
  matrix wider than the model (an error).
"""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200.data import collate_pool
from recoder_b200.synth import synthetic_csr
from tests.gpu_util import compact_oracle_grads, device_dataset, make_engine, make_model, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star: loss and gradient L2 norms, relative
TOL_ELEM = 2e-2     # relative Frobenius distance of whole gradient blocks (bf16 operands vs fp32)


def _oracle_loss_and_grads(tr, ob):
  """__compute_loss + backward of the reference (model.py:395-397) without the optimizer step (keeps the host memory
  of the 500K x 1024 case at parameters + gradients)."""
  for t in tr.params.values():
    t.grad = None
  loss = tr.compute_loss(ob)
  loss.backward()
  return float(loss.item()), {k: v.grad.detach() for k, v in tr.params.items()}


@pytest.mark.parametrize('name,I,H,B,nnz', [('c3', 200_000, 512, 2048, 100), ('c5', 500_000, 1024, 2048, 100)])
def test_step_at_benchmarked_shape_matches_oracle(name, I, H, B, nnz):
  U = 4096
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=1234)     # the bench's generator and seed (a user prefix)
  params = O.init_ae_params(I, [H], seed=0)
  g = torch.Generator().manual_seed(7)
  params[O.AE_EN_B] = torch.randn(H, generator=g) * 0.05
  params[O.AE_DE_B] = torch.randn(I, generator=g) * 0.05
  tr = O.OracleTrainer('ae', params, loss='logloss', optimizer='adam', lr=1e-3, activation='tanh')
  model = make_model('ae', I, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, 'logloss', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
  ds = device_dataset(indptr, indices, data, I)
  users = np.random.default_rng(1).permutation(U)[:B]
  pool = collate_pool(ds.device_csr(), users, True)
  ob = O.collate(indptr, indices, data, I, users, B, True)[0]
  assert np.array_equal(pool.items.cpu().numpy(), ob.items)                    # bit-exact item bookkeeping
  assert np.array_equal(pool.cols[:pool.nnz].cpu().numpy(), ob.indices[1].astype(np.int32))
  oloss, ograds = _oracle_loss_and_grads(tr, ob)
  eng.train_step(pool, 0, B)
  gloss = float(eng.losses(1)[0])
  assert gloss == pytest.approx(oloss, rel=TOL), name
  want = compact_oracle_grads('ae', ograds, ob)
  for key, wgt in want.items():
    got = eng.last[key].detach().cpu().numpy()
    assert np.linalg.norm(got) == pytest.approx(np.linalg.norm(wgt), rel=TOL), '%s %s norm' % (name, key)
    assert rel_err(got, wgt) < TOL_ELEM, '%s %s' % (name, key)


def test_c5_largest_batch_gradient_is_sum_of_row_blocks():
  I, H, B, U = 500_000, 1024, 8192, 8192
  indptr, indices, data = synthetic_csr(U, I, 100, seed=1234)
  params = O.init_ae_params(I, [H], seed=0)
  model = make_model('ae', I, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
  del params
  eng = make_engine(model, 'logloss', 0.0, 'adam', 0.0, 0.0, _native.GEMM_TCGEN05)   # lr = 0: parameters stay put
  ds = device_dataset(indptr, indices, data, I)
  pool = collate_pool(ds.device_csr(), np.arange(U), True)
  n = pool.n
  assert n > 300_000
  before = model.en_embedding_layer.weight.data[:1000].clone()
  eng.train_step(pool, 0, B)
  full_loss = float(eng.losses(1)[0])
  full = {k: eng.last[k].detach().clone() for k in ('dWe', 'dWd', 'dbd', 'dbe')}
  acc = {k: torch.zeros_like(v, dtype=torch.float64) for k, v in full.items()}
  loss_sum = 0.0
  for row0 in range(0, B, 2048):
    eng.train_step(pool, row0, 2048, global_rows=B)
    loss_sum += float(eng.losses(1)[0])
    for k in acc:
      acc[k] += eng.last[k].detach().double()
  assert torch.equal(model.en_embedding_layer.weight.data[:1000], before)
  assert loss_sum == pytest.approx(full_loss, rel=1e-5)
  for k in acc:
    # same bf16 operands and the same per-row softmax statistics on both sides: only the fp32 summation order differs
    assert rel_err(acc[k].cpu().numpy(), full[k].cpu().numpy()) < 2e-4, k
    assert float(acc[k].norm()) == pytest.approx(float(full[k].double().norm()), rel=1e-4), k


CURVES = [
  # name, U, I, nnz, H, B, loss, steps
  ('c1', 10_000, 5_000, 50, 128, 256, 'mse', 200),
  ('c2', 30_000, 26_744, 144, 200, 500, 'mse', 60),
  ('nll', 20_000, 8_000, 60, 256, 512, 'logloss', 200),
]


@pytest.mark.parametrize('name,U,I,nnz,H,B,loss,steps', CURVES)
def test_loss_curve_matches_oracle_without_teacher_forcing(name, U, I, nnz, H, B, loss, steps):
  """Both sides start from the same parameters and consume the same users in the same order; nothing is copied across
  afterwards.  bf16 GEMM operands make the two trajectories drift apart slowly: the gate is 2e-3 on every step's loss
  and 1e-3 on the mean of the curve (the figures observed on a B200 are printed)."""
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=1234)
  params = O.init_ae_params(I, [H], seed=0)
  lr, wd = 1e-3, 0.0
  tr = O.OracleTrainer('ae', params, loss=loss, optimizer='adam', lr=lr, weight_decay=wd, activation='tanh')
  model = make_model('ae', I, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, loss, 0.0, 'adam', lr, wd, _native.GEMM_TCGEN05)
  ds = device_dataset(indptr, indices, data, I)
  order = np.random.default_rng(11).permutation(U)
  ocurve = []
  for s in range(steps):
    users = order[(s * B) % (U - B):][:B]
    ob = O.collate(indptr, indices, data, I, users, B, True)[0]
    ocurve.append(tr.step(ob)[0])
    pool = collate_pool(ds.device_csr(), users, True)
    eng.train_step(pool, 0, B)
  gcurve = eng.losses(steps).numpy()
  ocurve = np.asarray(ocurve)
  rel = np.abs(gcurve - ocurve) / np.abs(ocurve)
  print('loss curve %s: %d steps, loss %.4f -> %.4f, max rel err %.2e, mean rel err %.2e, last-20 mean rel err %.2e'
        % (name, steps, ocurve[0], ocurve[-1], rel.max(), rel.mean(), rel[-20:].mean()))
  assert ocurve[-1] < ocurve[0]
  assert rel.max() < 2e-3
  assert abs(gcurve.mean() - ocurve.mean()) / ocurve.mean() < 1e-3
  # the trained parameters themselves (Adam normalises every gradient to +-lr per step, so element-wise agreement is
  # loose by construction; the norm of the update is what is comparable)
  for k, v in tr.state().items():
    got = dict(model.named_parameters())[k].detach().cpu().numpy()
    du_o = np.linalg.norm(v - params[k].numpy())
    du_g = np.linalg.norm(got - params[k].numpy())
    assert du_g == pytest.approx(du_o, rel=5e-2), k


def test_model_wider_than_the_matrix():
  """`Recoder(num_items=N)` with N larger than the matrix width (reference model.py:241): table rows beyond the matrix
  never receive a gradient but still move under dense Adam with weight decay, exactly like in the reference."""
  U, I, extra, H, B = 600, 3000, 517, 64, 256
  T = I + extra
  indptr, indices, data = synthetic_csr(U, I, 40, seed=5)
  params = O.init_ae_params(T, [H], seed=2)
  lr, wd = 1e-3, 1e-2
  tr = O.OracleTrainer('ae', params, loss='mse', optimizer='adam', lr=lr, weight_decay=wd, activation='tanh')
  model = make_model('ae', T, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, 'mse', 0.0, 'adam', lr, wd, _native.GEMM_TCGEN05)
  ds = device_dataset(indptr, indices, data, I)
  users = np.arange(B)
  with pytest.raises(ValueError, match='table_rows'):
    eng.train_step(collate_pool(ds.device_csr(), users, True), 0, B)       # collated for I items only
  for s in range(3):
    users = np.arange(s * 100, s * 100 + B)
    pool = collate_pool(ds.device_csr(), users, True, table_rows=T)
    assert pool.pos.numel() == T and bool((pool.pos[I:] == -1).all())
    ob = O.collate(indptr, indices, data, I, users, B, True)[0]
    oloss, _ = tr.step(ob)
    eng.train_step(pool, 0, B)
    assert float(eng.losses(1)[0]) == pytest.approx(oloss, rel=TOL)
  state = {n: p.detach().cpu().numpy() for n, p in model.named_parameters()}
  ost = tr.state()
  for k in (O.AE_EN_W, O.AE_DE_W):
    tail_o, tail_g = ost[k][I:], state[k][I:]
    assert np.abs(tail_o - params[k].numpy()[I:]).max() > 0        # weight decay moved the untouched rows
    np.testing.assert_allclose(tail_g, tail_o, rtol=1e-5, atol=1e-7)   # no gradient there: pure fp32 Adam arithmetic
    assert rel_err(state[k], ost[k]) < 5e-2
  # without negative sampling the columns are raw item ids and the decoder spans all T rows
  pool = collate_pool(ds.device_csr(), np.arange(B), False, table_rows=T)
  assert pool.pos.numel() == T


def test_matrix_wider_than_the_model_is_an_error():
  U, I, H, B = 300, 2000, 32, 128
  indptr, indices, data = synthetic_csr(U, I, 30, seed=6)
  params = O.init_ae_params(I - 100, [H], seed=2)
  model = make_model('ae', I - 100, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, 'mse', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
  ds = device_dataset(indptr, indices, data, I)
  with pytest.raises(ValueError, match='columns'):
    collate_pool(ds.device_csr(), np.arange(B), True, table_rows=I - 100)
  pool = collate_pool(ds.device_csr(), np.arange(B), True)
  with pytest.raises(ValueError, match='columns'):
    eng.train_step(pool, 0, B)
  # MF: a user id beyond the user table
  mparams = O.init_mf_params(I, 100, H, seed=1)
  mf = make_model('mf', I, 100, H, 'none', {k: v.numpy() for k, v in mparams.items()})
  meng = make_engine(mf, 'mse', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
  with pytest.raises(ValueError, match='user id'):
    meng.train_step(collate_pool(ds.device_csr(), np.arange(50, 50 + B), True), 0, B)

"""Evaluation path: the mask + top-k kernels against torch, `Recoder.recommend` against the
recommendations / scores / metrics the unmodified reference produced (tests/golden/eval/eval_golden.npz), and the
reference's own end-to-end golden-metric test (tests/test_model.py:14-84: Recall@20 0.40, Recall@50 0.43, NDCG@100
0.45, atol 0.01, then the same after a checkpoint round trip) on its own data (tests/golden/eval/ml_fixture.npz)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from recoder_b200 import _native
from recoder_b200._native import call, ptr
from recoder_b200.data import RecommendationDataset
from recoder_b200.metrics import NDCG, AveragePrecision, Recall
from recoder_b200.model import Recoder
from recoder_b200.nn import DynamicAutoencoder, MatrixFactorization

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'eval')


@pytest.mark.parametrize('rows,n,k', [(7, 100, 10), (64, 11466, 100), (33, 200000, 1024), (5, 50, 50), (3, 1000, 1)])
def test_topk_rows_matches_torch(rows, n, k):
  g = torch.Generator(device='cuda').manual_seed(rows * 31 + k)
  ld = (n + 7) // 8 * 8
  x = torch.randn(rows, ld, device='cuda', generator=g)
  x[:, :n // 7] = -float('inf')            # masked (seen) items
  vals = torch.empty(rows, k, device='cuda')
  idx = torch.empty(rows, k, dtype=torch.int64, device='cuda')
  call('rcd_topk_rows', ptr(x), ld, rows, n, k, ptr(vals), ptr(idx))
  tv, ti = torch.topk(x[:, :n], k, dim=1, sorted=True)
  assert torch.equal(vals, tv)
  finite = torch.isfinite(tv)
  assert torch.equal(idx[finite], ti[finite])          # random floats: no ties among the finite values
  assert bool(((idx >= 0) & (idx < n)).all())
  assert all(len(set(r)) == k for r in idx.tolist())   # no item twice


def test_topk_ties_prefer_lower_index():
  x = torch.zeros(2, 64, device='cuda')
  x[0, 10] = 1.0
  x[1, :] = 5.0
  vals = torch.empty(2, 4, device='cuda')
  idx = torch.empty(2, 4, dtype=torch.int64, device='cuda')
  call('rcd_topk_rows', ptr(x), 64, 2, 64, 4, ptr(vals), ptr(idx))
  assert idx.tolist() == [[10, 0, 1, 2], [0, 1, 2, 3]]
  assert vals.tolist() == [[1.0, 0.0, 0.0, 0.0], [5.0] * 4]


def test_mask_seen():
  rows, n = 5, 40
  row_ptr = torch.tensor([0, 2, 2, 5, 6, 9], dtype=torch.int32, device='cuda')
  items = torch.tensor([3, 7, 0, 1, 39, 5, 8, 9, 10], dtype=torch.int32, device='cuda')
  x = torch.zeros(rows, n, device='cuda')
  call('rcd_mask_seen', ptr(row_ptr), ptr(items), 0, rows, ptr(x), n)
  want = torch.zeros(rows, n)
  for r in range(rows):
    for p in range(int(row_ptr[r]), int(row_ptr[r + 1])):
      want[r, int(items[p])] = -float('inf')
  assert torch.equal(x.cpu(), want)


def _load_model(z, kind, U, I, H):
  if kind == 'ae':
    model = DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
  elif kind == 'ae2':
    model = DynamicAutoencoder(hidden_layers=[H, 16], activation_type='sigmoid')
  else:
    model = MatrixFactorization(embedding_size=H, activation_type='tanh')
  model.init_model(num_items=I, num_users=U)
  model = model.to('cuda')
  named = dict(model.named_parameters())
  with torch.no_grad():
    for n in z[kind + '/param_names']:
      named[str(n)].copy_(torch.from_numpy(z[kind + '/param/' + str(n)]).cuda())
  return model


@pytest.mark.parametrize('kind', ['ae', 'ae2', 'mf'])
def test_recommend_matches_reference(kind):
  z = np.load(os.path.join(GOLDEN_DIR, 'eval_golden.npz'))
  U, I, H, K = (int(v) for v in z['shape'])
  inp = sp.csr_matrix((z['in_data'], z['in_indices'], z['in_indptr']), shape=(U, I))
  tgt = sp.csr_matrix((z['tg_data'], z['tg_indices'], z['tg_indptr']), shape=(U, I))
  model = _load_model(z, kind, U, I, H)
  trainer = Recoder(model=model, use_cuda=True, optimizer_type='adam', loss='mse')
  ds = RecommendationDataset(inp, tgt)
  ui, _ = ds[np.arange(U)]
  recs = np.array(trainer.recommend(ui, K))
  ref = z[kind + '/recs']
  assert recs.shape == ref.shape
  # scores of the reference's recommended items through this implementation's forward (bf16 decoder operands)
  from recoder_b200.data import pool_of
  pool, _ = pool_of(ui, False)
  logits = model.forward_pool(pool).cpu().numpy()
  got_scores = np.take_along_axis(logits, ref, axis=1)
  want_scores = z[kind + '/rec_scores']
  scale = np.abs(want_scores).mean()
  assert np.abs(got_scores - want_scores).max() < 3e-2 * scale + 1e-3
  # seen items are never recommended
  for u in range(U):
    assert not set(recs[u]).intersection(inp.indices[inp.indptr[u]:inp.indptr[u + 1]])
  overlap = np.mean([len(set(recs[u]).intersection(ref[u])) / K for u in range(U)])
  top1 = np.mean(recs[:, 0] == ref[:, 0])
  assert overlap > 0.97 and top1 > 0.9, (overlap, top1)
  for m in (Recall(k=20, normalize=True), Recall(k=50, normalize=False), NDCG(k=50), AveragePrecision(k=10)):
    want = z[kind + '/metric/' + str(m)]
    got = np.array([m.evaluate(recs[u], tgt.indices[tgt.indptr[u]:tgt.indptr[u + 1]]) for u in range(U)])
    assert abs(got.mean() - want.mean()) < 5e-3, str(m)
  # the evaluator drives the same path
  res = trainer.evaluate(ds, num_recommendations=K, metrics=[NDCG(k=50)], batch_size=64)
  assert abs(np.mean(list(res.values())[0]) - z[kind + '/metric/NDCG@50'].mean()) < 5e-3


def _ml_matrices():
  z = np.load(os.path.join(GOLDEN_DIR, 'ml_fixture.npz'))
  out = []
  for name in ('train', 'val'):
    idx = z[name + '_indices'].astype(np.int32)
    shape = tuple(int(v) for v in z[name + '_shape'])
    out.append(sp.csr_matrix((np.ones(len(idx), dtype=np.float32), idx, z[name + '_indptr'].astype(np.int64)),
                             shape=shape))
  return out


@pytest.mark.parametrize('sparse,exp_recall_20,exp_recall_50,exp_ndcg_100', [
  (False, 0.40, 0.43, 0.45),
  (True, 0.40, 0.43, 0.45),
])
def test_model_golden_metrics(sparse, exp_recall_20, exp_recall_50, exp_ndcg_100, tmp_path):
  """The reference's tests/test_model.py, line for line, on this implementation."""
  train_matrix, val_matrix = _ml_matrices()
  train_dataset = RecommendationDataset(train_matrix)
  val_dataset = RecommendationDataset(val_matrix, train_matrix)
  torch.manual_seed(0)
  model = DynamicAutoencoder(hidden_layers=[200], activation_type='tanh', noise_prob=0.5, sparse=sparse)
  trainer = Recoder(model=model, use_cuda=True, optimizer_type='adam', loss='logloss')
  trainer.train(train_dataset=train_dataset, val_dataset=val_dataset, batch_size=500, lr=1e-3, weight_decay=2e-5,
                num_epochs=30, negative_sampling=True)
  recall_20, recall_50, ndcg_100 = Recall(k=20, normalize=True), Recall(k=50, normalize=True), NDCG(k=100)

  def check(tr):
    results = tr._evaluate(eval_dataset=val_dataset, num_recommendations=100, metrics=[recall_20, recall_50, ndcg_100],
                           batch_size=500)
    results = {m: np.mean(v) for m, v in results.items()}
    print({str(m): round(float(v), 4) for m, v in results.items()})
    assert np.isclose(results[recall_20], exp_recall_20, atol=0.01, rtol=0)
    assert np.isclose(results[recall_50], exp_recall_50, atol=0.01, rtol=0)
    assert np.isclose(results[ndcg_100], exp_ndcg_100, atol=0.01, rtol=0)

  check(trainer)
  state_file = trainer.save_state(str(tmp_path / 'test_model.model'))
  model = DynamicAutoencoder(sparse=sparse)
  trainer = Recoder(model=model, use_cuda=True, optimizer_type='adam', loss='logloss')
  trainer.init_from_model_file(state_file)
  check(trainer)


def test_resume_from_reference_checkpoint_matches_reference():
  """`init_from_model_file` on a checkpoint written by the unmodified reference, optimizer state hand-over included:
  the next training step must give the loss, gradients, parameters and Adam state the reference itself computed after
  resuming from the same file (tests/golden/make_checkpoint_fixture.py)."""
  from recoder_b200.data import collate_pool
  z = np.load(os.path.join(GOLDEN_DIR, 'ref_checkpoint_resume.npz'))
  lr, wd, batch = float(z['hyper'][0]), float(z['hyper'][1]), int(z['hyper'][2])
  matrix = sp.csr_matrix((z['csr_data'], z['csr_indices'], z['csr_indptr']))
  ds = RecommendationDataset(matrix)
  trainer = Recoder(model=DynamicAutoencoder(), use_cuda=True, optimizer_type='sgd', loss='mse')   # overwritten by the file
  trainer.init_from_model_file(os.path.join(GOLDEN_DIR, 'ref_checkpoint_epoch_2.model'))
  assert trainer.current_epoch == int(z['current_epoch']) and trainer.optimizer_type == 'adam' and trainer.loss == 'logloss'
  trainer._Recoder__init_training(train_dataset=ds, lr=lr, weight_decay=wd)
  pool = collate_pool(ds.device_csr(), z['users'], True)
  assert np.array_equal(pool.items.cpu().numpy(), z['items'])
  trainer.engine.train_step(pool, 0, batch)
  assert float(trainer.engine.losses(1)[0]) == pytest.approx(float(z['loss']), rel=1e-3)
  items = z['items']
  names = [str(n) for n in z['param_names']]
  grads = {'en_embedding_layer.weight': trainer.engine.last['dWe'], 'de_embedding_layer.weight': trainer.engine.last['dWd']}
  for n, g in grads.items():
    want = z['grad/' + n][items]
    got = g.detach().cpu().numpy()
    assert np.linalg.norm(got) == pytest.approx(np.linalg.norm(want), rel=3e-3), n
  params = dict(trainer.model.named_parameters())
  for n in names:
    got = params[n].detach().cpu().numpy()
    want = z['param/' + n]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-2, n
    st = trainer.optimizer.states[n]
    assert st.step == int(z['step/' + n])                       # the step counter continued from the checkpoint
    m_want, v_want = z['exp_avg/' + n], z['exp_avg_sq/' + n]
    assert np.linalg.norm(st.m.cpu().numpy() - m_want) / np.linalg.norm(m_want) < 2e-2, n
    assert np.linalg.norm(st.v.cpu().numpy() - v_want) / np.linalg.norm(v_want) < 2e-2, n

"""Pins the CPU oracle (oracle/recoder_oracle.py) against golden vectors produced by the unmodified reference,
against the reference's own collate property test (tests/test_data.py:129-165 there), against an independent
closed-form derivation, and — when /root/reference is mounted — against the live reference."""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from oracle import ref_shims
from tests.golden_util import Golden, case_names


def _trainer_from(g):
  m = g.meta
  params = {k: torch.from_numpy(v.copy()) for k, v in g.init_params().items()}
  return O.OracleTrainer(model=m['model'], params=params, loss=m['loss'],
                         confidence=m['loss_params'].get('confidence', 0.0), optimizer=m['opt'], lr=m['lr'],
                         weight_decay=m['wd'], activation=m['act'], sparse=m['sparse'],
                         is_constrained=bool(m.get('constrained', False)))


@pytest.mark.parametrize('name', case_names())
def test_oracle_matches_reference_golden(name):
  g = Golden(name)
  m = g.meta
  tr = _trainer_from(g)
  for users, steps in g.pools():
    batches = O.collate(g.indptr, g.indices, g.data, m['num_items'], users, m['batch'], m['neg'])
    assert len(batches) == len(steps)
    for b, s in zip(batches, steps):
      ref = g.step(s)
      # integer bookkeeping: bit-exact
      assert np.array_equal(b.users, ref['users'])
      if ref['items'] is None:
        assert b.items is None
      else:
        assert b.items.dtype == np.int64 and np.array_equal(b.items, ref['items'])
      assert np.array_equal(b.indices, ref['indices'])
      assert np.array_equal(b.values, ref['values'])
      assert tuple(b.size) == ref['size']
      loss, grads = tr.step(b, noise_keep=ref['noise_keep'], noise_prob=float(m.get('noise', 0.0)),
                            dropout_keep=ref['dropout_keep'], dropout_prob=float(m.get('dropout', 0.0)))
      assert loss == pytest.approx(ref['loss'], rel=1e-6, abs=1e-7)
      for n in g.param_names:
        np.testing.assert_allclose(grads[n].numpy(), ref['grads'][n], rtol=1e-5, atol=1e-7, err_msg=n)
      st = tr.state()
      for n in g.param_names:
        np.testing.assert_allclose(st[n], ref['params'][n], rtol=1e-5, atol=1e-7, err_msg=n)


@pytest.mark.parametrize('name', ['ae_mse_adam', 'ae_mse_conf_ratings', 'ae_nll_adam', 'ae_bce_adam', 'ae_nll_pool',
                                  'mf_mse_adam', 'mf_nll_sgd'])
def test_closed_form_matches_autograd(name):
  g = Golden(name)
  m = g.meta
  tr = _trainer_from(g)
  users, _ = next(iter(g.pools()))
  b = O.collate(g.indptr, g.indices, g.data, m['num_items'], users, m['batch'], m['neg'])[0]
  p0 = {k: v.copy() for k, v in g.init_params().items()}
  loss, grads = tr.step(b)
  conf = m['loss_params'].get('confidence', 0.0)
  if m['model'] == 'ae':
    L, cg = O.ae_closed_form(p0, b, m['loss'], conf, m['act'])
    np.testing.assert_allclose(grads[O.AE_EN_W].numpy()[b.items], cg['dWe'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(grads[O.AE_DE_W].numpy()[b.items], cg['dWd'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(grads[O.AE_EN_B].numpy(), cg['dbe'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(grads[O.AE_DE_B].numpy()[b.items], cg['dbd'], rtol=2e-4, atol=1e-6)
    # rows outside the batch get exactly zero gradient
    mask = np.ones(m['num_items'], dtype=bool)
    mask[b.items] = False
    assert not grads[O.AE_EN_W].numpy()[mask].any() and not grads[O.AE_DE_W].numpy()[mask].any()
  else:
    L, cg = O.mf_closed_form(p0, b, m['loss'], conf, m['act'])
    np.testing.assert_allclose(grads[O.MF_USER_W].numpy()[b.users], cg['dU'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(grads[O.MF_ITEM_W].numpy()[b.items], cg['dV'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(grads[O.MF_BIAS].numpy()[b.items], cg['dbias'], rtol=2e-4, atol=1e-6)
  assert loss == pytest.approx(L, rel=1e-5)


@pytest.mark.parametrize('batch_size', [1, 2, 5, 10, 13])
def test_collate_property(batch_size):
  """Restates the reference's tests/test_data.py:129-165 against the oracle collate."""
  rng = np.random.default_rng(batch_size)
  U, I = 100, 200
  rows = rng.integers(0, U, 1000)
  cols = rng.integers(0, I, 1000)
  pairs = np.unique(np.stack([rows, cols], 1), axis=0)
  import scipy.sparse as sp
  csr = sp.coo_matrix((np.ones(len(pairs), dtype=np.float32), (pairs[:, 0], pairs[:, 1])), shape=(U, I)).tocsr()
  users = np.arange(U)
  batches = O.collate(csr.indptr.astype(np.int64), csr.indices, csr.data, I, users, batch_size, True)
  assert len(batches) == int(np.ceil(U / batch_size))
  cur = 0
  for b in batches:
    dense = O.to_dense(b).numpy()
    sub = csr[cur:cur + batch_size]
    assert (dense > 0).sum(axis=1).tolist() == np.diff(sub.indptr).tolist()
    item_idx = {int(it): k for k, it in enumerate(b.items.tolist())}
    for r in range(sub.shape[0]):
      for c, v in zip(sub[r].indices, sub[r].data):
        assert dense[r, item_idx[int(c)]] == v
    cur += batch_size


@pytest.mark.skipif(not ref_shims.reference_available(), reason='reference tree not mounted')
def test_oracle_matches_live_reference():
  rdata, rnn, rlosses, rmodel = ref_shims.import_reference()
  import warnings
  warnings.simplefilter('ignore')
  from recoder_b200.synth import synthetic_csr, to_scipy
  U, I, H, B = 300, 257, 32, 64
  indptr, indices, data = synthetic_csr(U, I, 20, seed=5)
  csr = to_scipy(indptr, indices, data, I)
  ds = rdata.RecommendationDataset(csr)
  torch.manual_seed(0)
  model = rnn.DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
  trainer = rmodel.Recoder(model=model, use_cuda=False, optimizer_type='adam', loss='logloss')
  trainer._Recoder__init_training(train_dataset=ds, lr=1e-3, weight_decay=2e-5)
  params = {n: p.detach().clone() for n, p in model.named_parameters()}
  tr = O.OracleTrainer('ae', params, loss='logloss', optimizer='adam', lr=1e-3, weight_decay=2e-5)
  order = np.random.default_rng(0).permutation(U)
  for off in range(0, U, B):
    users = order[off:off + B]
    ui, _ = ds[users]
    rb = rdata.BatchCollator(B, True).collate(ui)[0]
    ob = O.collate(indptr, indices, data, I, users, B, True)[0]
    assert np.array_equal(rb.items.numpy(), ob.items) and np.array_equal(rb.indices.numpy(), ob.indices)
    trainer.optimizer.zero_grad()
    loss = trainer._Recoder__compute_loss(rb, None)
    loss.backward()
    trainer.optimizer.step()
    oloss, _ = tr.step(ob)
    assert oloss == pytest.approx(loss.item(), rel=1e-6)
  st = tr.state()
  for n, p in model.named_parameters():
    np.testing.assert_allclose(st[n], p.detach().numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.skipif(not ref_shims.reference_available(), reason='reference tree not mounted')
@pytest.mark.parametrize('hidden,tied,noise,dropout,loss,opt', [
  ([24, 12], False, 0.0, 0.0, 'mse', 'adagrad'),
  ([24, 12, 6], True, 0.0, 0.0, 'logloss', 'rmsprop'),
  ([32], False, 0.4, 0.3, 'logistic', 'adam'),
  ([24, 8], True, 0.25, 0.5, 'logloss', 'sgd'),
])
def test_oracle_matches_live_reference_general_models(hidden, tied, noise, dropout, loss, opt):
  """Multi-layer / tied autoencoders, input noise and bottleneck dropout, the remaining optimizers: the oracle against
  the live reference over several steps.  nn.Dropout's keep masks are obtained by replaying the draws the reference's
  forward is about to make on the global CPU generator (same shapes, same order) and rewinding it."""
  rdata, rnn, rlosses, rmodel = ref_shims.import_reference()
  import warnings
  warnings.simplefilter('ignore')
  from recoder_b200.synth import synthetic_csr, to_scipy
  U, I, B = 200, 181, 50
  indptr, indices, data = synthetic_csr(U, I, 15, seed=9)
  csr = to_scipy(indptr, indices, data, I)
  ds = rdata.RecommendationDataset(csr)
  torch.manual_seed(3)
  model = rnn.DynamicAutoencoder(hidden_layers=hidden, activation_type='tanh', is_constrained=tied, noise_prob=noise,
                                 dropout_prob=dropout)
  trainer = rmodel.Recoder(model=model, use_cuda=False, optimizer_type=opt, loss=loss)
  trainer._Recoder__init_training(train_dataset=ds, lr=1e-2, weight_decay=1e-4)
  model.train()
  params = {n: p.detach().clone() for n, p in model.named_parameters()}
  tr = O.OracleTrainer('ae', params, loss=loss, optimizer=opt, lr=1e-2, weight_decay=1e-4, is_constrained=tied)
  order = np.random.default_rng(1).permutation(U)
  for step, off in enumerate(range(0, U, B)):
    users = order[off:off + B]
    ui, _ = ds[users]
    rb = rdata.BatchCollator(B, True).collate(ui)[0]
    ob = O.collate(indptr, indices, data, I, users, B, True)[0]
    nk = dk = None
    torch.manual_seed(500 + step)
    if noise > 0:
      nk = (torch.nn.functional.dropout(torch.ones(tuple(rb.size)), noise, True) != 0).numpy().astype(np.float32)
    if dropout > 0:
      dk = (torch.nn.functional.dropout(torch.ones(rb.size[0], hidden[-1]), dropout, True) != 0).numpy().astype(np.float32)
    torch.manual_seed(500 + step)
    trainer.optimizer.zero_grad()
    ref_loss = trainer._Recoder__compute_loss(rb, None)
    ref_loss.backward()
    trainer.optimizer.step()
    oloss, _ = tr.step(ob, noise_keep=nk, noise_prob=noise, dropout_keep=dk, dropout_prob=dropout)
    assert oloss == pytest.approx(ref_loss.item(), rel=1e-5), step
  st = tr.state()
  for n, p in model.named_parameters():
    np.testing.assert_allclose(st[n], p.detach().numpy(), rtol=2e-4, atol=1e-6, err_msg=n)

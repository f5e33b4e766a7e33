"""Training on a dataset whose TARGET matrix differs from its input matrix
(`RecommendationDataset(interactions_matrix, target_interactions_matrix)`, reference recoder/data.py:41-45, 54-63;
`Recoder.__compute_loss` builds the dense target from the target batch, recoder/model.py:464-472): the encoder runs over
the input pool's items, the decoder / loss / decoder gradients over the target pool's items.  Both step paths (native
executor and Python launch sequence) against the CPU oracle, and against each other bit for bit."""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200.data import RecommendationDataset, collate_pool
from recoder_b200.synth import synthetic_csr, to_scipy
from tests.gpu_util import make_engine, make_model, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _matrices(U, I, nnz_in, nnz_t):
  a = synthetic_csr(U, I, nnz_in, seed=31)
  b = synthetic_csr(U, I, nnz_t, seed=32)
  return a, b


def _run(kind, loss, native, steps=3):
  U, I, H, B = 1500, 6000, 96, 256
  (ip, ii, idt), (tp, ti, tdt) = _matrices(U, I, 40, 25)
  params = O.init_ae_params(I, [H], seed=5) if kind == 'ae' else O.init_mf_params(I, U, H, seed=5)
  act = 'tanh' if kind == 'ae' else 'none'
  tr = O.OracleTrainer(kind, params, loss=loss, optimizer='adam', lr=1e-3, weight_decay=1e-4, activation=act)
  model = make_model(kind, I, U, [H] if kind == 'ae' else H, act, {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, loss, 0.0, 'adam', 1e-3, 1e-4, _native.GEMM_TCGEN05)
  eng.native_enabled = native
  ds = RecommendationDataset(to_scipy(ip, ii, idt, I), to_scipy(tp, ti, tdt, I))
  order = np.random.default_rng(2).permutation(U)
  out = []
  for s in range(steps):
    users = order[s * B:(s + 1) * B]
    pool = collate_pool(ds.device_csr(), users, True)
    tpool = collate_pool(ds.device_target_csr(), users, True)
    ob = O.collate(ip, ii, idt, I, users, B, True)[0]
    ot = O.collate(tp, ti, tdt, I, users, B, True)[0]
    assert np.array_equal(pool.items.cpu().numpy(), ob.items) and np.array_equal(tpool.items.cpu().numpy(), ot.items)
    assert pool.n != tpool.n
    with torch.no_grad():      # teacher forcing: the oracle steps from the GPU's current parameters
      for n2, p2 in model.named_parameters():
        tr.params[n2].copy_(p2.detach().cpu())
    oloss, ograds = tr.step(ob, ot)
    eng.train_step(pool, 0, B, target_pool=tpool)
    gloss = float(eng.losses(1)[0])
    assert gloss == pytest.approx(oloss, rel=TOL), 'loss step %d' % s
    if kind == 'ae':
      want = {'dWe': ograds[O.AE_EN_W].numpy()[ob.items], 'dWd': ograds[O.AE_DE_W].numpy()[ot.items],
              'dbd': ograds[O.AE_DE_B].numpy()[ot.items], 'dbe': ograds[O.AE_EN_B].numpy()}
    else:
      want = {'dV': ograds[O.MF_ITEM_W].numpy()[ot.items], 'dbias': ograds[O.MF_BIAS].numpy()[ot.items],
              'dU': ograds[O.MF_USER_W].numpy()[ob.users]}
    for key, w in want.items():
      got = eng.last[key].detach().cpu().numpy()
      assert got.shape == w.shape, key
      assert np.linalg.norm(got) == pytest.approx(np.linalg.norm(w), rel=TOL), '%s norm step %d' % (key, s)
      assert rel_err(got, w) < 2e-2, '%s step %d' % (key, s)
    out.append(gloss)
  val = eng.eval_loss(collate_pool(ds.device_csr(), order[:B], True), 0, B,
                      target_pool=collate_pool(ds.device_target_csr(), order[:B], True))
  state = {n: p.detach().clone() for n, p in model.named_parameters()}
  ost = tr.state()
  for n in ost:
    assert rel_err(state[n].cpu().numpy(), ost[n]) < 5e-2, n
  assert (eng._native is not None) == native
  return out, val, state


@pytest.mark.parametrize('kind,loss', [('ae', 'mse'), ('ae', 'logloss'), ('mf', 'mse')])
def test_separate_target_matrix_matches_oracle_on_both_step_paths(kind, loss):
  a = _run(kind, loss, native=True)
  b = _run(kind, loss, native=False)
  assert a[0] == b[0] and a[1] == b[1]
  for n in a[2]:
    assert torch.equal(a[2][n], b[2][n]), n

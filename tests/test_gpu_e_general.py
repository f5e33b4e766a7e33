"""Kernels of the generalised model against PyTorch fp32: the inner-layer SGEMM in all four
layouts, the Philox dropout (rate, determinism, explicit masks, backward = same mask), Adagrad / RMSprop steps against
torch.optim, and whole steps of multi-layer / tied / noisy autoencoders against the CPU oracle at realistic sizes."""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200._native import call, ptr
from recoder_b200.data import collate_pool
from recoder_b200.synth import synthetic_csr
from tests.gpu_util import device_dataset, inner_grads, make_engine, make_model, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('ta,tb', [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(64, 64, 16), (333, 200, 77), (2048, 512, 256), (5, 3, 1000)])
def test_sgemm_matches_torch(ta, tb, M, N, K):
  g = torch.Generator(device='cuda').manual_seed(M * 7 + N)
  A = torch.randn((K, M) if ta else (M, K), device='cuda', generator=g)
  B = torch.randn((N, K) if tb else (K, N), device='cuda', generator=g)
  bias = torch.randn(N, device='cuda', generator=g)
  C = torch.full((M, N), 3.0, device='cuda')
  want = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
  call('rcd_sgemm', ta, tb, M, N, K, ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(C), N, ptr(bias),
       _native.ACT_IDS['tanh'], 0)
  assert rel_err(C.cpu().numpy(), torch.tanh(want + bias.double()).cpu().numpy()) < 1e-5
  C2 = torch.full((M, N), 3.0, device='cuda')
  call('rcd_sgemm', ta, tb, M, N, K, ptr(A), A.shape[1], ptr(B), B.shape[1], ptr(C2), N, None,
       _native.ACT_IDS['none'], 1)
  assert rel_err(C2.cpu().numpy(), (want + 3.0).cpu().numpy()) < 1e-5


def test_dropout_kernel():
  n, p = 1 << 20, 0.3
  x = torch.rand(n, device='cuda') + 0.5
  y1, y2, y3 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
  call('rcd_dropout', ptr(x), n, p, 1234, 1, 0, None, ptr(y1))
  call('rcd_dropout', ptr(x), n, p, 1234, 1, 0, None, ptr(y2))
  call('rcd_dropout', ptr(x), n, p, 1235, 1, 0, None, ptr(y3))
  assert torch.equal(y1, y2), 'same key, same counter -> same mask'
  keep = (y1 != 0)
  assert abs(float(keep.float().mean()) - (1 - p)) < 3e-3
  assert torch.allclose(y1[keep], x[keep] / (1 - p), rtol=1e-6)
  assert float((keep != (y3 != 0)).float().mean()) > 0.3, 'another seed gives another mask'
  # a shifted window of the same stream reproduces the same decisions (masks do not depend on how rows are sharded)
  y4 = torch.empty(n - 1000, device='cuda')
  call('rcd_dropout', ptr(x[1000:]), n - 1000, p, 1234, 1, 1000, None, ptr(y4))
  assert torch.equal(y4 != 0, keep[1000:])
  mask = (torch.rand(n, device='cuda') > 0.5).to(torch.uint8)
  call('rcd_dropout', ptr(x), n, 0.5, 0, 0, 0, ptr(mask), ptr(y1))
  assert torch.allclose(y1, x * mask.float() * 2.0)


@pytest.mark.parametrize('opt', ['adagrad', 'rmsprop'])
def test_adagrad_rmsprop_kernels_match_torch(opt):
  I, H, n = 3000, 72, 700
  g = torch.Generator().manual_seed(5)
  p0 = torch.randn(I, H, generator=g) * 0.1
  ids = torch.randperm(I, generator=g)[:n].sort().values
  ref = torch.nn.Parameter(p0.clone())
  if opt == 'adagrad':
    o = torch.optim.Adagrad([{'params': ref, 'weight_decay': 1e-3}], lr=1e-2)
  else:
    o = torch.optim.RMSprop([{'params': ref, 'weight_decay': 1e-3}], lr=1e-2, momentum=0.9)
  p = p0.clone().cuda()
  s1, s2 = torch.zeros_like(p), torch.zeros_like(p)
  pos = torch.full((I,), -1, dtype=torch.int32)
  pos[ids] = torch.arange(n, dtype=torch.int32)
  pos = pos.cuda()
  for t in range(3):
    gr = torch.randn(n, H, generator=g)
    dense = torch.zeros(I, H)
    dense[ids] = gr
    ref.grad = dense
    o.step()
    grc = gr.cuda()
    if opt == 'adagrad':
      call('rcd_adagrad_step', ptr(p), ptr(s1), I, H, ptr(grc), H, ptr(pos), 1e-2, 1e-10, 1e-3)
    else:
      call('rcd_rmsprop_step', ptr(p), ptr(s1), ptr(s2), I, H, ptr(grc), H, ptr(pos), 1e-2, 0.99, 1e-8, 0.9, 1e-3)
    assert rel_err(p.cpu().numpy(), ref.detach().numpy()) < 2e-6, 'step %d' % t


CASES = [
  # hidden, tied, noise, dropout, loss, opt
  ([256, 64], False, 0.0, 0.0, 'mse', 'adam'),
  ([128, 64, 32], True, 0.0, 0.0, 'logloss', 'adam'),
  ([200], False, 0.5, 0.2, 'logloss', 'adam'),
  ([128, 48], False, 0.3, 0.5, 'logistic', 'rmsprop'),
  ([96], False, 0.0, 0.0, 'mse', 'adagrad'),
]


@pytest.mark.parametrize('hidden,tied,noise,dropout,loss,opt', CASES)
def test_general_autoencoder_step_matches_oracle(hidden, tied, noise, dropout, loss, opt):
  U, I, nnz, B = 3000, 12000, 80, 512
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=21)
  params = O.init_ae_params(I, hidden, seed=4, is_constrained=tied)
  for k in params:
    if 'bias' in k:
      params[k] = torch.randn(params[k].shape) * 0.05
  lr, wd = 1e-3, 1e-5
  tr = O.OracleTrainer('ae', params, loss=loss, optimizer=opt, lr=lr, weight_decay=wd, activation='tanh',
                       is_constrained=tied)
  model = make_model('ae', I, U, hidden, 'tanh', {k: v.numpy() for k, v in params.items()}, constrained=tied,
                     noise=noise, dropout=dropout)
  eng = make_engine(model, loss, 0.0, opt, lr, wd, _native.GEMM_TCGEN05)
  ds = device_dataset(indptr, indices, data, I)
  order = np.random.default_rng(9).permutation(U)
  rng = np.random.default_rng(77)
  for s in range(2):
    users = order[s * B:(s + 1) * B]
    pool = collate_pool(ds.device_csr(), users, True)
    ob = O.collate(indptr, indices, data, I, users, B, True)[0]
    with torch.no_grad():
      for n2, p2 in model.named_parameters():
        tr.params[n2].copy_(p2.detach().cpu())
    nk = (rng.random(ob.values.shape[0]) >= noise).astype(np.uint8) if noise > 0 else None
    dk = (rng.random((len(users), hidden[-1])) >= dropout).astype(np.uint8) if dropout > 0 else None
    oloss, ograds = tr.step(ob, noise_keep=nk, noise_prob=noise, dropout_keep=dk, dropout_prob=dropout)
    eng.debug_noise_keep = None if nk is None else torch.from_numpy(nk).cuda()
    eng.debug_dropout_keep = None if dk is None else torch.from_numpy(dk).cuda()
    eng.train_step(pool, 0, len(users))
    assert float(eng.losses(1)[0]) == pytest.approx(oloss, rel=1e-3)
    got = {'dWe': eng.last['dWe'], 'dbe': eng.last['dbe'], 'dbd': eng.last['dbd']}
    want = {'dWe': ograds[O.AE_EN_W].numpy()[ob.items], 'dbe': ograds[O.AE_EN_B].numpy(),
            'dbd': ograds[O.AE_DE_B].numpy()[ob.items]}
    if not tied:
      got['dWd'] = eng.last['dWd']
      want['dWd'] = ograds[O.AE_DE_W].numpy()[ob.items]
    for n2, gt in inner_grads(eng, model).items():
      got[n2] = gt
      want[n2] = ograds[n2].numpy()
    for key, w in want.items():
      gnp = got[key].detach().cpu().numpy()
      assert np.linalg.norm(gnp) == pytest.approx(np.linalg.norm(w), rel=2e-3), '%s norm step %d' % (key, s)
      assert rel_err(gnp, w) < 2e-2, '%s step %d' % (key, s)
  state = {n: p.detach().cpu().numpy() for n, p in model.named_parameters()}
  for n, w in tr.state().items():
    assert rel_err(state[n], w) < 5e-2, n


def test_philox_dropout_trains_and_differs_between_steps():
  """Without injected masks the engine draws its own Philox masks: finite loss, and the noise changes every step."""
  U, I, nnz, B, H = 2000, 5000, 50, 256, 64
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=2)
  params = O.init_ae_params(I, [H], seed=1)
  model = make_model('ae', I, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()}, noise=0.5, dropout=0.3)
  eng = make_engine(model, 'logloss', 0.0, 'adam', 0.0, 0.0, _native.GEMM_TCGEN05)   # lr 0: parameters frozen
  ds = device_dataset(indptr, indices, data, I)
  pool = collate_pool(ds.device_csr(), np.arange(B), True)
  losses = []
  for _ in range(3):
    eng.train_step(pool, 0, B)
    losses.append(float(eng.losses(1)[0]))
  assert all(np.isfinite(losses))
  assert len({round(x, 6) for x in losses}) == 3, losses   # same batch, same weights, different masks
  clean = eng.eval_loss(pool, 0, B)                         # eval mode: no noise, no dropout
  assert abs(clean - eng.eval_loss(pool, 0, B)) < 1e-9


def test_chunked_heavy_columns_match_single_pass():
  """Tall slices cut columns with more than 128 entries into chunks (k_heavy_setup / k_heavy_reduce); the result must
  equal the single-pass kernel (same fp32 terms, different association only inside the heavy columns)."""
  U, I, nnz, B, H = 3000, 4000, 80, 2048, 96       # item 0 sits in ~70 % of the rows: ~1400 entries in its column
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=8)
  params = O.init_ae_params(I, [H], seed=2)
  ds = device_dataset(indptr, indices, data, I)
  pool = collate_pool(ds.device_csr(), np.arange(B), True)
  got = []
  for threshold in (1 << 30, 0):
    model = make_model('ae', I, U, [H], 'tanh', {k: v.numpy() for k, v in params.items()})
    eng = make_engine(model, 'logloss', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
    eng.HEAVY_COLUMN_ROWS = threshold
    eng.train_step(pool, 0, B)
    loss = float(eng.losses(1)[0])
    got.append((loss, {k: eng.last[k].detach().cpu().numpy().copy() for k in ('dWe', 'dWd', 'dbd', 'dbe')}))
  assert got[0][0] == got[1][0]
  for k in got[0][1]:
    assert rel_err(got[1][1][k], got[0][1][k]) < 1e-6, k


class _SumHuber(torch.nn.Module):
  """A loss the library does not know: Huber with sum reduction (any nn.Module is accepted, recoder/model.py:88-89)."""

  def forward(self, input, target):
    return torch.nn.functional.smooth_l1_loss(input, target, reduction='sum', beta=0.5)


@pytest.mark.parametrize('kind,loss_module', [('ae', _SumHuber()), ('mf', _SumHuber()),
                                              ('ae', torch.nn.BCEWithLogitsLoss(reduction='sum',
                                                                               pos_weight=torch.tensor(3.0)))])
def test_custom_loss_module_step_matches_oracle(kind, loss_module):
  U, I, nnz, B, H = 2000, 6000, 60, 384, 64
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=13)
  params = O.init_ae_params(I, [H], seed=6) if kind == 'ae' else O.init_mf_params(I, U, H, seed=6)
  act = 'tanh' if kind == 'ae' else 'none'
  tr = O.OracleTrainer(kind, params, loss=loss_module, optimizer='adam', lr=1e-3, weight_decay=0.0, activation=act)
  model = make_model(kind, I, U, [H] if kind == 'ae' else H, act, {k: v.numpy() for k, v in params.items()})
  import copy
  eng = make_engine(model, copy.deepcopy(loss_module), 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)   # .to('cuda') moves a module in place
  ds = device_dataset(indptr, indices, data, I)
  users = np.random.default_rng(3).permutation(U)[:B]
  pool = collate_pool(ds.device_csr(), users, True)
  ob = O.collate(indptr, indices, data, I, users, B, True)[0]
  oloss, ograds = tr.step(ob)
  eng.train_step(pool, 0, B)
  assert float(eng.losses(1)[0]) == pytest.approx(oloss, rel=1e-3)
  if kind == 'ae':
    want = {'dWe': ograds[O.AE_EN_W].numpy()[ob.items], 'dWd': ograds[O.AE_DE_W].numpy()[ob.items],
            'dbd': ograds[O.AE_DE_B].numpy()[ob.items], 'dbe': ograds[O.AE_EN_B].numpy()}
  else:
    want = {'dV': ograds[O.MF_ITEM_W].numpy()[ob.items], 'dbias': ograds[O.MF_BIAS].numpy()[ob.items],
            'dU': ograds[O.MF_USER_W].numpy()[ob.users]}
  for key, w in want.items():
    got = eng.last[key].detach().cpu().numpy()
    # the whole dL/dlogits goes through bf16 here (no fp32 sparse part), hence the wider gate
    assert np.linalg.norm(got) == pytest.approx(np.linalg.norm(w), rel=5e-3), key
    assert rel_err(got, w) < 3e-2, key
  assert abs(eng.eval_loss(pool, 0, B) - float(tr.compute_loss(ob).item())) < 2e-3 * abs(oloss) + 1e-2


ACTS = ['tanh', 'sigmoid', 'relu', 'selu', 'celu', 'hardshrink', 'atan', 'sinh', 'asinh', 'expm1']


@pytest.mark.parametrize('act', ACTS)
def test_activation_forward_and_derivative_match_torch(act):
  """recoder/nn.py:6-9 applies `torch.<activation_type>`: the kernels' forward (rcd_bias_act) and their derivative
  expressed through the OUTPUT (rcd_act_grad) against torch and torch autograd."""
  rows, H = 257, 72
  g = torch.Generator(device='cuda').manual_seed(3)
  x = (torch.randn(rows, H, device='cuda', generator=g) * 1.5)
  bias = torch.randn(H, device='cuda', generator=g) * 0.1
  Z = torch.empty(rows, H, device='cuda')
  call('rcd_bias_act', ptr(x), ptr(bias), rows, H, _native.ACT_IDS[act], ptr(Z), None, H)
  pre = (x + bias).double().requires_grad_(True)
  want = getattr(torch, act)(pre)
  torch.testing.assert_close(Z.double(), want.detach(), rtol=2e-6, atol=2e-6)
  dy = torch.randn(rows, H, device='cuda', generator=g)
  want.backward(dy.double())
  dpre = torch.empty(rows, H, device='cuda')
  call('rcd_act_grad', ptr(dy), ptr(Z), rows * H, _native.ACT_IDS[act], ptr(dpre))
  torch.testing.assert_close(dpre.double(), pre.grad, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('act', ['selu', 'celu', 'atan'])
def test_step_with_other_torch_activations_matches_oracle(act):
  U, I, H, B = 2000, 6000, 96, 256
  indptr, indices, data = synthetic_csr(U, I, 40, seed=13)
  params = O.init_ae_params(I, [H], seed=3)
  tr = O.OracleTrainer('ae', params, loss='logloss', optimizer='adam', lr=1e-3, activation=act)
  model = make_model('ae', I, U, [H], act, {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, 'logloss', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
  ds = device_dataset(indptr, indices, data, I)
  users = np.arange(B)
  pool = collate_pool(ds.device_csr(), users, True)
  ob = O.collate(indptr, indices, data, I, users, B, True)[0]
  oloss, ograds = tr.step(ob)
  eng.train_step(pool, 0, B)
  assert float(eng.losses(1)[0]) == pytest.approx(oloss, rel=1e-3)
  for key, name in (('dWe', O.AE_EN_W), ('dWd', O.AE_DE_W)):
    want = ograds[name].numpy()[ob.items]
    got = eng.last[key].detach().cpu().numpy()
    assert np.linalg.norm(got) == pytest.approx(np.linalg.norm(want), rel=1e-3), key
    assert rel_err(got, want) < 2e-2, key


def test_unsupported_activation_is_refused_loudly():
  from recoder_b200.nn import DynamicAutoencoder
  with pytest.raises(NotImplementedError, match='pre-activation'):
    DynamicAutoencoder(hidden_layers=[8], activation_type='sin').init_model(num_items=10)

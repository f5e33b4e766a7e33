"""Deferred dense Adam (`rcd_adam_lazy_*`, Optimizer.enable_lazy): torch.optim.Adam's dense semantics — every row of an
embedding table moves on every step (recoder/model.py:118-135, 398-399) — with the updates of rows outside the batch
postponed until the row is next read.  The contract is BIT-identity with the dense kernel (`rcd_adam_step`), which is
itself held to torch.optim.Adam elsewhere (tests/test_gpu_b_kernels.py): parameters, exp_avg and exp_avg_sq after
>= 50 steps, with and without weight decay, across a learning-rate change, through both step paths and through the
public `Recoder` API."""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200._native import call, ptr
from recoder_b200.data import RecommendationDataset, collate_pool
from recoder_b200.engine import ADAM_BETAS, ADAM_EPS
from recoder_b200.model import Recoder
from recoder_b200.nn import DynamicAutoencoder, MatrixFactorization
from recoder_b200.synth import epoch_user_order, synthetic_csr, to_scipy
from tests.gpu_util import device_dataset, make_engine, make_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('wd', [0.0, 1e-2])
@pytest.mark.parametrize('H', [64, 50])
def test_lazy_kernels_replay_the_dense_kernel_bit_for_bit(wd, H):
  I, steps = 3000, 60
  lib = _native.load()
  g = torch.Generator(device='cuda').manual_seed(1)
  p0 = torch.randn(I, H, device='cuda', generator=g) * 0.1
  dense = [p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)]
  lazy = [p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)]
  ref = torch.nn.Parameter(p0.clone())
  topt = torch.optim.Adam([ref], lr=1e-3, weight_decay=wd)
  last = torch.zeros(I, dtype=torch.int32, device='cuda')
  cap = 256
  scal = torch.zeros(cap, 2, device='cuda')
  scal_host = torch.zeros(cap, 2)
  rng = np.random.default_rng(0)
  pos = torch.empty(I, dtype=torch.int32, device='cuda')
  lr = 1e-3
  for t in range(1, steps + 1):
    if t == 31:
      lr = 3e-4                     # MultiStepLR-style change: later scalars use the new rate, history keeps the old
      for grp in topt.param_groups:
        grp['lr'] = lr
    # heavy-tailed row choice: some rows every step, some once in a while, some never
    n = int(rng.integers(50, 400))
    ids_np = np.unique(np.minimum((I * rng.random(n) ** 3).astype(np.int64), I - 1))
    n = len(ids_np)
    ids = torch.from_numpy(ids_np).cuda()
    grad = torch.randn(n, H, device='cuda', generator=g) * 0.01
    # dense kernel
    pos.fill_(-1)
    pos[ids] = torch.arange(n, dtype=torch.int32, device='cuda')
    call('rcd_adam_step', ptr(dense[0]), ptr(dense[1]), ptr(dense[2]), I, H, ptr(grad), H, ptr(pos), lr,
         ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, wd, t)
    # torch reference
    full = torch.zeros(I, H, device='cuda')
    full[ids] = grad
    ref.grad = full
    topt.step()
    # lazy: publish the scalars of step t, catch the batch rows up to t-1, update them at t
    _native.check(lib.rcd_adam_scalars(lr, ADAM_BETAS[0], ADAM_BETAS[1], t, 1, scal_host[t:t + 1].data_ptr()), 'scalars')
    scal[t].copy_(scal_host[t])
    call('rcd_adam_lazy_catchup', ptr(lazy[0]), ptr(lazy[1]), ptr(lazy[2]), H, ptr(ids), n, ptr(last), t - 1, ptr(scal),
         0, cap, ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, wd, 1, None, None)
    assert bool((last[ids] == t - 1).all())
    call('rcd_adam_lazy_update', ptr(lazy[0]), ptr(lazy[1]), ptr(lazy[2]), H, ptr(ids), n, ptr(grad), H, ptr(last), lr,
         ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, wd, t)
    assert bool((last[ids] == t).all())
    if t % 20 == 0:
      # rows of the current batch are current and identical to the dense table even before any flush
      for a, b in zip(dense, lazy):
        assert torch.equal(a[ids], b[ids])
  stale = int((last < steps).sum())
  assert stale > I // 4, 'the test must leave a good share of rows deferred'
  call('rcd_adam_lazy_catchup', ptr(lazy[0]), ptr(lazy[1]), ptr(lazy[2]), H, None, I, ptr(last), steps, ptr(scal), 0, cap,
       ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, wd, 1, None, None)
  torch.cuda.synchronize()
  assert bool((last == steps).all())
  for name, a, b in zip(('p', 'exp_avg', 'exp_avg_sq'), dense, lazy):
    assert torch.equal(a, b), name
  torch.testing.assert_close(lazy[0], ref.data, rtol=2e-5, atol=1e-7)


CASES = [
  # kind, U, I, nnz, H, B, loss, negative sampling, native
  ('ae', 6000, 20000, 40, 64, 256, 'logloss', True, True),
  ('ae', 6000, 20000, 40, 64, 256, 'mse', True, False),
  ('mf', 6000, 8000, 40, 48, 256, 'mse', True, True),
  ('mf', 6000, 8000, 40, 48, 256, 'logloss', True, False),
  ('ae', 3000, 2000, 30, 32, 128, 'mse', False, True),       # no negative sampling: every row is in every batch
]


def _train(kind, U, I, nnz, H, B, loss, neg, native, lazy, steps=50):
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=8)
  params = O.init_ae_params(I, [H], seed=4) if kind == 'ae' else O.init_mf_params(I, U, H, seed=4)
  model = make_model(kind, I, U, [H] if kind == 'ae' else H, 'tanh' if kind == 'ae' else 'none',
                     {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, loss, 0.0, 'adam', 1e-3, 1e-3, _native.GEMM_TCGEN05)
  eng.native_enabled = native
  eng.lazy_adam = lazy
  ds = device_dataset(indptr, indices, data, I)
  order = np.random.default_rng(3).permutation(U)
  mid = None
  for s in range(steps):
    if s == steps // 2:
      eng.opt.lr = 5e-4
      mid = eng.eval_loss(collate_pool(ds.device_csr(), order[:B], neg), 0, B)   # reads rows in the middle of the run
    users = order[(s * B) % (U - B):][:B]
    eng.train_step(collate_pool(ds.device_csr(), users, neg), 0, B)
  deferred = {n: int((st.last < st.step).sum()) for n, st in eng.opt.states.items() if st.lazy}
  eng.opt.flush()
  torch.cuda.synchronize()
  state = {n: p.detach().clone() for n, p in model.named_parameters()}
  ostate = {n: (st.m.clone(), st.v.clone()) for n, st in eng.opt.states.items()}
  return eng.losses(steps).clone(), mid, state, ostate, deferred


@pytest.mark.parametrize('kind,U,I,nnz,H,B,loss,neg,native', CASES)
def test_engine_with_deferred_adam_is_bit_identical(kind, U, I, nnz, H, B, loss, neg, native):
  a = _train(kind, U, I, nnz, H, B, loss, neg, native, lazy=True)
  b = _train(kind, U, I, nnz, H, B, loss, neg, native, lazy=False)
  assert a[4] and not b[4], 'deferred mode was not active'
  if neg:
    assert max(a[4].values()) > 0, 'nothing was deferred'
  assert torch.equal(a[0], b[0])
  assert a[1] == b[1]
  for n in b[2]:
    assert torch.equal(a[2][n], b[2][n]), n
  for n in b[3]:
    assert torch.equal(a[3][n][0], b[3][n][0]) and torch.equal(a[3][n][1], b[3][n][1]), n


def test_auto_mode_defers_only_where_it_pays():
  indptr, indices, data = synthetic_csr(4000, 50000, 40, seed=8)
  params = O.init_ae_params(50000, [32], seed=4)
  model = make_model('ae', 50000, 4000, [32], 'tanh', {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, 'mse', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
  eng.lazy_adam = 'auto'
  ds = device_dataset(indptr, indices, data, 50000)
  eng.train_step(collate_pool(ds.device_csr(), np.arange(128), True), 0, 128)      # n << I
  assert len(eng.opt.lazy_names()) == 2
  indptr, indices, data = synthetic_csr(4000, 600, 40, seed=8)
  params = O.init_ae_params(600, [32], seed=4)
  model = make_model('ae', 600, 4000, [32], 'tanh', {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, 'mse', 0.0, 'adam', 1e-3, 0.0, _native.GEMM_TCGEN05)
  eng.lazy_adam = 'auto'
  ds = device_dataset(indptr, indices, data, 600)
  eng.train_step(collate_pool(ds.device_csr(), np.arange(512), True), 0, 512)      # n ~ I: the dense pass is cheaper
  assert eng.opt.lazy_names() == []


@pytest.mark.parametrize('kind', ['ae', 'mf'])
def test_recoder_api_sees_current_parameters(kind):
  """state_dict(), save_state and recommend() flush the deferred rows: a run with deferral on and one with it off give
  identical checkpoints and recommendations."""
  U, I, H, B = 5000, 12000, 32, 200
  indptr, indices, data = synthetic_csr(U, I, 30, seed=9)
  matrix = to_scipy(indptr, indices, data, I)
  out = []
  for lazy in (True, False):
    torch.manual_seed(0)
    model = DynamicAutoencoder(hidden_layers=[H]) if kind == 'ae' else MatrixFactorization(embedding_size=H)
    tr = Recoder(model=model, use_cuda=True, optimizer_type='adam', loss='logloss', lazy_adam=lazy)
    ds = RecommendationDataset(matrix)
    tr.train(ds, lr=1e-3, weight_decay=1e-4, num_epochs=2, iters_per_epoch=20, batch_size=B, negative_sampling=True,
             lr_milestones=[2], user_order=lambda e: epoch_user_order(U, e))
    assert bool(tr.optimizer.lazy_names()) == lazy
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    osd = tr.optimizer.state_dict(dense=True)
    rec = tr.recommend(ds[np.arange(64)][0], 10)
    out.append((sd, osd, rec, tr.last_epoch_losses.copy()))
  for k in out[0][0]:
    assert torch.equal(out[0][0][k], out[1][0][k]), k
  for i in out[1][1]['state']:
    for k, v in out[1][1]['state'][i].items():
      assert torch.equal(torch.as_tensor(v), torch.as_tensor(out[0][1]['state'][i][k])), (i, k)
  assert out[0][2] == out[1][2]
  assert np.array_equal(out[0][3], out[1][3])

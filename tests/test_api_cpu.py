"""Host-side behaviour of the reference's API surface that needs no GPU: argument validation and exceptions
(SURVEY.md §8b "Error conventions"), the learning-rate schedule against torch's MultiStepLR driven the way the reference
drives it (recoder/model.py:327-332, 364-366), optimizer construction rules (model.py:101-164), pool / slice bookkeeping
of the data loader (data.py:114-126, 166-167)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from recoder_b200.data import RecommendationDataLoader, RecommendationDataset
from recoder_b200.engine import Optimizer
from recoder_b200.losses import MSELoss, MultinomialNLLLoss
from recoder_b200.model import Recoder
from recoder_b200.nn import DynamicAutoencoder, FactorizationModel, MatrixFactorization


def _dataset(users=20, items=30):
  rng = np.random.default_rng(0)
  m = sp.random(users, items, density=0.2, format='csr', random_state=rng, data_rvs=lambda k: np.ones(k))
  return RecommendationDataset(m.astype(np.float32))


@pytest.mark.parametrize('milestones', [[3], [2, 5], [1, 2, 3], []])
def test_lr_schedule_matches_multisteplr(milestones):
  base, epochs = 0.01, 8
  p = torch.nn.Parameter(torch.zeros(1))
  opt = torch.optim.Adam([p], lr=base)
  sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=milestones, gamma=0.1, last_epoch=-1)  # model.py:328-330
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False)
  tr._base_lr, tr._lr_milestones = base, sorted(milestones)
  for epoch in range(1, epochs + 1):
    sched.step()                                   # stepped at the START of every epoch (model.py:364-366)
    assert tr._epoch_lr(epoch) == pytest.approx(opt.param_groups[0]['lr'], rel=1e-12), epoch
  tr._lr_milestones = None
  assert tr._epoch_lr(5) == base


def test_loss_selection_errors_before_touching_the_device():
  ds = _dataset()
  for bad, exc in ((None, ValueError), ('hinge', ValueError)):
    tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False, loss=bad)
    with pytest.raises(exc):                        # model.py:97, 99
      tr.train(ds, batch_size=4)
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False, loss='mse', loss_params={'confidense': 2})
  with pytest.raises(TypeError):                    # MSELoss(**loss_params) rejects unknown keywords (model.py:93)
    tr.train(ds, batch_size=4)
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False, loss='mse')
  with pytest.raises(RuntimeError, match='no CPU path'):
    tr.train(ds, batch_size=4)


def test_loss_resolution_fused_or_generic():
  spec = lambda loss, **kw: Recoder(model=DynamicAutoencoder(hidden_layers=[4]), loss=loss, **kw)._Recoder__loss_spec()  # noqa: E731
  assert spec('mse', loss_params={'confidence': 2})[:2] == ('mse', 2.0)
  assert spec('logloss')[0] == 'logloss' and spec('logistic')[0] == 'logistic'
  assert spec(MSELoss(confidence=1.5, reduction='sum'))[:2] == ('mse', 1.5)
  assert spec(MultinomialNLLLoss(reduction='sum'))[0] == 'logloss'
  assert spec(torch.nn.BCEWithLogitsLoss(reduction='sum'))[0] == 'logistic'
  # anything else is used as it is (recoder/model.py:88-89) through the generic path
  for module in (MSELoss(reduction='mean'), torch.nn.BCEWithLogitsLoss(reduction='sum', pos_weight=torch.ones(1)),
                 torch.nn.SmoothL1Loss(reduction='sum')):
    kind, _, got = spec(module)
    assert kind == 'custom' and got is module
  kind, _, module = spec('logistic', loss_params={'pos_weight': torch.ones(1)})
  assert kind == 'custom' and isinstance(module, torch.nn.BCEWithLogitsLoss) and module.reduction == 'sum'


def test_sampling_users_must_be_a_multiple_of_the_batch():
  ds = _dataset()
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False)
  with pytest.raises(AssertionError):               # model.py:310-311
    tr.train(ds, batch_size=4, num_sampling_users=6)
  with pytest.raises(AssertionError):
    tr.train(ds, batch_size=4, num_sampling_users=2)
  with pytest.raises(AssertionError):               # data.py:117
    RecommendationDataLoader(ds, batch_size=8, num_sampling_users=4)


def test_item_and_user_range_checks():
  ds = _dataset(users=20, items=30)
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False, num_items=10)
  with pytest.raises(AssertionError):               # model.py:241
    tr.train(ds, batch_size=4)
  tr = Recoder(model=MatrixFactorization(embedding_size=4), use_cuda=False, num_users=5)
  with pytest.raises(AssertionError):               # model.py:248
    tr.train(ds, batch_size=4)


def test_constructor_argument_validation():
  with pytest.raises(ValueError):
    Recoder(model=DynamicAutoencoder(hidden_layers=[4]), dp_exchange='ring')
  with pytest.raises(ValueError):
    Recoder(model=DynamicAutoencoder(hidden_layers=[4]), parallel='columns')
  for method, args in (('init_model', ()), ('model_params', ()), ('load_model_params', ({},)), ('forward', (None,))):
    with pytest.raises(NotImplementedError):        # nn.py:26, 36, 47, 65
      getattr(FactorizationModel(), method)(*args)
  with pytest.raises(Exception, match='No state file'):
    Recoder(model=DynamicAutoencoder()).init_from_model_file('/nonexistent/file.model')   # model.py:175


def test_optimizer_groups_follow_the_reference_rules():
  model = DynamicAutoencoder(hidden_layers=[8, 4])
  model.init_model(num_items=12)
  named = [(n, p.data) for n, p in model.named_parameters()]
  opt = Optimizer(named, 'adam', lr=0.1, weight_decay=0.01)
  for n, _ in named:                                # weight decay 0 for anything named *bias* (model.py:123-124)
    assert opt.states[n].weight_decay == (0 if 'bias' in n else 0.01), n
  assert [n for n, _ in named][:2] == ['en_embedding_layer.weight', '_DynamicAutoencoder__en_linear_embedding_layer.bias']
  with pytest.raises(Exception, match='Unknown optimizer kind'):                          # model.py:156
    Optimizer(named, 'lbfgs', 0.1, 0.0)
  for kind in ('sgd', 'adagrad', 'rmsprop'):        # sparse tables only work with Adam (model.py:142, 147, 152)
    with pytest.raises(ValueError, match='Sparse gradients'):
      Optimizer(named, kind, 0.1, 0.0, sparse_names=('en_embedding_layer.weight',))
  Optimizer(named, 'adam', 0.1, 0.0, sparse_names=('en_embedding_layer.weight',))


def test_loader_pool_and_slice_counts():
  ds = _dataset(users=23, items=30)
  order = np.arange(23)
  dl = RecommendationDataLoader(ds, batch_size=4, negative_sampling=True, num_sampling_users=8,
                                user_order=lambda e: order)
  assert len(dl) == 6                               # ceil(U / batch_size), data.py:166-167
  pools = list(dl.pools())
  assert [len(p) for p in pools] == [8, 8, 7]       # BatchSampler(num_sampling_users, drop_last=False), data.py:124-126
  assert np.array_equal(np.concatenate(pools), order)
  dl0 = RecommendationDataLoader(ds, batch_size=5)  # num_sampling_users = 0 -> batch_size (data.py:114-116)
  assert dl0.num_sampling_users == 5


def test_loss_modules_keep_the_reference_constructor():
  assert MSELoss(confidence=3, reduction='sum').confidence == 3
  assert MultinomialNLLLoss(reduction='sum').reduction == 'sum'
  x, t = torch.randn(3, 5), torch.rand(3, 5).round()
  want = ((1 + 3 * (t > 0).float()) * (x - t) ** 2).sum()
  assert torch.allclose(MSELoss(confidence=3, reduction='sum')(x, t), want)
  want = -(t * torch.log_softmax(x, dim=1)).sum()
  assert torch.allclose(MultinomialNLLLoss(reduction='sum')(x, t), want)


class _StubEngine:
  pg = None

  def __init__(self):
    self.steps_done = 0
    self.seen = []

  def train_step(self, pool, row0, rows, target_pool=None, global_rows=None):
    self.steps_done += 1
    self.seen.append(pool)

  def losses(self, k):
    return torch.arange(self.steps_done - k, self.steps_done, dtype=torch.float64)

  def join(self):
    pass


class _StubOptimizer:
  lr = 0.0
  flushes = 0

  def flush(self):
    self.flushes += 1


@pytest.mark.parametrize('iters_per_epoch,epochs,want', [
  (None, 2, [7, 7]),            # whole passes
  (3, 5, [3, 3, 1, 3, 3]),      # the pass continues across epochs; its ragged end is a short epoch, then a new pass
  (10, 2, [7, 7]),              # more than one pass per epoch is capped at a pass
  (7, 3, [7, 7, 7]),
])
def test_epoch_and_pass_bookkeeping_follows_the_reference(iters_per_epoch, epochs, want):
  """reference model.py:354-383,417-418: `iters_per_epoch` steps per epoch taken from ONE running iterator over the
  dataloader; a new pass starts only when the previous one was consumed.  Driven with a stub engine, so the host logic
  of `Recoder._train` runs on the CPU; every pass must see its steps in order, and the deferred optimizer is flushed at
  every epoch end."""
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False)
  ds = _dataset(users=7 * 4 - 1, items=30)            # 7 steps of 4 users, the last one ragged
  loader = RecommendationDataLoader(ds, batch_size=4)
  tr.engine, tr.optimizer, tr._ip = _StubEngine(), _StubOptimizer(), None
  tr._base_lr, tr._lr_milestones, tr._step_callback, tr._sync_loss_every_step = 0.01, None, None, False
  passes = []

  def pool_steps(dataloader, batch_size):
    passes.append(0)
    tr._host_timing = {'wait': 0.0, 'launch': 0.0, 'step': 0.0, 'n': 0}
    for i in range(len(dataloader)):
      passes[-1] += 1
      yield type('Pool', (), {'n': 5, 'tag': (len(passes), i)})(), None, 0, 4, 4
  tr._pool_steps = pool_steps
  per_epoch = []
  tr._step_callback = lambda done: per_epoch[-1].append(done)
  orig = tr._epoch_lr

  def epoch_lr(epoch):
    per_epoch.append([])
    return orig(epoch)
  tr._epoch_lr = epoch_lr
  tr._train(loader, None, num_epochs=epochs, current_epoch=1, batch_size=4, model_checkpoint_prefix=None,
            checkpoint_freq=0, eval_freq=0, metrics=None, eval_num_recommendations=None,
            iters_per_epoch=iters_per_epoch, eval_num_users=None, eval_batch_size=4)
  assert [len(e) for e in per_epoch] == want
  assert tr.optimizer.flushes == epochs and tr.current_epoch == epochs
  tags = [p.tag for p in tr.engine.seen]
  for k in range(1, len(passes) + 1):                  # every started pass delivers its steps 0, 1, 2, ... in order
    mine = [i for (pk, i) in tags if pk == k]
    assert mine == list(range(len(mine))) and (k == len(passes) or len(mine) == 7)
  assert len(tr.last_epoch_losses) == want[-1]


@pytest.mark.parametrize('depth', [1, 2, 3])
def test_pool_pipeline_order_and_slices(monkeypatch, depth):
  """`Recoder._pool_steps`: the collates of the next RCD_POOL_PIPELINE pools are enqueued before the steps of the
  current one, a pool is waited for only when its steps are about to be yielded, and every pool is cut into
  `batch_size` slices in order (reference data.py:138-144).  The collate itself is stubbed: host logic only."""
  import recoder_b200.model as M
  events = []

  class FakePool:
    def __init__(self, index):
      self.index, self.num_rows, self.n = index, len(index), 3

  def launch(csr, index, ns, **kw):
    assert kw['stream'] is None and kw['table_rows'] == 30 and kw['stage_shard'] is None
    events.append(('launch', int(index[0])))
    return FakePool(index)

  def finish(pool):
    events.append(('finish', int(pool.index[0])))
    return pool
  monkeypatch.setattr(M, 'collate_pool_launch', launch)
  monkeypatch.setattr(M, 'collate_pool_finish', finish)
  monkeypatch.setenv('RCD_OVERLAP', '0')
  monkeypatch.setenv('RCD_POOL_PIPELINE', str(depth))
  ds = _dataset(users=43, items=30)
  ds.device_csr = lambda: 'csr'
  ds.device_target_csr = lambda: None
  loader = RecommendationDataLoader(ds, batch_size=4, num_sampling_users=12, user_order=lambda e: np.arange(43))
  tr = Recoder(model=DynamicAutoencoder(hidden_layers=[4]), use_cuda=False)
  tr.engine, tr._ip, tr.num_items = None, None, 30
  got = []
  for pool, tpool, row0, rows, grows in tr._pool_steps(loader, 4):
    events.append(('step', int(pool.index[0]), row0))
    got.append((int(pool.index[0]), row0, rows, grows))
    assert tpool is None
  # 43 users in pools of 12: pools start at users 0, 12, 24, 36; the last has 7 rows -> slices of 4 and 3
  assert got == [(p, r, 4, 4) for p in (0, 12, 24) for r in (0, 4, 8)] + [(36, 0, 4, 4), (36, 4, 3, 3)]
  assert len(got) == len(loader)
  starts = [0, 12, 24, 36]
  for k, p in enumerate(starts):
    first_step = events.index(('step', p, 0))
    assert events.index(('finish', p)) < first_step
    for ahead in starts[k + 1:k + 1 + depth]:              # the next `depth` pools are already in flight
      assert events.index(('launch', ahead)) < first_step
    for later in starts[k + 1 + depth:]:                   # and nothing beyond them
      assert events.index(('launch', later)) > first_step
    if k + 1 < len(starts):                                # a pool is not waited for before its turn
      assert events.index(('finish', starts[k + 1])) > events.index(('step', p, 8 if k < 3 else 4))

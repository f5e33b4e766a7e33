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

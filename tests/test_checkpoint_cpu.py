"""Checkpoint interchange with the reference, CPU part: a checkpoint written by the UNMODIFIED
reference's `Recoder.save_state` (tests/golden/eval/ref_checkpoint_epoch_2.model, made by
tests/golden/make_checkpoint_fixture.py) must load into this implementation's containers — same `state_dict()` keys and
shapes, same `named_parameters()` order, optimizer state in torch.optim's layout — and survive a round trip."""
import os

import numpy as np
import torch

from recoder_b200.engine import Optimizer
from recoder_b200.nn import DynamicAutoencoder

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(HERE, 'golden', 'eval', 'ref_checkpoint_epoch_2.model')


def _load():
  return torch.load(CKPT, map_location='cpu', weights_only=False)


def test_reference_checkpoint_layout_is_understood():
  ck = _load()
  assert set(ck) >= {'recoder_version', 'model_params', 'last_epoch', 'model', 'optimizer_type', 'optimizer', 'items',
                     'users', 'num_items', 'num_users', 'loss', 'loss_params'}      # recoder/model.py:204-222
  model = DynamicAutoencoder()
  model.load_model_params(ck['model_params'])
  model.init_model(ck['num_items'], ck['num_users'])
  ours = model.state_dict()
  assert list(ours.keys()) == list(ck['model'].keys())
  for k, v in ck['model'].items():
    assert tuple(ours[k].shape) == tuple(v.shape), k
  assert model.load_state_dict(ck['model']).missing_keys == []
  for k, v in ck['model'].items():
    assert torch.equal(model.state_dict()[k], v), k
  assert model.model_params() == ck['model_params']


def test_reference_optimizer_state_round_trips():
  ck = _load()
  model = DynamicAutoencoder()
  model.load_model_params(ck['model_params'])
  model.init_model(ck['num_items'], ck['num_users'])
  named = [(n, p.data) for n, p in model.named_parameters()]
  ref_opt = ck['optimizer']
  assert len(ref_opt['param_groups']) == len(named)                                # one group per parameter, in order
  opt = Optimizer(named, ck['optimizer_type'], lr=0.5, weight_decay=ref_opt['param_groups'][0]['weight_decay'])
  opt.load_state_dict(ref_opt, dense=True)
  assert opt.lr == ref_opt['param_groups'][0]['lr']
  back = opt.state_dict(dense=True)
  assert set(back['state'].keys()) == set(ref_opt['state'].keys())
  for i, entry in ref_opt['state'].items():
    assert float(back['state'][i]['step']) == float(entry['step'])
    assert torch.equal(back['state'][i]['exp_avg'], entry['exp_avg'])
    assert torch.equal(back['state'][i]['exp_avg_sq'], entry['exp_avg_sq'])
  for g_ref, g in zip(ref_opt['param_groups'], back['param_groups']):
    for key in ('lr', 'weight_decay', 'betas', 'eps', 'params'):
      assert g[key] == g_ref[key] or np.allclose(g[key], g_ref[key]), key
  # bias groups carry no weight decay (recoder/model.py:123-124)
  names = [n for n, _ in named]
  for n, g in zip(names, back['param_groups']):
    assert (g['weight_decay'] == 0) == ('bias' in n)


def test_resume_keeps_the_checkpoint_learning_rates():
  """torch's `optimizer.load_state_dict` restores every group's lr and MultiStepLR's `initial_lr` (reference
  model.py:158-164, 327-332): a resumed or continued run trains on with the rates it was saved with, whatever lr is
  passed to `train()`.  The fused Optimizer keeps both and writes `initial_lr` into its own state dicts."""
  import torch
  from recoder_b200.engine import Optimizer
  p = torch.zeros(6, 4)
  opt = Optimizer([('w', p)], 'adam', lr=1e-2, weight_decay=0.0)
  opt.base_lr, opt.lr = 1e-2, 1e-3          # after one milestone
  opt._ensure(opt.states['w'])
  opt.states['w'].step = 7
  sd = opt.state_dict(dense=True)
  assert sd['param_groups'][0]['lr'] == 1e-3 and sd['param_groups'][0]['initial_lr'] == 1e-2
  torch.optim.Adam([torch.nn.Parameter(p.clone())], lr=0.5).load_state_dict(
    {'state': {0: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in sd['state'][0].items()}},
     'param_groups': [dict(sd['param_groups'][0])]})   # torch accepts the layout
  fresh = Optimizer([('w', p.clone())], 'adam', lr=0.5, weight_decay=0.0)
  assert not fresh.resumed
  fresh.load_state_dict(sd, dense=True)
  assert fresh.resumed and fresh.lr == 1e-3 and fresh.base_lr == 1e-2 and fresh.states['w'].step == 7

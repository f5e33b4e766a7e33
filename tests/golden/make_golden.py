"""Generates tests/golden/*.npz from the UNMODIFIED reference (/root/reference, amoussawi/recoder @ a9ed3e8).

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case it drives the reference's own objects — `RecommendationDataset.__getitem__`
(recoder/data.py:50-61), `BatchCollator.collate` (data.py:203-251), `Recoder.__init_training`
(model.py:226-254), `Recoder.__compute_loss` (model.py:454-485), `loss.backward()` and the optimizer steps
(model.py:397-402) — on an explicit user order, and records the collate outputs, loss, dense gradients and
post-step parameters of each step.  Dropout/noise RNG streams cannot be matched across implementations
(SURVEY.md §7.1): for the cases that use them the keep masks the reference drew are recorded per step
(`noise_keep` at the stored non-zeros, `dropout_keep` [B, width]) and injected into the implementation under test.

    python tests/golden/make_golden.py [case names...]     (no names: regenerate everything)
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

CASES = [
  # name, model, loss, loss_params, optimizer, sparse, hidden/emb, act, wd, neg, batch, pool, values
  dict(name='ae_mse_adam', model='ae', loss='mse', loss_params={}, opt='adam', sparse=False, hidden=[16],
       act='tanh', wd=1e-2, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_mse_conf_ratings', model='ae', loss='mse', loss_params={'confidence': 3.0}, opt='adam', sparse=False,
       hidden=[16], act='tanh', wd=0.0, neg=True, batch=24, pool=24, ratings=True),
  dict(name='ae_nll_adam', model='ae', loss='logloss', loss_params={}, opt='adam', sparse=False, hidden=[16],
       act='tanh', wd=2e-5, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_bce_adam', model='ae', loss='logistic', loss_params={}, opt='adam', sparse=False, hidden=[16],
       act='tanh', wd=0.0, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_mse_sgd', model='ae', loss='mse', loss_params={}, opt='sgd', sparse=False, hidden=[16],
       act='tanh', wd=1e-3, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_nll_sparseadam', model='ae', loss='logloss', loss_params={}, opt='adam', sparse=True, hidden=[16],
       act='tanh', wd=2e-5, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_mse_noneg', model='ae', loss='mse', loss_params={}, opt='adam', sparse=False, hidden=[16],
       act='tanh', wd=0.0, neg=False, batch=24, pool=24, ratings=False),
  dict(name='ae_nll_pool', model='ae', loss='logloss', loss_params={}, opt='adam', sparse=False, hidden=[16],
       act='sigmoid', wd=0.0, neg=True, batch=16, pool=48, ratings=False),
  dict(name='ae_mse_deep', model='ae', loss='mse', loss_params={}, opt='adam', sparse=False, hidden=[16, 8],
       act='tanh', wd=1e-3, neg=True, batch=24, pool=24, ratings=False),
  dict(name='mf_mse_adam', model='mf', loss='mse', loss_params={}, opt='adam', sparse=False, hidden=16,
       act='none', wd=1e-2, neg=True, batch=24, pool=24, ratings=False),
  dict(name='mf_nll_sgd', model='mf', loss='logloss', loss_params={}, opt='sgd', sparse=False, hidden=16,
       act='tanh', wd=0.0, neg=True, batch=24, pool=24, ratings=False),
  # remaining optimizers, tied / deeper autoencoders, input noise and dropout (keep masks recorded)
  dict(name='ae_mse_adagrad', model='ae', loss='mse', loss_params={}, opt='adagrad', sparse=False, hidden=[16],
       act='tanh', wd=1e-3, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_nll_rmsprop', model='ae', loss='logloss', loss_params={}, opt='rmsprop', sparse=False, hidden=[16],
       act='tanh', wd=1e-4, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_nll_deep_tied', model='ae', loss='logloss', loss_params={}, opt='adam', sparse=False, hidden=[16, 8],
       act='tanh', wd=1e-4, neg=True, batch=24, pool=24, ratings=False, constrained=True),
  dict(name='ae_bce_deep3', model='ae', loss='logistic', loss_params={}, opt='adam', sparse=False, hidden=[16, 12, 8],
       act='sigmoid', wd=0.0, neg=True, batch=24, pool=24, ratings=False),
  dict(name='ae_nll_noise_dropout', model='ae', loss='logloss', loss_params={}, opt='adam', sparse=False, hidden=[16],
       act='tanh', wd=2e-5, neg=True, batch=24, pool=24, ratings=False, noise=0.5, dropout=0.25),
  dict(name='ae_mse_deep_dropout', model='ae', loss='mse', loss_params={}, opt='adam', sparse=False, hidden=[16, 8],
       act='tanh', wd=0.0, neg=True, batch=24, pool=24, ratings=True, noise=0.3, dropout=0.5),
  dict(name='mf_mse_dropout', model='mf', loss='mse', loss_params={}, opt='adam', sparse=False, hidden=16,
       act='tanh', wd=0.0, neg=True, batch=24, pool=24, ratings=False, dropout=0.4),
]

NUM_USERS, NUM_ITEMS, NNZ = 80, 101, 9
LR = 1e-2


def make_matrix(seed, ratings):
  rng = np.random.default_rng(seed)
  rows, cols = [], []
  for u in range(NUM_USERS):
    k = int(rng.integers(1, 2 * NNZ))
    items = np.unique(np.minimum((NUM_ITEMS * rng.random(k) ** 2).astype(np.int64), NUM_ITEMS - 1))
    rows.extend([u] * len(items))
    cols.extend(items.tolist())
  vals = rng.integers(1, 6, len(rows)).astype(np.float32) if ratings else np.ones(len(rows), dtype=np.float32)
  m = sp.coo_matrix((vals, (rows, cols)), shape=(NUM_USERS, NUM_ITEMS)).tocsr()
  m.sort_indices()
  return m


def run_case(case, rdata, rnn, rmodel):
  torch.manual_seed(1234)
  csr = make_matrix(7, case['ratings'])
  dataset = rdata.RecommendationDataset(csr)
  noise, dropout = float(case.get('noise', 0.0)), float(case.get('dropout', 0.0))
  if case['model'] == 'ae':
    model = rnn.DynamicAutoencoder(hidden_layers=case['hidden'], activation_type=case['act'], sparse=case['sparse'],
                                   is_constrained=bool(case.get('constrained', False)), noise_prob=noise,
                                   dropout_prob=dropout)
  else:
    model = rnn.MatrixFactorization(embedding_size=case['hidden'], activation_type=case['act'], sparse=case['sparse'],
                                    dropout_prob=dropout)
  model.train()
  trainer = rmodel.Recoder(model=model, use_cuda=False, optimizer_type=case['opt'], loss=case['loss'],
                           loss_params=case['loss_params'])
  trainer._Recoder__init_training(train_dataset=dataset, lr=LR, weight_decay=case['wd'])
  # give biases non-zero values so bias paths are actually exercised
  with torch.no_grad():
    for name, p in model.named_parameters():
      if 'bias' in name:
        p.copy_(torch.randn(p.shape) * 0.1)
  out = {'meta': json.dumps({**case, 'lr': LR, 'num_users': NUM_USERS, 'num_items': NUM_ITEMS}),
         'csr_indptr': csr.indptr.astype(np.int64), 'csr_indices': csr.indices.astype(np.int32),
         'csr_data': csr.data.astype(np.float32)}
  names = [n for n, _ in model.named_parameters()]
  out['param_names'] = np.array(names)
  for n, p in model.named_parameters():
    out['init/' + n] = p.detach().numpy().copy()
  order = torch.randperm(NUM_USERS, generator=torch.Generator().manual_seed(1)).numpy()
  out['user_order'] = order.astype(np.int64)
  collator = rdata.BatchCollator(batch_size=case['batch'], negative_sampling=case['neg'])
  step = 0
  for off in range(0, NUM_USERS, case['pool']):
    pool_users = order[off:off + case['pool']]
    ui, _ = dataset[pool_users]
    batches = collator.collate(ui)
    for b in batches:
      if trainer.optimizer is not None:
        trainer.optimizer.zero_grad()
      if trainer.sparse_optimizer is not None:
        trainer.sparse_optimizer.zero_grad()
      pre = 'step%d/' % step
      if noise > 0 or dropout > 0:
        # nn.Dropout draws from the global CPU generator: replay the draws the forward is about to make (same shapes,
        # same order: input noise on the dense [B, n] input, then the bottleneck / user-embedding dropout) to record
        # the keep masks, then rewind the generator so the reference's forward sees exactly those draws.
        torch.manual_seed(1000 + step)
        B_, n_ = int(b.size[0]), int(b.size[1])
        if noise > 0:
          keep = torch.nn.functional.dropout(torch.ones(B_, n_), noise, True) != 0
          idx = b.indices.numpy()
          out[pre + 'noise_keep'] = keep.numpy()[idx[0], idx[1]].astype(np.uint8)
        if dropout > 0:
          width = case['hidden'][-1] if case['model'] == 'ae' else case['hidden']
          keep = torch.nn.functional.dropout(torch.ones(B_, width), dropout, True) != 0
          out[pre + 'dropout_keep'] = keep.numpy().astype(np.uint8)
        torch.manual_seed(1000 + step)
      loss = trainer._Recoder__compute_loss(b, None)
      loss.backward()
      out[pre + 'users'] = b.users.numpy().astype(np.int64)
      out[pre + 'items'] = (b.items.numpy().astype(np.int64) if b.items is not None else np.zeros(0, dtype=np.int64))
      out[pre + 'has_items'] = np.array(b.items is not None)
      out[pre + 'indices'] = b.indices.numpy().astype(np.int64)
      out[pre + 'values'] = b.values.numpy().astype(np.float32)
      out[pre + 'size'] = np.array(list(b.size), dtype=np.int64)
      out[pre + 'loss'] = np.array(loss.item(), dtype=np.float64)
      for n, p in model.named_parameters():
        g = p.grad
        g = g.to_dense() if g.is_sparse else g
        out[pre + 'grad/' + n] = g.detach().numpy().copy()
      if trainer.optimizer is not None:
        trainer.optimizer.step()
      if trainer.sparse_optimizer is not None:
        trainer.sparse_optimizer.step()
      for n, p in model.named_parameters():
        out[pre + 'param/' + n] = p.detach().numpy().copy()
      step += 1
  out['num_steps'] = np.array(step)
  return out


def main():
  rdata, rnn, rlosses, rmodel = ref_shims.import_reference()
  import warnings
  warnings.simplefilter('ignore')
  only = set(sys.argv[1:])
  for case in CASES:
    if only and case['name'] not in only:
      continue
    out = run_case(case, rdata, rnn, rmodel)
    path = os.path.join(HERE, case['name'] + '.npz')
    np.savez_compressed(path, **out)
    print('%-24s steps=%d  %6.1f KB' % (case['name'], int(out['num_steps']), os.path.getsize(path) / 1024))


if __name__ == '__main__':
  main()

"""Generates the evaluation fixtures from the UNMODIFIED reference (build container only):

  tests/golden/eval/ml_fixture.npz   the train / validation interaction matrices of the reference's golden-metric test
                                (tests/test_model.py:18-36: tests/data/{train,val}.csv through
                                recoder.utils.dataframe_to_csr_matrix, validation items restricted to training items),
                                stored as CSR with uint16 item ids (all values are 1)
  tests/golden/eval/eval_golden.npz  `Recoder.recommend` (model.py:525-544) + Recall/NDCG/AP (metrics.py) of the reference on
                                a seeded random DynamicAutoencoder and MatrixFactorization: the inputs, the parameters,
                                the top-k lists and the per-user metric values

    python tests/golden/make_eval_fixture.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

REF = '/root/reference'


def ml_fixture():
  import pandas as pd
  ref_shims.import_reference()
  from recoder.utils import dataframe_to_csr_matrix
  train_df = pd.read_csv(os.path.join(REF, 'tests/data/train.csv'))
  val_df = pd.read_csv(os.path.join(REF, 'tests/data/val.csv'))
  val_df = val_df[val_df.sid.isin(train_df.sid.unique())]
  train, item_map, user_map = dataframe_to_csr_matrix(train_df, user_col='uid', item_col='sid', inter_col='watched')
  val, _, _ = dataframe_to_csr_matrix(val_df, user_col='uid', item_col='sid', inter_col='watched',
                                      item_id_map=item_map, user_id_map=user_map)
  out = {}
  for name, m in (('train', train), ('val', val)):
    m = m.tocsr()
    m.sum_duplicates()
    m.sort_indices()
    assert m.shape[1] < 65536 and np.all(m.data == 1)
    out[name + '_indptr'] = m.indptr.astype(np.int32)
    out[name + '_indices'] = m.indices.astype(np.uint16)
    out[name + '_shape'] = np.array(m.shape, dtype=np.int64)
  path = os.path.join(HERE, 'eval', 'ml_fixture.npz')
  np.savez_compressed(path, **out)
  print('ml_fixture: train %s nnz %d, val %s nnz %d, %.0f KB' % (train.shape, train.nnz, val.shape, val.nnz,
                                                                 os.path.getsize(path) / 1024))


def eval_golden():
  rdata, rnn, rlosses, rmodel = ref_shims.import_reference()
  from recoder.metrics import AveragePrecision, NDCG, Recall, RecommenderEvaluator
  from recoder.recommender import InferenceRecommender
  from recoder_b200.synth import synthetic_csr, to_scipy
  U, I, H, K = 300, 2000, 32, 50
  indptr, indices, data = synthetic_csr(U, I, 30, seed=5)
  inp = to_scipy(indptr, indices, data, I)
  tptr, tidx, tdat = synthetic_csr(U, I, 20, seed=6)
  tgt = to_scipy(tptr, tidx, tdat, I)
  out = {'in_indptr': indptr, 'in_indices': indices, 'in_data': data, 'tg_indptr': tptr, 'tg_indices': tidx,
         'tg_data': tdat, 'shape': np.array([U, I, H, K])}
  metrics = [Recall(k=20, normalize=True), Recall(k=50, normalize=False), NDCG(k=50), AveragePrecision(k=10)]
  for kind in ('ae', 'ae2', 'mf'):
    torch.manual_seed(7)
    if kind == 'ae':
      model = rnn.DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
    elif kind == 'ae2':
      model = rnn.DynamicAutoencoder(hidden_layers=[H, 16], activation_type='sigmoid')
    else:
      model = rnn.MatrixFactorization(embedding_size=H, activation_type='tanh')
    trainer = rmodel.Recoder(model=model, use_cuda=False, optimizer_type='adam', loss='mse')
    ds = rdata.RecommendationDataset(inp, tgt)
    trainer._Recoder__init_training(train_dataset=ds, lr=1e-3, weight_decay=0)
    with torch.no_grad():   # spread the logits so that near-ties are rare
      for n, p in model.named_parameters():
        p.copy_(torch.randn(p.shape) * (0.5 if p.dim() > 1 else 0.2))
    model.eval()
    users = np.arange(U)
    ui, ti = ds[users]
    recs = np.array(trainer.recommend(ui, K), dtype=np.int64)
    out[kind + '/recs'] = recs
    with torch.no_grad():
      scores = trainer.predict(ui)
      scores = scores[0] if isinstance(scores, tuple) else scores
    out[kind + '/rec_scores'] = np.take_along_axis(scores.numpy().astype(np.float32), recs, axis=1)
    for n, p in model.named_parameters():
      out[kind + '/param/' + n] = p.detach().numpy().copy()
    out[kind + '/param_names'] = np.array([n for n, _ in model.named_parameters()])
    per_user = {str(m): [] for m in metrics}
    for u in range(U):
      y = tgt[u].nonzero()[1]
      for m in metrics:
        per_user[str(m)].append(m.evaluate(recs[u], y))
    for k2, v in per_user.items():
      out[kind + '/metric/' + k2] = np.array(v, dtype=np.float64)
  path = os.path.join(HERE, 'eval', 'eval_golden.npz')
  np.savez_compressed(path, **out)
  print('eval_golden: %.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
  ml_fixture()
  eval_golden()

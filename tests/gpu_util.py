"""Shared helpers for the GPU parity tests (CUDA path vs the CPU oracle)."""
import numpy as np
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200.data import RecommendationDataset, collate_pool
from recoder_b200.engine import Optimizer, TrainEngine
from recoder_b200.nn import DynamicAutoencoder, MatrixFactorization
from recoder_b200.synth import to_scipy


def make_model(kind, num_items, num_users, hidden, act, params, sparse=False, constrained=False, noise=0.0,
               dropout=0.0):
  """Builds a recoder_b200 model on cuda:0 and loads `params` (reference state_dict names -> numpy/tensor)."""
  if kind == 'ae':
    model = DynamicAutoencoder(hidden_layers=hidden if isinstance(hidden, list) else [hidden],
                               activation_type=act, sparse=sparse, is_constrained=constrained, noise_prob=noise,
                               dropout_prob=dropout)
  else:
    model = MatrixFactorization(embedding_size=hidden, activation_type=act, sparse=sparse, dropout_prob=dropout)
  model.init_model(num_items=num_items, num_users=num_users)
  model = model.to('cuda')
  named = dict(model.named_parameters())
  with torch.no_grad():
    for k, v in params.items():
      named[k].copy_(torch.as_tensor(np.asarray(v)).to('cuda'))
  return model


def make_engine(model, loss, confidence, opt_type, lr, wd, gemm_engine):
  kind, roles, act, tied = model._engine_spec()
  named = [(n, p.data) for n, p in model.named_parameters()]
  opt = Optimizer(named, opt_type, lr, wd, sparse_names=model._sparse_param_names())
  if isinstance(loss, torch.nn.Module):   # user-supplied module: the generic (autograd) loss path
    return TrainEngine(kind, roles, 'custom', confidence, act, opt, gemm_engine=gemm_engine, tied=tied,
                       loss_module=loss.to('cuda'))
  return TrainEngine(kind, roles, loss, confidence, act, opt, gemm_engine=gemm_engine, tied=tied)


def inner_grads(eng, model):
  """{parameter name: gradient} of the inner dense layers of the last step (views into the gradient slab)."""
  lay = eng.last.get('inner_layout')
  if not lay or not lay['size']:
    return {}
  flat = eng.last['inner']
  out = {}
  for L, (o_w, o_b, out_f, in_f) in zip(eng.enc_layers, lay['enc']):
    out[L['w'][0]] = flat[o_w:o_w + out_f * in_f].view(out_f, in_f)
    out[L['b'][0]] = flat[o_b:o_b + out_f]
  for L, (o_w, o_b, out_f, in_f) in zip(eng.dec_layers, lay['dec']):
    if L['w'] is not None:
      out[L['w'][0]] = flat[o_w:o_w + out_f * in_f].view(out_f, in_f)
    out[L['b'][0]] = flat[o_b:o_b + out_f]
  return out


def device_dataset(indptr, indices, data, num_items):
  return RecommendationDataset(to_scipy(indptr, indices, data, num_items))


def rel_err(a, b):
  a = np.asarray(a, dtype=np.float64)
  b = np.asarray(b, dtype=np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def compact_oracle_grads(kind, grads, batch):
  """Oracle dense grads -> the compact row blocks the CUDA step produces."""
  items = batch.items if batch.items is not None else None
  sel = (lambda g: g.numpy()[items]) if items is not None else (lambda g: g.numpy())
  if kind == 'ae':
    return {'dWe': sel(grads[O.AE_EN_W]), 'dWd': sel(grads[O.AE_DE_W]), 'dbd': sel(grads[O.AE_DE_B]),
            'dbe': grads[O.AE_EN_B].numpy()}
  return {'dV': sel(grads[O.MF_ITEM_W]), 'dbias': sel(grads[O.MF_BIAS]), 'dU': grads[O.MF_USER_W].numpy()[batch.users]}

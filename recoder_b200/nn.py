"""Factorization models with the reference's interface (recoder/nn.py), running on the recoder_b200 kernels.

`FactorizationModel`, `DynamicAutoencoder` and `MatrixFactorization` keep the reference's constructor
arguments, `init_model / model_params / load_model_params / forward` methods, attribute names and — so that
checkpoints interchange with the reference — the exact `state_dict()` keys and `named_parameters()` order
(SURVEY.md §8b).  `torch.nn.Module` / `nn.Parameter` are used purely as parameter containers: nothing here
calls `torch.nn.Linear`, `F.linear`, `nn.Embedding.forward` or autograd.  Training does not go through
`forward()` at all (see engine.TrainEngine); `forward()` is the inference path and also launches only
C-ABI kernels.

Attribution: the public interface of this module (class / method names, argument lists and their documentation, log
messages, checkpoint keys) mirrors amoussawi/recoder (MIT License, Copyright (c) 2018 Abdallah Moussawi) so that it is
a drop-in for that library; see LICENSE.  The implementation underneath is original.
"""
import math

import torch
from torch import nn

from . import _native
from ._native import call, ptr

SUPPORTED_ACTIVATIONS = tuple(_native.ACT_IDS.keys())


def _check_activation(act):
  if act not in _native.ACT_IDS:
    raise NotImplementedError("activation '%s' has no B200 kernel: the backward kernels keep the activation output "
                              "only, so torch.<name> functions whose derivative needs the pre-activation are not "
                              "covered; supported: %s" % (act, SUPPORTED_ACTIVATIONS))


def _xavier_uniform_(t):
  """nn.init.xavier_uniform_ on a 2-D tensor (recoder/nn.py:186,211,219): U(-a, a), a = sqrt(6/(fan_in+fan_out))."""
  bound = math.sqrt(6.0 / (t.shape[0] + t.shape[1]))
  with torch.no_grad():
    t.uniform_(-bound, bound)
  return t


class EmbeddingTable(nn.Module):
  """Parameter container standing where the reference has an `nn.Embedding` (same attribute names)."""

  def __init__(self, num_embeddings, embedding_dim, sparse=False):
    super().__init__()
    self.num_embeddings = num_embeddings
    self.embedding_dim = embedding_dim
    self.sparse = sparse
    self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim), requires_grad=False)


class DenseLayer(nn.Module):
  """Parameter container standing where the reference has an inner `nn.Linear` (weight [out, in], bias [out])."""

  def __init__(self, in_features, out_features):
    super().__init__()
    self.in_features = in_features
    self.out_features = out_features
    self.weight = nn.Parameter(torch.empty(out_features, in_features), requires_grad=False)
    self.bias = nn.Parameter(torch.empty(out_features), requires_grad=False)


class LinearEmbedding(nn.Module):
  """Container mirroring recoder/nn.py:256-267 (a bias plus an alias of the embedding table)."""

  def __init__(self, embedding_layer, input_based=True, bias=True):
    super().__init__()
    self.embedding_layer = embedding_layer
    self.input_based = input_based
    self.in_features = embedding_layer.num_embeddings if input_based else embedding_layer.embedding_dim
    self.out_features = embedding_layer.embedding_dim if input_based else embedding_layer.num_embeddings
    if bias:
      self.bias = nn.Parameter(torch.empty(self.out_features), requires_grad=False)
    else:
      self.bias = None


class FactorizationModel(nn.Module):
  """
  Base class for factorization models (reference recoder/nn.py:12-65). All subclasses should implement
  the following methods.
  """

  def init_model(self, num_items=None, num_users=None):
    """Initializes the model with the number of users and items to be represented."""
    raise NotImplementedError

  def model_params(self):
    """Returns the model hyper-parameters stored in a snapshot file by :class:`recoder_b200.model.Recoder`."""
    raise NotImplementedError

  def load_model_params(self, model_params):
    """Loads the ``model_params`` into the model."""
    raise NotImplementedError

  def forward(self, input, input_users=None, input_items=None, target_users=None, target_items=None):
    """Applies a forward pass of the input on the latent factor model (dense ``input`` as in the reference)."""
    raise NotImplementedError

  # hook used by Recoder: returns ('ae'|'mf', role->(name, tensor), activation, tied) or raises
  def _engine_spec(self):
    raise NotImplementedError


def _dense_input_to_csr(input):
  """K1': dense [B, n] fp32 CUDA tensor -> CSR of its non-zeros + row statistics."""
  _native.require_cuda()
  if not input.is_cuda:
    raise RuntimeError('recoder_b200 models run on CUDA tensors only (there is no CPU path)')
  x = input.contiguous().float()
  B, n = x.shape
  dev = x.device
  lib = _native.load()
  row_ptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
  cap = max(int(torch.count_nonzero(x).item()), 1)
  cols = torch.empty(cap, dtype=torch.int32, device=dev)
  vals = torch.empty(cap, dtype=torch.float32, device=dev)
  rin = torch.empty(B, dtype=torch.float32, device=dev)
  rsum = torch.empty(B, dtype=torch.float32, device=dev)
  nnz = torch.zeros(1, dtype=torch.int32, device=dev)
  sb = lib.rcd_dense_to_csr_scratch_bytes(B)
  scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
  call('rcd_dense_to_csr', ptr(x), B, n, n, cap, ptr(row_ptr), ptr(cols), ptr(vals), ptr(rin), ptr(rsum), ptr(nnz),
       ptr(scratch), sb)
  return row_ptr, cols, vals, rin, rsum


class DynamicAutoencoder(FactorizationModel):
  """
  An Autoencoder module that processes variable size vectors (reference recoder/nn.py:68-253): the encoder and
  decoder weights are item-embedding tables, gathered for the items present in a batch, which makes
  mini-batch negative sampling cheap.

  Args:
    hidden_layers (list): autoencoder hidden layers sizes. only the encoder layers.
    activation_type (str, optional): activation function to use for hidden layers ('tanh', 'sigmoid',
      'relu' or 'none' have B200 kernels).
    is_constrained (bool, optional): constraining model by using the encoder weights in the
      decoder (tying the weights).
    dropout_prob (float, optional): dropout probability at the bottleneck layer
    noise_prob (float, optional): dropout (noise) probability at the input layer
    sparse (bool, optional): if True, the embedding tables are updated by the row-sparse Adam
      (``torch.optim.SparseAdam`` semantics) instead of the dense optimizer.
  """

  def __init__(self, hidden_layers=None, activation_type='tanh', is_constrained=False, dropout_prob=0.0,
               noise_prob=0.0, sparse=False):
    super().__init__()
    self.activation_type = activation_type
    self.is_constrained = is_constrained
    self.hidden_layers = hidden_layers
    self.dropout_prob = dropout_prob
    self.noise_prob = noise_prob
    self.sparse = sparse

    self.num_items = None
    self.num_embeddings = None
    self.noise_layer = None
    self.dropout_layer = None

  def init_model(self, num_items=None, num_users=None):
    self.num_items = num_items
    self.num_embeddings = num_items
    _check_activation(self.activation_type)
    self.__create_encoding_layers()
    self.__create_decoding_layers()

  def model_params(self):
    return {
      'hidden_layers': self.hidden_layers,
      'activation_type': self.activation_type,
      'is_constrained': self.is_constrained,
      'dropout_prob': self.dropout_prob,
      'noise_prob': self.noise_prob
    }

  def load_model_params(self, model_params):
    self.hidden_layers = model_params['hidden_layers']
    self.activation_type = model_params['activation_type']
    self.is_constrained = model_params['is_constrained']
    self.dropout_prob = model_params['dropout_prob']
    self.noise_prob = model_params['noise_prob']

  # registration order == reference (nn.py:179-187, 189-220) so named_parameters()/state_dict() match
  def __create_encoding_layers(self):
    self.en_embedding_layer = EmbeddingTable(self.num_embeddings, self.hidden_layers[0], sparse=self.sparse)
    self.__en_linear_embedding_layer = LinearEmbedding(self.en_embedding_layer, input_based=True)
    self.encoding_layers = nn.Sequential(*self.__create_coding_layers(self.hidden_layers))
    _xavier_uniform_(self.en_embedding_layer.weight)
    nn.init.constant_(self.__en_linear_embedding_layer.bias, 0)

  def __create_decoding_layers(self):
    _decoding_layers = self.__create_coding_layers(list(reversed(self.hidden_layers)))
    if self.is_constrained:
      for decoding_layer in _decoding_layers:
        del decoding_layer.weight           # only the decoding biases are parameters (nn.py:192-195)
      self.de_embedding_layer = self.en_embedding_layer
    else:
      self.de_embedding_layer = EmbeddingTable(self.num_embeddings, self.hidden_layers[0], sparse=self.sparse)
    self.decoding_layers = nn.Sequential(*_decoding_layers)
    self.__de_linear_embedding_layer = LinearEmbedding(self.de_embedding_layer, input_based=False)
    _xavier_uniform_(self.de_embedding_layer.weight)
    nn.init.constant_(self.__de_linear_embedding_layer.bias, 0)

  def __create_coding_layers(self, layer_sizes):
    layers = []
    for ind, layer_size in enumerate(layer_sizes[1:], 1):
      layer = DenseLayer(layer_sizes[ind - 1], layer_size)
      layers.append(layer)
      _xavier_uniform_(layer.weight)
      nn.init.constant_(layer.bias, 0)
    return layers

  @property
  def en_bias(self):
    return self.__en_linear_embedding_layer.bias

  @property
  def de_bias(self):
    return self.__de_linear_embedding_layer.bias

  def _engine_spec(self):
    sd_names = {id(p): n for n, p in self.named_parameters()}
    named = lambda p: (sd_names[id(p)], p.data)  # noqa: E731
    roles = {
      'en_w': named(self.en_embedding_layer.weight),
      'en_b': named(self.en_bias),
      'de_w': named(self.de_embedding_layer.weight),
      'de_b': named(self.de_bias),
      # inner dense layers (nn.py:189-226): a tied decoding layer has no weight of its own ('w': None)
      'enc_layers': [{'w': named(l.weight), 'b': named(l.bias)} for l in self.encoding_layers],
      'dec_layers': [{'w': None if self.is_constrained else named(l.weight), 'b': named(l.bias)}
                     for l in self.decoding_layers],
      'noise_prob': float(self.noise_prob),
      'dropout_prob': float(self.dropout_prob),
    }
    return 'ae', roles, self.activation_type, self.is_constrained

  def _sparse_param_names(self):
    if not self.sparse:
      return ()
    names = {id(p): n for n, p in self.named_parameters()}
    return tuple({names[id(self.en_embedding_layer.weight)], names[id(self.de_embedding_layer.weight)]})

  def forward(self, input, input_users=None, input_items=None, target_users=None, target_items=None):
    """Inference forward on a dense ``[B, n]`` CUDA input (reference nn.py:228-253), fp32 logits out.
    Noise / dropout layers are identities outside training, as in the reference's eval mode."""
    row_ptr, cols, vals, rin, _ = _dense_input_to_csr(input)
    if input_items is not None:
      raw = input_items.to(input.device)[cols.long()].to(torch.int32)
    else:
      raw = cols
    return self._forward_csr(row_ptr, raw, vals, rin, 0, input.shape[0], target_items)

  def forward_pool(self, pool, target_items=None):
    """Inference forward straight from a collated pool (data.PoolBatch, negative_sampling=False): logits fp32
    [rows, num_items] — the dense [B, I] input of the reference's `predict` (model.py:502-510) never exists."""
    return self._forward_csr(pool.row_ptr, pool.raw_items, pool.vals, pool.row_inv_norm, 0, pool.num_rows, target_items)

  def _forward_csr(self, row_ptr, raw, vals, rin, row0, B, target_items):
    We, Wd = self.en_embedding_layer.weight.data, self.de_embedding_layer.weight.data
    dev = We.device
    H = We.shape[1]
    ldh = (H + 7) // 8 * 8
    Z = torch.empty(B, H, dtype=torch.float32, device=dev)
    Zb = torch.empty(B, ldh, dtype=torch.bfloat16, device=dev)
    act = _native.ACT_IDS[self.activation_type]
    call('rcd_ae_encoder_fwd', ptr(We), H, ptr(self.en_bias.data), ptr(row_ptr), ptr(raw), ptr(vals), ptr(rin), row0, B,
         act, ptr(Z), ptr(Zb), ldh)
    if len(self.encoding_layers):
      # inner encoding / decoding layers, activation after every one (nn.py:242-249); dropout is off in eval mode
      z = Z
      enc = list(self.encoding_layers)
      for layer in enc:
        W = layer.weight.data
        y = torch.empty(B, W.shape[0], dtype=torch.float32, device=dev)
        call('rcd_sgemm', 0, 1, B, W.shape[0], W.shape[1], ptr(z), W.shape[1], ptr(W), W.shape[1], ptr(y), W.shape[0],
             ptr(layer.bias.data), act, 0)
        z = y
      for j, layer in enumerate(self.decoding_layers):
        if self.is_constrained:   # weight = encoding layer^T (nn.py:224-226): y = z @ W_enc
          W = enc[len(enc) - 1 - j].weight.data
          y = torch.empty(B, W.shape[1], dtype=torch.float32, device=dev)
          call('rcd_sgemm', 0, 0, B, W.shape[1], W.shape[0], ptr(z), W.shape[0], ptr(W), W.shape[1], ptr(y), W.shape[1],
               ptr(layer.bias.data), act, 0)
        else:
          W = layer.weight.data
          y = torch.empty(B, W.shape[0], dtype=torch.float32, device=dev)
          call('rcd_sgemm', 0, 1, B, W.shape[0], W.shape[1], ptr(z), W.shape[1], ptr(W), W.shape[1], ptr(y), W.shape[0],
               ptr(layer.bias.data), act, 0)
        z = y
      call('rcd_f32_to_bf16_rows', ptr(z), B, H, ptr(Zb), ldh)
    return _decode_all(Zb, ldh, Wd, self.de_bias.data, target_items, B, H)


def _decode_all(Zb, ldh, table, bias, target_items, B, H):
  """logits fp32 [B, m] = Zb @ table[target_items].T + bias[target_items] (all rows when target_items is None)."""
  dev = Zb.device
  ids = None if target_items is None else target_items.to(dev).to(torch.int64).contiguous()
  m = table.shape[0] if ids is None else ids.numel()
  Wg = torch.empty(m, ldh, dtype=torch.bfloat16, device=dev)
  bg = torch.empty(m, dtype=torch.float32, device=dev)
  call('rcd_gather_rows', ptr(table), H, ptr(ids), m, 0, ptr(Wg), ldh, None)
  call('rcd_gather_vec', ptr(bias), ptr(ids), m, ptr(bg))
  ldo = (m + 7) // 8 * 8
  out = torch.empty(B, ldo, dtype=torch.float32, device=dev)
  call('rcd_decoder_fwd', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bg), B, m, H, None, ptr(out), ldo, None, None,
       _native.GEMM_TCGEN05)
  return out[:, :m]


class MatrixFactorization(FactorizationModel):
  """
  Matrix Factorization model for collaborative filtering (reference recoder/nn.py:283-362).

  Args:
    embedding_size (int): embedding size (rank) of the latent factors of users and items
    activation_type (str, optional): activation function to be applied on the user embedding.
    dropout_prob (float, optional): dropout probability to be applied on the user embedding
    sparse (bool, optional): row-sparse Adam updates for the embedding tables.
  """

  def __init__(self, embedding_size, activation_type='none', dropout_prob=0, sparse=False):
    super().__init__()
    self.embedding_size = embedding_size
    self.activation_type = activation_type
    self.dropout_prob = dropout_prob

    self.num_users = None
    self.num_items = None
    self.user_embedding_layer = None
    self.item_embedding_layer = None
    self.bias = None
    self.dropout_layer = None
    self.sparse = sparse

  def init_model(self, num_items=None, num_users=None):
    self.num_users = num_users
    self.num_items = num_items
    _check_activation(self.activation_type)
    # registration order bias -> user -> item matches the reference's named_parameters()
    self.bias = nn.Parameter(torch.empty(self.num_items), requires_grad=False)
    self.user_embedding_layer = EmbeddingTable(self.num_users, self.embedding_size, sparse=self.sparse)
    self.item_embedding_layer = EmbeddingTable(self.num_items, self.embedding_size, sparse=self.sparse)
    _xavier_uniform_(self.user_embedding_layer.weight)
    _xavier_uniform_(self.item_embedding_layer.weight)
    nn.init.constant_(self.bias, 0)

  def model_params(self):
    return {
      'embedding_size': self.embedding_size,
      'activation_type': self.activation_type,
      'dropout_prob': self.dropout_prob,
    }

  def load_model_params(self, model_params):
    self.embedding_size = model_params['embedding_size']
    self.activation_type = model_params['activation_type']
    self.dropout_prob = model_params['dropout_prob']

  def _engine_spec(self):
    roles = {
      'bias': ('bias', self.bias.data),
      'user_w': ('user_embedding_layer.weight', self.user_embedding_layer.weight.data),
      'item_w': ('item_embedding_layer.weight', self.item_embedding_layer.weight.data),
      'dropout_prob': float(self.dropout_prob),
    }
    return 'mf', roles, self.activation_type, False

  def _sparse_param_names(self):
    return ('user_embedding_layer.weight', 'item_embedding_layer.weight') if self.sparse else ()

  def forward(self, input, input_users=None, input_items=None, target_users=None, target_items=None):
    """Inference forward (reference nn.py:344-362); ``input`` is unused there as well."""
    _native.require_cuda()
    U, V = self.user_embedding_layer.weight.data, self.item_embedding_layer.weight.data
    dev = U.device
    D = V.shape[1]
    ldd = (D + 7) // 8 * 8
    users = torch.as_tensor(input_users).to(dev).to(torch.int64).contiguous()
    B = users.numel()
    Ub = torch.empty(B, ldd, dtype=torch.bfloat16, device=dev)
    call('rcd_gather_rows', ptr(U), D, ptr(users), B, _native.ACT_IDS[self.activation_type], ptr(Ub), ldd, None)
    return _decode_all(Ub, ldd, V, self.bias.data, target_items, B, D)

  def forward_pool(self, pool, target_items=None):
    """Inference forward for the users of a collated pool (data.PoolBatch)."""
    return self.forward(None, input_users=pool.users, target_items=target_items)

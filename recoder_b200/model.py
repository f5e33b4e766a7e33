"""`Recoder` trainer with the reference's interface (recoder/model.py), driving the B200 step engine.

Same constructor / `train()` keyword arguments, optimizer and loss selection rules, epoch / iteration
bookkeeping, learning-rate milestones, checkpoint file layout and exceptions as the reference.  What changed
underneath: the interaction matrix lives in HBM, every mini-batch is collated by CUDA kernels, and one
training step is a fixed sequence of C-ABI kernel launches (engine.TrainEngine) instead of
`__compute_loss` → autograd → torch.optim (recoder/model.py:383-404, 454-485).  The loss is kept on the
device and read back only when the progress bar refreshes (the reference syncs with `loss.item()` every
step, model.py:404).

Attribution: the public interface of this module (class / method names, argument lists and their documentation, log
messages, checkpoint keys) mirrors amoussawi/recoder (MIT License, Copyright (c) 2018 Abdallah Moussawi) so that it is
a drop-in for that library; see LICENSE.  The implementation underneath is original.
"""
import logging
import os
import time

import numpy as np
import torch
from torch.nn import BCEWithLogitsLoss

from . import __version__
from . import _native
from .data import (RecommendationDataLoader, BatchCollator, collate_pool, collate_pool_launch, collate_pool_finish,
                   pool_of, PoolRing)
from .engine import Optimizer, TrainEngine, shard_rows
from .losses import MSELoss, MultinomialNLLLoss
from .nn import FactorizationModel

log = logging.getLogger('recoder_b200')

try:
  from tqdm import tqdm
except ImportError:  # pragma: no cover
  tqdm = None


class _NoBar:
  def __init__(self, *a, **k): pass
  def set_postfix(self, *a, **k): pass
  def update(self, *a, **k): pass
  def close(self): pass


def steps_per_pass(num_users, global_step_rows, world, rows_sharded):
  """Optimizer steps one pass over `num_users` users takes: the reference's `len(dataloader)` = ceil(U / batch)
  (recoder/data.py:166-167; pools are whole multiples of the step, so only the very last slice is ragged).  A
  row-parallel run gives every rank floor(rows / world) rows of a slice (`engine.shard_rows`): a last slice with fewer
  users than ranks is no step at all."""
  steps = -(-int(num_users) // int(global_step_rows))
  tail = int(num_users) % int(global_step_rows)
  if rows_sharded and 0 < tail < world:
    steps -= 1
  return steps


class Recoder(object):
  """
  Module to train/evaluate a recommendation :class:`recoder_b200.nn.FactorizationModel`.

  Args:
    model (FactorizationModel): the factorization model to train.
    num_items (int, optional): the number of items to represent. If None, it will
      be computed from the first training dataset passed to ``train()``.
    num_users (int, optional): the number of users to represent. If not provided, it will
      be computed from the first training dataset passed to ``train()``.
    optimizer_type (str, optional): optimizer type (one of 'sgd', 'adam', 'adagrad', 'rmsprop').
    loss (str or torch.nn.Module, optional): `mse` for ``MSELoss``, `logistic` for ``BCEWithLogitsLoss``,
      `logloss` for ``MultinomialNLLLoss``, or an instance of one of those modules.
    loss_params (dict, optional): loss function extra params based on loss module if ``loss`` is a ``str``.
    use_cuda (bool, optional): must be True to train — this implementation has no CPU path.
    user_based (bool, optional): raise on inconsistencies between model users and dataset users.
    item_based (bool, optional): raise on inconsistencies between model items and dataset items.
    process_group (optional, extension): torch.distributed group for data-parallel training; defaults to the
      WORLD group when torch.distributed is initialised with more than one rank.
    gemm_engine (optional, extension): _native.GEMM_TCGEN05 (default) or _native.GEMM_SIMT (validation).
    dp_exchange (optional, extension): how data-parallel ranks combine a step — 'p2p': embedding tables and the
      gradient slab live in CUDA-IPC peer memory and one fused kernel per table does reduce-scatter -> Adam ->
      all-gather over NVLink (dense Adam, untied weights); 'nccl': one all-reduce of the slab, then a full Adam
      pass on every rank; 'auto' (default): 'p2p' when it applies, else 'nccl'.
    parallel (optional, extension): what the ranks of a multi-GPU run split — 'rows' (default): the users of the
      global batch (data parallel, gradients exchanged as above); 'items': the item axis (embedding tables, optimizer
      state, matrix columns, decoder GEMM width; `itempar.py`) — every rank processes all rows of the global batch
      and only two [B_global, H] activations cross NVLink per step.  Single-hidden-layer autoencoders without
      noise / dropout / tied weights and with a dense optimizer; anything else falls back to 'rows'.
  """

  def __init__(self, model: FactorizationModel,
               num_items=None, num_users=None,
               optimizer_type='sgd', loss='mse',
               loss_params=None, use_cuda=False,
               user_based=True, item_based=True,
               process_group=None, gemm_engine=None, dp_exchange='auto', parallel='rows', lazy_adam='auto'):

    self.model = model
    self.num_items = num_items
    self.num_users = num_users
    self.optimizer_type = optimizer_type
    self.loss = loss
    self.loss_params = loss_params if loss_params else {}
    self.use_cuda = use_cuda
    self.user_based = user_based
    self.item_based = item_based
    self.process_group = process_group
    self.gemm_engine = gemm_engine
    if dp_exchange not in ('auto', 'p2p', 'nccl'):
      raise ValueError("dp_exchange must be 'auto', 'p2p' or 'nccl'")
    self.dp_exchange = dp_exchange
    if parallel not in ('rows', 'items'):
      raise ValueError("parallel must be 'rows' or 'items'")
    self.parallel = parallel
    # deferred dense Adam (engine.Optimizer.enable_lazy): 'auto' (per table, when the batches touch a small enough
    # share of its rows), True or False; RCD_LAZY_ADAM=0|1|auto overrides.  Results are bit-identical either way.
    env = os.environ.get('RCD_LAZY_ADAM')
    if env is not None:
      lazy_adam = {'0': False, '1': True}.get(env, 'auto')
    self.lazy_adam = lazy_adam
    self._p2p = None
    self._ip = None

    if self.use_cuda:
      self.device = torch.device('cuda')
    else:
      self.device = torch.device('cpu')

    self.optimizer = None
    self.sparse_optimizer = None  # kept for interface parity: the fused Optimizer handles both kinds
    self.engine = None
    self.current_epoch = 1
    self.items = None
    self.users = None
    self.__model_initialized = False
    self.__optimizer_state_dict = None
    self.__sparse_optimizer_state_dict = None

  # ---------------------------------------------------------------------------------------------------
  def __require_cuda(self):
    if self.device.type != 'cuda':
      raise RuntimeError('recoder_b200 has no CPU path: construct Recoder(..., use_cuda=True)')
    _native.require_cuda()
    _native.load()

  def __init_model(self):
    if self.__model_initialized:
      return
    self.model.init_model(self.num_items, self.num_users)
    self.model = self.model.to(device=self.device)
    self.__model_initialized = True

  def __loss_spec(self):
    """Resolves `self.loss` like `__init_loss_module` (reference model.py:87-99) into (kind, confidence, module).
    'mse' / 'logloss' / 'logistic' with sum reduction are fused into the decoder GEMM's epilogue; any other
    `nn.Module` (the reference accepts every module with sum reduction, model.py:36-37, 88-89) takes the generic
    path: dense fp32 logits from the tcgen05 GEMM, the module and its gradient through torch autograd, then the
    same backward kernels (kind 'custom')."""
    loss = self.loss
    if isinstance(loss, torch.nn.Module):
      if isinstance(loss, MSELoss) and loss.reduction == 'sum':
        return 'mse', float(loss.confidence), None
      if isinstance(loss, MultinomialNLLLoss) and loss.reduction == 'sum':
        return 'logloss', 0.0, None
      if (isinstance(loss, BCEWithLogitsLoss) and loss.reduction == 'sum' and loss.weight is None
          and loss.pos_weight is None):
        return 'logistic', 0.0, None
      return 'custom', 0.0, loss
    elif loss == 'logistic':
      if self.loss_params:
        return 'custom', 0.0, BCEWithLogitsLoss(reduction='sum', **self.loss_params)   # model.py:91
      return 'logistic', 0.0, None
    elif loss == 'mse':
      unknown = set(self.loss_params) - {'confidence'}
      if unknown:
        raise TypeError("__init__() got an unexpected keyword argument '%s'" % sorted(unknown)[0])
      return 'mse', float(self.loss_params.get('confidence', 0)), None
    elif loss == 'logloss':
      return 'logloss', 0.0, None
    elif loss is None:
      raise ValueError('No loss function defined')
    else:
      raise ValueError('Unknown loss function {}'.format(loss))

  def __init_optimizer(self, lr, weight_decay):
    # When continuing training on the same Recoder instance (reference model.py:103-107)
    if self.optimizer is not None:
      self.__optimizer_state_dict = self.optimizer.state_dict(dense=True)
      self.__sparse_optimizer_state_dict = self.optimizer.state_dict(dense=False)

    sparse_names = self.model._sparse_param_names()
    named = [(n, p.data) for n, p in self.model.named_parameters()]
    if self._ip is not None:   # item-parallel: the optimizer owns the local shards of the item-indexed tensors
      named = [(n, self._ip.sharded.get(n, t)) for n, t in named]
    self.optimizer = Optimizer(named, self.optimizer_type, lr, weight_decay, sparse_names=sparse_names)
    if self._ip is not None and self.__optimizer_state_dict is not None:
      self.__optimizer_state_dict = self.__shard_optimizer_state(self.__optimizer_state_dict)

    if self.__optimizer_state_dict is not None:
      # like torch's load_state_dict in the reference (model.py:158-160): the checkpoint's lr / initial_lr replace the
      # lr argument — continuing or resuming a run trains on with the rates it was saved with
      self.optimizer.load_state_dict(self.__optimizer_state_dict, dense=True)
      self.__optimizer_state_dict = None
    if self.__sparse_optimizer_state_dict is not None:
      self.optimizer.load_state_dict(self.__sparse_optimizer_state_dict, dense=False)
      self.__sparse_optimizer_state_dict = None

  def __resolve_pg(self):
    pg = self.process_group
    if pg is None:
      import torch.distributed as dist
      if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        pg = dist.group.WORLD
    return pg

  def __init_dp(self):
    """Data-parallel set-up (once): replicas start from rank 0's parameters; for the peer-memory exchange the
    embedding tables move into CUDA-IPC allocations every rank maps (`p2p.SharedBuffer`)."""
    pg = self.__resolve_pg()
    if pg is None or getattr(self, '_dp_ready', False):
      return
    import torch.distributed as dist
    if dist.get_world_size(pg) < 2:
      return
    for p in self.model.parameters():
      dist.broadcast(p.data, src=dist.get_global_rank(pg, 0), group=pg)
    self._dp_ready = True
    self._p2p_buffers = {}
    if self.parallel == 'items':
      kind, roles, _, tied = self.model._engine_spec()
      ok = (kind == 'ae' and not tied and not roles['enc_layers'] and roles['noise_prob'] == 0.0
            and roles['dropout_prob'] == 0.0 and not self.model._sparse_param_names()
            and self.__loss_spec()[0] != 'custom')
      if ok:
        from .itempar import ItemParallel
        from .p2p import P2PContext
        self._ip = ItemParallel(pg)
        if os.environ.get('RCD_IP_COLLECTIVES', 'p2p') != 'nccl' and P2PContext.available(pg):
          self._ip.p2p = P2PContext(pg)
        for role in ('en_w', 'de_w', 'de_b'):
          name, full = roles[role]
          self._ip.shard(name, full)
        return
      log.warning("parallel='items' does not cover this model configuration; falling back to 'rows'")
    if self.dp_exchange == 'nccl':
      return
    from .p2p import P2PContext
    _, roles, _, tied = self.model._engine_spec()
    applies = (self.optimizer_type == 'adam' and not tied and not self.model._sparse_param_names()
               and P2PContext.available(pg))
    if not applies:
      if self.dp_exchange == 'p2p':
        raise RuntimeError("dp_exchange='p2p' needs dense Adam, untied weights and one NCCL rank per GPU with "
                           'peer access between all GPUs')
      return
    self._p2p = P2PContext(pg)
    named = dict(self.model.named_parameters())
    for role in ('en_w', 'de_w', 'user_w', 'item_w'):
      if role in roles:
        name = roles[role][0]
        buf, view = self._p2p.shared_like(named[name].data)
        named[name].data = view
        self._p2p_buffers[name] = buf

  def __shard_optimizer_state(self, sd):
    """Full-size optimizer state tensors (a checkpoint) -> this rank's item shard."""
    names = [n for n, _ in self.model.named_parameters()]
    out = {'state': {}, 'param_groups': sd.get('param_groups')}
    for i, entry in sd['state'].items():
      name = names[i]
      if name in self._ip.sharded:
        entry = {k: (self._ip.shard_like(v) if torch.is_tensor(v) and v.dim() >= 1 and
                     v.shape[0] == dict(self.model.named_parameters())[name].shape[0] else v)
                 for k, v in entry.items()}
      out['state'][i] = entry
    return out

  def __full_optimizer_state(self, sd):
    """This rank's optimizer state with the item-sharded tensors gathered to full size (collective)."""
    names = [n for n, _ in self.model.named_parameters()]
    full_rows = {n: p.shape[0] for n, p in self.model.named_parameters()}
    for i, entry in sd['state'].items():
      name = names[i]
      if name in self._ip.sharded:
        for k, v in list(entry.items()):
          if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == self._ip.sharded[name].shape[0]:
            entry[k] = self._ip.gather_full(v.to(self.device), full_rows[name]).cpu()
    return sd

  def flush_parameters(self):
    """Deferred dense Adam: brings every table row up to date (a no-op when nothing is deferred).  Called before
    anything reads whole tables: evaluation, recommendations, checkpoints, the end of `train()`."""
    if self.optimizer is not None:
      self.optimizer.flush()

  def sync_parameters(self):
    """Item-parallel mode: gathers the item shards back into the model's (full-size) parameters — called before
    evaluation and checkpoints; collective.  A no-op otherwise."""
    self.flush_parameters()
    if self._ip is not None:
      if self.engine is not None:
        self.engine.join()
      self._ip.sync_to_full({n: p.data for n, p in self.model.named_parameters()})

  def __init_engine(self):
    kind, roles, activation, tied = self.model._engine_spec()
    if self._ip is not None:
      for role in ('en_w', 'de_w', 'de_b'):
        name, _ = roles[role]
        roles[role] = (name, self._ip.sharded[name])
    loss_kind, confidence, loss_module = self.__loss_spec()
    if loss_module is not None:
      loss_module = loss_module.to(self.device)
    pg = self.__resolve_pg()
    for name, buf in getattr(self, '_p2p_buffers', {}).items():
      self.optimizer.states[name].shared = buf
    self.engine = TrainEngine(kind, roles, loss_kind, confidence, activation, self.optimizer,
                              gemm_engine=self.gemm_engine, process_group=pg, tied=tied, p2p=self._p2p,
                              item_parallel=self._ip, loss_module=loss_module, lazy_adam=self.lazy_adam)
    if not getattr(self, '_flush_hook', None):
      # anything that serialises the model sees current rows
      self._flush_hook = self.model.register_state_dict_pre_hook(lambda *a, **k: self.flush_parameters())

  def init_from_model_file(self, model_file):
    """
    Initializes the model from a pre-trained model (reference model.py:166-191; same file layout).

    Args:
       model_file (str): the pre-trained model file path
    """
    log.info('Loading model from: {}'.format(model_file))
    if not os.path.isfile(model_file):
      raise Exception('No state file found in {}'.format(model_file))
    model_saved_state = torch.load(model_file, map_location='cpu', weights_only=False)
    model_params = model_saved_state['model_params']
    self.current_epoch = model_saved_state['last_epoch']
    self.loss = model_saved_state.get('loss', self.loss)
    self.loss_params = model_saved_state.get('loss_params', self.loss_params)
    self.optimizer_type = model_saved_state['optimizer_type']
    self.items = model_saved_state.get('items', None)
    self.users = model_saved_state.get('users', None)
    self.num_items = model_saved_state.get('num_items', None)
    self.num_users = model_saved_state.get('num_users', None)
    self.__optimizer_state_dict = model_saved_state['optimizer']
    self.__sparse_optimizer_state_dict = model_saved_state.get('sparse_optimizer', None)

    self.model.load_model_params(model_params)
    self.__init_model()
    self.model.load_state_dict(model_saved_state['model'])

  def save_state(self, model_checkpoint_prefix):
    """
    Saves the model state in the path starting with ``model_checkpoint_prefix`` and appending it
    with the model current training epoch (reference model.py:193-224; same keys).

    Returns:
      the model state file path
    """
    checkpoint_file = "{}_epoch_{}.model".format(model_checkpoint_prefix, self.current_epoch)
    log.info("Saving model to {}".format(checkpoint_file))
    if self._p2p is not None and self.optimizer is not None:
      self.optimizer.gather_shards(self._p2p)   # collective: every rank must call save_state
    self.sync_parameters()
    optimizer_state = self.optimizer.state_dict(dense=True)
    if self._ip is not None:
      optimizer_state = self.__full_optimizer_state(optimizer_state)
    current_state = {
      'recoder_version': __version__,
      'model_params': self.model.model_params(),
      'last_epoch': self.current_epoch,
      'model': {k: v.detach().cpu() for k, v in self.model.state_dict().items()},
      'optimizer_type': self.optimizer_type,
      'optimizer': optimizer_state,
      'items': self.items,
      'users': self.users,
      'num_items': self.num_items,
      'num_users': self.num_users
    }
    if any(s.sparse for s in self.optimizer.states.values()):
      # the reference reads this key on load (model.py:187) but never writes it; writing it loses nothing
      current_state['sparse_optimizer'] = self.optimizer.state_dict(dense=False)

    if type(self.loss) is str:
      current_state['loss'] = self.loss
      current_state['loss_params'] = self.loss_params

    if self._world()[1] == 0:
      torch.save(current_state, checkpoint_file)
    return checkpoint_file

  def __init_training(self, train_dataset, lr, weight_decay):
    if self.items is None:
      self.items = train_dataset.items
    else:
      self.items = np.unique(np.append(self.items, train_dataset.items))

    if self.users is None:
      self.users = train_dataset.users
    else:
      self.users = np.unique(np.append(self.users, train_dataset.users))

    if self.item_based and self.num_items is None:
      self.num_items = int(np.max(self.items)) + 1
    elif self.item_based:
      assert self.num_items >= int(np.max(self.items)) + 1, \
        'The largest item id should be smaller than number of items.' \
        'If your model is not based on items, set item_based to False in Recoder constructor.'

    if self.user_based and self.num_users is None:
      self.num_users = int(np.max(self.users)) + 1
    elif self.user_based:
      assert self.num_users >= int(np.max(self.users)) + 1, \
        'The largest user id should be smaller than number of users.' \
        'If your model is not based on users, set user_based to False in Recoder constructor.'

    self.__loss_spec()  # raises ValueError for unknown / missing losses before touching the device
    self.__require_cuda()
    self.__init_model()
    self.__init_dp()
    self.__init_optimizer(lr=lr, weight_decay=weight_decay)
    self.__init_engine()

  def _world(self):
    pg = self.engine.pg if self.engine is not None else None
    if pg is None:
      return 1, 0
    import torch.distributed as dist
    return dist.get_world_size(pg), dist.get_rank(pg)

  def train(self, train_dataset, val_dataset=None,
            lr=0.001, weight_decay=0, num_epochs=1,
            iters_per_epoch=None, batch_size=64, lr_milestones=None,
            negative_sampling=False, num_sampling_users=0, num_data_workers=0,
            model_checkpoint_prefix=None, checkpoint_freq=0,
            eval_freq=0, eval_num_recommendations=None,
            eval_num_users=None, metrics=None, eval_batch_size=None, user_order=None, step_callback=None,
            sync_loss_every_step=False):
    """
    Trains the model (reference model.py:256-347; same arguments).  In data-parallel runs ``batch_size`` is the
    per-rank batch: one optimizer step consumes ``batch_size * world_size`` users and is numerically the
    reference's step with that global batch size.  Extensions used by tests and benchmarks: ``user_order`` is an
    ``epoch -> index array`` callable replacing the random sampler; ``step_callback(step_count)`` is called after
    every optimizer step has been enqueued; ``sync_loss_every_step=True`` reads the loss back to the host after
    every step like the reference's ``loss.item()`` (model.py:404) instead of every 50 steps.
    """
    log.info('{} Mode'.format('CPU' if self.device.type == 'cpu' else 'GPU'))
    model_params = self.model.model_params()
    for param in model_params:
      log.info('Model {}: {}'.format(param, model_params[param]))
    log.info('Initial Learning Rate: {}'.format(lr))
    log.info('Weight decay: {}'.format(weight_decay))
    log.info('Batch Size: {}'.format(batch_size))
    log.info('Optimizer: {}'.format(self.optimizer_type))
    log.info('LR milestones: {}'.format(lr_milestones))
    log.info('Loss Function: {}'.format(self.loss))
    for param in self.loss_params:
      log.info('Loss {}: {}'.format(param, self.loss_params[param]))

    if num_sampling_users == 0:
      num_sampling_users = batch_size

    if eval_batch_size is None:
      eval_batch_size = batch_size

    assert num_sampling_users >= batch_size and num_sampling_users % batch_size == 0, \
      "number of sampling users should be a multiple of the batch size"

    self.__init_training(train_dataset=train_dataset, lr=lr, weight_decay=weight_decay)
    world, _ = self._world()

    train_dataloader = RecommendationDataLoader(train_dataset, batch_size=batch_size * world,
                                                negative_sampling=negative_sampling,
                                                num_sampling_users=num_sampling_users * world,
                                                num_workers=num_data_workers, user_order=user_order)
    if val_dataset is not None:
      val_dataloader = RecommendationDataLoader(val_dataset, batch_size=batch_size,
                                                negative_sampling=negative_sampling,
                                                num_sampling_users=num_sampling_users,
                                                num_workers=num_data_workers)
    else:
      val_dataloader = None

    self._base_lr = self.optimizer.base_lr if self.optimizer.resumed else lr
    self.optimizer.base_lr = self._base_lr
    self._lr_milestones = sorted(lr_milestones) if lr_milestones is not None else None
    self._step_callback = step_callback
    self._sync_loss_every_step = bool(sync_loss_every_step)

    self._train(train_dataloader=train_dataloader,
                val_dataloader=val_dataloader,
                num_epochs=num_epochs,
                current_epoch=self.current_epoch,
                batch_size=batch_size,
                model_checkpoint_prefix=model_checkpoint_prefix,
                checkpoint_freq=checkpoint_freq,
                eval_freq=eval_freq,
                metrics=metrics,
                eval_num_recommendations=eval_num_recommendations,
                iters_per_epoch=iters_per_epoch,
                eval_num_users=eval_num_users,
                eval_batch_size=eval_batch_size)

  def _epoch_lr(self, epoch):
    """MultiStepLR(gamma=0.1) stepped at every epoch start (reference model.py:327-332, 364-366):
    during epoch e the dense optimizer runs at lr * 0.1 ** #{milestones <= e}."""
    if self._lr_milestones is None:
      # no scheduler: the optimizer's own rate (the checkpoint's after a resume)
      return self.optimizer.lr if self.optimizer is not None else self._base_lr
    k = sum(1 for m in self._lr_milestones if m <= epoch)
    return self._base_lr * (0.1 ** k)

  def _pool_steps(self, dataloader, batch_size):
    """Generator over (pool, target_pool, row0, rows, global_rows): one item per optimizer step, in the order the
    reference's `_default_data_generator` yields slices (data.py:138-144)."""
    world, rank = self._world()
    ds = dataloader.dataset
    if self._ip is not None:
      if ds.target_interactions_matrix is not None:
        raise NotImplementedError("parallel='items' trains on datasets whose input is their own target")
      csr, tcsr = ds.item_shard_csr(rank, world), None
    else:
      csr = ds.device_csr()
      tcsr = ds.device_target_csr()
    gstep = batch_size * world
    ns = dataloader.negative_sampling

    # the collate of the NEXT pool runs on its own stream, underneath the current pool's training steps
    aux = torch.cuda.Stream() if os.environ.get('RCD_OVERLAP', '1') != '0' else None

    ring, tring = PoolRing(), PoolRing()

    # the model may represent more items than the matrix has columns (reference model.py:241)
    table_rows = self.num_items if (self._ip is None or self.num_items is None) else \
      self._ip.local_rows(self.num_items)

    # row-parallel runs on a host-resident matrix: every rank stages its own block of the pool, NVLink does the rest
    shard = (rank, world, self.engine.pg) if (world > 1 and self._ip is None) else None

    def launch(index):
      after = (self.engine._side,) if self.engine is not None else ()
      pool = collate_pool_launch(csr, index, ns, stream=aux, after=after, ring=ring, table_rows=table_rows,
                                 stage_shard=shard)
      tpool = collate_pool_launch(tcsr, index, ns, stream=aux, after=after, ring=tring, table_rows=table_rows,
                                  stage_shard=shard) if tcsr is not None else None
      return pool, tpool

    # Software pipeline, RCD_POOL_PIPELINE pools deep (default 2): the collates of pools i+1 and i+2 are enqueued BEFORE
    # the training steps of pool i.  The (n, nnz) read-back of a pool — the only host sync of the data path — then
    # completes a whole pool before the host asks for it, so the host never waits for the GPU to reach a collate and
    # can run a step ahead of the device even when enqueueing a pool takes most of a step (8 ranks staging 16 K rows
    # each from host memory: 1.4 ms of host time per pool against a 2.6 ms step, profiles/README.md r02u).  With one
    # pool in flight the read-back sat behind the previous step: host wait + enqueue time was the step time.
    import collections
    depth = max(1, int(os.environ.get('RCD_POOL_PIPELINE', '2')))
    pools = iter(dataloader.pools())
    queue = collections.deque()

    def fill():
      while len(queue) < depth:
        index = next(pools, None)
        if index is None:
          return
        queue.append(launch(index))

    fill()
    import time as _time
    ht = self._host_timing = getattr(self, '_host_timing', {'wait': 0.0, 'launch': 0.0, 'step': 0.0, 'n': 0})
    while queue:
      pool, tpool = queue.popleft()
      t0 = _time.perf_counter()
      collate_pool_finish(pool)
      if tpool is not None:
        collate_pool_finish(tpool)
      t1 = _time.perf_counter()
      fill()
      nxt = queue[0] if queue else None
      ht['wait'] += t1 - t0                      # blocked on the GPU (counts of the collated pool)
      ht['launch'] += _time.perf_counter() - t1  # host time to enqueue the next pool's collate
      if self._ip is not None:   # every rank takes all rows of the global slice; the item axis is what is split
        slices = [(goff, min(gstep, pool.num_rows - goff), min(gstep, pool.num_rows - goff))
                  for goff in range(0, pool.num_rows, gstep)]
        tpool = None
      else:
        slices = list(shard_rows(pool.num_rows, gstep, world, rank))
      if slices:
        # the engine may bring the NEXT pool's table rows up to date underneath the last step of this pool
        pool.next_hint = nxt
        pool.last_slice_row0 = slices[-1][0]
      for row0, rows, global_rows in slices:
        yield pool, tpool, row0, rows, global_rows

  def _train(self, train_dataloader, val_dataloader,
             num_epochs, current_epoch,
             batch_size, model_checkpoint_prefix, checkpoint_freq,
             eval_freq, metrics, eval_num_recommendations, iters_per_epoch,
             eval_num_users, eval_batch_size):
    world, rank = self._world()
    # optimizer steps of one pass over the data: len(dataloader), minus the step row-parallel runs skip when the last
    # slice has fewer users than there are ranks (engine.shard_rows)
    num_batches = steps_per_pass(len(train_dataloader.dataset), batch_size * world, world,
                                 rows_sharded=world > 1 and self._ip is None)

    iters_processed = 0
    if iters_per_epoch is None:
      iters_per_epoch = num_batches
    refresh_every = 50
    iterator = None

    for epoch in range(current_epoch, num_epochs + 1):
      self.current_epoch = epoch
      self.model.train()
      lr = self._epoch_lr(epoch)
      self.optimizer.lr = lr
      description = 'Epoch {}/{} (lr={})'.format(epoch, num_epochs, lr)

      if iters_processed == 0 or iters_processed == num_batches:
        # starting from scratch, or the whole dataloader was consumed: new pass (reference model.py:371-376)
        iters_processed = 0
        iterator = enumerate(self._pool_steps(train_dataloader, batch_size), 1)

      iters_to_process = min(iters_per_epoch, num_batches - iters_processed)
      iters_processed += iters_to_process

      progress_bar = tqdm(range(iters_to_process), desc=description) if (tqdm is not None and rank == 0) \
        else _NoBar()
      steps_this_epoch = 0
      last_loss = None
      num_items = None
      for batch_itr, (pool, tpool, row0, rows, global_rows) in iterator:
        _t0 = time.perf_counter()
        self.engine.train_step(pool, row0, rows, target_pool=tpool, global_rows=global_rows)
        self._host_timing['step'] += time.perf_counter() - _t0   # host time to enqueue one step
        self._host_timing['n'] += 1
        steps_this_epoch += 1
        num_items = (tpool or pool).n
        if self._sync_loss_every_step:
          # the reference's per-step `loss.item()` (model.py:404), read one step late so that the host never waits
          # for the step it has just enqueued
          prev = self.engine.loss_to_host_deferred()
          last_loss = prev if prev is not None else last_loss
        if steps_this_epoch % refresh_every == 0:
          if not self._sync_loss_every_step:
            last_loss = float(self.engine.losses(1)[0])
          progress_bar.set_postfix(loss=last_loss, num_items=num_items, refresh=False)
        if self._step_callback is not None:
          self._step_callback(self.engine.steps_done)
        progress_bar.update()
        if batch_itr % iters_per_epoch == 0:
          break

      if self._sync_loss_every_step:
        self.engine.drain_deferred_loss()
      self.flush_parameters()     # the model's parameters are current at every epoch boundary
      if self._ip is not None and ((eval_freq > 0 and epoch % eval_freq == 0) or epoch == num_epochs or
                                   (checkpoint_freq > 0 and epoch % checkpoint_freq == 0)):
        self.sync_parameters()
      if steps_this_epoch:
        last_loss = float(self.engine.losses(1)[0])
      self.last_epoch_losses = self.engine.losses(steps_this_epoch).numpy() if steps_this_epoch else np.zeros(0)
      postfix = {'loss': last_loss}
      if eval_freq > 0 and epoch % eval_freq == 0 and val_dataloader is not None:
        if self._ip is None:   # (item-parallel engines hold shards; the validation loss is skipped there)
          val_loss = self._validate(val_dataloader)
          postfix['val_loss'] = val_loss
        if metrics is not None and eval_num_recommendations is not None:
          results = self._evaluate(val_dataloader.dataset,
                                   num_recommendations=eval_num_recommendations,
                                   metrics=metrics, batch_size=eval_batch_size,
                                   num_users=eval_num_users)
          for metric in results:
            postfix[str(metric)] = np.mean(results[metric])

      progress_bar.set_postfix(postfix)
      progress_bar.close()

      if model_checkpoint_prefix and \
          ((checkpoint_freq > 0 and epoch % checkpoint_freq == 0) or epoch == num_epochs):
        self.save_state(model_checkpoint_prefix)   # every rank calls it; rank 0 writes the file

  def _validate(self, val_dataloader):
    """Average loss over the validation batches (reference model.py:439-452); forward + loss kernels only."""
    self.model.eval()
    total_loss = 0.0
    num_batches = 1
    ds = val_dataloader.dataset
    csr, tcsr = ds.device_csr(), ds.device_target_csr()
    itr = 0
    for index in val_dataloader.pools():
      pool = collate_pool(csr, index, val_dataloader.negative_sampling, table_rows=self.num_items)
      tpool = collate_pool(tcsr, index, val_dataloader.negative_sampling, table_rows=self.num_items) \
        if tcsr is not None else None
      for off in range(0, pool.num_rows, val_dataloader.batch_size):
        rows = min(val_dataloader.batch_size, pool.num_rows - off)
        total_loss += self.engine.eval_loss(pool, off, rows, target_pool=tpool)
        itr += 1
        num_batches = itr
    return total_loss / num_batches

  def predict(self, users_interactions, return_input=False):
    """
    Predicts the user interactions with all items (reference model.py:487-511).

    Returns:
      if ``return_input`` is ``True`` a tuple of the predictions and the dense input batch, otherwise the
      predictions.
    """
    if self.model is None:
      raise Exception('Model not initialized.')
    self.__require_cuda()
    self.flush_parameters()
    self.model.eval()
    batch_collator = BatchCollator(batch_size=len(users_interactions.users), negative_sampling=False)
    batch = batch_collator.collate(users_interactions)[0]
    input_dense = torch.zeros(tuple(batch.size), dtype=torch.float32, device=batch.values.device)
    idx = batch.indices
    if idx.numel():
      input_dense[idx[0], idx[1]] = batch.values
    output = self.model(input_dense, input_users=batch.users)
    return (output, input_dense) if return_input else output

  def _evaluate(self, eval_dataset, num_recommendations, metrics, batch_size=1, num_users=None):
    if self.model is None:
      raise Exception('Model not initialized')
    from .metrics import RecommenderEvaluator
    from .recommender import InferenceRecommender
    self.model.eval()
    recommender = InferenceRecommender(self, num_recommendations)
    evaluator = RecommenderEvaluator(recommender, metrics)
    return evaluator.evaluate(eval_dataset, batch_size=batch_size, num_users=num_users)

  def recommend(self, users_interactions, num_recommendations):
    """
    Generate list of recommendations for each user in ``users_interactions`` (reference model.py:525-544).

    Returns:
      list: list of recommended items for each user in users_interactions.
    """
    if self.model is None:
      raise Exception('Model not initialized.')
    self.__require_cuda()
    if self.engine is not None:
      self.engine.join()
    self.flush_parameters()
    self.model.eval()
    # the pool is collated on the GPU (no negative sampling: columns are raw item ids) and goes through the encoder as
    # CSR; the dense [B, I] input and the boolean mask pass of the reference (model.py:502-510, 541) never exist
    pool, _ = pool_of(users_interactions, False)
    if self.num_items is not None and pool.num_items > self.num_items:
      raise ValueError('recoder_b200: the interactions matrix has %d columns but the model represents only %d items'
                       % (pool.num_items, self.num_items))
    logits = self.model.forward_pool(pool)
    rows, n = logits.shape
    ld = logits.stride(0)
    if pool.nnz:
      _native.call('rcd_mask_seen', _native.ptr(pool.row_ptr), _native.ptr(pool.raw_items), 0, rows,
                   _native.ptr(logits), ld)
    k = int(num_recommendations)
    top_val = torch.empty(rows, k, dtype=torch.float32, device=logits.device)
    top_ind = torch.empty(rows, k, dtype=torch.int64, device=logits.device)
    _native.call('rcd_topk_rows', _native.ptr(logits), ld, rows, n, k, _native.ptr(top_val), _native.ptr(top_ind))
    return top_ind.tolist()

  def evaluate(self, eval_dataset, num_recommendations, metrics, batch_size=1, num_users=None):
    """Evaluates the current model given an evaluation dataset (reference model.py:546-559)."""
    results = self._evaluate(eval_dataset, num_recommendations, metrics,
                             batch_size=batch_size, num_users=num_users)
    for metric in results:
      log.info('{}: {}'.format(metric, np.mean(results[metric])))
    return results

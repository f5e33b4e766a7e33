"""Step engine: one training step of the hot path as a fixed sequence of C-ABI kernel launches.

Replaces the per-batch body of `Recoder._train` (recoder/model.py:383-404) and `Recoder.__compute_loss`
(recoder/model.py:454-485): no dense [B, n] fp32 input/target is ever materialised, no autograd graph is
built, the loss stays on the device.

Launch sequence of an autoencoder step (names are include/recoder_b200.h entry points; DESIGN.md §5 "Streams"):
  aux stream   rcd_slice_csc (CSC views of the slice, needed late) ; rcd_collate of the NEXT pool
  main stream  rcd_gather_rows / rcd_gather_vec (W_d, b_d of the n batch items) -> rcd_ae_encoder_fwd ->
               [inner layers / dropout: rcd_sgemm, rcd_dropout] -> rcd_sddmm (logits at the stored targets, softmax
               reference) -> rcd_decoder_fwd_loss (tcgen05 GEMM, loss / dlogits epilogue) -> rcd_loss_finish ->
               rcd_sparse_dgrad -> rcd_decoder_wgrad (+ bias gradient side product) -> rcd_csc_rows_accumulate ->
               rcd_decoder_dgrad -> rcd_dz_act -> [inner backward] -> rcd_ae_encoder_wgrad -> W_e / b_e update
  update stream  W_d / b_d update as soon as dW_d is complete (rcd_adam_step | rcd_sgd_step | ...), underneath the
               dgrad GEMM and the encoder backward

Multi-GPU: rows of a global batch sharded over ranks with the gradient slab exchanged by
`rcd_adam_step_p2p` (fused reduce-scatter -> Adam -> all-gather over peer memory) or one NCCL all-reduce; or the item
axis sharded (`_ae_step_items`, itempar.py) with four small peer-memory collectives per step.
"""
import math
import os

import torch

from . import _native
from ._native import call, ptr

ADAM_BETAS = (0.9, 0.999)   # torch.optim.Adam defaults used at recoder/model.py:135
ADAM_EPS = 1e-8
SGD_MOMENTUM = 0.9          # recoder/model.py:149


def _round_up(x, m):
  return (x + m - 1) // m * m


def reduce_slab(slab, loss_slot, pg):
  """Data-parallel exchange: ONE sum all-reduce over the contiguous fp32 gradient slab of a step.
  The step loss (float64 [1]) rides in the slab's last two floats as a hi/lo split, so nothing is lost to fp32
  and no second collective is needed.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""
  import torch.distributed as dist
  tail = slab[-2:]
  hi = loss_slot.to(torch.float32)
  lo = (loss_slot - hi.to(torch.float64)).to(torch.float32)
  tail[0:1].copy_(hi)
  tail[1:2].copy_(lo)
  dist.all_reduce(slab, op=dist.ReduceOp.SUM, group=pg)
  loss_slot.copy_(tail[0:1].to(torch.float64) + tail[1:2].to(torch.float64))


def timed_all_reduce(tag, tensor, op, pg):
  """dist.all_reduce with the same optional CUDA-event timing as the C-ABI entry points (bench.py breakdown)."""
  import torch.distributed as dist
  if _native.PROFILE is not None and (_native.PROFILE == 'all' or tag in _native.PROFILE):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    dist.all_reduce(tensor, op=op, group=pg)
    end.record()
    _native.TIMINGS.setdefault(tag, []).append((start, end))
  else:
    dist.all_reduce(tensor, op=op, group=pg)


def shard_rows(pool_rows, global_step_rows, world, rank):
  """Row ranges of one collated pool for data-parallel training: the pool is cut into global steps of
  `global_step_rows` rows (the reference's `batch_size` slices, recoder/data.py:231-249) and every global step
  into `world` equal contiguous blocks.  Yields (row0, rows, global_rows) for `rank`; a ragged tail keeps
  floor(rows/world) rows per rank (the remainder rows are dropped so that every rank runs the same shapes)."""
  for goff in range(0, pool_rows, global_step_rows):
    grows = min(global_step_rows, pool_rows - goff)
    per = grows // world
    if per == 0:
      continue
    yield goff + rank * per, per, per * world


class _Buffers:
  """Grow-only named device buffers (a step never allocates once shapes have been seen)."""

  def __init__(self, device):
    self.device = device
    self._store = {}
    self.retired = []   # outgrown buffers: kernels on the update / auxiliary streams may still read them

  def get(self, name, numel, dtype):
    t = self._store.get(name)
    if t is None or t.numel() < numel or t.dtype != dtype:
      if t is not None:
        # do not hand the old block back to the allocator yet: it would be reused in main-stream order while a kernel
        # of the previous step may still be reading it on another stream (released in TrainEngine.join)
        self.retired.append(t)
      cap = int(numel * 1.2) + 64 if t is not None else int(numel)
      t = torch.empty(max(cap, 1), dtype=dtype, device=self.device)
      self._store[name] = t
    return t[:numel]


class ParamState:
  """Optimizer state of one parameter tensor viewed as a [rows, H] matrix."""

  def __init__(self, name, tensor, weight_decay, sparse):
    self.name = name
    self.p = tensor
    self.weight_decay = weight_decay
    self.sparse = sparse
    self.step = 0
    self.m = None   # Adam exp_avg / SGD momentum buffer
    self.v = None   # Adam exp_avg_sq
    self.shared = None  # p2p.SharedBuffer holding `p` when the table lives in peer-mapped memory
    self.lazy = False   # deferred dense Adam (rcd_adam_lazy_*): `last[r]` = step up to which row r is current
    self.last = None

  def view2d(self):
    t = self.p
    if t.dim() == 1:
      return t.numel(), 1
    return t.shape[0], t.shape[1]


class Optimizer:
  """Fused replacements of the torch.optim objects Recoder.__init_optimizer builds (recoder/model.py:101-164):
  one group per parameter, weight decay 0 for parameters with 'bias' in their name (model.py:123-124), a
  separate row-sparse Adam for `sparse=True` embedding tables (model.py:110-115,137-138)."""

  def __init__(self, named_params, optimizer_type, lr, weight_decay, sparse_names=()):
    if optimizer_type not in ('adam', 'sgd', 'adagrad', 'rmsprop'):
      raise Exception('Unknown optimizer kind')                       # model.py:156
    self.type = optimizer_type
    self.lr = lr            # dense optimizer lr (MultiStepLR acts on it, model.py:327-332)
    self.base_lr = lr       # MultiStepLR's `initial_lr`: the schedule's base, restored from a checkpoint on resume
    self.resumed = False    # True once a state dict has been loaded: its lr / initial_lr win over train(lr=...)
    self.sparse_lr = lr     # SparseAdam lr never decays (reference quirk)
    self.states = {}
    for name, p in named_params:
      wd = 0 if 'bias' in name else weight_decay
      sp = name in sparse_names
      if sp and optimizer_type != 'adam':  # sgd / adagrad / rmsprop
        raise ValueError('Sparse gradients optimization not supported with %s' % optimizer_type)  # model.py:142-152
      self.states[name] = ParamState(name, p, wd, sp)

  # --- deferred dense Adam (include/recoder_b200.h, rcd_adam_lazy_*) ---------------------------------------------------
  SCAL_CHUNK = 4096       # per-step scalars are computed on the host this many steps ahead
  FLUSH_EVERY = 16384     # every table is brought up to date at least this often (bounds the scalar table)

  def enable_lazy(self, names, allow_shared=False):
    """Defers the dense-Adam update of rows outside the batch for the given embedding tables (bit-identical results;
    HBM traffic per step proportional to the batch's rows instead of the whole table)."""
    if self.type != 'adam':
      return
    for n in names:
      st = self.states[n]
      if st.sparse or (st.shared is not None and not allow_shared) or st.p.dim() != 2:
        continue
      self._ensure(st)
      st.lazy = True
      st.last = torch.full((st.p.shape[0],), st.step, dtype=torch.int32, device=st.p.device)
    if any(s.lazy for s in self.states.values()) and getattr(self, '_scal', None) is None:
      dev = next(s.p.device for s in self.states.values() if s.lazy)
      self._scal_cap = self.FLUSH_EVERY + 2 * self.SCAL_CHUNK
      self._scal = torch.zeros(self._scal_cap, 2, dtype=torch.float32, device=dev)
      self._scal_pin = torch.zeros(self.SCAL_CHUNK, 2, dtype=torch.float32).pin_memory()
      self._scal_base = 0        # step number of table entry 0
      self._scal_upto = 0        # entries for steps <= _scal_upto are final (steps already taken)
      self._scal_filled = 0      # entries for steps <= _scal_filled hold values for `_scal_lr`
      self._scal_lr = None
      self._scal_event = None

  def lazy_names(self):
    return [n for n, s in self.states.items() if s.lazy]

  def _lazy_T(self):
    """Steps taken so far by the lazy tables (they are stepped together)."""
    steps = {s.step for s in self.states.values() if s.lazy}
    assert len(steps) <= 1, 'deferred Adam tables must be stepped together'
    return steps.pop() if steps else 0

  def _ensure_scal(self, t):
    """Makes the scalar-table entry of step `t` (the step about to be taken) valid for the current learning rate."""
    if t - self._scal_base >= self._scal_cap - 1:
      self.flush()     # rebase: everything current, history before `t` no longer needed
      self._scal_base = t - 1
      self._scal_filled = 0
    if self._scal_lr != self.lr or t > self._scal_filled:
      # (re)compute the entries of steps t .. t+CHUNK-1 with the current lr; entries of earlier steps are history
      count = min(self.SCAL_CHUNK, self._scal_cap - (t - self._scal_base))
      if self._scal_event is not None:
        self._scal_event.synchronize()      # the previous upload has left the pinned buffer
      lib = _native.load()
      _native.check(lib.rcd_adam_scalars(float(self.lr), ADAM_BETAS[0], ADAM_BETAS[1], int(t), int(count),
                                         self._scal_pin.data_ptr()), 'rcd_adam_scalars')
      lo = t - self._scal_base
      self._scal[lo:lo + count].copy_(self._scal_pin[:count], non_blocking=True)
      self._scal_event = torch.cuda.Event()
      self._scal_event.record()
      self._scal_lr = self.lr
      self._scal_filled = t + count - 1

  def begin_step(self):
    """Called on the main stream before a training step forks work to other streams: publishes the per-step scalars of
    the step about to be taken (deferred tables replay them later)."""
    if getattr(self, '_scal', None) is not None:
      self._ensure_scal(self._lazy_T() + 1)

  def catch_up(self, name, ids, n):
    """Brings rows `ids[0:n]` (int64, device) of a lazy table up to date before they are read."""
    st = self.states[name]
    if not st.lazy or st.step == 0:
      return
    _, H = st.view2d()
    call('rcd_adam_lazy_catchup', ptr(st.p), ptr(st.m), ptr(st.v), H, ptr(ids), int(n), ptr(st.last), st.step,
         ptr(self._scal), self._scal_base, self._scal_cap, ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS,
         float(st.weight_decay), 1, None, None)

  def flush(self):
    """Brings every row of every lazy table up to date (before evaluation, checkpoints, or anything else that reads
    whole tables).  Ordered after the updates still running on the engine's other streams (`_join`, set by the
    engine): a flush that overtakes the last row update would replay rows whose step is still in flight."""
    if not any(st.lazy for st in self.states.values()):
      return
    join = getattr(self, '_join', None)
    if join is not None:
      join()
    for st in self.states.values():
      if st.lazy and st.step > 0:
        rows, H = st.view2d()
        call('rcd_adam_lazy_catchup', ptr(st.p), ptr(st.m), ptr(st.v), H, None, int(rows), ptr(st.last), st.step,
             ptr(self._scal), self._scal_base, self._scal_cap, ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS,
             float(st.weight_decay), 1, None, None)

  def _ensure(self, st):
    if st.m is None:
      st.m = torch.zeros_like(st.p)     # adam exp_avg | sgd momentum buffer | adagrad sum | rmsprop square_avg
      if self.type in ('adam', 'rmsprop'):
        st.v = torch.zeros_like(st.p)   # adam exp_avg_sq | rmsprop momentum buffer

  def step_param(self, name, grad, ldg, pos=None, ids=None, n_ids=0):
    """grad: compact rows [*, ldg] (see rcd_adam_step); pos: int32 map row->grad row or None for dense grads;
    ids: int64 row ids for the row-sparse Adam."""
    st = self.states[name]
    self._ensure(st)
    st.step += 1
    rows, H = st.view2d()
    if self.type == 'adam':
      if st.sparse:
        call('rcd_sparse_adam_step', ptr(st.p), ptr(st.m), ptr(st.v), H, ptr(grad), ldg, ptr(ids), int(n_ids),
             float(self.sparse_lr), ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, st.step)
      elif st.lazy:
        # rows of the batch only (they were caught up before the forward); the rest is replayed when next touched
        assert self._scal_filled >= st.step and self._scal_lr == self.lr, 'Optimizer.begin_step() was not called'
        call('rcd_adam_lazy_update', ptr(st.p), ptr(st.m), ptr(st.v), H, ptr(ids), int(n_ids), ptr(grad), ldg,
             ptr(st.last), float(self.lr), ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, float(st.weight_decay), st.step)
      else:
        call('rcd_adam_step', ptr(st.p), ptr(st.m), ptr(st.v), rows, H, ptr(grad), ldg, ptr(pos), float(self.lr),
             ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS, float(st.weight_decay), st.step)
    elif self.type == 'sgd':
      call('rcd_sgd_step', ptr(st.p), ptr(st.m), rows, H, ptr(grad), ldg, ptr(pos), float(self.lr), SGD_MOMENTUM,
           float(st.weight_decay))
    elif self.type == 'adagrad':   # torch.optim.Adagrad defaults (model.py:144): lr_decay 0, eps 1e-10
      call('rcd_adagrad_step', ptr(st.p), ptr(st.m), rows, H, ptr(grad), ldg, ptr(pos), float(self.lr), 1e-10,
           float(st.weight_decay))
    else:                          # torch.optim.RMSprop(momentum=0.9) (model.py:154): alpha 0.99, eps 1e-8
      call('rcd_rmsprop_step', ptr(st.p), ptr(st.m), ptr(st.v), rows, H, ptr(grad), ldg, ptr(pos), float(self.lr),
           0.99, 1e-8, SGD_MOMENTUM, float(st.weight_decay))

  def step_param_p2p(self, name, ctx, grads_table, ldg, pos, grad_block_rows=0, grads_mc=None):
    """Fused reduce-scatter -> Adam -> all-gather over peer memory (`rcd_adam_step_p2p`): this rank updates its row
    shard of the table from the sum of all ranks' compact gradients and stores the new rows into every replica.
    m / v are full-size tensors of which only the owned rows are live (`gather_shards` completes them)."""
    st = self.states[name]
    assert self.type == 'adam' and not st.sparse and st.shared is not None
    self._ensure(st)
    st.step += 1
    rows, H = st.view2d()
    lo, hi = ctx.owned_rows(rows)
    call('rcd_adam_step_p2p', st.shared.ptr_table(), ptr(st.m), ptr(st.v), lo, hi, H, grads_table, ldg, ptr(pos),
         int(grad_block_rows), ctx.rank, ctx.world, float(self.lr), ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS,
         float(st.weight_decay), st.step, grads_mc, st.shared.mc())

  def gather_shards(self, ctx):
    """Completes the row-sharded Adam state (m, v) of peer-memory tables on every rank (before a checkpoint)."""
    import torch.distributed as dist
    for st in self.states.values():
      if st.shared is None or st.m is None or st.lazy:   # (deferred tables keep complete state on every replica)
        continue
      rows, _ = st.view2d()
      per = (rows + ctx.world - 1) // ctx.world
      for q in range(ctx.world):
        lo, hi = min(q * per, rows), min((q + 1) * per, rows)
        if hi > lo:
          dist.broadcast(st.m[lo:hi], src=dist.get_global_rank(ctx.pg, q), group=ctx.pg)
          dist.broadcast(st.v[lo:hi], src=dist.get_global_rank(ctx.pg, q), group=ctx.pg)

  # --- torch.optim-compatible state_dict (param index = position in named_parameters(), model.py:208-215) ----
  def state_dict(self, dense=True):
    self.flush()
    names = [n for n, s in self.states.items() if s.sparse != dense]
    state, groups = {}, []
    for i, n in enumerate(names):
      s = self.states[n]
      if s.m is not None:
        if self.type == 'adam':
          state[i] = {'step': torch.tensor(float(s.step)), 'exp_avg': s.m.detach().cpu().clone(),
                      'exp_avg_sq': s.v.detach().cpu().clone()}
        elif self.type == 'adagrad':
          state[i] = {'step': torch.tensor(float(s.step)), 'sum': s.m.detach().cpu().clone()}
        elif self.type == 'rmsprop':
          state[i] = {'step': s.step, 'square_avg': s.m.detach().cpu().clone(),
                      'momentum_buffer': s.v.detach().cpu().clone()}
        else:
          state[i] = {'momentum_buffer': s.m.detach().cpu().clone()}
      g = {'params': [i], 'lr': self.lr if dense else self.sparse_lr, 'weight_decay': s.weight_decay}
      if dense:
        g['initial_lr'] = self.base_lr    # what torch's MultiStepLR adds to every group it schedules
      if self.type == 'adam':
        g.update({'betas': ADAM_BETAS, 'eps': ADAM_EPS})
      elif self.type == 'adagrad':
        g.update({'lr_decay': 0, 'eps': 1e-10, 'initial_accumulator_value': 0})
      elif self.type == 'rmsprop':
        g.update({'momentum': SGD_MOMENTUM, 'alpha': 0.99, 'eps': 1e-8, 'centered': False})
      else:
        g.update({'momentum': SGD_MOMENTUM, 'dampening': 0, 'nesterov': False})
      groups.append(g)
    return {'state': state, 'param_groups': groups}

  def load_state_dict(self, sd, dense=True):
    names = [n for n, s in self.states.items() if s.sparse != dense]
    for i, n in enumerate(names):
      s = self.states[n]
      entry = sd['state'].get(i)
      if entry is None:
        continue
      dev = s.p.device
      if self.type == 'adam':
        s.m = entry['exp_avg'].to(dev, torch.float32).clone()
        s.v = entry['exp_avg_sq'].to(dev, torch.float32).clone()
        s.step = int(float(entry['step']))
      elif self.type == 'adagrad':
        s.m = entry['sum'].to(dev, torch.float32).clone()
        s.step = int(float(entry['step']))
      elif self.type == 'rmsprop':
        s.m = entry['square_avg'].to(dev, torch.float32).clone()
        s.v = entry['momentum_buffer'].to(dev, torch.float32).clone()
        s.step = int(float(entry['step']))
      else:
        s.m = entry['momentum_buffer'].to(dev, torch.float32).clone()
    if sd.get('param_groups'):
      lr = sd['param_groups'][0].get('lr')
      if lr is not None:
        if dense:
          # torch's optimizer.load_state_dict restores the groups' lr (and MultiStepLR's initial_lr): on resume the
          # reference trains on with the CHECKPOINT's rates, whatever lr is passed to train() (model.py:158-164, 327-332)
          self.lr = lr
          self.base_lr = sd['param_groups'][0].get('initial_lr', lr)
          self.resumed = True
        else:
          self.sparse_lr = lr


class NativeStep:
  """Host side of the native step executor (`rcd_step_run`, include/recoder_b200.h "K12"): the launch sequence of
  `TrainEngine._ae_step` / `_mf_step` / `_ae_step_items` issued from C++ in ONE call — same entry points, same order,
  same streams, bit-identical results — instead of ~40 ctypes calls (about 1 ms of interpreter time per step, which
  bounds the small configurations and lets one slow host stall every rank at the item-parallel barriers).  Owns one
  grow-only workspace carved by capacities, so nothing is allocated once the shapes have been seen."""

  GROW = 1.25

  def __init__(self, engine):
    import ctypes
    self.ctypes = ctypes
    self.eng = engine
    self.lib = engine.lib
    ctx = ctypes.c_void_p()
    _native.check(self.lib.rcd_step_create(ctypes.byref(ctx)), 'rcd_step_create')
    self.ctx = ctx
    self.args = _native.RcdStepArgs()
    self.ws = None
    self.caps = {'rows': 0, 'n': 0, 'n_in': 0, 'nnz': 0, 'tnnz': 0}
    self._static_done = False
    self._prof_state = None
    self._names = ctypes.create_string_buffer(4096)
    self._ms = (ctypes.c_float * 64)()
    self._cnt = (ctypes.c_int * 64)()

  def close(self):
    if self.ctx is not None:
      self.lib.rcd_step_destroy(self.ctx)
      self.ctx = None

  def __del__(self):  # pragma: no cover
    try:
      self.close()
    except Exception:
      pass

  @staticmethod
  def _fill_param(dst, st, t):
    rows, cols = st.view2d()
    dst.p = st.p.data_ptr()
    dst.s1 = st.m.data_ptr() if st.m is not None else None
    dst.s2 = st.v.data_ptr() if st.v is not None else None
    dst.rows, dst.cols = rows, cols
    dst.weight_decay = float(st.weight_decay)
    dst.t = t
    dst.last = st.last.data_ptr() if st.lazy else None

  @staticmethod
  def _fill_pool(dst, pb, row0, rows):
    dst.row_ptr = pb.row_ptr.data_ptr()
    dst.raw_items = pb.raw_items.data_ptr()
    dst.cols = pb.cols.data_ptr()
    dst.vals = pb.vals.data_ptr()
    dst.row_inv_norm = pb.row_inv_norm.data_ptr()
    dst.row_sum = pb.row_sum.data_ptr()
    dst.pos = pb.pos.data_ptr()
    dst.items = pb.items_buf.data_ptr() if pb.negative_sampling else None
    dst.users = pb.users.data_ptr()
    dst.nnz_slice = int(pb.row_ptr_host[row0 + rows] - pb.row_ptr_host[row0])
    dst.n = pb.n

  def _static(self):
    e, a = self.eng, self.args
    a.abi = _native.STEP_ABI
    a.kind = _native.MODEL_IDS[e.kind]
    a.act = e.act
    a.loss = e.loss_id
    a.optimizer = _native.OPT_IDS[e.opt.type]
    a.confidence = e.confidence
    a.bad_flag = e.bad_flag.data_ptr()
    a.redo_flag = e.redo_flag.data_ptr()
    if e.kind == 'ae':
      a.H = e.params['en_w'][1].shape[1]
    else:
      a.H = e.params['item_w'][1].shape[1]
      U = e.params['user_w'][1]
      self.user_pos = torch.full((U.shape[0],), -1, dtype=torch.int32, device=U.device)
      a.user_pos = self.user_pos.data_ptr()
    if e.ip is not None:
      ctx = e.ip.p2p
      a.ip.enabled, a.ip.rank, a.ip.world = 1, ctx.rank, ctx.world
      a.ip.flags_host = self.ctypes.cast(ctx._flag_table, self.ctypes.c_void_p)
      a.ip.seq_host = self.ctypes.cast(self.ctypes.pointer(ctx._seq), self.ctypes.c_void_p)
      a.ip.barrier_timeout_s = float(ctx.barrier_timeout_s)
    self._static_done = True

  def _names_of(self):
    e = self.eng
    if e.kind == 'ae':
      return e.params['en_w'][0], e.params['en_b'][0], e.params['de_w'][0], e.params['de_b'][0]
    return e.params['user_w'][0], None, e.params['item_w'][0], e.params['bias'][0]

  def _sync_profile(self):
    """Mirrors `_native.PROFILE` (bench.py) into the executor's own CUDA-event timing."""
    want = _native.PROFILE
    key = None if want is None else ('all' if want == 'all' else tuple(sorted(want)))
    if key == self._prof_state:
      return
    self._prof_state = key
    if key is None:
      self.lib.rcd_step_profile(self.ctx, 0, None)
    elif key == 'all' or len(key) != 1:
      self.lib.rcd_step_profile(self.ctx, 1, None)
    else:
      self.lib.rcd_step_profile(self.ctx, 2, key[0].encode())

  def read_profile(self):
    """{entry point: (total ms, launches)} recorded since the last read (synchronises on the recorded events)."""
    k = self.lib.rcd_step_profile_read(self.ctx, self._names, 4096, self._ms, self._cnt, 64)
    if k < 0:
      _native.check(k, 'rcd_step_profile_read')
    names = self._names.value.decode().split('\n') if k else []
    return {names[i]: (float(self._ms[i]), int(self._cnt[i])) for i in range(k)}

  def run(self, pool, tpool, row0, rows, inv_b, loss_slot, train):
    e, a = self.eng, self.args
    if not self._static_done:
      self._static()
    self._sync_profile()
    same = tpool is pool
    n_in_name, b_in_name, n_out_name, b_out_name = self._names_of()
    opt = e.opt
    states = [opt.states[n_in_name], opt.states[b_in_name] if b_in_name else None, opt.states[n_out_name],
              opt.states[b_out_name]]
    for st in states:
      if st is not None:
        opt._ensure(st)
    for dst, st in zip((a.table_in, a.bias_in, a.table_out, a.bias_out), states):
      if st is not None:
        self._fill_param(dst, st, st.step + 1)
    self._fill_pool(a.pool_in, pool, row0, rows)
    self._fill_pool(a.pool_tgt, tpool, row0, rows)
    a.same_pool = int(same)
    a.row0, a.rows = int(row0), int(rows)
    a.inv_b = inv_b
    a.lr = float(opt.lr)
    a.train = int(train)
    a.overlap = int(e.overlap)
    a.loss_acc = loss_slot.data_ptr()
    if getattr(opt, '_scal', None) is not None:
      a.scal, a.scal_base, a.scal_len = opt._scal.data_ptr(), opt._scal_base, opt._scal_cap
    # workspace capacities: grow-only with head-room; the layout is a function of the capacities alone
    c = self.caps
    need = {'rows': rows, 'n': tpool.n, 'n_in': pool.n, 'nnz': a.pool_in.nnz_slice, 'tnnz': a.pool_tgt.nnz_slice}
    grow = self.ws is None or any(need[k] > c[k] for k in c)
    if grow:
      for k in c:
        if need[k] > c[k]:
          c[k] = int(need[k]) if k == 'rows' else int(need[k] * self.GROW) + 16
      # a batch never holds more items than the table has rows
      if tpool.negative_sampling:
        c['n'] = max(min(c['n'], int(a.table_out.rows)), tpool.n)
      if e.kind == 'ae' and pool.negative_sampling:
        c['n_in'] = max(min(c['n_in'], int(a.table_in.rows)), pool.n)
    a.cap_rows, a.cap_n, a.cap_n_in, a.cap_nnz, a.cap_tnnz = c['rows'], c['n'], c['n_in'], c['nnz'], c['tnnz']
    if grow:
      nbytes = int(self.lib.rcd_step_workspace_bytes(self.ctypes.byref(a)))
      if nbytes == 0:
        raise RuntimeError('recoder_b200: rcd_step_workspace_bytes rejected the step description')
      if self.ws is None or self.ws.numel() < nbytes:
        torch.cuda.synchronize()     # kernels of earlier steps may still use the old workspace on any stream
        self.ws = None
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=e.device)
    a.ws = self.ws.data_ptr()
    a.ws_bytes = self.ws.numel()
    # streams
    if e.overlap:
      if e._side is None:
        e._side = torch.cuda.Stream(device=e.device)
      if e._aux is None:
        e._aux = torch.cuda.Stream(device=e.device)
      a.stream_side = e._side.cuda_stream
      a.stream_aux = e._aux.cuda_stream
      e._keep_for_side(pool, None if same else tpool)
    a.stream_main = _native.stream_ptr()
    # deferred Adam, opt-in (RCD_LAZY_PREFETCH=1): catch the next pool's rows up ahead of time on the side stream (last
    # slice of this pool only; see rcd_step_args).  Measured a LOSS (profiles/README.md r02m: C3 2.15 -> 2.28 ms, C5/512
    # 2.67 -> 2.84 ms): the SIMT catch-up CTAs share the SMs' issue slots with the single MMA / TMA threads of the
    # tensor-core kernels they overlap, which slows those down by more than the catch-up saves at the start of a step.
    a.next_items_in = a.next_items_out = None
    nxt = getattr(pool, 'next_hint', None)
    if (train and e.overlap and nxt is not None and row0 == getattr(pool, 'last_slice_row0', -1)
        and os.environ.get('RCD_LAZY_PREFETCH', '0') == '1'):
      npool, ntpool = nxt
      ntpool = ntpool or npool
      lazy_in = e.kind == 'ae' and opt.states[n_in_name].lazy
      lazy_out = opt.states[n_out_name].lazy
      if (lazy_in or lazy_out) and npool.negative_sampling and npool._pending is not None:
        e._side.wait_event(npool._pending[1])         # the next pool's collate (trainer's auxiliary stream)
        if ntpool is not npool and ntpool._pending is not None:
          e._side.wait_event(ntpool._pending[1])
        if lazy_out:
          a.next_items_out = ntpool.items_buf.data_ptr()
          a.next_n_out = ntpool.counts.data_ptr()
          a.next_cap_out = ntpool.items_buf.numel()
        if lazy_in:
          a.next_items_in = npool.items_buf.data_ptr()
          a.next_n_in = npool.counts.data_ptr()
          a.next_cap_in = npool.items_buf.numel()
    if e.ip is not None:
      xb = e._ip_buffers(rows, a.H)
      sh = xb['shared']
      o_z, o_dz, o_ref, o_sum = xb['off']
      a.ip.shared_local = sh.local_ptr
      self._ip_table = sh.ptr_table()
      a.ip.shared_host = self.ctypes.cast(self._ip_table, self.ctypes.c_void_p)
      a.ip.shared_mc = sh.mc()
      a.ip.off_z, a.ip.off_dz, a.ip.off_ref, a.ip.off_sum = o_z // 4, o_dz // 4, o_ref // 4, o_sum // 4
    status = self.lib.rcd_step_run(self.ctx, self.ctypes.byref(a))
    if status != 0:
      _native.check(status, 'rcd_step_run')
    if train:
      for st in states:
        if st is not None:
          st.step += 1
      H = a.H
      f32 = self.ws.view(torch.float32)

      def view(off, numel):
        return f32[off // 4:off // 4 + numel]
      if e.kind == 'ae':
        e.last = {'n': tpool.n, 'n_in': pool.n, 'dWe': view(a.out_dW_in, pool.n * H).view(pool.n, H),
                  'dWd': view(a.out_dW_out, tpool.n * H).view(tpool.n, H), 'dbd': view(a.out_db_out, tpool.n),
                  'dbe': view(a.out_db_in, H), 'inner': f32[0:0], 'inner_layout': e._inner_layout()}
      else:
        e.last = {'n': tpool.n, 'dV': view(a.out_dW_out, tpool.n * H).view(tpool.n, H),
                  'dbias': view(a.out_db_out, tpool.n), 'dU': view(a.out_dW_in, rows * H).view(rows, H)}

  def join(self):
    _native.check(self.lib.rcd_step_join(self.ctx, _native.stream_ptr()), 'rcd_step_join')


class TrainEngine:
  """Executes training steps for a single-hidden-layer DynamicAutoencoder ('ae') or a MatrixFactorization ('mf')."""

  LOSS_RING = 4096

  def __init__(self, kind, params, loss, confidence, activation, optimizer: Optimizer, gemm_engine=None,
               process_group=None, tied=False, p2p=None, item_parallel=None, loss_module=None, lazy_adam=False):
    _native.require_cuda()
    # deferred dense Adam for the embedding tables: False (default for directly constructed engines: parameters are
    # always current), True, or 'auto' = per table when the batches touch a small enough share of its rows
    self.lazy_adam = lazy_adam
    self._lazy_decided = False
    self.kind = kind
    self.params = params          # dict of role -> (name, tensor)
    self.loss_module = loss_module   # kind 'custom': any nn.Module with sum reduction (recoder/model.py:88-89)
    self.loss_id = _native.LOSS_IDS[loss] if loss != 'custom' else -1
    if loss == 'custom' and loss_module is None:
      raise ValueError("loss kind 'custom' needs the loss module")
    self.confidence = float(confidence)
    self.act = _native.ACT_IDS[activation]
    self.opt = optimizer
    self.gemm = _native.GEMM_TCGEN05 if gemm_engine is None else gemm_engine
    self.pg = process_group
    self.tied = tied
    self.p2p = p2p                # p2p.P2PContext: exchange through peer memory instead of an NCCL all-reduce
    self.ip = item_parallel       # itempar.ItemParallel: every rank sees all rows, the item axis is sharded
    self._ip_shared = None
    self._slab_shared = None
    self.overlap = os.environ.get('RCD_OVERLAP', '1') != '0'
    self._side = None
    self._aux = None
    self._ready = {}
    dev = (params['en_w'] if kind == 'ae' else params['item_w'])[1].device
    self.device = dev
    # generalised model: inner dense layers, input noise, bottleneck / user-embedding dropout
    self.enc_layers = params.get('enc_layers', [])   # [{'w': (name, [out,in]), 'b': (name, [out])}, ...]
    self.dec_layers = params.get('dec_layers', [])   # 'w' is None for tied layers (weight = enc layer^T)
    self.noise_prob = float(params.get('noise_prob', 0.0))
    self.dropout_prob = float(params.get('dropout_prob', 0.0))
    # Philox key of the dropout / noise masks: drawn from torch's global generator when the engine is built (like
    # RandomSampler seeds itself, data.py), so successive train() calls and different models get fresh mask streams
    # while torch.manual_seed() still makes a run reproducible; data-parallel ranks share rank 0's key (the masks are
    # indexed by GLOBAL element, so all ranks must agree)
    seed = int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFFFFFFFFFF
    if process_group is not None:
      import torch.distributed as dist
      if dist.get_world_size(process_group) > 1:
        t = torch.tensor([seed], dtype=torch.int64, device=dev if dist.get_backend(process_group) == 'nccl' else 'cpu')
        dist.broadcast(t, src=dist.get_global_rank(process_group, 0), group=process_group)
        seed = int(t.item())
    self.rng_seed = seed
    self.debug_noise_keep = None     # tests: explicit uint8 keep masks instead of Philox
    self.debug_dropout_keep = None
    self.buf = _Buffers(dev)
    self.loss_ring = torch.zeros(self.LOSS_RING, dtype=torch.float64, device=dev)
    self.steps_done = 0
    self.lib = _native.load()
    self.tile_n = self.lib.rcd_decoder_tile_n()
    self.last = {}                # views of the last step's compact gradients (tests / telemetry)
    self.bad_flag = torch.zeros(1, dtype=torch.int32, device=dev)   # set by rcd_loss_finish on non-finite rows
    self.redo_flag = torch.zeros(1, dtype=torch.int32, device=dev)  # NLL rows to redo with their true maximum
    self._loss_host = None
    self.opt._join = self.join     # Optimizer.flush orders itself after the update / side streams
    self._native = None            # NativeStep, created on first use
    self.native_enabled = os.environ.get('RCD_NATIVE_STEP', '1') != '0'
    self._used_python_path = False

  LAZY_GAIN = 0.9   # 'auto': defer when the estimated traffic is below this share of the dense update's

  def _decide_lazy(self, pool, tpool, rows):
    """Enables the deferred dense Adam (Optimizer.enable_lazy) per table on the first training step.  Estimate of the
    traffic per parameter: dense 24 B over all T rows + 4 B over the n gradient rows; deferred 28 B over the n batch rows
    + 24 B over the batch rows that were not in the previous batch (about n * (1 - n/T) of them)."""
    self._lazy_decided = True
    if not self.lazy_adam or self.opt.type != 'adam' or self.tied:
      return
    if self.p2p is not None:
      # the fused peer-memory exchange updates row SHARDS densely and pushes them to every replica — except for the MF
      # user table, whose gradient rows belong to the B_global users of the step: deferred, every replica applies the
      # same B_global row updates itself and the table (1M x 256 at C4) never crosses NVLink
      if self.kind == 'mf':
        name = self.params['user_w'][0]
        T = self.opt.states[name].p.shape[0]
        n = rows * self._world()[0]
        if self.lazy_adam is True or (28.0 * n + 24.0 * n) / (24.0 * T + 4.0 * n) < self.LAZY_GAIN:
          self.opt.enable_lazy([name], allow_shared=True)
      return
    # row-parallel runs with the NCCL exchange qualify: every rank applies the same catch-ups and the same row updates
    # from the same all-reduced gradients, so the replicas stay bit-identical
    world = self._world()[0] if self.ip is None else 1
    if self.kind == 'ae':
      cand = [(self.params['en_w'][0], pool.n), (self.params['de_w'][0], tpool.n)]
    else:
      cand = [(self.params['user_w'][0], rows * world), (self.params['item_w'][0], tpool.n)]
    names = []
    for name, n in cand:
      T = self.opt.states[name].p.shape[0]
      est = (28.0 * n + 24.0 * n * (1.0 - n / T)) / (24.0 * T + 4.0 * n)
      if self.lazy_adam is True or est < self.LAZY_GAIN:
        names.append(name)
    self.opt.enable_lazy(names)

  def _native_ok(self, pool, tpool, train):
    """True when the step can go through the native executor (`rcd_step_run`): single-hidden-layer autoencoder or
    matrix factorisation, fused loss, dense optimizer, tcgen05 engine, one GPU or the item-parallel mode with
    peer-memory collectives.  Everything else (inner layers, noise / dropout, tied weights, SparseAdam, custom loss
    modules, row-parallel exchanges, the SIMT validation engine) keeps the Python launch sequence."""
    if not self.native_enabled or self.loss_module is not None or self.gemm != _native.GEMM_TCGEN05:
      return False
    if self.enc_layers or self.dec_layers or self.tied:
      return False
    if train and (self.noise_prob > 0.0 or self.dropout_prob > 0.0):
      return False
    if any(st.sparse for st in self.opt.states.values()):
      return False
    if self.ip is not None:
      return train and self.ip.p2p is not None
    return self.pg is None

  def _native_run(self, pool, tpool, row0, rows, inv_b, loss_slot, train):
    if self._native is None:
      self._native = NativeStep(self)
    if self._used_python_path:     # order the executor after Python-path work still pending on the other streams
      self.join()
      self._used_python_path = False
    self._native.run(pool, tpool, row0, rows, inv_b, loss_slot, train)

  # ------------------------------------------------------------------------------------------------------
  def _loss_slot(self):
    i = self.steps_done % self.LOSS_RING
    slot = self.loss_ring[i:i + 1]
    slot.zero_()
    return slot

  def losses(self, last_k):
    """The last `last_k` step losses as a CPU float64 tensor (one sync)."""
    k = min(last_k, self.steps_done, self.LOSS_RING)
    self.join()
    self.check_finite()
    idx = [(self.steps_done - k + j) % self.LOSS_RING for j in range(k)]
    return self.loss_ring[torch.tensor(idx, device=self.device, dtype=torch.long)].cpu() if k else torch.zeros(0)

  def check_finite(self):
    """Raises if a step produced a non-finite loss / softmax row sum (synchronises)."""
    self.join()
    flag = int(self.bad_flag.item())
    if flag:
      self.bad_flag.zero_()
      if flag & 4:
        raise RuntimeError('recoder_b200: a peer rank did not reach the data-parallel barrier in time (flag %d)' % flag)
      raise FloatingPointError('recoder_b200: non-finite loss in a training step (flag %d: 1 = softmax row sum, '
                               '2 = loss value)' % flag)

  def loss_to_host_deferred(self):
    """Per-step loss read-back without stalling the pipeline: enqueues the D2H copy of the loss of the step just
    launched (8 bytes, pinned ring) and returns the loss of the PREVIOUS step (None on the first call)."""
    from . import data as _data
    if self._loss_host is None:
      self._loss_host = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(4)]
      self._loss_pending = []
    i = (self.steps_done - 1) % self.LOSS_RING
    if self.p2p is not None:
      self.join()   # the global loss is produced on the update stream
    buf = self._loss_host[self.steps_done % len(self._loss_host)]
    buf.copy_(self.loss_ring[i:i + 1], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _data.TRANSFER_BYTES['d2h'] += 8
    self._loss_pending.append((buf, ev))
    if len(self._loss_pending) < 2:
      return None
    pbuf, pev = self._loss_pending.pop(0)
    pev.synchronize()
    return float(pbuf[0])

  def drain_deferred_loss(self):
    out = None
    while getattr(self, '_loss_pending', None):
      pbuf, pev = self._loss_pending.pop(0)
      pev.synchronize()
      out = float(pbuf[0])
    return out

  def _world(self):
    if self.pg is None:
      return 1, 0
    import torch.distributed as dist
    return dist.get_world_size(self.pg), dist.get_rank(self.pg)

  # ------------------------------------------------------------------------------------------------------
  def train_step(self, pool, row0, rows, target_pool=None, global_rows=None):
    """One optimizer step on rows [row0, row0+rows) of `pool` (a data.PoolBatch).  `target_pool` is the collate
    of the dataset's target matrix for the same users (recoder/model.py:464-472) or None when the input is its
    own target (model.py:473-476).  `global_rows` is the number of rows of the whole (all-rank) batch the loss
    is averaged over (model.py:483-484); defaults to `rows`."""
    inv_b = 1.0 / float(global_rows or rows)
    self._check_pool(pool, target_pool)
    if not self._lazy_decided:
      self._decide_lazy(pool, target_pool or pool, rows)
    self.opt.begin_step()
    loss_slot = self._loss_slot()
    if self._native_ok(pool, target_pool or pool, True):
      self._native_run(pool, target_pool or pool, row0, rows, inv_b, loss_slot, True)
      self.steps_done += 1
      return
    self._python_path()
    if self.ip is not None:
      self._ae_step_items(pool, row0, rows, inv_b, loss_slot, train=True)
    elif self.kind == 'ae':
      self._ae_step(pool, target_pool or pool, row0, rows, inv_b, loss_slot, train=True)
    else:
      self._mf_step(pool, target_pool or pool, row0, rows, inv_b, loss_slot, train=True)
    self.steps_done += 1

  def eval_loss(self, pool, row0, rows, target_pool=None):
    """Loss of one batch without touching the parameters (`Recoder._validate`, recoder/model.py:439-452)."""
    self._check_pool(pool, target_pool)
    slot = self.buf.get('eval_loss', 1, torch.float64)
    self.join()
    slot.zero_()
    if self._native_ok(pool, target_pool or pool, False):
      self._native_run(pool, target_pool or pool, row0, rows, 1.0 / rows, slot, False)
      return float(slot.item())
    self._python_path()
    if self.ip is not None:
      self._ae_step_items(pool, row0, rows, 1.0 / rows, slot, train=False)
    elif self.kind == 'ae':
      self._ae_step(pool, target_pool or pool, row0, rows, 1.0 / rows, slot, train=False)
    else:
      self._mf_step(pool, target_pool or pool, row0, rows, 1.0 / rows, slot, train=False)
    return float(slot.item())

  def _check_pool(self, pool, target_pool=None):
    """The kernels index the embedding tables by raw item id (encoder, sparse dgrad, gathers) and read the pool's
    item -> column map `pos` for EVERY table row (optimizers): the pool must have been collated against a matrix no
    wider than the tables and with `table_rows` = the table height (`data.collate_pool_launch`).  The reference
    allows `Recoder(num_items=N)` with N larger than the matrix width (recoder/model.py:241) and fails inside its
    embedding lookup when the matrix is wider."""
    if self.kind == 'ae':
      tables = [(pool, self.params['en_w'][1].shape[0]), (target_pool or pool, self.params['de_w'][1].shape[0])]
    else:
      tables = [(target_pool or pool, self.params['item_w'][1].shape[0])]
      users = self.params['user_w'][1].shape[0]
      if pool.max_user >= users:
        raise ValueError('recoder_b200: user id %d in the batch but the model represents %d users' %
                         (pool.max_user, users))
    for pb, rows in tables:
      if pb.num_items > rows:
        raise ValueError('recoder_b200: the interactions matrix has %d columns but the model represents only %d items'
                         % (pb.num_items, rows))
      if pb.pos.numel() < rows:
        raise ValueError('recoder_b200: the pool was collated for %d items but the model represents %d: collate it '
                         'with table_rows=%d' % (pb.pos.numel(), rows, rows))

  def _aux_stream(self):
    """Context manager: the column-major views of the slice (needed only by the weight gradients, late in the step)
    are built on an auxiliary stream underneath the forward kernels."""
    import contextlib
    if not self.overlap:
      return contextlib.nullcontext()
    if self._aux is None:
      self._aux = torch.cuda.Stream(device=self.device)
    self._aux.wait_stream(torch.cuda.current_stream())
    return torch.cuda.stream(self._aux)

  def _aux_done(self):
    if self._aux is not None:
      ev = torch.cuda.Event()
      ev.record(self._aux)
      self._ready['csc'] = ev

  HEAVY_COLUMN_ROWS = 4096

  def _heavy_scratch(self, n, nnz, H, rows):
    """Workspace of the chunked heavy-column path of the column-major accumulations (rcd_csc_heavy_scratch_bytes).
    A column holds at most one entry per row, so slices of up to HEAVY_COLUMN_ROWS rows keep the single-pass kernel
    (measured at 2048 rows: 115 / 163 us single-pass against 134 / 189 us chunked); taller slices — the item-parallel
    mode processes the whole global batch — take the chunked path (16384 rows: 0.54 / 0.64 ms -> 0.08 / 0.09 ms)."""
    if rows <= self.HEAVY_COLUMN_ROWS:
      return None, 0, int(nnz)
    sbytes = self.lib.rcd_csc_heavy_scratch_bytes(int(n), int(nnz), int(H))
    return self.buf.get('heavy_scratch', sbytes, torch.uint8), sbytes, int(nnz)

  def _slice_csc(self, pool, row0, rows, n, tag):
    nnz = int(pool.row_ptr_host[row0 + rows] - pool.row_ptr_host[row0])
    b = self.buf
    csc_ptr = b.get(tag + 'csc_ptr', n + 1, torch.int32)
    csc_row = b.get(tag + 'csc_row', max(nnz, 1), torch.int32)
    csc_val = b.get(tag + 'csc_val', max(nnz, 1), torch.float32)
    csc_src = b.get(tag + 'csc_src', max(nnz, 1), torch.int32)
    sbytes = self.lib.rcd_slice_csc_scratch_bytes(n, max(nnz, 1))
    scratch = b.get('csc_scratch', sbytes, torch.uint8)
    call('rcd_slice_csc', ptr(pool.row_ptr), ptr(pool.cols), ptr(pool.vals), row0, rows, n, ptr(csc_ptr),
         ptr(csc_row), ptr(csc_val), ptr(csc_src), ptr(scratch), sbytes)
    return csc_ptr, csc_row, csc_val, csc_src

  def _decoder_and_loss(self, Zb, ldh, Zf32, Wg, bias_g, rows, n, H, inv_b, tpool, row0, loss_slot, train):
    """K5 (sparse side) + K4 (fused decoder GEMM / loss epilogue) + row finish.
    Returns (G bf16 [rows, ldn], ldn, corr fp32 [nnz], alpha fp32 [rows] or None, Zs bf16 operand for dW_d)."""
    b = self.buf
    ldn = _round_up(n, 8)
    nnz = max(int(tpool.row_ptr_host[row0 + rows] - tpool.row_ptr_host[row0]), 1)
    if self.loss_module is not None:
      return self._custom_loss(Zb, ldh, Wg, bias_g, rows, n, H, inv_b, tpool, row0, loss_slot, train, ldn, nnz)
    nll = self.loss_id == _native.LOSS_IDS['logloss']
    G = b.get('G', rows * ldn, torch.bfloat16)
    o_nnz = b.get('o_nnz', nnz, torch.float32)
    corr = b.get('corr', nnz, torch.float32)
    row_ref = b.get('row_ref', rows, torch.float32) if nll else None
    call('rcd_sddmm', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias_g), H, ptr(tpool.row_ptr), ptr(tpool.cols),
         ptr(tpool.vals), row0, rows, self.loss_id, self.confidence, inv_b, ptr(o_nnz), ptr(corr), ptr(row_ref))
    stat_cols = self.lib.rcd_decoder_stat_cols(n)
    stat = b.get('stat', rows * stat_cols, torch.float32)
    alpha = b.get('alpha', rows, torch.float32) if nll else None
    Zs = b.get('Zs', rows * ldh, torch.bfloat16) if (nll and train) else None
    nblocks = self.lib.rcd_loss_finish_blocks(rows)
    loss_blocks = b.get('loss_blocks', nblocks, torch.float64)
    row_redo = b.get('row_redo', rows, torch.int32) if nll else None

    def fused(mode, cond):
      call('rcd_decoder_fwd_loss', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias_g), rows, n, H, self.loss_id, inv_b,
           ptr(row_ref), ptr(G), ldn, ptr(stat), stat_cols, mode, ptr(cond))

    def finish(redo_flag, redo_rows, cond):
      call('rcd_loss_finish', ptr(stat), stat_cols, stat_cols, rows, self.loss_id, self.confidence, inv_b,
           ptr(row_ref), ptr(tpool.row_sum), ptr(tpool.row_ptr), ptr(tpool.vals), ptr(o_nnz), row0, ptr(alpha),
           ptr(Zf32), H, ptr(Zs), ldh, None, ptr(self.bad_flag), 0, ptr(loss_blocks), ptr(redo_flag), ptr(redo_rows),
           ptr(cond))

    fused(_native.DEC_MODE_LOSS, None)
    if nll:
      # F.log_softmax is stable for any logits (recoder/losses.py:69).  The fused pass takes the largest TARGET logit as
      # the softmax reference; a row in which some other logit towers over it has its exponentials clamped, is flagged
      # by the row finish and redone with its true maximum — four launches that return at once while the flag is clear
      flag = self.redo_flag
      finish(flag, row_redo, None)
      fused(_native.DEC_MODE_ROWMAX, flag)
      call('rcd_nll_ref_fix', ptr(stat), stat_cols, stat_cols, rows, ptr(row_redo), ptr(row_ref), ptr(flag))
      fused(_native.DEC_MODE_LOSS, flag)
      finish(None, None, flag)
      call('rcd_loss_sum', ptr(loss_blocks), nblocks, ptr(loss_slot), ptr(flag))
    else:
      finish(None, None, None)
      call('rcd_loss_sum', ptr(loss_blocks), nblocks, ptr(loss_slot), None)
    return G, ldn, corr, alpha, (Zs if Zs is not None else Zb)

  def _custom_loss(self, Zb, ldh, Wg, bias_g, rows, n, H, inv_b, tpool, row0, loss_slot, train, ldn, nnz):
    """Generic loss path for user-supplied `nn.Module`s (sum reduction): fp32 logits [rows, n] from the tcgen05 GEMM,
    the dense target scattered from the pool's CSR, loss and dL/dlogits through torch autograd (the module is opaque),
    dL/dlogits rounded to bf16 for the same backward GEMMs as the fused losses.  Materialises three [rows, n] fp32
    matrices, like the reference does (recoder/model.py:457-458, 473-484) — the price of an arbitrary module."""
    b = self.buf
    ldo = ldn
    O = b.get('custom_O', rows * ldo, torch.float32).view(rows, ldo)
    call('rcd_decoder_fwd', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias_g), rows, n, H, None, ptr(O), ldo, None, None,
         _native.GEMM_TCGEN05)
    T = b.get('custom_T', rows * n, torch.float32).view(rows, n)
    T.zero_()
    lo, hi = int(tpool.row_ptr_host[row0]), int(tpool.row_ptr_host[row0 + rows])
    if hi > lo:
      coo = b.get('custom_coo', 2 * (hi - lo), torch.int64).view(2, hi - lo)
      call('rcd_collate_coo', ptr(tpool.row_ptr), ptr(tpool.cols), int(row0), int(rows), ptr(coo))
      T[coo[0], coo[1]] = tpool.vals[lo:hi]
    logits = O[:, :n].detach().requires_grad_(train)
    with torch.enable_grad():
      loss = self.loss_module(logits, T) * inv_b                      # `/ B` of recoder/model.py:483-484
      if train:
        loss.backward()
    loss_slot.add_(loss.detach().to(torch.float64).reshape(1))
    G = b.get('G', rows * ldn, torch.bfloat16)
    corr = b.get('corr', nnz, torch.float32)
    if train:
      call('rcd_f32_to_bf16_rows', ptr(logits.grad.contiguous()), rows, n, ptr(G), ldn)
      corr.zero_()                                                    # the whole gradient is in G: no sparse part
    return G, ldn, corr, None, Zb

  def _sparse_dgrad(self, corr, W_master, tpool, row0, rows, n, H):
    """fp32 sparse part of dZ (reads the MASTER table, so it is issued before that table's optimizer update).
    Returns (partials buffer, splits): slot `splits` holds the sparse part, slots [0, splits) are for the GEMM."""
    splits = self.lib.rcd_decoder_dgrad_splits(rows, n, H)
    partials = self.buf.get('dz_partials', (splits + 1) * rows * H, torch.float32)
    sparse_slot = partials[splits * rows * H:]
    call('rcd_sparse_dgrad', ptr(W_master), H, ptr(tpool.row_ptr), ptr(tpool.raw_items), ptr(corr), row0, rows,
         ptr(sparse_slot), H)
    return partials, splits

  def _dgrad(self, G, ldn, alpha, Wg, ldh, partials, splits, rows, n, H, Zf32, act, dA, db):
    """dZ = alpha * (dense part: tcgen05 split-K GEMM over bf16 G) + sparse part (already in partials) -> dA, db."""
    call('rcd_decoder_dgrad', ptr(G), ldn, ptr(Wg), ldh, rows, n, H, splits, ptr(partials), H, self.gemm)
    call('rcd_dz_act', ptr(partials), splits + 1, splits, ptr(alpha), H, ptr(Zf32), rows, H, act, ptr(dA), ptr(db))

  def _wgrad(self, G, ldn, Zs, ldh, Zf32, csc, corr, alpha, rows, n, H, dW, db):
    """dW rows = dense part (tcgen05 GEMM, bias gradient as a side product) + sparse part (fp32 rank-1 updates at
    the stored targets)."""
    call('rcd_decoder_wgrad', ptr(G), ldn, ptr(Zs), ldh, rows, n, H, ptr(dW), H, ptr(alpha), ptr(db), self.gemm)
    csc_ptr, csc_row, _, csc_src = csc
    scratch, sbytes, nnz = self._heavy_scratch(n, csc_row.numel(), H, rows)
    call('rcd_csc_rows_accumulate', ptr(Zf32), H, ptr(csc_ptr), ptr(csc_row), ptr(csc_src), ptr(corr), n, ptr(dW),
         ptr(db), ptr(scratch), sbytes, nnz)

  # --- update stream: optimizer / exchange kernels overlap the rest of the backward ------------------------------
  def _update_stream(self):
    """Context manager: kernels launched inside run on the side stream, ordered after everything enqueued so far on
    the main stream (RCD_OVERLAP=0: they stay on the main stream)."""
    import contextlib
    if not self.overlap:
      return contextlib.nullcontext()
    if self._side is None:
      self._side = torch.cuda.Stream(device=self.device)
    ev = torch.cuda.Event()
    ev.record()
    self._side.wait_event(ev)
    return torch.cuda.stream(self._side)

  def _keep_for_side(self, *pools):
    """Pool tensors the update stream reads (pos / items) must not be recycled before its kernels have run."""
    if self._side is None:
      return
    for pb in pools:
      if pb is None:
        continue
      for t in (pb.pos, pb.items_buf, pb.users):
        if t is not None and t.is_cuda:
          t.record_stream(self._side)

  def _mark_ready(self, tag):
    ev = torch.cuda.Event()
    ev.record()          # on the current stream (the side stream inside `_update_stream`)
    self._ready[tag] = ev

  def _wait_ready(self, tag):
    ev = self._ready.pop(tag, None)
    if ev is not None:
      torch.cuda.current_stream().wait_event(ev)

  def _python_path(self):
    """Called before a step takes the Python launch sequence: orders it after native-executor work in flight."""
    if self._native is not None:
      self._native.join()
    self._used_python_path = True

  def join(self):
    """Makes the current stream wait for every update still running on the side stream."""
    if self._native is not None:
      self._native.join()
    for tag in list(self._ready):
      self._wait_ready(tag)
    if self.buf.retired:
      # everything enqueued on the other streams so far is now ordered before later main-stream work, which is the
      # order the caching allocator assumes when it recycles a block
      if self._side is not None:
        torch.cuda.current_stream().wait_stream(self._side)
      if self._aux is not None:
        torch.cuda.current_stream().wait_stream(self._aux)
      self.buf.retired.clear()

  def _stash_loss(self, slab, loss_slot):
    """The step loss (float64) rides in the slab's last two floats as a hi/lo split."""
    tail = slab[-2:]
    hi = loss_slot.to(torch.float32)
    tail[0:1].copy_(hi)
    tail[1:2].copy_((loss_slot - hi.to(torch.float64)).to(torch.float32))

  # --- generalised model: inner dense layers, dropout --------------------------------------
  def _inner_layout(self):
    """Offsets of the inner-layer gradients inside the slab: {'size', 'enc': [(o_w, o_b, out, in)], 'dec': [...]}
    (a tied decoding layer has no weight gradient of its own: o_w is the encoder layer's)."""
    lay = getattr(self, '_inner_layout_cache', None)
    if lay is not None:
      return lay
    off = 0
    enc, dec = [], []
    for L in self.enc_layers:
      out_f, in_f = L['w'][1].shape
      enc.append((off, off + out_f * in_f, out_f, in_f))
      off += out_f * in_f + _round_up(out_f, 4)
      off = _round_up(off, 4)
    k = len(self.enc_layers)
    for j, L in enumerate(self.dec_layers):
      if L['w'] is None:
        e_out, e_in = self.enc_layers[k - 1 - j]['w'][1].shape
        out_f, in_f = e_in, e_out
        dec.append((enc[k - 1 - j][0], off, out_f, in_f))
        off += _round_up(out_f, 4)
      else:
        out_f, in_f = L['w'][1].shape
        dec.append((off, off + out_f * in_f, out_f, in_f))
        off += out_f * in_f + _round_up(out_f, 4)
        off = _round_up(off, 4)
    self._inner_layout_cache = {'size': _round_up(off, 4), 'enc': enc, 'dec': dec}
    return self._inner_layout_cache

  def _dropout(self, x, y, p, rng_stream, index_base, keep_mask):
    """y = dropout(x) with the step's Philox key (or an explicit keep mask, tests)."""
    seed = (self.rng_seed + 0x9E3779B97F4A7C15 * (self.steps_done + 1)) & 0xFFFFFFFFFFFFFFFF
    call('rcd_dropout', ptr(x), int(x.numel()), float(p), seed, int(rng_stream), int(index_base), ptr(keep_mask),
         ptr(y))

  def _mid_forward(self, Z0, rows, row0, H, train):
    """Everything between the embedding encoder and the embedding decoder (nn.py:242-249): inner encoding layers,
    bottleneck dropout, inner decoding layers (activation after EVERY layer).  Returns (Y [rows, H], saved)."""
    b = self.buf
    acts = [Z0.view(rows, H)]
    for i, L in enumerate(self.enc_layers):
      W, bias = L['w'][1], L['b'][1]
      out_f, in_f = W.shape
      y = b.get('enc_act%d' % i, rows * out_f, torch.float32).view(rows, out_f)
      call('rcd_sgemm', 0, 1, rows, out_f, in_f, ptr(acts[-1]), in_f, ptr(W), in_f, ptr(y), out_f, ptr(bias), self.act, 0)
      acts.append(y)
    top = acts[-1]
    drop = train and self.dropout_prob > 0.0
    y0 = top
    if drop:
      y0 = b.get('drop_out', top.numel(), torch.float32).view(top.shape)
      self._dropout(top, y0, self.dropout_prob, 2, row0 * top.shape[1], self.debug_dropout_keep)
    dec_acts = [y0]
    k = len(self.enc_layers)
    for j, L in enumerate(self.dec_layers):
      bias = L['b'][1]
      if L['w'] is None:     # tied: weight = enc layer^T, y = x @ W_enc (no transpose of the stored [out,in] needed)
        We_l = self.enc_layers[k - 1 - j]['w'][1]
        in_f, out_f = We_l.shape
        y = b.get('dec_act%d' % j, rows * out_f, torch.float32).view(rows, out_f)
        call('rcd_sgemm', 0, 0, rows, out_f, in_f, ptr(dec_acts[-1]), in_f, ptr(We_l), out_f, ptr(y), out_f, ptr(bias),
             self.act, 0)
      else:
        W = L['w'][1]
        out_f, in_f = W.shape
        y = b.get('dec_act%d' % j, rows * out_f, torch.float32).view(rows, out_f)
        call('rcd_sgemm', 0, 1, rows, out_f, in_f, ptr(dec_acts[-1]), in_f, ptr(W), in_f, ptr(y), out_f, ptr(bias),
             self.act, 0)
      dec_acts.append(y)
    Y = dec_acts[-1]
    assert Y.shape[1] == H
    return Y.reshape(-1), {'acts': acts, 'dec_acts': dec_acts, 'drop': drop}

  def _mid_backward(self, dY, mid, rows, row0, H, dA, inner_grads, lay):
    """Backward of `_mid_forward`.  dY: gradient w.r.t. the PRE-activation of the last decoding layer when inner
    decoding layers exist, else w.r.t. the decoder input itself.  Writes dA = gradient w.r.t. the pre-activation of
    the embedding encoder and the inner-layer weight / bias gradients into `inner_grads`."""
    b = self.buf
    acts, dec_acts, drop = mid['acts'], mid['dec_acts'], mid['drop']
    k = len(self.enc_layers)
    none = _native.ACT_IDS['none']
    g = dY.view(rows, -1)[:, :dec_acts[-1].shape[1]]
    tied_done = set()
    for j in range(len(self.dec_layers) - 1, -1, -1):      # g = dPre of decoding layer j
      L = self.dec_layers[j]
      o_w, o_b, out_f, in_f = lay['dec'][j]
      x = dec_acts[j]
      call('rcd_colsum', ptr(g), rows, out_f, out_f, ptr(inner_grads[o_b:]))
      gx = b.get('dec_gx%d' % j, rows * in_f, torch.float32).view(rows, in_f)
      if L['w'] is None:
        We_l = self.enc_layers[k - 1 - j]['w'][1]          # [in_f, out_f] stored; dec weight = We_l^T
        # d(We_l) (+)= x^T @ g   ([in_f, out_f]); accumulated with the encoder-side contribution below
        call('rcd_sgemm', 1, 0, in_f, out_f, rows, ptr(x), in_f, ptr(g), out_f, ptr(inner_grads[o_w:]), out_f, None,
             none, 0)
        tied_done.add(k - 1 - j)
        call('rcd_sgemm', 0, 1, rows, in_f, out_f, ptr(g), out_f, ptr(We_l), out_f, ptr(gx), in_f, None, none, 0)
      else:
        W = L['w'][1]
        call('rcd_sgemm', 1, 0, out_f, in_f, rows, ptr(g), out_f, ptr(x), in_f, ptr(inner_grads[o_w:]), in_f, None,
             none, 0)
        call('rcd_sgemm', 0, 0, rows, in_f, out_f, ptr(g), out_f, ptr(W), in_f, ptr(gx), in_f, None, none, 0)
      if j > 0:
        call('rcd_act_grad', ptr(gx), ptr(dec_acts[j]), int(gx.numel()), self.act, ptr(gx))
      g = gx
    # g = gradient w.r.t. dec_acts[0] (the dropped bottleneck)
    top = acts[-1]
    if drop:
      gt = b.get('drop_grad', top.numel(), torch.float32).view(top.shape)
      self._dropout(g, gt, self.dropout_prob, 2, row0 * top.shape[1], self.debug_dropout_keep)
      g = gt
    out = dA.view(rows, H) if k == 0 else b.get('enc_gpre%d' % k, top.numel(), torch.float32).view(top.shape)
    call('rcd_act_grad', ptr(g), ptr(top), int(top.numel()), self.act, ptr(out))
    g = out
    for i in range(k - 1, -1, -1):                         # g = dPre of encoding layer i
      W = self.enc_layers[i]['w'][1]
      o_w, o_b, out_f, in_f = lay['enc'][i]
      x = acts[i]
      call('rcd_colsum', ptr(g), rows, out_f, out_f, ptr(inner_grads[o_b:]))
      call('rcd_sgemm', 1, 0, out_f, in_f, rows, ptr(g), out_f, ptr(x), in_f, ptr(inner_grads[o_w:]), in_f, None, none,
           1 if i in tied_done else 0)
      gx = dA.view(rows, H) if i == 0 else b.get('enc_gpre%d' % i, rows * in_f, torch.float32).view(rows, in_f)
      call('rcd_sgemm', 0, 0, rows, in_f, out_f, ptr(g), out_f, ptr(W), in_f, ptr(gx), in_f, None, none, 0)
      call('rcd_act_grad', ptr(gx), ptr(x), int(gx.numel()), self.act, ptr(gx))
      g = gx

  def _step_inner(self, inner_grads, lay):
    """Optimizer steps of the inner dense layers (dense gradients)."""
    for L, (o_w, o_b, out_f, in_f) in zip(self.enc_layers, lay['enc']):
      self.opt.step_param(L['w'][0], inner_grads[o_w:o_w + out_f * in_f], in_f)
      self.opt.step_param(L['b'][0], inner_grads[o_b:o_b + out_f], 1)
    for L, (o_w, o_b, out_f, in_f) in zip(self.dec_layers, lay['dec']):
      if L['w'] is not None:
        self.opt.step_param(L['w'][0], inner_grads[o_w:o_w + out_f * in_f], in_f)
      self.opt.step_param(L['b'][0], inner_grads[o_b:o_b + out_f], 1)

  def _slab(self, numel, capacity):
    """The step's gradient slab: a grow-only private buffer, or (peer-memory exchange) a view of ONE shared
    allocation of `capacity` floats that every rank has mapped."""
    if self.p2p is None:
      return self.buf.get('slab', numel, torch.float32)
    if self._slab_shared is None or self._slab_shared.nbytes < 4 * numel:
      # collective: every rank sees the same shapes, so every rank (re)allocates at the same step
      torch.cuda.synchronize()
      if self._slab_shared is not None:
        import torch.distributed as dist
        dist.barrier(group=self.p2p.pg)      # no peer is still reading the outgrown slab
        self._slab_shared.close()
      self._slab_shared = self.p2p.shared(4 * max(numel, capacity))
    return self._slab_shared.view(torch.float32, numel)

  def _p2p_reduce(self, name, offset, count):
    """Sum over ranks of slab[offset : offset+count] into a private buffer (small replicated tensors)."""
    out = self.buf.get(name, count, torch.float32)
    call('rcd_p2p_reduce', self._slab_shared.ptr_table(), self.p2p.world, int(offset), int(count), ptr(out), 0)
    return out

  def _reduce_slab(self, slab, loss_slot):
    """Data-parallel exchange: ONE all-reduce over the gradient slab; the loss rides in its last 2 floats
    (hi/lo split of the double, so nothing is lost to fp32)."""
    if self.pg is None:
      return
    reduce_slab(slab, loss_slot, self.pg)

  # ------------------------------------------------------------------------------------------------------
  def _ae_step(self, pool, tpool, row0, rows, inv_b, loss_slot, train):
    b = self.buf
    (en_name, We), (enb_name, be) = self.params['en_w'], self.params['en_b']
    (de_name, Wd), (deb_name, bd) = self.params['de_w'], self.params['de_b']
    H = We.shape[1]
    ldh = _round_up(H, 8)
    n_in, n = pool.n, tpool.n
    same = tpool is pool
    if self.tied and not same:
      raise NotImplementedError('tied weights with a separate target matrix are not supported')
    t_items = tpool.items if tpool.negative_sampling else None
    csc_t = csc_in = None
    if train:
      with self._aux_stream():
        csc_t = self._slice_csc(tpool, row0, rows, n, 't_')
        csc_in = csc_t if same else self._slice_csc(pool, row0, rows, n_in, 'i_')
      self._aux_done()

    # gradient slab: [dWe_rows n_in*H | dWd_rows n*H | dbd n (pad 4) | dbe H (pad 4) | inner layers | pad 2 | loss hi, lo]
    n4, h4 = _round_up(n, 4), _round_up(H, 4)
    inner = self._inner_layout()
    o_wd = n_in * H
    o_bd = o_wd + n * H
    o_be = o_bd + n4
    o_in = o_be + h4
    slab = self._slab(o_in + inner['size'] + 4, 2 * We.shape[0] * H + _round_up(We.shape[0], 4) + h4 + inner['size'] + 4)
    dWe, dWd = slab[0:o_wd], slab[o_wd:o_bd]
    dbd, dbe = slab[o_bd:o_bd + n], slab[o_be:o_be + H]
    inner_grads = slab[o_in:o_in + inner['size']]

    Wg = b.get('Wg', n * ldh, torch.bfloat16)
    bg = b.get('bias_g', n, torch.float32)
    self._wait_ready('de')   # the previous step's W_d / b_d update (side stream; peers' pushes) has landed
    self.opt.catch_up(de_name, t_items, n)      # deferred dense Adam: the batch's rows are brought up to date
    call('rcd_gather_rows', ptr(Wd), H, ptr(t_items), n, 0, ptr(Wg), ldh, None)
    call('rcd_gather_vec', ptr(bd), ptr(t_items), n, ptr(bg))

    general = bool(self.enc_layers) or (train and self.dropout_prob > 0.0)
    base = int(pool.row_ptr_host[row0])
    nnz_in = int(pool.row_ptr_host[row0 + rows]) - base
    in_vals = pool.vals
    if train and self.noise_prob > 0.0 and nnz_in > 0:
      # input noise (nn.py:236-237): dropout on the normalised input == dropout on the stored non-zeros
      in_vals = b.get('noised_vals', pool.vals.numel(), torch.float32)
      self._dropout(pool.vals[base:base + nnz_in], in_vals[base:base + nnz_in], self.noise_prob, 1,
                    base, self.debug_noise_keep)

    Z = b.get('Z', rows * H, torch.float32)
    Zb = b.get('Zb', rows * ldh, torch.bfloat16)
    self._wait_ready('en')
    self.opt.catch_up(en_name, pool.items if pool.negative_sampling else None, n_in)
    call('rcd_ae_encoder_fwd', ptr(We), H, ptr(be), ptr(pool.row_ptr), ptr(pool.raw_items), ptr(in_vals),
         ptr(pool.row_inv_norm), row0, rows, self.act, ptr(Z), None if general else ptr(Zb), ldh)
    mid = None
    Y = Z
    if general:
      Y, mid = self._mid_forward(Z, rows, row0, H, train)
      call('rcd_f32_to_bf16_rows', ptr(Y), rows, H, ptr(Zb), ldh)

    G, ldn, corr, alpha, Zs = self._decoder_and_loss(Zb, ldh, Y, Wg, bg, rows, n, H, inv_b, tpool, row0, loss_slot,
                                                     train)
    if not train:
      return

    # Backward, ordered so that the decoder-side update (HBM- or NVLink-bound) runs on the side stream underneath the
    # tensor-core dgrad GEMM and the encoder backward: sparse dgrad (reads master W_d) -> dW_d -> [W_d, b_d update]
    # || dgrad GEMM -> dA -> dW_e -> [W_e, b_e update].
    partials, splits = self._sparse_dgrad(corr, Wd, tpool, row0, rows, n, H)
    self._wait_ready('csc')
    self._wgrad(G, ldn, Zs, ldh, Y, csc_t, corr, alpha, rows, n, H, dWd, dbd)
    self.last = {'n': n, 'n_in': n_in, 'dWe': dWe.view(n_in, H), 'dWd': dWd.view(n, H), 'dbd': dbd, 'dbe': dbe,
                 'inner': inner_grads, 'inner_layout': inner}
    sequential = self.tied or (self.pg is not None and self.p2p is None)
    if not sequential:
      with self._update_stream():
        self._keep_for_side(pool, tpool)
        if self.p2p is not None:
          self.p2p.barrier(self.bad_flag)          # every rank's dW_d / db_d is complete
          self.opt.step_param_p2p(de_name, self.p2p, self._slab_shared.ptr_table(4 * o_wd), H, tpool.pos,
                                  grads_mc=self._slab_shared.mc(4 * o_wd))
          dbd_sum = self._p2p_reduce('tail_de', o_bd, n4)
          self.opt.step_param(deb_name, dbd_sum[0:n], 1, pos=tpool.pos)
        else:
          self.opt.step_param(de_name, dWd, H, pos=tpool.pos, ids=tpool.items_buf, n_ids=n)
          self.opt.step_param(deb_name, dbd, 1, pos=tpool.pos)
          self._mark_ready('de')

    dA = b.get('dA', rows * H, torch.float32)
    if not general:
      self._dgrad(G, ldn, alpha, Wg, ldh, partials, splits, rows, n, H, Z, self.act, dA, dbe)
    else:
      # gradient w.r.t. the decoder input Y: through its activation when Y is an inner decoding layer's output
      dY = b.get('dY', rows * H, torch.float32)
      self._dgrad(G, ldn, alpha, Wg, ldh, partials, splits, rows, n, H, Y,
                  self.act if self.dec_layers else _native.ACT_IDS['none'], dY, None)
      self._mid_backward(dY, mid, rows, row0, H, dA, inner_grads, inner)
      call('rcd_colsum', ptr(dA), rows, H, H, ptr(dbe))
    csc_ptr, csc_row, csc_val, csc_src = csc_in
    scratch, sbytes, nnz_c = self._heavy_scratch(n_in, csc_row.numel(), H, rows)
    if in_vals is pool.vals:
      call('rcd_ae_encoder_wgrad', ptr(dA), H, ptr(csc_ptr), ptr(csc_row), ptr(csc_val), ptr(pool.row_inv_norm), row0,
           n_in, ptr(dWe), None, None, ptr(scratch), sbytes, nnz_c)
    else:
      call('rcd_ae_encoder_wgrad', ptr(dA), H, ptr(csc_ptr), ptr(csc_row), ptr(csc_val), ptr(pool.row_inv_norm), row0,
           n_in, ptr(dWe), ptr(csc_src), ptr(in_vals[base:]), ptr(scratch), sbytes, nnz_c)

    if self.p2p is not None:
      self._stash_loss(slab, loss_slot)
      with self._update_stream():
        self.p2p.barrier(self.bad_flag)            # dW_e / db_e / loss complete everywhere; W_d pushes have landed
        self._mark_ready('de')
        self.opt.step_param_p2p(en_name, self.p2p, self._slab_shared.ptr_table(0), H, pool.pos,
                                grads_mc=self._slab_shared.mc(0))
        tail = self._p2p_reduce('tail_en', o_be, h4 + inner['size'] + 4)
        self.opt.step_param(enb_name, tail[0:H], 1)
        self._step_inner(tail[h4:h4 + inner['size']], inner)
        loss_slot.copy_(tail[-2:-1].to(torch.float64) + tail[-1:].to(torch.float64))
        self.p2p.barrier(self.bad_flag)            # W_e pushes have landed; the slabs may be overwritten
        self._mark_ready('en')
      return
    if sequential:
      self._reduce_slab(slab, loss_slot)
      if self.tied:  # is_constrained: one table receives both gradients (recoder/nn.py:200)
        dWe.add_(dWd)
        self.opt.step_param(en_name, dWe, H, pos=pool.pos, ids=pool.items_buf, n_ids=n_in)
      else:
        self.opt.step_param(en_name, dWe, H, pos=pool.pos, ids=pool.items_buf, n_ids=n_in)
        self.opt.step_param(de_name, dWd, H, pos=tpool.pos, ids=tpool.items_buf, n_ids=n)
      self.opt.step_param(enb_name, dbe, 1)
      self.opt.step_param(deb_name, dbd, 1, pos=tpool.pos)
      self._step_inner(inner_grads, inner)
      return
    self.opt.step_param(en_name, dWe, H, pos=pool.pos, ids=pool.items_buf, n_ids=n_in)
    self.opt.step_param(enb_name, dbe, 1)
    self._step_inner(inner_grads, inner)

  # --- item-parallel collectives -----------------------------------------------------------------------------------
  def _ip_buffers(self, rows, H):
    """Buffers that cross ranks in the item-parallel step — encoder partial sums Zp [rows*H], dL/dZ (+ loss tail),
    softmax reference and row sums [rows] — in peer-mapped memory when the peer-memory collectives are on."""
    f32 = torch.float32
    nz, r4 = _round_up(rows * H, 4), _round_up(rows, 4)
    ctx = self.ip.p2p
    if ctx is None:
      b = self.buf
      return {'Zp': b.get('Zp', nz, f32), 'dZ': b.get('dZp', nz + 4, f32), 'row_ref': b.get('row_ref', rows, f32),
              'ssum': b.get('stat_sum', rows, f32), 'shared': None, 'off': (0, 0, 0, 0)}
    need = 4 * (nz + nz + 4 + 2 * r4) + 64
    if self._ip_shared is None or self._ip_shared.nbytes < need:
      torch.cuda.synchronize()          # collective: every rank sees the same shapes at the same step
      if self._ip_shared is not None:
        import torch.distributed as dist
        dist.barrier(group=ctx.pg)
        self._ip_shared.close()
      self._ip_shared = ctx.shared(need)
    sh = self._ip_shared
    o_z, o_dz = 0, 4 * nz
    o_ref = o_dz + 4 * (nz + 4)
    o_sum = o_ref + 4 * r4
    return {'Zp': sh.view(f32, nz, o_z), 'dZ': sh.view(f32, nz + 4, o_dz), 'row_ref': sh.view(f32, rows, o_ref),
            'ssum': sh.view(f32, rows, o_sum), 'shared': sh, 'off': (o_z, o_dz, o_ref, o_sum)}

  def _ip_allreduce(self, tag, t, shared, off):
    """In-place sum of `t` over ranks: two-shot peer-memory all-reduce between two barriers, or NCCL."""
    import torch.distributed as dist
    if shared is None:
      timed_all_reduce(tag, t, dist.ReduceOp.SUM, self.ip.pg)
      return
    ctx = self.ip.p2p
    ctx.barrier(self.bad_flag)
    call('rcd_p2p_allreduce', shared.ptr_table(off), shared.mc(off), int(t.numel()), ctx.rank, ctx.world)
    ctx.barrier(self.bad_flag)

  def _ip_reduce_small(self, tag, t, shared, off, op_max):
    """Sum / max of a [rows] vector over ranks; returns the tensor holding the result (a private buffer when the
    peers read `t` directly, `t` itself with NCCL)."""
    import torch.distributed as dist
    if shared is None:
      timed_all_reduce(tag, t, dist.ReduceOp.MAX if op_max else dist.ReduceOp.SUM, self.ip.pg)
      return t
    ctx = self.ip.p2p
    out = self.buf.get(tag + '_out', t.numel(), torch.float32)
    ctx.barrier(self.bad_flag)
    call('rcd_p2p_reduce', shared.ptr_table(off), ctx.world, 0, int(t.numel()), ptr(out), 1 if op_max else 0)
    return out

  # ------------------------------------------------------------------------------------------------------
  def _ae_step_items(self, pool, row0, rows, inv_b, loss_slot, train):
    """Item-parallel autoencoder step (itempar.py): `pool` is the collate of the GLOBAL batch over this rank's item
    shard of the matrix (local item ids; row_inv_norm / row_sum hold the whole-row constants), every parameter tensor
    here is the local shard.  Four collectives: sum of the encoder partials [rows, H], max of the softmax reference
    [rows], sum of the softmax row sums [rows] (NLL only), sum of dL/dZ [rows, H] (+ the loss in its tail)."""
    import torch.distributed as dist
    b = self.buf
    pg = self.ip.pg
    (en_name, We), (enb_name, be) = self.params['en_w'], self.params['en_b']
    (de_name, Wd), (deb_name, bd) = self.params['de_w'], self.params['de_b']
    H = We.shape[1]
    ldh = _round_up(H, 8)
    n = pool.n
    none = _native.ACT_IDS['none']
    nll = self.loss_id == _native.LOSS_IDS['logloss']
    csc = None
    if train:
      with self._aux_stream():
        csc = self._slice_csc(pool, row0, rows, n, 't_')
      self._aux_done()

    Wg = b.get('Wg', n * ldh, torch.bfloat16)
    bg = b.get('bias_g', n, torch.float32)
    self._wait_ready('de')
    self.opt.catch_up(de_name, pool.items, n)
    call('rcd_gather_rows', ptr(Wd), H, ptr(pool.items), n, 0, ptr(Wg), ldh, None)
    call('rcd_gather_vec', ptr(bd), ptr(pool.items), n, ptr(bg))

    # encoder: partial sums over this rank's items -> all-reduce -> bias + activation
    xb = self._ip_buffers(rows, H)
    sh, (o_z, o_dz, o_ref, o_sum) = xb['shared'], xb['off']
    Zp = xb['Zp']
    zero_bias = b.get('zero_bias', H, torch.float32)
    if not getattr(self, '_zero_bias_init', False):
      zero_bias.zero_()
      self._zero_bias_init = True
    self._wait_ready('en')
    self.opt.catch_up(en_name, pool.items, n)
    call('rcd_ae_encoder_fwd', ptr(We), H, ptr(zero_bias), ptr(pool.row_ptr), ptr(pool.raw_items), ptr(pool.vals),
         ptr(pool.row_inv_norm), row0, rows, none, ptr(Zp), None, ldh)
    self._ip_allreduce('allreduce_Z', Zp, sh, o_z)
    Z = b.get('Z', rows * H, torch.float32)
    Zb = b.get('Zb', rows * ldh, torch.bfloat16)
    call('rcd_bias_act', ptr(Zp), ptr(be), rows, H, self.act, ptr(Z), ptr(Zb), ldh)

    # decoder over the local items, loss with the softmax statistics combined across the shards
    ldn = _round_up(n, 8)
    nnz = max(int(pool.row_ptr_host[row0 + rows] - pool.row_ptr_host[row0]), 1)
    G = b.get('G', rows * ldn, torch.bfloat16)
    o_nnz = b.get('o_nnz', nnz, torch.float32)
    corr = b.get('corr', nnz, torch.float32)
    row_ref = xb['row_ref'] if nll else None
    call('rcd_sddmm', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bg), H, ptr(pool.row_ptr), ptr(pool.cols), ptr(pool.vals), row0,
         rows, self.loss_id, self.confidence, inv_b, ptr(o_nnz), ptr(corr), ptr(row_ref))
    if nll:
      row_ref = self._ip_reduce_small('allreduce_rowmax', row_ref, sh, o_ref, True)
    stat_cols = self.lib.rcd_decoder_stat_cols(n)
    stat = b.get('stat', rows * stat_cols, torch.float32)
    call('rcd_decoder_fwd_loss', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bg), rows, n, H, self.loss_id, inv_b, ptr(row_ref),
         ptr(G), ldn, ptr(stat), stat_cols, _native.DEC_MODE_LOSS, None)
    alpha = b.get('alpha', rows, torch.float32) if nll else None
    Zs = b.get('Zs', rows * ldh, torch.bfloat16) if (nll and train) else None
    if nll:
      ssum = xb['ssum']
      call('rcd_rowsum', ptr(stat), rows, stat_cols, stat_cols, ptr(ssum))
      ssum = self._ip_reduce_small('allreduce_rowsum', ssum, sh, o_sum, False)
      stat, stat_ld, stat_n = ssum, 1, 1
    else:
      stat_ld = stat_n = stat_cols
    call('rcd_loss_finish', ptr(stat), stat_ld, stat_n, rows, self.loss_id, self.confidence, inv_b, ptr(row_ref),
         ptr(pool.row_sum), ptr(pool.row_ptr), ptr(pool.vals), ptr(o_nnz), row0, ptr(alpha), ptr(Z), H, ptr(Zs), ldh,
         ptr(loss_slot), ptr(self.bad_flag), 1, None, None, None, None)
    if not train:
      dist.all_reduce(loss_slot, op=dist.ReduceOp.SUM, group=pg)
      return
    Zs = Zs if Zs is not None else Zb

    # backward: local dW_d, then its update on the side stream underneath the dgrad GEMM / all-reduce / encoder backward
    grads = b.get('ip_grads', 2 * n * H + _round_up(n, 4) + _round_up(H, 4), torch.float32)
    dWe, dWd = grads[0:n * H], grads[n * H:2 * n * H]
    dbd = grads[2 * n * H:2 * n * H + n]
    dbe = grads[2 * n * H + _round_up(n, 4):2 * n * H + _round_up(n, 4) + H]
    partials, splits = self._sparse_dgrad(corr, Wd, pool, row0, rows, n, H)
    self._wait_ready('csc')
    self._wgrad(G, ldn, Zs, ldh, Z, csc, corr, alpha, rows, n, H, dWd, dbd)
    self.last = {'n': n, 'n_in': n, 'dWe': dWe.view(n, H), 'dWd': dWd.view(n, H), 'dbd': dbd, 'dbe': dbe}
    with self._update_stream():
      self._keep_for_side(pool)
      self.opt.step_param(de_name, dWd, H, pos=pool.pos, ids=pool.items_buf, n_ids=n)
      self.opt.step_param(deb_name, dbd, 1, pos=pool.pos)
      self._mark_ready('de')

    dZ = xb['dZ']
    self._dgrad(G, ldn, alpha, Wg, ldh, partials, splits, rows, n, H, Z, none, dZ, None)
    self._stash_loss(dZ, loss_slot)            # the loss shares ride in the tail of the dL/dZ all-reduce
    self._ip_allreduce('allreduce_dZ', dZ, sh, o_dz)
    loss_slot.copy_(dZ[-2:-1].to(torch.float64) + dZ[-1:].to(torch.float64))
    dA = b.get('dA', rows * H, torch.float32)
    call('rcd_act_grad', ptr(dZ), ptr(Z), rows * H, self.act, ptr(dA))
    call('rcd_colsum', ptr(dA), rows, H, H, ptr(dbe))
    csc_ptr, csc_row, csc_val, _ = csc
    scratch, sbytes, nnz_c = self._heavy_scratch(n, csc_row.numel(), H, rows)
    call('rcd_ae_encoder_wgrad', ptr(dA), H, ptr(csc_ptr), ptr(csc_row), ptr(csc_val), ptr(pool.row_inv_norm), row0, n,
         ptr(dWe), None, None, ptr(scratch), sbytes, nnz_c)
    self.opt.step_param(en_name, dWe, H, pos=pool.pos, ids=pool.items_buf, n_ids=n)
    self.opt.step_param(enb_name, dbe, 1)

  # ------------------------------------------------------------------------------------------------------
  def _mf_step(self, pool, tpool, row0, rows, inv_b, loss_slot, train):
    b = self.buf
    (u_name, U), (v_name, V), (bias_name, bias) = self.params['user_w'], self.params['item_w'], self.params['bias']
    D = V.shape[1]
    ldd = _round_up(D, 8)
    n = tpool.n
    world, rank = self._world()
    t_items = tpool.items if tpool.negative_sampling else None
    csc = None
    if train:
      with self._aux_stream():
        csc = self._slice_csc(tpool, row0, rows, n, 't_')
      self._aux_done()
    users = pool.users[row0:row0 + rows]

    # slab: [dV_rows n*D | dbias n (pad 4) | dU rows of ALL ranks (other ranks' blocks zero) | pad 2 | loss hi, lo]
    n4 = _round_up(n, 4)
    all_rows = rows * world
    o_b = n * D
    o_u = o_b + n4
    slab = self._slab(o_u + all_rows * D + 4, V.shape[0] * D + _round_up(V.shape[0], 4) + all_rows * D + 4)
    dV, dbias = slab[0:o_b], slab[o_b:o_b + n]
    dU_all = slab[o_u:o_u + all_rows * D]
    lazy_users = self.opt.states[u_name].lazy
    if world > 1 and self.p2p is None:
      dU_all.zero_()       # the other ranks' blocks: the all-reduce of the slab then is a gather
    dU = dU_all[rank * rows * D:(rank + 1) * rows * D]

    Vg = b.get('Wg', n * ldd, torch.bfloat16)
    bg = b.get('bias_g', n, torch.float32)
    self._wait_ready('item')
    self.opt.catch_up(v_name, t_items, n)
    call('rcd_gather_rows', ptr(V), D, ptr(t_items), n, 0, ptr(Vg), ldd, None)
    call('rcd_gather_vec', ptr(bias), ptr(t_items), n, ptr(bg))
    Ue = b.get('Z', rows * D, torch.float32)
    Ub = b.get('Zb', rows * ldd, torch.bfloat16)
    self._wait_ready('user')
    # (data parallel: the user rows of ALL ranks' blocks are updated on every replica, so all of them are caught up)
    self.opt.catch_up(u_name, pool.users[row0 - rank * rows:row0 - rank * rows + all_rows], all_rows)
    if world > 1 and self.p2p is not None and lazy_users:
      # peer-memory exchange with the deferred user table: the peers SUM the slabs' user blocks (a gather, since block q
      # is non-zero in rank q's slab only).  Zeroed here, after the waits above: the peers have finished reading the
      # previous step's slab (barrier at the end of its exchange)
      dU_all.zero_()
    call('rcd_gather_rows', ptr(U), D, ptr(users), rows, self.act, ptr(Ub), ldd, ptr(Ue))
    drop = train and self.dropout_prob > 0.0
    Y = Ue
    if drop:   # dropout on the (activated) user embedding, nn.py:351-352
      Y = b.get('drop_out', rows * D, torch.float32)
      self._dropout(Ue, Y, self.dropout_prob, 2, row0 * D, self.debug_dropout_keep)
      call('rcd_f32_to_bf16_rows', ptr(Y), rows, D, ptr(Ub), ldd)

    G, ldn, corr, alpha, Us = self._decoder_and_loss(Ub, ldd, Y, Vg, bg, rows, n, D, inv_b, tpool, row0, loss_slot,
                                                     train)
    if not train:
      return
    # same schedule as the autoencoder: sparse dgrad (reads master V) -> dV -> [V, bias update on the side stream]
    # || dgrad GEMM -> dU -> [user-table update]
    partials, splits = self._sparse_dgrad(corr, V, tpool, row0, rows, n, D)
    self._wait_ready('csc')
    self._wgrad(G, ldn, Us, ldd, Y, csc, corr, alpha, rows, n, D, dV, dbias)
    self.last = {'n': n, 'dV': dV.view(n, D), 'dbias': dbias, 'dU': dU.view(rows, D)}
    sequential = self.pg is not None and self.p2p is None
    if not sequential:
      with self._update_stream():
        self._keep_for_side(pool, tpool)
        if self.p2p is not None:
          self.p2p.barrier(self.bad_flag)
          self.opt.step_param_p2p(v_name, self.p2p, self._slab_shared.ptr_table(0), D, tpool.pos,
                                  grads_mc=self._slab_shared.mc(0))
          dbias_sum = self._p2p_reduce('tail_de', o_b, n4)
          self.opt.step_param(bias_name, dbias_sum[0:n], 1, pos=tpool.pos)
        else:
          self.opt.step_param(v_name, dV, D, pos=tpool.pos, ids=tpool.items_buf, n_ids=n)
          self.opt.step_param(bias_name, dbias, 1, pos=tpool.pos)
          self._mark_ready('item')

    if not drop:
      self._dgrad(G, ldn, alpha, Vg, ldd, partials, splits, rows, n, D, Ue, self.act, dU, None)
    else:
      dY = b.get('dY', rows * D, torch.float32)
      self._dgrad(G, ldn, alpha, Vg, ldd, partials, splits, rows, n, D, Y, _native.ACT_IDS['none'], dY, None)
      self._dropout(dY, dY, self.dropout_prob, 2, row0 * D, self.debug_dropout_keep)
      call('rcd_act_grad', ptr(dY), ptr(Ue), rows * D, self.act, ptr(dU))

    # In DP the pool holds the GLOBAL batch and rank r works on its r-th block of `rows` rows, so the users
    # of all ranks are the pool rows of the whole global slice.
    base = row0 - rank * rows
    all_users = pool.users[base:base + all_rows]
    upos = b.get('user_pos', U.shape[0], torch.int32)
    if not getattr(self, '_upos_init', False):
      upos.fill_(-1)
      self._upos_init = True
    if self.p2p is not None:
      # user-row gradients are produced by exactly one rank each: block q of dU_all lives in rank q's slab
      self._stash_loss(slab, loss_slot)
      with self._update_stream():
        self.p2p.barrier(self.bad_flag)            # dU / loss complete everywhere; V pushes have landed
        self._mark_ready('item')
        if lazy_users:
          # deferred user table: gather the B_global gradient rows (block q is non-zero in rank q's slab only, so the
          # rank-ordered sum IS the gather) and update those rows on this replica; nothing else of the table moves
          dU_g = self._p2p_reduce('dU_gathered', o_u, all_rows * D)
          self.opt.step_param(u_name, dU_g, D, ids=all_users, n_ids=all_rows)
        else:
          call('rcd_scatter_pos', ptr(all_users), all_rows, ptr(upos), 0)
          self.opt.step_param_p2p(u_name, self.p2p, self._slab_shared.ptr_table(4 * o_u), D, upos,
                                  grad_block_rows=rows)
          call('rcd_scatter_pos', ptr(all_users), all_rows, ptr(upos), 1)
        lsum = self._p2p_reduce('tail_en', slab.numel() - 2, 2)
        loss_slot.copy_(lsum[0:1].to(torch.float64) + lsum[1:2].to(torch.float64))
        self.p2p.barrier(self.bad_flag)            # user-row pushes have landed; the slabs may be overwritten
        self._mark_ready('user')
      return
    if sequential:
      self._reduce_slab(slab, loss_slot)
    call('rcd_scatter_pos', ptr(all_users), all_rows, ptr(upos), 0)
    self.opt.step_param(u_name, dU_all, D, pos=upos, ids=all_users, n_ids=all_rows)
    call('rcd_scatter_pos', ptr(all_users), all_rows, ptr(upos), 1)
    if sequential:
      self.opt.step_param(v_name, dV, D, pos=tpool.pos, ids=tpool.items_buf, n_ids=n)
      self.opt.step_param(bias_name, dbias, 1, pos=tpool.pos)

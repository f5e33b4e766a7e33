"""Data pipeline of the hot path with the reference's interface (recoder/data.py), collate on the GPU.

Same class names, constructor arguments, field names and assertion behaviour as the reference:
`UsersInteractions` (data.py:14-25), `RecommendationDataset` (data.py:28-83), `RecommendationDataLoader`
(data.py:86-167), `Batch` (data.py:170-187), `BatchCollator` (data.py:190-251).  The difference is where the
work happens: the interaction matrix lives in HBM as CSR and `BatchCollator.collate` launches the K1 kernels
(`rcd_collate`), so a `Batch` holds CUDA tensors (plus a handle on the compute layout the trainer consumes).

Attribution: the public interface of this module (class / method names, argument lists and their documentation, log
messages, checkpoint keys) mirrors amoussawi/recoder (MIT License, Copyright (c) 2018 Abdallah Moussawi) so that it is
a drop-in for that library; see LICENSE.  The implementation underneath is original.
"""
import numbers

import numpy as np
import scipy.sparse as sparse
import torch

from . import _native


def _issequence(t):
  return (isinstance(t, (list, tuple)) and (len(t) == 0 or np.isscalar(t[0]))) or \
         (isinstance(t, np.ndarray) and t.ndim == 1)


def _isintlike(x):
  if isinstance(x, numbers.Integral):
    return True
  return isinstance(x, np.ndarray) and x.ndim == 0 and np.issubdtype(x.dtype, np.integer)


# bytes moved over PCIe by the data path since import (bench.py reports them per step)
TRANSFER_BYTES = {'h2d': 0, 'd2h': 0}


class DeviceCSR:
  """A CSR matrix resident in HBM: indptr int64[U+1], indices int32[nnz], data fp32[nnz]."""

  def __init__(self, matrix: sparse.csr_matrix, device=None):
    _native.require_cuda()
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    self.shape = matrix.shape
    self.indptr_host = np.ascontiguousarray(matrix.indptr, dtype=np.int64)
    self.indptr = torch.from_numpy(self.indptr_host).to(self.device)
    self.indices = torch.from_numpy(np.ascontiguousarray(matrix.indices, dtype=np.int32)).to(self.device)
    self.data = torch.from_numpy(np.ascontiguousarray(matrix.data, dtype=np.float32)).to(self.device)
    if self.indices.numel() == 0:  # keep pointers valid
      self.indices = torch.zeros(1, dtype=torch.int32, device=self.device)
      self.data = torch.zeros(1, dtype=torch.float32, device=self.device)

  def pool_nnz(self, users: np.ndarray) -> int:
    return int((self.indptr_host[users + 1] - self.indptr_host[users]).sum())


class HostStagedCSR:
  """A CSR matrix kept in HOST memory (the reference's layout: the dataset is a SciPy matrix and every batch
  is shipped to the device, recoder/model.py:457-462).  `stage(users)` copies the pool's rows into pinned
  staging buffers (K0, `rcd_host_stage_rows`) and sends them H2D; the result is a pool-local DeviceCSR whose
  row r is user `users[r]`."""

  RING = 4   # pool i trains while pools i+1, i+2 are staged / collated; one spare so a slot is never rewritten in flight

  def __init__(self, matrix: sparse.csr_matrix, device=None):
    _native.require_cuda()
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    self.shape = matrix.shape
    self.indptr_host = np.ascontiguousarray(matrix.indptr, dtype=np.int64)
    self.indices_host = np.ascontiguousarray(matrix.indices, dtype=np.int32)
    self.data_host = np.ascontiguousarray(matrix.data, dtype=np.float32)
    self._ring = [None] * self.RING
    self._turn = 0

  def _slot(self, P, nnz):
    i = self._turn % self.RING
    self._turn += 1
    slot = self._ring[i]
    if slot is None or slot['ptr'].numel() < P + 1 or slot['idx'].numel() < nnz:
      cap_p, cap_n = int(P * 1.25) + 1, int(nnz * 1.25) + 16
      slot = {'ptr': torch.empty(cap_p, dtype=torch.int64).pin_memory(),
              'idx': torch.empty(cap_n, dtype=torch.int32).pin_memory(),
              'val': torch.empty(cap_n, dtype=torch.float32).pin_memory(),
              # device side of the staging slot: grow-only, so a step allocates nothing
              'd_ptr': torch.empty(cap_p, dtype=torch.int64, device=self.device),
              'd_idx': torch.empty(cap_n, dtype=torch.int32, device=self.device),
              'd_val': torch.empty(cap_n, dtype=torch.float32, device=self.device),
              'event': None}
      self._ring[i] = slot
    elif slot['event'] is not None:
      slot['event'].synchronize()  # the previous H2D copy out of this slot has finished
    return slot

  def stage_sharded(self, users: np.ndarray, rank: int, world: int, pg):
    """Data-parallel staging: every rank needs the WHOLE pool on its device (the item set of a step is that of the
    global batch) but copies only ITS block of rows over PCIe; the blocks are then exchanged device to
    device (one NCCL all-gather over NVLink).  Layout of the result: block q occupies [q*M, q*M + nnz_q) of the index /
    value arrays (M = the largest block), and the pool-local indptr carries one dummy row per block that covers the
    padding, so the rows the collate gathers are contiguous as usual.  Returns (mini CSR, positions int64[P] of the pool
    rows inside it)."""
    import torch.distributed as dist
    lib = _native.load()
    P = int(users.size)
    per = P // world
    assert per * world == P
    lens = self.indptr_host[users + 1] - self.indptr_host[users]
    block_nnz = lens.reshape(world, per).sum(axis=1)
    M = int(max(int(block_nnz.max()), 1))
    mine = users[rank * per:(rank + 1) * per]
    slot = self._slot(per, M)
    got = lib.rcd_host_stage_rows(self.indptr_host.ctypes.data, self.indices_host.ctypes.data,
                                  self.data_host.ctypes.data, mine.ctypes.data, per, int(self.shape[0]),
                                  int(slot['idx'].numel()), slot['ptr'].data_ptr(), slot['idx'].data_ptr(),
                                  slot['val'].data_ptr())
    if got < 0:
      _native.check(int(got), 'rcd_host_stage_rows')
    sh = slot.get('shard')
    if sh is None or sh['idx'].numel() < world * M or sh['ptr'].numel() < P + world + 1:
      cap_m, cap_p = int(M * 1.25) + 16, P + world + 1
      sh = {'cap_m': cap_m,
            'send': torch.empty(2 * cap_m, dtype=torch.int32, device=self.device),
            'recv': torch.empty(2 * cap_m * world, dtype=torch.int32, device=self.device),
            'idx': torch.empty(cap_m * world, dtype=torch.int32, device=self.device),
            'val': torch.empty(cap_m * world, dtype=torch.float32, device=self.device),
            'ptr': torch.empty(cap_p, dtype=torch.int64, device=self.device),
            'ptr_pin': torch.empty(cap_p, dtype=torch.int64).pin_memory()}
      slot['shard'] = sh
    # own block -> device (indices and value bit patterns packed into one buffer), then one all-gather
    nnz_r = int(block_nnz[rank])
    send = sh['send'][:2 * M]
    send[:nnz_r].copy_(slot['idx'][:nnz_r], non_blocking=True)
    send[M:M + nnz_r].copy_(slot['val'][:nnz_r].view(torch.int32), non_blocking=True)
    recv = sh['recv'][:2 * M * world]
    dist.all_gather_into_tensor(recv, send, group=pg)
    rv = recv.view(world, 2, M)
    idx, val = sh['idx'][:world * M], sh['val'][:world * M]
    idx.view(world, M).copy_(rv[:, 0, :])
    val.view(world, M).copy_(rv[:, 1, :].view(torch.float32))
    # pool-local indptr with one dummy row per block (covers the padding up to the next block)
    ptr = np.empty(P + world + 1, dtype=np.int64)
    starts = np.zeros((world, per + 1), dtype=np.int64)
    np.cumsum(lens.reshape(world, per), axis=1, out=starts[:, 1:])
    starts += (np.arange(world, dtype=np.int64) * M)[:, None]
    ptr[:-1] = starts.reshape(-1)
    ptr[-1] = world * M
    sh['ptr_pin'][:P + world + 1].copy_(torch.from_numpy(ptr))
    d_ptr = sh['ptr'][:P + world + 1]
    d_ptr.copy_(sh['ptr_pin'][:P + world + 1], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    slot['event'] = ev
    mini = DeviceCSR.__new__(DeviceCSR)
    mini.device = self.device
    mini.shape = (P + world, self.shape[1])
    mini.indptr_host = ptr
    mini.indptr = d_ptr
    mini.indices = idx
    mini.data = val
    TRANSFER_BYTES['h2d'] += 2 * max(nnz_r, 1) * 4 + (P + world + 1) * 8
    positions = np.arange(P, dtype=np.int64) + np.arange(P, dtype=np.int64) // per
    return mini, positions

  def stage(self, users: np.ndarray):
    lib = _native.load()
    P = int(users.size)
    nnz = int((self.indptr_host[users + 1] - self.indptr_host[users]).sum())
    slot = self._slot(P, max(nnz, 1))
    got = lib.rcd_host_stage_rows(self.indptr_host.ctypes.data, self.indices_host.ctypes.data,
                                  self.data_host.ctypes.data, users.ctypes.data, P, int(self.shape[0]),
                                  int(slot['idx'].numel()), slot['ptr'].data_ptr(), slot['idx'].data_ptr(),
                                  slot['val'].data_ptr())
    if got < 0:
      _native.check(int(got), 'rcd_host_stage_rows')
    mini = DeviceCSR.__new__(DeviceCSR)
    mini.device = self.device
    mini.shape = (P, self.shape[1])
    m = max(nnz, 1)
    mini.indptr_host = slot['ptr'][:P + 1].numpy()
    mini.indptr = slot['d_ptr'][:P + 1]
    mini.indices = slot['d_idx'][:m]
    mini.data = slot['d_val'][:m]
    mini.indptr.copy_(slot['ptr'][:P + 1], non_blocking=True)
    mini.indices.copy_(slot['idx'][:m], non_blocking=True)
    mini.data.copy_(slot['val'][:m], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    slot['event'] = ev
    TRANSFER_BYTES['h2d'] += (P + 1) * 8 + 2 * max(nnz, 1) * 4
    return mini


class UsersInteractions:
  """
  Holds the interactions of a set of users in an interactions sparse matrix (reference data.py:14-25).

  Args:
    users (np.array): users being represented.
    interactions_matrix (scipy.sparse.csr_matrix): user-item interactions matrix, where ``interactions_matrix[i]``
      correspond to the interactions of ``users[i]``.
  """

  def __init__(self, users, interactions_matrix=None, _source=None):
    self.users = users
    self._matrix = interactions_matrix
    self._source = _source  # (host csr, DeviceCSR getter, row index array) when it comes from a dataset

  @property
  def interactions_matrix(self):
    if self._matrix is None:
      host_csr, _, index = self._source
      self._matrix = RecommendationDataset._extract(host_csr, index)
    return self._matrix

  @property
  def num_rows(self):
    if self._source is not None:
      return len(self._source[2])
    return self._matrix.shape[0]

  @property
  def num_cols(self):
    if self._source is not None:
      return self._source[0].shape[1]
    return self._matrix.shape[1]


class RecommendationDataset:
  """
  Iterates through the users interactions with items (reference data.py:28-83).  Indexing returns a
  :class:`UsersInteractions` of the users in the index; the CSR itself is uploaded to HBM once, on first use.

  Args:
    interactions_matrix (scipy.sparse.csr_matrix): the user-item interactions matrix.
    target_interactions_matrix (scipy.sparse.csr_matrix, optional): the target user-item interactions
      matrix. Mainly used for evaluation, representing the items to recommend.
  """

  def __init__(self, interactions_matrix, target_interactions_matrix=None, device_resident=True):
    self.interactions_matrix = interactions_matrix
    self.target_interactions_matrix = target_interactions_matrix
    # extension: False keeps the matrix in host memory and stages every pool over PCIe (HostStagedCSR)
    self.device_resident = device_resident
    self.users = np.arange(self.interactions_matrix.shape[0])
    self.items = np.arange(self.interactions_matrix.shape[1])
    self._device_csr = None
    self._device_target_csr = None

  def __len__(self):
    return self.interactions_matrix.shape[0]

  def device_csr(self):
    if self._device_csr is None:
      cls = DeviceCSR if self.device_resident else HostStagedCSR
      self._device_csr = cls(self.interactions_matrix)
    return self._device_csr

  def device_target_csr(self):
    if self.target_interactions_matrix is None:
      return None
    if self._device_target_csr is None:
      cls = DeviceCSR if self.device_resident else HostStagedCSR
      self._device_target_csr = cls(self.target_interactions_matrix)
    return self._device_target_csr

  def item_shard_csr(self, rank, world):
    """Item-parallel mode (itempar.py): this rank's column shard of the matrix in HBM (or host-staged), carrying the
    whole-row constants (1/||x_u||, sum_j x_uj) as device vectors indexed by user."""
    key = (rank, world)
    if getattr(self, '_item_shard_key', None) != key:
      from .itempar import shard_matrix_by_items
      local, inv_norm, row_sum = shard_matrix_by_items(self.interactions_matrix.tocsr(), rank, world)
      cls = DeviceCSR if self.device_resident else HostStagedCSR
      csr = cls(local)
      csr.inv_norm_all = torch.from_numpy(inv_norm).to(csr.device)
      csr.row_sum_all = torch.from_numpy(row_sum).to(csr.device)
      self._item_shard = csr
      self._item_shard_key = key
    return self._item_shard

  def __getitem__(self, index):
    assert _issequence(index) or _isintlike(index)  # data.py:51
    users = np.array(index).reshape(-1,)
    input = UsersInteractions(users=users, _source=(self.interactions_matrix, self.device_csr, users))
    if self.target_interactions_matrix is None:
      return input, None
    target = UsersInteractions(users=users, _source=(self.target_interactions_matrix, self.device_target_csr, users))
    return input, target

  @staticmethod
  def _extract(sparse_matrix, index):
    """Host-side row extraction (only used when somebody asks a UsersInteractions for its SciPy matrix)."""
    index = np.asarray(index).reshape(-1)
    return sparse_matrix[index]


class Batch:
  """
  Represents a sparse batch of users and items interactions (reference data.py:170-187).

  Args:
    users (torch.LongTensor): users that are in the batch
    items (torch.LongTensor): items that are in the batch
    indices (torch.LongTensor): the indices of the interactions in the sparse matrix
    values (torch.LongTensor): the values of the interactions
    size (torch.Size): the size of the sparse interactions matrix
  """

  def __init__(self, users, items, indices, values, size, _pool=None, _row0=0):
    self.users = users
    self.items = items
    self._indices = indices
    self.values = values
    self.size = size
    self._pool = _pool  # PoolBatch: compute layout shared by the slices of one pool
    self._row0 = _row0

  @property
  def indices(self):
    """COO indices int64[2, nnz] (data.py:244); built on demand — the trainer consumes the CSR layout."""
    if self._indices is None:
      p = self._pool
      rows = self.size[0]
      lo, hi = p.row_ptr_host[self._row0], p.row_ptr_host[self._row0 + rows]
      out = torch.empty((2, hi - lo), dtype=torch.int64, device=p.cols.device)
      if hi > lo:
        _native.call('rcd_collate_coo', _native.ptr(p.row_ptr), _native.ptr(p.cols), int(self._row0), int(rows),
                     _native.ptr(out))
      self._indices = out
    return self._indices


class PoolBatch:
  """Compute layout of one collated pool (outputs of `rcd_collate`), shared by all its slices."""

  def __init__(self, users_dev, num_rows, num_items, negative_sampling):
    self.users = users_dev
    self.num_rows = num_rows
    self.num_items = num_items
    self.negative_sampling = negative_sampling
    self.row_ptr = self.raw_items = self.cols = self.vals = None
    self.row_inv_norm = self.row_sum = self.pos = self.items_buf = self.counts = None
    self.n = 0
    self.nnz = 0
    self.max_user = -1
    self.row_ptr_host = None
    self._pending = None
    self.next_hint = None          # (next PoolBatch, next target PoolBatch) while that pool is being collated
    self.last_slice_row0 = -1

  @property
  def items(self):
    return self.items_buf[:self.n]


class PoolRing:
  """Grow-only device buffers for the collate outputs of successive pools (4 slots: the pool being trained on, the
  two collated / being collated ahead of it, one spare).  The training loop collates thousands of pools of nearly equal size: without the
  ring every pool costs a dozen allocator round trips on the host, which is what bounds small configurations."""

  SLOTS = 4

  def __init__(self):
    self._slots = [dict() for _ in range(self.SLOTS)]
    self._turn = 0

  def next_slot(self):
    self._turn += 1
    return self._slots[self._turn % self.SLOTS]

  @staticmethod
  def take(slot, name, numel, dtype, device):
    t = slot.get(name)
    if t is None or t.numel() < numel or t.dtype != dtype:
      t = torch.empty(max(int(numel * 1.25) + 16, 1), dtype=dtype, device=device)
      slot[name] = t
    return t[:numel]


def collate_pool_launch(csr, users, negative_sampling: bool, stream=None, after=(), ring=None,
                        row_constants=None, table_rows=None, stage_shard=None) -> PoolBatch:
  """Enqueues K1 on the rows `users` of `csr` (DeviceCSR, or HostStagedCSR: staged over PCIe first) and an
  asynchronous read-back of the two counts (n, nnz) every downstream shape depends on.  The returned PoolBatch is
  usable after `collate_pool_finish`.  Launching the collate of pool i+1 before the training step of pool i is
  enqueued hides the read-back behind that step (Recoder._pool_steps).

  `stream`: run the collate kernels on this (auxiliary) CUDA stream so that they overlap the training step the
  caller enqueues next on the current stream; the outputs are allocated on the CURRENT stream's pool and the auxiliary
  stream first waits for the current stream and for every stream in `after`, so recycled memory is never written
  while an earlier kernel still reads it.  `collate_pool_finish` makes the current stream wait for the collate.

  `row_constants` = (inv_norm_all, row_sum_all), device vectors indexed by USER id: item-parallel mode, where `csr`
  holds only this rank's columns and the row statistics of the collate must be replaced by whole-row values.

  `table_rows`: rows of the embedding tables the pool will train (the model's `num_items`).  The reference allows a
  model wider than the matrix (`assert num_items >= max item id + 1`, recoder/model.py:241): the item -> column map
  `pos`, which the optimizer kernels read for EVERY table row, is then `table_rows` long with -1 beyond the matrix
  width.  A matrix wider than the model is an error (the reference fails inside its embedding lookup).

  `stage_shard` = (rank, world, process group): data-parallel runs on a host-resident matrix — every rank stages only
  its own block of the pool's rows and the blocks are all-gathered device to device (HostStagedCSR.stage_sharded)."""
  users = np.ascontiguousarray(np.asarray(users).reshape(-1), dtype=np.int64)
  assert users.size > 0
  assert users.min() >= 0 and users.max() < csr.shape[0], 'user index out of range'
  dev = csr.device
  P, I = int(users.size), int(csr.shape[1])
  if table_rows is not None and I > int(table_rows):
    raise ValueError('the interactions matrix has %d columns but the model represents only %d items' % (I, table_rows))
  pos_len = max(I, int(table_rows or 0))
  max_user = int(users.max())
  if row_constants is None and hasattr(csr, 'inv_norm_all'):
    row_constants = (csr.inv_norm_all, csr.row_sum_all)
  slot = ring.next_slot() if ring is not None else None

  def buf(name, numel, dtype):
    if slot is None:
      return torch.empty(max(numel, 1), dtype=dtype, device=dev)[:numel]
    return PoolRing.take(slot, name, numel, dtype, dev)

  import contextlib
  ctx = contextlib.nullcontext
  if stream is not None:
    stream.wait_stream(torch.cuda.current_stream())
    for other in after:
      if other is not None:
        stream.wait_stream(other)
    ctx = lambda: torch.cuda.stream(stream)  # noqa: E731
  # with ring buffers nothing below allocates, so the H2D copies can ride on the auxiliary stream as well
  copy_ctx = ctx if slot is not None else contextlib.nullcontext
  with copy_ctx():
    if slot is None:
      users_dev = torch.from_numpy(users).to(dev, non_blocking=True)
    else:
      pin = slot.get('users_pin')
      if pin is None or pin.numel() < P:
        pin = torch.empty(int(P * 1.25) + 16, dtype=torch.int64).pin_memory()
        slot['users_pin'] = pin
      pin[:P].copy_(torch.from_numpy(users))
      users_dev = buf('users', P, torch.int64)
      users_dev.copy_(pin[:P], non_blocking=True)
    TRANSFER_BYTES['h2d'] += P * 8
    rows_dev = users_dev
    if isinstance(csr, HostStagedCSR):
      if stage_shard is not None and stage_shard[1] > 1 and P % stage_shard[1] == 0:
        csr, users = csr.stage_sharded(users, *stage_shard)
        rows_dev = _positions_cached(P, stage_shard[1], dev)
      else:
        csr = csr.stage(users)
        users = np.arange(P, dtype=np.int64)
        rows_dev = _arange_cached(P, dev)
  lens = csr.indptr_host[users + 1] - csr.indptr_host[users]
  nnz = int(lens.sum())
  assert nnz < 2 ** 31, 'pool too large'
  pb = PoolBatch(users_dev, P, I, negative_sampling)
  row_ptr_host = np.zeros(P + 1, dtype=np.int64)
  np.cumsum(lens, out=row_ptr_host[1:])
  pb.row_ptr_host = row_ptr_host
  cap = max(nnz, 1)
  pb.row_ptr = buf('row_ptr', P + 1, torch.int32)
  pb.raw_items = buf('raw_items', cap, torch.int32)
  pb.cols = buf('cols', cap, torch.int32)
  pb.vals = buf('vals', cap, torch.float32)
  pb.row_inv_norm = buf('row_inv_norm', P, torch.float32)
  pb.row_sum = buf('row_sum', P, torch.float32)
  pb.pos = buf('pos', pos_len, torch.int32)
  pb.max_user = max_user
  pb.items_buf = buf('items_buf', min(I, cap) if negative_sampling else I, torch.int64)
  pb.counts = buf('counts', 2, torch.int32)
  lib = _native.load()
  sbytes = lib.rcd_collate_scratch_bytes(P, I)
  scratch = buf('scratch', sbytes, torch.uint8)
  counts_host = _pinned_counts()
  if stream is not None and slot is None:
    stream.wait_stream(torch.cuda.current_stream())   # the copies above went to the current stream
  with ctx():
    pb.counts.zero_()
    if pos_len > I:
      pb.pos[I:].fill_(-1)   # table rows beyond the matrix width never receive a gradient
    _native.call('rcd_collate', _native.ptr(csr.indptr), _native.ptr(csr.indices), _native.ptr(csr.data),
                 _native.ptr(rows_dev), P, I, int(bool(negative_sampling)), cap, _native.ptr(pb.row_ptr),
                 _native.ptr(pb.raw_items), _native.ptr(pb.cols), _native.ptr(pb.vals), _native.ptr(pb.row_inv_norm),
                 _native.ptr(pb.row_sum), _native.ptr(pb.pos), _native.ptr(pb.items_buf), _native.ptr(pb.counts),
                 _native.ptr(scratch), sbytes)
    if row_constants is not None:
      _native.call('rcd_gather_vec', _native.ptr(row_constants[0]), _native.ptr(users_dev), P,
                   _native.ptr(pb.row_inv_norm))
      _native.call('rcd_gather_vec', _native.ptr(row_constants[1]), _native.ptr(users_dev), P, _native.ptr(pb.row_sum))
    counts_host.copy_(pb.counts, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
  TRANSFER_BYTES['d2h'] += 8
  pb._pending = (counts_host, ev, nnz, (csr, rows_dev, scratch))  # keeps the kernel inputs alive until finish
  return pb


_ARANGE_CACHE = {}


def _arange_cached(n, dev):
  t = _ARANGE_CACHE.get(dev)
  if t is None or t.numel() < n:
    t = torch.arange(int(n * 1.25) + 16, dtype=torch.int64, device=dev)
    _ARANGE_CACHE[dev] = t
  return t[:n]


_POSITIONS_CACHE = {}


def _positions_cached(P, world, dev):
  """Positions of the pool rows inside a sharded staging layout (one dummy row after every block)."""
  key = (P, world, dev)
  t = _POSITIONS_CACHE.get(key)
  if t is None:
    per = P // world
    t = (torch.arange(P, dtype=torch.int64) + torch.arange(P, dtype=torch.int64) // per).to(dev)
    _POSITIONS_CACHE[key] = t
  return t


_COUNTS_RING = []
_COUNTS_TURN = [0]


def _pinned_counts():
  if not _COUNTS_RING:
    for _ in range(8):
      _COUNTS_RING.append(torch.zeros(2, dtype=torch.int32).pin_memory())
  _COUNTS_TURN[0] += 1
  return _COUNTS_RING[_COUNTS_TURN[0] % len(_COUNTS_RING)]


def collate_pool_finish(pb: PoolBatch) -> PoolBatch:
  """Waits for the counts of a launched collate (the one host sync of the collate)."""
  if pb._pending is None:
    return pb
  counts_host, ev, nnz, _ = pb._pending
  torch.cuda.current_stream().wait_event(ev)   # no-op when the collate ran on the current stream
  ev.synchronize()
  pb.n = int(counts_host[0])
  pb.nnz = int(counts_host[1])
  pb._pending = None
  assert pb.nnz == nnz, 'device/host nnz mismatch'
  return pb


def collate_pool(csr, users, negative_sampling: bool, table_rows=None) -> PoolBatch:
  """Runs K1 on the rows `users` of `csr` and returns the pool's compute layout."""
  return collate_pool_finish(collate_pool_launch(csr, users, negative_sampling, table_rows=table_rows))


def pool_of(users_interactions, negative_sampling):
  """Collates a whole :class:`UsersInteractions` on the GPU; returns (PoolBatch, the CSR it came from)."""
  if users_interactions._source is not None:
    _, get_csr, index = users_interactions._source
    csr, rows = get_csr(), index
  else:
    m = users_interactions.interactions_matrix.tocsr()
    csr, rows = DeviceCSR(m), np.arange(m.shape[0])
  return collate_pool(csr, rows, negative_sampling), csr


class BatchCollator:
  """
  Collator of :class:`UsersInteractions` into multiple :class:`Batch` based on ``batch_size``
  (reference data.py:190-251), executed on the GPU.

  Args:
    batch_size (int): number of samples per batch
    negative_sampling (bool, optional): whether to apply mini-batch based negative sampling or not.
  """

  def __init__(self, batch_size, negative_sampling=False):
    self.batch_size = batch_size
    self.negative_sampling = negative_sampling

  def collate(self, users_interactions):
    """
    Collates :class:`UsersInteractions` into batches of size ``batch_size``.

    Returns:
      list[Batch]: list of batches (``items`` shared by all slices, data.py:246).
    """
    pb, csr = pool_of(users_interactions, self.negative_sampling)
    batch_users = torch.as_tensor(np.asarray(users_interactions.users), dtype=torch.int64, device=csr.device)
    vector_dim = pb.n if self.negative_sampling else csr.shape[1]
    items = pb.items if self.negative_sampling else None
    slices = []
    P = pb.num_rows
    for offset in range(0, P, self.batch_size):
      hi = min(offset + self.batch_size, P)
      lo_n, hi_n = int(pb.row_ptr_host[offset]), int(pb.row_ptr_host[hi])
      slices.append(Batch(users=batch_users[offset:hi], items=items, indices=None, values=pb.vals[lo_n:hi_n],
                          size=torch.Size([hi - offset, vector_dim]), _pool=pb, _row0=offset))
    return slices


class RecommendationDataLoader:
  """
  Generates batches with mini-batch negative sampling (reference data.py:86-167).

  The sampling order follows the reference: ``RandomSampler`` over the dataset (torch global RNG), grouped into
  pools of ``num_sampling_users`` (data.py:124-126); every pool is collated at once and yielded one
  ``batch_size`` slice at a time (data.py:138-144).  ``num_workers`` is accepted for interface compatibility;
  the collate runs on the GPU, so no worker processes are forked.

  Args:
    dataset (RecommendationDataset): dataset from which to load the data
    batch_size (int): number of samples per batch
    negative_sampling (bool, optional): whether to apply mini-batch based negative sampling or not.
    num_sampling_users (int, optional): number of users to consider for mini-batch based negative
      sampling. If 0, then num_sampling_users will be equal to batch_size.
    num_workers (int, optional): ignored (see above).
    collate_fn (callable, optional): A function that transforms a :class:`UsersInteractions` into a mini-batch.
    user_order (callable, optional): ``epoch -> np.ndarray`` explicit user order (tests, benchmarks).
  """

  def __init__(self, dataset, batch_size, negative_sampling=False, num_sampling_users=0, num_workers=0,
               collate_fn=None, user_order=None):
    self.dataset = dataset
    self.num_sampling_users = num_sampling_users
    self.num_workers = num_workers
    self.batch_size = batch_size
    self.negative_sampling = negative_sampling
    if self.num_sampling_users == 0:
      self.num_sampling_users = batch_size
    assert self.num_sampling_users >= batch_size, 'num_sampling_users should be at least equal to the batch_size'
    self.batch_collator = BatchCollator(batch_size=self.batch_size, negative_sampling=self.negative_sampling)
    if collate_fn is None:
      self._collate_fn = self.batch_collator.collate
      self._use_default_data_generator = True
    else:
      self._collate_fn = collate_fn
      self._use_default_data_generator = False
    self._user_order = user_order
    self._epoch = 0

  def _sample_order(self):
    """torch.utils.data.RandomSampler semantics: a generator seeded from the global RNG, then randperm."""
    self._epoch += 1
    if self._user_order is not None:
      return np.asarray(self._user_order(self._epoch), dtype=np.int64)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    gen = torch.Generator()
    gen.manual_seed(seed)
    return torch.randperm(len(self.dataset), generator=gen).numpy()

  def pools(self):
    """Yields the index list of every sampling pool of one epoch (data.py:124-126, drop_last=False)."""
    order = self._sample_order()
    for off in range(0, len(order), self.num_sampling_users):
      yield order[off:off + self.num_sampling_users]

  def _collated(self):
    for index in self.pools():
      input_ui, target_ui = self.dataset[index]
      input = self._collate_fn(input_ui)
      target = None if target_ui is None else self._collate_fn(target_ui)
      yield input, target

  def _default_data_generator(self):
    for input, target in self._collated():
      for batch_ind in range(len(input)):
        if target is None:
          yield input[batch_ind], None
        else:
          yield input[batch_ind], target[batch_ind]

  def __iter__(self):
    if self._use_default_data_generator:
      return self._default_data_generator()
    return self._collated()

  def __len__(self):
    return int(np.ceil(len(self.dataset) / self.batch_collator.batch_size))

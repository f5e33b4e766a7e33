"""Deterministic synthetic user x item interaction matrices.

Power-law item popularity (item = floor(I * u**gamma), gamma = 2), binary fp32 values, sorted column indices,
no duplicates and no stored zeros (the reference crashes on explicit zeros: recoder/data.py:215 vs :238-242).
Rows are generated in fixed-size chunks whose RNG stream depends only on (seed, chunk index), so any user
prefix of a larger matrix is bit-identical to generating that prefix alone — the CPU baseline can run on a
prefix of the very matrix the GPU run uses.
"""
import numpy as np

CHUNK_USERS = 16384


def synthetic_csr(num_users: int, num_items: int, nnz_per_user: int, seed: int = 1234, gamma: float = 2.0):
  """Returns (indptr int64 [U+1], indices int32 [nnz], data float32 [nnz])."""
  counts = np.zeros(num_users, dtype=np.int64)
  idx_chunks = []
  for c, start in enumerate(range(0, num_users, CHUNK_USERS)):
    rows = min(CHUNK_USERS, num_users - start)
    rng = np.random.default_rng([seed, c])
    u = rng.random((CHUNK_USERS, nnz_per_user), dtype=np.float64)[:rows]
    items = np.minimum(np.floor(num_items * np.power(u, gamma)), num_items - 1).astype(np.int32)
    items.sort(axis=1)
    keep = np.ones(items.shape, dtype=bool)
    keep[:, 1:] = items[:, 1:] != items[:, :-1]
    counts[start:start + rows] = keep.sum(axis=1)
    idx_chunks.append(items[keep])
  indptr = np.zeros(num_users + 1, dtype=np.int64)
  np.cumsum(counts, out=indptr[1:])
  indices = np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, dtype=np.int32)
  data = np.ones(indices.shape[0], dtype=np.float32)
  return indptr, indices, data


def to_scipy(indptr, indices, data, num_items):
  import scipy.sparse as sp
  m = sp.csr_matrix((data, indices, indptr), shape=(len(indptr) - 1, num_items))
  m.has_sorted_indices = True
  return m


def epoch_user_order(num_users: int, epoch: int) -> np.ndarray:
  """Per-epoch user permutation shared by the GPU run and the CPU baseline."""
  import torch
  return torch.randperm(num_users, generator=torch.Generator().manual_seed(epoch)).numpy()

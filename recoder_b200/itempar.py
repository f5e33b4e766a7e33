"""Item-parallel training (the scaling mode of the multi-GPU path).

Rows (users) of a global batch are processed by EVERY rank; the item axis — the embedding tables W_e, W_d, b_d, their
optimizer state, the interaction matrix columns, and hence the decoder GEMM's N dimension — is sharded cyclically
(item i lives on rank i % R as local id i // R).  What crosses NVLink per step is then two [B_global, H] activations
(the encoder partial sums and dL/dZ) and two [B_global] vectors (softmax reference max and row sums) instead of the
2·n·H gradient slab and the updated tables: 67 MB instead of ≈1.6 GB per step at C3 / 8 GPUs, and the Adam pass
touches I/R rows.  Numerically it is the reference's step with `batch_size = B_global`, like the row-sharded mode.
"""
import numpy as np
import scipy.sparse as sparse
import torch


def shard_matrix_by_items(matrix: sparse.csr_matrix, rank: int, world: int):
  """Columns i with i % world == rank of a CSR matrix, renumbered i // world (stored order kept), plus the per-user
  constants that need the WHOLE row: 1/max(||x_u||_2, 1e-12) (F.normalize, recoder/nn.py:235) and sum_j x_uj."""
  indptr = np.asarray(matrix.indptr, dtype=np.int64)
  indices = np.asarray(matrix.indices)
  data = np.asarray(matrix.data, dtype=np.float32)
  U, I = matrix.shape
  sq = np.concatenate([[0.0], np.cumsum(data.astype(np.float64) ** 2)])
  sm = np.concatenate([[0.0], np.cumsum(data.astype(np.float64))])
  norm = np.sqrt(sq[indptr[1:]] - sq[indptr[:-1]])
  # per-row sums in float64 differences are exact enough for fp32 (rows hold O(100) small values)
  inv_norm = (1.0 / np.maximum(norm, 1e-12)).astype(np.float32)
  row_sum = (sm[indptr[1:]] - sm[indptr[:-1]]).astype(np.float32)
  keep = (indices % world) == rank
  kept_before = np.concatenate([[0], np.cumsum(keep, dtype=np.int64)])
  local_indptr = kept_before[indptr]
  local = sparse.csr_matrix((data[keep], (indices[keep] // world).astype(np.int32), local_indptr),
                            shape=(U, (I - rank + world - 1) // world))
  local.has_sorted_indices = bool(getattr(matrix, 'has_sorted_indices', False))
  return local, inv_norm, row_sum


class ItemParallel:
  """Per-process state of the item-parallel mode: rank / world, local shards of the item-indexed parameters."""

  def __init__(self, pg):
    import torch.distributed as dist
    self.pg = pg
    self.world = dist.get_world_size(pg)
    self.rank = dist.get_rank(pg)
    self.sharded = {}     # parameter name -> local shard tensor
    self.p2p = None       # p2p.P2PContext: the step's four collectives run as peer-memory kernels instead of NCCL

  def local_rows(self, rows):
    return (rows - self.rank + self.world - 1) // self.world

  def shard(self, name, full):
    """Local shard (rows rank, rank+R, ...) of an item-indexed tensor, as its own contiguous storage."""
    t = full[self.rank::self.world].contiguous()
    self.sharded[name] = t
    return t

  def shard_like(self, full):
    return full[self.rank::self.world].contiguous()

  def gather_full(self, local, rows):
    """All ranks' shards interleaved back into the full item-indexed tensor (collective)."""
    import torch.distributed as dist
    per = (rows + self.world - 1) // self.world
    shape = (per,) + tuple(local.shape[1:])
    padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(self.world)]
    dist.all_gather(parts, padded, group=self.pg)
    full = torch.empty((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for q in range(self.world):
      nq = (rows - q + self.world - 1) // self.world
      full[q::self.world] = parts[q][:nq]
    return full

  def sync_to_full(self, named_full):
    """Writes the current shards back into the full parameters (before evaluation / checkpoints)."""
    for name, local in self.sharded.items():
      full = named_full[name]
      full.copy_(self.gather_full(local, full.shape[0]))

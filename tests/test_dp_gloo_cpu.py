"""Host-side logic of the data-parallel path on CPU with the gloo backend, world_size 2: the row sharding of a
global batch and the single slab all-reduce that also carries the float64 loss."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recoder_b200.engine import reduce_slab, shard_rows


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    g = torch.Generator().manual_seed(100 + rank)
    slab = torch.randn(1001 + 4, generator=g, dtype=torch.float32)
    mine = slab.clone()
    loss = torch.tensor([1234.5678901234567 * (rank + 1) + 1e-9 * rank], dtype=torch.float64)
    mine_loss = loss.clone()
    reduce_slab(slab, loss, dist.group.WORLD)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    losses = [torch.zeros_like(mine_loss) for _ in range(world)]
    dist.all_gather(losses, mine_loss)
    want = sum(gathered)[:-2]
    ok_grad = torch.allclose(slab[:-2], want, rtol=0, atol=1e-6)
    want_loss = float(sum(losses).item())
    ok_loss = abs(float(loss.item()) - want_loss) <= 1e-9 * abs(want_loss)  # far below fp32 resolution (6e-8)
    # sharding: both ranks agree on shapes, ranges are disjoint and contiguous
    mine_rows = list(shard_rows(100, 32, world, rank))
    out[rank] = (ok_grad, ok_loss, mine_rows)
  finally:
    dist.destroy_process_group()


def test_slab_allreduce_and_sharding_world2():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  assert all(out[r][0] for r in range(world)), 'gradient slab sum mismatch'
  assert all(out[r][1] for r in range(world)), 'loss hi/lo transport lost precision'
  r0, r1 = out[0][2], out[1][2]
  assert len(r0) == len(r1) == 4
  for (a0, an, ag), (b0, bn, bg) in zip(r0, r1):
    assert an == bn and ag == bg == an * 2
    assert b0 == a0 + an
  assert r0[0] == (0, 16, 32) and r0[-1] == (96, 2, 4)


@pytest.mark.parametrize('P,g,world', [(2048, 2048, 1), (4096, 1024, 4), (1000, 512, 8), (7, 16, 8)])
def test_shard_rows_partition(P, g, world):
  covered = []
  shapes = None
  for rank in range(world):
    rows = list(shard_rows(P, g, world, rank))
    sig = [(n, gr) for _, n, gr in rows]
    shapes = sig if shapes is None else shapes
    assert sig == shapes
    for r0, n, _ in rows:
      covered.extend(range(r0, r0 + n))
  assert len(covered) == len(set(covered))
  assert all(0 <= r < P for r in covered)
  dropped = P - len(covered)
  assert dropped < world * ((P + g - 1) // g)


# ---- item-parallel host logic (recoder_b200/itempar.py) -----------------------------------------------------------
def _items_worker(rank, world, port, out):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from recoder_b200.itempar import ItemParallel
    ip = ItemParallel(dist.group.WORLD)
    rows = 11                      # not a multiple of the world size: the last shard is shorter
    full = torch.arange(rows * 3, dtype=torch.float32).view(rows, 3)
    local = ip.shard('table', full)
    ok_shard = torch.equal(local, full[rank::world]) and ip.local_rows(rows) == local.shape[0]
    local.mul_(2.0)                # every rank updates only its own rows
    back = ip.gather_full(local, rows)
    ok_gather = torch.equal(back, full * 2.0)
    target = {'table': torch.zeros_like(full)}
    ip.sync_to_full(target)
    out[rank] = (ok_shard, ok_gather, torch.equal(target['table'], full * 2.0))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_item_shards_gather_back(world):
  mgr = mp.Manager()
  out = mgr.dict()
  mp.spawn(_items_worker, args=(world, _free_port(), out), nprocs=world, join=True)
  assert all(all(out[r]) for r in range(world)), dict(out)


def test_shard_matrix_by_items_partitions_the_matrix():
  from recoder_b200.itempar import shard_matrix_by_items
  from recoder_b200.synth import synthetic_csr, to_scipy
  indptr, indices, data = synthetic_csr(400, 257, 25, seed=3)
  data = (1 + np.arange(len(data)) % 5).astype(np.float32)           # ratings, so the row constants are not trivial
  m = to_scipy(indptr, indices, data, 257)
  dense = m.toarray()
  world = 4
  total = 0
  for r in range(world):
    local, inv_norm, row_sum = shard_matrix_by_items(m, r, world)
    assert local.shape == (400, len(range(r, 257, world)))
    assert np.array_equal(local.toarray(), dense[:, r::world])        # columns r, r+R, ... renumbered 0, 1, ...
    assert np.all(np.diff(local.indices[local.indptr[5]:local.indptr[6]]) > 0)   # stored order kept (sorted)
    np.testing.assert_allclose(inv_norm, 1.0 / np.maximum(np.linalg.norm(dense, axis=1), 1e-12), rtol=1e-6)
    np.testing.assert_allclose(row_sum, dense.sum(axis=1), rtol=1e-6)
    total += local.nnz
  assert total == m.nnz


@pytest.mark.parametrize('U,batch,pool_batches,world', [(1000, 64, 1, 1), (1000, 64, 2, 2), (1027, 128, 2, 8),
                                                        (1025, 128, 1, 2), (1024, 128, 4, 8), (5, 2, 1, 4), (130, 16, 3, 4)])
def test_steps_per_pass_counts_what_the_trainer_yields(U, batch, pool_batches, world):
  """`Recoder._train` bounds its epochs by the number of optimizer steps in one pass over the data: it has to equal what
  `_pool_steps` yields — pools of num_sampling_users * world users cut by `shard_rows` — also when the last slice has
  fewer users than there are ranks (the step every rank skips)."""
  from recoder_b200.model import steps_per_pass
  g, S = batch * world, batch * pool_batches * world
  for rank in range(world):
    yielded = sum(len(list(shard_rows(min(S, U - off), g, world, rank))) for off in range(0, U, S))
    assert yielded == steps_per_pass(U, g, world, rows_sharded=world > 1)
  assert steps_per_pass(U, g, world, rows_sharded=False) == int(np.ceil(U / g))     # the reference's len(dataloader)

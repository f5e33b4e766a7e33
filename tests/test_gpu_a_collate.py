"""K1 parity: the GPU collate is bit-exact against the oracle restatement of BatchCollator.collate and against
the golden vectors recorded from the reference."""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200.data import BatchCollator, RecommendationDataset, RecommendationDataLoader
from recoder_b200.synth import synthetic_csr, to_scipy
from tests.golden_util import Golden, case_names

pytestmark = pytest.mark.gpu


def _check(batches, ref_batches):
  assert len(batches) == len(ref_batches)
  for b, r in zip(batches, ref_batches):
    assert b.users.dtype == torch.int64 and np.array_equal(b.users.cpu().numpy(), r.users)
    if r.items is None:
      assert b.items is None
    else:
      assert b.items.dtype == torch.int64 and np.array_equal(b.items.cpu().numpy(), r.items)
    assert b.indices.dtype == torch.int64 and np.array_equal(b.indices.cpu().numpy(), r.indices)
    assert b.values.dtype == torch.float32 and np.array_equal(b.values.cpu().numpy(), r.values)
    assert tuple(b.size) == tuple(r.size)


@pytest.mark.parametrize('name', case_names())
def test_collate_matches_golden(name):
  g = Golden(name)
  m = g.meta
  ds = RecommendationDataset(to_scipy(g.indptr, g.indices, g.data, m['num_items']))
  collator = BatchCollator(batch_size=m['batch'], negative_sampling=m['neg'])
  for users, steps in g.pools():
    ui, _ = ds[users]
    batches = collator.collate(ui)
    assert len(batches) == len(steps)
    for b, s in zip(batches, steps):
      ref = g.step(s)
      assert np.array_equal(b.users.cpu().numpy(), ref['users'])
      if ref['items'] is None:
        assert b.items is None
      else:
        assert np.array_equal(b.items.cpu().numpy(), ref['items'])
      assert np.array_equal(b.indices.cpu().numpy(), ref['indices'])
      assert np.array_equal(b.values.cpu().numpy(), ref['values'])
      assert tuple(b.size) == ref['size']


@pytest.mark.parametrize('U,I,nnz,batch,pool,neg', [
  (100, 200, 10, 13, 13, True), (100, 200, 10, 5, 10, True), (100, 200, 10, 1, 1, True),
  (257, 1000, 40, 64, 128, True), (300, 5000, 3, 100, 100, False), (2000, 30000, 120, 512, 512, True),
])
def test_collate_matches_oracle(U, I, nnz, batch, pool, neg):
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=U + I)
  rng = np.random.default_rng(0)
  data = rng.integers(1, 6, size=data.shape[0]).astype(np.float32)  # non-binary values
  ds = RecommendationDataset(to_scipy(indptr, indices, data, I))
  order = rng.permutation(U)
  collator = BatchCollator(batch_size=batch, negative_sampling=neg)
  for off in range(0, min(U, 4 * pool), pool):
    users = order[off:off + pool]
    ui, _ = ds[users]
    _check(collator.collate(ui), O.collate(indptr, indices, data, I, users, batch, neg))


def test_collate_unsorted_rows_and_empty_rows():
  # stored (unsorted) column order inside a row must be preserved (data.py:236-242), empty rows are legal
  indptr = np.array([0, 3, 3, 5, 9], dtype=np.int64)
  indices = np.array([7, 2, 5, 9, 0, 3, 8, 1, 2], dtype=np.int32)
  data = np.arange(1, 10, dtype=np.float32)
  import scipy.sparse as sp
  m = sp.csr_matrix((data, indices, indptr), shape=(4, 10))
  ds = RecommendationDataset(m)
  users = np.array([3, 1, 0, 2])
  ui, _ = ds[users]
  _check(BatchCollator(2, True).collate(ui), O.collate(indptr, indices, data, 10, users, 2, True))


def test_dataloader_pool_semantics():
  """Restates the reference's tests/test_data.py:89-126 for the GPU loader."""
  indptr, indices, data = synthetic_csr(103, 300, 8, seed=3)
  ds = RecommendationDataset(to_scipy(indptr, indices, data, 300), to_scipy(indptr, indices, data, 300))
  for batch_size, nsu in [(5, 0), (5, 10)]:
    dl = RecommendationDataLoader(ds, batch_size=batch_size, negative_sampling=True, num_sampling_users=nsu)
    count = 0
    for batch_idx, (inp, tgt) in enumerate(dl, 1):
      assert tgt is not None
      assert inp.size[0] == batch_size or (batch_idx == len(dl) and inp.size[0] == len(ds) % batch_size)
      assert inp.size[1] == len(inp.items)
      dense = torch.zeros(tuple(inp.size), device='cuda')
      dense[inp.indices[0], inp.indices[1]] = inp.values
      assert int((dense > 0).sum()) == inp.values.numel()
      count += 1
    assert count == len(dl)

"""The C-ABI library must load on a CPU-only box and export every entry point include/recoder_b200.h declares
(no compute calls here: there is no GPU)."""
import ctypes
import os
import re

import pytest

from recoder_b200 import _native
from recoder_b200.csrc import build as csrc_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'recoder_b200.h')


def declared_symbols():
  text = open(HEADER).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(rcd_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
  csrc_build.build()
  return ctypes.CDLL(_native.LIB_PATH)


def test_header_declares_entry_points():
  syms = declared_symbols()
  assert len(syms) >= 25
  for must in ('rcd_collate', 'rcd_gather_rows', 'rcd_ae_encoder_fwd', 'rcd_decoder_fwd', 'rcd_decoder_fwd_loss', 'rcd_sddmm', 'rcd_loss_finish',
               'rcd_decoder_dgrad', 'rcd_decoder_wgrad', 'rcd_adam_step', 'rcd_sgd_step', 'rcd_host_stage_rows'):
    assert must in syms


def test_library_exports_every_declared_symbol(lib):
  for name in declared_symbols():
    assert hasattr(lib, name), 'librecoder_b200.so does not export %s' % name


def test_python_binding_covers_header():
  assert sorted(_native.EXPORTED_SYMBOLS) == declared_symbols()


def test_abi_version_and_error_string(lib):
  lib.rcd_abi_version.restype = ctypes.c_int
  assert lib.rcd_abi_version() == 1
  lib.rcd_last_error.restype = ctypes.c_char_p
  assert isinstance(lib.rcd_last_error(), bytes)


def test_host_staging_copies_rows(lib):
  """K0 is host code: it can be checked without a GPU against NumPy row gathering."""
  import numpy as np
  h = _native.load()
  rng = np.random.default_rng(0)
  U = 50
  lens = rng.integers(0, 9, U)
  indptr = np.zeros(U + 1, dtype=np.int64)
  np.cumsum(lens, out=indptr[1:])
  nnz = int(indptr[-1])
  indices = rng.integers(0, 1000, nnz).astype(np.int32)
  data = rng.random(nnz).astype(np.float32)
  users = rng.permutation(U)[:17].astype(np.int64)
  out_ptr = np.zeros(len(users) + 1, dtype=np.int64)
  out_idx = np.zeros(nnz, dtype=np.int32)
  out_val = np.zeros(nnz, dtype=np.float32)
  got = h.rcd_host_stage_rows(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, users.ctypes.data,
                              len(users), U, nnz, out_ptr.ctypes.data, out_idx.ctypes.data, out_val.ctypes.data)
  want_idx = np.concatenate([indices[indptr[u]:indptr[u + 1]] for u in users])
  want_val = np.concatenate([data[indptr[u]:indptr[u + 1]] for u in users])
  assert got == len(want_idx)
  assert np.array_equal(out_idx[:got], want_idx) and np.array_equal(out_val[:got], want_val)
  assert np.array_equal(np.diff(out_ptr), lens[users])
  # errors come back as negative status + message, never as an exception or a crash
  bad = np.array([U + 3], dtype=np.int64)
  rc = h.rcd_host_stage_rows(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, bad.ctypes.data, 1, U, nnz,
                             out_ptr.ctypes.data, out_idx.ctypes.data, out_val.ctypes.data)
  assert rc < 0 and b'out of range' in h.rcd_last_error()


def test_product_path_fails_loudly_without_cuda():
  import torch
  if torch.cuda.is_available():
    pytest.skip('CUDA present')
  with pytest.raises(RuntimeError):
    _native.require_cuda()
  from recoder_b200.model import Recoder
  from recoder_b200.nn import DynamicAutoencoder
  import numpy as np
  import scipy.sparse as sp
  from recoder_b200.data import RecommendationDataset
  m = sp.random(20, 30, density=0.2, format='csr', dtype=np.float32, random_state=0)
  m.data[:] = 1
  tr = Recoder(DynamicAutoencoder([8]), use_cuda=False, optimizer_type='adam', loss='mse')
  with pytest.raises(RuntimeError):
    tr.train(RecommendationDataset(m), batch_size=4)


def test_host_staging_large_pool_on_several_threads(lib):
  """Pools of 2048+ rows copy their rows on several threads (RCD_STAGE_THREADS): same bytes as NumPy row gathering."""
  import numpy as np
  h = _native.load()
  rng = np.random.default_rng(1)
  U = 20000
  lens = rng.integers(0, 30, U)
  indptr = np.zeros(U + 1, dtype=np.int64)
  np.cumsum(lens, out=indptr[1:])
  nnz = int(indptr[-1])
  indices = rng.integers(0, 1000, nnz).astype(np.int32)
  data = rng.random(nnz).astype(np.float32)
  users = rng.permutation(U)[:5000].astype(np.int64)
  want_i = np.concatenate([indices[indptr[u]:indptr[u + 1]] for u in users])
  want_d = np.concatenate([data[indptr[u]:indptr[u + 1]] for u in users])
  cap = len(want_i) + 7
  rp = np.full(len(users) + 1, -1, dtype=np.int64)
  oi = np.zeros(cap, dtype=np.int32)
  od = np.zeros(cap, dtype=np.float32)
  got = h.rcd_host_stage_rows(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, users.ctypes.data, len(users), U,
                              cap, rp.ctypes.data, oi.ctypes.data, od.ctypes.data)
  assert got == len(want_i)
  assert np.array_equal(oi[:got], want_i) and np.array_equal(od[:got], want_d)
  assert rp[0] == 0 and rp[-1] == got and np.array_equal(np.diff(rp), lens[users])
  # capacity and range errors are reported, not written past
  assert h.rcd_host_stage_rows(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, users.ctypes.data, len(users),
                               U, got - 1, rp.ctypes.data, oi.ctypes.data, od.ctypes.data) < 0
  bad = users.copy()
  bad[17] = U
  assert h.rcd_host_stage_rows(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, bad.ctypes.data, len(bad), U,
                               cap, rp.ctypes.data, oi.ctypes.data, od.ctypes.data) < 0


def test_adam_scalar_table_matches_the_dense_kernels_host_arithmetic():
  """`rcd_adam_scalars` (the per-step scalars the deferred Adam replays) forms lr/(1-b1^t) and 1/sqrt(1-b2^t) in double
  precision and rounds once to float — what `rcd_adam_step` does on the host for its own launch."""
  import math
  import numpy as np
  h = _native.load()
  out = np.zeros((50, 2), dtype=np.float32)
  assert h.rcd_adam_scalars(1e-3, 0.9, 0.999, 7, 50, out.ctypes.data) == 0
  for i in range(50):
    t = 7 + i
    assert out[i, 0] == np.float32(1e-3 / (1.0 - math.pow(0.9, t)))
    assert out[i, 1] == np.float32(1.0 / math.sqrt(1.0 - math.pow(0.999, t)))
  assert h.rcd_adam_scalars(1e-3, 0.9, 0.999, 0, 5, out.ctypes.data) < 0


def test_native_step_workspace_is_sized_by_capacities():
  """`rcd_step_workspace_bytes` is host arithmetic: the layout depends on the capacities alone (so it does not move from
  step to step) and grows with each of them."""
  import ctypes as C
  h = _native.load()
  assert h.rcd_step_args_size() == C.sizeof(_native.RcdStepArgs)

  def size(**kw):
    a = _native.RcdStepArgs()
    a.abi, a.kind, a.H, a.same_pool = _native.STEP_ABI, 0, 512, 1
    a.cap_rows, a.cap_n, a.cap_n_in, a.cap_nnz, a.cap_tnnz = 2048, 120000, 120000, 210000, 210000
    for k, v in kw.items():
      setattr(a, k, v)
    return int(h.rcd_step_workspace_bytes(C.byref(a)))

  base = size()
  assert 1.0e9 < base < 2.5e9          # C3: G 0.49 GB + slab 0.49 GB + Wg, partials, ...
  assert size(rows=17, row0=3) == base  # the actual step shape does not enter
  assert size(cap_n=130000, cap_n_in=130000) > base and size(cap_rows=4096) > base and size(cap_tnnz=400000) > base
  assert size(same_pool=0) > base
  assert size(abi=_native.STEP_ABI + 1) == 0 and size(H=0) == 0

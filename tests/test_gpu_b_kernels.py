"""Kernel-level parity: GEMM engines (tcgen05 vs SIMT vs fp32 torch reference on the same bf16 operands),
optimizer kernels vs torch.optim, gather / encoder kernels vs the oracle formulas."""
import numpy as np
import pytest
import torch

from recoder_b200 import _native
from recoder_b200._native import call, ptr

pytestmark = pytest.mark.gpu


def _gemm(mode, A, B, M, N, K, engine):
  C = torch.full((M, N), float('nan'), dtype=torch.float32, device='cuda')
  call('rcd_gemm_bf16', mode, ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, ptr(C), N, engine)
  torch.cuda.synchronize()
  return C


def _operands(mode, M, N, K, seed):
  g = torch.Generator(device='cuda').manual_seed(seed)
  def rnd(r, c):
    ld = (c + 7) // 8 * 8
    t = torch.zeros(r, ld, dtype=torch.bfloat16, device='cuda')
    t[:, :c] = torch.randn(r, c, generator=g, device='cuda').to(torch.bfloat16)
    return t[:, :c]
  if mode == 0:
    A, B = rnd(M, K), rnd(N, K)
    ref = A.float() @ B.float().t()
  elif mode == 1:
    A, B = rnd(M, K), rnd(K, N)
    ref = A.float() @ B.float()
  else:
    A, B = rnd(K, M), rnd(K, N)
    ref = A.float().t() @ B.float()
  return A, B, ref


SHAPES = [(128, 256, 64), (128, 256, 512), (256, 512, 128), (100, 300, 200), (500, 1000, 200), (37, 77, 24),
          (1024, 2048, 512), (130, 16, 1000)]


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_simt_matches_fp32(mode, M, N, K):
  A, B, ref = _operands(mode, M, N, K, seed=mode * 100 + M)
  C = _gemm(mode, A, B, M, N, K, _native.GEMM_SIMT)
  torch.testing.assert_close(C, ref, rtol=1e-4, atol=1e-3 * (K ** 0.5))


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_tcgen05_matches_fp32(mode, M, N, K):
  A, B, ref = _operands(mode, M, N, K, seed=mode * 100 + M)
  C = _gemm(mode, A, B, M, N, K, _native.GEMM_TCGEN05)
  assert not torch.isnan(C).any()
  torch.testing.assert_close(C, ref, rtol=1e-4, atol=1e-3 * (K ** 0.5))


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('M,N,K', SHAPES + [(2048, 512, 4096), (3001, 200, 333)])
def test_gemm_tcgen05_256row_tiles_match_fp32(mode, M, N, K):
  """256-row CTA tiles (two M=128 sub-tiles sharing the B stage, accumulators filling all 512 TMEM columns)."""
  A, B, ref = _operands(mode, M, N, K, seed=mode * 100 + M)
  C = _gemm(mode, A, B, M, N, K, _native.GEMM_TCGEN05 | (2 << 8))
  assert not torch.isnan(C).any()
  torch.testing.assert_close(C, ref, rtol=1e-4, atol=1e-3 * (K ** 0.5))


def test_adam_kernel_matches_torch():
  torch.manual_seed(0)
  I, H, n = 300, 32, 57
  p = torch.randn(I, H, device='cuda')
  ref = p.clone().requires_grad_(True)
  opt = torch.optim.Adam([{'params': ref, 'weight_decay': 1e-2}], lr=1e-2)
  m = torch.zeros_like(p); v = torch.zeros_like(p)
  for t in range(1, 6):
    ids = torch.randperm(I, device='cuda')[:n].sort().values
    g = torch.randn(n, H, device='cuda')
    pos = torch.full((I,), -1, dtype=torch.int32, device='cuda')
    pos[ids] = torch.arange(n, dtype=torch.int32, device='cuda')
    dense = torch.zeros(I, H, device='cuda'); dense[ids] = g
    ref.grad = dense
    opt.step()
    call('rcd_adam_step', ptr(p), ptr(m), ptr(v), I, H, ptr(g), H, ptr(pos), 1e-2, 0.9, 0.999, 1e-8, 1e-2, t)
  torch.testing.assert_close(p, ref.detach(), rtol=2e-6, atol=2e-7)


def test_sgd_and_sparse_adam_kernels_match_torch():
  torch.manual_seed(1)
  I, H, n = 200, 16, 33
  p = torch.randn(I, H, device='cuda'); ref = p.clone().requires_grad_(True)
  opt = torch.optim.SGD([{'params': ref, 'weight_decay': 1e-3}], lr=1e-2, momentum=0.9)
  buf = torch.zeros_like(p)
  p2 = torch.randn(I, H, device='cuda'); ref2 = p2.clone().requires_grad_(True)
  opt2 = torch.optim.SparseAdam([ref2], lr=1e-2)
  m2 = torch.zeros_like(p2); v2 = torch.zeros_like(p2)
  for t in range(1, 5):
    ids = torch.randperm(I, device='cuda')[:n].sort().values
    g = torch.randn(n, H, device='cuda')
    pos = torch.full((I,), -1, dtype=torch.int32, device='cuda')
    pos[ids] = torch.arange(n, dtype=torch.int32, device='cuda')
    dense = torch.zeros(I, H, device='cuda'); dense[ids] = g
    ref.grad = dense
    opt.step()
    call('rcd_sgd_step', ptr(p), ptr(buf), I, H, ptr(g), H, ptr(pos), 1e-2, 0.9, 1e-3)
    ref2.grad = torch.sparse_coo_tensor(ids.unsqueeze(0), g, (I, H)).coalesce()
    opt2.step()
    call('rcd_sparse_adam_step', ptr(p2), ptr(m2), ptr(v2), H, ptr(g), H, ptr(ids), n, 1e-2, 0.9, 0.999, 1e-8, t)
  torch.testing.assert_close(p, ref.detach(), rtol=2e-6, atol=2e-7)
  torch.testing.assert_close(p2, ref2.detach(), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize('H', [16, 200, 512, 30])
def test_gather_rows(H):
  torch.manual_seed(2)
  I, n = 1000, 333
  table = torch.randn(I, H, device='cuda')
  ids = torch.randperm(I, device='cuda')[:n].sort().values
  ld = (H + 7) // 8 * 8
  out = torch.full((n, ld), 7.0, dtype=torch.bfloat16, device='cuda')
  f32 = torch.empty(n, H, device='cuda')
  call('rcd_gather_rows', ptr(table), H, ptr(ids), n, _native.ACT_IDS['tanh'], ptr(out), ld, ptr(f32))
  ref = torch.tanh(table[ids])
  torch.testing.assert_close(f32, ref, rtol=1e-6, atol=1e-6)
  assert torch.equal(out[:, :H], ref.to(torch.bfloat16)) or \
      (out[:, :H].float() - ref).abs().max() < 1e-2
  assert (out[:, H:] == 0).all()


# ---------------------------------------------------------------------------------------------------------------
# fused decoder forward / loss epilogue, sparse loss side, CSC view, bias-gradient side product
# ---------------------------------------------------------------------------------------------------------------
def _bf16_mat(r, c, g, scale=1.0):
  ld = (c + 7) // 8 * 8
  t = torch.zeros(r, ld, dtype=torch.bfloat16, device='cuda')
  t[:, :c] = (torch.randn(r, c, generator=g, device='cuda') * scale).to(torch.bfloat16)
  return t, ld


def _sparse_targets(rows, n, per_row, g, ratings=False):
  """Random CSR targets with unique sorted columns per row (row 3 left empty)."""
  ptr_, cols, vals = [0], [], []
  for r in range(rows):
    k = 0 if r == 3 else int(torch.randint(1, per_row * 2, (1,), generator=g).item())
    c = torch.randperm(n, generator=g)[:min(k, n)].sort().values
    cols.append(c)
    vals.append(torch.randint(1, 6, (len(c),), generator=g).float() if ratings else torch.ones(len(c)))
    ptr_.append(ptr_[-1] + len(c))
  return (torch.tensor(ptr_, dtype=torch.int32, device='cuda'), torch.cat(cols).to(torch.int32).cuda(),
          torch.cat(vals).cuda())


@pytest.mark.parametrize('loss', ['mse', 'logloss', 'logistic'])
@pytest.mark.parametrize('rows,n,H', [(128, 256, 64), (333, 3001, 72), (1024, 20000, 512), (40, 100, 16)])
def test_decoder_fwd_loss_and_finish(loss, rows, n, H):
  gcpu = torch.Generator().manual_seed(rows + n)
  g = torch.Generator(device='cuda').manual_seed(rows * 7 + H)
  Zb, ldh = _bf16_mat(rows, H, g, 0.5)
  Wg, _ = _bf16_mat(n, H, g, 0.3)
  bias = torch.randn(n, generator=g, device='cuda') * 0.2
  row_ptr, cols, vals = _sparse_targets(rows, n, 12, gcpu, ratings=(loss == 'mse'))
  nnz = int(row_ptr[-1])
  conf = 2.0 if loss == 'mse' else 0.0
  inv_b = 1.0 / rows
  lid = _native.LOSS_IDS[loss]
  lib = _native.load()
  # reference on the same bf16 operands, fp32 math
  O = Zb[:, :H].float() @ Wg[:, :H].float().t() + bias
  T = torch.zeros(rows, n, device='cuda')
  rix = torch.repeat_interleave(torch.arange(rows, device='cuda'), (row_ptr[1:] - row_ptr[:-1]).long())
  T[rix, cols.long()] = vals
  if loss == 'mse':
    w = 1 + conf * (T > 0).float()
    ref_loss = (w * (O - T) ** 2).sum() * inv_b
    ref_dense = 2 * O * inv_b
    ref_full = 2 * w * (O - T) * inv_b
  elif loss == 'logloss':
    ref_loss = (-T * torch.log_softmax(O, dim=1)).sum() * inv_b
    ref_full = (torch.softmax(O, dim=1) * T.sum(1, keepdim=True) - T) * inv_b
    ref_dense = None
  else:
    ref_loss = torch.nn.functional.binary_cross_entropy_with_logits(O, T, reduction='sum') * inv_b
    ref_dense = torch.sigmoid(O) * inv_b
    ref_full = (torch.sigmoid(O) - T) * inv_b
  o_nnz = torch.empty(max(nnz, 1), device='cuda'); corr = torch.empty(max(nnz, 1), device='cuda')
  row_ref = torch.empty(rows, device='cuda')
  call('rcd_sddmm', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias), H, ptr(row_ptr), ptr(cols), ptr(vals), 0, rows, lid, conf,
       inv_b, ptr(o_nnz), ptr(corr), ptr(row_ref))
  torch.testing.assert_close(o_nnz[:nnz], O[rix, cols.long()], rtol=1e-4, atol=1e-4)
  ldn = (n + 7) // 8 * 8
  G = torch.full((rows, ldn), float('nan'), dtype=torch.bfloat16, device='cuda')
  sc = lib.rcd_decoder_stat_cols(n)
  stat = torch.full((rows, sc), float('nan'), device='cuda')
  call('rcd_decoder_fwd_loss', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias), rows, n, H, lid, inv_b,
       ptr(row_ref) if loss == 'logloss' else None, ptr(G), ldn, ptr(stat), sc, _native.DEC_MODE_LOSS, None)
  alpha = torch.empty(rows, device='cuda')
  Zf = Zb[:, :H].float().contiguous()
  Zs = torch.empty(rows, ldh, dtype=torch.bfloat16, device='cuda')
  acc = torch.zeros(1, dtype=torch.float64, device='cuda')
  bad = torch.zeros(1, dtype=torch.int32, device='cuda')
  row_sum = T.sum(1).contiguous()
  call('rcd_loss_finish', ptr(stat), sc, sc, rows, lid, conf, inv_b, ptr(row_ref), ptr(row_sum), ptr(row_ptr),
       ptr(vals), ptr(o_nnz), 0, ptr(alpha), ptr(Zf), H, ptr(Zs), ldh, ptr(acc), ptr(bad), 0, None, None, None, None)
  torch.cuda.synchronize()
  assert int(bad.item()) == 0
  assert not torch.isnan(G[:, :n].float()).any()
  assert float(acc.item()) == pytest.approx(float(ref_loss.item()), rel=2e-4)
  # full dL/dlogits = alpha * G + sparse part scattered at the stored targets
  full = G[:, :n].float() * (alpha[:, None] if loss == 'logloss' else 1.0)
  full[rix, cols.long()] += corr[:nnz]
  err = (full - ref_full).norm() / ref_full.norm()
  assert err < 4e-3, err   # bf16 storage of G: 2^-9 relative per element
  if ref_dense is not None:
    torch.testing.assert_close(G[:, :n].float(), ref_dense, rtol=1e-2, atol=1e-6)
  if loss == 'logloss':
    torch.testing.assert_close(Zs[:, :H].float(), (alpha[:, None] * Zf), rtol=1e-2, atol=1e-7)
    assert (Zs[:, H:] == 0).all()


@pytest.mark.parametrize('gap', [0.0, 30.0, 100.0, 400.0])
def test_nll_is_stable_for_any_logits(gap):
  """F.log_softmax (recoder/losses.py:69) is finite for any logits.  The fused epilogue takes the largest TARGET logit
  as softmax reference; rows in which a NON-target logit sits `gap` above every target (gap 100 and 400 overflow
  exp() against that reference) must be flagged, redone on the device with their true maximum, and match."""
  rows, n, H = 300, 1500, 64
  gcpu = torch.Generator().manual_seed(3)
  g = torch.Generator(device='cuda').manual_seed(4)
  Zb, ldh = _bf16_mat(rows, H, g, 0.5)
  Wg, _ = _bf16_mat(n, H, g, 0.3)
  bias = torch.randn(n, generator=g, device='cuda') * 0.2
  row_ptr, cols, vals = _sparse_targets(rows, n, 12, gcpu, ratings=False)
  nnz = int(row_ptr[-1])
  rix = torch.repeat_interleave(torch.arange(rows, device='cuda'), (row_ptr[1:] - row_ptr[:-1]).long())
  T = torch.zeros(rows, n, device='cuda')
  T[rix, cols.long()] = vals
  # adversarial columns: items nobody in the slice interacted with get a huge decoder bias
  free = (T.sum(0) == 0).nonzero().flatten()
  assert free.numel() >= 3
  if gap > 0:
    bias[free[0]] += gap
    bias[free[1]] += gap * 0.5
  inv_b = 1.0 / rows
  lid = _native.LOSS_IDS['logloss']
  lib = _native.load()
  O = Zb[:, :H].float() @ Wg[:, :H].float().t() + bias
  ref_loss = (-T.double() * torch.log_softmax(O.double(), dim=1)).sum() * inv_b
  ref_full = (torch.softmax(O.double(), dim=1) * T.sum(1, keepdim=True).double() - T.double()) * inv_b
  assert torch.isfinite(ref_loss)
  o_nnz = torch.empty(nnz, device='cuda'); corr = torch.empty(nnz, device='cuda')
  row_ref = torch.empty(rows, device='cuda')
  call('rcd_sddmm', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias), H, ptr(row_ptr), ptr(cols), ptr(vals), 0, rows, lid, 0.0,
       inv_b, ptr(o_nnz), ptr(corr), ptr(row_ref))
  ldn = (n + 7) // 8 * 8
  G = torch.zeros((rows, ldn), dtype=torch.bfloat16, device='cuda')
  sc = lib.rcd_decoder_stat_cols(n)
  stat = torch.zeros((rows, sc), device='cuda')
  alpha = torch.empty(rows, device='cuda')
  Zf = Zb[:, :H].float().contiguous()
  acc = torch.zeros(1, dtype=torch.float64, device='cuda')
  bad = torch.zeros(1, dtype=torch.int32, device='cuda')
  flag = torch.zeros(1, dtype=torch.int32, device='cuda')
  row_redo = torch.zeros(rows, dtype=torch.int32, device='cuda')
  nb = lib.rcd_loss_finish_blocks(rows)
  blocks = torch.zeros(nb, dtype=torch.float64, device='cuda')
  row_sum = T.sum(1).contiguous()

  def fused(mode, cond):
    call('rcd_decoder_fwd_loss', ptr(Zb), ldh, ptr(Wg), ldh, ptr(bias), rows, n, H, lid, inv_b, ptr(row_ref), ptr(G),
         ldn, ptr(stat), sc, mode, ptr(cond))

  def finish(rf, rr, cond):
    call('rcd_loss_finish', ptr(stat), sc, sc, rows, lid, 0.0, inv_b, ptr(row_ref), ptr(row_sum), ptr(row_ptr),
         ptr(vals), ptr(o_nnz), 0, ptr(alpha), ptr(Zf), H, None, ldh, None, ptr(bad), 0, ptr(blocks), ptr(rf), ptr(rr),
         ptr(cond))

  fused(_native.DEC_MODE_LOSS, None)
  finish(flag, row_redo, None)
  torch.cuda.synchronize()
  flagged = int(flag.item())
  assert flagged == (1 if gap >= 100 else 0)      # 2^64 = e^44: gaps of 0 and 30 stay below the clamp
  assert torch.isfinite(G.float()).all()          # the clamp keeps the first pass finite whatever the logits are
  fused(_native.DEC_MODE_ROWMAX, flag)
  call('rcd_nll_ref_fix', ptr(stat), sc, sc, rows, ptr(row_redo), ptr(row_ref), ptr(flag))
  fused(_native.DEC_MODE_LOSS, flag)
  finish(None, None, flag)
  call('rcd_loss_sum', ptr(blocks), nb, ptr(acc), ptr(flag))
  torch.cuda.synchronize()
  assert int(bad.item()) == 0 and int(flag.item()) == 0
  if flagged:
    assert int(row_redo.sum().item()) == rows
    torch.testing.assert_close(row_ref, O.max(dim=1).values, rtol=1e-3, atol=1e-2)
  assert float(acc.item()) == pytest.approx(float(ref_loss.item()), rel=2e-4)
  full = G[:, :n].float() * alpha[:, None]
  full[rix, cols.long()] += corr
  err = (full.double() - ref_full).norm() / ref_full.norm()
  assert err < 4e-3, err


@pytest.mark.parametrize('rows,n,per_row', [(64, 50, 10), (2048, 300, 40), (500, 4000, 30)])
def test_slice_csc_matches_scipy(rows, n, per_row):
  import scipy.sparse as sp
  g = torch.Generator().manual_seed(rows)
  row_ptr, cols, vals = _sparse_targets(rows, n, per_row, g, ratings=True)
  nnz = int(row_ptr[-1])
  lib = _native.load()
  csc_ptr = torch.empty(n + 1, dtype=torch.int32, device='cuda')
  csc_row = torch.empty(nnz, dtype=torch.int32, device='cuda')
  csc_val = torch.empty(nnz, device='cuda')
  csc_src = torch.empty(nnz, dtype=torch.int32, device='cuda')
  sb = lib.rcd_slice_csc_scratch_bytes(n, nnz)
  scratch = torch.empty(sb, dtype=torch.uint8, device='cuda')
  call('rcd_slice_csc', ptr(row_ptr), ptr(cols), ptr(vals), 0, rows, n, ptr(csc_ptr), ptr(csc_row), ptr(csc_val),
       ptr(csc_src), ptr(scratch), sb)
  m = sp.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), row_ptr.cpu().numpy()), shape=(rows, n)).tocsc()
  m.sort_indices()
  assert np.array_equal(csc_ptr.cpu().numpy(), m.indptr)
  assert np.array_equal(csc_row.cpu().numpy(), m.indices)
  assert np.array_equal(csc_val.cpu().numpy(), m.data)
  src = csc_src.cpu().numpy()
  assert np.array_equal(vals.cpu().numpy()[src], m.data) and np.array_equal(cols.cpu().numpy()[src],
                                                                           np.repeat(np.arange(n), np.diff(m.indptr)))


@pytest.mark.parametrize('engine', [_native.GEMM_SIMT, _native.GEMM_TCGEN05])
@pytest.mark.parametrize('rows,n,H,weighted', [(256, 1000, 64, True), (300, 777, 200, False), (2048, 5000, 512, True)])
def test_wgrad_bias_side_product(engine, rows, n, H, weighted):
  g = torch.Generator(device='cuda').manual_seed(n)
  G, ldn = _bf16_mat(rows, n, g, 0.1)
  Zs, ldh = _bf16_mat(rows, H, g, 0.5)
  w = torch.rand(rows, generator=g, device='cuda') if weighted else None
  dW = torch.full((n, H), float('nan'), device='cuda')
  db = torch.full((n,), float('nan'), device='cuda')
  call('rcd_decoder_wgrad', ptr(G), ldn, ptr(Zs), ldh, rows, n, H, ptr(dW), H, ptr(w), ptr(db), engine)
  torch.cuda.synchronize()
  ref_dW = G[:, :n].float().t() @ Zs[:, :H].float()
  ref_db = (G[:, :n].float() * (w[:, None] if weighted else 1.0)).sum(0)
  torch.testing.assert_close(dW, ref_dW, rtol=1e-4, atol=1e-3 * rows ** 0.5)
  torch.testing.assert_close(db, ref_db, rtol=1e-4, atol=1e-4)

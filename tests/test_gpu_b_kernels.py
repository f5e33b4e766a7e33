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

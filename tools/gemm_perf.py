"""GEMM throughput probe at the C3 step shapes: the tcgen05 engine (128-row and 256-row CTA tiles) against cuBLAS
(torch.matmul, bf16 -> fp32 accumulate, bf16 out) on the same operands, each timed as `iters` back-to-back launches
with CUDA events (burst) — plus the same after a 2 s tensor-core soak (sustained, power-capped clocks)."""
import sys
import time

import torch

sys.path.insert(0, '.')
from recoder_b200 import _native  # noqa: E402
from recoder_b200._native import call, ptr  # noqa: E402

ROWS, N_ITEMS, H = 2048, 112896, 512


def timed(fn, iters):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters


def soak(seconds):
  x = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
  t0 = time.time()
  while time.time() - t0 < seconds:
    for _ in range(20):
      x @ x
    torch.cuda.synchronize()


def main():
  g = torch.Generator(device='cuda').manual_seed(0)
  rnd = lambda r, c: torch.randn(r, c, device='cuda', generator=g).to(torch.bfloat16)  # noqa: E731
  Zb, Wg, G = rnd(ROWS, H), rnd(N_ITEMS, H), rnd(ROWS, N_ITEMS)
  flops = 2.0 * ROWS * N_ITEMS * H
  shapes = {
    # name: (mode, A, B, M, N, K, torch reference)
    'fwd   O=Zb.Wg^T  [2048 x 113K x 512]': (0, Zb, Wg, ROWS, N_ITEMS, H, lambda: Zb @ Wg.t()),
    'dgrad dZ=G.Wg    [2048 x 512 x 113K]': (1, G, Wg, ROWS, H, N_ITEMS, lambda: G @ Wg),
    'wgrad dW=G^T.Zb  [113K x 512 x 2048]': (2, G, Zb, N_ITEMS, H, ROWS, lambda: G.t() @ Zb),
  }
  for phase in ('burst', 'sustained'):
    if phase == 'sustained':
      soak(2.0)
    for name, (mode, A, B, M, N, K, ref) in shapes.items():
      C = torch.empty(M, N, device='cuda', dtype=torch.float32)
      res = {}
      for tag, eng in (('tc128', _native.GEMM_TCGEN05), ('tc256', _native.GEMM_TCGEN05 | (2 << 8))):
        ms = timed(lambda: call('rcd_gemm_bf16', mode, ptr(A), A.stride(0), ptr(B), B.stride(0), M, N, K, ptr(C), N,
                                eng), 20)
        res[tag] = flops / ms / 1e9
      ms = timed(ref, 20)
      res['cublas'] = flops / ms / 1e9
      print('%-9s %s  ' % (phase, name) + '  '.join('%s %.0f TF/s' % kv for kv in res.items()), flush=True)


if __name__ == '__main__':
  main()

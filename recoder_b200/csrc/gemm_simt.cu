// SIMT validation engine: the same tiles, split-K ranges and row epilogue as the tcgen05 engine, with the
// accumulator produced by plain fp32 FMAs over the same bf16 operands.  It exists so that the tests can tell a
// tensor-core / TMA / descriptor fault from a fault anywhere else in the step; it is never the default engine.
#include "gemm_internal.cuh"

namespace rcd {

__device__ __forceinline__ float bf16_to_f(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }

static __global__ void __launch_bounds__(kTileM)
    k_gemm_simt(GemmProblem g, EpiParams e, int m_tiles, int n_tiles, int kblocks) {
  const int units = m_tiles * n_tiles * g.splits;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const UnitCoord u = decode_unit(unit, m_tiles, n_tiles, g.splits, kblocks, g.n_fastest, g.split_major);
    const int row = u.mt * kTileM + threadIdx.x;
    const int k0 = u.kb0 * kTileK, k1 = min(u.kb1 * kTileK, g.K);
    RowEpilogue epi;
    epi.begin();
    for (int cb = 0; cb < g.bn; cb += 32) {
      const int col_base = u.nt * g.bn + cb;
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      if (col_base < g.N) {
        for (int k = k0; k < k1; ++k) {
          float a = 0.f;
          if (row < g.M) a = bf16_to_f(g.mode == 2 ? g.A[(size_t)k * g.lda + row] : g.A[(size_t)row * g.lda + k]);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = col_base + i;
            if (c < g.N) {
              float b = bf16_to_f(g.mode == 0 ? g.B[(size_t)c * g.ldb + k] : g.B[(size_t)k * g.ldb + c]);
              acc[i] = fmaf(a, b, acc[i]);
            }
          }
        }
      }
      epi.chunk32(e, row, col_base, u.split, acc);
    }
    epi.end(e, row, u.nt);
  }
}

int gemm_simt_launch(const GemmProblem& g, const EpiParams& e, cudaStream_t st) {
  const int m_tiles = rcd_div_up(g.M, kTileM), n_tiles = rcd_div_up(g.N, g.bn);
  const int kblocks = rcd_div_up(g.K, kTileK);
  const int units = m_tiles * n_tiles * g.splits;
  int grid = units < 148 * 8 ? units : 148 * 8;
  k_gemm_simt<<<grid, kTileM, 0, st>>>(g, e, m_tiles, n_tiles, kblocks);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

}  // namespace rcd

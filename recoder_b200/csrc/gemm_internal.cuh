// Internal GEMM interface shared by the tcgen05 engine (gemm_tc.cu), the SIMT validation engine
// (gemm_simt.cu) and the public entry points (gemm.cu).
//
// Tile semantics common to both engines: a work unit is (m-tile of 128 rows, n-tile of `bn` columns, k-split);
// inside a unit, thread "lane-of-128" owns one output row and receives the accumulator 32 columns at a time —
// exactly how tcgen05.ld.32x32b hands TMEM to a warp (TMEM lane == row).  The row-wise epilogue below is
// therefore written once and used by both engines.
#pragma once
#include "common.cuh"

struct CUtensorMap_st;

namespace rcd {

constexpr int kTileM = 128;
constexpr int kTileK = 64;     // one 128-byte swizzle atom of bf16
constexpr int kTileNMax = 256;
constexpr int kDecoderTileN = 256;

enum { EPI_F32 = 0, EPI_DECODER = 1 };

// modes (see rcd_gemm_bf16): 0 = A K-major, B K-major; 1 = A K-major, B MN-major; 2 = both MN-major
struct GemmProblem {
  int mode;
  const uint16_t* A; int lda;
  const uint16_t* B; int ldb;
  int M, N, K;
  int bn;        // n-tile width (multiple of 16, <= 256)
  int splits;    // k-splits (each non-empty)
  int n_fastest; // tile order
  int m_sub;     // tcgen05 engine: 128-row sub-tiles per CTA (2 = 256-row tiles sharing the B stage; 0/1 = one)
  int split_major; // 1: unit = split * tiles + tile — CTAs in flight work on the SAME k-range of different tiles, so
                   // the A/B k-slabs they stream are shared through L2 (split-K over a K that does not fit L2)
};

struct EpiParams {
  int kind;
  int M, N;
  // EPI_F32: C[split][row][col]
  float* C; int ldc; long long split_stride;
  // EPI_DECODER
  const float* bias;
  uint16_t* Obf; float* Of32; int ldo;
  float* stat_max; float* stat_sum;  // [n_tiles, M]
  // side product of mode 2 (A MN-major): colsum[m] = sum_k colw[k] * A[k, m]  (colw == nullptr: weights 1)
  const float* colw; float* colsum;
};

struct RowEpilogue {
  float m_run, s_run;
  __device__ __forceinline__ void begin() {
    m_run = -INFINITY;
    s_run = 0.f;
  }

  // v: accumulator values of columns [col_base, col_base+32) of row `row`
  __device__ __forceinline__ void chunk32(const EpiParams& p, int row, int col_base, int split, float (&v)[32]) {
    if (col_base >= p.N) return;
    const bool row_ok = row < p.M;
    if (p.kind == EPI_F32) {
      if (!row_ok) return;
      float* dst = p.C + (size_t)split * p.split_stride + (size_t)row * p.ldc + col_base;
      const bool vec_ok = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                          ((p.split_stride & 3) == 0);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const int c = col_base + i;
        if (vec_ok && c + 3 < p.N) {
          *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c + k < p.N) dst[i + k] = v[i + k];
        }
      }
      return;
    }
    // EPI_DECODER: logits = acc + bias, rounded to bf16 (the value every later kernel sees)
    uint32_t packed[16];
    float cmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int c = col_base + i;
      float o0 = (c < p.N) ? v[i] + __ldg(p.bias + c) : 0.f;
      float o1 = (c + 1 < p.N) ? v[i + 1] + __ldg(p.bias + c + 1) : 0.f;
      v[i] = o0;
      v[i + 1] = o1;
      packed[i >> 1] = pack_bf16x2(o0, o1);
      if (c < p.N) cmax = fmaxf(cmax, bf16_lo(packed[i >> 1]));
      if (c + 1 < p.N) cmax = fmaxf(cmax, bf16_hi(packed[i >> 1]));
    }
    if (p.stat_max) {
      const float new_m = fmaxf(m_run, cmax);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int c = col_base + i;
        if (c < p.N) s += __expf(bf16_lo(packed[i >> 1]) - new_m);
        if (c + 1 < p.N) s += __expf(bf16_hi(packed[i >> 1]) - new_m);
      }
      s_run = s_run * __expf(m_run - new_m) + s;  // m_run == -inf on the first chunk: exp(-inf) = 0
      m_run = new_m;
    }
    if (!row_ok) return;
    if (p.Obf) {
      uint16_t* dst = p.Obf + (size_t)row * p.ldo + col_base;
      const int n8 = (p.N + 7) & ~7;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col_base + g * 8 < n8)
          *reinterpret_cast<uint4*>(dst + g * 8) =
              make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
      }
    }
    if (p.Of32) {
      float* dst = p.Of32 + (size_t)row * p.ldo + col_base;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col_base + i < p.N) dst[i] = v[i];
    }
  }

  __device__ __forceinline__ void end(const EpiParams& p, int row, int n_tile) {
    if (p.kind == EPI_DECODER && p.stat_max && row < p.M) {
      p.stat_max[(size_t)n_tile * p.M + row] = m_run;
      p.stat_sum[(size_t)n_tile * p.M + row] = s_run;
    }
  }
};

// unit -> (m_tile, n_tile, split, k-block range); shared by both engines
struct UnitCoord {
  int mt, nt, split, kb0, kb1;
};
__device__ __forceinline__ UnitCoord decode_unit(int unit, int m_tiles, int n_tiles, int splits, int kblocks,
                                                 int n_fastest, int split_major = 0) {
  UnitCoord u;
  const int tiles = m_tiles * n_tiles;
  const int tile = split_major ? unit % tiles : unit / splits;
  u.split = split_major ? unit / tiles : unit % splits;
  if (n_fastest) {
    u.nt = tile % n_tiles;
    u.mt = tile / n_tiles;
  } else {
    u.mt = tile % m_tiles;
    u.nt = tile / m_tiles;
  }
  const int per = (kblocks + splits - 1) / splits;
  u.kb0 = u.split * per;
  u.kb1 = min(u.kb0 + per, kblocks);
  return u;
}

int gemm_tc_launch(const GemmProblem& g, const EpiParams& e, cudaStream_t st);
// TMA descriptor of a 2D bf16 tensor [outer, inner] (row-major, leading dimension ld elements), 128B swizzle
int encode_map(::CUtensorMap_st* map, const void* base, long long inner, long long outer, long long ld, int box_inner,
               int box_outer);
int gemm_simt_launch(const GemmProblem& g, const EpiParams& e, cudaStream_t st);

}  // namespace rcd

// Inner dense layers, dropout and small elementwise pieces of the generalised model:
//   rcd_sgemm        : fp32 C = op(A) * op(B) (+ bias[col]) -> act            inner `nn.Linear` layers of a
//                      multi-layer DynamicAutoencoder (recoder/nn.py:242-243, 248-249) and their backward `mm`s.
//                      These layers are [B, h_i] x [h_i, h_j] with h <= a few thousand: ~1e9 flops per step against
//                      ~1e12 in the embedding GEMMs, so a plain SIMT fp32 kernel (exact fp32, like the reference)
//                      is the right tool; the tcgen05 engine stays reserved for the item-axis GEMMs.
//   rcd_dropout      : y = x * keep / (1 - p), keep ~ Bernoulli(1 - p) from Philox4x32-10 keyed by (seed, stream) and
//                      counted by the element's GLOBAL index, or from an explicit keep mask (tests).  nn.Dropout at
//                      recoder/nn.py:236-237 (input noise: applied to the stored non-zeros only — zeros stay zero),
//                      nn.py:245-246 (bottleneck) and nn.py:351-352 (MF user embedding).  The same call applied to
//                      the gradient is the backward.
//   rcd_act_grad     : dpre = dy * act'(y) (derivative through the activation OUTPUT)
//   rcd_colsum       : out[h] = sum_r x[r,h] (bias gradients)
//   rcd_f32_to_bf16_rows : fp32 [rows,H] -> bf16 [rows,ld] zero padded (decoder GEMM operand)
#include "common.cuh"

namespace rcd {

// ---- Philox4x32-10 (Salmon et al., SC'11) ------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// one uniform in [0,1) per element index (4 elements share one Philox block)
__device__ __forceinline__ float philox_uniform(unsigned long long seed, uint32_t stream, unsigned long long idx) {
  const unsigned long long blk = idx >> 2;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), stream, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t lane = (uint32_t)idx & 3u;
  const uint32_t v = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  return (float)(v >> 8) * (1.0f / 16777216.0f);
}

static __global__ void k_dropout(const float* __restrict__ x, long long count, float p, unsigned long long seed,
                                 uint32_t stream, long long index_base, const uint8_t* __restrict__ keep_mask,
                                 float* __restrict__ y) {
  const float scale = 1.0f / (1.0f - p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const bool keep = keep_mask ? (keep_mask[i] != 0) : (philox_uniform(seed, stream, (unsigned long long)(index_base + i)) >= p);
    y[i] = keep ? x[i] * scale : 0.f;
  }
}

static __global__ void k_act_grad(const float* __restrict__ dy, const float* __restrict__ y, long long count, int act,
                                  float* __restrict__ dpre) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    dpre[i] = dy[i] * act_grad_from_out(y[i], act);
}

// 32 columns per block, 32 warps stride the rows (tall inputs: the item-parallel mode sums over the GLOBAL batch), four
// independent row loads in flight per warp, fixed-order reduction through shared memory
constexpr int kColsumWarps = 32;
static __global__ void __launch_bounds__(kColsumWarps * 32)
    k_colsum(const float* __restrict__ x, int rows, int H, int ld, float* __restrict__ out) {
  __shared__ float part[kColsumWarps][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int h = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (h < H) {
    int r = w;
    for (; r + 3 * kColsumWarps < rows; r += 4 * kColsumWarps) {
      s0 += x[(size_t)r * ld + h];
      s1 += x[(size_t)(r + kColsumWarps) * ld + h];
      s2 += x[(size_t)(r + 2 * kColsumWarps) * ld + h];
      s3 += x[(size_t)(r + 3 * kColsumWarps) * ld + h];
    }
    for (; r < rows; r += kColsumWarps) s0 += x[(size_t)r * ld + h];
  }
  part[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && h < H) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kColsumWarps; ++k) t += part[k][lane];  // fixed order
    out[h] = t;
  }
}

static __global__ void k_f32_to_bf16_rows(const float* __restrict__ x, int rows, int H, uint16_t* __restrict__ out,
                                          int ld) {
  const long long total = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), h = (int)(i % ld);
    const float v = (h < H) ? x[(size_t)r * H + h] : 0.f;
    reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  }
}

// ---- fp32 SIMT GEMM: 64x64 tile, K step 16, 256 threads, 4x4 outputs per thread -------------------------------------
// A(m,k) = TA ? A[k*lda + m] : A[m*lda + k];  B(k,n) = TB ? B[n*ldb + k] : B[k*ldb + n]
template <bool TA, bool TB>
static __global__ void __launch_bounds__(256)
    k_sgemm(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
            int M, int N, int K, const float* __restrict__ bias, int act, int accumulate) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int e = threadIdx.x + t * 256;  // 0..1023
      {
        // A tile element (mm, kk): choose the mapping whose fastest index is contiguous in memory
        const int kk = TA ? e / 64 : e % 16, mm = TA ? e % 64 : e / 16;
        const int m = m0 + mm, k = k0 + kk;
        float v = 0.f;
        if (m < M && k < K) v = TA ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k];
        As[kk][mm] = v;
      }
      {
        const int kk = TB ? e % 16 : e / 64, nn = TB ? e / 16 : e % 64;
        const int n = n0 + nn, k = k0 + kk;
        float v = 0.f;
        if (n < N && k < K) v = TB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      v = act_apply(v, act);
      if (accumulate) v += C[(size_t)m * ldc + n];
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// z = act(x + bias[h]) -> fp32 Z and (optional) bf16 copy [rows, ld] zero padded: finishes the encoder after the
// partial sums of the item shards have been all-reduced (item-parallel mode)
static __global__ void k_bias_act(const float* __restrict__ x, const float* __restrict__ bias, int rows, int H, int act,
                                  float* __restrict__ Z, uint16_t* __restrict__ Zb, int ld) {
  const long long total = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), h = (int)(i % ld);
    float z = 0.f;
    if (h < H) {
      z = act_apply(x[(size_t)r * H + h] + bias[h], act);
      Z[(size_t)r * H + h] = z;
    }
    if (Zb) reinterpret_cast<__nv_bfloat16*>(Zb)[i] = __float2bfloat16_rn(z);
  }
}

// out[r] = sum_c x[r, c] (one warp per row, fixed order)
static __global__ void k_rowsum(const float* __restrict__ x, int rows, int cols, int ld, float* __restrict__ out) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += x[(size_t)r * ld + c];
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

static inline int ew_grid(long long count) {
  long long b = (count + 255) / 256;
  long long cap = (long long)rcd_num_sms() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_sgemm(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B,
                         int ldb, float* C, int ldc, const float* bias, int act, int accumulate, void* stream) {
  RCD_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, "bad arguments");
  RCD_CHECK_ARG(lda >= (trans_a ? M : K) && ldb >= (trans_b ? K : N), "bad leading dimension");
  dim3 grid(rcd_div_up(N, 64), rcd_div_up(M, 64));
  cudaStream_t st = (cudaStream_t)stream;
  if (!trans_a && !trans_b) k_sgemm<false, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, act, accumulate);
  else if (!trans_a && trans_b) k_sgemm<false, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, act, accumulate);
  else if (trans_a && !trans_b) k_sgemm<true, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, act, accumulate);
  else k_sgemm<true, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, act, accumulate);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_dropout(const float* x, long long count, float p, unsigned long long seed, unsigned int rng_stream,
                           long long index_base, const uint8_t* keep_mask, float* y, void* stream) {
  RCD_CHECK_ARG(x && y && count > 0 && p >= 0.f && p < 1.f && index_base >= 0, "bad arguments");
  k_dropout<<<ew_grid(count), 256, 0, (cudaStream_t)stream>>>(x, count, p, seed, rng_stream, index_base, keep_mask, y);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_act_grad(const float* dy, const float* y, long long count, int act, float* dpre, void* stream) {
  RCD_CHECK_ARG(dy && y && dpre && count > 0, "bad arguments");
  k_act_grad<<<ew_grid(count), 256, 0, (cudaStream_t)stream>>>(dy, y, count, act, dpre);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_colsum(const float* x, int rows, int H, int ld, float* out, void* stream) {
  RCD_CHECK_ARG(x && out && rows > 0 && H > 0 && ld >= H, "bad arguments");
  k_colsum<<<rcd_div_up(H, 32), kColsumWarps * 32, 0, (cudaStream_t)stream>>>(x, rows, H, ld, out);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_f32_to_bf16_rows(const float* x, int rows, int H, uint16_t* out, int ld, void* stream) {
  RCD_CHECK_ARG(x && out && rows > 0 && H > 0 && ld >= H, "bad arguments");
  k_f32_to_bf16_rows<<<ew_grid((long long)rows * ld), 256, 0, (cudaStream_t)stream>>>(x, rows, H, out, ld);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_bias_act(const float* x, const float* bias, int rows, int H, int act, float* Z, uint16_t* Zb, int ld,
                            void* stream) {
  RCD_CHECK_ARG(x && bias && Z && rows > 0 && H > 0 && (!Zb || ld >= H), "bad arguments");
  const int ldd = Zb ? ld : H;
  k_bias_act<<<ew_grid((long long)rows * ldd), 256, 0, (cudaStream_t)stream>>>(x, bias, rows, H, act, Z, Zb, ldd);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_rowsum(const float* x, int rows, int cols, int ld, float* out, void* stream) {
  RCD_CHECK_ARG(x && out && rows > 0 && cols > 0 && ld >= cols, "bad arguments");
  k_rowsum<<<rcd_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, out);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

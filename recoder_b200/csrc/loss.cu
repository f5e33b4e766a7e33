// K5: loss and dL/dlogits from the bf16 logits written by the decoder GEMM.
//
// rcd_loss_grad is a column-strip streaming kernel (HBM-bound: reads B*n*2 B of logits, writes B*n*2 B of
// dlogits).  A block owns 64 columns (one 128-byte line of bf16 per row) and sweeps all rows:
//   phase 1: dense formula with target 0 for every element, fp32 column sums and loss partial in registers;
//   phase 2: for the strip's non-zero targets (contiguous range of the slice CSC) the exact-minus-dense
//            correction is computed in fp32, one warp per column, and written to csc_corr; column-sum and loss
//            deltas are reduced in a fixed order (deterministic results).  dO itself keeps the dense value: the
//            large, clustered gradient entries at the non-zeros would otherwise be quantised to bf16 with a
//            systematic bias; they are applied in fp32 by the sparse kernels (rcd_sparse_dgrad,
//            rcd_csc_rows_accumulate), i.e. dL/dO = bf16 dense part + fp32 sparse part.
// This replaces recoder/losses.py:43-47,68-71, BCEWithLogitsLoss (recoder/model.py:91), the /B at
// recoder/model.py:483-484 and autograd's backward through them; the sparse target never gets densified
// (recoder/model.py:457-458,473-476 materialise it as a dense [B,n] fp32 matrix).
#include "common.cuh"

namespace rcd {

constexpr int kStripCols = 64;
constexpr int kStripWarps = 8;

__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// dense-part gradient (target = 0) and loss term
template <int LOSS>
__device__ __forceinline__ void dense_term(float o, float inv_b, float lse, float rsum, float& d, float& l) {
  if (LOSS == RCD_LOSS_MSE) {
    d = 2.0f * o * inv_b;
    l = o * o;
  } else if (LOSS == RCD_LOSS_NLL) {
    d = expf(o - lse) * rsum * inv_b;
    l = 0.f;
  } else {
    d = sigmoid_f(o) * inv_b;
    l = softplus_f(o);
  }
}
// exact gradient and loss term at a stored target t
template <int LOSS>
__device__ __forceinline__ void exact_term(float o, float t, float conf, float inv_b, float lse, float rsum, float& d,
                                           float& l) {
  if (LOSS == RCD_LOSS_MSE) {
    float w = 1.0f + (t > 0.f ? conf : 0.f);  // losses.py:44
    d = 2.0f * w * (o - t) * inv_b;
    l = w * (o - t) * (o - t);
  } else if (LOSS == RCD_LOSS_NLL) {
    d = (expf(o - lse) * rsum - t) * inv_b;
    l = -t * o;  // the +t*lse part is added per row by rcd_softmax_lse
  } else {
    d = (sigmoid_f(o) - t) * inv_b;
    l = softplus_f(o) - t * o;
  }
}

template <int LOSS>
static __global__ void __launch_bounds__(kStripWarps * 32)
    k_loss_grad(const uint16_t* __restrict__ O, int ldo, int rows, int n, float conf, float inv_b,
                const float* __restrict__ lse, const float* __restrict__ row_sum, const int32_t* __restrict__ csc_ptr,
                const int32_t* __restrict__ csc_row, const float* __restrict__ csc_val, uint16_t* __restrict__ dO,
                int lddo, float* __restrict__ csc_corr, float* __restrict__ db, double* __restrict__ loss_acc) {
  __shared__ float s_col[kStripWarps][kStripCols];
  __shared__ float s_loss[kStripWarps];
  __shared__ float s_delta[kStripCols];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * kStripCols;
  const int c = c0 + 2 * lane;  // this lane's column pair
  const bool ok0 = c < n, ok1 = c + 1 < n;
  float cs0 = 0.f, cs1 = 0.f, lsum = 0.f;
  if (ok0) {
    for (int r = w; r < rows; r += kStripWarps) {
      const uint32_t packed = *reinterpret_cast<const uint32_t*>(O + (size_t)r * ldo + c);
      float l_r = 0.f, rs_r = 0.f;
      if (LOSS == RCD_LOSS_NLL) {
        l_r = lse[r];
        rs_r = row_sum[r];
      }
      float d0, d1, l0, l1;
      dense_term<LOSS>(bf16_lo(packed), inv_b, l_r, rs_r, d0, l0);
      dense_term<LOSS>(bf16_hi(packed), inv_b, l_r, rs_r, d1, l1);
      if (!ok1) {
        d1 = 0.f;
        l1 = 0.f;
      }
      cs0 += d0;
      cs1 += d1;
      lsum += l0 + l1;
      *reinterpret_cast<uint32_t*>(dO + (size_t)r * lddo + c) = pack_bf16x2(d0, d1);
    }
  }
  s_col[w][2 * lane] = cs0;
  s_col[w][2 * lane + 1] = cs1;
  lsum = warp_sum(lsum);
  if (lane == 0) s_loss[w] = lsum;
  __syncthreads();

  // phase 2: exact values at the stored targets; warp w owns columns w, w+8, ... of the strip
  float ldelta = 0.f;
  for (int j = w; j < kStripCols; j += kStripWarps) {
    const int cc = c0 + j;
    float dsum = 0.f;
    if (cc < n) {
      const int s = csc_ptr[cc], e = csc_ptr[cc + 1];
      for (int p = s + lane; p < e; p += 32) {
        const int r = csc_row[p];
        const float t = csc_val[p];
        const float o = __uint_as_float((uint32_t)O[(size_t)r * ldo + cc] << 16);
        float l_r = 0.f, rs_r = 0.f;
        if (LOSS == RCD_LOSS_NLL) {
          l_r = lse[r];
          rs_r = row_sum[r];
        }
        float d_dense, l_dense, d, l;
        dense_term<LOSS>(o, inv_b, l_r, rs_r, d_dense, l_dense);
        exact_term<LOSS>(o, t, conf, inv_b, l_r, rs_r, d, l);
        const float corr = sparse_corr(LOSS, o, t, conf, inv_b);  // == d - d_dense in exact arithmetic
        csc_corr[p] = corr;
        dsum += corr;
        ldelta += l - l_dense;
      }
    }
    dsum = warp_sum(dsum);
    if (lane == 0) s_delta[j] = dsum;
  }
  ldelta = warp_sum(ldelta);
  if (lane == 0) s_loss[w] += ldelta;
  __syncthreads();
  if (threadIdx.x < kStripCols) {
    const int cc = c0 + threadIdx.x;
    if (cc < n) {
      float t = s_delta[threadIdx.x];
#pragma unroll
      for (int k = 0; k < kStripWarps; ++k) t += s_col[k][threadIdx.x];
      db[cc] = t;
    }
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kStripWarps; ++k) t += s_loss[k];
    atomicAdd(loss_acc, (double)t * (double)inv_b);
  }
}

// lse[r] = log sum_c exp(O[r,c]) from per-(n-tile,row) (max, sumexp) partials; loss += sum_r lse[r]*row_sum[r]/B
static __global__ void k_softmax_lse(const float* __restrict__ stat_max, const float* __restrict__ stat_sum,
                                     int n_tiles, int rows, const float* __restrict__ row_sum, float inv_b,
                                     float* __restrict__ lse, double* __restrict__ loss_acc) {
  __shared__ float s_part[8];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (r < rows) {
    float m = -INFINITY;
    for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, stat_max[(size_t)t * rows + r]);
    float s = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      float mt = stat_max[(size_t)t * rows + r];
      if (mt > -INFINITY) s += stat_sum[(size_t)t * rows + r] * expf(mt - m);
    }
    float l = m + logf(s);
    lse[r] = l;
    contrib = l * row_sum[r];
  }
  contrib = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_part[k];
    if (loss_acc) atomicAdd(loss_acc, (double)t * (double)inv_b);
  }
}

static __global__ void k_sumsq(const float* __restrict__ x, long long rows, int cols, int ld,
                               double* __restrict__ out) {
  __shared__ double s_part[8];
  double acc = 0.0;
  const long long total = rows * (long long)cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i / cols;
    int c = (int)(i % cols);
    double v = x[r * ld + c];
    acc += v * v;
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_part[k];
    atomicAdd(out, t);
  }
}

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_softmax_lse(const float* stat_max, const float* stat_sum, int n_tiles, int rows,
                               const float* row_sum, float inv_b, float* lse, double* loss_acc, void* stream) {
  RCD_CHECK_ARG(stat_max && stat_sum && row_sum && lse, "null pointer");
  RCD_CHECK_ARG(n_tiles > 0 && rows > 0, "bad shape");
  k_softmax_lse<<<rcd_div_up(rows, 256), 256, 0, (cudaStream_t)stream>>>(stat_max, stat_sum, n_tiles, rows, row_sum,
                                                                        inv_b, lse, loss_acc);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_loss_grad(const uint16_t* O_bf16, int ldo, int rows, int n, int loss, float confidence,
                             float inv_b, const float* lse, const float* row_sum, const int32_t* csc_ptr,
                             const int32_t* csc_row, const float* csc_val, uint16_t* dO, int lddo, float* csc_corr,
                             float* db, double* loss_acc, void* stream) {
  RCD_CHECK_ARG(O_bf16 && csc_ptr && csc_row && csc_val && dO && csc_corr && db && loss_acc, "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0, "bad shape");
  RCD_CHECK_ARG(ldo % 2 == 0 && lddo % 2 == 0 && ldo >= n && lddo >= n, "ldo/lddo must be even and >= n");
  RCD_CHECK_ARG(loss != RCD_LOSS_NLL || (lse && row_sum), "NLL needs lse and row_sum");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = rcd_div_up(n, kStripCols);
  const int threads = kStripWarps * 32;
  switch (loss) {
    case RCD_LOSS_MSE:
      k_loss_grad<RCD_LOSS_MSE><<<blocks, threads, 0, st>>>(O_bf16, ldo, rows, n, confidence, inv_b, lse, row_sum,
                                                             csc_ptr, csc_row, csc_val, dO, lddo, csc_corr, db, loss_acc);
      break;
    case RCD_LOSS_NLL:
      k_loss_grad<RCD_LOSS_NLL><<<blocks, threads, 0, st>>>(O_bf16, ldo, rows, n, confidence, inv_b, lse, row_sum,
                                                             csc_ptr, csc_row, csc_val, dO, lddo, csc_corr, db, loss_acc);
      break;
    case RCD_LOSS_LOGISTIC:
      k_loss_grad<RCD_LOSS_LOGISTIC><<<blocks, threads, 0, st>>>(O_bf16, ldo, rows, n, confidence, inv_b, lse,
                                                                  row_sum, csc_ptr, csc_row, csc_val, dO, lddo,
                                                                  csc_corr, db, loss_acc);
      break;
    default:
      rcd_set_error("rcd_loss_grad: unknown loss id %d", loss);
      return RCD_ERR_INVALID;
  }
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_sumsq(const float* x, long long rows, int cols, int ld, double* out_sq, void* stream) {
  RCD_CHECK_ARG(x && out_sq && rows > 0 && cols > 0 && ld >= cols, "bad arguments");
  long long total = rows * (long long)cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_sumsq<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, out_sq);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

// Sparse side of the loss (fp32, HBM/L2-bound): everything that touches only the stored interactions.
//   rcd_sddmm       : logits at the stored targets, o_p = Zb[r,:] . Wg[col_p,:] + bias[col_p]  (same bf16 operands as
//                     the tensor-core GEMM, fp32 accumulate), the fp32 sparse part of dL/dlogits (exact gradient minus
//                     the dense target-free formula the GEMM epilogue writes) and the per-row softmax reference
//                     ref[r] = max_p o_p used by the fused multinomial-NLL epilogue (decoder_tc.cu).
//   rcd_loss_finish : per row: reduce the epilogue's partials, finish the loss (recoder/losses.py:43-47, 68-71,
//                     BCEWithLogitsLoss; the /B of recoder/model.py:483-484), produce the softmax row scale
//                     alpha[r] = S_r / (B * sum_c exp(o_rc - ref_r)) and the bf16 copy of alpha*Z that feeds dW_d.
#include "common.cuh"

namespace rcd {

constexpr int kSpWarps = 8;

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b, float acc) {
  acc = fmaf(bf16_lo(a.x), bf16_lo(b.x), acc);
  acc = fmaf(bf16_hi(a.x), bf16_hi(b.x), acc);
  acc = fmaf(bf16_lo(a.y), bf16_lo(b.y), acc);
  acc = fmaf(bf16_hi(a.y), bf16_hi(b.y), acc);
  acc = fmaf(bf16_lo(a.z), bf16_lo(b.z), acc);
  acc = fmaf(bf16_hi(a.z), bf16_hi(b.z), acc);
  acc = fmaf(bf16_lo(a.w), bf16_lo(b.w), acc);
  acc = fmaf(bf16_hi(a.w), bf16_hi(b.w), acc);
  return acc;
}

// kSdWarpsPerRow warps per row (a row holds O(100) stored targets; one warp alone would be latency-bound on its
// serial chain of embedding-row fetches); lanes stride the hidden dimension in 16-byte chunks; every warp keeps four
// independent Wg rows in flight.
constexpr int kSdWarpsPerRow = 4;
constexpr int kSdRowsPerBlock = kSpWarps / kSdWarpsPerRow;

template <int NCH>  // 16-byte chunks per lane kept in registers (H <= 256 * NCH)
static __global__ void __launch_bounds__(kSpWarps * 32)
    k_sddmm(const uint16_t* __restrict__ Zb, int ldzb, const uint16_t* __restrict__ Wg, int ldw,
            const float* __restrict__ bias_g, int nchunks, const int32_t* __restrict__ row_ptr,
            const int32_t* __restrict__ cols, const float* __restrict__ vals, int row0, int rows, int loss, float conf,
            float inv_b, float* __restrict__ o_nnz, float* __restrict__ corr, float* __restrict__ row_ref) {
  __shared__ float s_max[kSpWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kSdRowsPerBlock + w / kSdWarpsPerRow;
  const int sub = w % kSdWarpsPerRow;
  float rmax = -INFINITY;
  if (r < rows) {
    const int base = row_ptr[row0];
    const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
    uint4 z[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int ch = lane + 32 * k;
      z[k] = (ch < nchunks) ? __ldg(reinterpret_cast<const uint4*>(Zb + (size_t)r * ldzb) + ch) : make_uint4(0, 0, 0, 0);
    }
    for (int p0 = s + sub * 4; p0 < e; p0 += 4 * kSdWarpsPerRow) {
      int c[4];
      float a[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j] = (p0 + j < e) ? cols[p0 + j] : -1;
        a[j] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int ch = lane + 32 * k;
        if (ch < nchunks) {
          uint4 wv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            wv[j] = (c[j] >= 0) ? __ldg(reinterpret_cast<const uint4*>(Wg + (size_t)c[j] * ldw) + ch)
                                : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = dot8_bf16(z[k], wv[j], a[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = warp_sum(a[j]);
      if (lane < 4 && c[lane < 4 ? lane : 0] >= 0) {
        // lane j finishes target p0 + j
        const float acc = lane == 0 ? a[0] : lane == 1 ? a[1] : lane == 2 ? a[2] : a[3];
        const int cj = lane == 0 ? c[0] : lane == 1 ? c[1] : lane == 2 ? c[2] : c[3];
        const float o = acc + bias_g[cj];
        o_nnz[p0 + lane - base] = o;
        corr[p0 + lane - base] = sparse_corr(loss, o, vals[p0 + lane], conf, inv_b);
        rmax = fmaxf(rmax, o);
      }
    }
  }
  rmax = warp_max(rmax);
  if (lane == 0) s_max[w] = rmax;
  __syncthreads();
  if (row_ref && r < rows && sub == 0 && lane == 0) {
    float m = s_max[w];
#pragma unroll
    for (int j = 1; j < kSdWarpsPerRow; ++j) m = fmaxf(m, s_max[w + j]);
    row_ref[r] = (m == -INFINITY) ? 0.f : m;
  }
}

static __global__ void __launch_bounds__(kSpWarps * 32)
    k_loss_finish(const float* __restrict__ stat, int stat_ld, int stat_cols, int rows, int loss, float conf,
                  float inv_b, const float* __restrict__ row_ref, const float* __restrict__ row_sum,
                  const int32_t* __restrict__ row_ptr, const float* __restrict__ vals,
                  const float* __restrict__ o_nnz, int row0, float* __restrict__ row_scale,
                  const float* __restrict__ Z, int H, uint16_t* __restrict__ Zs, int ldzs,
                  double* __restrict__ loss_acc, int32_t* __restrict__ bad, int local_targets,
                  double* __restrict__ loss_blocks, int32_t* __restrict__ redo_flag, int32_t* __restrict__ row_redo,
                  const int32_t* __restrict__ cond) {
  if (cond != nullptr && *cond == 0) return;
  __shared__ double s_part[kSpWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kSpWarps + w;
  double row_loss = 0.0;
  if (r < rows) {
    float s = 0.f;
    for (int t = lane; t < stat_cols; t += 32) s += stat[(size_t)r * stat_ld + t];
    s = warp_sum(s);  // fixed order: deterministic
    const int base = row_ptr[row0];
    const int ps = row_ptr[row0 + r], pe = row_ptr[row0 + r + 1];
    float sp = 0.f, st = 0.f;
    for (int p = ps + lane; p < pe; p += 32) {
      const float o = o_nnz[p - base], t = vals[p];
      st += t;
      if (loss == RCD_LOSS_MSE) {
        const float wgt = 1.0f + (t > 0.f ? conf : 0.f);  // losses.py:44
        sp += wgt * (o - t) * (o - t) - o * o;
      } else {
        sp -= t * o;
      }
    }
    sp = warp_sum(sp);
    st = warp_sum(st);
    float alpha = 1.0f, lrow;
    if (loss == RCD_LOSS_NLL) {
      const float S = row_sum[row0 + r];
      // a row sum at the clamp means some logit sat >= 64 octaves above the reference: the row is redone with its
      // true maximum (decoder_tc.cu); without the redo machinery (item-parallel mode) it is an error
      const bool clamped = s >= 18446744073709551616.0f;  // 2^RCD_NLL_CLAMP_LOG2
      if (row_redo && lane == 0) row_redo[r] = clamped ? 1 : 0;
      const float T = local_targets ? st : S;
      if (T != 0.f) {
        const float lse = row_ref[r] + logf(s);
        // item-parallel mode: `stat` holds the all-reduced row sum, the stored targets are this rank's item shard —
        // the rank's loss share is (sum of ITS targets) * lse + sp, so that the shares add up to S * lse + sum sp
        lrow = T * lse + sp;
      } else {  // no target mass: the log-sum-exp term vanishes whatever the logits are (never 0 * log 0)
        lrow = sp;
      }
      alpha = (S != 0.f) ? S * inv_b / s : 0.f;
      if (clamped && redo_flag) {
        if (lane == 0) atomicOr(redo_flag, 1);
        lrow = 0.f;   // recomputed by the redo pass
      } else if (clamped || ((S != 0.f || T != 0.f) && (!(s > 0.f) || !isfinite(s)))) {
        if (lane == 0) atomicOr(bad, 1);
      }
    } else {
      lrow = s + sp;
    }
    if (!isfinite(lrow) && lane == 0) atomicOr(bad, 2);
    row_loss = (double)lrow;
    if (lane == 0 && row_scale) row_scale[r] = alpha;
    if (Zs) {
      for (int h = lane; h < ldzs; h += 32) {
        const float z = (h < H) ? alpha * Z[(size_t)r * H + h] : 0.f;
        reinterpret_cast<__nv_bfloat16*>(Zs)[(size_t)r * ldzs + h] = __float2bfloat16_rn(z);
      }
    }
  }
  if (lane == 0) s_part[w] = row_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kSpWarps; ++k) t += s_part[k];
    if (loss_blocks) loss_blocks[blockIdx.x] = t * (double)inv_b;
    else atomicAdd(loss_acc, t * (double)inv_b);
  }
}

// row_ref[r] = ln2 * max_t stat[r, t] for flagged rows (stat holds per-tile maxima of logit * log2 e)
static __global__ void k_nll_ref_fix(const float* __restrict__ stat, int stat_ld, int stat_cols, int rows,
                                     const int32_t* __restrict__ row_redo, float* __restrict__ row_ref,
                                     const int32_t* __restrict__ cond) {
  if (cond != nullptr && *cond == 0) return;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kSpWarps + w;
  if (r >= rows || row_redo[r] == 0) return;
  float m = -INFINITY;
  for (int t = lane; t < stat_cols; t += 32) m = fmaxf(m, stat[(size_t)r * stat_ld + t]);
  m = warp_max(m);
  if (lane == 0) row_ref[r] = m * 0.6931471805599453f;
}

// loss_acc += sum_b loss_blocks[b] in index order (one thread: a few hundred doubles); clears the redo flag
static __global__ void k_loss_sum(const double* __restrict__ loss_blocks, int nblocks, double* __restrict__ loss_acc,
                                  int32_t* __restrict__ redo_flag) {
  __shared__ double part[32];
  // 32 lanes take contiguous chunks, then a fixed-order sum of the 32 partials: deterministic
  const int lane = threadIdx.x;
  const int per = (nblocks + 31) / 32;
  double t = 0.0;
  for (int i = lane * per; i < min((lane + 1) * per, nblocks); ++i) t += loss_blocks[i];
  part[lane] = t;
  __syncwarp();
  if (lane == 0) {
    double s = 0.0;
    for (int k = 0; k < 32; ++k) s += part[k];
    *loss_acc += s;
    if (redo_flag) *redo_flag = 0;
  }
}

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_sddmm(const uint16_t* Zb, int ldzb, const uint16_t* Wg, int ldw, const float* bias_g, int H,
                         const int32_t* row_ptr, const int32_t* cols, const float* vals, int row0, int rows, int loss,
                         float confidence, float inv_b, float* o_nnz, float* corr, float* row_ref, void* stream) {
  RCD_CHECK_ARG(Zb && Wg && bias_g && row_ptr && cols && vals && o_nnz && corr, "null pointer");
  RCD_CHECK_ARG(rows > 0 && H > 0 && row0 >= 0, "bad shape");
  RCD_CHECK_ARG(ldzb % 8 == 0 && ldw % 8 == 0 && ldzb >= H && ldw >= H, "ld must be a multiple of 8 and >= H");
  RCD_CHECK_ARG(((reinterpret_cast<uintptr_t>(Zb) | reinterpret_cast<uintptr_t>(Wg)) & 15) == 0, "unaligned operand");
  const int nchunks = rcd_div_up(H, 8);
  const int blocks = rcd_div_up(rows, kSdRowsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
#define RCD_SDDMM(NCH)                                                                                              \
  k_sddmm<NCH><<<blocks, kSpWarps * 32, 0, st>>>(Zb, ldzb, Wg, ldw, bias_g, nchunks, row_ptr, cols, vals, row0, rows, \
                                                 loss, confidence, inv_b, o_nnz, corr, row_ref)
  if (nchunks <= 32) RCD_SDDMM(1);
  else if (nchunks <= 64) RCD_SDDMM(2);
  else if (nchunks <= 128) RCD_SDDMM(4);
  else if (nchunks <= 256) RCD_SDDMM(8);
  else {
    rcd_set_error("rcd_sddmm: hidden size %d too large", H);
    return RCD_ERR_UNSUPPORTED;
  }
#undef RCD_SDDMM
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_loss_finish(const float* stat, int stat_ld, int stat_cols, int rows, int loss, float confidence,
                               float inv_b, const float* row_ref, const float* row_sum, const int32_t* row_ptr,
                               const float* vals, const float* o_nnz, int row0, float* row_scale, const float* Z, int H,
                               uint16_t* Zs, int ldzs, double* loss_acc, int32_t* bad_flag, int local_targets,
                               double* loss_blocks, int32_t* redo_flag, int32_t* row_redo, const int32_t* cond,
                               void* stream) {
  RCD_CHECK_ARG(stat && row_ptr && vals && o_nnz && (loss_acc || loss_blocks) && bad_flag, "null pointer");
  RCD_CHECK_ARG(!redo_flag || row_redo, "redo_flag needs row_redo");
  RCD_CHECK_ARG(rows > 0 && stat_cols > 0 && stat_ld >= stat_cols && row0 >= 0, "bad shape");
  RCD_CHECK_ARG(loss != RCD_LOSS_NLL || (row_ref && row_sum && row_scale), "NLL needs row_ref, row_sum and row_scale");
  RCD_CHECK_ARG(!Zs || (Z && ldzs >= H), "Zs needs Z and ldzs >= H");
  k_loss_finish<<<rcd_div_up(rows, kSpWarps), kSpWarps * 32, 0, (cudaStream_t)stream>>>(
      stat, stat_ld, stat_cols, rows, loss, confidence, inv_b, row_ref, row_sum, row_ptr, vals, o_nnz, row0, row_scale,
      Z, H, Zs, ldzs, loss_acc, bad_flag, local_targets, loss_blocks, redo_flag, row_redo, cond);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_loss_finish_blocks(int rows) { return rcd_div_up(rows > 0 ? rows : 1, kSpWarps); }

RCD_EXPORT int rcd_nll_ref_fix(const float* stat, int stat_ld, int stat_cols, int rows, const int32_t* row_redo,
                               float* row_ref, const int32_t* cond, void* stream) {
  RCD_CHECK_ARG(stat && row_redo && row_ref && rows > 0 && stat_cols > 0 && stat_ld >= stat_cols, "bad arguments");
  k_nll_ref_fix<<<rcd_div_up(rows, kSpWarps), kSpWarps * 32, 0, (cudaStream_t)stream>>>(stat, stat_ld, stat_cols, rows,
                                                                                      row_redo, row_ref, cond);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_loss_sum(const double* loss_blocks, int nblocks, double* loss_acc, int32_t* redo_flag,
                            void* stream) {
  RCD_CHECK_ARG(loss_blocks && loss_acc && nblocks > 0, "bad arguments");
  k_loss_sum<<<1, 32, 0, (cudaStream_t)stream>>>(loss_blocks, nblocks, loss_acc, redo_flag);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

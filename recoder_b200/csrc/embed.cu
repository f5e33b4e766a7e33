// K2/K3/K7: embedding-row kernels (HBM-bound, fp32 exact).
//   gather         : fp32 master rows -> bf16 GEMM operand (+ optional fp32 copy, optional activation)
//   encoder fwd    : Z = act(row_scale * sum_p x_p * We[item_p,:] + be)     (sparse formulation of nn.py:235-240)
//   dz_act         : dA = (sum of split-K partials) * act'(Z), db = column sums
//   encoder wgrad  : dWe_rows[c,:] = sum_{entries of column c} x * row_scale * dA[row,:]
// Thread mapping: a group of `tpr` threads owns one output row and strides over its H columns with 16-byte
// vectors; consecutive threads touch consecutive 16 B so every warp request is fully coalesced.
#include "common.cuh"

namespace rcd {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void fma(float a, const Vec<4>& o) {
    v.x = fmaf(a, o.v.x, v.x);
    v.y = fmaf(a, o.v.y, v.y);
    v.z = fmaf(a, o.v.z, v.z);
    v.w = fmaf(a, o.v.w, v.w);
  }
  __device__ __forceinline__ float get(int i) const { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
  __device__ __forceinline__ void fma(float a, const Vec<1>& o) { v = fmaf(a, o.v, v); }
  __device__ __forceinline__ float get(int) const { return v; }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
};

constexpr int kEmbThreads = 256;
constexpr int kHeavyChunk = 128;  // entries per chunk of a heavy column (see k_heavy_setup)

// acc[k] (k < NV) += sum_e coef(e) * M[idx(e), (t + k*tpr)*VEC ...]
template <int VEC, int NV, class Entry>
__device__ __forceinline__ void seg_accumulate(Vec<VEC> (&acc)[NV], const float* __restrict__ M, int H, int t, int tpr,
                                               int s, int e, Entry entry) {
  int p = s;
  for (; p + 4 <= e; p += 4) {  // 4 independent row fetches in flight
    int i0, i1, i2, i3;
    float c0, c1, c2, c3;
    entry(p, i0, c0);
    entry(p + 1, i1, c1);
    entry(p + 2, i2, c2);
    entry(p + 3, i3, c3);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int h = (t + k * tpr) * VEC;
      if (h < H) {
        Vec<VEC> w0, w1, w2, w3;
        w0.load(M + (size_t)i0 * H + h);
        w1.load(M + (size_t)i1 * H + h);
        w2.load(M + (size_t)i2 * H + h);
        w3.load(M + (size_t)i3 * H + h);
        acc[k].fma(c0, w0);
        acc[k].fma(c1, w1);
        acc[k].fma(c2, w2);
        acc[k].fma(c3, w3);
      }
    }
  }
  for (; p < e; ++p) {
    int i0;
    float c0;
    entry(p, i0, c0);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int h = (t + k * tpr) * VEC;
      if (h < H) {
        Vec<VEC> w0;
        w0.load(M + (size_t)i0 * H + h);
        acc[k].fma(c0, w0);
      }
    }
  }
}

template <int VEC, int NV>
static __global__ void __launch_bounds__(kEmbThreads)
    k_encoder_fwd(const float* __restrict__ We, int H, const float* __restrict__ be,
                  const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ raw_items,
                  const float* __restrict__ vals, const float* __restrict__ row_inv_norm, int row0, int rows, int act,
                  int tpr, float* __restrict__ Z, uint16_t* __restrict__ Zb, int ldzb) {
  const int rpb = kEmbThreads / tpr;
  const int r = blockIdx.x * rpb + threadIdx.x / tpr;
  const int t = threadIdx.x % tpr;
  if (r >= rows) return;
  const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
  Vec<VEC> acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k].zero();
  seg_accumulate<VEC, NV>(acc, We, H, t, tpr, s, e, [&](int p, int& idx, float& c) {
    idx = raw_items[p];
    c = vals[p];
  });
  const float scale = row_inv_norm[row0 + r];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int h = (t + k * tpr) * VEC;
    if (h < H) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float z = act_apply(fmaf(scale, acc[k].get(i), be[h + i]), act);
        Z[(size_t)r * H + h + i] = z;
        if (Zb) reinterpret_cast<__nv_bfloat16*>(Zb)[(size_t)r * ldzb + h + i] = __float2bfloat16_rn(z);
      }
    }
  }
  if (Zb)
    for (int h = H + t; h < ldzb; h += tpr) Zb[(size_t)r * ldzb + h] = 0;
}

// out[c,:] (+)= sum_{entries e of column c} coef(e) * M[csc_row[e],:], coef(e) = csc_val[e] * row_scale[row0+row]  (encoder
// weight gradient) or, when `src` is given, coef(e) = csc_val[src[e]] looked up through the CSC->CSR permutation
// (sparse part of the decoder weight gradient); db[c] += sum_e coef(e) when db is given.
template <int VEC, int NV>
static __global__ void __launch_bounds__(kEmbThreads, (NV <= 1) ? 8 : 1)  // latency-bound gather: keep 8 CTAs per SM
    k_encoder_wgrad(const float* __restrict__ dA, int H, const int32_t* __restrict__ csc_ptr,
                    const int32_t* __restrict__ csc_row, const float* __restrict__ csc_val,
                    const int32_t* __restrict__ src, const float* __restrict__ row_inv_norm, int row0, int n, int tpr,
                    float* __restrict__ out, int accumulate, float* __restrict__ db,
                    const int32_t* __restrict__ slot_base, int heavy_blocks, const int32_t* __restrict__ counter,
                    const int32_t* __restrict__ chunk_col, int cap, float* __restrict__ partial,
                    float* __restrict__ partial_db) {
  const int rpb = kEmbThreads / tpr;
  const int t = threadIdx.x % tpr;
  if ((int)blockIdx.x < heavy_blocks) {
    // ---- one 128-entry chunk of a heavy column (scheduled first, so the long columns overlap the light ones) ----
    const int q = blockIdx.x * rpb + threadIdx.x / tpr;
    const int total = min(*counter, cap);
    if (q >= total) return;
    const int c = chunk_col[q];
    const int j = q - slot_base[c];
    const int s = csc_ptr[c] + j * kHeavyChunk, e = min(s + kHeavyChunk, csc_ptr[c + 1]);
    Vec<VEC> acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k].zero();
    seg_accumulate<VEC, NV>(acc, dA, H, t, tpr, s, e, [&](int p, int& idx, float& cf) {
      idx = csc_row[p];
      cf = src ? csc_val[src[p]] : csc_val[p];
      if (row_inv_norm) cf *= row_inv_norm[row0 + idx];
    });
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      int h = (t + k * tpr) * VEC;
      if (h < H) acc[k].store(partial + (size_t)q * H + h);
    }
    if (partial_db && t == 0) {
      float sum = 0.f;
      for (int p = s; p < e; ++p) sum += src ? csc_val[src[p]] : csc_val[p];
      partial_db[q] = sum;
    }
    return;
  }
  const int c = ((int)blockIdx.x - heavy_blocks) * rpb + threadIdx.x / tpr;
  if (c >= n) return;
  if (slot_base && slot_base[c] >= 0) return;  // heavy column: chunked path above
  const int s = csc_ptr[c], e = csc_ptr[c + 1];
  Vec<VEC> acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k].zero();
  seg_accumulate<VEC, NV>(acc, dA, H, t, tpr, s, e, [&](int p, int& idx, float& cf) {
    idx = csc_row[p];
    cf = src ? csc_val[src[p]] : csc_val[p];
    if (row_inv_norm) cf *= row_inv_norm[row0 + idx];
  });
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int h = (t + k * tpr) * VEC;
    if (h < H) {
      if (accumulate) {
        Vec<VEC> prev;
        prev.load(out + (size_t)c * H + h);
        acc[k].fma(1.0f, prev);
      }
      acc[k].store(out + (size_t)c * H + h);
    }
  }
  if (db && t == 0) {
    float sum = 0.f;
    for (int p = s; p < e; ++p) sum += src ? csc_val[src[p]] : csc_val[p];  // fixed order
    db[c] += sum;
  }
}

// ---- heavy columns ---------------------------------------------------------------------------------------------------
// Item popularity is a power law: a few columns of a slice hold an entry in almost every row (up to `rows` entries, 16K
// in the item-parallel mode), and one thread group walking such a column alone is a serial chain of L2 round trips that
// outlasts the whole rest of the kernel.  Columns with more than kHeavyChunk entries are therefore cut into chunks of
// kHeavyChunk entries, one thread group per chunk writes a partial row, and the first chunk's group adds the partials up
// in chunk order (deterministic).  Chunk slots are handed out with one atomic counter per launch.
static __global__ void k_heavy_setup(const int32_t* __restrict__ csc_ptr, int n, int32_t* __restrict__ counter,
                                     int32_t* __restrict__ slot_base, int32_t* __restrict__ chunk_col, int cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int cnt = csc_ptr[c + 1] - csc_ptr[c];
  if (cnt <= kHeavyChunk) {
    slot_base[c] = -1;
    return;
  }
  const int nch = (cnt + kHeavyChunk - 1) / kHeavyChunk;
  const int base = atomicAdd(counter, nch);
  slot_base[c] = base;
  for (int j = 0; j < nch; ++j)
    if (base + j < cap) chunk_col[base + j] = c;
}

static __global__ void __launch_bounds__(kEmbThreads)
    k_heavy_reduce(const int32_t* __restrict__ csc_ptr, int H, int tpr, const int32_t* __restrict__ counter,
                   const int32_t* __restrict__ slot_base, const int32_t* __restrict__ chunk_col, int cap,
                   const float* __restrict__ partial, const float* __restrict__ partial_db, float* __restrict__ out,
                   int accumulate, float* __restrict__ db) {
  const int rpb = kEmbThreads / tpr;
  const int q = blockIdx.x * rpb + threadIdx.x / tpr;
  const int t = threadIdx.x % tpr;
  const int total = min(*counter, cap);
  if (q >= total) return;
  const int c = chunk_col[q];
  if (slot_base[c] != q) return;  // the first chunk of a column does the reduction
  const int nch = (csc_ptr[c + 1] - csc_ptr[c] + kHeavyChunk - 1) / kHeavyChunk;
  for (int h = t; h < H; h += tpr) {
    float s = 0.f;
    for (int j = 0; j < nch; ++j) s += partial[(size_t)(q + j) * H + h];  // fixed order
    if (accumulate) s += out[(size_t)c * H + h];
    out[(size_t)c * H + h] = s;
  }
  if (db && partial_db && t == 0) {
    float s = 0.f;
    for (int j = 0; j < nch; ++j) s += partial_db[q + j];
    db[c] += s;
  }
}

// out[r,:] = sum_p corr[p] * W[raw_items[p],:]  — the fp32 sparse part of dZ = dO @ W (corr from rcd_sddmm,
// indexed relative to the first stored entry of the slice)
template <int VEC, int NV>
static __global__ void __launch_bounds__(kEmbThreads)
    k_sparse_dgrad(const float* __restrict__ W, int H, const int32_t* __restrict__ row_ptr,
                   const int32_t* __restrict__ raw_items, const float* __restrict__ corr, int row0, int rows, int tpr,
                   float* __restrict__ out, int ldp) {
  const int rpb = kEmbThreads / tpr;
  const int r = blockIdx.x * rpb + threadIdx.x / tpr;
  const int t = threadIdx.x % tpr;
  if (r >= rows) return;
  const int base = row_ptr[row0];
  const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
  Vec<VEC> acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k].zero();
  seg_accumulate<VEC, NV>(acc, W, H, t, tpr, s, e, [&](int p, int& idx, float& cf) {
    idx = raw_items[p];
    cf = corr[p - base];
  });
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int h = (t + k * tpr) * VEC;
    if (h < H) acc[k].store(out + (size_t)r * ldp + h);
  }
}

// flat mapping: one thread per 8 output columns of one row
static __global__ void k_gather_rows(const float* __restrict__ table, int H, const int64_t* __restrict__ ids, int n,
                                     int act, uint16_t* __restrict__ out_bf16, int ld_out,
                                     float* __restrict__ out_f32, int vec_ok) {
  const int cpr = ld_out >> 3;  // 8-column chunks per row
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * cpr) return;
  const int r = (int)(i / cpr), ch = (int)(i % cpr);
  const long long src = ids ? ids[r] : r;
  const float* row = table + (size_t)src * H;
  const int h0 = ch * 8;
  float v[8];
  if (vec_ok && h0 + 8 <= H) {
    float4 a = __ldg(reinterpret_cast<const float4*>(row + h0));
    float4 b = __ldg(reinterpret_cast<const float4*>(row + h0 + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (h0 + k < H) ? __ldg(row + h0 + k) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = (h0 + k < H) ? act_apply(v[k], act) : 0.f;
  if (out_f32) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (h0 + k < H) out_f32[(size_t)r * H + h0 + k] = v[k];
  }
  if (out_bf16) {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out_bf16 + (size_t)r * ld_out + h0) = o;
  }
}

static __global__ void k_gather_vec(const float* __restrict__ vec, const int64_t* __restrict__ ids, int n,
                                    float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = vec[ids ? ids[i] : i];
}

// dA = (row_scale[r] * sum_{k < n_scaled} partials[k] + sum_{k >= n_scaled} partials[k]) * act'(Z)
static __global__ void k_dz_act(const float* __restrict__ partials, int splits, int n_scaled,
                                const float* __restrict__ row_scale, long long split_stride, int ldp,
                                const float* __restrict__ Z, int rows, int H, int act, float* __restrict__ dA) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * H) return;
  const int r = (int)(i / H), h = (int)(i % H);
  float s = 0.f;
  for (int k = 0; k < n_scaled; ++k) s += partials[k * split_stride + (size_t)r * ldp + h];  // fixed order
  if (row_scale) s *= row_scale[r];
  for (int k = n_scaled; k < splits; ++k) s += partials[k * split_stride + (size_t)r * ldp + h];
  dA[i] = s * act_grad_from_out(Z[i], act);
}

// db[h] = sum_r x[r,h]; one block per 32 columns, 32 warps stride the rows with four loads in flight, fixed-order
// smem reduction.
constexpr int kColsumWarpsE = 32;
static __global__ void __launch_bounds__(kColsumWarpsE * 32)
    k_colsum_f32(const float* __restrict__ x, int rows, int H, float* __restrict__ db) {
  __shared__ float part[kColsumWarpsE][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int h = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (h < H) {
    int r = w;
    for (; r + 3 * kColsumWarpsE < rows; r += 4 * kColsumWarpsE) {
      s0 += x[(size_t)r * H + h];
      s1 += x[(size_t)(r + kColsumWarpsE) * H + h];
      s2 += x[(size_t)(r + 2 * kColsumWarpsE) * H + h];
      s3 += x[(size_t)(r + 3 * kColsumWarpsE) * H + h];
    }
    for (; r < rows; r += kColsumWarpsE) s0 += x[(size_t)r * H + h];
  }
  part[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && h < H) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kColsumWarpsE; ++k) t += part[k][lane];
    db[h] = t;
  }
}

static inline int pick_tpr(int units) {  // units = number of VEC-wide chunks per row
  int t = 1;
  while (t < units && t < kEmbThreads) t <<= 1;
  return t;
}

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_gather_rows(const float* table, int H, const int64_t* ids, int n, int act, uint16_t* out_bf16,
                               int ld_out, float* out_f32, void* stream) {
  RCD_CHECK_ARG(table && n > 0 && H > 0, "null table or empty gather");
  RCD_CHECK_ARG(out_bf16 || out_f32, "no output");
  if (!out_bf16) ld_out = (H + 7) / 8 * 8;
  RCD_CHECK_ARG(ld_out % 8 == 0 && ld_out >= H, "ld_out must be a multiple of 8 and >= H");
  long long threads = (long long)n * (ld_out / 8);
  int vec_ok = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(table) & 15) == 0);
  k_gather_rows<<<rcd_div_up(threads, 256), 256, 0, (cudaStream_t)stream>>>(table, H, ids, n, act, out_bf16, ld_out,
                                                                             out_f32, vec_ok);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_gather_vec(const float* vec, const int64_t* ids, int n, float* out, void* stream) {
  RCD_CHECK_ARG(vec && out && n > 0, "null pointer or empty gather");
  k_gather_vec<<<rcd_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(vec, ids, n, out);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

#define RCD_DISPATCH_NV(KERNEL, VECW, units, tpr, ...)                                              \
  do {                                                                                              \
    int _nv = rcd_div_up(units, tpr);                                                               \
    if (_nv <= 1) KERNEL<VECW, 1> __VA_ARGS__;                                                       \
    else if (_nv <= 2) KERNEL<VECW, 2> __VA_ARGS__;                                                  \
    else if (_nv <= 4) KERNEL<VECW, 4> __VA_ARGS__;                                                  \
    else if (_nv <= 8) KERNEL<VECW, 8> __VA_ARGS__;                                                  \
    else {                                                                                          \
      rcd_set_error("%s: hidden size %d too large", __func__, H);                                   \
      return RCD_ERR_UNSUPPORTED;                                                                   \
    }                                                                                               \
  } while (0)

RCD_EXPORT int rcd_ae_encoder_fwd(const float* We, int H, const float* be, const int32_t* row_ptr,
                                  const int32_t* raw_items, const float* vals, const float* row_inv_norm, int row0,
                                  int rows, int act, float* Z, uint16_t* Zb, int ldzb, void* stream) {
  RCD_CHECK_ARG(We && be && row_ptr && raw_items && vals && row_inv_norm && Z, "null pointer");
  RCD_CHECK_ARG(rows > 0 && H > 0 && row0 >= 0, "bad shape");
  RCD_CHECK_ARG(!Zb || ldzb >= H, "ldzb < H");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(We) & 15) == 0);
  const int units = vec ? H / 4 : H;
  const int tpr = pick_tpr(units);
  const int blocks = rcd_div_up(rows, kEmbThreads / tpr);
  if (vec)
    RCD_DISPATCH_NV(k_encoder_fwd, 4, units, tpr, <<<blocks, kEmbThreads, 0, st>>>(
        We, H, be, row_ptr, raw_items, vals, row_inv_norm, row0, rows, act, tpr, Z, Zb, ldzb));
  else
    RCD_DISPATCH_NV(k_encoder_fwd, 1, units, tpr, <<<blocks, kEmbThreads, 0, st>>>(
        We, H, be, row_ptr, raw_items, vals, row_inv_norm, row0, rows, act, tpr, Z, Zb, ldzb));
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

// shared driver of the two column-major accumulations (encoder weight gradient / sparse part of dW_d)
static int csc_accumulate(const float* M, int H, const int32_t* csc_ptr, const int32_t* csc_row, const float* vals,
                          const int32_t* src, const float* row_inv_norm, int row0, int n, float* out, int accumulate,
                          float* db, void* scratch, size_t scratch_bytes, long long nnz_hint, cudaStream_t st,
                          const char* who) {
  const bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(M) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int units = vec ? H / 4 : H;
  const int tpr = pick_tpr(units);
  const int rpb = kEmbThreads / tpr;
  const int nv = rcd_div_up(units, tpr);
  if (nv > 8) {
    rcd_set_error("%s: hidden size %d too large", who, H);
    return RCD_ERR_UNSUPPORTED;
  }
  // heavy-column workspace: counter (4 ints) | slot_base[n] | chunk_col[cap] | partial[cap*H] | partial_db[cap]
  int32_t *counter = nullptr, *slot_base = nullptr, *chunk_col = nullptr;
  float *partial = nullptr, *partial_db = nullptr;
  int cap = 0;
  if (scratch) {
    const size_t head = (size_t)(4 + n) * sizeof(int32_t);
    if (scratch_bytes > head + 1024) {
      const size_t per_chunk = sizeof(int32_t) + ((size_t)H + 1) * sizeof(float);
      size_t c = (scratch_bytes - head - 64) / per_chunk;
      const size_t want = (size_t)(2 * (nnz_hint > 0 ? nnz_hint : 0) / kHeavyChunk + 2);
      cap = (int)(c < want ? c : want);
      counter = reinterpret_cast<int32_t*>(scratch);
      slot_base = counter + 4;
      chunk_col = slot_base + n;
      partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(chunk_col + cap) + 15) & ~(uintptr_t)15);
      partial_db = partial + (size_t)cap * H;
    }
  }
  if (cap > 0) {
    RCD_CUDA(cudaMemsetAsync(counter, 0, 4 * sizeof(int32_t), st));
    k_heavy_setup<<<rcd_div_up(n, 256), 256, 0, st>>>(csc_ptr, n, counter, slot_base, chunk_col, cap);
    RCD_LAUNCH_CHECK();
  } else {
    slot_base = nullptr;
  }
  const int blocks = rcd_div_up(n, rpb);
  const int hb = cap > 0 ? rcd_div_up(cap, rpb) : 0;
#define RCD_CSC_LAUNCH(VECW, NVV)                                                                                     \
  do {                                                                                                                \
    k_encoder_wgrad<VECW, NVV><<<hb + blocks, kEmbThreads, 0, st>>>(                                                  \
        M, H, csc_ptr, csc_row, vals, src, row_inv_norm, row0, n, tpr, out, accumulate, db, slot_base, hb, counter,   \
        chunk_col, cap, partial, db ? partial_db : nullptr);                                                          \
    RCD_LAUNCH_CHECK();                                                                                               \
    if (cap > 0) {                                                                                                    \
      k_heavy_reduce<<<hb, kEmbThreads, 0, st>>>(csc_ptr, H, tpr, counter, slot_base, chunk_col, cap, partial,        \
                                                 db ? partial_db : nullptr, out, accumulate, db);                     \
      RCD_LAUNCH_CHECK();                                                                                             \
    }                                                                                                                 \
  } while (0)
  if (vec) {
    if (nv <= 1) RCD_CSC_LAUNCH(4, 1);
    else if (nv <= 2) RCD_CSC_LAUNCH(4, 2);
    else if (nv <= 4) RCD_CSC_LAUNCH(4, 4);
    else RCD_CSC_LAUNCH(4, 8);
  } else {
    if (nv <= 1) RCD_CSC_LAUNCH(1, 1);
    else if (nv <= 2) RCD_CSC_LAUNCH(1, 2);
    else if (nv <= 4) RCD_CSC_LAUNCH(1, 4);
    else RCD_CSC_LAUNCH(1, 8);
  }
#undef RCD_CSC_LAUNCH
  return RCD_OK;
}

RCD_EXPORT size_t rcd_csc_heavy_scratch_bytes(int n, long long nnz_slice, int H) {
  const size_t cap = (size_t)(2 * (nnz_slice > 0 ? nnz_slice : 0) / kHeavyChunk + 2);
  return (size_t)(4 + n) * sizeof(int32_t) + cap * (sizeof(int32_t) + ((size_t)H + 1) * sizeof(float)) + 2048;
}

RCD_EXPORT int rcd_ae_encoder_wgrad(const float* dA, int H, const int32_t* csc_ptr, const int32_t* csc_row,
                                    const float* csc_val, const float* row_inv_norm, int row0, int n,
                                    float* dWe_rows, const int32_t* csc_src, const float* csr_vals, void* scratch,
                                    size_t scratch_bytes, long long nnz_slice, void* stream) {
  RCD_CHECK_ARG(dA && csc_ptr && csc_row && csc_val && row_inv_norm && dWe_rows, "null pointer");
  RCD_CHECK_ARG((csc_src == nullptr) == (csr_vals == nullptr), "csc_src and csr_vals come together");
  if (csr_vals) csc_val = csr_vals;  // input values in CSR order of the slice (noised), reached through csc_src
  RCD_CHECK_ARG(n > 0 && H > 0 && row0 >= 0, "bad shape");
  return csc_accumulate(dA, H, csc_ptr, csc_row, csc_val, csc_src, row_inv_norm, row0, n, dWe_rows, 0, nullptr, scratch,
                        scratch_bytes, nnz_slice, (cudaStream_t)stream, "rcd_ae_encoder_wgrad");
}

RCD_EXPORT int rcd_csc_rows_accumulate(const float* M, int H, const int32_t* csc_ptr, const int32_t* csc_row,
                                       const int32_t* csc_src, const float* coef, int n, float* out, float* db,
                                       void* scratch, size_t scratch_bytes, long long nnz_slice, void* stream) {
  RCD_CHECK_ARG(M && csc_ptr && csc_row && coef && out, "null pointer");
  RCD_CHECK_ARG(n > 0 && H > 0, "bad shape");
  return csc_accumulate(M, H, csc_ptr, csc_row, coef, csc_src, nullptr, 0, n, out, 1, db, scratch, scratch_bytes,
                        nnz_slice, (cudaStream_t)stream, "rcd_csc_rows_accumulate");
}

RCD_EXPORT int rcd_sparse_dgrad(const float* W, int H, const int32_t* row_ptr, const int32_t* raw_items,
                                const float* corr, int row0, int rows, float* out, int ldp, void* stream) {
  RCD_CHECK_ARG(W && row_ptr && raw_items && corr && out, "null pointer");
  RCD_CHECK_ARG(rows > 0 && H > 0 && row0 >= 0 && ldp >= H, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (H % 4 == 0) && (ldp % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int units = vec ? H / 4 : H;
  const int tpr = pick_tpr(units);
  const int blocks = rcd_div_up(rows, kEmbThreads / tpr);
  if (vec)
    RCD_DISPATCH_NV(k_sparse_dgrad, 4, units, tpr, <<<blocks, kEmbThreads, 0, st>>>(
        W, H, row_ptr, raw_items, corr, row0, rows, tpr, out, ldp));
  else
    RCD_DISPATCH_NV(k_sparse_dgrad, 1, units, tpr, <<<blocks, kEmbThreads, 0, st>>>(
        W, H, row_ptr, raw_items, corr, row0, rows, tpr, out, ldp));
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_dz_act(const float* partials, int splits, int n_scaled, const float* row_scale, int ldp,
                          const float* Z, int rows, int H, int act, float* dA, float* db, void* stream) {
  RCD_CHECK_ARG(partials && Z && dA, "null pointer");
  RCD_CHECK_ARG(rows > 0 && H > 0 && splits > 0 && ldp >= H && n_scaled >= 0 && n_scaled <= splits, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  long long total = (long long)rows * H;
  k_dz_act<<<rcd_div_up(total, 256), 256, 0, st>>>(partials, splits, row_scale ? n_scaled : splits, row_scale,
                                                   (long long)rows * ldp, ldp, Z, rows, H, act, dA);
  RCD_LAUNCH_CHECK();
  if (db) {
    k_colsum_f32<<<rcd_div_up(H, 32), kColsumWarpsE * 32, 0, st>>>(dA, rows, H, db);
    RCD_LAUNCH_CHECK();
  }
  return RCD_OK;
}

// Inference side of the path: `Recoder.recommend` (recoder/model.py:525-544) masks the items a
// user has already interacted with to -inf and takes torch.topk(k, sorted=True) of the full-width logits.
//   rcd_mask_seen : logits[r, item] = -inf for every stored interaction of row r (straight from the pool's CSR; the
//                   dense `output[input > 0] = -inf` pass over [B, I] never happens)
//   rcd_topk_rows : per row, the k largest logits in descending order (ties: lower item id first), values + int64 ids.
//                   One CTA per row: 4-pass 8-bit radix select on order-preserving uint32 keys finds the k-th key,
//                   one more pass collects the winners into shared memory, a bitonic sort orders them.
//                   HBM/L2-bound: 5 passes over a row of n floats; k <= 1024.
#include "common.cuh"

namespace rcd {

static __global__ void k_mask_seen(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ items, int row0,
                                   int rows, float* __restrict__ logits, long long ld) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
  for (int p = s + threadIdx.x; p < e; p += blockDim.x) logits[(size_t)r * ld + items[p]] = -INFINITY;
}

__device__ __forceinline__ uint32_t float_key(float f) {  // larger float <=> larger key; NaN sorts above +inf
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int kTopkThreads = 256;
constexpr int kTopkMax = 1024;

static __global__ void __launch_bounds__(kTopkThreads)
    k_topk_rows(const float* __restrict__ logits, long long ld, int n, int k, float* __restrict__ out_val,
                int64_t* __restrict__ out_idx) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_need, s_count_gt, s_count_eq;
  __shared__ unsigned long long cand[kTopkMax];  // (key << 32) | ~index : sorting descending gives key desc, index asc
  const float* row = logits + (size_t)blockIdx.x * ld;
  const int t = threadIdx.x;

  // ---- radix select: find the key of the k-th largest element --------------------------------------------------
  uint32_t prefix = 0, mask = 0, need = (uint32_t)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[t] = 0;
    __syncthreads();
    for (int i = t; i < n; i += kTopkThreads) {
      const uint32_t key = float_key(row[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (t == 0) {
      uint32_t acc = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((uint32_t)b << shift);
      s_need = need - acc;  // how many of the elements inside bin b are still wanted
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    mask |= 0xffu << shift;
    __syncthreads();
  }
  const uint32_t kth = prefix;  // `need` elements equal to kth belong to the top k (lowest indices first)

  // ---- collect: every element above kth, then the first `need` elements equal to it (index order) ----------------
  if (t == 0) {
    s_count_gt = 0;
    s_count_eq = 0;
  }
  __syncthreads();
  for (int i = t; i < n; i += kTopkThreads) {
    const uint32_t key = float_key(row[i]);
    if (key > kth) {
      const uint32_t slot = atomicAdd(&s_count_gt, 1u);
      cand[slot] = ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)i);
    }
  }
  __syncthreads();
  const uint32_t n_gt = s_count_gt;  // == k - need
  // ties: ordered scan in chunks so that the lowest indices win deterministically
  for (int base = 0; base < n && s_count_eq < need; base += kTopkThreads) {
    const int i = base + t;
    const bool eq = (i < n) && (float_key(row[i]) == kth);
    const unsigned ball = __ballot_sync(0xffffffffu, eq);
    __shared__ uint32_t warp_cnt[kTopkThreads / 32];
    if ((t & 31) == 0) warp_cnt[t >> 5] = __popc(ball);
    __syncthreads();
    uint32_t before = s_count_eq;
    for (int w = 0; w < (t >> 5); ++w) before += warp_cnt[w];
    before += __popc(ball & ((1u << (t & 31)) - 1u));
    if (eq && before < need) cand[n_gt + before] = ((unsigned long long)kth << 32) | (uint32_t)(~(uint32_t)i);
    __syncthreads();
    if (t == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < kTopkThreads / 32; ++w) tot += warp_cnt[w];
      s_count_eq += tot;
    }
    __syncthreads();
  }

  // ---- bitonic sort of the k candidates, descending ----------------------------------------------------------------
  int m = 1;
  while (m < k) m <<= 1;
  for (int i = k + t; i < m; i += kTopkThreads) cand[i] = 0ull;  // padding sorts last
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = t; i < m; i += kTopkThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = ((i & size) == 0);
          const unsigned long long a = cand[i], b = cand[j];
          if (desc ? (a < b) : (a > b)) {
            cand[i] = b;
            cand[j] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = t; i < k; i += kTopkThreads) {
    const uint32_t idx = ~(uint32_t)(cand[i] & 0xffffffffull);
    out_idx[(size_t)blockIdx.x * k + i] = (int64_t)idx;
    out_val[(size_t)blockIdx.x * k + i] = row[idx];
  }
}

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_mask_seen(const int32_t* row_ptr, const int32_t* items, int row0, int rows, float* logits,
                             long long ld, void* stream) {
  RCD_CHECK_ARG(row_ptr && items && logits && rows > 0 && row0 >= 0 && ld > 0, "bad arguments");
  k_mask_seen<<<rows, 128, 0, (cudaStream_t)stream>>>(row_ptr, items, row0, rows, logits, ld);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_topk_rows(const float* logits, long long ld, int rows, int n, int k, float* out_val,
                             int64_t* out_idx, void* stream) {
  RCD_CHECK_ARG(logits && out_val && out_idx && rows > 0 && n > 0 && ld >= n, "bad arguments");
  RCD_CHECK_ARG(k > 0 && k <= n && k <= kTopkMax, "k must be in [1, min(n, 1024)]");
  k_topk_rows<<<rows, kTopkThreads, 0, (cudaStream_t)stream>>>(logits, ld, n, k, out_val, out_idx);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

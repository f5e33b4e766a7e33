// K1: GPU collate over an HBM-resident CSR — integer bookkeeping, bit-exact against
// BatchCollator.collate (recoder/data.py:203-251) and RecommendationDataset._extract (recoder/data.py:63-83).
//
// All kernels are HBM/latency bound; the unit of parallelism is one warp per pool row (rows hold O(100) nnz)
// and one thread per item for the flag scan.  Algorithmic bytes per pool nnz: 8 B read (index+value) +
// 12 B written (raw id, column, value); per item 12 B (flag, rank, pos) — see DESIGN.md.
#include "common.cuh"
#include "scan.cuh"

namespace rcd {

constexpr int kRowWarps = 8;  // warps per block for warp-per-row kernels

// One warp per pool row: row length, L2 norm, value sum, and the "item present" flags.
static __global__ void k_rows_mark(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const float* __restrict__ data, const int64_t* __restrict__ users, int P,
                                   int* __restrict__ row_len, float* __restrict__ row_inv_norm,
                                   float* __restrict__ row_sum, int* __restrict__ flag) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= P) return;
  const int lane = threadIdx.x & 31;
  const int64_t u = users[r];
  const int64_t s = indptr[u], e = indptr[u + 1];
  float sq = 0.f, sm = 0.f;
  for (int64_t p = s + lane; p < e; p += 32) {
    float v = data[p];
    sq += v * v;
    sm += v;
    if (flag) flag[indices[p]] = 1;  // benign race: every writer stores 1
  }
  sq = warp_sum(sq);
  sm = warp_sum(sm);
  if (lane == 0) {
    row_len[r] = (int)(e - s);
    row_inv_norm[r] = 1.0f / fmaxf(sqrtf(sq), 1e-12f);  // F.normalize(p=2, eps=1e-12), nn.py:235
    row_sum[r] = sm;
  }
}

// items[rank[i]] = i and pos[i] = rank[i] for flagged items, pos[i] = -1 otherwise.
static __global__ void k_compact_items(const int* __restrict__ flag, const int* __restrict__ rank, int I,
                                       int64_t* __restrict__ items, int32_t* __restrict__ pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= I) return;
  if (flag[i]) {
    int k = rank[i];
    items[k] = i;
    pos[i] = k;
  } else {
    pos[i] = -1;
  }
}

static __global__ void k_identity_items(int I, int64_t* __restrict__ items, int32_t* __restrict__ pos,
                                        int32_t* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) counts[0] = I;
  if (i >= I) return;
  if (items) items[i] = i;
  pos[i] = i;
}

// One warp per pool row: copy the row's nnz in stored order, remapping the column through pos.
static __global__ void k_emit_rows(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const float* __restrict__ data, const int64_t* __restrict__ users, int P,
                                   const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ pos,
                                   int nnz_capacity, int32_t* __restrict__ raw_items, int32_t* __restrict__ cols,
                                   float* __restrict__ vals) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= P) return;
  const int lane = threadIdx.x & 31;
  const int64_t u = users[r];
  const int64_t s = indptr[u];
  const int len = (int)(indptr[u + 1] - s);
  const int o = row_ptr[r];
  if (o + len > nnz_capacity) return;  // host validates capacity; never write out of bounds
  for (int p = lane; p < len; p += 32) {
    int it = indices[s + p];
    raw_items[o + p] = it;
    cols[o + p] = pos[it];
    vals[o + p] = data[s + p];
  }
}

static __global__ void k_set_nnz(const int32_t* __restrict__ row_ptr, int P, int32_t* __restrict__ counts) {
  counts[1] = row_ptr[P];
}

static __global__ void k_coo(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ cols, int row0,
                             int rows, int64_t* __restrict__ out) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int base = row_ptr[row0];
  const long long nnz = row_ptr[row0 + rows] - base;
  const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
  for (int p = s + lane; p < e; p += 32) {
    out[p - base] = r;
    out[nnz + (p - base)] = cols[p];
  }
}

// ---- slice CSC ----------------------------------------------------------------------------------------
static __global__ void k_col_count(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                   int row0, int rows, int* __restrict__ cnt) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
  for (int p = s + lane; p < e; p += 32) atomicAdd(&cnt[cols[p]], 1);
}

static __global__ void k_col_fill(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                                  const float* __restrict__ vals, int row0, int rows,
                                  const int32_t* __restrict__ csc_ptr, int* __restrict__ cursor,
                                  int32_t* __restrict__ tmp_row, int32_t* __restrict__ tmp_src) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int base = row_ptr[row0];
  const int s = row_ptr[row0 + r], e = row_ptr[row0 + r + 1];
  for (int p = s + lane; p < e; p += 32) {
    int c = cols[p];
    int slot = csc_ptr[c] + atomicAdd(&cursor[c], 1);
    tmp_row[slot] = r;
    tmp_src[slot] = p - base;
  }
}

// One warp per column: order the column's entries by row so every later reduction is deterministic.
// (row, col) pairs are unique, so the rank of an entry = number of entries of the column with a smaller row.
// Short columns: all-pairs compare through shuffles.  Long columns (popular items: up to `rows` entries): a
// per-warp row bitmap in shared memory, rank = popcount of the bits below the row (O(L + rows/32)).
static __global__ void k_col_sort(const int32_t* __restrict__ csc_ptr, int n, int rows, int words,
                                  const int32_t* __restrict__ tmp_row, const int32_t* __restrict__ tmp_src,
                                  const int32_t* __restrict__ row_ptr, int row0, const float* __restrict__ vals,
                                  int32_t* __restrict__ csc_row, float* __restrict__ csc_val,
                                  int32_t* __restrict__ csc_src) {
  extern __shared__ uint32_t s_bits[];  // [kRowWarps][2*words]: bitmap, then exclusive popcount prefix
  int c = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (c >= n) return;
  const int lane = threadIdx.x & 31;
  const float* vals_slice = vals + row_ptr[row0];
  const int s = csc_ptr[c], L = csc_ptr[c + 1] - s;
  if (L <= 32) {
    int row = (lane < L) ? tmp_row[s + lane] : 0x7fffffff;
    int src = (lane < L) ? tmp_src[s + lane] : 0;
    int rank = 0;
    for (int j = 0; j < L; ++j) rank += (__shfl_sync(0xffffffffu, row, j) < row) ? 1 : 0;
    if (lane < L) {
      csc_row[s + rank] = row;
      csc_val[s + rank] = vals_slice[src];
      if (csc_src) csc_src[s + rank] = src;
    }
  } else if (words > 0) {
    uint32_t* bm = s_bits + (size_t)(threadIdx.x >> 5) * 2 * words;
    uint32_t* pre = bm + words;
    for (int i = lane; i < words; i += 32) bm[i] = 0u;
    __syncwarp();
    for (int i = lane; i < L; i += 32) {
      const int row = tmp_row[s + i];
      atomicOr(&bm[row >> 5], 1u << (row & 31));
    }
    __syncwarp();
    int carry = 0;
    for (int b0 = 0; b0 < words; b0 += 32) {
      const int i = b0 + lane;
      const int cnt = (i < words) ? __popc(bm[i]) : 0;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (i < words) pre[i] = (uint32_t)(carry + incl - cnt);
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    for (int i = lane; i < L; i += 32) {
      const int row = tmp_row[s + i], src = tmp_src[s + i];
      const int rank = (int)pre[row >> 5] + __popc(bm[row >> 5] & ((1u << (row & 31)) - 1u));
      csc_row[s + rank] = row;
      csc_val[s + rank] = vals_slice[src];
      if (csc_src) csc_src[s + rank] = src;
    }
  } else {
    const int iters = (L + 31) / 32;  // all lanes iterate together (shuffles below are warp-wide)
    for (int it = 0; it < iters; ++it) {
      const int i = it * 32 + lane;
      int row = (i < L) ? tmp_row[s + i] : 0x7fffffff;
      int src = (i < L) ? tmp_src[s + i] : 0;
      int rank = 0;
      for (int j0 = 0; j0 < L; j0 += 32) {
        int other = (j0 + lane < L) ? tmp_row[s + j0 + lane] : 0x7fffffff;
        int lim = min(32, L - j0);
        for (int j = 0; j < lim; ++j) rank += (__shfl_sync(0xffffffffu, other, j) < row) ? 1 : 0;
      }
      if (i < L) {
        csc_row[s + rank] = row;
        csc_val[s + rank] = vals_slice[src];
        if (csc_src) csc_src[s + rank] = src;
      }
    }
  }
}

// ---- dense -> CSR (FactorizationModel.forward(input=dense), recoder/nn.py:228-235) ------------------------
static __global__ void k_dense_count(const float* __restrict__ x, int rows, int n, int ld, int* __restrict__ row_len,
                                     float* __restrict__ row_inv_norm, float* __restrict__ row_sum) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)r * ld;
  int cnt = 0;
  float sq = 0.f, sm = 0.f;
  for (int c = lane; c < n; c += 32) {
    float v = xr[c];
    cnt += (v != 0.f);
    sq += v * v;
    sm += v;
  }
  cnt = warp_sum_i(cnt);
  sq = warp_sum(sq);
  sm = warp_sum(sm);
  if (lane == 0) {
    row_len[r] = cnt;
    row_inv_norm[r] = 1.0f / fmaxf(sqrtf(sq), 1e-12f);
    row_sum[r] = sm;
  }
}

static __global__ void k_dense_fill(const float* __restrict__ x, int rows, int n, int ld,
                                    const int32_t* __restrict__ row_ptr, int nnz_capacity, int32_t* __restrict__ cols,
                                    float* __restrict__ vals) {
  int r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (size_t)r * ld;
  int o = row_ptr[r];
  if (row_ptr[r + 1] > nnz_capacity) return;
  for (int c0 = 0; c0 < n; c0 += 32) {
    int c = c0 + lane;
    float v = (c < n) ? xr[c] : 0.f;
    unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
    if (v != 0.f) {
      int k = o + __popc(m & ((1u << lane) - 1u));
      cols[k] = c;
      vals[k] = v;
    }
    o += __popc(m);
  }
}

}  // namespace rcd

using namespace rcd;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

RCD_EXPORT size_t rcd_collate_scratch_bytes(int pool_rows, int num_items) {
  size_t ints = (size_t)(pool_rows + 1) + (size_t)num_items + (size_t)(num_items + 1) +
                scan_scratch_ints(num_items > pool_rows ? num_items : pool_rows) + 64;
  return align_up(ints * sizeof(int), 256);
}

RCD_EXPORT int rcd_collate(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* users,
                           int pool_rows, int num_items, int negative_sampling, int nnz_capacity, int32_t* row_ptr,
                           int32_t* raw_items, int32_t* cols, float* vals, float* row_inv_norm, float* row_sum,
                           int32_t* pos, int64_t* items, int32_t* counts, void* scratch, size_t scratch_bytes,
                           void* stream) {
  RCD_CHECK_ARG(indptr && indices && data && users, "null CSR / users pointer");
  RCD_CHECK_ARG(pool_rows > 0 && num_items > 0, "empty pool or item space");
  RCD_CHECK_ARG(row_ptr && raw_items && cols && vals && row_inv_norm && row_sum && pos && counts, "null output");
  RCD_CHECK_ARG(!negative_sampling || items, "items output required with negative sampling");
  RCD_CHECK_ARG(scratch && scratch_bytes >= rcd_collate_scratch_bytes(pool_rows, num_items), "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  int* row_len = (int*)scratch;
  int* flag = row_len + (pool_rows + 1);
  int* rank = flag + num_items;
  int* scan_tmp = rank + (num_items + 1);
  const int P = pool_rows, I = num_items;
  const int row_blocks = rcd_div_up(P, kRowWarps);
  if (negative_sampling) RCD_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * (size_t)I, st));
  k_rows_mark<<<row_blocks, kRowWarps * 32, 0, st>>>(indptr, indices, data, users, P, row_len, row_inv_norm, row_sum,
                                                    negative_sampling ? flag : nullptr);
  RCD_LAUNCH_CHECK();
  RCD_CUDA(exclusive_scan_i32(row_len, P, row_ptr, nullptr, scan_tmp, st));
  k_set_nnz<<<1, 1, 0, st>>>(row_ptr, P, counts);
  RCD_LAUNCH_CHECK();
  if (negative_sampling) {
    RCD_CUDA(exclusive_scan_i32(flag, I, rank, counts, scan_tmp, st));
    k_compact_items<<<rcd_div_up(I, 256), 256, 0, st>>>(flag, rank, I, items, pos);
  } else {
    k_identity_items<<<rcd_div_up(I, 256), 256, 0, st>>>(I, items, pos, counts);
  }
  RCD_LAUNCH_CHECK();
  k_emit_rows<<<row_blocks, kRowWarps * 32, 0, st>>>(indptr, indices, data, users, P, row_ptr, pos, nnz_capacity,
                                                    raw_items, cols, vals);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_collate_coo(const int32_t* row_ptr, const int32_t* cols, int row0, int rows, int64_t* indices_out,
                               void* stream) {
  RCD_CHECK_ARG(row_ptr && cols && indices_out && rows > 0 && row0 >= 0, "bad slice");
  k_coo<<<rcd_div_up(rows, kRowWarps), kRowWarps * 32, 0, (cudaStream_t)stream>>>(row_ptr, cols, row0, rows,
                                                                                 indices_out);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT size_t rcd_slice_csc_scratch_bytes(int n, int nnz_slice) {
  size_t ints = (size_t)(n + 1) * 2 + (size_t)nnz_slice * 2 + scan_scratch_ints(n + 1) + 64;
  return align_up(ints * sizeof(int), 256);
}

RCD_EXPORT int rcd_slice_csc(const int32_t* row_ptr, const int32_t* cols, const float* vals, int row0, int rows,
                             int n, int32_t* csc_ptr, int32_t* csc_row, float* csc_val, int32_t* csc_src,
                             void* scratch, size_t scratch_bytes, void* stream) {
  RCD_CHECK_ARG(row_ptr && cols && vals && csc_ptr && csc_row && csc_val, "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0 && row0 >= 0, "bad slice");
  RCD_CHECK_ARG(scratch && scratch_bytes >= 256, "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: cnt[n+1] | cursor[n+1] | scan tmp | tmp_row[nnz] | tmp_src[nnz]   (nnz bounded by scratch size)
  int* cnt = (int*)scratch;
  int* cursor = cnt + (n + 1);
  int* scan_tmp = cursor + (n + 1);
  int* tmp_row = scan_tmp + scan_scratch_ints(n + 1);
  size_t used_ints = (size_t)(tmp_row - cnt);
  RCD_CHECK_ARG(scratch_bytes / sizeof(int) > used_ints, "scratch too small");
  size_t nnz_cap = (scratch_bytes / sizeof(int) - used_ints) / 2;
  int* tmp_src = tmp_row + nnz_cap;
  RCD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(n + 1) * 2, st));
  const int row_blocks = rcd_div_up(rows, kRowWarps);
  k_col_count<<<row_blocks, kRowWarps * 32, 0, st>>>(row_ptr, cols, row0, rows, cnt);
  RCD_LAUNCH_CHECK();
  RCD_CUDA(exclusive_scan_i32(cnt, n, csc_ptr, nullptr, scan_tmp, st));
  k_col_fill<<<row_blocks, kRowWarps * 32, 0, st>>>(row_ptr, cols, vals, row0, rows, csc_ptr, cursor, tmp_row,
                                                   tmp_src);
  RCD_LAUNCH_CHECK();
  int words = rcd_div_up(rows, 32);
  size_t smem = (size_t)kRowWarps * 2 * words * sizeof(uint32_t);
  if (smem > 40 * 1024) {  // very tall slices: all-pairs fallback inside the kernel
    words = 0;
    smem = 0;
  }
  k_col_sort<<<rcd_div_up(n, kRowWarps), kRowWarps * 32, smem, st>>>(csc_ptr, n, rows, words, tmp_row, tmp_src,
                                                                    row_ptr, row0, vals, csc_row, csc_val, csc_src);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT size_t rcd_dense_to_csr_scratch_bytes(int rows) {
  return align_up(((size_t)rows + 1 + scan_scratch_ints(rows) + 64) * sizeof(int), 256);
}

RCD_EXPORT int rcd_dense_to_csr(const float* dense, int rows, int n, int ld, int nnz_capacity, int32_t* row_ptr,
                                int32_t* cols, float* vals, float* row_inv_norm, float* row_sum, int32_t* nnz_out,
                                void* scratch, size_t scratch_bytes, void* stream) {
  RCD_CHECK_ARG(dense && row_ptr && cols && vals && row_inv_norm && row_sum, "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0 && ld >= n, "bad shape");
  RCD_CHECK_ARG(scratch && scratch_bytes >= rcd_dense_to_csr_scratch_bytes(rows), "scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  int* row_len = (int*)scratch;
  int* scan_tmp = row_len + rows + 1;
  const int row_blocks = rcd_div_up(rows, kRowWarps);
  k_dense_count<<<row_blocks, kRowWarps * 32, 0, st>>>(dense, rows, n, ld, row_len, row_inv_norm, row_sum);
  RCD_LAUNCH_CHECK();
  RCD_CUDA(exclusive_scan_i32(row_len, rows, row_ptr, nnz_out, scan_tmp, st));
  k_dense_fill<<<row_blocks, kRowWarps * 32, 0, st>>>(dense, rows, n, ld, row_ptr, nnz_capacity, cols, vals);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

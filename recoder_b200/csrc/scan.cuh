// Device-wide exclusive scan of int32 (three small kernels; the arrays scanned here are O(items) or
// O(pool rows), i.e. at most a few MB, so a decoupled look-back scan would buy nothing).
#pragma once
#include "common.cuh"

namespace rcd {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;  // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*>= 33 ints*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  int res = incl - v + smem[warp];
  if (total) *total = smem[32];
  __syncthreads();
  return res;
}

static __global__ void k_scan_tile_sums(const int* __restrict__ in, int N, int* __restrict__ tile_sums) {
  __shared__ int smem[33];
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < N) s += in[base + i];
  int total;
  block_exclusive_scan(s, &total, smem);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums in place, total -> *total_out (may be NULL) and out_last (may be NULL)
static __global__ void k_scan_tile_offsets(int* __restrict__ tile_sums, int nb, int* __restrict__ total_out,
                                           int* __restrict__ out_last) {
  __shared__ int smem[33];
  int carry = 0;
  for (int start = 0; start < nb; start += kScanThreads) {
    int i = start + threadIdx.x;
    int v = (i < nb) ? tile_sums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total, smem);
    if (i < nb) tile_sums[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = carry;
    if (out_last) *out_last = carry;
  }
}

static __global__ void k_scan_final(const int* __restrict__ in, int N, const int* __restrict__ tile_offsets,
                                    int* __restrict__ out) {
  __shared__ int smem[33];
  int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < N) ? in[base + i] : 0;
    s += v[i];
  }
  int ex = block_exclusive_scan(s, nullptr, smem) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < N) out[base + i] = ex;
    ex += v[i];
  }
}

static inline size_t scan_scratch_ints(int N) { return (size_t)rcd_div_up(N, kScanTile) + 8; }

// out[0..N-1] = exclusive scan of in, out[N] = total (out must hold N+1 ints); total_out optional extra copy.
static inline cudaError_t exclusive_scan_i32(const int* in, int N, int* out, int* total_out, int* scratch,
                                             cudaStream_t st) {
  int nb = rcd_div_up(N > 0 ? N : 1, kScanTile);
  k_scan_tile_sums<<<nb, kScanThreads, 0, st>>>(in, N, scratch);
  k_scan_tile_offsets<<<1, kScanThreads, 0, st>>>(scratch, nb, total_out, out + N);
  k_scan_final<<<nb, kScanThreads, 0, st>>>(in, N, scratch, out);
  g_rcd_launches += 3;
  return cudaGetLastError();
}

}  // namespace rcd

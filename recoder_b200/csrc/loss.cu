// Telemetry: squared L2 norm of a strided fp32 matrix (double accumulation).  The loss itself is computed by the
// fused decoder epilogue (decoder_tc.cu) and finished per row in sparse.cu.
#include "common.cuh"

namespace rcd {

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

RCD_EXPORT int rcd_sumsq(const float* x, long long rows, int cols, int ld, double* out_sq, void* stream) {
  RCD_CHECK_ARG(x && out_sq && rows > 0 && cols > 0 && ld >= cols, "bad arguments");
  long long total = rows * (long long)cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_sumsq<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, out_sq);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

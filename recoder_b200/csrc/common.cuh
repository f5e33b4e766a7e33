// Shared helpers for the recoder_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/recoder_b200.h"

#define RCD_EXPORT extern "C" __attribute__((visibility("default")))

// ---- error reporting: C ABI returns int status, message via rcd_last_error() -------------------
void rcd_set_error(const char* fmt, ...);

#define RCD_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      rcd_set_error("%s: invalid argument: %s", __func__, msg);    \
      return RCD_ERR_INVALID;                                      \
    }                                                              \
  } while (0)

#define RCD_CUDA(expr)                                                               \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      rcd_set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(_e));   \
      return RCD_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// every kernel launch of the library goes through this macro; the counter feeds bench.py's `gpu_launches`
extern unsigned long long g_rcd_launches;
#define RCD_LAUNCH_CHECK()             \
  do {                                 \
    ++g_rcd_launches;                  \
    RCD_CUDA(cudaGetLastError());      \
  } while (0)

static inline int rcd_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

int rcd_num_sms();

// per-rank base pointers of a buffer every rank has mapped (CUDA IPC): index = rank
struct PeerPtrs {
  void* p[RCD_MAX_PEERS];
};
int rcd_fill_peers(PeerPtrs* out, const void* const* ptrs_host, int world, const char* who);

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Activation ids shared by host and device.  recoder/nn.py:6-9 accepts any `torch.<name>`; the backward kernels keep the
// activation OUTPUT only, so the supported set is the unary torch functions whose derivative is a function of the output:
// the ones the reference's scripts, docs and tests use (tanh, sigmoid, relu) plus selu, celu, hardshrink, atan, sinh,
// asinh, expm1.  (abs, square, sin, erf, ... need the pre-activation and raise NotImplementedError in nn.py.)
__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case RCD_ACT_TANH: return tanhf(x);
    case RCD_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case RCD_ACT_RELU: return fmaxf(x, 0.0f);
    case RCD_ACT_SELU: return 1.0507009873554804934193349852946f *
                              (x > 0.f ? x : 1.6732632423543772848170429916717f * expm1f(x));
    case RCD_ACT_CELU: return x > 0.f ? x : expm1f(x);
    case RCD_ACT_HARDSHRINK: return fabsf(x) > 0.5f ? x : 0.f;
    case RCD_ACT_ATAN: return atanf(x);
    case RCD_ACT_SINH: return sinhf(x);
    case RCD_ACT_ASINH: return asinhf(x);
    case RCD_ACT_EXPM1: return expm1f(x);
    default: return x;
  }
}
// derivative expressed through the OUTPUT y = act(x)
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  switch (act) {
    case RCD_ACT_TANH: return 1.0f - y * y;
    case RCD_ACT_SIGMOID: return y * (1.0f - y);
    case RCD_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
    // selu: scale for x > 0, scale*alpha*e^x = y + scale*alpha otherwise
    case RCD_ACT_SELU: return y > 0.0f ? 1.0507009873554804934193349852946f
                                       : y + 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f;
    case RCD_ACT_CELU: return y > 0.0f ? 1.0f : y + 1.0f;          // e^x = y + 1
    case RCD_ACT_HARDSHRINK: return y != 0.0f ? 1.0f : 0.0f;
    case RCD_ACT_ATAN: { const float c = cosf(y); return c * c; }  // 1/(1+x^2) = cos^2(atan x)
    case RCD_ACT_SINH: return sqrtf(fmaf(y, y, 1.0f));             // cosh x
    case RCD_ACT_ASINH: return 1.0f / coshf(y);                    // 1/sqrt(1+x^2)
    case RCD_ACT_EXPM1: return y + 1.0f;
    default: return 1.0f;
  }
}

// fp32 sparse part of dL/dlogits at a stored target t: exact gradient minus the dense (target = 0) formula.
//   MSE      2*w*(o-t)/B - 2*o/B = 2*((w-1)*o - w*t)/B,  w = 1 + conf*[t>0]   (recoder/losses.py:44-47)
//   NLL      (p*S - t)/B - p*S/B = -t/B                                       (recoder/losses.py:69-71)
//   LOGISTIC (sigmoid(o) - t)/B - sigmoid(o)/B = -t/B
__device__ __forceinline__ float sparse_corr(int loss, float o, float t, float conf, float inv_b) {
  if (loss == RCD_LOSS_MSE) {
    const float w = 1.0f + (t > 0.f ? conf : 0.f);
    return 2.0f * inv_b * ((w - 1.0f) * o - w * t);
  }
  return -t * inv_b;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// K9: peer-memory plumbing of the data-parallel exchange (the reference has no multi-GPU code).
// One process per GPU; buffers that peers must reach (embedding tables, the gradient slab, barrier flags) are
// plain cudaMalloc allocations exported with CUDA IPC and mapped into every other rank, so kernels address peer
// HBM directly over NVLink 5 / NVSwitch with ordinary ld/st — see k_adam_p2p (optim.cu), the fused
// reduce-scatter -> Adam -> all-gather kernel.
//   rcd_p2p_alloc/free/export/open/close : IPC allocation and mapping (host calls, synchronous)
//   rcd_p2p_barrier : stream-ordered barrier of all ranks (release/acquire flags in peer memory, bounded spin)
//   rcd_p2p_reduce  : dst[i] = sum_q src_q[i] in fixed rank order (small replicated tensors: biases, loss)
#include <string.h>

#include "common.cuh"

namespace rcd {

static __global__ void k_p2p_barrier(PeerPtrs flags, int rank, int world, uint32_t seq, int32_t* __restrict__ bad,
                                     long long timeout_cycles) {
  const int q = threadIdx.x;
  if (q >= world) return;
  // everything this rank wrote before the barrier (own HBM and peer HBM) is ordered before the flag
  __threadfence_system();
  uint32_t* remote = reinterpret_cast<uint32_t*>(flags.p[q]) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(seq) : "memory");
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[rank]) + q;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - seq) >= 0) break;
    if (clock64() - t0 > timeout_cycles) {  // a peer died or fell behind by seconds: flag it instead of hanging the GPU
      atomicOr(bad, 4);
      break;
    }
    __nanosleep(100);
  }
}

static __global__ void __launch_bounds__(256)
    k_p2p_reduce(PeerPtrs src, int world, long long offset, long long count, float* __restrict__ dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < world; ++q) s += __ldcg(reinterpret_cast<const float*>(src.p[q]) + offset + i);  // fixed order
    dst[i] = s;
  }
}

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_p2p_alloc(size_t bytes, void** out_host) {
  RCD_CHECK_ARG(out_host && bytes > 0, "bad arguments");
  void* p = nullptr;
  RCD_CUDA(cudaMalloc(&p, bytes));
  RCD_CUDA(cudaMemset(p, 0, bytes));
  RCD_CUDA(cudaDeviceSynchronize());
  *out_host = p;
  return RCD_OK;
}

RCD_EXPORT int rcd_p2p_free(void* p) {
  if (p) RCD_CUDA(cudaFree(p));
  return RCD_OK;
}

RCD_EXPORT int rcd_p2p_export(const void* p, unsigned char* handle_host) {
  RCD_CHECK_ARG(p && handle_host, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == RCD_P2P_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  RCD_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
  memcpy(handle_host, &h, sizeof(h));
  return RCD_OK;
}

RCD_EXPORT int rcd_p2p_open(const unsigned char* handle_host, void** out_host) {
  RCD_CHECK_ARG(handle_host && out_host, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  RCD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out_host = p;
  return RCD_OK;
}

RCD_EXPORT int rcd_p2p_close(void* p) {
  if (p) RCD_CUDA(cudaIpcCloseMemHandle(p));
  return RCD_OK;
}

static int fill_peers(PeerPtrs* out, const void* const* ptrs_host, int world, const char* who) {
  if (!ptrs_host || world < 1 || world > RCD_MAX_PEERS) {
    rcd_set_error("%s: world size %d out of range (1..%d) or null pointer table", who, world, RCD_MAX_PEERS);
    return RCD_ERR_INVALID;
  }
  for (int q = 0; q < RCD_MAX_PEERS; ++q) out->p[q] = nullptr;
  for (int q = 0; q < world; ++q) {
    if (!ptrs_host[q]) {
      rcd_set_error("%s: null pointer for rank %d", who, q);
      return RCD_ERR_INVALID;
    }
    out->p[q] = const_cast<void*>(ptrs_host[q]);
  }
  return RCD_OK;
}

int rcd_fill_peers(PeerPtrs* out, const void* const* ptrs_host, int world, const char* who) {
  return fill_peers(out, ptrs_host, world, who);
}

RCD_EXPORT int rcd_p2p_barrier(void* const* flags_host, int rank, int world, unsigned int seq, int32_t* bad_flag,
                               double timeout_s, void* stream) {
  RCD_CHECK_ARG(bad_flag && rank >= 0 && rank < world, "bad arguments");
  PeerPtrs f;
  int rc = fill_peers(&f, flags_host, world, "rcd_p2p_barrier");
  if (rc != RCD_OK) return rc;
  const long long cycles = (long long)((timeout_s > 0 ? timeout_s : 30.0) * 1.9e9);
  k_p2p_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, (uint32_t)seq, bad_flag, cycles);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_p2p_reduce(const float* const* src_host, int world, long long offset, long long count, float* dst,
                              void* stream) {
  RCD_CHECK_ARG(dst && offset >= 0 && count > 0, "bad arguments");
  PeerPtrs s;
  int rc = fill_peers(&s, reinterpret_cast<const void* const*>(src_host), world, "rcd_p2p_reduce");
  if (rc != RCD_OK) return rc;
  const int blocks = rcd_div_up(count, 256) < 4 * rcd_num_sms() ? rcd_div_up(count, 256) : 4 * rcd_num_sms();
  k_p2p_reduce<<<blocks, 256, 0, (cudaStream_t)stream>>>(s, world, offset, count, dst);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

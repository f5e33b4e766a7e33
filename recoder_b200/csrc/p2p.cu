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
    k_p2p_reduce(PeerPtrs src, int world, long long offset, long long count, float* __restrict__ dst, int op) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float s = __ldcg(reinterpret_cast<const float*>(src.p[0]) + offset + i);
    for (int q = 1; q < world; ++q) {  // fixed rank order
      const float v = __ldcg(reinterpret_cast<const float*>(src.p[q]) + offset + i);
      s = (op == RCD_REDUCE_MAX) ? fmaxf(s, v) : s + v;
    }
    dst[i] = s;
  }
}

// Two-shot all-reduce (sum, fp32) in place over a buffer every rank has mapped: this rank reduces its 1/world slice —
// one multimem.ld_reduce per 16 bytes (summed inside the NVSwitch) or world plain loads over NVLink in rank order —
// and writes the result into every rank's copy (one multimem.st, or world plain stores).  Callers bracket it with
// rcd_p2p_barrier.  Moves count*4*(world-1)/world bytes each way per GPU; latency = two barriers + one short kernel,
// which is what the item-parallel mode needs for its [rows, H] activations (NCCL measured 0.27-1.6 ms for 17-34 MB).
static __global__ void __launch_bounds__(256)
    k_p2p_allreduce(PeerPtrs bufs, float* __restrict__ mc, long long count4, int rank, int world) {
  const long long per = (count4 + world - 1) / world;
  const long long lo = (long long)rank * per;
  const long long hi = lo + per < count4 ? lo + per : count4;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
       i += (long long)gridDim.x * blockDim.x) {
    if (mc) {
      float4 r;
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                   : "l"(mc + 4 * i)
                   : "memory");
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + 4 * i), "f"(r.x),
                   "f"(r.y), "f"(r.z), "f"(r.w)
                   : "memory");
    } else {
      float4 r = __ldcg(reinterpret_cast<const float4*>(bufs.p[0]) + i);
      for (int q = 1; q < world; ++q) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(bufs.p[q]) + i);
        r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
      }
      for (int q = 0; q < world; ++q) __stcg(reinterpret_cast<float4*>(bufs.p[q]) + i, r);
    }
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

RCD_EXPORT int rcd_p2p_allreduce(float* const* bufs_host, float* mc, long long count, int rank, int world, void* stream) {
  RCD_CHECK_ARG(count > 0 && count % 4 == 0 && rank >= 0 && rank < world, "count must be a positive multiple of 4");
  PeerPtrs b;
  int rc = fill_peers(&b, reinterpret_cast<const void* const*>(bufs_host), world, "rcd_p2p_allreduce");
  if (rc != RCD_OK) return rc;
  for (int q = 0; q < world; ++q) RCD_CHECK_ARG((reinterpret_cast<uintptr_t>(b.p[q]) & 15) == 0, "unaligned buffer");
  RCD_CHECK_ARG((reinterpret_cast<uintptr_t>(mc) & 15) == 0, "unaligned multicast address");
  const long long count4 = count / 4;
  const long long per = (count4 + world - 1) / world;
  long long blocks = (per + 255) / 256;
  const long long cap = (long long)rcd_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_p2p_allreduce<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(b, mc, count4, rank, world);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_p2p_reduce(const float* const* src_host, int world, long long offset, long long count, float* dst,
                              int op, void* stream) {
  RCD_CHECK_ARG(dst && offset >= 0 && count > 0 && (op == RCD_REDUCE_SUM || op == RCD_REDUCE_MAX), "bad arguments");
  PeerPtrs s;
  int rc = fill_peers(&s, reinterpret_cast<const void* const*>(src_host), world, "rcd_p2p_reduce");
  if (rc != RCD_OK) return rc;
  const int blocks = rcd_div_up(count, 256) < 4 * rcd_num_sms() ? rcd_div_up(count, 256) : 4 * rcd_num_sms();
  k_p2p_reduce<<<blocks, 256, 0, (cudaStream_t)stream>>>(s, world, offset, count, dst, op);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

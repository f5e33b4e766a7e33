// K8: fused optimizer steps (HBM-bound streaming kernels; 24 B/param for Adam: read+write of p, m, v).
// Arithmetic follows torch.optim's single-tensor kernels as configured by Recoder.__init_optimizer
// (recoder/model.py:101-164): Adam(lr, betas=(0.9,0.999), eps=1e-8, L2 weight decay added to the gradient),
// SGD(momentum=0.9, dampening=0), SparseAdam (touched rows only, no weight decay).
// Gradients arrive compact: row i of the table has gradient grad_rows[pos[i]] if pos[i] >= 0 else 0, so the
// scatter of embedding_dense_backward is folded into the optimizer read.
#include <stdlib.h>

#include "common.cuh"

namespace rcd {

struct AdamScalars {
  float beta2, eps, wd;
  float omb1, omb2;       // 1 - beta1, 1 - beta2 (formed in double on the host, like torch's Python scalars)
  float step_size;        // lr / (1 - beta1^t)
  float inv_bc2_sqrt;     // 1 / sqrt(1 - beta2^t)
};

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamScalars& a) {
  g = fmaf(a.wd, p, g);                                  // grad.add(param, alpha=weight_decay)
  m = fmaf(a.omb1, g - m, m);                            // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.omb2, g * g, a.beta2 * v);                  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps; // (sqrt(v) / sqrt(bc2)).add_(eps)
  p = p - a.step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// VEC = 4: one thread per float4; requires H % 4 == 0 and 16-byte aligned pointers.  p/m/v are touched exactly once
// per step and are far larger than L2: streaming (evict-first) loads and stores keep them from flushing the
// operands the GEMMs re-read.  Two independent vectors per thread are in flight.
__device__ __forceinline__ void adam_vec4(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                          size_t off, const float4& gg, const AdamScalars& a, float4 pp, float4 mm,
                                          float4 vv) {
  float4 g = gg;
  adam_update(pp.x, mm.x, vv.x, g.x, a);
  adam_update(pp.y, mm.y, vv.y, g.y, a);
  adam_update(pp.z, mm.z, vv.z, g.z, a);
  adam_update(pp.w, mm.w, vv.w, g.w, a);
  __stcs(reinterpret_cast<float4*>(p + off), pp);
  __stcs(reinterpret_cast<float4*>(m + off), mm);
  __stcs(reinterpret_cast<float4*>(v + off), vv);
}

template <int VEC>
static __global__ void __launch_bounds__(256)
    k_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, long long rows, int H,
           const float* __restrict__ grad_rows, int ldg, const int32_t* __restrict__ pos, AdamScalars a) {
  const int vpr = H / VEC;
  const long long total = rows * vpr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (VEC == 4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < total; i += 2 * stride) {
      const long long i1 = i + stride;
      const long long r0 = i / vpr, r1 = i1 / vpr;
      const int h0 = (int)(i - r0 * vpr) * 4, h1 = (int)(i1 - r1 * vpr) * 4;
      const size_t o0 = (size_t)r0 * H + h0, o1 = (size_t)r1 * H + h1;
      const long long g0 = pos ? (long long)__ldg(pos + r0) : r0, g1 = pos ? (long long)__ldg(pos + r1) : r1;
      const float4 p0 = __ldcs(reinterpret_cast<const float4*>(p + o0));
      const float4 p1 = __ldcs(reinterpret_cast<const float4*>(p + o1));
      const float4 m0 = __ldcs(reinterpret_cast<const float4*>(m + o0));
      const float4 m1 = __ldcs(reinterpret_cast<const float4*>(m + o1));
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(v + o0));
      const float4 v1 = __ldcs(reinterpret_cast<const float4*>(v + o1));
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
      if (g0 >= 0 && grad_rows) ga = __ldcs(reinterpret_cast<const float4*>(grad_rows + (size_t)g0 * ldg + h0));
      if (g1 >= 0 && grad_rows) gb = __ldcs(reinterpret_cast<const float4*>(grad_rows + (size_t)g1 * ldg + h1));
      adam_vec4(p, m, v, o0, ga, a, p0, m0, v0);
      adam_vec4(p, m, v, o1, gb, a, p1, m1, v1);
    }
    if (i < total) {
      const long long r0 = i / vpr;
      const int h0 = (int)(i - r0 * vpr) * 4;
      const size_t o0 = (size_t)r0 * H + h0;
      const long long g0 = pos ? (long long)__ldg(pos + r0) : r0;
      const float4 p0 = __ldcs(reinterpret_cast<const float4*>(p + o0));
      const float4 m0 = __ldcs(reinterpret_cast<const float4*>(m + o0));
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(v + o0));
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g0 >= 0 && grad_rows) ga = __ldcs(reinterpret_cast<const float4*>(grad_rows + (size_t)g0 * ldg + h0));
      adam_vec4(p, m, v, o0, ga, a, p0, m0, v0);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long r = i / vpr;
      const int h = (int)(i % vpr);
      const long long gr = pos ? (long long)pos[r] : r;
      const size_t off = (size_t)r * H + h;
      float pp = p[off], mm = m[off], vv = v[off];
      float gg = (gr >= 0 && grad_rows) ? grad_rows[(size_t)gr * ldg + h] : 0.f;
      adam_update(pp, mm, vv, gg, a);
      p[off] = pp;
      m[off] = mm;
      v[off] = vv;
    }
  }
}

// Fused reduce-scatter -> Adam -> all-gather over peer memory (data-parallel exchange).
// This rank owns table rows [row_begin, row_begin + nrows): for each it sums the compact gradient rows of ALL ranks
// (ld over NVLink from the peers' slabs, fixed rank order -> deterministic), applies the torch.optim.Adam update to its
// shard of p / m / v, and stores the new parameter row into every rank's replica (st over NVLink).  Replaces
// ncclAllReduce(slab) + a full-table Adam pass on every rank: the optimizer's HBM traffic drops by the world size and
// no rank ever holds a reduced copy of the slab.  grad_block_rows > 0: gradient row g was produced by exactly one
// rank, g / grad_block_rows (MF user rows), so only that peer is read.
// NVLS (NVSwitch multicast) forms: one instruction reduces the same address across every rank's copy inside the
// switch / stores to every rank's copy — per-GPU link traffic drops from (world-1) x to 1 x per element.
__device__ __forceinline__ float4 multimem_ld_reduce_add_v4(const float* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_v4(float* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

template <int VEC>
static __global__ void __launch_bounds__(256)
    k_adam_p2p(PeerPtrs tables, float* __restrict__ m, float* __restrict__ v, long long row_begin, long long nrows,
               int H, PeerPtrs grads, int ldg, const int32_t* __restrict__ pos, int grad_block_rows, int rank, int world,
               AdamScalars a, const float* __restrict__ grads_mc, float* __restrict__ table_mc) {
  const int vpr = H / VEC;
  const long long total = nrows * vpr;
  float* __restrict__ p_local = reinterpret_cast<float*>(tables.p[rank]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = row_begin + i / vpr;
    const int h = (int)(i % vpr) * VEC;
    const size_t off = (size_t)r * H + h;
    const long long gr = pos ? (long long)__ldg(pos + r) : r;
    if (VEC == 4) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr >= 0) {
        const size_t goff = (size_t)gr * ldg + h;
        if (grad_block_rows > 0) {
          g = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grads.p[gr / grad_block_rows]) + goff));
        } else if (grads_mc) {
          g = multimem_ld_reduce_add_v4(grads_mc + goff);   // summed across all ranks' slabs inside the switch
        } else {
          for (int q0 = 0; q0 < world; q0 += 4) {  // up to four peer loads in flight, summed in rank order
            float4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              t[j] = (q0 + j < world)
                         ? __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(grads.p[q0 + j]) + goff))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              g.x += t[j].x; g.y += t[j].y; g.z += t[j].z; g.w += t[j].w;
            }
          }
        }
      }
      float4 pp = __ldcs(reinterpret_cast<const float4*>(p_local + off));
      float4 mm = __ldcs(reinterpret_cast<const float4*>(m + off));
      float4 vv = __ldcs(reinterpret_cast<const float4*>(v + off));
      adam_update(pp.x, mm.x, vv.x, g.x, a);
      adam_update(pp.y, mm.y, vv.y, g.y, a);
      adam_update(pp.z, mm.z, vv.z, g.z, a);
      adam_update(pp.w, mm.w, vv.w, g.w, a);
      __stcs(reinterpret_cast<float4*>(m + off), mm);
      __stcs(reinterpret_cast<float4*>(v + off), vv);
      if (table_mc) {
        multimem_st_v4(table_mc + off, pp);   // one store, replicated to every rank's table by the switch
      } else {
        for (int q = 0; q < world; ++q)
          __stcg(reinterpret_cast<float4*>(reinterpret_cast<float*>(tables.p[q]) + off), pp);
      }
    } else {
      float g = 0.f;
      if (gr >= 0) {
        const size_t goff = (size_t)gr * ldg + h;
        if (grad_block_rows > 0) {
          g = __ldcg(reinterpret_cast<const float*>(grads.p[gr / grad_block_rows]) + goff);
        } else {
          for (int q = 0; q < world; ++q) g += __ldcg(reinterpret_cast<const float*>(grads.p[q]) + goff);
        }
      }
      float pp = p_local[off], mm = m[off], vv = v[off];
      adam_update(pp, mm, vv, g, a);
      m[off] = mm;
      v[off] = vv;
      for (int q = 0; q < world; ++q) __stcg(reinterpret_cast<float*>(tables.p[q]) + off, pp);
    }
  }
}


// ---- deferred ("lazy") dense Adam -----------------------------------------------------------------------------------
// torch.optim.Adam is dense: every row of an embedding table moves on every step, rows outside the batch with g = wd*p
// (momentum, weight decay).  That is 24 B/param of HBM traffic per step over the WHOLE table (65 % of the step's bytes at
// C3, 80 % at C5 / 512 users, the 1M-row user table of MF) for rows whose update does not depend on the batch at all.
// Those updates are DEFERRED here, exactly: `last[r]` is the step up to which row r is current; before a step reads a
// row (forward of a batch that contains it, evaluation, checkpoint) the skipped steps last[r]+1 .. T are replayed with
// the very instruction sequence of k_adam (same fp32 operations in the same order, the per-step scalars lr/(1-b1^t) and
// 1/sqrt(1-b2^t) from a device table the host fills in double precision like rcd_adam_step does) — the result is
// bit-identical to the dense kernel's, the traffic is proportional to the rows the batch touches.
//   k_adam_lazy_catchup : rows ids[0..n) (or all rows): replay steps last[r]+1 .. T with zero gradient
//   k_adam_lazy_mark    : last[r] = T for the same rows (second launch: every thread of a row must have read last[r])
//   k_adam_lazy_update  : rows ids[i] (current at T): the step T+1 with gradient row i; last[ids[i]] = T+1
struct LazyConsts {
  float beta2, eps, wd, omb1, omb2;
};

__device__ __forceinline__ void lazy_replay4(float4& pp, float4& mm, float4& vv, int t0, int T,
                                             const float2* __restrict__ scal, int scal_base, AdamScalars& a) {
  for (int t = t0 + 1; t <= T; ++t) {
    const float2 s = __ldg(scal + (t - scal_base));
    a.step_size = s.x;
    a.inv_bc2_sqrt = s.y;
    adam_update(pp.x, mm.x, vv.x, 0.f, a);
    adam_update(pp.y, mm.y, vv.y, 0.f, a);
    adam_update(pp.z, mm.z, vv.z, 0.f, a);
    adam_update(pp.w, mm.w, vv.w, 0.f, a);
  }
}

template <int VEC>
static __global__ void __launch_bounds__(256)
    k_adam_lazy_catchup(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int H,
                        const int64_t* __restrict__ ids, long long n, const int32_t* __restrict__ last, int T,
                        const float2* __restrict__ scal, int scal_base, LazyConsts c, const int32_t* __restrict__ n_dev,
                        const int32_t* __restrict__ exclude) {
  const int vpr = H / VEC;
  if (n_dev) {  // row count known on the device only (the collate of the NEXT pool has not been read back yet)
    const long long nd = (long long)__ldg(n_dev);
    n = nd < n ? nd : n;
  }
  const long long total = n * vpr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  AdamScalars a;
  a.beta2 = c.beta2; a.eps = c.eps; a.wd = c.wd; a.omb1 = c.omb1; a.omb2 = c.omb2;
  if (VEC == 4) {
    // two items per iteration: the id -> last[] -> p/m/v chain of dependent loads is latency-bound, and only the stale
    // rows (a fraction of the batch) carry any payload — the second item's loads are in flight while the first replays
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
      const long long i1 = i + stride;
      const bool has1 = i1 < total;
      const long long k0 = i / vpr, k1 = has1 ? i1 / vpr : 0;
      const long long r0 = ids ? ids[k0] : k0;
      const long long r1 = has1 ? (ids ? ids[k1] : k1) : 0;
      const int t00 = (exclude && exclude[r0] >= 0) ? T : last[r0];
      const int t01 = (has1 && !(exclude && exclude[r1] >= 0)) ? last[r1] : T;
      const bool do0 = t00 < T, do1 = t01 < T;
      const size_t off0 = (size_t)r0 * H + (size_t)(i - k0 * vpr) * 4;
      const size_t off1 = (size_t)r1 * H + (size_t)(i1 - k1 * vpr) * 4;
      float4 p0, m0, v0, p1, m1, v1;
      if (do0) {
        p0 = *reinterpret_cast<const float4*>(p + off0);
        m0 = *reinterpret_cast<const float4*>(m + off0);
        v0 = *reinterpret_cast<const float4*>(v + off0);
      }
      if (do1) {
        p1 = *reinterpret_cast<const float4*>(p + off1);
        m1 = *reinterpret_cast<const float4*>(m + off1);
        v1 = *reinterpret_cast<const float4*>(v + off1);
      }
      if (do0) {
        lazy_replay4(p0, m0, v0, t00, T, scal, scal_base, a);
        *reinterpret_cast<float4*>(p + off0) = p0;
        *reinterpret_cast<float4*>(m + off0) = m0;
        *reinterpret_cast<float4*>(v + off0) = v0;
      }
      if (do1) {
        lazy_replay4(p1, m1, v1, t01, T, scal, scal_base, a);
        *reinterpret_cast<float4*>(p + off1) = p1;
        *reinterpret_cast<float4*>(m + off1) = m1;
        *reinterpret_cast<float4*>(v + off1) = v1;
      }
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long k = i / vpr;
      const long long r = ids ? ids[k] : k;
      if (exclude && exclude[r] >= 0) continue;
      const int t0 = last[r];
      if (t0 >= T) continue;
      const size_t off = (size_t)r * H + (size_t)(i - k * vpr);
      float pp = p[off], mm = m[off], vv = v[off];
      for (int t = t0 + 1; t <= T; ++t) {
        const float2 s = __ldg(scal + (t - scal_base));
        a.step_size = s.x;
        a.inv_bc2_sqrt = s.y;
        adam_update(pp, mm, vv, 0.f, a);
      }
      p[off] = pp;
      m[off] = mm;
      v[off] = vv;
    }
  }
}

static __global__ void k_adam_lazy_mark(const int64_t* __restrict__ ids, long long n, int32_t* __restrict__ last, int T,
                                        const int32_t* __restrict__ n_dev, const int32_t* __restrict__ exclude) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) {
    const long long nd = (long long)__ldg(n_dev);
    n = nd < n ? nd : n;
  }
  if (i >= n) return;
  const long long r = ids ? ids[i] : i;
  if (exclude && exclude[r] >= 0) return;
  last[r] = T;
}

template <int VEC>
static __global__ void __launch_bounds__(256)
    k_adam_lazy_update(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int H,
                       const int64_t* __restrict__ ids, long long n, const float* __restrict__ grad_rows, int ldg,
                       int32_t* __restrict__ last, int t_new, AdamScalars a) {
  const int vpr = H / VEC;
  const long long total = n * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / vpr;
    const long long r = ids ? ids[k] : k;
    const int h = (int)(i - k * vpr) * VEC;
    const size_t off = (size_t)r * H + h;
    if (VEC == 4) {
      float4 pp = __ldcs(reinterpret_cast<const float4*>(p + off));
      float4 mm = __ldcs(reinterpret_cast<const float4*>(m + off));
      float4 vv = __ldcs(reinterpret_cast<const float4*>(v + off));
      const float4 g = __ldcs(reinterpret_cast<const float4*>(grad_rows + (size_t)k * ldg + h));
      adam_vec4(p, m, v, off, g, a, pp, mm, vv);
    } else {
      float pp = p[off], mm = m[off], vv = v[off];
      adam_update(pp, mm, vv, grad_rows[(size_t)k * ldg + h], a);
      p[off] = pp;
      m[off] = mm;
      v[off] = vv;
    }
    if (h == 0) last[r] = t_new;
  }
}

template <int VEC>
static __global__ void __launch_bounds__(256)
    k_sgd(float* __restrict__ p, float* __restrict__ buf, long long rows, int H, const float* __restrict__ grad_rows,
          int ldg, const int32_t* __restrict__ pos, float lr, float momentum, float wd) {
  const int vpr = H / VEC;
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int h = (int)(i % vpr) * VEC;
    long long gr = pos ? (long long)pos[r] : r;
    const size_t off = (size_t)r * H + h;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float pp = p[off + k], bb = buf[off + k];
      float g = (gr >= 0 && grad_rows) ? grad_rows[(size_t)gr * ldg + h + k] : 0.f;
      g = fmaf(wd, pp, g);
      bb = fmaf(momentum, bb, g);  // buf.mul_(momentum).add_(grad); zero-initialised buf == clone(grad) on step 1
      p[off + k] = pp - lr * bb;
      buf[off + k] = bb;
    }
  }
}

// torch.optim.Adagrad (lr_decay 0, initial accumulator 0, eps 1e-10) and torch.optim.RMSprop (alpha .99, eps 1e-8,
// momentum .9, not centered) as built at recoder/model.py:140-144, 150-154; dense semantics like k_adam / k_sgd.
//   KIND 0 Adagrad : sum += g*g; p -= lr * g / (sqrt(sum) + eps)
//   KIND 1 RMSprop : sq = alpha*sq + (1-alpha)*g*g; buf = momentum*buf + g / (sqrt(sq) + eps); p -= lr * buf
template <int KIND>
static __global__ void __launch_bounds__(256)
    k_opt_accum(float* __restrict__ p, float* __restrict__ s1, float* __restrict__ s2, long long rows, int H,
                const float* __restrict__ grad_rows, int ldg, const int32_t* __restrict__ pos, float lr, float alpha,
                float eps, float momentum, float wd) {
  const long long total = rows * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H;
    const int h = (int)(i - r * H);
    const long long gr = pos ? (long long)pos[r] : r;
    float pp = p[i];
    float g = (gr >= 0 && grad_rows) ? grad_rows[(size_t)gr * ldg + h] : 0.f;
    g = fmaf(wd, pp, g);
    if (KIND == 0) {
      const float sum = fmaf(g, g, s1[i]);
      s1[i] = sum;
      p[i] = pp - lr * (g / (sqrtf(sum) + eps));
    } else {
      const float sq = fmaf(1.0f - alpha, g * g, alpha * s1[i]);
      s1[i] = sq;
      const float buf = fmaf(momentum, s2[i], g / (sqrtf(sq) + eps));
      s2[i] = buf;
      p[i] = pp - lr * buf;
    }
  }
}

// torch.optim.SparseAdam on the n rows `ids` (sparse_adam functional): no weight decay, eps outside the
// bias correction: p -= lr*sqrt(bc2)/bc1 * m / (sqrt(v) + eps)
static __global__ void __launch_bounds__(256)
    k_sparse_adam(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int H,
                  const float* __restrict__ grad_rows, int ldg, const int64_t* __restrict__ ids, int n, float omb1,
                  float omb2, float eps, float step_size) {
  const long long total = (long long)n * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / H), h = (int)(i % H);
    const size_t off = (size_t)ids[r] * H + h;
    const float g = grad_rows[(size_t)r * ldg + h];
    float mm = m[off], vv = v[off];
    mm = mm + omb1 * (g - mm);                    // exp_avg.add_(make_sparse(grad.sub(m).mul_(1-beta1)))
    vv = vv + omb2 * (g * g - vv);                // exp_avg_sq.add_(make_sparse(grad^2.sub(v).mul_(1-beta2)))
    m[off] = mm;
    v[off] = vv;
    p[off] = p[off] - step_size * (mm / (sqrtf(vv) + eps));
  }
}

static __global__ void k_scatter_pos(const int64_t* __restrict__ ids, int n, int32_t* __restrict__ pos, int reset) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos[ids[i]] = reset ? -1 : i;
}

// Resident CTAs per SM of the streaming optimizer kernels (grid-stride beyond that).  Measured on C3 (profiles/README.md
// r02b): 8 x 256 threads per SM reach 6.2 TB/s alone; 4 per SM, which would leave thread slots for the tensor-core CTAs
// of the main stream, reach 5.3 TB/s alone and do NOT shorten the step (2.50 vs 2.34 ms: the two streams compete for
// HBM either way), 3 -> 2.35 ms, 2 -> 2.59 ms.  RCD_STREAM_CTAS_PER_SM overrides the default of 8.
static inline int stream_ctas_per_sm() {
  static int v = 0;
  if (v > 0) return v;
  const char* e = getenv("RCD_STREAM_CTAS_PER_SM");
  int x = e ? atoi(e) : 8;
  v = (x >= 1 && x <= 8) ? x : 8;
  return v;
}
static inline int stream_grid(long long work_items) {
  long long b = (work_items + 255) / 256;
  long long cap = (long long)rcd_num_sms() * stream_ctas_per_sm();
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace rcd

using namespace rcd;

RCD_EXPORT int rcd_adam_step(float* p, float* m, float* v, long long rows, int H, const float* grad_rows, int ldg,
                             const int32_t* pos, double lr, double beta1, double beta2, double eps,
                             double weight_decay, long long t, void* stream) {
  RCD_CHECK_ARG(p && m && v && rows > 0 && H > 0 && t >= 1, "bad arguments");
  RCD_CHECK_ARG(!grad_rows || ldg >= H, "ldg < H");
  AdamScalars a;
  a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)t);
  const double bc2 = 1.0 - pow(beta2, (double)t);
  a.step_size = (float)(lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (H % 4 == 0) && aligned16(p) && aligned16(m) && aligned16(v) &&
                   (!grad_rows || (aligned16(grad_rows) && ldg % 4 == 0));
  if (vec)
    k_adam<4><<<stream_grid(rows * (H / 4)), 256, 0, st>>>(p, m, v, rows, H, grad_rows, ldg, pos, a);
  else
    k_adam<1><<<stream_grid(rows * H), 256, 0, st>>>(p, m, v, rows, H, grad_rows, ldg, pos, a);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}


RCD_EXPORT int rcd_adam_lazy_catchup(float* p, float* m, float* v, int H, const int64_t* ids, long long n,
                                     int32_t* last, long long T, const float* scal, long long scal_base,
                                     long long scal_len, double beta1, double beta2, double eps, double weight_decay,
                                     int mark, const int32_t* n_dev, const int32_t* exclude_pos, void* stream) {
  RCD_CHECK_ARG(p && m && v && last && scal && H > 0 && n >= 0, "bad arguments");
  RCD_CHECK_ARG(T >= 0 && scal_base >= 0 && T - scal_base < scal_len, "step outside the scalar table");
  if (n == 0) return RCD_OK;
  LazyConsts c;
  c.beta2 = (float)beta2; c.eps = (float)eps; c.wd = (float)weight_decay;
  c.omb1 = (float)(1.0 - beta1); c.omb2 = (float)(1.0 - beta2);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (H % 4 == 0) && aligned16(p) && aligned16(m) && aligned16(v);
  const float2* sc = reinterpret_cast<const float2*>(scal);
  if (vec)
    k_adam_lazy_catchup<4><<<stream_grid(n * (H / 4)), 256, 0, st>>>(p, m, v, H, ids, n, last, (int)T, sc, (int)scal_base, c,
                                                                     n_dev, exclude_pos);
  else
    k_adam_lazy_catchup<1><<<stream_grid(n * H), 256, 0, st>>>(p, m, v, H, ids, n, last, (int)T, sc, (int)scal_base, c, n_dev,
                                                               exclude_pos);
  RCD_LAUNCH_CHECK();
  if (mark) {
    k_adam_lazy_mark<<<rcd_div_up(n, 256), 256, 0, st>>>(ids, n, last, (int)T, n_dev, exclude_pos);
    RCD_LAUNCH_CHECK();
  }
  return RCD_OK;
}

RCD_EXPORT int rcd_adam_lazy_update(float* p, float* m, float* v, int H, const int64_t* ids, long long n,
                                    const float* grad_rows, int ldg, int32_t* last, double lr, double beta1,
                                    double beta2, double eps, double weight_decay, long long t, void* stream) {
  RCD_CHECK_ARG(p && m && v && last && grad_rows && H > 0 && n >= 0 && t >= 1 && ldg >= H, "bad arguments");
  if (n == 0) return RCD_OK;
  AdamScalars a;
  a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)t);
  const double bc2 = 1.0 - pow(beta2, (double)t);
  a.step_size = (float)(lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (H % 4 == 0) && (ldg % 4 == 0) && aligned16(p) && aligned16(m) && aligned16(v) && aligned16(grad_rows);
  if (vec)
    k_adam_lazy_update<4><<<stream_grid(n * (H / 4)), 256, 0, st>>>(p, m, v, H, ids, n, grad_rows, ldg, last, (int)t, a);
  else
    k_adam_lazy_update<1><<<stream_grid(n * H), 256, 0, st>>>(p, m, v, H, ids, n, grad_rows, ldg, last, (int)t, a);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

/* host helper: the per-step scalars of torch.optim.Adam exactly as rcd_adam_step forms them (double -> float) */
RCD_EXPORT int rcd_adam_scalars(double lr, double beta1, double beta2, long long t_first, int count, float* out_host) {
  RCD_CHECK_ARG(out_host && count > 0 && t_first >= 1, "bad arguments");
  for (int i = 0; i < count; ++i) {
    const double t = (double)(t_first + i);
    out_host[2 * i] = (float)(lr / (1.0 - pow(beta1, t)));
    out_host[2 * i + 1] = (float)(1.0 / sqrt(1.0 - pow(beta2, t)));
  }
  return RCD_OK;
}

RCD_EXPORT int rcd_adam_step_p2p(float* const* tables_host, float* m, float* v, long long row_begin, long long row_end,
                                 int H, const float* const* grads_host, int ldg, const int32_t* pos, int grad_block_rows,
                                 int rank, int world, double lr, double beta1, double beta2, double eps,
                                 double weight_decay, long long t, const float* grads_mc, float* table_mc,
                                 void* stream) {
  RCD_CHECK_ARG(m && v && H > 0 && t >= 1 && row_begin >= 0 && row_end >= row_begin, "bad arguments");
  RCD_CHECK_ARG(rank >= 0 && rank < world && ldg >= H && grad_block_rows >= 0, "bad arguments");
  PeerPtrs tabs, grads;
  int rc = rcd_fill_peers(&tabs, reinterpret_cast<const void* const*>(tables_host), world, "rcd_adam_step_p2p");
  if (rc != RCD_OK) return rc;
  rc = rcd_fill_peers(&grads, reinterpret_cast<const void* const*>(grads_host), world, "rcd_adam_step_p2p");
  if (rc != RCD_OK) return rc;
  const long long nrows = row_end - row_begin;
  if (nrows == 0) return RCD_OK;
  AdamScalars a;
  a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)t);
  const double bc2 = 1.0 - pow(beta2, (double)t);
  a.step_size = (float)(lr / bc1);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  bool vec = (H % 4 == 0) && (ldg % 4 == 0) && aligned16(m) && aligned16(v);
  for (int q = 0; q < world; ++q) vec = vec && aligned16(tabs.p[q]) && aligned16(grads.p[q]);
  vec = vec && aligned16(grads_mc) && aligned16(table_mc);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    k_adam_p2p<4><<<stream_grid(nrows * (H / 4)), 256, 0, st>>>(tabs, m, v, row_begin, nrows, H, grads, ldg, pos,
                                                               grad_block_rows, rank, world, a, grads_mc, table_mc);
  else  // scalar fallback: unicast only
    k_adam_p2p<1><<<stream_grid(nrows * H), 256, 0, st>>>(tabs, m, v, row_begin, nrows, H, grads, ldg, pos,
                                                         grad_block_rows, rank, world, a, nullptr, nullptr);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_sgd_step(float* p, float* buf, long long rows, int H, const float* grad_rows, int ldg,
                            const int32_t* pos, double lr, double momentum, double weight_decay,
                            void* stream) {
  RCD_CHECK_ARG(p && buf && rows > 0 && H > 0, "bad arguments");
  RCD_CHECK_ARG(!grad_rows || ldg >= H, "ldg < H");
  cudaStream_t st = (cudaStream_t)stream;
  if (H % 4 == 0)
    k_sgd<4><<<stream_grid(rows * (H / 4)), 256, 0, st>>>(p, buf, rows, H, grad_rows, ldg, pos, (float)lr,
                                                         (float)momentum, (float)weight_decay);
  else
    k_sgd<1><<<stream_grid(rows * H), 256, 0, st>>>(p, buf, rows, H, grad_rows, ldg, pos, (float)lr, (float)momentum,
                                                   (float)weight_decay);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_adagrad_step(float* p, float* sum, long long rows, int H, const float* grad_rows, int ldg,
                                const int32_t* pos, double lr, double eps, double weight_decay, void* stream) {
  RCD_CHECK_ARG(p && sum && rows > 0 && H > 0, "bad arguments");
  RCD_CHECK_ARG(!grad_rows || ldg >= H, "ldg < H");
  k_opt_accum<0><<<stream_grid(rows * H), 256, 0, (cudaStream_t)stream>>>(p, sum, nullptr, rows, H, grad_rows, ldg, pos,
                                                                         (float)lr, 0.f, (float)eps, 0.f,
                                                                         (float)weight_decay);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_rmsprop_step(float* p, float* square_avg, float* buf, long long rows, int H, const float* grad_rows,
                                int ldg, const int32_t* pos, double lr, double alpha, double eps, double momentum,
                                double weight_decay, void* stream) {
  RCD_CHECK_ARG(p && square_avg && buf && rows > 0 && H > 0, "bad arguments");
  RCD_CHECK_ARG(!grad_rows || ldg >= H, "ldg < H");
  k_opt_accum<1><<<stream_grid(rows * H), 256, 0, (cudaStream_t)stream>>>(p, square_avg, buf, rows, H, grad_rows, ldg,
                                                                         pos, (float)lr, (float)alpha, (float)eps,
                                                                         (float)momentum, (float)weight_decay);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_sparse_adam_step(float* p, float* m, float* v, int H, const float* grad_rows, int ldg,
                                    const int64_t* ids, int n, double lr, double beta1, double beta2, double eps,
                                    long long t, void* stream) {
  RCD_CHECK_ARG(p && m && v && grad_rows && ids && n > 0 && H > 0 && t >= 1 && ldg >= H, "bad arguments");
  const double bc1 = 1.0 - pow(beta1, (double)t);
  const double bc2 = 1.0 - pow(beta2, (double)t);
  const float step_size = (float)(lr * sqrt(bc2) / bc1);
  k_sparse_adam<<<stream_grid((long long)n * H), 256, 0, (cudaStream_t)stream>>>(p, m, v, H, grad_rows, ldg, ids, n,
                                                                                (float)(1.0 - beta1),
                                                                                (float)(1.0 - beta2), (float)eps,
                                                                                step_size);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

RCD_EXPORT int rcd_scatter_pos(const int64_t* ids, int n, int32_t* pos, int reset, void* stream) {
  RCD_CHECK_ARG(ids && pos && n > 0, "bad arguments");
  k_scatter_pos<<<rcd_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(ids, n, pos, reset);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

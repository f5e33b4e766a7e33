// K4+K5 fused: decoder forward GEMM with the loss / dL/dlogits epilogue, tcgen05 + TMEM + TMA, sm_100a.
//
//   acc[r,c] = sum_h Zb[r,h] * Wg[c,h]          (bf16 operands, fp32 accumulate in TMEM), o = acc + bias[c]
//   G[r,c]   = NLL      exp(o - ref[r])          (unnormalised softmax numerator; dL/dO = alpha[r]*G - t/B)
//              MSE      2*o/B                     (dense, target-free part of dL/dO)
//              LOGISTIC sigmoid(o)/B
//   stat[r, 2*n_tile+group] = row partial of      NLL: sum G   MSE: sum o^2   LOGISTIC: sum softplus(o)
//
// The logits never go to memory: this replaces F.linear (recoder/nn.py:280, :361), the loss modules
// (recoder/losses.py:43-47, 68-71; BCEWithLogitsLoss, recoder/model.py:91) and the first node of their backward.
// For the multinomial NLL the usual two passes (row max/sum, then softmax) collapse into one because any per-row
// reference value `ref` gives exp(o-ref)/sum exp(o-ref) = softmax; the caller passes the largest logit among the
// row's own positives (rcd_sddmm), so the row sum is >= 1.  That reference is not an upper bound of the row: a
// non-target logit far above it would overflow exp().  The exponent is therefore clamped at 2^kNllClampLog2 (G and
// the row sums stay finite whatever the logits are), a row whose sum reaches that value is flagged by
// rcd_loss_finish, and the flagged step is redone on the device with the TRUE row maxima as reference: MAXMODE
// instantiation below (same GEMM, the epilogue reduces max instead of exp/sum and stores nothing) ->
// rcd_nll_ref_fix -> this kernel again.  All three extra launches return at once while the flag is clear, which
// is every step of a sane model: F.log_softmax's unconditional stability (recoder/losses.py:69) for a few
// microseconds of empty launches.
//
// One persistent CTA per SM, 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (tcgen05.mma cta_group::1,
// M=128, N=256, K=16) + TMEM allocator, warps 2-9 = epilogue.  Tile 128 rows x 256 items, K = H in 64-wide blocks
// through a 4-stage smem ring (192 KB of operands in flight: the MMAs wait for operand latency, not for the tensor pipe —
// 3 stages measured 0.268 ms for the C3 forward, 4 stages 0.231 ms); accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Epilogue warp (q = warp%4 -> TMEM lanes 32q.., group = (warp-2)/4 -> columns 128*group..):
// tcgen05.ld 32x32b (lane = row) -> math -> bf16 -> 128B-swizzled smem box [32 rows x 64 cols] -> TMA store
// (full 128-byte lines to HBM instead of row-per-thread 16-byte stores).
#include <cuda.h>
#include <stdlib.h>

#include "gemm_internal.cuh"
#include "tc_ptx.cuh"

#ifndef RCD_DEC_STAGES
#define RCD_DEC_STAGES 4
#endif

namespace rcd {

constexpr int kDecStages = RCD_DEC_STAGES;   // 3 (double-buffered staging boxes) or 4 (single: +48 KB of operands in flight)
// 8 epilogue warps.  Measured (profiles/README.md r02g): 16 warps (4 per scheduler, 64 columns each) are SLOWER — fused
// forward 0.291 ms against 0.268 at C3 — so the epilogue is not what the MMAs wait for; bytes in flight per SM are.
constexpr int kDecEpiWarps = 8;
constexpr int kDecColGroups = kDecEpiWarps / 4;             // column groups of a tile: warp e -> lanes 32*(warp%4), group e/4
constexpr int kDecThreads = 64 + 32 * kDecEpiWarps;         // 320
constexpr int kDecTileN = 256;
constexpr int kDecAStage = kTileM * kTileK * 2;       // 16 KB
constexpr int kDecBStage = kDecTileN * kTileK * 2;    // 32 KB
constexpr int kDecStage = kDecAStage + kDecBStage;
constexpr int kDecBoxBytes = 32 * 64 * 2;             // staging box: 32 rows x 64 bf16 columns
constexpr int kDecColsPerWarp = 256 / kDecColGroups;         // 128
constexpr int kDecBoxes = kDecColsPerWarp / 64;             // staging boxes per warp and tile
constexpr int kDecStageBufs = (RCD_DEC_STAGES >= 4) ? 1 : 2;
constexpr int kDecStaging = kDecEpiWarps * kDecStageBufs * kDecBoxBytes;
constexpr int kDecBiasBytes = 2 * kDecTileN * 4;
constexpr int kDecSmem = kDecStages * kDecStage + kDecStaging + kDecBiasBytes + 256;   // aligned below: no slack
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNllClampLog2 = 64.0f;   // == RCD_NLL_CLAMP_LOG2 (rcd_loss_finish redoes rows with sum >= 2^64)

struct DecFusedParams {
  int M, N;             // rows, items
  float inv_b;
  const float* bias;    // [N]
  const float* row_ref; // [M] or nullptr (NLL only)
  float* stat;          // [M, stat_ld]
  int stat_ld;
  const int32_t* cond;  // device flag: the kernel returns at once when *cond == 0 (nullptr: always run)
};

// 32 accumulator values of one row -> 32 outputs (packed bf16x2) + row partial.  MASK: columns >= n_valid are dead.
template <int LOSS, bool MASK, bool MAXMODE>
__device__ __forceinline__ void dec_chunk32(const float (&v)[32], const float* __restrict__ bias_s, float m2,
                                            float scale, int n_valid, uint32_t (&packed)[16], float& racc) {
  if (MAXMODE) {  // row maximum of the logits in log2 units (o * log2 e); nothing is stored
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float o2 = fmaf(v[i], kLog2e, bias_s[i]);
      if (!MASK || i < n_valid) racc = fmaxf(racc, o2);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias_s + i);  // warp-wide broadcast read
    const float bb[4] = {b.x, b.y, b.z, b.w};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = v[i + k];
      float out, part;
      if (LOSS == RCD_LOSS_NLL) {
        // bias and ref arrive pre-multiplied by log2(e); clamped: see the header comment
        out = ex2_approx(fminf(fmaf(a, kLog2e, bb[k]) - m2, kNllClampLog2));
        part = out;
      } else if (LOSS == RCD_LOSS_MSE) {
        const float o = a + bb[k];
        out = o * scale;  // 2/B
        part = o * o;
      } else {
        const float o = a + bb[k];
        const float e = ex2_approx(-fabsf(o) * kLog2e);  // exp(-|o|) in (0, 1]
        const float r = rcp_approx(1.0f + e);
        out = (o >= 0.f ? r : e * r) * scale;            // sigmoid(o)/B
        part = fmaxf(o, 0.f) + kLn2 * lg2_approx(1.0f + e);
      }
      if (MASK && i + k >= n_valid) {
        out = 0.f;
        part = 0.f;
      }
      g[k] = out;
      racc += part;
    }
    packed[i >> 1] = pack_bf16x2(g[0], g[1]);
    packed[(i >> 1) + 1] = pack_bf16x2(g[2], g[3]);
  }
}

template <int LOSS, bool MAXMODE>
static __global__ void __launch_bounds__(kDecThreads, 1)
    k_decoder_fused(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmG, DecFusedParams p, int m_tiles, int n_tiles, int kblocks,
                    uint32_t idesc) {
  if (p.cond != nullptr && __ldg(p.cond) == 0) return;  // uniform over the grid: nobody has touched a barrier yet
  extern __shared__ __align__(1024) uint8_t smem_pair_raw[];
  const uint32_t raw_addr = smem_u32(smem_pair_raw);
  if ((raw_addr & 1023u) != 0) {   // 128B-swizzle atoms need 1024 B alignment; the budget has no room for slack
    if (threadIdx.x == 0) printf("recoder_b200: dynamic shared memory is not 1024-byte aligned (0x%x)\n", raw_addr);
    __trap();
  }
  const uint32_t tiles = raw_addr;
  uint8_t* smem = smem_pair_raw;
  const uint32_t staging = tiles + kDecStages * kDecStage;
  float* bias_s = reinterpret_cast<float*>(smem + kDecStages * kDecStage + kDecStaging);
  const uint32_t bars = staging + kDecStaging + kDecBiasBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kDecStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kDecStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kDecStages + 2 + a); };
  uint32_t* tmem_slot =
      reinterpret_cast<uint32_t*>(smem + kDecStages * kDecStage + kDecStaging + kDecBiasBytes + 8 * (2 * kDecStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kDecStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), kDecEpiWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = m_tiles * n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int mt = unit % m_tiles, nt = unit / m_tiles;  // m fastest: concurrent CTAs share the Wg tile in L2
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 10);
          mbar_arrive_expect_tx(full_bar(stage), (uint32_t)kDecStage);
          const uint32_t a_dst = tiles + stage * kDecStage;
          tma_load_2d(a_dst, &tmA, full_bar(stage), kb * kTileK, mt * kTileM);
          tma_load_2d(a_dst + kDecAStage, &tmB, full_bar(stage), kb * kTileK, nt * kDecTileN);
          if (++stage == kDecStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 11);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kDecTileN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase, 12);
          tc_fence_after();
          const uint32_t a_addr = tiles + stage * kDecStage, b_addr = a_addr + kDecAStage;
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k)
            tc_mma_bf16(tmem_d, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(b_addr + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (++stage == kDecStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull_bar(acc));
      }
    }
  } else {
    // ---------------- epilogue warps 2..9 ----------------
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const int colg = e >> 2;  // column group of the tile
    const uint32_t my_staging = staging + (uint32_t)(e * kDecStageBufs * kDecBoxBytes);
    const int et = threadIdx.x - 64;  // 0..32*kDecEpiWarps-1
    const float scale = (LOSS == RCD_LOSS_MSE) ? 2.0f * p.inv_b : p.inv_b;
    int sbuf = 0;
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const int mt = unit % m_tiles, nt = unit / m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int n0 = nt * kDecTileN;
      if (et < kDecTileN) {
        const int c = n0 + et;
        float b = (c < p.N) ? __ldg(p.bias + c) : 0.f;
        if (LOSS == RCD_LOSS_NLL) b *= kLog2e;
        bias_s[acc * kDecTileN + et] = b;
      }
      named_bar_sync(1, kDecEpiWarps * 32);
      const int row = mt * kTileM + q * 32 + lane;
      float m2 = 0.f;
      if (LOSS == RCD_LOSS_NLL && p.row_ref && row < p.M) m2 = __ldg(p.row_ref + row) * kLog2e;
      mbar_wait(tfull_bar(acc), acc_phase, 13);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kDecTileN + colg * kDecColsPerWarp);
      float racc = MAXMODE ? -INFINITY : 0.f;
#pragma unroll 1
      for (int box = 0; box < kDecBoxes; ++box) {
        const int col0 = n0 + colg * kDecColsPerWarp + box * 64;
        if (col0 >= p.N) break;  // warp-uniform
        const uint32_t sdst = my_staging + (uint32_t)(sbuf * kDecBoxBytes);
        if (!MAXMODE) {
          if (lane == 0) tma_store_wait_read<kDecStageBufs - 1>();  // the store that last used this buffer has read it
          __syncwarp();
        }
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          float v[32];
          tc_ld_32x32(taddr + (uint32_t)(box * 64 + c32 * 32), v);
          uint32_t packed[16];
          const float* bs = bias_s + acc * kDecTileN + colg * kDecColsPerWarp + box * 64 + c32 * 32;
          const int n_valid = p.N - (col0 + c32 * 32);
          if (n_valid >= 32) dec_chunk32<LOSS, false, MAXMODE>(v, bs, m2, scale, 32, packed, racc);
          else dec_chunk32<LOSS, true, MAXMODE>(v, bs, m2, scale, n_valid, packed, racc);
          if (MAXMODE) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t chunk = (uint32_t)(c32 * 4 + j);
            const uint32_t addr = sdst + (uint32_t)lane * 128u + ((chunk ^ ((uint32_t)lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[j * 4]),
                         "r"(packed[j * 4 + 1]), "r"(packed[j * 4 + 2]), "r"(packed[j * 4 + 3])
                         : "memory");
          }
        }
        if (MAXMODE) continue;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmG, sdst, col0, mt * kTileM + q * 32);
          tma_store_commit();
        }
        if (kDecStageBufs > 1) sbuf ^= 1;
      }
      if (row < p.M) p.stat[(size_t)row * p.stat_ld + nt * kDecColGroups + colg] = racc;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (!MAXMODE && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- CTA-pair variant (cta_group::2) ------------------------------------------------------------------------------------
// Two CTAs on the two SMs of a TPC compute a 256-row x 256-item tile with ONE M=256 MMA per k-step: each CTA stages its
// own 128 rows of Zb and only HALF of the Wg tile (128 items); the tensor cores read the other half from the partner's
// shared memory.  Per 128x256x64 MMA a CTA therefore loads 32 KB instead of 48 KB — the single-CTA kernel runs at the
// L2->SM return bandwidth (2.80 GB per launch for a 0.24 GB operand set, tensor pipe 44 % active: profiles r02e), not at
// the tensor pipe.  Accumulators stay double-buffered in each CTA's TMEM (its 128 lanes x 2 x 256 columns), so the loss
// epilogue of tile i still overlaps the MMAs of tile i+1.  Protocol: both producers' TMA loads complete on the LEADER's
// full barrier (expect_tx = both CTAs' bytes); the leader's MMA lane issues tcgen05.mma.cta_group::2 and releases stages
// / publishes accumulators with multicast commits that arrive in BOTH CTAs; both CTAs' epilogue warps arrive on the
// leader's accumulator-empty barrier.  5-stage ring of 32 KB: 160 KB in flight per SM for two thirds of the bytes per MMA
// (single-CTA kernel: 144 KB) — the fused forward is bound by operand bytes in flight, not by the tensor pipe or the
// epilogue (r02f / r02g).
constexpr int kPairStages = (RCD_DEC_STAGES >= 4) ? 6 : 5;
constexpr int kPairAStage = kTileM * kTileK * 2;            // 16 KB: this CTA's 128 rows
constexpr int kPairBStage = (kDecTileN / 2) * kTileK * 2;   // 16 KB: this CTA's 128 items
constexpr int kPairStage = kPairAStage + kPairBStage;
constexpr int kPairSmem = kPairStages * kPairStage + kDecStaging + kDecBiasBytes + 256;  // no slack: aligned below

template <int LOSS, bool MAXMODE>
static __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDecThreads, 1)
    k_decoder_fused_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmG, DecFusedParams p, int m_tiles, int n_tiles, int kblocks,
                    uint32_t idesc) {
  if (p.cond != nullptr && __ldg(p.cond) == 0) return;  // uniform over the grid: nobody has touched a barrier yet
  extern __shared__ __align__(1024) uint8_t smem_pair_raw[];
  const uint32_t raw_addr = smem_u32(smem_pair_raw);
  if ((raw_addr & 1023u) != 0) {   // 128B-swizzle atoms need 1024 B alignment; the budget has no room for slack
    if (threadIdx.x == 0) printf("recoder_b200: dynamic shared memory is not 1024-byte aligned (0x%x)\n", raw_addr);
    __trap();
  }
  const uint32_t tiles = raw_addr;
  uint8_t* smem = smem_pair_raw;
  const uint32_t staging = tiles + kPairStages * kPairStage;
  float* bias_s = reinterpret_cast<float*>(smem + kPairStages * kPairStage + kDecStaging);
  const uint32_t bars = staging + kDecStaging + kDecBiasBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kPairStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kPairStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kPairStages + 2 + a); };
  uint32_t* tmem_slot =
      reinterpret_cast<uint32_t*>(smem + kPairStages * kPairStage + kDecStaging + kDecBiasBytes + 8 * (2 * kPairStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();   // 0 = leader: owns the operand-ready barriers and issues the MMAs
  const bool leader = cta_rank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmG);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kPairStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 2 * kDecEpiWarps);   // the epilogue warps of BOTH CTAs release the leader's accumulator
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised and TMEM is allocated before anyone signals across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = m_tiles * n_tiles;           // m_tiles counts 256-row PAIR tiles
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = pair_id; unit < units; unit += num_pairs) {
        const int mt = unit % m_tiles, nt = unit / m_tiles;  // m fastest: concurrent pairs share the Wg tile in L2
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 10);
          // each CTA loads ITS 128 rows of A and ITS half (128 items) of B; all four boxes count on the leader's barrier
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * (uint32_t)kPairStage);
          const uint32_t lead_bar = mapa_shared(full_bar(stage), 0);
          const uint32_t a_dst = tiles + stage * kPairStage;
          tma_load_2d_pair(a_dst, &tmA, lead_bar, kb * kTileK, mt * 2 * kTileM + (int)cta_rank * kTileM);
          tma_load_2d_pair(a_dst + kPairAStage, &tmB, lead_bar, kb * kTileK,
                           nt * kDecTileN + (int)cta_rank * (kDecTileN / 2));
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int unit = pair_id; unit < units; unit += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 11);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kDecTileN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase, 12);
          tc_fence_after();
          const uint32_t a_addr = tiles + stage * kPairStage, b_addr = a_addr + kPairAStage;
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k)
            tc_mma_bf16_pair(tmem_d, make_smem_desc(a_addr + k * 32, 16, 1024),
                             make_smem_desc(b_addr + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit_pair(empty_bar(stage));   // the slot is free in BOTH CTAs once these MMAs retire
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit_pair(tfull_bar(acc));   // both CTAs' epilogues read their own 128 lanes
      }
    }
  } else {
    // ---------------- epilogue warps 2..9 ----------------
    const int e = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const int colg = e >> 2;  // column group of the tile
    const uint32_t my_staging = staging + (uint32_t)(e * kDecStageBufs * kDecBoxBytes);
    const int et = threadIdx.x - 64;  // 0..32*kDecEpiWarps-1
    const float scale = (LOSS == RCD_LOSS_MSE) ? 2.0f * p.inv_b : p.inv_b;
    int sbuf = 0;
    int it = 0;
    const uint32_t lead_tempty[2] = {mapa_shared(tempty_bar(0), 0), mapa_shared(tempty_bar(1), 0)};
    for (int unit = pair_id; unit < units; unit += num_pairs, ++it) {
      const int mt = unit % m_tiles, nt = unit / m_tiles;
      const int row_base = mt * 2 * kTileM + (int)cta_rank * kTileM;   // this CTA's 128 rows of the pair tile
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int n0 = nt * kDecTileN;
      if (et < kDecTileN) {
        const int c = n0 + et;
        float b = (c < p.N) ? __ldg(p.bias + c) : 0.f;
        if (LOSS == RCD_LOSS_NLL) b *= kLog2e;
        bias_s[acc * kDecTileN + et] = b;
      }
      named_bar_sync(1, kDecEpiWarps * 32);
      const int row = row_base + q * 32 + lane;
      float m2 = 0.f;
      if (LOSS == RCD_LOSS_NLL && p.row_ref && row < p.M) m2 = __ldg(p.row_ref + row) * kLog2e;
      mbar_wait(tfull_bar(acc), acc_phase, 13);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kDecTileN + colg * kDecColsPerWarp);
      float racc = MAXMODE ? -INFINITY : 0.f;
#pragma unroll 1
      for (int box = 0; box < kDecBoxes; ++box) {
        const int col0 = n0 + colg * kDecColsPerWarp + box * 64;
        if (col0 >= p.N) break;  // warp-uniform
        const uint32_t sdst = my_staging + (uint32_t)(sbuf * kDecBoxBytes);
        if (!MAXMODE) {
          if (lane == 0) tma_store_wait_read<kDecStageBufs - 1>();  // the store that last used this buffer has read it
          __syncwarp();
        }
#pragma unroll
        for (int c32 = 0; c32 < 2; ++c32) {
          float v[32];
          tc_ld_32x32(taddr + (uint32_t)(box * 64 + c32 * 32), v);
          uint32_t packed[16];
          const float* bs = bias_s + acc * kDecTileN + colg * kDecColsPerWarp + box * 64 + c32 * 32;
          const int n_valid = p.N - (col0 + c32 * 32);
          if (n_valid >= 32) dec_chunk32<LOSS, false, MAXMODE>(v, bs, m2, scale, 32, packed, racc);
          else dec_chunk32<LOSS, true, MAXMODE>(v, bs, m2, scale, n_valid, packed, racc);
          if (MAXMODE) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t chunk = (uint32_t)(c32 * 4 + j);
            const uint32_t addr = sdst + (uint32_t)lane * 128u + ((chunk ^ ((uint32_t)lane & 7u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[j * 4]),
                         "r"(packed[j * 4 + 1]), "r"(packed[j * 4 + 2]), "r"(packed[j * 4 + 3])
                         : "memory");
          }
        }
        if (MAXMODE) continue;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmG, sdst, col0, row_base + q * 32);
          tma_store_commit();
        }
        if (kDecStageBufs > 1) sbuf ^= 1;
      }
      if (row < p.M) p.stat[(size_t)row * p.stat_ld + nt * kDecColGroups + colg] = racc;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lead_tempty[acc]);
    }
    if (!MAXMODE && lane == 0) tma_store_wait_all<0>();
  }

  // neither CTA may leave (or free TMEM) while the other still reads its shared memory / signals its barriers
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace rcd

using namespace rcd;

// CTA-pair (cta_group::2) variant of the fused kernel for slices of more than 128 rows; RCD_GEMM_PAIR=0 switches it off.
// Measured (profiles/README.md r02k): C3 forward 0.225 ms against 0.232 ms single-CTA, C5 / 8192 users 4.28 against 4.57 ms.
static bool decoder_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RCD_GEMM_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

RCD_EXPORT int rcd_decoder_stat_cols(int n) { return kDecColGroups * rcd_div_up(n > 0 ? n : 1, kDecTileN); }

RCD_EXPORT int rcd_decoder_fwd_loss(const uint16_t* Zb, int ldzb, const uint16_t* Wg, int ldw, const float* bias,
                                    int rows, int n, int H, int loss, float inv_b, const float* row_ref, uint16_t* G,
                                    int ldg, float* stat, int stat_ld, int mode, const int32_t* cond, void* stream) {
  RCD_CHECK_ARG(Zb && Wg && bias && G && stat, "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0 && H > 0, "bad shape");
  RCD_CHECK_ARG(ldzb >= H && ldw >= H && ldg >= n && ldg % 8 == 0, "bad leading dimension");
  RCD_CHECK_ARG(stat_ld >= rcd_decoder_stat_cols(n), "stat_ld too small");
  RCD_CHECK_ARG(mode == RCD_DEC_MODE_LOSS || (mode == RCD_DEC_MODE_ROWMAX && loss == RCD_LOSS_NLL),
                "mode must be RCD_DEC_MODE_LOSS, or RCD_DEC_MODE_ROWMAX with the multinomial NLL");
  CUtensorMap tmA, tmB, tmG;
  int rc = encode_map(&tmA, Zb, H, rows, ldzb, kTileK, kTileM);
  if (rc != RCD_OK) return rc;
  rc = encode_map(&tmB, Wg, H, n, ldw, kTileK, kDecTileN);
  if (rc != RCD_OK) return rc;
  rc = encode_map(&tmG, G, n, rows, ldg, 64, 32);
  if (rc != RCD_OK) return rc;
  const int m_tiles = rcd_div_up(rows, kTileM), n_tiles = rcd_div_up(n, kDecTileN), kblocks = rcd_div_up(H, kTileK);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kDecTileN >> 3) << 17) |
                         ((uint32_t)(kTileM >> 4) << 24);
  DecFusedParams p{};
  p.M = rows; p.N = n; p.inv_b = inv_b; p.bias = bias; p.row_ref = row_ref; p.stat = stat; p.stat_ld = stat_ld;
  p.cond = cond;
  cudaStream_t st = (cudaStream_t)stream;
  if (decoder_pair_enabled() && rows > kTileM) {
    // CTA pairs: 256-row tiles, each CTA of a pair stages half of the Wg tile
    CUtensorMap tmBh;
    rc = encode_map(&tmBh, Wg, H, n, ldw, kTileK, kDecTileN / 2);
    if (rc != RCD_OK) return rc;
    const int m_tiles2 = rcd_div_up(rows, 2 * kTileM);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kDecTileN >> 3) << 17) |
                            ((uint32_t)((2 * kTileM) >> 4) << 24);
    const int units2 = m_tiles2 * n_tiles;
    static int max_pairs = 0;
    static bool pair_attr[4] = {false, false, false, false};
#define RCD_DEC_PAIR_LAUNCH(L, MAXM, SLOT)                                                                          \
  do {                                                                                                              \
    if (!pair_attr[SLOT]) {                                                                                         \
      RCD_CUDA(cudaFuncSetAttribute(k_decoder_fused_pair<L, MAXM>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                    kPairSmem));                                                                    \
      pair_attr[SLOT] = true;                                                                                       \
    }                                                                                                               \
    if (max_pairs == 0) {                                                                                           \
      cudaLaunchConfig_t cfg = {};                                                                                  \
      cfg.gridDim = dim3(rcd_num_sms() & ~1);                                                                       \
      cfg.blockDim = dim3(kDecThreads);                                                                             \
      cfg.dynamicSmemBytes = kPairSmem;                                                                             \
      int nc = 0;                                                                                                   \
      if (cudaOccupancyMaxActiveClusters(&nc, k_decoder_fused_pair<L, MAXM>, &cfg) != cudaSuccess || nc <= 0)       \
        nc = (rcd_num_sms() / 2) - 2;                                                                               \
      max_pairs = nc;                                                                                               \
    }                                                                                                               \
    const int pairs = units2 < max_pairs ? units2 : max_pairs;                                                      \
    k_decoder_fused_pair<L, MAXM><<<2 * pairs, kDecThreads, kPairSmem, st>>>(tmA, tmBh, tmG, p, m_tiles2, n_tiles,  \
                                                                            kblocks, idesc2);                       \
  } while (0)
    switch (loss) {
      case RCD_LOSS_MSE: RCD_DEC_PAIR_LAUNCH(RCD_LOSS_MSE, false, 0); break;
      case RCD_LOSS_NLL:
        if (mode == RCD_DEC_MODE_ROWMAX) RCD_DEC_PAIR_LAUNCH(RCD_LOSS_NLL, true, 3);
        else RCD_DEC_PAIR_LAUNCH(RCD_LOSS_NLL, false, 1);
        break;
      case RCD_LOSS_LOGISTIC: RCD_DEC_PAIR_LAUNCH(RCD_LOSS_LOGISTIC, false, 2); break;
      default:
        rcd_set_error("rcd_decoder_fwd_loss: unknown loss id %d", loss);
        return RCD_ERR_INVALID;
    }
#undef RCD_DEC_PAIR_LAUNCH
    RCD_LAUNCH_CHECK();
    return RCD_OK;
  }
  const int units = m_tiles * n_tiles;
  const int sms = rcd_num_sms();
  const int grid = units < sms ? units : sms;
  static bool attr_set[4] = {false, false, false, false};
#define RCD_DEC_LAUNCH(L, MAXM, SLOT)                                                                             \
  do {                                                                                                            \
    if (!attr_set[SLOT]) {                                                                                        \
      RCD_CUDA(cudaFuncSetAttribute(k_decoder_fused<L, MAXM>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    kDecSmem));                                                                   \
      attr_set[SLOT] = true;                                                                                      \
    }                                                                                                             \
    k_decoder_fused<L, MAXM><<<grid, kDecThreads, kDecSmem, st>>>(tmA, tmB, tmG, p, m_tiles, n_tiles, kblocks,    \
                                                                  idesc);                                         \
  } while (0)
  switch (loss) {
    case RCD_LOSS_MSE: RCD_DEC_LAUNCH(RCD_LOSS_MSE, false, 0); break;
    case RCD_LOSS_NLL:
      if (mode == RCD_DEC_MODE_ROWMAX) RCD_DEC_LAUNCH(RCD_LOSS_NLL, true, 3);
      else RCD_DEC_LAUNCH(RCD_LOSS_NLL, false, 1);
      break;
    case RCD_LOSS_LOGISTIC: RCD_DEC_LAUNCH(RCD_LOSS_LOGISTIC, false, 2); break;
    default:
      rcd_set_error("rcd_decoder_fwd_loss: unknown loss id %d", loss);
      return RCD_ERR_INVALID;
  }
#undef RCD_DEC_LAUNCH
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

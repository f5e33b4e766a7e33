// tcgen05 / TMEM / TMA GEMM engine for sm_100a (bf16 x bf16 -> fp32), hand-written PTX.
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0 (one lane)  TMA producer : cp.async.bulk.tensor 2D boxes (128B swizzle) -> 4-stage smem ring
//   warp 1 (one lane)  MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, M=128, N=bn<=256, K=16 per instruction;
//                                      accumulators double-buffered in TMEM (2 x 256 columns); also owns
//                                      tcgen05.alloc/dealloc
//   warps 2-5          epilogue     : tcgen05.ld 32x32b (TMEM lane == output row) -> row epilogue -> global
//   warps 6-9          side product : (mode 2 only, optional) weighted column sums of the A operand read from the
//                                      smem stages the MMA consumes: colsum[m] = sum_k colw[k]*A[k,m] — the decoder
//                                      bias gradient db = dO^T alpha without another pass over dO
// Three mbarrier pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue).
//
// Operand layouts (smem, 128B swizzle, one 64-wide K block per stage):
//   K-major  operand X[rows, K]  : one TMA box {64 k, rows}; UMMA desc SBO = 1024 B (8 rows x 128 B),
//                                  K step of 16 elements = +32 B on the start address
//   MN-major operand Y[K, cols]  : boxes {64 cols, 64 k} of 8 KB each; UMMA desc LBO = 8192 B (next 64 columns),
//                                  SBO = 1024 B (next 8 k), K step of 16 = +2048 B
// Out-of-range parts of a box are zero-filled by TMA, so ragged M/N/K need no special casing before the
// epilogue, which masks rows >= M and columns >= N.
#include <cuda.h>

#include "gemm_internal.cuh"
#include "tc_ptx.cuh"

namespace rcd {

// MT = 128-row sub-tiles per CTA.  MT = 1: 128 x bn tile, accumulators double-buffered in TMEM (the epilogue of tile i
// overlaps the MMAs of tile i+1), 4-stage ring.  MT = 2: 256 x bn tile as two M=128 MMAs that share the B stage — one
// third less operand traffic per MMA (A 2x16 KB + B 32 KB per 2x512 MMA cycles instead of 16 + 32 KB per 512): the
// GEMMs of this path run at the L2->SM return bandwidth (~48 B/clk/SM measured), not at the tensor pipe, so bytes per
// MMA is what sets their speed.  The two accumulators fill all 512 TMEM columns (no double buffering: the epilogue is
// exposed, which costs little for the long units of dgrad (split-K) and wgrad (K = batch rows)); 3-stage ring.
constexpr int kASubBytes = kTileM * kTileK * 2;       // 16 KB per 128-row sub-tile
constexpr int kBStageBytes = kTileNMax * kTileK * 2;  // 32 KB
constexpr int kBarrierBytes = 256;
constexpr int kSideWarps = 4;
constexpr int kSideSmemBytes = 8 * kTileM * 4;  // [8 k-groups][128 columns] partial sums
constexpr int kGemmThreads = 320;  // warps 0/1 TMA/MMA, 2-5 epilogue, 6-9 column-sum side product
constexpr int kMaxStages = 4;

template <int MT>
struct TcCfg {
  static constexpr int kStages = (MT == 1) ? 4 : 3;
  static constexpr int kAStageBytes = MT * kASubBytes;
  static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
  static constexpr int kAccBufs = (MT == 1) ? 2 : 1;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + kSideSmemBytes + 1024;  // + alignment slack
};

constexpr int kTmemCols = 512;
constexpr int kBoxBytes = 64 * 64 * 2;  // MN-major box

template <int MT>
static __global__ void __launch_bounds__(kGemmThreads, 1)
    k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmProblem g,
              EpiParams e, int m_tiles, int n_tiles, int kblocks, uint32_t idesc, int b_boxes,
              uint32_t stage_tx_bytes) {
  constexpr int kStages = TcCfg<MT>::kStages;
  constexpr int kAStageBytes = TcCfg<MT>::kAStageBytes;
  constexpr int kStageBytes = TcCfg<MT>::kStageBytes;
  constexpr int kAccBufs = TcCfg<MT>::kAccBufs;
  constexpr int kRowsPerCta = MT * kTileM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles = (raw_addr + 1023u) & ~1023u;  // 128B-swizzle atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (tiles - raw_addr);
  const uint32_t bars = tiles + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + 8 * (2 * kStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool do_colsum = (e.colsum != nullptr) && (g.mode == 2);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), do_colsum ? 1 + kSideWarps : 1);  // tcgen05.commit (+ one arrive per side warp)
      }
      for (int a = 0; a < kAccBufs; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = m_tiles * n_tiles * g.splits;
  const bool a_mn = (g.mode == 2), b_mn = (g.mode != 0);

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const UnitCoord u = decode_unit(unit, m_tiles, n_tiles, g.splits, kblocks, g.n_fastest, g.split_major);
        const int m0 = u.mt * kRowsPerCta, n0 = u.nt * g.bn;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 0);
          mbar_arrive_expect_tx(full_bar(stage), stage_tx_bytes);
          const uint32_t a_dst = tiles + stage * kStageBytes, b_dst = a_dst + kAStageBytes;
          const int k = kb * kTileK;
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {  // rows beyond M: TMA zero-fills the box
            const uint32_t a_sub = a_dst + sub * kASubBytes;
            const int ms = m0 + sub * kTileM;
            if (a_mn) {
              tma_load_2d(a_sub, &tmA, full_bar(stage), ms, k);
              tma_load_2d(a_sub + kBoxBytes, &tmA, full_bar(stage), ms + 64, k);
            } else {
              tma_load_2d(a_sub, &tmA, full_bar(stage), k, ms);
            }
          }
          if (b_mn) {
            for (int j = 0; j < b_boxes; ++j) tma_load_2d(b_dst + j * kBoxBytes, &tmB, full_bar(stage), n0 + 64 * j, k);
          } else {
            tma_load_2d(b_dst, &tmB, full_bar(stage), k, n0);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
        const UnitCoord u = decode_unit(unit, m_tiles, n_tiles, g.splits, kblocks, g.n_fastest, g.split_major);
        const int acc = it % kAccBufs;
        const uint32_t acc_phase = (uint32_t)(it / kAccBufs) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kTileNMax);
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 2);
          tc_fence_after();
          const uint32_t a_addr = tiles + stage * kStageBytes, b_addr = a_addr + kAStageBytes;
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) {
            const uint64_t bdesc = b_mn ? make_smem_desc(b_addr + k * 2048, kBoxBytes, 1024)
                                        : make_smem_desc(b_addr + k * 32, 16, 1024);
#pragma unroll
            for (int sub = 0; sub < MT; ++sub) {  // the sub-tiles share the B stage; accumulator `sub` at column sub*256
              const uint32_t a_sub = a_addr + sub * kASubBytes;
              const uint64_t adesc = a_mn ? make_smem_desc(a_sub + k * 2048, kBoxBytes, 1024)
                                          : make_smem_desc(a_sub + k * 32, 16, 1024);
              tc_mma_bf16(tmem_d + (uint32_t)(sub * kTileNMax), adesc, bdesc, idesc, (kb > u.kb0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull_bar(acc));  // accumulator complete
      }
    }
  } else if (warp >= 6) {
    // ---------------- side product warps (6..9): colsum[m] = sum_k colw[k] * A[k, m] ----------------
    // 128 threads; thread t reads the 16-byte chunk (8 columns) c = t%16 of the k rows kg, kg+8, ... (kg = t/16) of
    // every A stage, straight from the swizzled smem the MMA consumes; the 8 k-groups are summed through smem once per
    // tile.  Only the nt == 0 unit of a tile does the work (the other n-tiles see the same A slab).
    if (do_colsum) {
      const int t = (warp - 6) * 32 + lane;  // 0..127
      const uint32_t c = (uint32_t)t & 15u, kgrp = (uint32_t)t >> 4;
      const uint32_t box_off = (c >> 3) * kBoxBytes;
      float* side = reinterpret_cast<float*>(smem + kStages * kStageBytes + kBarrierBytes);
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const UnitCoord u = decode_unit(unit, m_tiles, n_tiles, g.splits, kblocks, g.n_fastest, g.split_major);
        float acc[MT][8];
#pragma unroll
        for (int sub = 0; sub < MT; ++sub)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[sub][j] = 0.f;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(full_bar(stage), phase, 4);
          if (u.nt == 0) {
            const int kg0 = kb * kTileK + lane, kg1 = kg0 + 32;
            const float w_lo = (kg0 < g.K) ? (e.colw ? __ldg(e.colw + kg0) : 1.0f) : 0.f;
            const float w_hi = (kg1 < g.K) ? (e.colw ? __ldg(e.colw + kg1) : 1.0f) : 0.f;
#pragma unroll
            for (int sub = 0; sub < MT; ++sub) {
              const uint8_t* a_ptr = smem + stage * kStageBytes + sub * kASubBytes + box_off;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t k = kgrp + 8u * (uint32_t)i;
                const float w = __shfl_sync(0xffffffffu, (i < 4) ? w_lo : w_hi, (int)(k & 31u));
                const uint4 v = *reinterpret_cast<const uint4*>(a_ptr + k * 128u + (((c & 7u) ^ (k & 7u)) << 4));
                acc[sub][0] = fmaf(w, bf16_lo(v.x), acc[sub][0]);
                acc[sub][1] = fmaf(w, bf16_hi(v.x), acc[sub][1]);
                acc[sub][2] = fmaf(w, bf16_lo(v.y), acc[sub][2]);
                acc[sub][3] = fmaf(w, bf16_hi(v.y), acc[sub][3]);
                acc[sub][4] = fmaf(w, bf16_lo(v.z), acc[sub][4]);
                acc[sub][5] = fmaf(w, bf16_hi(v.z), acc[sub][5]);
                acc[sub][6] = fmaf(w, bf16_lo(v.w), acc[sub][6]);
                acc[sub][7] = fmaf(w, bf16_hi(v.w), acc[sub][7]);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (u.nt == 0) {
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
            for (int j = 0; j < 8; ++j) side[kgrp * kTileM + c * 8 + j] = acc[sub][j];
            named_bar_sync(2, kSideWarps * 32);
            float sum = 0.f;
#pragma unroll
            for (int q2 = 0; q2 < 8; ++q2) sum += side[q2 * kTileM + t];  // fixed order
            const int m = u.mt * kRowsPerCta + sub * kTileM + t;
            if (m < g.M) e.colsum[m] = sum;
            named_bar_sync(2, kSideWarps * 32);
          }
        }
      }
    }
  } else {
    // ---------------- epilogue warps (2..5): TMEM lane quarter = warp % 4 ----------------
    const int q = warp & 3;
    int it = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++it) {
      const UnitCoord u = decode_unit(unit, m_tiles, n_tiles, g.splits, kblocks, g.n_fastest, g.split_major);
      const int acc = it % kAccBufs;
      const uint32_t acc_phase = (uint32_t)(it / kAccBufs) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase, 3);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        const int row = u.mt * kRowsPerCta + sub * kTileM + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc + sub) * kTileNMax);
        RowEpilogue epi;
        epi.begin();
        for (int cb = 0; cb < g.bn; cb += 32) {
          float v[32];
          tc_ld_32x32(taddr + (uint32_t)cb, v);
          epi.chunk32(e, row, u.nt * g.bn + cb, u.split, v);
        }
        epi.end(e, row, u.nt);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols)
                 : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2D bf16 tensor [outer, inner] row-major with leading dimension ld (elements), box {box_inner, box_outer}
int encode_map(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_inner,
               int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    rcd_set_error("gemm_tc: cuTensorMapEncodeTiled not available from the driver");
    return RCD_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) {
    rcd_set_error("gemm_tc: operand base must be 16-byte aligned and ld a multiple of 8 elements (ld=%lld)", ld);
    return RCD_ERR_INVALID;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    rcd_set_error("gemm_tc: cuTensorMapEncodeTiled failed (CUresult %d; inner=%lld outer=%lld ld=%lld box=%dx%d)",
                  (int)r, inner, outer, ld, box_inner, box_outer);
    return RCD_ERR_CUDA;
  }
  return RCD_OK;
}

int gemm_tc_launch(const GemmProblem& g, const EpiParams& e, cudaStream_t st) {
  if (g.bn % 16 != 0 || g.bn < 16 || g.bn > kTileNMax) {
    rcd_set_error("gemm_tc: bad n-tile %d", g.bn);
    return RCD_ERR_INVALID;
  }
  // 256-row CTA tiles (two sub-tiles sharing the B stage) whenever the epilogue is the plain fp32 store and there
  // are at least two sub-tiles of rows; the decoder-logits epilogue keeps its per-(n-tile,row) statistics layout.
  const int mt_sub = (g.m_sub == 2 && e.kind == EPI_F32 && g.M > kTileM) ? 2 : 1;
  const int m_tiles = rcd_div_up(g.M, kTileM * mt_sub), n_tiles = rcd_div_up(g.N, g.bn);
  const int kblocks = rcd_div_up(g.K, kTileK);
  if (n_tiles > 1 && g.bn % 32 != 0) {
    rcd_set_error("gemm_tc: n-tile %d must be a multiple of 32 when N spans several tiles", g.bn);
    return RCD_ERR_INVALID;
  }
  if (g.splits < 1 || (long long)(g.splits - 1) * rcd_div_up(kblocks, g.splits) >= kblocks) {
    rcd_set_error("gemm_tc: %d k-splits over %d k-blocks leaves an empty split", g.splits, kblocks);
    return RCD_ERR_INVALID;
  }
  CUtensorMap tmA, tmB;
  int rc;
  const bool a_mn = (g.mode == 2), b_mn = (g.mode != 0);
  if (a_mn) rc = encode_map(&tmA, g.A, g.M, g.K, g.lda, 64, 64);
  else rc = encode_map(&tmA, g.A, g.K, g.M, g.lda, kTileK, kTileM);
  if (rc != RCD_OK) return rc;
  if (b_mn) rc = encode_map(&tmB, g.B, g.N, g.K, g.ldb, 64, 64);
  else rc = encode_map(&tmB, g.B, g.K, g.N, g.ldb, kTileK, kTileNMax);
  if (rc != RCD_OK) return rc;
  const int b_boxes = rcd_div_up(g.bn, 64);
  const uint32_t tx = (uint32_t)(mt_sub * kASubBytes) + (b_mn ? (uint32_t)(b_boxes * kBoxBytes) : (uint32_t)kBStageBytes);
  // instruction descriptor: D=f32 [4,6)=1 | A=bf16 [7,10)=1 | B=bf16 [10,13)=1 | a_major [15] | b_major [16] |
  // N>>3 [17,23) | M>>4 [24,29)
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                         ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  static bool attr_set = false;
  if (!attr_set) {
    RCD_CUDA(cudaFuncSetAttribute(k_gemm_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<1>::kSmemBytes));
    RCD_CUDA(cudaFuncSetAttribute(k_gemm_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<2>::kSmemBytes));
    attr_set = true;
  }
  const int units = m_tiles * n_tiles * g.splits;
  const int sms = rcd_num_sms();
  const int grid = units < sms ? units : sms;
  if (mt_sub == 2)
    k_gemm_tc<2><<<grid, kGemmThreads, TcCfg<2>::kSmemBytes, st>>>(tmA, tmB, g, e, m_tiles, n_tiles, kblocks, idesc,
                                                                  b_boxes, tx);
  else
    k_gemm_tc<1><<<grid, kGemmThreads, TcCfg<1>::kSmemBytes, st>>>(tmA, tmB, g, e, m_tiles, n_tiles, kblocks, idesc,
                                                                  b_boxes, tx);
  RCD_LAUNCH_CHECK();
  return RCD_OK;
}

}  // namespace rcd

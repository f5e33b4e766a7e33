// Public GEMM entry points (K4 decoder forward, K6 decoder backward, plain GEMM) — shape bookkeeping and engine
// dispatch.  The tcgen05 engine is the product; RCD_GEMM_SIMT selects the validation engine.
#include "gemm_internal.cuh"

using namespace rcd;

namespace {

int launch(const GemmProblem& g, const EpiParams& e, int engine, cudaStream_t st) {
  if (engine == RCD_GEMM_TCGEN05) return gemm_tc_launch(g, e, st);
  if (engine == RCD_GEMM_SIMT) return gemm_simt_launch(g, e, st);
  rcd_set_error("unknown GEMM engine %d", engine);
  return RCD_ERR_INVALID;
}

// n-tile for an output of width N that is not the decoder's item axis (N = hidden size)
int pick_bn(int N) {
  const int n_tiles = rcd_div_up(N, kTileNMax);
  if (n_tiles == 1) return (N + 15) / 16 * 16;
  const int per = rcd_div_up(N, n_tiles);
  return (per + 31) / 32 * 32;
}

}  // namespace

RCD_EXPORT int rcd_decoder_tile_n(void) { return kDecoderTileN; }

RCD_EXPORT int rcd_decoder_fwd(const uint16_t* Zb, int ldzb, const uint16_t* Wg, int ldw, const float* bias, int rows,
                               int n, int H, uint16_t* O_bf16, float* out_f32, int ldo, float* stat_max,
                               float* stat_sum, int engine, void* stream) {
  RCD_CHECK_ARG(Zb && Wg && bias && (O_bf16 || out_f32), "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0 && H > 0, "bad shape");
  RCD_CHECK_ARG(ldzb >= H && ldw >= H && ldo >= n && ldo % 8 == 0, "bad leading dimension");
  RCD_CHECK_ARG((stat_max == nullptr) == (stat_sum == nullptr), "stat_max/stat_sum must come together");
  GemmProblem g{};
  g.mode = 0; g.A = Zb; g.lda = ldzb; g.B = Wg; g.ldb = ldw; g.M = rows; g.N = n; g.K = H;
  g.bn = kDecoderTileN; g.splits = 1; g.n_fastest = 0;  // m fastest: CTAs in flight share the same Wg tile in L2
  EpiParams e{};
  e.kind = EPI_DECODER; e.M = rows; e.N = n; e.bias = bias; e.Obf = O_bf16; e.Of32 = out_f32; e.ldo = ldo;
  e.stat_max = stat_max; e.stat_sum = stat_sum;
  return launch(g, e, engine, (cudaStream_t)stream);
}

RCD_EXPORT int rcd_decoder_dgrad_splits(int rows, int n, int H) {
  if (rows <= 0 || n <= 0 || H <= 0) return 1;
  const int bn = pick_bn(H);
  const int tiles = rcd_div_up(rows, rows > kTileM ? 2 * kTileM : kTileM) * rcd_div_up(H, bn);  // 256-row CTA tiles
  const int kblocks = rcd_div_up(n, kTileK);
  int splits = rcd_div_up(2 * rcd_num_sms(), tiles);  // about two waves of work units
  if (splits > kblocks) splits = kblocks;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  const int per = rcd_div_up(kblocks, splits);
  return rcd_div_up(kblocks, per);  // every split non-empty
}

RCD_EXPORT int rcd_decoder_dgrad(const uint16_t* dO, int lddo, const uint16_t* Wg, int ldw, int rows, int n, int H,
                                 int splits, float* partials, int ldp, int engine, void* stream) {
  RCD_CHECK_ARG(dO && Wg && partials, "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0 && H > 0 && splits > 0, "bad shape");
  RCD_CHECK_ARG(lddo >= n && ldw >= H && ldp >= H, "bad leading dimension");
  GemmProblem g{};
  g.mode = 1; g.A = dO; g.lda = lddo; g.B = Wg; g.ldb = ldw; g.M = rows; g.N = H; g.K = n;
  g.bn = pick_bn(H); g.splits = splits; g.n_fastest = 1; g.split_major = 1; g.m_sub = 2;
  EpiParams e{};
  e.kind = EPI_F32; e.M = rows; e.N = H; e.C = partials; e.ldc = ldp; e.split_stride = (long long)rows * ldp;
  return launch(g, e, engine, (cudaStream_t)stream);
}

// validation-engine path of the wgrad side product: db[c] = sum_r w[r] * G[r,c]
static __global__ void k_colsum_weighted(const uint16_t* __restrict__ G, int ldg, int rows, int n,
                                         const float* __restrict__ w, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r)
    s = fmaf(w ? w[r] : 1.0f, __uint_as_float((uint32_t)G[(size_t)r * ldg + c] << 16), s);
  db[c] = s;
}

RCD_EXPORT int rcd_decoder_wgrad(const uint16_t* G, int ldg, const uint16_t* Zs, int ldzs, int rows, int n, int H,
                                 float* dW, int lddw, const float* col_weight, float* db, int engine, void* stream) {
  RCD_CHECK_ARG(G && Zs && dW, "null pointer");
  RCD_CHECK_ARG(rows > 0 && n > 0 && H > 0, "bad shape");
  RCD_CHECK_ARG(ldg >= n && ldzs >= H && lddw >= H, "bad leading dimension");
  GemmProblem g{};
  g.mode = 2; g.A = G; g.lda = ldg; g.B = Zs; g.ldb = ldzs; g.M = n; g.N = H; g.K = rows;
  g.bn = pick_bn(H); g.splits = 1; g.n_fastest = 1;  // the n-tiles of one G^T tile run back to back (L2 reuse)
  g.m_sub = 2;
  EpiParams e{};
  e.kind = EPI_F32; e.M = n; e.N = H; e.C = dW; e.ldc = lddw; e.split_stride = 0;
  if (engine == RCD_GEMM_TCGEN05) {
    e.colw = col_weight; e.colsum = db;
  } else if (db) {
    k_colsum_weighted<<<rcd_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(G, ldg, rows, n, col_weight, db);
    RCD_LAUNCH_CHECK();
  }
  return launch(g, e, engine, (cudaStream_t)stream);
}

RCD_EXPORT int rcd_gemm_bf16(int mode, const uint16_t* A, int lda, const uint16_t* B, int ldb, int M, int N, int K,
                             float* C, int ldc, int engine, void* stream) {
  RCD_CHECK_ARG(A && B && C, "null pointer");
  RCD_CHECK_ARG(mode >= 0 && mode <= 2 && M > 0 && N > 0 && K > 0 && ldc >= N, "bad shape");
  GemmProblem g{};
  g.mode = mode; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K;
  g.bn = (mode == 0) ? kTileNMax : pick_bn(N); g.splits = 1; g.n_fastest = 1;
  g.m_sub = (engine >> 8) & 3;   // tests: engine | (2 << 8) selects the 256-row CTA tile of the tcgen05 engine
  engine &= 0xff;
  EpiParams e{};
  e.kind = EPI_F32; e.M = M; e.N = N; e.C = C; e.ldc = ldc; e.split_stride = 0;
  return launch(g, e, engine, (cudaStream_t)stream);
}

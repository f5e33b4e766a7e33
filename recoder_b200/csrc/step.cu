// K12: native step executor — one host call per optimizer step (include/recoder_b200.h, "K12").
//
// The per-batch body of Recoder._train (recoder/model.py:383-404: zero_grad -> __compute_loss -> backward ->
// optimizer.step) as a fixed sequence of this library's own entry points, issued from C++ on three CUDA streams.  It is
// the same sequence, in the same order, with the same arguments as recoder_b200/engine.py issues through ctypes
// (TrainEngine._ae_step / _mf_step / _ae_step_items) — the two paths are bit-identical and the tests hold them to that —
// minus ~1 ms of interpreter time per step, which is what bounds the small configurations (C1 / C2: 0.1-0.3 ms of GPU
// work per step) and what lets one slow host stall eight GPUs at the barriers of the item-parallel mode.
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace rcd {

struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};

struct StepCtx {
  cudaEvent_t ev_fork = nullptr;  // main -> side / aux hand-over points
  cudaEvent_t ev_csc = nullptr;   // aux: column-major views ready
  cudaEvent_t ev_out = nullptr;   // side: output-table update done
  bool out_pending = false;
  int prof_mode = 0;
  std::string prof_name;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;  // recycled timing events
};

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int rup(long long x, int m) { return (int)((x + m - 1) / m * m); }

// Workspace layout from the CAPACITIES of the step (stable across steps).
struct Layout {
  size_t csc_ptr[2], csc_row[2], csc_val[2], csc_src[2], csc_scratch;  // [0] target pool, [1] input pool (if different)
  size_t slab, Wg, bg, Z, Zb, G, o_nnz, corr, row_ref, row_ref2, ssum2, stat, alpha, Zs, loss_blocks, row_redo, partials, dA,
      heavy, zero_bias, total;
  size_t heavy_bytes, csc_scratch_bytes;
  int stat_cols_cap;
};

static Layout make_layout(const rcd_step_args& a) {
  Layout L{};
  const size_t rows = (size_t)a.cap_rows, n = (size_t)a.cap_n, n_in = (size_t)a.cap_n_in;
  const size_t nnz = (size_t)(a.cap_tnnz > 0 ? a.cap_tnnz : 1), nnz_in = (size_t)(a.cap_nnz > 0 ? a.cap_nnz : 1);
  const size_t H = (size_t)a.H, ldh = (size_t)rup(a.H, 8), ldn = (size_t)rup((long long)n, 8);
  size_t at = 0;
  auto take = [&](size_t bytes) {
    const size_t o = at;
    at += al256(bytes);
    return o;
  };
  L.csc_scratch_bytes = rcd_slice_csc_scratch_bytes((int)(n > n_in ? n : n_in), (int)(nnz > nnz_in ? nnz : nnz_in));
  for (int k = 0; k < 2; ++k) {
    const size_t nn = k == 0 ? n : n_in, zz = k == 0 ? nnz : nnz_in;
    if (k == 1 && a.same_pool) {
      L.csc_ptr[1] = L.csc_ptr[0]; L.csc_row[1] = L.csc_row[0]; L.csc_val[1] = L.csc_val[0]; L.csc_src[1] = L.csc_src[0];
      break;
    }
    L.csc_ptr[k] = take((nn + 1) * 4);
    L.csc_row[k] = take(zz * 4);
    L.csc_val[k] = take(zz * 4);
    L.csc_src[k] = take(zz * 4);
  }
  L.csc_scratch = take(L.csc_scratch_bytes);
  // gradient slab: AE [dW_in n_in*H | dW_out n*H | db_out n4 | db_in h4]; MF [dV n*D | dbias n4 | dU rows*D]
  const size_t n4 = (size_t)rup((long long)n, 4), h4 = (size_t)rup(a.H, 4);
  const size_t slab_f = a.kind == RCD_MODEL_AE ? n_in * H + n * H + n4 + h4 + 4 : n * H + n4 + rows * H + 4;
  L.slab = take(slab_f * 4);
  L.Wg = take(n * ldh * 2);
  L.bg = take(n * 4);
  L.Z = take(rows * H * 4);
  L.Zb = take(rows * ldh * 2);
  L.G = take(rows * ldn * 2);
  L.o_nnz = take(nnz * 4);
  L.corr = take(nnz * 4);
  L.row_ref = take(rows * 4);
  L.row_ref2 = take(rows * 4);
  L.ssum2 = take(rows * 4);
  L.stat_cols_cap = rcd_decoder_stat_cols((int)n);
  L.stat = take(rows * (size_t)L.stat_cols_cap * 4);
  L.alpha = take(rows * 4);
  L.Zs = take(rows * ldh * 2);
  L.loss_blocks = take((size_t)rcd_loss_finish_blocks((int)rows) * 8);
  L.row_redo = take(rows * 4);
  // split-K partials: the split count depends on the actual shape; 64 is the engine's cap (+1 slot for the sparse part)
  size_t max_splits = 64;
  {
    const size_t kb = (n + 63) / 64;
    if (kb < max_splits) max_splits = kb;
    if (max_splits < 1) max_splits = 1;
  }
  L.partials = take((max_splits + 1) * rows * H * 4);
  L.dA = take(rows * H * 4);
  L.heavy_bytes = rows > 4096 ? rcd_csc_heavy_scratch_bytes((int)(n > n_in ? n : n_in), (long long)(nnz > nnz_in ? nnz : nnz_in),
                                                            a.H)
                              : 0;
  L.heavy = take(L.heavy_bytes);
  L.zero_bias = take(h4 * 4);
  L.total = at;
  return L;
}

static __global__ void k_stash_loss(const double* __restrict__ loss, float* __restrict__ tail) {
  const double v = *loss;
  const float hi = (float)v;
  tail[0] = hi;
  tail[1] = (float)(v - (double)hi);
}
static __global__ void k_unstash_loss(const float* __restrict__ tail, double* __restrict__ loss) {
  *loss = (double)tail[0] + (double)tail[1];
}

}  // namespace rcd

using namespace rcd;

// ---- profiling helpers ------------------------------------------------------------------------------------------------
static cudaEvent_t prof_event(StepCtx* c) {
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// CALL(name, stream, expr): runs `expr` (an int-status entry point taking `stream` last) with optional event timing.
#define STEP_CALL(NAME, STREAM, EXPR)                                                     \
  do {                                                                                    \
    const bool _prof = c->prof_mode == 1 || (c->prof_mode == 2 && c->prof_name == NAME);  \
    ProfRec _r{NAME, nullptr, nullptr};                                                   \
    if (_prof) {                                                                          \
      _r.a = prof_event(c);                                                               \
      _r.b = prof_event(c);                                                               \
      cudaEventRecord(_r.a, (cudaStream_t)(STREAM));                                      \
    }                                                                                     \
    const int _rc = (EXPR);                                                               \
    if (_prof) {                                                                          \
      cudaEventRecord(_r.b, (cudaStream_t)(STREAM));                                      \
      c->prof.push_back(_r);                                                              \
    }                                                                                     \
    if (_rc != RCD_OK) return _rc;                                                        \
  } while (0)

static int opt_step(StepCtx* c, const rcd_step_args& a, const rcd_param& p, const float* grad, int ldg, const int32_t* pos,
                    void* st, const int64_t* ids = nullptr, long long n_ids = 0) {
  const long long rows = p.rows;
  const int H = p.cols;
  if (p.last) {  // deferred dense Adam: the batch's rows only (ids NULL: the identity, i.e. every row)
    RCD_CHECK_ARG(a.optimizer == RCD_OPT_ADAM, "deferred updates are an Adam mode");
    STEP_CALL("rcd_adam_lazy_update", st,
              rcd_adam_lazy_update(p.p, p.s1, p.s2, H, ids, n_ids, grad, ldg, p.last, a.lr, 0.9, 0.999, 1e-8,
                                   p.weight_decay, p.t, st));
    return RCD_OK;
  }
  switch (a.optimizer) {
    case RCD_OPT_ADAM:
      STEP_CALL("rcd_adam_step", st,
                rcd_adam_step(p.p, p.s1, p.s2, rows, H, grad, ldg, pos, a.lr, 0.9, 0.999, 1e-8, p.weight_decay, p.t, st));
      break;
    case RCD_OPT_SGD:
      STEP_CALL("rcd_sgd_step", st, rcd_sgd_step(p.p, p.s1, rows, H, grad, ldg, pos, a.lr, 0.9, p.weight_decay, st));
      break;
    case RCD_OPT_ADAGRAD:
      STEP_CALL("rcd_adagrad_step", st,
                rcd_adagrad_step(p.p, p.s1, rows, H, grad, ldg, pos, a.lr, 1e-10, p.weight_decay, st));
      break;
    case RCD_OPT_RMSPROP:
      STEP_CALL("rcd_rmsprop_step", st,
                rcd_rmsprop_step(p.p, p.s1, p.s2, rows, H, grad, ldg, pos, a.lr, 0.99, 1e-8, 0.9, p.weight_decay, st));
      break;
    default:
      rcd_set_error("rcd_step_run: unknown optimizer %d", a.optimizer);
      return RCD_ERR_INVALID;
  }
  return RCD_OK;
}

// deferred dense Adam: rows ids[0..n) of table `p` are brought up to date (steps <= p.t - 1) before they are read
static int catch_up(StepCtx* c, const rcd_step_args& a, const rcd_param& p, const int64_t* ids, long long n, void* st) {
  if (!p.last || p.t <= 1) return RCD_OK;
  RCD_CHECK_ARG(a.scal, "deferred Adam needs the scalar table");
  STEP_CALL("rcd_adam_lazy_catchup", st,
            rcd_adam_lazy_catchup(p.p, p.s1, p.s2, p.cols, ids, n, p.last, p.t - 1, a.scal, a.scal_base, a.scal_len, 0.9,
                                  0.999, 1e-8, p.weight_decay, 1, nullptr, nullptr, st));
  return RCD_OK;
}

static int ip_barrier(StepCtx* c, const rcd_step_args& a, void* st) {
  const unsigned int seq = ++(*a.ip.seq_host);
  STEP_CALL("rcd_p2p_barrier", st,
            rcd_p2p_barrier(a.ip.flags_host, a.ip.rank, a.ip.world, seq, a.bad_flag, a.ip.barrier_timeout_s, st));
  return RCD_OK;
}

// per-rank pointer table of the shared block shifted by `off` floats (host array on the stack of the caller)
static void shifted(const rcd_step_args& a, long long off, float** out) {
  for (int q = 0; q < a.ip.world; ++q) out[q] = a.ip.shared_host[q] + off;
}

RCD_EXPORT size_t rcd_step_args_size(void) { return sizeof(rcd_step_args); }

RCD_EXPORT int rcd_step_create(void** ctx_out) {
  RCD_CHECK_ARG(ctx_out, "null pointer");
  StepCtx* c = new StepCtx();
  RCD_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  RCD_CUDA(cudaEventCreateWithFlags(&c->ev_csc, cudaEventDisableTiming));
  RCD_CUDA(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
  *ctx_out = c;
  return RCD_OK;
}

RCD_EXPORT int rcd_step_destroy(void* ctx) {
  StepCtx* c = reinterpret_cast<StepCtx*>(ctx);
  if (!c) return RCD_OK;
  for (auto& r : c->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_csc);
  cudaEventDestroy(c->ev_out);
  delete c;
  return RCD_OK;
}

RCD_EXPORT size_t rcd_step_workspace_bytes(const rcd_step_args* args) {
  if (!args || args->abi != RCD_STEP_ABI || args->H <= 0 || args->cap_rows <= 0 || args->cap_n <= 0) return 0;
  return make_layout(*args).total + 256;
}

RCD_EXPORT int rcd_step_join(void* ctx, void* stream) {
  StepCtx* c = reinterpret_cast<StepCtx*>(ctx);
  RCD_CHECK_ARG(c, "null context");
  if (c->out_pending) {
    RCD_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_out, 0));
    c->out_pending = false;
  }
  return RCD_OK;
}

RCD_EXPORT int rcd_step_profile(void* ctx, int mode, const char* name) {
  StepCtx* c = reinterpret_cast<StepCtx*>(ctx);
  RCD_CHECK_ARG(c && mode >= 0 && mode <= 2 && (mode != 2 || name), "bad arguments");
  c->prof_mode = mode;
  c->prof_name = name ? name : "";
  return RCD_OK;
}

RCD_EXPORT int rcd_step_profile_read(void* ctx, char* names_out, int names_cap, float* ms_out, int* count_out,
                                     int max_names) {
  StepCtx* c = reinterpret_cast<StepCtx*>(ctx);
  RCD_CHECK_ARG(c && names_out && ms_out && count_out && names_cap > 0 && max_names > 0, "bad arguments");
  std::vector<std::string> names;
  std::vector<float> ms;
  std::vector<int> cnt;
  for (auto& r : c->prof) {
    RCD_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    RCD_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    size_t k = 0;
    for (; k < names.size(); ++k)
      if (names[k] == r.name) break;
    if (k == names.size()) {
      names.push_back(r.name);
      ms.push_back(0.f);
      cnt.push_back(0);
    }
    ms[k] += t;
    cnt[k] += 1;
    c->ev_pool.push_back(r.a);
    c->ev_pool.push_back(r.b);
  }
  c->prof.clear();
  std::string joined;
  int out = 0;
  for (size_t k = 0; k < names.size() && out < max_names; ++k) {
    if ((int)(joined.size() + names[k].size() + 2) > names_cap) break;
    if (out) joined += "\n";
    joined += names[k];
    ms_out[out] = ms[k];
    count_out[out] = cnt[k];
    ++out;
  }
  memcpy(names_out, joined.c_str(), joined.size() + 1);
  return out;
}

// NLL / MSE / logistic: sparse side, fused decoder GEMM + loss epilogue, row finish (+ the on-device NLL redo), loss sum.
// Mirrors TrainEngine._decoder_and_loss.  Xb / X: bf16 / fp32 decoder input [rows, ldh / H].
static int decoder_and_loss(StepCtx* c, const rcd_step_args& a, const Layout& L, uint8_t* ws, const uint16_t* Xb,
                            const float* X, const float* row_ref_in, bool ip, void* st, const float** alpha_out,
                            const uint16_t** Zs_out) {
  const int H = a.H, ldh = rup(H, 8), rows = a.rows, n = a.tgt.n, ldn = rup(n, 8);
  const bool nll = a.loss == RCD_LOSS_NLL;
  const uint16_t* Wg = reinterpret_cast<uint16_t*>(ws + L.Wg);
  const float* bg = reinterpret_cast<float*>(ws + L.bg);
  uint16_t* G = reinterpret_cast<uint16_t*>(ws + L.G);
  float* o_nnz = reinterpret_cast<float*>(ws + L.o_nnz);
  float* corr = reinterpret_cast<float*>(ws + L.corr);
  float* stat = reinterpret_cast<float*>(ws + L.stat);
  float* alpha = nll ? reinterpret_cast<float*>(ws + L.alpha) : nullptr;
  uint16_t* Zs = (nll && a.train) ? reinterpret_cast<uint16_t*>(ws + L.Zs) : nullptr;
  double* loss_blocks = reinterpret_cast<double*>(ws + L.loss_blocks);
  int32_t* row_redo = nll ? reinterpret_cast<int32_t*>(ws + L.row_redo) : nullptr;
  const int stat_cols = rcd_decoder_stat_cols(n);
  const int nblocks = rcd_loss_finish_blocks(rows);
  float* row_ref = nullptr;
  if (nll) row_ref = ip ? const_cast<float*>(row_ref_in) : reinterpret_cast<float*>(ws + L.row_ref);
  const rcd_pool_view& t = a.tgt;
  if (!ip) {
    STEP_CALL("rcd_sddmm", st,
              rcd_sddmm(Xb, ldh, Wg, ldh, bg, H, t.row_ptr, t.cols, t.vals, a.row0, rows, a.loss, a.confidence, a.inv_b,
                        o_nnz, corr, row_ref, st));
  }
  STEP_CALL("rcd_decoder_fwd_loss", st,
            rcd_decoder_fwd_loss(Xb, ldh, Wg, ldh, bg, rows, n, H, a.loss, a.inv_b, row_ref, G, ldn, stat, stat_cols,
                                 RCD_DEC_MODE_LOSS, nullptr, st));
  *alpha_out = alpha;
  *Zs_out = Zs ? Zs : Xb;
  if (ip) return RCD_OK;  // the item-parallel step finishes the loss after combining the shards' statistics
  if (nll) {
    int32_t* flag = a.redo_flag;
    STEP_CALL("rcd_loss_finish", st,
              rcd_loss_finish(stat, stat_cols, stat_cols, rows, a.loss, a.confidence, a.inv_b, row_ref, t.row_sum,
                              t.row_ptr, t.vals, o_nnz, a.row0, alpha, X, H, Zs, ldh, nullptr, a.bad_flag, 0, loss_blocks,
                              flag, row_redo, nullptr, st));
    STEP_CALL("rcd_decoder_fwd_loss", st,
              rcd_decoder_fwd_loss(Xb, ldh, Wg, ldh, bg, rows, n, H, a.loss, a.inv_b, row_ref, G, ldn, stat, stat_cols,
                                   RCD_DEC_MODE_ROWMAX, flag, st));
    STEP_CALL("rcd_nll_ref_fix", st, rcd_nll_ref_fix(stat, stat_cols, stat_cols, rows, row_redo, row_ref, flag, st));
    STEP_CALL("rcd_decoder_fwd_loss", st,
              rcd_decoder_fwd_loss(Xb, ldh, Wg, ldh, bg, rows, n, H, a.loss, a.inv_b, row_ref, G, ldn, stat, stat_cols,
                                   RCD_DEC_MODE_LOSS, flag, st));
    STEP_CALL("rcd_loss_finish", st,
              rcd_loss_finish(stat, stat_cols, stat_cols, rows, a.loss, a.confidence, a.inv_b, row_ref, t.row_sum,
                              t.row_ptr, t.vals, o_nnz, a.row0, alpha, X, H, Zs, ldh, nullptr, a.bad_flag, 0, loss_blocks,
                              nullptr, nullptr, flag, st));
    STEP_CALL("rcd_loss_sum", st, rcd_loss_sum(loss_blocks, nblocks, a.loss_acc, flag, st));
  } else {
    STEP_CALL("rcd_loss_finish", st,
              rcd_loss_finish(stat, stat_cols, stat_cols, rows, a.loss, a.confidence, a.inv_b, nullptr, t.row_sum,
                              t.row_ptr, t.vals, o_nnz, a.row0, nullptr, X, H, nullptr, ldh, nullptr, a.bad_flag, 0,
                              loss_blocks, nullptr, nullptr, nullptr, st));
    STEP_CALL("rcd_loss_sum", st, rcd_loss_sum(loss_blocks, nblocks, a.loss_acc, nullptr, st));
  }
  return RCD_OK;
}

static int slice_csc(StepCtx* c, const rcd_step_args& a, const Layout& L, uint8_t* ws, const rcd_pool_view& p, int k,
                     void* st) {
  const int nnz = (int)(p.nnz_slice > 0 ? p.nnz_slice : 1);
  STEP_CALL("rcd_slice_csc", st,
            rcd_slice_csc(p.row_ptr, p.cols, p.vals, a.row0, a.rows, p.n, reinterpret_cast<int32_t*>(ws + L.csc_ptr[k]),
                          reinterpret_cast<int32_t*>(ws + L.csc_row[k]), reinterpret_cast<float*>(ws + L.csc_val[k]),
                          reinterpret_cast<int32_t*>(ws + L.csc_src[k]), ws + L.csc_scratch,
                          rcd_slice_csc_scratch_bytes(p.n, nnz), st));
  return RCD_OK;
}

RCD_EXPORT int rcd_step_run(void* ctx, rcd_step_args* args) {
  StepCtx* c = reinterpret_cast<StepCtx*>(ctx);
  RCD_CHECK_ARG(c && args, "null pointer");
  rcd_step_args& a = *args;
  RCD_CHECK_ARG(a.abi == RCD_STEP_ABI, "rcd_step_args ABI mismatch");
  RCD_CHECK_ARG(a.kind == RCD_MODEL_AE || a.kind == RCD_MODEL_MF, "unknown model kind");
  RCD_CHECK_ARG(a.H > 0 && a.rows > 0 && a.row0 >= 0 && a.in.n > 0 && a.tgt.n > 0, "bad shape");
  RCD_CHECK_ARG(a.rows <= a.cap_rows && a.tgt.n <= a.cap_n && a.in.n <= a.cap_n_in && a.tgt.nnz_slice <= a.cap_tnnz &&
                    a.in.nnz_slice <= a.cap_nnz,
                "the step exceeds the workspace capacities");
  RCD_CHECK_ARG(a.ws && a.loss_acc && a.bad_flag && a.redo_flag, "null pointer");
  RCD_CHECK_ARG(!a.ip.enabled || (a.kind == RCD_MODEL_AE && a.same_pool && a.train && a.ip.world > 1 &&
                                  a.ip.world <= RCD_MAX_PEERS && a.ip.flags_host && a.ip.seq_host && a.ip.shared_host),
                "item-parallel block incomplete");
  const Layout L = make_layout(a);
  RCD_CHECK_ARG(a.ws_bytes >= L.total, "workspace too small (rcd_step_workspace_bytes)");
  uint8_t* ws = reinterpret_cast<uint8_t*>(a.ws);
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) {
    rcd_set_error("rcd_step_run: workspace must be 256-byte aligned");
    return RCD_ERR_INVALID;
  }
  cudaStream_t sm = (cudaStream_t)a.stream_main;
  cudaStream_t ss = a.overlap ? (cudaStream_t)a.stream_side : sm;
  cudaStream_t sa = a.overlap ? (cudaStream_t)a.stream_aux : sm;
  void* st = (void*)sm;
  const int H = a.H, ldh = rup(H, 8), rows = a.rows, row0 = a.row0;
  const int n = a.tgt.n, n_in = a.in.n, ldn = rup(n, 8);
  const bool ip = a.ip.enabled != 0;
  const bool nll = a.loss == RCD_LOSS_NLL;
  const bool ae = a.kind == RCD_MODEL_AE;
  const int tnnz = (int)(a.tgt.nnz_slice > 0 ? a.tgt.nnz_slice : 1);
  const int innz = (int)(a.in.nnz_slice > 0 ? a.in.nnz_slice : 1);

  // ---- aux stream: column-major views of the slice (needed late, by the weight gradients) ----------------------------
  if (a.train) {
    if (sa != sm) {
      RCD_CUDA(cudaEventRecord(c->ev_fork, sm));
      RCD_CUDA(cudaStreamWaitEvent(sa, c->ev_fork, 0));
      // the layout is shared with the previous step: its side-stream update may still be reading the slab
      if (c->out_pending) RCD_CUDA(cudaStreamWaitEvent(sa, c->ev_out, 0));
    }
    int rc = slice_csc(c, a, L, ws, a.tgt, 0, (void*)sa);
    if (rc != RCD_OK) return rc;
    if (ae && !a.same_pool) {
      rc = slice_csc(c, a, L, ws, a.in, 1, (void*)sa);
      if (rc != RCD_OK) return rc;
    }
    if (sa != sm) RCD_CUDA(cudaEventRecord(c->ev_csc, sa));
  }

  // ---- gradient slab ---------------------------------------------------------------------------------------------------
  float* slab = reinterpret_cast<float*>(ws + L.slab);
  const size_t n4 = (size_t)rup(n, 4), h4 = (size_t)rup(H, 4);
  float *dW_in, *dW_out, *db_out, *db_in;
  if (ae) {
    dW_in = slab;
    dW_out = slab + (size_t)n_in * H;
    db_out = dW_out + (size_t)n * H;
    db_in = db_out + n4;
  } else {
    dW_out = slab;                     // dV
    db_out = slab + (size_t)n * H;     // dbias
    dW_in = db_out + n4;               // dU [rows, D]
    db_in = nullptr;
  }
  a.out_dW_in = (long long)((uint8_t*)dW_in - ws);
  a.out_dW_out = (long long)((uint8_t*)dW_out - ws);
  a.out_db_out = (long long)((uint8_t*)db_out - ws);
  a.out_db_in = db_in ? (long long)((uint8_t*)db_in - ws) : -1;

  uint16_t* Wg = reinterpret_cast<uint16_t*>(ws + L.Wg);
  float* bg = reinterpret_cast<float*>(ws + L.bg);
  float* Z = reinterpret_cast<float*>(ws + L.Z);
  uint16_t* Zb = reinterpret_cast<uint16_t*>(ws + L.Zb);
  uint16_t* G = reinterpret_cast<uint16_t*>(ws + L.G);
  float* corr = reinterpret_cast<float*>(ws + L.corr);
  float* o_nnz = reinterpret_cast<float*>(ws + L.o_nnz);
  float* partials = reinterpret_cast<float*>(ws + L.partials);
  float* dA = reinterpret_cast<float*>(ws + L.dA);
  void* heavy = (rows > 4096 && L.heavy_bytes) ? (void*)(ws + L.heavy) : nullptr;

  // ---- forward ---------------------------------------------------------------------------------------------------------
  if (c->out_pending) {  // the previous step's output-table update (side stream) has landed
    if (ss != sm) RCD_CUDA(cudaStreamWaitEvent(sm, c->ev_out, 0));
    c->out_pending = false;
  }
  {
    int rc = catch_up(c, a, a.table_out, a.tgt.items, a.tgt.items ? n : a.table_out.rows, st);
    if (rc != RCD_OK) return rc;
    if (ae) rc = catch_up(c, a, a.table_in, a.in.items, a.in.items ? n_in : a.table_in.rows, st);
    else rc = catch_up(c, a, a.table_in, a.in.users + row0, rows, st);
    if (rc != RCD_OK) return rc;
  }
  STEP_CALL("rcd_gather_rows", st, rcd_gather_rows(a.table_out.p, H, a.tgt.items, n, 0, Wg, ldh, nullptr, st));
  STEP_CALL("rcd_gather_vec", st, rcd_gather_vec(a.bias_out.p, a.tgt.items, n, bg, st));
  float* row_ref_ip = nullptr;
  float* shared_q[RCD_MAX_PEERS];
  if (ae && !ip) {
    STEP_CALL("rcd_ae_encoder_fwd", st,
              rcd_ae_encoder_fwd(a.table_in.p, H, a.bias_in.p, a.in.row_ptr, a.in.raw_items, a.in.vals,
                                 a.in.row_inv_norm, row0, rows, a.act, Z, Zb, ldh, st));
  } else if (ae) {
    // item-parallel: partial sums over this rank's items -> all-reduce -> bias + activation
    float* zero_bias = reinterpret_cast<float*>(ws + L.zero_bias);
    RCD_CUDA(cudaMemsetAsync(zero_bias, 0, h4 * 4, sm));
    float* Zp = a.ip.shared_local + a.ip.off_z;
    STEP_CALL("rcd_ae_encoder_fwd", st,
              rcd_ae_encoder_fwd(a.table_in.p, H, zero_bias, a.in.row_ptr, a.in.raw_items, a.in.vals, a.in.row_inv_norm,
                                 row0, rows, RCD_ACT_NONE, Zp, nullptr, ldh, st));
    int rc = ip_barrier(c, a, st);
    if (rc != RCD_OK) return rc;
    shifted(a, a.ip.off_z, shared_q);
    STEP_CALL("rcd_p2p_allreduce", st,
              rcd_p2p_allreduce(shared_q, a.ip.shared_mc ? a.ip.shared_mc + a.ip.off_z : nullptr,
                                (long long)rup((long long)rows * H, 4), a.ip.rank, a.ip.world, st));
    rc = ip_barrier(c, a, st);
    if (rc != RCD_OK) return rc;
    STEP_CALL("rcd_bias_act", st, rcd_bias_act(Zp, a.bias_in.p, rows, H, a.act, Z, Zb, ldh, st));
    // sparse side with the softmax reference combined across the shards
    float* ref_local = a.ip.shared_local + a.ip.off_ref;
    STEP_CALL("rcd_sddmm", st,
              rcd_sddmm(Zb, ldh, Wg, ldh, bg, H, a.tgt.row_ptr, a.tgt.cols, a.tgt.vals, row0, rows, a.loss, a.confidence,
                        a.inv_b, o_nnz, corr, nll ? ref_local : nullptr, st));
    if (nll) {
      rc = ip_barrier(c, a, st);
      if (rc != RCD_OK) return rc;
      row_ref_ip = reinterpret_cast<float*>(ws + L.row_ref2);
      shifted(a, a.ip.off_ref, shared_q);
      STEP_CALL("rcd_p2p_reduce", st,
                rcd_p2p_reduce(shared_q, a.ip.world, 0, rows, row_ref_ip, RCD_REDUCE_MAX, st));
    }
  } else {
    // MF: user rows -> activation (recoder/nn.py:348-349)
    STEP_CALL("rcd_gather_rows", st,
              rcd_gather_rows(a.table_in.p, H, a.in.users + row0, rows, a.act, Zb, ldh, Z, st));
  }

  const float* alpha = nullptr;
  const uint16_t* Zs = nullptr;
  {
    int rc = decoder_and_loss(c, a, L, ws, Zb, Z, row_ref_ip, ip, st, &alpha, &Zs);
    if (rc != RCD_OK) return rc;
  }
  if (ip) {
    float* stat = reinterpret_cast<float*>(ws + L.stat);
    const int stat_cols = rcd_decoder_stat_cols(n);
    const float* stat_use = stat;
    int stat_ld = stat_cols, stat_n = stat_cols;
    if (nll) {
      float* ssum = a.ip.shared_local + a.ip.off_sum;
      STEP_CALL("rcd_rowsum", st, rcd_rowsum(stat, rows, stat_cols, stat_cols, ssum, st));
      int rc = ip_barrier(c, a, st);
      if (rc != RCD_OK) return rc;
      float* ssum_out = reinterpret_cast<float*>(ws + L.ssum2);
      shifted(a, a.ip.off_sum, shared_q);
      STEP_CALL("rcd_p2p_reduce", st, rcd_p2p_reduce(shared_q, a.ip.world, 0, rows, ssum_out, RCD_REDUCE_SUM, st));
      stat_use = ssum_out;
      stat_ld = stat_n = 1;
    }
    STEP_CALL("rcd_loss_finish", st,
              rcd_loss_finish(stat_use, stat_ld, stat_n, rows, a.loss, a.confidence, a.inv_b, row_ref_ip, a.tgt.row_sum,
                              a.tgt.row_ptr, a.tgt.vals, o_nnz, row0, const_cast<float*>(alpha), Z, H,
                              (nll && a.train) ? reinterpret_cast<uint16_t*>(ws + L.Zs) : nullptr, ldh, a.loss_acc,
                              a.bad_flag, 1, nullptr, nullptr, nullptr, nullptr, st));
  }
  if (!a.train) return RCD_OK;

  // ---- backward --------------------------------------------------------------------------------------------------------
  const int splits = rcd_decoder_dgrad_splits(rows, n, H);
  float* sparse_slot = partials + (size_t)splits * rows * H;
  STEP_CALL("rcd_sparse_dgrad", st,
            rcd_sparse_dgrad(a.table_out.p, H, a.tgt.row_ptr, a.tgt.raw_items, corr, row0, rows, sparse_slot, H, st));
  if (sa != sm) RCD_CUDA(cudaStreamWaitEvent(sm, c->ev_csc, 0));
  const int32_t* csc_ptr_t = reinterpret_cast<int32_t*>(ws + L.csc_ptr[0]);
  const int32_t* csc_row_t = reinterpret_cast<int32_t*>(ws + L.csc_row[0]);
  const int32_t* csc_src_t = reinterpret_cast<int32_t*>(ws + L.csc_src[0]);
  STEP_CALL("rcd_decoder_wgrad", st,
            rcd_decoder_wgrad(G, ldn, Zs, ldh, rows, n, H, dW_out, H, alpha, db_out, RCD_GEMM_TCGEN05, st));
  {
    const size_t hb = heavy ? rcd_csc_heavy_scratch_bytes(n, tnnz, H) : 0;
    STEP_CALL("rcd_csc_rows_accumulate", st,
              rcd_csc_rows_accumulate(Z, H, csc_ptr_t, csc_row_t, csc_src_t, corr, n, dW_out, db_out, heavy, hb, tnnz,
                                      st));
  }
  // output-table update on the side stream, underneath the dgrad GEMM and the encoder backward
  if (ss != sm) {
    RCD_CUDA(cudaEventRecord(c->ev_fork, sm));
    RCD_CUDA(cudaStreamWaitEvent(ss, c->ev_fork, 0));
  }
  {
    int rc = opt_step(c, a, a.table_out, dW_out, H, a.tgt.pos, (void*)ss, a.tgt.items,
                      a.tgt.items ? n : a.table_out.rows);
    if (rc != RCD_OK) return rc;
    rc = opt_step(c, a, a.bias_out, db_out, 1, a.tgt.pos, (void*)ss);
    if (rc != RCD_OK) return rc;
  }
  // deferred Adam: rows of the NEXT pool that are not in this batch are replayed up to THIS step now, on the side
  // stream, underneath the dgrad GEMM / encoder backward (they are disjoint from the rows this step updates)
  if (a.next_items_out && a.table_out.last && a.scal) {
    STEP_CALL("rcd_adam_lazy_catchup", (void*)ss,
              rcd_adam_lazy_catchup(a.table_out.p, a.table_out.s1, a.table_out.s2, a.table_out.cols, a.next_items_out,
                                    a.next_cap_out, a.table_out.last, a.table_out.t, a.scal, a.scal_base, a.scal_len, 0.9,
                                    0.999, 1e-8, a.table_out.weight_decay, 1, a.next_n_out, a.tgt.pos, (void*)ss));
  }
  if (ae && a.next_items_in && a.table_in.last && a.scal) {
    STEP_CALL("rcd_adam_lazy_catchup", (void*)ss,
              rcd_adam_lazy_catchup(a.table_in.p, a.table_in.s1, a.table_in.s2, a.table_in.cols, a.next_items_in,
                                    a.next_cap_in, a.table_in.last, a.table_in.t, a.scal, a.scal_base, a.scal_len, 0.9,
                                    0.999, 1e-8, a.table_in.weight_decay, 1, a.next_n_in, a.in.pos, (void*)ss));
  }
  if (ss != sm) {
    RCD_CUDA(cudaEventRecord(c->ev_out, ss));
    c->out_pending = true;
  }

  if (ae && !ip) {
    STEP_CALL("rcd_decoder_dgrad", st,
              rcd_decoder_dgrad(G, ldn, Wg, ldh, rows, n, H, splits, partials, H, RCD_GEMM_TCGEN05, st));
    STEP_CALL("rcd_dz_act", st, rcd_dz_act(partials, splits + 1, splits, alpha, H, Z, rows, H, a.act, dA, db_in, st));
  } else if (ae) {
    float* dZ = a.ip.shared_local + a.ip.off_dz;
    const long long nz = rup((long long)rows * H, 4);
    STEP_CALL("rcd_decoder_dgrad", st,
              rcd_decoder_dgrad(G, ldn, Wg, ldh, rows, n, H, splits, partials, H, RCD_GEMM_TCGEN05, st));
    STEP_CALL("rcd_dz_act", st,
              rcd_dz_act(partials, splits + 1, splits, alpha, H, Z, rows, H, RCD_ACT_NONE, dZ, nullptr, st));
    k_stash_loss<<<1, 1, 0, sm>>>(a.loss_acc, dZ + nz + 2);   // the loss shares ride in the tail of the dL/dZ all-reduce
    RCD_LAUNCH_CHECK();
    int rc = ip_barrier(c, a, st);
    if (rc != RCD_OK) return rc;
    shifted(a, a.ip.off_dz, shared_q);
    STEP_CALL("rcd_p2p_allreduce", st,
              rcd_p2p_allreduce(shared_q, a.ip.shared_mc ? a.ip.shared_mc + a.ip.off_dz : nullptr, nz + 4, a.ip.rank,
                                a.ip.world, st));
    rc = ip_barrier(c, a, st);
    if (rc != RCD_OK) return rc;
    k_unstash_loss<<<1, 1, 0, sm>>>(dZ + nz + 2, a.loss_acc);
    RCD_LAUNCH_CHECK();
    STEP_CALL("rcd_act_grad", st, rcd_act_grad(dZ, Z, (long long)rows * H, a.act, dA, st));
    STEP_CALL("rcd_colsum", st, rcd_colsum(dA, rows, H, H, db_in, st));
  } else {
    // MF: dU = dZ * act'(Ue) straight into the slab's user block
    STEP_CALL("rcd_decoder_dgrad", st,
              rcd_decoder_dgrad(G, ldn, Wg, ldh, rows, n, H, splits, partials, H, RCD_GEMM_TCGEN05, st));
    STEP_CALL("rcd_dz_act", st, rcd_dz_act(partials, splits + 1, splits, alpha, H, Z, rows, H, a.act, dW_in, nullptr, st));
  }

  if (ae) {
    const int k = a.same_pool ? 0 : 1;
    const size_t hb = heavy ? rcd_csc_heavy_scratch_bytes(n_in, innz, H) : 0;
    STEP_CALL("rcd_ae_encoder_wgrad", st,
              rcd_ae_encoder_wgrad(dA, H, reinterpret_cast<int32_t*>(ws + L.csc_ptr[k]),
                                   reinterpret_cast<int32_t*>(ws + L.csc_row[k]),
                                   reinterpret_cast<float*>(ws + L.csc_val[k]), a.in.row_inv_norm, row0, n_in, dW_in,
                                   nullptr, nullptr, heavy, hb, innz, st));
    int rc = opt_step(c, a, a.table_in, dW_in, H, a.in.pos, st, a.in.items, a.in.items ? n_in : a.table_in.rows);
    if (rc != RCD_OK) return rc;
    rc = opt_step(c, a, a.bias_in, db_in, 1, nullptr, st);
    if (rc != RCD_OK) return rc;
  } else {
    RCD_CHECK_ARG(a.user_pos, "MF needs user_pos");
    const int64_t* users = a.in.users + row0;
    STEP_CALL("rcd_scatter_pos", st, rcd_scatter_pos(users, rows, a.user_pos, 0, st));
    int rc = opt_step(c, a, a.table_in, dW_in, H, a.user_pos, st, users, rows);
    if (rc != RCD_OK) return rc;
    STEP_CALL("rcd_scatter_pos", st, rcd_scatter_pos(users, rows, a.user_pos, 1, st));
  }
  return RCD_OK;
}

/* recoder_b200 — C ABI of the B200-native Recoder training hot path.
 *
 * The reference (amoussawi/recoder @ a9ed3e8) has no FFI of its own: its extension points are Python classes
 * (SURVEY.md §8b).  This header is the boundary a maintainer binds underneath those classes (ctypes stub in
 * INTEGRATION.md).  Every entry point cites the reference code it replaces (paths relative to the reference
 * root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - every function is asynchronous and ordered on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, negative = error (RCD_ERR_*); `rcd_last_error()` gives the message of the last
 *     failure on the calling thread; nothing throws, nothing allocates caller-visible memory;
 *   - scratch memory is provided by the caller (`*_scratch_bytes` helpers give the size);
 *   - row-major everywhere; "ld" arguments are leading dimensions in ELEMENTS;
 *   - bf16 operands are `uint16_t` bit patterns (same layout as __nv_bfloat16 / torch.bfloat16).
 */
#ifndef RECODER_B200_H_
#define RECODER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCD_ABI_VERSION 1

#define RCD_OK 0
#define RCD_ERR_INVALID (-1) /* bad argument */
#define RCD_ERR_CUDA (-2)    /* CUDA runtime / driver error */
#define RCD_ERR_UNSUPPORTED (-3)

/* activation ids — recoder/nn.py:6-9 `activation(x, act)` */
#define RCD_ACT_NONE 0
#define RCD_ACT_TANH 1
#define RCD_ACT_SIGMOID 2
#define RCD_ACT_RELU 3
/* further `torch.<name>` unary functions whose derivative is a function of the OUTPUT (what the backward kernels keep) */
#define RCD_ACT_SELU 4
#define RCD_ACT_CELU 5       /* alpha = 1 (torch.celu default) */
#define RCD_ACT_HARDSHRINK 6 /* lambda = 0.5 */
#define RCD_ACT_ATAN 7
#define RCD_ACT_SINH 8
#define RCD_ACT_ASINH 9
#define RCD_ACT_EXPM1 10

/* loss ids — recoder/model.py:87-99 `__init_loss_module` */
#define RCD_LOSS_MSE 0      /* recoder/losses.py:16-47  MSELoss(confidence, 'sum') */
#define RCD_LOSS_NLL 1      /* recoder/losses.py:50-71  MultinomialNLLLoss('sum') ('logloss') */
#define RCD_LOSS_LOGISTIC 2 /* torch BCEWithLogitsLoss('sum') ('logistic'), recoder/model.py:90-91 */

/* rcd_decoder_fwd_loss modes; multinomial-NLL exponent clamp (see K4 / K5 below) */
#define RCD_DEC_MODE_LOSS 0    /* loss / dL/dlogits epilogue */
#define RCD_DEC_MODE_ROWMAX 1  /* NLL redo pass: stat = per-tile row maxima of logit*log2(e), nothing else is written */
#define RCD_NLL_CLAMP_LOG2 64  /* G = exp(o - ref) is clamped at 2^64; rows whose sum reaches it are redone */

/* GEMM engines: the tcgen05/TMA kernels are the product; the SIMT engine is a slow reference of the same
 * math (same bf16 operands, fp32 accumulation) used by the tests to localise faults. */
#define RCD_GEMM_TCGEN05 0
#define RCD_GEMM_SIMT 1

int rcd_abi_version(void);
const char* rcd_last_error(void);
/* number of SMs of the current device (148 on B200); <0 on error */
int rcd_device_sms(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py's `gpu_launches`) */
long long rcd_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * K0  host staging — replaces the SciPy fancy row indexing of RecommendationDataset._extract
 *     (recoder/data.py:66-81) when the interaction matrix stays in HOST memory: copies the CSR rows `users`
 *     (stored order) into caller-provided (pinned) buffers that are then shipped H2D and collated by K1 with
 *     users = 0..P-1.  ALL pointers are HOST pointers.  Returns the pool's nnz, or a negative RCD_ERR_*.
 * ------------------------------------------------------------------------------------------------------- */
long long rcd_host_stage_rows(const int64_t* indptr_host, const int32_t* indices_host, const float* data_host,
                              const int64_t* users_host, int pool_rows, long long num_users, long long capacity,
                              int64_t* row_ptr_out_host, int32_t* indices_out_host, float* data_out_host);

/* ---------------------------------------------------------------------------------------------------------
 * K1  collate — replaces RecommendationDataset.__getitem__/_extract (recoder/data.py:50-83) and
 *     BatchCollator.collate (recoder/data.py:203-251) plus the COO->dense scatter at recoder/model.py:457-462.
 *
 * Inputs : dataset CSR (indptr int64[U+1], indices int32[], data fp32[]), the pool's user ids (int64[P]).
 * Outputs: row_ptr int32[P+1] (exclusive scan of the pool rows' nnz), raw_items int32[nnz] (global item id),
 *          cols int32[nnz] (position of the item in `items`, or the raw id without negative sampling),
 *          vals fp32[nnz] (stored CSR order inside each row — data.py:236-242),
 *          row_inv_norm fp32[P] = 1/max(||x_u||_2, 1e-12)  (F.normalize, recoder/nn.py:235),
 *          row_sum fp32[P] = sum_j x_uj (needed by the multinomial NLL gradient),
 *          pos int32[num_items] (item -> column or -1), items int64[>=n] sorted ascending unique
 *          (np.unique, data.py:220), counts int32[2] = {n, nnz}.
 * `nnz_capacity` is the capacity of raw_items/cols/vals; the pool's nnz must not exceed it.
 * ------------------------------------------------------------------------------------------------------- */
size_t rcd_collate_scratch_bytes(int pool_rows, int num_items);
int rcd_collate(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* users,
                int pool_rows, int num_items, int negative_sampling, int nnz_capacity, int32_t* row_ptr,
                int32_t* raw_items, int32_t* cols, float* vals, float* row_inv_norm, float* row_sum, int32_t* pos,
                int64_t* items, int32_t* counts, void* scratch, size_t scratch_bytes, void* stream);

/* Batch.indices of one slice (recoder/data.py:244): indices int64[2, nnz_slice], rows relative to row0. */
int rcd_collate_coo(const int32_t* row_ptr, const int32_t* cols, int row0, int rows, int64_t* indices_out,
                    void* stream);

/* Column-major view (CSC) of one slice of the pool, used by the weight gradients:
 * csc_ptr int32[n+1], csc_row int32[nnz_slice] (row relative to row0, ascending inside a column),
 * csc_val fp32[nnz_slice] (raw interaction value), csc_src int32[nnz_slice] (position of the entry in the slice's
 * CSR order, i.e. relative to row_ptr[row0]; optional). */
size_t rcd_slice_csc_scratch_bytes(int n, int nnz_slice);
int rcd_slice_csc(const int32_t* row_ptr, const int32_t* cols, const float* vals, int row0, int rows, int n,
                  int32_t* csc_ptr, int32_t* csc_row, float* csc_val, int32_t* csc_src, void* scratch,
                  size_t scratch_bytes, void* stream);

/* dense [rows, n] fp32 input (the `input` argument of FactorizationModel.forward, recoder/nn.py:49-65)
 * -> CSR of its non-zeros (row_ptr int32[rows+1], cols, vals, row_inv_norm, row_sum). */
size_t rcd_dense_to_csr_scratch_bytes(int rows);
int rcd_dense_to_csr(const float* dense, int rows, int n, int ld, int nnz_capacity, int32_t* row_ptr,
                     int32_t* cols, float* vals, float* row_inv_norm, float* row_sum, int32_t* nnz_out,
                     void* scratch, size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K2  embedding gather — replaces nn.Embedding lookups / index_select in LinearEmbedding.forward
 *     (recoder/nn.py:271-272) and MatrixFactorization.forward (recoder/nn.py:348,358-359).
 *     out_bf16[r, 0:H] = bf16(act(table[ids[r], :])), zero padded to ld_out; optional fp32 copy.
 *     ids == NULL means the identity (full table, `items is None`, nn.py:273-275).
 * ------------------------------------------------------------------------------------------------------- */
int rcd_gather_rows(const float* table, int H, const int64_t* ids, int n, int act, uint16_t* out_bf16,
                    int ld_out, float* out_f32, void* stream);
int rcd_gather_vec(const float* vec, const int64_t* ids, int n, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K3  autoencoder encoder forward — replaces F.normalize + LinearEmbedding(en) + activation
 *     (recoder/nn.py:235-240, 269-278) and the dense materialisation of the input (recoder/model.py:457-458).
 *     Z[r,:] = act( keep_scale * row_inv_norm[r] * sum_p vals[p] * We[raw_items[p], :] + be )
 *     for the slice rows [row0, row0+rows).  Sparse formulation: 2*nnz*H executed flops, fp32 exact.
 *     Writes Z fp32 [rows, H] and a bf16 copy [rows, ldzb] (zero padded) that feeds the decoder GEMM.
 * ------------------------------------------------------------------------------------------------------- */
int rcd_ae_encoder_fwd(const float* We, int H, const float* be, const int32_t* row_ptr, const int32_t* raw_items,
                       const float* vals, const float* row_inv_norm, int row0, int rows, int act, float* Z,
                       uint16_t* Zb, int ldzb, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K4  decoder forward fused with the loss — replaces LinearEmbedding(de).forward
 *     `F.linear(z, W_d[items], b_d[items])` (recoder/nn.py:280; MF: recoder/nn.py:361), MSELoss.forward
 *     (recoder/losses.py:43-47), MultinomialNLLLoss.forward (recoder/losses.py:68-71), BCEWithLogitsLoss('sum')
 *     (recoder/model.py:91), the `/ B` (recoder/model.py:483-484) and the first node of autograd's backward.
 *     Logits o[r,c] = sum_h Zb[r,h]*Wg[c,h] + bias[c] (bf16 operands, fp32 accumulate in TMEM) never reach memory;
 *     the epilogue writes G bf16 [rows, ldg], the DENSE (target-free) part of dL/dlogits:
 *        MSE      G = 2*o/B                      dL/dO = G + sparse
 *        LOGISTIC G = sigmoid(o)/B               dL/dO = G + sparse
 *        NLL      G = exp(o - row_ref[r])        dL/dO = alpha[r]*G + sparse,  alpha[r] = S_r/(B*sum_c G[r,c])
 *     and per-row partials stat fp32 [rows, stat_ld] (rcd_decoder_stat_cols(n) of them: 2 per 256-column tile): sum G
 *     (NLL), sum o^2 (MSE),
 *     sum softplus(o) (LOGISTIC).  The SPARSE part at the stored targets (fp32, never quantised to bf16) comes
 *     from rcd_sddmm:  MSE 2*((w-1)*o - w*t)/B with w = 1+conf*[t>0];  NLL / LOGISTIC -t/B.
 *     row_ref (NLL): any per-row reference; rcd_sddmm supplies the largest logit among the row's own targets.
 *     That is not an upper bound of the row, so G is clamped at 2^RCD_NLL_CLAMP_LOG2 (everything stays finite) and
 *     rows whose sum reaches the clamp are REDONE ON THE DEVICE with their true maximum as reference, which makes
 *     the loss as unconditionally stable as F.log_softmax (recoder/losses.py:69):
 *        rcd_decoder_fwd_loss(mode LOSS) -> rcd_loss_finish(redo_flag, row_redo)            every step
 *        rcd_decoder_fwd_loss(mode ROWMAX, cond = redo_flag) -> rcd_nll_ref_fix(cond) ->
 *        rcd_decoder_fwd_loss(mode LOSS, cond) -> rcd_loss_finish(cond)                     no-ops while *cond == 0
 *        rcd_loss_sum (adds the per-block loss partials to loss_acc in fixed order, clears redo_flag)
 *     `cond` (device int32, may be NULL = always run): the kernel returns immediately when *cond == 0.
 *
 *     rcd_decoder_fwd is the plain logits GEMM (fp32 or bf16 out, optional online-softmax partials laid out
 *     [n_tiles, rows]) used by the inference path (recoder/model.py:487-511) and the kernel tests.
 * ------------------------------------------------------------------------------------------------------- */
int rcd_decoder_tile_n(void); /* n-tile width of rcd_decoder_fwd's stats (256) */
int rcd_decoder_fwd(const uint16_t* Zb, int ldzb, const uint16_t* Wg, int ldw, const float* bias, int rows, int n,
                    int H, uint16_t* O_bf16, float* out_f32, int ldo, float* stat_max, float* stat_sum, int engine,
                    void* stream);
int rcd_decoder_stat_cols(int n); /* number of per-row partials rcd_decoder_fwd_loss writes for n items */
int rcd_decoder_fwd_loss(const uint16_t* Zb, int ldzb, const uint16_t* Wg, int ldw, const float* bias, int rows, int n,
                         int H, int loss, float inv_b, const float* row_ref, uint16_t* G, int ldg, float* stat,
                         int stat_ld, int mode, const int32_t* cond, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K5  sparse side of the loss (fp32) — the stored targets of the slice rows [row0, row0+rows).
 *     rcd_sddmm       : o_nnz[p] = Zb[r,:].Wg[cols[p],:] + bias_g[cols[p]], corr[p] = sparse part of dL/dlogits,
 *                       row_ref[r] = max_p o_nnz[p] (0 for an empty row; optional).  p is relative to
 *                       row_ptr[row0].  cols index Wg/bias_g (gathered rows).
 *     rcd_loss_finish : reduces stat, finishes the loss of the slice (divided by B), writes
 *                       row_scale[r] = alpha[r] (1 for MSE/LOGISTIC) and, when Zs != NULL, Zs = bf16(alpha*Z)
 *                       [rows, ldzs] — the operand of the decoder weight gradient.  The loss goes to
 *                       loss_blocks[block] (double, one per 8 rows; summed by rcd_loss_sum in fixed order) or, when
 *                       loss_blocks == NULL, straight into loss_acc with one atomic per block.
 *                       NLL rows whose row sum reached 2^RCD_NLL_CLAMP_LOG2 (a logit far above the reference): with
 *                       redo_flag != NULL they set row_redo[r] = 1 and *redo_flag (see K4); otherwise, and for any
 *                       row sum that is not positive/finite, bad_flag |= 1.  bad_flag |= 2 when the loss is not
 *                       finite.  Rows without targets contribute 0 (never 0 * log 0).  local_targets != 0
 *                       (item-parallel mode): stat is the row sum over ALL item shards, the stored targets are this
 *                       rank's shard only, and the loss added is this rank's share.
 *     rcd_nll_ref_fix : row_ref[r] = ln2 * max_t stat[r,t] for the rows with row_redo[r] != 0 (after a ROWMAX pass)
 *     rcd_loss_sum    : loss_acc += sum of loss_blocks[0..nblocks) in index order; *redo_flag = 0 (if given)
 *     rcd_sparse_dgrad: out[r,0:H] = sum_p corr[p] * W[raw_items[p],:]  (fp32 master table; out fp32 [rows, ldp])
 *     rcd_csc_rows_accumulate: out[c,0:H] += sum_{e in column c} coef[csc_src[e]] * M[csc_row[e],:] and
 *                       db[c] += sum_e coef[csc_src[e]]  (csc_src == NULL: coef is already in CSC order)
 * ------------------------------------------------------------------------------------------------------- */
int rcd_sddmm(const uint16_t* Zb, int ldzb, const uint16_t* Wg, int ldw, const float* bias_g, int H,
              const int32_t* row_ptr, const int32_t* cols, const float* vals, int row0, int rows, int loss,
              float confidence, float inv_b, float* o_nnz, float* corr, float* row_ref, void* stream);
int rcd_loss_finish(const float* stat, int stat_ld, int stat_cols, int rows, int loss, float confidence, float inv_b,
                    const float* row_ref, const float* row_sum, const int32_t* row_ptr, const float* vals,
                    const float* o_nnz, int row0, float* row_scale, const float* Z, int H, uint16_t* Zs, int ldzs,
                    double* loss_acc, int32_t* bad_flag, int local_targets, double* loss_blocks, int32_t* redo_flag,
                    int32_t* row_redo, const int32_t* cond, void* stream);
int rcd_loss_finish_blocks(int rows); /* number of loss_blocks entries rcd_loss_finish writes for `rows` rows */
int rcd_nll_ref_fix(const float* stat, int stat_ld, int stat_cols, int rows, const int32_t* row_redo, float* row_ref,
                    const int32_t* cond, void* stream);
int rcd_loss_sum(const double* loss_blocks, int nblocks, double* loss_acc, int32_t* redo_flag, void* stream);
int rcd_sparse_dgrad(const float* W, int H, const int32_t* row_ptr, const int32_t* raw_items, const float* corr,
                     int row0, int rows, float* out, int ldp, void* stream);
int rcd_csc_rows_accumulate(const float* M, int H, const int32_t* csc_ptr, const int32_t* csc_row,
                            const int32_t* csc_src, const float* coef, int n, float* out, float* db, void* scratch,
                            size_t scratch_bytes, long long nnz_slice, void* stream);
/* Workspace of the two column-major accumulations (rcd_csc_rows_accumulate, rcd_ae_encoder_wgrad): item popularity is
 * a power law, so a few columns hold an entry in nearly every row; columns with more than 128 entries are processed
 * in 128-entry chunks by separate thread groups and summed in chunk order.  scratch == NULL: no chunking. */
size_t rcd_csc_heavy_scratch_bytes(int n, long long nnz_slice, int H);

/* ---------------------------------------------------------------------------------------------------------
 * K6  decoder backward GEMMs — replace autograd's `mm` nodes of F.linear.
 *     rcd_decoder_dgrad : P[r,h]  = sum_c G[r,c] * Wg[c,h]      (K = n, split-K; partials fp32
 *                         [splits, rows, ldp]); the caller reduces/scales them with rcd_dz_act.
 *     rcd_decoder_wgrad : dW[c,h] = sum_r G[r,c] * Zs[r,h]      (K = rows) -> fp32 [n, H] compact row grads;
 *                         side product (optional): db[c] = sum_r col_weight[r] * G[r,c]  (col_weight NULL = 1),
 *                         computed from the operand tiles already staged in shared memory.
 * ------------------------------------------------------------------------------------------------------- */
int rcd_decoder_dgrad_splits(int rows, int n, int H);
int rcd_decoder_dgrad(const uint16_t* G, int ldg, const uint16_t* Wg, int ldw, int rows, int n, int H, int splits,
                      float* partials, int ldp, int engine, void* stream);
int rcd_decoder_wgrad(const uint16_t* G, int ldg, const uint16_t* Zs, int ldzs, int rows, int n, int H, float* dW,
                      int lddw, const float* col_weight, float* db, int engine, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K7  encoder backward — replaces tanh_backward + the `mm`/`sum` nodes of the encoder F.linear and
 *     embedding_dense_backward.
 *     rcd_dz_act        : dA = (row_scale[r] * sum_{s<n_scaled} partials[s] + sum_{s>=n_scaled} partials[s]) * act'(Z)
 *                         -> fp32 [rows, H] (row_scale NULL = 1); db_e[h] = sum_r dA[r,h]
 *     rcd_ae_encoder_wgrad : dWe_rows[c,:] = sum_{(r,x) in column c} x * row_inv_norm[row0+r] * dA[r,:]
 *                         (csc_src + csr_vals, optional: take x from csr_vals[csc_src[e]] — the input values after
 *                         the input-noise dropout, in the slice's CSR order — instead of csc_val[e])
 * ------------------------------------------------------------------------------------------------------- */
int rcd_dz_act(const float* partials, int splits, int n_scaled, const float* row_scale, int ldp, const float* Z,
               int rows, int H, int act, float* dA, float* db, void* stream);
int rcd_ae_encoder_wgrad(const float* dA, int H, const int32_t* csc_ptr, const int32_t* csc_row,
                         const float* csc_val, const float* row_inv_norm, int row0, int n, float* dWe_rows,
                         const int32_t* csc_src, const float* csr_vals, void* scratch, size_t scratch_bytes,
                         long long nnz_slice, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K8  optimizers — replace torch.optim.{Adam,SGD,SparseAdam}.step as configured by
 *     Recoder.__init_optimizer (recoder/model.py:101-164); per-parameter weight decay (0 for biases,
 *     model.py:123-124).  Gradients arrive as compact row blocks: the gradient of table row i is
 *     grad_rows[pos[i], :] when pos[i] >= 0 and exactly zero otherwise (pos == NULL: grad is dense [rows,H]).
 *     Dense semantics: EVERY row of the table is updated (momentum and weight decay move untouched rows).
 *     `t` is the 1-based step count of this parameter.
 * ------------------------------------------------------------------------------------------------------- */
int rcd_adam_step(float* p, float* m, float* v, long long rows, int H, const float* grad_rows, int ldg,
                  const int32_t* pos, double lr, double beta1, double beta2, double eps, double weight_decay,
                  long long t, void* stream);
int rcd_sgd_step(float* p, float* buf, long long rows, int H, const float* grad_rows, int ldg, const int32_t* pos,
                 double lr, double momentum, double weight_decay, void* stream);
/* Deferred dense Adam — the SAME arithmetic and results as rcd_adam_step (bit-identical), with HBM traffic proportional
 * to the rows a batch touches instead of the whole table.  Rows outside a batch have gradient wd*p whatever the batch
 * is, so their updates are postponed: last int32[rows] holds the step up to which each row is current.
 *   rcd_adam_lazy_catchup : replays the skipped steps last[r]+1 .. T (zero data gradient) for rows ids[0..n) (ids NULL:
 *                           rows 0..n-1, i.e. a flush of the table) from the per-step scalars scal float[2*scal_len]
 *                           = {lr_t/(1-beta1^t), 1/sqrt(1-beta2^t)} of steps scal_base .. scal_base+scal_len-1
 *                           (rcd_adam_scalars); mark != 0: then sets last[r] = T.  Call it before anything reads the rows.
 *                           n_dev (optional): device int32 holding the row count (n is then an upper bound: the
 *                           collate of the NEXT pool has not been read back when its rows are caught up ahead of
 *                           time); exclude_pos (optional): rows r with exclude_pos[r] >= 0 are skipped — the rows of
 *                           the batch in flight, whose step is still being applied.
 *   rcd_adam_lazy_update  : step t on rows ids[i] (current at t-1) with gradient row grad_rows[i,:]; last[ids[i]] = t.
 *   rcd_adam_scalars      : HOST helper filling out_host[2*count] for steps t_first .. t_first+count-1 exactly as
 *                           rcd_adam_step forms its scalars (double arithmetic, rounded to float). */
int rcd_adam_lazy_catchup(float* p, float* m, float* v, int H, const int64_t* ids, long long n, int32_t* last,
                          long long T, const float* scal, long long scal_base, long long scal_len, double beta1,
                          double beta2, double eps, double weight_decay, int mark, const int32_t* n_dev,
                          const int32_t* exclude_pos, void* stream);
int rcd_adam_lazy_update(float* p, float* m, float* v, int H, const int64_t* ids, long long n, const float* grad_rows,
                         int ldg, int32_t* last, double lr, double beta1, double beta2, double eps, double weight_decay,
                         long long t, void* stream);
int rcd_adam_scalars(double lr, double beta1, double beta2, long long t_first, int count, float* out_host);
/* torch.optim.Adagrad(lr) and torch.optim.RMSprop(lr, momentum=0.9) with their defaults (lr_decay 0, eps 1e-10 /
 * alpha 0.99, eps 1e-8, not centered) — recoder/model.py:140-144, 150-154; dense semantics as above. */
int rcd_adagrad_step(float* p, float* sum, long long rows, int H, const float* grad_rows, int ldg, const int32_t* pos,
                     double lr, double eps, double weight_decay, void* stream);
int rcd_rmsprop_step(float* p, float* square_avg, float* buf, long long rows, int H, const float* grad_rows, int ldg,
                     const int32_t* pos, double lr, double alpha, double eps, double momentum, double weight_decay,
                     void* stream);
/* torch.optim.SparseAdam on the n rows `ids` (no weight decay; recoder/model.py:137-138) */
int rcd_sparse_adam_step(float* p, float* m, float* v, int H, const float* grad_rows, int ldg, const int64_t* ids,
                         int n, double lr, double beta1, double beta2, double eps, long long t, void* stream);
/* pos[ids[r]] = r (or -1 to reset) — inverse map for row-indexed gradients (MF user table) */
int rcd_scatter_pos(const int64_t* ids, int n, int32_t* pos, int reset, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K9  data-parallel exchange over peer memory (NVLink 5 / NVSwitch) — no reference counterpart: the reference
 *     is single-device.  One process per GPU; buffers peers must reach are cudaMalloc
 *     allocations shared with CUDA IPC.  Pointer tables (`*_host`) are HOST arrays indexed by rank holding the
 *     address of the same buffer as mapped in THIS process (own rank: the local allocation).
 *     rcd_p2p_alloc/free   : zero-filled device allocation that can be exported (synchronous host calls)
 *     rcd_p2p_export/open/close : 64-byte IPC handle of an allocation / mapping of a peer's handle
 *     rcd_p2p_barrier      : stream-ordered barrier of all ranks; `flags` = per-rank uint32[RCD_MAX_PEERS] arrays,
 *                            `seq` strictly increasing per call; bad_flag |= 4 if a peer does not arrive in timeout_s
 *     rcd_p2p_reduce       : dst[i] = sum (or max) over ranks q (ascending) of src_q[offset + i], i < count (fp32)
 *     rcd_adam_step_p2p    : fused reduce-scatter -> Adam -> all-gather.  This rank owns table rows
 *                            [row_begin, row_end): g = sum_q grads_q[pos[row], :] (rank order; only rank
 *                            pos[row]/grad_block_rows when grad_block_rows > 0), torch.optim.Adam update of the local
 *                            p/m/v rows (same arithmetic as rcd_adam_step), new p row stored into EVERY rank's table.
 *                            Callers bracket it with rcd_p2p_barrier (all slabs written / all pushes landed).
 *                            grads_mc / table_mc (optional, NULL = off): NVSwitch MULTICAST addresses of the gradient
 *                            block / the table (same memory as the per-rank pointers, bound to one multicast object):
 *                            the sum becomes one `multimem.ld_reduce` (reduced inside the switch) and the replica
 *                            update one `multimem.st` — link traffic per GPU drops from (world-1) x to 1 x.
 * ------------------------------------------------------------------------------------------------------- */
#define RCD_MAX_PEERS 16
#define RCD_P2P_HANDLE_BYTES 64
int rcd_p2p_alloc(size_t bytes, void** out_host);
int rcd_p2p_free(void* p);
int rcd_p2p_export(const void* p, unsigned char* handle_host);
int rcd_p2p_open(const unsigned char* handle_host, void** out_host);
int rcd_p2p_close(void* p);
int rcd_p2p_barrier(void* const* flags_host, int rank, int world, unsigned int seq, int32_t* bad_flag,
                    double timeout_s, void* stream);
#define RCD_REDUCE_SUM 0
#define RCD_REDUCE_MAX 1
int rcd_p2p_reduce(const float* const* src_host, int world, long long offset, long long count, float* dst, int op,
                   void* stream);
/* two-shot all-reduce (sum, fp32) in place over a buffer every rank has mapped (count a multiple of 4, 16-byte aligned):
 * this rank reduces its 1/world slice (multimem.ld_reduce through `mc`, the multicast address, or plain peer loads when
 * mc == NULL) and stores the result into every rank's copy.  Bracket with rcd_p2p_barrier. */
int rcd_p2p_allreduce(float* const* bufs_host, float* mc, long long count, int rank, int world, void* stream);
int rcd_adam_step_p2p(float* const* tables_host, float* m, float* v, long long row_begin, long long row_end, int H,
                      const float* const* grads_host, int ldg, const int32_t* pos, int grad_block_rows, int rank,
                      int world, double lr, double beta1, double beta2, double eps, double weight_decay, long long t,
                      const float* grads_mc, float* table_mc, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K10 generalised model pieces: inner dense layers of a multi-layer DynamicAutoencoder
 *     (recoder/nn.py:189-226, 242-249), input noise / bottleneck / user-embedding dropout (nn.py:236-237, 245-246,
 *     351-352) and the elementwise nodes of their backward.
 *     rcd_sgemm   : fp32 C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N]) -> act; trans_a: A stored [K, lda>=M];
 *                   trans_b: B stored [N, ldb>=K] (an nn.Linear weight [out, in] is B with trans_b = 1);
 *                   accumulate != 0: C += result (tied weights receive two gradient contributions)
 *     rcd_dropout : y[i] = keep_i ? x[i]/(1-p) : 0; keep_i from Philox4x32-10(seed, rng_stream) at element index
 *                   index_base + i, or from keep_mask[i] (uint8, tests).  Apply it to the gradient for the backward.
 *     rcd_act_grad: dpre = dy * act'(y) with y the activation OUTPUT (in place allowed)
 *     rcd_colsum  : out[h] = sum_r x[r, h]
 *     rcd_f32_to_bf16_rows : bf16 copy [rows, ld] (zero padded) of an fp32 [rows, H] matrix
 * ------------------------------------------------------------------------------------------------------- */
int rcd_sgemm(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
              int ldc, const float* bias, int act, int accumulate, void* stream);
int rcd_dropout(const float* x, long long count, float p, unsigned long long seed, unsigned int rng_stream,
                long long index_base, const uint8_t* keep_mask, float* y, void* stream);
int rcd_act_grad(const float* dy, const float* y, long long count, int act, float* dpre, void* stream);
int rcd_colsum(const float* x, int rows, int H, int ld, float* out, void* stream);
int rcd_f32_to_bf16_rows(const float* x, int rows, int H, uint16_t* out, int ld, void* stream);
/* item-parallel mode helpers: Z = act(x + bias) (fp32 [rows,H] + optional bf16 copy [rows, ld]) after the all-reduce
 * of the encoder partial sums; out[r] = sum_c x[r, c] */
int rcd_bias_act(const float* x, const float* bias, int rows, int H, int act, float* Z, uint16_t* Zb, int ld,
                 void* stream);
int rcd_rowsum(const float* x, int rows, int cols, int ld, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * K11 recommendation — replaces the tail of Recoder.recommend (recoder/model.py:525-544):
 *     `output[input > 0] = -inf` and `torch.topk(output, k, dim=1, sorted=True)`.
 *     rcd_mask_seen : logits[r, items[p]] = -inf for the stored interactions p of pool rows [row0, row0+rows)
 *                     (row_ptr / items = the collate outputs row_ptr / raw_items); logits fp32 [rows, ld]
 *     rcd_topk_rows : per row the k largest of n logits, descending (ties: lower index first);
 *                     out_val fp32 [rows, k], out_idx int64 [rows, k]; 1 <= k <= min(n, 1024)
 * ------------------------------------------------------------------------------------------------------- */
int rcd_mask_seen(const int32_t* row_ptr, const int32_t* items, int row0, int rows, float* logits, long long ld,
                  void* stream);
int rcd_topk_rows(const float* logits, long long ld, int rows, int n, int k, float* out_val, int64_t* out_idx,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Telemetry / tests: L2 norm squared of a strided fp32 matrix (double accumulation), out_sq[0] += ...
 * ------------------------------------------------------------------------------------------------------- */
int rcd_sumsq(const float* x, long long rows, int cols, int ld, double* out_sq, void* stream);

/* Plain bf16 GEMM entry used by the kernel unit tests (and by the inner MLP layers):
 *   mode 0: C[M,N] = A[M,K] * B[N,K]^T        (A, B K-major)
 *   mode 1: C[M,N] = A[M,K] * B[K,N]          (B MN-major)
 *   mode 2: C[M,N] = A[K,M]^T * B[K,N]        (A, B MN-major)
 * C fp32 [M, ldc]. */
int rcd_gemm_bf16(int mode, const uint16_t* A, int lda, const uint16_t* B, int ldb, int M, int N, int K, float* C,
                  int ldc, int engine, void* stream);


/* ---------------------------------------------------------------------------------------------------------
 * K12 native step executor — replaces the per-batch body of Recoder._train (recoder/model.py:383-404:
 *     zero_grad -> __compute_loss -> backward -> optimizer.step) as ONE host call.  The Python engine issues the
 *     same ~40 entry points of this header one ctypes call at a time (≈1 ms of host time per step: what bounds
 *     the small configurations, where the GPU needs 0.1-0.3 ms); rcd_step_run issues them from C++ in the same
 *     order on the same three streams, so both paths are bit-identical.
 *     Covers single-hidden-layer autoencoders and matrix factorisation with a fused loss (MSE / NLL / logistic)
 *     and a dense optimizer on one GPU, and the item-parallel autoencoder step of the multi-GPU mode (`ip`).
 *     Streams: main (forward, GEMMs, encoder backward, input-table update), side (output-table update, as soon
 *     as its gradient is complete), aux (column-major views of the slice).  overlap == 0: everything on main.
 *     Workspace: ONE caller-provided device buffer, carved by CAPACITIES (cap_rows, cap_n, cap_nnz) so that the
 *     layout is stable from step to step; rcd_step_workspace_bytes says how large it must be.  The caller grows
 *     it (after a device synchronisation) when a step exceeds a capacity.
 *     rcd_step_run fills `out_*` with BYTE OFFSETS into the workspace of the step's compact gradients.
 *     rcd_step_join makes `stream` wait for the side-stream work of the last step.
 *     rcd_step_profile(ctx, mode): 0 off, 1 CUDA events around every entry point, 2 around the entry point named
 *     `name` only; rcd_step_profile_read sums the event pairs per entry point (synchronises) and clears them.
 * ------------------------------------------------------------------------------------------------------- */
#define RCD_STEP_ABI 4
#define RCD_MODEL_AE 0
#define RCD_MODEL_MF 1
#define RCD_OPT_ADAM 0
#define RCD_OPT_SGD 1
#define RCD_OPT_ADAGRAD 2
#define RCD_OPT_RMSPROP 3

typedef struct rcd_param {
  float* p;            /* parameter, [rows, cols] row-major */
  float* s1;           /* Adam exp_avg | SGD momentum buffer | Adagrad sum | RMSprop square_avg */
  float* s2;           /* Adam exp_avg_sq | RMSprop momentum buffer | unused */
  long long rows;
  int cols;
  int pad_;
  double weight_decay;
  long long t;         /* 1-based step count of THIS step (Adam bias correction) */
  int32_t* last;       /* deferred dense Adam (rcd_adam_lazy_*): per-row step counters, or NULL = dense update */
} rcd_param;

typedef struct rcd_pool_view { /* outputs of rcd_collate for one pool (all device pointers) */
  const int32_t* row_ptr;
  const int32_t* raw_items;
  const int32_t* cols;
  const float* vals;
  const float* row_inv_norm;
  const float* row_sum;
  const int32_t* pos;
  const int64_t* items;   /* NULL without negative sampling (columns are raw item ids) */
  const int64_t* users;
  long long nnz_slice;    /* stored entries of the slice rows [row0, row0+rows) */
  int n;                  /* number of columns (batch items) */
  int pad_;
} rcd_pool_view;

typedef struct rcd_step_ip { /* item-parallel mode (world > 1): peer-mapped buffers of the four collectives */
  int enabled, rank, world, use_nccl_;     /* use_nccl_ must be 0 */
  void* const* flags_host;                 /* barrier flags, per-rank pointer table (HOST array) */
  unsigned int* seq_host;                  /* HOST counter shared with the caller's other barriers */
  double barrier_timeout_s;
  float* shared_local;                     /* this rank's copy of the shared block */
  float* const* shared_host;               /* per-rank pointer table of the shared block (HOST array) */
  float* shared_mc;                        /* multicast address of the shared block or NULL */
  long long off_z, off_dz, off_ref, off_sum;  /* FLOAT offsets inside the shared block: Zp [rows*H], dZ [rows*H + 4],
                                                 row_ref [rows], row sums [rows] */
} rcd_step_ip;

typedef struct rcd_step_args {
  int abi;               /* RCD_STEP_ABI */
  int kind;              /* RCD_MODEL_AE | RCD_MODEL_MF */
  int H, act, loss, optimizer, train, overlap;
  float confidence, inv_b;
  double lr;
  rcd_param table_in;    /* AE: W_e [I,H]      MF: user table [U,D] */
  rcd_param bias_in;     /* AE: b_e [H]        MF: unused (p == NULL) */
  rcd_param table_out;   /* AE: W_d [I,H]      MF: item table [I,D] */
  rcd_param bias_out;    /* AE: b_d [I]        MF: bias [I] */
  rcd_pool_view in;      /* input pool */
  rcd_pool_view tgt;     /* target pool (a copy of `in` when the input is its own target, recoder/model.py:473-476) */
  int same_pool;
  int row0, rows;
  int cap_rows, cap_n, cap_n_in;
  long long cap_nnz, cap_tnnz;
  void* ws;
  size_t ws_bytes;
  double* loss_acc;
  int32_t* bad_flag;
  int32_t* redo_flag;
  int32_t* user_pos;     /* MF: int32 [num_users], all -1 between steps */
  const float* scal;     /* deferred dense Adam: per-step scalars (rcd_adam_lazy_catchup), steps scal_base .. +scal_len-1 */
  long long scal_base, scal_len;
  /* deferred Adam, optional: the rows of the NEXT pool that are not in this batch are caught up ahead of time on the side
   * stream (underneath the dgrad GEMM and the encoder backward) instead of at the start of the next step.  items: the next
   * pool's sorted item ids (device), n: its item count ON THE DEVICE (not read back yet), cap: an upper bound of it.  The
   * caller has made the side stream wait for the next pool's collate. */
  const int64_t* next_items_in;
  const int32_t* next_n_in;
  long long next_cap_in;
  const int64_t* next_items_out;
  const int32_t* next_n_out;
  long long next_cap_out;
  void* stream_main;
  void* stream_side;
  void* stream_aux;
  rcd_step_ip ip;
  /* outputs: byte offsets into ws (-1: not produced) */
  long long out_dW_in, out_db_in, out_dW_out, out_db_out;
} rcd_step_args;

size_t rcd_step_args_size(void); /* sizeof(rcd_step_args), checked by the binding */
int rcd_step_create(void** ctx_out);
int rcd_step_destroy(void* ctx);
size_t rcd_step_workspace_bytes(const rcd_step_args* args);
int rcd_step_run(void* ctx, rcd_step_args* args);
int rcd_step_join(void* ctx, void* stream);
int rcd_step_profile(void* ctx, int mode, const char* name);
/* names_out: '\n'-separated entry-point names (names_cap bytes), ms_out / count_out: per name; returns the number of
 * names or a negative RCD_ERR_* */
int rcd_step_profile_read(void* ctx, char* names_out, int names_cap, float* ms_out, int* count_out, int max_names);

#ifdef __cplusplus
}
#endif
#endif /* RECODER_B200_H_ */

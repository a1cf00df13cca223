/*
 * b200rec — C-ABI of the B200-native two-tower hot path.
 *
 * The reference (yxyxcyx/Real-Time-Recommendation-System-with-Feature-Store) is pure Python and has no FFI;
 * each entry point below replaces the LIBRARY CALL the reference makes at the cited line.  Conventions:
 *   - every pointer is a DEVICE pointer unless its name ends in _host; sizes are element counts;
 *   - `stream` is a cudaStream_t passed as void*; every launch is asynchronous on it;
 *   - nothing is allocated or freed inside; workspaces are caller-owned (query *_workspace_bytes first);
 *   - return 0 on success, non-zero on argument / CUDA error; b200rec_last_error() explains (thread-local);
 *   - bf16 operands are raw uint16 storage, row-major, leading dimension in ELEMENTS, a multiple of 64.
 */
#ifndef B200REC_H
#define B200REC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* b200rec_last_error(void);
int b200rec_version(void);
/* number of kernels launched by this library since load (bench.py's `gpu_launches`) */
int64_t b200rec_launch_count(void);

/* ---------------------------------------------------------------- operand preparation
 * fp32 -> bf16 tensor-core operand.  terms==1: plain round-to-nearest bf16 (bf16 mode).
 * terms==3: split-bf16 "fp32 mode": x = hi + lo; side 0 (left operand) writes [hi|lo|hi], side 1 (right
 * operand) writes [hi|hi|lo], each block `kpad` wide, so that  left . right^T = hi*hi + lo*hi + hi*lo.
 * transpose!=0 reads src as [cols, rows] (writes the transposed operand).  dst is [rows, terms*kpad], zero padded.
 * Replaces nothing in the reference (ATen consumes fp32 directly: two_tower.py:129,276,470). */
int b200rec_split_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, int transpose,
                       void* dst, int64_t kpad, int terms, int side, void* stream);

/* faiss.normalize_L2 (retrieval.py:86,167,214) / F.normalize(p=2,eps=1e-12) (two_tower.py:132,279), fused with the
 * operand cast: dst_f32 (nullable) gets the normalised fp32 rows, dst_bf16 (nullable) the bf16 operand
 * [rows, terms*kpad] as b200rec_split_bf16.  normalize==0 only casts.  faiss_zero_rule!=0 leaves zero rows
 * untouched (faiss); otherwise divides by max(norm, eps) (torch).  norms_out (nullable) receives max(norm,eps). */
int b200rec_normalize_rows(const float* src, int64_t rows, int64_t cols, int64_t ld_src, int normalize,
                           int faiss_zero_rule, float* dst_f32, int64_t ld_dst, float* norms_out, void* dst_bf16,
                           int64_t kpad, int terms, int side, void* stream);

/* ---------------------------------------------------------------- tensor-core GEMM (tcgen05 + TMA)
 * C[M,N] (+)= alpha * A[M,K] . B[N,K]^T + bias[N]   A,B bf16 K-major (lda/ldb elements, K multiple of 64),
 * C fp32 row-major.  k_splits>1 accumulates partial sums with fp32 atomics (C must be pre-zeroed by the caller).
 * Replaces ATen addmm/matmul: nn.Linear fwd/bwd two_tower.py:62,70,129,276 and matmul :470. */
int b200rec_gemm_bf16_tn(const void* A, int64_t lda, int64_t M, const void* B, int64_t ldb, int64_t N, int64_t K,
                         float* C, int64_t ldc, const float* bias, float alpha, int k_splits, void* stream);

/* ---------------------------------------------------------------- exact inner-product top-K (K4)
 * Replaces faiss.IndexFlatIP.search(q, k) (retrieval.py:171) and np.dot+argsort (scripts/evaluate_model.py:217-232).
 * catalogue: bf16 [N, ld] (ld = padded dim, multiple of 64), queries: bf16 [Q, ld].  out_scores fp32 [Q,k]
 * descending, out_ids int64 [Q,k] = row + row_offset; unfilled slots: -FLT_MAX / -1.  Order: score desc, row asc.
 * exclude_* (nullable): CSR of per-query catalogue rows (LOCAL row numbers, sorted ascending per query) that must
 * never be returned — the eval twin's -inf masking of train items. */
size_t b200rec_topk_workspace_bytes(int64_t N, int64_t ld, int64_t Q, int k);
int b200rec_flat_ip_topk(const void* catalogue, int64_t N, int64_t ld, const void* queries, int64_t Q, int k,
                         int64_t row_offset, const int64_t* exclude_indptr, const int32_t* exclude_rows,
                         float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);
/* k-way merge of `parts` sorted-or-not candidate lists laid out [parts][Q][k_in] (score fp32, id int64, id<0 = empty)
 * into the global top k_out under the same order.  The exchange step after the all-gather of per-GPU results. */
int b200rec_topk_merge(const float* scores, const int64_t* ids, int parts, int64_t Q, int k_in, int k_out,
                       float* out_scores, int64_t* out_ids, void* stream);

/* ---------------------------------------------------------------- embedding bags (K1)
 * Fused multi-field gather: out[b, col_off[f] : col_off[f]+width[f]] = table_f[idx_f[b], :width[f]] for F fields,
 * plus the numerical block copied to out[b, 0:num_cols].  Replaces nn.Embedding fwd + 2x torch.cat
 * (two_tower.py:113-126, :254-273).  tables/indices are device arrays of F device pointers. */
int b200rec_gather_concat(const float* numerical, int64_t num_cols, int64_t ld_num, const float* const* tables,
                          const int64_t* const* indices, const int32_t* widths, const int32_t* table_ld,
                          const int32_t* col_off, int F, int64_t B, float* out, int64_t ld_out, void* stream);
/* Sparse row-gradient of one table: given idx[B] and dY (the field's column block of d(out), row stride ld_dy),
 * emits the coalesced (unique_rows[U] ascending, grad_rows[U,width]) with duplicates summed and the padding row
 * (idx 0) dropped — the non-zero rows of ATen embedding_dense_backward (nn.Embedding(padding_idx=0), two_tower.py:46-50).
 * n_unique_out is a device int32.  B <= 16384 per call in this version. */
size_t b200rec_sparse_grad_workspace_bytes(int64_t B);
int b200rec_embedding_sparse_grad(const int64_t* idx, int64_t B, const float* dY, int64_t ld_dy, int width,
                                  int64_t* unique_rows, float* grad_rows, int32_t* n_unique_out, void* workspace,
                                  size_t workspace_bytes, void* stream);
/* Row-sparse Adam on the touched rows only (replaces dense torch.optim.Adam on the table, trainers/two_tower.py:60-64). */
int b200rec_sparse_adam(float* table, float* exp_avg, float* exp_avg_sq, int64_t ld, int width,
                        const int64_t* rows, const float* grad_rows, const int32_t* n_rows, int64_t max_rows,
                        float lr, float beta1, float beta2, float eps, float bias_c1, float bias_c2,
                        float grad_scale, void* stream);

/* ---------------------------------------------------------------- tower MLP pieces (K2)
 * act: 0 relu, 1 gelu(erf), 2 leaky_relu(0.1), 3 tanh, 4 sigmoid, 5 identity  (two_tower.py:77-86).
 * Column statistics of a = act(z) over B rows: sums[0:H] = sum a, sums[H:2H] = sum a^2 (fp32, pre-zeroed). */
int b200rec_bn_stats(const float* z, int64_t B, int64_t H, int64_t ld, int act, float* sums, void* stream);
/* y = ((act(z) - mean) * invstd * gamma + beta) * dropout_mask/(1-p); writes y fp32 (nullable) and the next layer's
 * bf16 operand (nullable).  mean/invstd are [H].  Dropout mask = Philox(seed, element index) < keep. */
int b200rec_bn_apply(const float* z, int64_t B, int64_t H, int64_t ld, int act, const float* mean,
                     const float* invstd, const float* gamma, const float* beta, float drop_p, uint64_t seed,
                     float* y, int64_t ld_y, void* y_bf16, int64_t kpad, int terms, int side, void* stream);
/* BatchNorm backward, pass 1: sums[0:H] = sum dy' , sums[H:2H] = sum dy'*xhat  with dy' = dy*mask/(1-p). */
int b200rec_bn_bwd_stats(const float* dy, int64_t ld_dy, const float* z, int64_t ld_z, int64_t B, int64_t H, int act,
                         const float* mean, const float* invstd, float drop_p, uint64_t seed, float* sums,
                         void* stream);
/* pass 2: dz = act'(z) * invstd*gamma/n_total * (n_total*dy' - sum1 - xhat*sum2)  (training) or dy'*gamma*invstd (eval,
 * n_total==0).  Also dbias[h] += sum_b dz (fp32 atomics, pre-zeroed). */
int b200rec_bn_bwd_apply(const float* dy, int64_t ld_dy, const float* z, int64_t ld_z, int64_t B, int64_t H, int act,
                         const float* mean, const float* invstd, const float* gamma, const float* sums,
                         float n_total, float drop_p, uint64_t seed, float* dz, int64_t ld_dz, float* dbias,
                         void* stream);
/* F.normalize backward: dO = (dE - E * rowsum(E*dE)) / norm. */
int b200rec_normalize_bwd(const float* dE, const float* E, const float* norms, int64_t B, int64_t D, float* dO,
                          void* stream);
/* column sums: out[h] += sum_b x[b,h] (pre-zeroed) — Linear bias gradients. */
int b200rec_colsum(const float* x, int64_t B, int64_t H, int64_t ld, float* out, void* stream);

/* ---------------------------------------------------------------- losses (K3)
 * In-batch softmax cross-entropy, logits never materialised (two_tower.py:467-479):
 *   loss = mean_i( logsumexp_j(U_i.I_j / T) - U_i.I_i / T ).
 * U,I given as bf16 operands (terms*kpad wide: left side for U, right side for I).  lse_out [B] is kept for backward.
 * Under data parallelism I may hold more rows (NI >= B, all-gathered); `diag_offset` is this rank's first row in I. */
size_t b200rec_inbatch_ce_workspace_bytes(int64_t B, int64_t NI);
int b200rec_inbatch_ce_fwd(const void* U, const void* I, int64_t ld, int64_t B, int64_t NI, int64_t diag_offset,
                           float inv_temperature, float* lse_out, float* loss_sum_out, void* workspace,
                           size_t workspace_bytes, void* stream);
/* Backward by tile-wise recomputation: dU[B,E] = g/(T*Btot) * (softmax(S) - onehot) . I ;
 * dI[NI,E] = g/(T*Btot) * (softmax(S) - onehot)^T . U.   I_t / U_t are the same matrices as MN-major-free transposed
 * bf16 operands ([E', NI] / [E', B]); see DESIGN.md. */
int b200rec_inbatch_ce_bwd(const void* U, const void* I, int64_t ld, int64_t B, int64_t NI, int64_t diag_offset,
                           float inv_temperature, const float* lse, float grad_scale, const float* U_f32,
                           const float* I_f32, int64_t E, float* dU, float* dI, void* workspace,
                           size_t workspace_bytes, void* stream);
/* Explicit-negative CE (two_tower.py:422-451): logits[b] = [u.p/T + ub + ib, u.n_{b,0..R-1}/T]; label 0; sum over b of
 * the per-row loss is ADDED to loss_sum_out.  If dU != NULL also writes gradients scaled by grad_scale/B
 * (dU, dP [B,E], dN [B*R,E], dbias += sum_b dlogit0). */
int b200rec_explicit_ce(const float* U, const float* P, const float* Nn, int64_t B, int64_t R, int64_t E,
                        float inv_temperature, float bias_sum, float* loss_sum_out, float grad_scale, float* dU,
                        float* dP, float* dN, float* dbias, void* stream);

/* ---------------------------------------------------------------- optimiser helpers (K5)
 * sum of squares of n floats added to out (pre-zeroed) — the pieces of clip_grad_norm_. */
int b200rec_sumsq(const float* x, int64_t n, float* out, void* stream);
/* Dense Adam with L2-coupled weight decay on a flat fp32 buffer; `clip_coef_dev` (device float, nullable) scales g. */
int b200rec_adam_dense(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float bias_c1, float bias_c2, const float* clip_coef_dev,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_H */

/*
 * okge_b200.h — C ABI of the B200-native OpenKGE hot path (libokge_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. Every pointer is a
 * DEVICE pointer unless stated otherwise; `stream` is a cudaStream_t passed as void*. All calls
 * are asynchronous on `stream`, never synchronise the host, never allocate or free caller memory
 * and return 0 on success (OKGE_OK) or a non-zero status; okge_last_error() gives the message.
 *
 * The reference (samuelbroscheit/open_knowledge_graph_embeddings) has no native layer: each entry
 * point replaces a group of PyTorch ATen calls on the reference's hot path. The reference
 * file:line each one replaces is cited at the declaration (paths relative to the reference root).
 *
 * Conventions
 *   B  rows (prefix queries) of a batch, po rows first then sp rows (openkge/dataset.py:885-932)
 *   N  candidate entities (= entities_size - 2 in 1-vs-all mode, openkge/dataset.py:872)
 *   D  real embedding width (ComplEx: first D/2 real, last D/2 imaginary, openkge/model.py:200-203)
 *   L  token slots per mention/relation row (max_lengths_tuple, openkge/model.py:576-595)
 *   ids are int32 (torch.IntTensor on the reference's wire), matrices fp32 row-major with an
 *   explicit leading dimension in ELEMENTS; the operands of the scoring contractions are fp16 copies
 *   (section 3). Row pitches must be multiples of 16 bytes and base pointers 16-byte aligned wherever
 *   a matrix feeds the tensor-core path (TMA constraint).
 */
#ifndef OKGE_B200_H_
#define OKGE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OKGE_OK 0
#define OKGE_ERR_INVALID 1     /* bad argument (shape, alignment, null pointer)            */
#define OKGE_ERR_CUDA 2        /* a CUDA runtime/driver call failed                         */
#define OKGE_ERR_UNSUPPORTED 3 /* device is not sm_100 or a required driver symbol missing */

#define OKGE_ABI_VERSION 2

typedef void* okge_stream_t; /* cudaStream_t */

/* Pooling modes of UnigramPoolingRelationEmbedder (openkge/model.py:768-774). */
#define OKGE_POOL_SUM 0
#define OKGE_POOL_MEAN 1
#define OKGE_POOL_MAX 2

/* Query folding kinds: the 1-vs-all ("prefix") score of each scorer written as ONE row vector q
 * so that score[b, n] = <q[b, :], E[n, :]>.
 *   COMPLEX_SP  a = subj, b = rel : q = [a1*b1 - a2*b2 ; a2*b1 + a1*b2]   (openkge/model.py:206-209)
 *   COMPLEX_PO  a = obj,  b = rel : q = [a1*b1 + a2*b2 ; a2*b1 - a1*b2]   (openkge/model.py:212-215)
 *   DISTMULT    a = subj|obj, b = rel : q = a*b                           (openkge/model.py:270-272) */
#define OKGE_FOLD_COMPLEX_SP 0
#define OKGE_FOLD_COMPLEX_PO 1
#define OKGE_FOLD_DISTMULT 2

/* ---- library ------------------------------------------------------------------------------ */

int okge_abi_version(void);
/* Message of the last failing call on this thread ("" if none). HOST pointer, owned by the library. */
const char* okge_last_error(void);
/* 0 if the current CUDA device can run the kernels (compute capability 10.x), else a status. */
int okge_device_check(void);

/* ---- (1) embedding gather and token pooling ------------------------------------------------ */

/* Lookup gather, out[i, :] = table[ids[i], :].
 * Replaces nn.Embedding.forward in LookupBaseRelationEmbedder._encode (openkge/model.py:457-458). */
int okge_gather_rows(const float* table, int64_t ld_table, const int32_t* ids, int64_t n, int64_t D,
                     float* out, int64_t ld_out, okge_stream_t stream);

/* grad_table[ids[i], :] += grad[i, :] (rows with ids[i] == skip_id are skipped; pass -1 for none).
 * Replaces embedding_dense_backward of the same lookup (autograd of openkge/model.py:458). */
int okge_scatter_add_rows(const float* grad, int64_t ld_grad, const int32_t* ids, int64_t n,
                          int64_t D, int32_t skip_id, float* grad_table, int64_t ld_table,
                          okge_stream_t stream);

/* Token gather + pooling without the [n, L, D] intermediate:
 *   row = ids ? ids[i] : id_start + i;  tok = id_rows[row, 0..L);  out[i,:] = pool_l tok_table[tok[l], :]
 * All L slots are gathered, PAD slots (token 0) included, exactly like the reference where the
 * PAD row of the token table is a trained-from-init non-zero vector (openkge/model.py:633-634);
 * MEAN divides by (#tok>0 + 1e-12) (openkge/model.py:770-772).
 * Replaces UnigramPoolingRelationEmbedder._map_to_tokens/_encode (openkge/model.py:762-774). */
int okge_gather_pool_fwd(const float* tok_table, int64_t ld_table, const int32_t* id_rows, int32_t L,
                         const int32_t* ids, int64_t id_start, int64_t n, int64_t D, int32_t mode,
                         float* out, int64_t ld_out, okge_stream_t stream);

/* Backward of okge_gather_pool_fwd: scatter-add of the pooled-row gradients into the token table
 * gradient (caller zeroes/owns grad_tok_table). Token 0 (padding_idx) never receives gradient
 * (openkge/model.py:597-608). `tok_table` is only read for OKGE_POOL_MAX (arg-max recomputation).
 * Replaces embedding_dense_backward + sum/div/max backward (autograd of openkge/model.py:767-774). */
int okge_gather_pool_bwd(const float* grad_out, int64_t ld_grad, const float* tok_table,
                         int64_t ld_table, const int32_t* id_rows, int32_t L, const int32_t* ids,
                         int64_t id_start, int64_t n, int64_t D, int32_t mode, float* grad_tok_table,
                         okge_stream_t stream);

/* The same backward with a COMPACT destination: token t is accumulated into row slot_map[t] of slot_grad (slots assigned by
 * okge_row_slots_build over the flattened token ids of the gathered rows). Used when far fewer rows than the table has
 * receive a gradient: the dense [V, D] gradient is never zero-filled, written or read; okge_adagrad_slot_table applies it. */
int okge_gather_pool_bwd_slots(const float* grad_out, int64_t ld_grad, const float* tok_table, int64_t ld_table,
                               const int32_t* id_rows, int32_t L, const int32_t* ids, int64_t id_start, int64_t n,
                               int64_t D, int32_t mode, const int32_t* slot_map, float* slot_grad, okge_stream_t stream);

/* Inverted dropout with a counter-based generator: out = x * keep / (1 - p), keep ~ Bernoulli(1-p)
 * drawn from Philox4x32-10(seed, element index + offset). Applying the same (seed, offset) to a
 * gradient is the backward pass. p == 0 is a copy. Streams cannot match torch's generator, so
 * parity runs use p = 0. Replaces F.dropout (openkge/model.py:461-470, 783-786). */
int okge_dropout(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset, float* out,
                 okge_stream_t stream);

/* The same with the stream position split in two: `offset` (a launch parameter) + (*step_dev << 44) read from device
 * memory, so that a launch replayed from a CUDA graph draws a fresh mask every step (the host program increments
 * *step_dev once per step; the backward launch of the same step reads the same value). */
int okge_dropout_step(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* step_dev,
                      float* out, okge_stream_t stream);

/* ---- (1b) batch normalisation of encoded rows --------------------------------------------------
 * torch.nn.BatchNorm1d over the rows of an [n, D] operand as the embedders apply it (openkge/model.py:463-465 Lookup,
 * :597-612 and :777-780 token models; momentum 0.1, eps 1e-5). The reference normalises every encode call of a batch
 * separately (openkge/trainer.py:69-87: candidates, po rel, po obj, sp subj, sp rel); here the row SEGMENTS of one call
 * are device data: seg (int32[2 * n_seg] on the device: one [begin, end) row range per segment, disjoint, in the order
 * of the reference's calls; rows outside every range are left untouched) or NULL for the single segment
 * [0, n_rows). n_rows is the host-side upper bound of the rows any segment covers (grid sizing only). Empty segments are
 * skipped; segments update the running statistics one after the other, in order (biased variance for the
 * normalisation, unbiased for running_var, *num_batches_tracked += number of non-empty segments). D % 4 == 0. */

/* Fused inverted dropout (the F.dropout that follows the normalisation in the embedders, openkge/model.py:783-786):
 * drop_p > 0 makes okge_bn_train_fwd write mask * y / (1 - p) and okge_bn_train_bwd read mask * dy / (1 - p), with the mask
 * okge_dropout / okge_dropout_step would draw for (drop_seed, drop_offset [, *drop_step_dev]) over the flattened [n, D]
 * output (ld_y must be D); the normalised-but-not-dropped tensor and the mask never reach memory. drop_p = 0: off. */

/* Bytes of scratch okge_bn_train_fwd / okge_bn_train_bwd need for this shape. */
int64_t okge_bn_workspace_bytes(int64_t n_rows, int D, int n_seg);

/* Training mode: y = (x - mean_seg) * invstd_seg * gamma + beta; save_mean / save_invstd [n_seg, D] keep the statistics
 * for the backward. gamma / beta may be NULL (1 / 0); running_mean / running_var / num_batches_tracked may be NULL. */
int okge_bn_train_fwd(const float* x, int64_t ld_x, const int32_t* seg, int n_seg, int64_t n_rows, int D,
                      const float* gamma, const float* beta, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, float momentum, float eps, float* y, int64_t ld_y,
                      float* save_mean, float* save_invstd, float drop_p, uint64_t drop_seed, uint64_t drop_offset,
                      const uint64_t* drop_step_dev, void* workspace, okge_stream_t stream);

/* dx = gamma * invstd * (dy - mean(dy) - xhat * mean(dy * xhat)) per segment; dgamma[D] = sum(dy * xhat),
 * dbeta[D] = sum(dy) over all segments (written, not accumulated). dx / dgamma / dbeta may be NULL. */
int okge_bn_train_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const int32_t* seg, int n_seg,
                      int64_t n_rows, int D, const float* gamma, const float* save_mean, const float* save_invstd,
                      float* dx, int64_t ld_dx, float* dgamma, float* dbeta, float drop_p, uint64_t drop_seed,
                      uint64_t drop_offset, const uint64_t* drop_step_dev, void* workspace, okge_stream_t stream);

/* Eval mode: y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta (the all-rows eval cache of the token
 * models, openkge/model.py:670-712). */
int okge_bn_eval_fwd(const float* x, int64_t ld_x, int64_t n_rows, int D, const float* gamma, const float* beta,
                     const float* running_mean, const float* running_var, float eps, float* y, int64_t ld_y,
                     okge_stream_t stream);

/* The three phases of the batch norm as separate calls, single segment, for callers whose rows are partitioned over
 * several GPUs (synchronised batch norm: the column sums are all-reduced between phase 1 and phase 2).
 *   okge_bn_col_sums      sums[0:D] = sum_r a[r, :], sums[D:2D] = sum_r a[r, :]^2 (x == NULL: forward statistics of a = x)
 *                         or sum_r a[r, :] * xhat[r, :] with xhat = (x - mean) * invstd (backward: a = dy); fp64
 *   okge_bn_normalize     y = (x - mean) * invstd * gamma + beta
 *   okge_bn_normalize_bwd dx = gamma * invstd * (dy - coef[0:D] - xhat * coef[D:2D])   (coef = global sums / global n)
 * workspace: okge_bn_workspace_bytes(n_rows, D, 1) bytes. */
int okge_bn_col_sums(const float* a, int64_t ld_a, const float* x, int64_t ld_x, const float* mean, const float* invstd,
                     int64_t n_rows, int D, double* sums, void* workspace, okge_stream_t stream);
int okge_bn_normalize(const float* x, int64_t ld_x, int64_t n_rows, int D, const float* mean, const float* invstd,
                      const float* gamma, const float* beta, float* y, int64_t ld_y, okge_stream_t stream);
int okge_bn_normalize_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, int64_t n_rows, int D,
                          const float* mean, const float* invstd, const float* coef, const float* gamma, float* dx,
                          int64_t ld_dx, okge_stream_t stream);

/* ---- (1c) LSTM token encoder: point-wise cell ---------------------------------------------------
 * LSTMRelationEmbedder (openkge/model.py:912-998): single-layer torch.nn.LSTM over the token embeddings of a mention
 * (gate order i, f, g, o; h0 = c0 = 0), output = hidden state at the last real token, last_state[row] =
 * (#tokens > 0) - 1. The gate pre-activations come from okge_gemm_tf32_nt (gx = x_t W_ih^T, gh = h_{t-1} W_hh^T); these
 * two calls are the element-wise rest of one time step. D % 4 == 0, all operands 16-byte aligned. */

/* act[n, 4D] = (sigmoid(i), sigmoid(f), tanh(g), sigmoid(o)) of gx + gh + b_ih + b_hh (gh NULL at t = 0),
 * c = f * c_prev + i * g (c_prev NULL = 0), h = o * tanh(c); rows with last_state[row] == t also write h to out.
 * act may be NULL (inference). */
int okge_lstm_cell_fwd(const float* gx, int64_t ld_gx, const float* gh, int64_t ld_gh, const float* b_ih,
                       const float* b_hh, const float* c_prev, int64_t n, int64_t D, int32_t t,
                       const int32_t* last_state, float* act, float* c, float* h, float* out, okge_stream_t stream);

/* Back-propagation of one time step: dh (NULL = 0) is the gradient arriving from step t + 1 through W_hh, grad_out the
 * gradient of the encoder output (added for rows with last_state[row] == t), dc [n, D] the running cell gradient
 * (in: d loss / d c_t from later steps, out: d loss / d c_{t-1}). dgates[n, 4D] = gradient of the gate pre-activations. */
int okge_lstm_cell_bwd(const float* act, const float* c_prev, const float* c, const float* grad_out,
                       const int32_t* last_state, int32_t t, const float* dh, float* dc, int64_t n, int64_t D,
                       float* dgates, okge_stream_t stream);

/* ---- (2) query folding ----------------------------------------------------------------------- */

/* q[b, :] = fold(kind, a[b, :], b[b, :]), see OKGE_FOLD_*. D must be even for ComplEx. */
int okge_fold_query(int32_t kind, const float* a, const float* b, int64_t Bq, int64_t D, float* q,
                    okge_stream_t stream);
/* (grad_a, grad_b) from grad_q for the same fold. */
int okge_fold_query_bwd(int32_t kind, const float* a, const float* b, const float* grad_q,
                        int64_t Bq, int64_t D, float* grad_a, float* grad_b, okge_stream_t stream);

/* The two ComplEx folds with the kind of every row read from device memory (kinds[r] = OKGE_FOLD_COMPLEX_SP | _PO): the po
 * and sp rows of a batch in one launch whose shape does not depend on the po / sp split (CUDA-graph replay). */
int okge_fold_query_rows(const int32_t* kinds, const float* a, const float* b, int64_t Bq, int64_t D, float* q,
                         okge_stream_t stream);
int okge_fold_query_rows_bwd(const int32_t* kinds, const float* a, const float* b, const float* grad_q, int64_t Bq,
                             int64_t D, float* grad_a, float* grad_b, okge_stream_t stream);

/* Per-batch bookkeeping of a graph-replayed training step, from device data (no host involvement): the po rows of a batch
 * come first (openkge/dataset.py:794-811), *n_po_dev of them. kinds[r] = r < n_po ? kind_po : kind_sp (nullable; the input
 * of okge_fold_query_rows); segments (nullable) = the [begin, end) row ranges of the batch-norm statistics in the
 * reference's call order -- Lookup models (token_model = 0): {0, n_po, n_po, rows}; token models, which encode the
 * candidates and the query rows in one call: {0, count, n_cols, n_cols + n_po, n_cols + n_po, n_cols + rows, 0, n_po, n_po,
 * rows} with count = *count_dev (the real length of a padded batch-shared list; NULL = n_cols). *step_counter += 1
 * (nullable): the dropout stream position of okge_dropout_step. */
int okge_batch_layout(const int32_t* n_po_dev, const int32_t* count_dev, int64_t rows, int64_t n_cols, int32_t kind_po,
                      int32_t kind_sp, int32_t* kinds, int32_t* segments, int32_t token_model, uint64_t* step_counter,
                      okge_stream_t stream);

/* ---- (3) 1-vs-all scoring on the tensor cores (tcgen05, FP16 inputs, FP32 accumulate) ---------- */

/* FP16 operands. The reference contracts in fp32 (torch.mm). The tensor cores take fp16 here: the same 10-bit mantissa
 * as TF32, rounded to nearest, at twice the TF32 rate and half the bytes. Every operand is an IEEE binary16 matrix
 * x16 = fp16(x * scale) with ONE power-of-two scale per operand and its inverse in device memory (`*_inv`, nullable = 1);
 * the epilogues multiply the fp32 accumulator by the inverse scales, which is exact. okge_f16_quantize produces such
 * operands (dynamic scale: largest element in [128, 256); conversions saturate). Resulting score error on B200:
 * about 1e-4 * ||q|| * ||e|| (tests use 1e-3). Split precision: an operand may carry a second plane
 * lo = fp16(x * scale - hi) (`*_lo`, nullable, same leading dimension, at a higher address than hi); with both lo planes
 * the contraction computes q_hi e_hi + q_hi e_lo + q_lo e_hi in ONE accumulator (three passes over K), which reproduces
 * the fp32 product to ~1e-6 norm-wise: the evaluation path ranks with it. */
typedef uint16_t okge_half_t; /* IEEE binary16 bit pattern */

/* Number of per-block partial maxima okge_f16_absmax writes (floats). */
#define OKGE_F16_ABSMAX_PARTS 256
/* partials[0 .. OKGE_F16_ABSMAX_PARTS) = per-block max |x| over the [rows, cols] fp32 matrix x (row pitch ld). */
int okge_f16_absmax(const float* x, int64_t ld, int64_t rows, int64_t cols, float* partials, okge_stream_t stream);
/* hi[r, c] = fp16(x[r, c] * scale), lo[r, c] = fp16(x[r, c] * scale - hi[r, c]) if lo != NULL (row pitch ld16 halves),
 * inv_scale[0] = 1 / scale (device, nullable). scale = 2^(8 - e) for max |x| in [2^(e-1), 2^e) taken from `partials`
 * (okge_f16_absmax of the same matrix), or `fixed_scale` when partials == NULL (a table whose fp16 copy is maintained
 * incrementally, see okge_gemm_adagrad, keeps its scale). Replaces nothing in the reference: it is the operand
 * preparation of the tensor-core form of openkge/model.py:206-215, 270-272. */
int okge_f16_quantize(const float* x, int64_t ld, int64_t rows, int64_t cols, const float* partials, float fixed_scale,
                      okge_half_t* hi, okge_half_t* lo, int64_t ld16, float* inv_scale, okge_stream_t stream);

/* Operand layouts of the contractions. ROW_MAJOR: [rows, K] with a leading dimension. K_PANELS: the same logical
 * matrix stored as [ceil(K/P)][rows][P] elements, P = 64 (fp16) or 32 (fp32) = one 128-byte line (panel p holds columns
 * P p .. P p + P - 1 of every row, the tail of the last panel is zero): every 128-row TMA box is then ONE contiguous
 * 16 KB block of memory instead of 128 strips that are a whole row pitch (2 MB at K = 10^6) apart. The loss gradient
 * dS is produced directly in this layout. */
#define OKGE_ROW_MAJOR 0
#define OKGE_K_PANELS 1
/* MN-major forms (tcgen05 reads them through 128B-swizzled shared memory, no transpose pass):
 * COL_MAJOR: the logical [rows, K] operand lives in memory as [K][ld] with the rows contiguous, i.e. a row-major
 * matrix X[K, rows] used as X^T (ld >= rows, multiple of 16 bytes) -- how E[N, D] enters dQ = dS E and Q[B, D] enters
 * dE = dS^T Q. MN_PANELS: [ceil(rows/P)][K][P] = the K_PANELS storage of the transposed matrix -- how the dS panels
 * enter dE = dS^T Q, so dS^T is never written. */
#define OKGE_COL_MAJOR 2
#define OKGE_MN_PANELS 3

/* C[M, N] = alpha * A[M, K] * B[N, K]^T on fp32 operands (TF32 inputs: tcgen05 truncates the low 13 mantissa bits);
 * each operand in any of the four layouts above. alpha_dev (nullable, device scalar) multiplies alpha. splits > 1
 * splits K over CTAs: partials go to split_ws[splits, M, N] (caller provided) and are summed deterministically into C.
 * The generic contraction for fp32 operands outside the 1-vs-all path (LSTM gate products,
 * openkge/model.py:963-987). */
int okge_gemm_tf32_nt(const float* A, int64_t lda, int32_t a_layout, const float* B, int64_t ldb,
                      int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* alpha_dev,
                      float* C, int64_t ldc, int32_t splits, float* split_ws, okge_stream_t stream);

/* Dropout of an fp16 operand without leaving fp16: dst[r, c] = keep ? src[r, c] : 0 for the inverted-dropout mask
 * okge_dropout[_step] draws for (p, seed, offset [, *step_dev]) over the flattened [rows, cols] matrix (cols % 4 == 0);
 * the 1 / (1 - p) of the kept elements goes into the operand's scale: *dst_inv_scale = *src_inv_scale / (1 - p)
 * (src_inv_scale NULL = 1). With src = the fp16 shadow of the entity table this is the candidate operand of a 1-vs-all
 * training step under input dropout (openkge/model.py:461-470 via _get_all, :512-514) in ONE 4 B/element pass instead of
 * dropout (8 B/element) + absmax (4) + quantize (6) of the fp32 rows. */
int okge_f16_mask_dropout(const okge_half_t* src, int64_t ld_src, int64_t rows, int64_t cols, float p, uint64_t seed,
                          uint64_t offset, const uint64_t* step_dev, const float* src_inv_scale, okge_half_t* dst,
                          int64_t ld_dst, float* dst_inv_scale, okge_stream_t stream);

/* The same on fp16 operands: C = alpha * scale0 * scale1 * scale2 * A B^T (device scalars, nullable: the operands'
 * inverse scales and a gradient scale, applied without a host sync). This is the one tensor-core kernel; every score /
 * gradient contraction below is an instance:
 *   scores = Q E^T          (openkge/model.py:206-215, 270-272: the 4-mm ComplEx form and the DistMult mm)
 *   dQ = dS E, dE = dS^T Q  (autograd of the same mm calls) */
int okge_gemm_f16_nt(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                     int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                     const float* scale1, const float* scale2, float* C, int64_t ldc, int32_t splits, float* split_ws,
                     okge_stream_t stream);

/* scores[B, N] = q e^T materialised (debug / parity / reference-compatible all_outputs; split precision with both lo
 * planes). Replaces ComplexRelationScorer._score / DistmultRelationScorer._score with prefix=True
 * (openkge/model.py:181-229, 248-278) after okge_fold_query + okge_f16_quantize. */
int okge_score_store(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                     const okge_half_t* e_lo, int64_t lde, int64_t B, int64_t N, int64_t D, const float* q_inv,
                     const float* e_inv, float* scores, int64_t lds, okge_stream_t stream);

/* Fused 1-vs-all scoring + BCE-with-logits(sum) loss; the score matrix never reaches memory.
 * Labels are sparse: row b has positives pos_idx[pos_ptr[b] .. pos_ptr[b+1]) (candidate-local column
 * indices, unique per row; entries outside [0, N) are ignored, which is how a column-sharded caller marks the
 * positives that live on other shards without changing the row pointer). Every label is y_base except positives which are y_pos
 * (bce_label_smoothing eps: y_base = (1-eps)/N, y_pos = (1+1/N)(1-eps); openkge/trainer.py:103-105).
 *   loss_sum[0]  = sum_{b,n} softplus(s) - s*y                      (double, overwritten)
 *   dS [B, N]    = fp16(ds_scale * (sigmoid(s) - y))  if dS != NULL, OKGE_K_PANELS layout: ceil(N/64)*B*64 halves,
 *                  128-byte aligned; the A operand of the dQ / dE contractions (ds_scale: power of two, e.g. 4096).
 * n_cols_dev (nullable, device int32): when the caller pads the candidate list to a fixed capacity N (batch-shared
 * candidates of openkge/dataset.py:813-860 replayed from a CUDA graph), *n_cols_dev is the number of real candidates:
 * columns at or beyond it contribute no loss and get a zero gradient.
 * Replaces torch.cat + BCEWithLogitsLoss(reduction='sum') (openkge/trainer.py:91-106) and the
 * sigmoid/sub of its backward. */
int okge_score_bce(const okge_half_t* q, int64_t ldq, const okge_half_t* e, int64_t lde, int64_t B, int64_t N,
                   int64_t D, const float* q_inv, const float* e_inv, const int32_t* pos_ptr, const int32_t* pos_idx,
                   float y_base, float y_pos, const int32_t* n_cols_dev, double* loss_sum, okge_half_t* dS,
                   float ds_scale, okge_stream_t stream);

/* Evaluation step in ONE pass over the candidates (openkge/trainer.py:259-272 computes the loss and then
 * compute_metrics on the same scores): okge_score_bce's loss sum plus, for up to 4 ranked answers per query row, the
 * counts of okge_score_rank. thresh4 [B, 4] fp32 (16-byte aligned; +inf in unused slots), greater4 / equal4 [B, 4] int32
 * are ADDED to; the kernel looks at the first n_slots (1, 2 or 4) slots of every row: each slot costs four instructions
 * per score, so the caller picks the smallest count that fits its batch. Prefix rows with more ranked answers than
 * slots: q may carry B_extra further rows behind the B prefix rows
 * (copies of the query vectors of those prefixes) whose slots hold the further thresholds; they are scored and
 * counted like the others but contribute no loss (q, thresh4, greater4, equal4 then have B + B_extra rows; labels B
 * rows). Scores are formed exactly as in okge_score_rank / okge_score_store, so the counts are bit-identical to the
 * two-pass path. */
int okge_score_bce_rank(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                        const okge_half_t* e_lo, int64_t lde, int64_t B, int64_t B_extra, int64_t N, int64_t D,
                        const float* q_inv, const float* e_inv, const int32_t* pos_ptr, const int32_t* pos_idx,
                        float y_base, float y_pos, const float* thresh4, int32_t* greater4, int32_t* equal4,
                        int32_t n_slots, double* loss_sum, okge_stream_t stream);

/* Fused scoring + row-wise log-sum-exp for the softmax/KL loss (openkge/trainer.py:99-100, 106):
 *   row_lse[b]      = log sum_n exp(s[b, n])
 *   pos_score[p]    = s[b, pos_idx[p]] for every CSR entry p
 * so that KLDivLoss(sum)(log_softmax(s), y) = sum_b npos_b*row_lse[b] - sum_p pos_score[p].
 * part_ws: caller workspace of okge_score_lse_ws_floats(B, N) floats. */
int64_t okge_score_lse_ws_floats(int64_t B, int64_t N);
int okge_score_lse(const okge_half_t* q, int64_t ldq, const okge_half_t* e, int64_t lde, int64_t B, int64_t N,
                   int64_t D, const float* q_inv, const float* e_inv, const int32_t* pos_ptr, const int32_t* pos_idx,
                   float* row_lse, float* pos_score, float* part_ws, okge_stream_t stream);

/* Gradient of the softmax/KL loss w.r.t. the scores, recomputed tile by tile:
 *   dS[b, n] = fp16(ds_scale * (row_weight[b] * exp(s[b, n] - row_lse[b]) - y[b, n]))   (row_weight = sum of y in row b)
 * written as fp16 K-panels like okge_score_bce. */
int okge_score_softmax_grad(const okge_half_t* q, int64_t ldq, const okge_half_t* e, int64_t lde, int64_t B,
                            int64_t N, int64_t D, const float* q_inv, const float* e_inv, const int32_t* pos_ptr,
                            const int32_t* pos_idx, const float* row_lse, const float* row_weight, okge_half_t* dS,
                            float ds_scale, okge_stream_t stream);

/* ---- (4) filtered ranking ---------------------------------------------------------------------- */

/* Filtered rank counts over a MATERIALISED score matrix, bit-exact restatement of
 * OneToNMentionRelationDataset.compute_metrics (openkge/dataset.py:423-445) for Q ranked answers:
 *   answer j belongs to prefix row ans_row[j]; its alternative mentions are alt_idx[alt_ptr[j]..alt_ptr[j+1])
 *   true[j]    = max_a scores[row, a]                       (unmasked, :436-438)
 *   masked     = scores[row, :] with filt_idx[filt_ptr[row]..filt_ptr[row+1]) set to -1e8   (:440)
 *   greater[j] = #{n : true[j] <  masked[n]}                (:441-443)
 *   equal[j]   = #{n : true[j] == masked[n]}                (:444)
 * rank = greater + equal/2 is left to the caller (:445). filt_idx must be unique per row. */
int okge_rank_count(const float* scores, int64_t lds, int64_t B, int64_t N, const int32_t* ans_row,
                    const int32_t* alt_ptr, const int32_t* alt_idx, int64_t Q, const int32_t* filt_ptr,
                    const int32_t* filt_idx, float* true_score, int32_t* greater, int32_t* equal,
                    okge_stream_t stream);

/* Fused scoring + counting: row j of q is the query of ranked answer j and thresh[j] its true score;
 *   greater[j] += #{n in [0,N) : thresh[j] <  s[j, n]},  equal[j] += #{n : thresh[j] == s[j, n]}
 * over ALL candidate columns of this (shard of the) entity table, unmasked; counters are int32 and
 * accumulated with integer atomics, so sharded partial counts add up exactly. The caller zeroes them. */
int okge_score_rank(const okge_half_t* q, const okge_half_t* q_lo, int64_t ldq, const okge_half_t* e,
                    const okge_half_t* e_lo, int64_t lde, int64_t Q, int64_t N, int64_t D, const float* q_inv,
                    const float* e_inv, const float* thresh, int32_t* greater, int32_t* equal, okge_stream_t stream);

/* true_score[j] = max(true_score[j], max_a sel_scores[ans_row[j], alt_pos[a]]) over a in
 * alt_ptr[j]..alt_ptr[j+1) with alt_pos[a] >= 0 (negative = not on this shard). */
int okge_rank_true_score(const float* sel_scores, int64_t lds, const int32_t* ans_row,
                         const int32_t* alt_ptr, const int32_t* alt_pos, int64_t Q, float* true_score,
                         okge_stream_t stream);

/* Filter correction after okge_score_rank: for every filtered candidate f of the answer's prefix row
 * (filt_pos[...] >= 0 is its column in sel_scores, negative = not on this shard):
 *   greater[j] -= [thresh[j] < s_f];  equal[j] -= [thresh[j] == s_f]
 * and, when add_mask_terms != 0, adds n_f * [thresh[j] < -1e8] / n_f * [thresh[j] == -1e8] — the
 * contribution of the -1e8 fill value itself (openkge/dataset.py:440) — exactly once per answer. */
int okge_rank_filter_correct(const float* sel_scores, int64_t lds, const int32_t* ans_row,
                             int64_t Q, const int32_t* filt_ptr, const int32_t* filt_pos,
                             const float* thresh, int32_t add_mask_terms, int32_t* greater,
                             int32_t* equal, okge_stream_t stream);

/* ---- (5) optimizer updates ---------------------------------------------------------------------- */

/* torch.optim.Adagrad step as the reference effectively runs it (utils/optim.py:139-160, 194-201;
 * eps inherited from the bootstrap Adam): g' = g + wd*p; G += g'^2; p -= clr * g' / (sqrt(G) + eps),
 * clr = lr / (1 + (step-1)*lr_decay) computed by the caller. Dense over n contiguous elements. */
int okge_adagrad_dense(float* param, const float* grad, float* state_sum, int64_t n, float clr,
                       float eps, float weight_decay, okge_stream_t stream);

/* Row-wise (sparse) Adagrad: the same update applied to rows row_ids[i] only, gradient rows given densely as
 * grad_rows[i, :] -- what torch.optim.Adagrad does with the sparse gradient of nn.Embedding(sparse=True)
 * (openkge/model.py:390-391). Exact w.r.t. the dense step iff wd == 0 (torch refuses sparse gradients with weight
 * decay). slot_map == NULL: row_ids must be unique. slot_map != NULL (okge_row_slots_build over the same ids, gradient
 * rows summed per slot by okge_row_slots_accumulate): ids may repeat, position i is applied iff slot_map[row_ids[i]] == i;
 * negative ids are skipped. */
int okge_adagrad_rows(float* param, float* state_sum, int64_t ld, const float* grad_rows,
                      int64_t ld_grad, const int32_t* row_ids, const int32_t* slot_map, int64_t n_rows, int64_t D,
                      float clr, float eps, float weight_decay, okge_stream_t stream);

/* torch.optim.Adam step (no amsgrad): m = b1*m + (1-b1)g'; v = b2*v + (1-b2)g'^2;
 * p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps); bias corrections computed by the caller. */
int okge_adam_dense(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    float bias_correction1, float bias_correction2, okge_stream_t stream);

/* Fused gradient contraction + Adagrad step: the gradient tile never reaches memory.
 *   g[m, n]   = alpha * scale0 * scale1 * scale2 * (A B^T)[m, n]  +  (extra_map && extra_map[m] >= 0 ? extra[extra_map[m], n] : 0)
 *   param[m, n], state_sum[m, n] <- Adagrad(param, g, state_sum)   exactly like okge_adagrad_dense
 *   shadow[m, n] = fp16(param[m, n] / *shadow_inv_scale)           if shadow != NULL (the scale okge_f16_quantize published)
 * With A = dS^T (OKGE_MN_PANELS view of the loss gradient) and B = Q (OKGE_COL_MAJOR), both fp16, this is dE = dS^T Q of
 * the 1-vs-all backward (autograd of openkge/model.py:206-215, 270-272) followed by torch.optim.Adagrad.step on the
 * candidate rows of the entity table (utils/optim.py:194-201) in ONE pass: the table and its accumulator move through
 * HBM once each way (16 B/element + 2 B for the fp16 copy that the NEXT step's scoring pass reads + the dS read) instead
 * of dE being written, re-read and the table streamed again (28 B/element). param / state_sum: [M, N] row-major fp32 with
 * row pitch ld (multiple of 4, 16-byte aligned), updated in place; shadow: [M, N] fp16 with row pitch ld_shadow
 * (multiple of 8, N % 8 == 0). extra rows are the lookup gradients of the batch's own entities, see okge_row_slots_*. */
int okge_gemm_adagrad(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                      int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                      const float* scale1, const float* scale2, const int32_t* extra_map, const float* extra,
                      int64_t ld_extra, float* param, float* state_sum, int64_t ld, okge_half_t* shadow,
                      int64_t ld_shadow, const float* shadow_inv_scale, float clr, float eps, float weight_decay,
                      okge_stream_t stream);

/* The same when the contraction is the gradient of a DROPPED-OUT candidate operand (the reference drops the whole
 * candidate matrix of a 1-vs-all step: _get_all -> encode_obj -> F.dropout, openkge/model.py:461-470, 512-514): the mask
 * okge_dropout[_step] draws for (drop_p, drop_seed, drop_offset [, *drop_step_dev]) over the flattened [M, N] matrix is
 * applied to the gradient tile (d raw = mask / (1 - p) * d dropped) before the extra rows are added and the update runs.
 * drop_p == 0: identical to okge_gemm_adagrad. N % 4 == 0. */
int okge_gemm_adagrad_dropout(const okge_half_t* A, int64_t lda, int32_t a_layout, const okge_half_t* B, int64_t ldb,
                              int32_t b_layout, int64_t M, int64_t N, int64_t K, float alpha, const float* scale0,
                              const float* scale1, const float* scale2, const int32_t* extra_map, const float* extra,
                              int64_t ld_extra, float* param, float* state_sum, int64_t ld, okge_half_t* shadow,
                              int64_t ld_shadow, const float* shadow_inv_scale, float clr, float eps, float weight_decay,
                              float drop_p, uint64_t drop_seed, uint64_t drop_offset, const uint64_t* drop_step_dev,
                              okge_stream_t stream);

/* Slots for the extra gradient rows of okge_gemm_adagrad. slot_map is a persistent [table rows] int32 buffer holding -1.
 *   build:       slot_map[ids[i]] = max_i i                    (one slot per distinct id, ids == skip_id ignored)
 *   accumulate:  extra[slot_map[ids[i]], :] += grad[i, :]      (extra: caller-zeroed [n, D])
 *   clear:       slot_map[ids[i]] = -1                         (restores the buffer after the step)
 * Together they are embedding_dense_backward of the batch lookups (autograd of openkge/model.py:458) in sparse form. */
int okge_row_slots_build(const int32_t* ids, int64_t n, int32_t skip_id, int32_t* slot_map, okge_stream_t stream);
int okge_row_slots_accumulate(const float* grad, int64_t ld_grad, const int32_t* ids, int64_t n, int64_t D,
                              int32_t skip_id, const int32_t* slot_map, float* extra, int64_t ld_extra,
                              okge_stream_t stream);
int okge_row_slots_clear(const int32_t* ids, int64_t n, int32_t skip_id, int32_t* slot_map, okge_stream_t stream);
/* The same Adagrad step for the first n_rows rows of a table, whose gradient is only their slot row (or zero): the PAD /
 * UNK rows of an entity table, which are not candidates and therefore not covered by okge_gemm_adagrad. */
int okge_adagrad_slot_rows(float* param, float* state_sum, int64_t ld, int64_t n_rows, int64_t D,
                           const int32_t* slot_map, const float* extra, int64_t ld_extra, float clr, float eps,
                           float weight_decay, okge_stream_t stream);

/* torch.optim.Adagrad's dense step over a whole [n_rows, D] table whose gradient is a compact slot table:
 * g[r] = slot_map[r] >= 0 ? slot_grad[slot_map[r], :] : 0 (rows without a slot still take the weight-decay step, like the
 * dense reference update). 16 B/element instead of 20 + the zero-fill and scatter of a dense gradient. D % 4 == 0. */
int okge_adagrad_slot_table(float* param, float* state_sum, int64_t n_rows, int64_t D, const int32_t* slot_map,
                            const float* slot_grad, float clr, float eps, float weight_decay, okge_stream_t stream);

/* Row-wise Adam over the listed rows (same list conventions as okge_adagrad_rows): the lazy / sparse form of the step
 * (moments of unlisted rows do not decay); torch.optim.Adam itself rejects sparse gradients, this is SparseAdam-like. */
int okge_adam_rows(float* param, float* exp_avg, float* exp_avg_sq, int64_t ld,
                   const float* grad_rows, int64_t ld_grad, const int32_t* row_ids, const int32_t* slot_map, int64_t n_rows,
                   int64_t D, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float bias_correction1, float bias_correction2, okge_stream_t stream);

/* ---- batch-shared collate of a training batch on the device ----------------------------------------------------------
 * Replaces the host-side collate of OneToNMentionRelationDataset with use_batch_shared_entities=True for TRAINING batches
 * (openkge/dataset.py:813-868: candidate list = the answers that occur in the batch, topped up to min_size_batch_labels
 * with sampled negatives; :885-932: po rows first, then sp rows, labels as columns of that list). Input: `rows` = n_rows
 * indices into the prefix index (device int64), the index itself in HBM (lab_ptr int64 [P + 1], lab_idx int32 = entity
 * id - id_offset with every answer list ascending, prefix int32 [P, 2], slot int32 [P]: 0 = po row, else sp row).
 * Output (all device): ent / rel / is_po [n_rows] int32 in the new row order; CSR labels ptr [n_rows + 1], idx [cap_nnz]
 * (columns of the candidate list, ascending within a row, -1 behind the last label); cand [cap_cols] entity ids (+
 * id_offset; positives ascending by id, then the negatives; entries behind `count` repeat id_offset); scalars (int64
 * [OKGE_COLLATE_SCALARS], see below; OVERFLOW, NNZ_TOTAL and CALLS accumulate over calls and must start at 0);
 * count_out (int32, optional) and inv_norm = 1 / (n_rows * count) (optional), the reference's 1 / normalizer_loss.
 * Capacities replace data-dependent shapes (nothing synchronises with the host, the call can be captured in a CUDA
 * graph): labels beyond cap_nnz and candidates beyond cap_cols are dropped and the batch is counted in OVERFLOW.
 * Negatives: n_draw Philox draws keyed by (seed, CALLS); the first min_size - n_unique of them that are neither positives
 * nor repeats are appended (n_draw ~ 2 * min_size + 64 makes a short list practically impossible).
 * Workspace (device): bitmap uint32 [ceil(n_entities / 32)], word_prefix int32 [same], tile_sum int32 [ceil(words / 1024)],
 * first_draw int32 [n_entities] filled with INT32_MAX before the first call (the call restores that), e_flat int32
 * [cap_nnz + n_draw], row_start int64 [n_rows]. */
#define OKGE_COLLATE_B_PO 0       /* po rows of the batch */
#define OKGE_COLLATE_COUNT 1      /* length of the candidate list */
#define OKGE_COLLATE_NNZ 2        /* positive labels kept (= ptr[n_rows]) */
#define OKGE_COLLATE_N_UNIQUE 3   /* distinct answers of the batch (before the capacity) */
#define OKGE_COLLATE_OVERFLOW 4   /* += 1 for a batch that lost labels / candidates to a capacity or got a short list */
#define OKGE_COLLATE_NNZ_TOTAL 5  /* += NNZ */
#define OKGE_COLLATE_CALLS 6      /* += 1; the Philox key of the negatives */
#define OKGE_COLLATE_LABELS_ALL 7 /* positive labels of the batch before the capacity */
#define OKGE_COLLATE_SCALARS 8
int okge_collate_shared(const int64_t* rows, int64_t n_rows, const int64_t* lab_ptr, const int32_t* lab_idx,
                        const int32_t* prefix, const int32_t* slot, int64_t n_entities, int32_t id_offset,
                        int64_t min_size, int64_t cap_nnz, int64_t cap_cols, int64_t n_draw, uint64_t seed,
                        uint32_t* bitmap, int32_t* word_prefix, int32_t* tile_sum, int32_t* first_draw, int32_t* e_flat,
                        int64_t* row_start, int32_t* ent, int32_t* rel, int32_t* is_po, int32_t* ptr, int32_t* idx,
                        int32_t* cand, int64_t* scalars, int32_t* count_out, float* inv_norm, okge_stream_t stream);

/* ---- 1-vs-all collate of training batches on the HOST (no GPU involved) ---------------------------------------------
 * Replaces the per-batch Python of OneToNMentionRelationDataset_collate_func for training batches in 1-vs-all mode
 * (openkge/dataset.py:794-811 groups the rows by slot, po rows first; :885-932 writes the prefix id columns; :873, 921
 * the label matrix -- here as CSR). k batches of B prefix rows each (rows [k * B], indices into the prefix table) are
 * written into ONE caller-allocated (pinned) int32 buffer; batch b owns [starts[b], starts[b + 1]) in the layout
 *   [ entity ids (B, padded to a multiple of 4) | relation ids (same) | row pointer (B + 1 entries), n_po, padding |
 *     label columns (count, padded to a multiple of 4) ]
 * i.e. section offsets 0, o_rel, o_ptr, o_idx with o_idx = o_ptr + pad4(B + 2): exactly the device-side staging buffer of
 * the graphed training step, which therefore receives a batch with a single H2D copy. Rows are stably partitioned (po
 * rows keep their order in front, sp rows behind). Plain C loops (a few microseconds per batch); the binding releases
 * the interpreter lock, so a loader thread runs next to the thread that launches the GPU work.
 *   okge_host_collate_plan: counts[b] = labels of batch b, starts[0 .. k] (int64) = buffer offsets; returns the total size
 *                           in starts[k].
 *   okge_host_collate_fill: fills `packed` (int32 [starts[k]]) and n_po[b].
 * row_is_sp uint8 [P], row_ent / row_rel int32 [P] (entity / relation id of every prefix row), lab_ptr int64 [P + 1],
 * lab_idx int32 (ascending candidate-local columns per row). */
int okge_host_collate_plan(const int64_t* rows, int64_t k, int64_t B, const int64_t* lab_ptr, int64_t n_prefix_rows,
                           int64_t* counts, int64_t* starts);
int okge_host_collate_fill(const int64_t* rows, int64_t k, int64_t B, const uint8_t* row_is_sp, const int32_t* row_ent,
                           const int32_t* row_rel, const int64_t* lab_ptr, const int32_t* lab_idx, const int64_t* starts,
                           int32_t* packed, int32_t* n_po);

#ifdef __cplusplus
}
#endif

#endif /* OKGE_B200_H_ */

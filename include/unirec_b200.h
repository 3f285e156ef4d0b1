/* unirec_b200.h -- C ABI of the B200-native hot path for microsoft/UniRec's sequential-recommender training loop.
 *
 * The reference has no FFI: its boundary is the Python model/trainer contract.  This header is the
 * NEW native boundary below that contract.  Every entry point names the reference code it replaces (paths relative
 * to the reference repository root).  Conventions:
 *   - plain device pointers and sizes; no framework types; `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, UR_ERR_* (<0) for bad arguments, -(1000+cudaError_t) for a launch error;
 *   - no allocation, no ownership transfer, no host synchronisation, re-entrant across streams;
 *   - all matrices row-major fp32; feature dimensions (d, hidden sizes, leading dimensions) are multiples of 4 and
 *     pointers 16-byte aligned; id 0 is the padding row of every table (reference: nn.Embedding(padding_idx=0),
 *     unirec/model/base/reco_abc.py:167-170).
 */
#ifndef UNIREC_B200_H_
#define UNIREC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UR_ABI_VERSION 4

enum { UR_LOSS_SOFTMAX = 0, UR_LOSS_BPR = 1 };
enum { UR_ACT_NONE = 0, UR_ACT_SWISH = 1, UR_ACT_GELU = 2, UR_ACT_RELU = 3, UR_ACT_TANH = 4, UR_ACT_SIGMOID = 5 };
enum { UR_OPT_ADAM = 0, UR_OPT_ADAMW = 1, UR_OPT_SGD = 2, UR_OPT_SQNORM = 3 };

int ur_version(void);
/* 1 when the library was built with the tcgen05 GEMM path */
int ur_has_tensor_core_gemm(void);

/* ---- K1/K2: embedding row gather.  out[i,:] = table[idx[i],:], bit-exact.
 * replaces: item_embedding(items) unirec/model/base/recommender.py:67, :137; user_embedding(user_id) :43 */
int ur_gather_rows_f32(const float* table, int64_t n_rows, int d, const void* idx, int idx_bits /*32|64*/, int64_t n,
                       float* out, void* stream);

/* scalar tables: out[idx[e / idx_group] & idx_mask] += src[e], e < n.  replaces: autograd of `+ user_bias[user_id] + item_bias[item_id]`,
 * unirec/model/base/recommender.py:79-90 (idx_group 1: item bias, one id per entry; idx_group N: user bias, one id per sample) */
int ur_scatter_add_scalar_f32(float* out, const void* idx, int idx_bits, int64_t idx_group, int64_t idx_mask, const float* src, int64_t n,
                              void* stream);

/* BF16 table copy (opt-in reduced-precision mode, north_star 1e-2 bar): out[i,:] = float(table_bf16[idx[i],:]), d % 8 == 0 */
int ur_gather_rows_bf16(const void* table_bf16, int64_t n_rows, int d, const void* idx, int idx_bits /*32|64*/, int64_t n, float* out,
                        void* stream);

/* ---- K11 (exact-dense mode): grad[idx[i],:] += coef * src[i / src_group,:], rows with idx == pad_id skipped.
 * replaces: embedding_dense_backward under accelerator.backward, unirec/facility/trainer.py:346 */
int ur_scatter_add_rows_f32(float* grad, int64_t n_rows, int d, const void* idx, int idx_bits, int64_t n, const float* src,
                            int64_t src_group, const float* coef /*nullable*/, int64_t coef_group, int64_t pad_id,
                            void* stream);

/* ---- K10: sum-pool user tower. user_emb[b] = (user_table ? user_table[user_id[b]] : 0) + (len[b]+1)^-alpha * sum_l table[seq[b,l]]
 * replaces: AvgHist.forward_user_emb unirec/model/sequential/avghist.py:34-42; SVDPlusPlus.forward_user_emb svdplusplus.py:31-39 */
int ur_pool_sum_fwd_f32(const float* table, int d, const int32_t* item_seq, int64_t B, int L, const int64_t* item_seq_len,
                        float alpha, const float* user_table /*nullable*/, const int64_t* user_id, float* user_emb,
                        float* coeff_out /*[B], nullable*/, int world, int rank /*row-sharded tables: partial sum over the owned rows*/,
                        void* stream);

/* ---- K1+K3: Y = LayerNorm(table[item_seq] + pos[0..L-1]); saves per-row mean / rstd.
 * replaces: SASRec.forward_user_emb prologue, unirec/model/sequential/sasrec.py:60-69 */
int ur_seq_prep_ln_fwd_f32(const float* table, const float* pos /*nullable*/, const float* gamma, const float* beta, float eps,
                           const int32_t* item_seq, int64_t B, int L, int d, float* Y, float* mean, float* rstd,
                           const int32_t* tok_src /*nullable*/, const int32_t* n_tok_dev /*nullable*/,
                           const void* shard_ptrs /*nullable: W peer pointers (int64) to the row shards*/, int shard_world,
                           const int64_t* rng /*nullable*/, float drop_p, int drop_site, void* stream);
/* dX[B*L,d] = gradient wrt the gathered rows (feeds the row-sparse table update); dgamma/dbeta/dpos are ACCUMULATED */
int ur_seq_prep_ln_bwd_f32(const float* table, const float* pos, const float* gamma, const int32_t* item_seq, int64_t B, int L,
                           int d, const float* mean, const float* rstd, const float* dY, float* dX, float* dgamma, float* dbeta,
                           float* dpos /*nullable*/, const int32_t* tok_inv /*nullable*/, const void* shard_ptrs /*nullable*/,
                           int shard_world, const int64_t* rng /*nullable*/, float drop_p, int drop_site, void* stream);

/* ---- K6: post-LN residual.  X <- dropout(X) + R (kept for backward), Y = LayerNorm(X).
 * replaces: LayerNorm(dropout(hidden) + input) unirec/model/modules.py:313-314, :352-353
 * Dropout (all ops below): rng = device int64[2] (seed, training step), drop_p in [0,1), drop_site = id of the nn.Dropout
 * instance; rng == NULL or drop_p == 0 -> identity.  The mask of an element depends only on (seed, step, site, ORIGINAL token
 * position, column) -- csrc/dropout.cuh -- so backward kernels regenerate it; the original position of row r is row_pos[r]
 * (packed token map) or r * pos_mul + pos_add. */
int ur_add_ln_fwd_f32(float* X, int64_t ldx, const float* R /*nullable*/, int64_t ldr, const float* gamma, const float* beta,
                      float eps, int64_t rows, int d, float* Y, int64_t ldy, float* mean, float* rstd,
                      const int32_t* rows_dev /*nullable*/, const int64_t* rng /*nullable*/, float drop_p, int drop_site,
                      const int32_t* row_pos /*nullable*/, int64_t pos_mul, int64_t pos_add, void* stream);
/* dZ = LN'(Z) (dY + dExtra); dgamma/dbeta ACCUMULATED; dZ may alias dY; dzsum (nullable) += column sums of the gradient of the
 * linear layer's output, i.e. its bias gradient.  With dropout, dZdrop = dZ * mask is that gradient (dZ keeps flowing down the
 * residual branch); without, dZdrop is NULL and dZ serves both. */
int ur_add_ln_bwd_f32(const float* Z, int64_t ldz, const float* gamma, const float* mean, const float* rstd, const float* dY,
                      int64_t lddy, const float* dExtra /*nullable*/, int64_t ldde, int64_t rows, int d, float* dZ, int64_t lddz,
                      float* dgamma, float* dbeta, float* dzsum /*nullable*/, const int32_t* rows_dev /*nullable*/,
                      float* dZdrop /*nullable*/, int64_t lddzd, const int64_t* rng /*nullable*/, float drop_p, int drop_site,
                      const int32_t* row_pos /*nullable*/, int64_t pos_mul, int64_t pos_add, void* stream);

/* ---- K4/K7: C (+)= act(op(A) op(B) + bias); preact (nullable) receives the pre-activation for backward.
 * replaces: nn.Linear calls unirec/model/modules.py:285-287,312,348-351; gru.py:30-31
 * ur_gemm_f32 picks the tcgen05 tensor-core kernel when the shape qualifies and precision != 0, else the exact-fp32 SIMT kernel.
 * precision: 0 = fp32 FMA (exact path), 1 = TF32 tensor cores, 2 = BF16 (reserved: routed to the exact path),
 *            3 = 3xTF32 split on tensor cores (fp32-class accuracy). */
int ur_gemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
                float* C, int64_t ldc, const float* bias /*nullable*/, int act, float* preact /*nullable*/, int64_t ldp,
                int accumulate, int precision, void* stream);
int ur_gemm_simt_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                     void* stream);
/* tcgen05 / TMEM / TMA kernel behind ur_gemm_f32 (csrc/gemm_tc.cu): TF32 operands read in place from fp32, fp32 accumulate.
 * Supports NT (A [M,K], B [N,K]), NN (B stored [K,N]) and TN (A stored [K,M], B stored [K,N], split-K with atomic accumulate)
 * when N % 128 == 0 and K % 32 == 0; returns UR_ERR_UNSUPPORTED otherwise. */
int ur_gemm_tc_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                   int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                   int precision, void* stream);
/* ur_gemm_f32 with two fused epilogue stages (csrc/gemm_tc.cu; unfused SIMT route otherwise):
 *   dact   (nullable): C = (op(A) op(B)) * act'(dact[row, col])      -- replaces: autograd of the FFN activation, modules.py:348-351
 *   colsum (nullable): colsum[col] += sum_rows C[row, col]           -- the bias gradient of the layer that consumes C as dy
 * precision 3 = 3xTF32 split (hi/lo operand split in shared memory, fp32-class results from the tensor pipe).
 * rows_dev (nullable): device-resident live token count of the packed sequence layout (ur_pack_tokens); the launch is sized for
 *   the static M / K and the kernels stop at the live count, so no host synchronisation is needed. */
int ur_gemm_fused_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                      int64_t ldb, float* C, int64_t ldc, const float* bias /*nullable*/, int act, float* preact /*nullable*/,
                      int64_t ldp, int accumulate, int precision, const float* dact /*nullable*/, int64_t ldd,
                      float* colsum /*nullable*/, const int32_t* rows_dev /*nullable*/, int rows_dim, void* stream);
/* exact-fp32 kernel with the same device-resident token count (rows_dim 1: M = min(M, *rows_dev); 2: K = min(K, *rows_dev)) */
int ur_gemm_simt_rows_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                          int64_t ldb, float* C, int64_t ldc, const float* bias, int act, float* preact, int64_t ldp, int accumulate,
                          const int32_t* rows_dev, int rows_dim, void* stream);
/* 3xTF32 weight lo planes (replaces the per-tile in-kernel split of the B operand; the encoder weights of
 * unirec/model/modules.py:254-356 are constant within a step): ur_split_lo_f32 writes lo[i] = tf32(x[i] - trunc_tf32(x[i])) for a
 * whole parameter buffer; ur_gemm_set_lo_plane registers (buffer, plane, n floats) -- a later tensor-core GEMM in 3xTF32 mode whose B
 * operand lies inside the buffer fetches B's lo term from the plane.  n = 0 clears the registration.  The caller refreshes the plane
 * whenever the weights change (unirec_b200.engine does so at the start of every forward pass). */
int ur_split_lo_f32(const float* x, float* lo, int64_t n, void* stream);
int ur_gemm_set_lo_plane(const float* base, const float* lo_plane, int64_t n);
/* out[c][r] = in[r][c] for small weight matrices (dx = dy W is issued as an NT product on W^T) */
int ur_transpose_f32(const float* in, int64_t rows, int64_t cols, float* out, void* stream);
/* dY *= act'(preact) */
int ur_act_bwd_f32(float* dY, const float* preact, int64_t n, int act, void* stream);
/* out[n] += sum_m X[m,n] (bias gradients) */
int ur_colsum_accum_f32(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, const int32_t* rows_dev /*nullable*/, void* stream);

/* ---- K5: fused attention on packed QKV [B*L, 3d] with the SASRec additive mask (-10000), L <= 256, head dim in
 * {2,4,8,16,32,64,128} (stock SASRec.yaml: n_heads 16), attention-probability dropout (modules.py:307) by counter.
 * replaces: MultiHeadAttention.forward unirec/model/modules.py:289-311 and SASRec._get_attention_mask sasrec.py:40-57 */
/* offs / tok_src (nullable, both or none): packed token layout of ur_pack_tokens (sample b owns rows [offs[b], offs[b+1]) of qkv / ctx).
 * q_last (nullable, with q_only_last): the single query per sample comes from a compact [B, d] buffer, ctx / dctx are compact
 * [B, d] too and dQ goes to dq_last [B, d]. */
int ur_attn_fwd_f32(const float* qkv, const int32_t* item_seq, int64_t B, int L, int H, int dh, int causal, int q_only_last,
                    float* ctx, float* lse /*[B,H,L]*/, const int32_t* offs, const int32_t* tok_src, const float* q_last,
                    const int64_t* rng /*nullable*/, float drop_p, int drop_site, void* stream);
int ur_attn_bwd_f32(const float* qkv, const int32_t* item_seq, int64_t B, int L, int H, int dh, int causal, int q_only_last,
                    const float* ctx, const float* lse, const float* dctx, float* dqkv, const int32_t* offs, const int32_t* tok_src,
                    const float* q_last, float* dq_last, const int64_t* rng /*nullable*/, float drop_p, int drop_site, void* stream);

/* ---- token packing for the sequence towers (csrc/pack.cu): live positions = real items + position L-1 (+ every position of a
 * sequence without real items).  replaces: nothing in the reference -- it computes all B*L positions (sasrec.py:59-76); the dead ones
 * cannot reach the loss.  keep_all = 1 yields the identity map. */
int ur_pack_tokens(const int32_t* item_seq, int64_t B, int L, int keep_all, int32_t* offs /*[B+1]*/, int32_t* tok_src /*[B*L]*/,
                   int32_t* tok_inv /*[B*L]*/, int32_t* last_tok /*[B]*/, int32_t* n_tok /*[1]*/, void* stream);
/* zero rows [*n_dev, roundup32(*n_dev)) of X[rows_cap, width] */
int ur_zero_tail_rows_f32(float* X, int64_t ld, int width, const int32_t* n_dev, int64_t rows_cap, void* stream);

/* ---- dropout outside the fused LN / attention kernels (csrc/dropout.cu).
 * replaces: nn.Dropout on the gathered item embeddings, unirec/model/sequential/gru.py:29 (forward on the rows, backward on their
 * gradient: the same call).  ur_dropout_mask_f32 writes the multipliers (0 or 1/(1-p)) of a site: flat = 0 -> [rows, d] by row
 * position; flat = 1 -> rows*d consecutive elements (attention site: ((b*H+h)*L + i)*L + j).  Test hook: the CPU oracle takes them
 * as explicit factors.  ur_rng_advance: rng[1] += 1 (one mask set per training step). */
int ur_dropout_rows_f32(float* X, int64_t ld, int64_t rows, int d, const int32_t* row_pos /*nullable*/, int64_t pos_mul,
                        int64_t pos_add, const int32_t* rows_dev /*nullable*/, const int64_t* rng, float p, int site, void* stream);
int ur_dropout_mask_f32(float* out, int64_t rows, int d, const int32_t* row_pos /*nullable*/, int64_t pos_mul, int64_t pos_add,
                        const int64_t* rng, float p, int site, int flat, void* stream);
int ur_rng_advance(int64_t* rng, void* stream);

/* ---- K9: GRU cell pointwise parts (matrix products go through ur_gemm_f32).
 * replaces: nn.GRU inside GRU.forward_user_emb unirec/model/sequential/gru.py:30 */
int ur_gru_gate_fwd_f32(const float* gi, int64_t ld_gi, const float* gh, const float* h_prev, float* h_out, float* save /*[B,4H]*/,
                        int64_t B, int H, void* stream);
int ur_gru_gate_bwd_f32(const float* dh, const float* save, const float* h_prev, float* dgi, int64_t ld_dgi, float* dgh,
                        float* dh_prev, int64_t B, int H, void* stream);

/* Persistent recurrence: ONE launch per direction runs all L steps (a CTA owns 16 batch rows for the whole sequence, hidden tile in
 * shared memory, W_hh streamed from L2; csrc/gru.cu).  gi / dgi [B, L, 3H] batch-major; hs [L+1, B, H], hs[0] = initial state; save
 * [L, B, 4H]; whh_t = W_hh^T [H, 3H]; dgh_all [L, B, 3H] feeds the W_hh / b_hh gradient reductions.  H % 32 == 0, H <= 768 (stock GRU.yaml: 768).
 * replaces: nn.GRU (cuDNN RNN / ATen step loop) unirec/model/sequential/gru.py:30 and its autograd */
int ur_gru_seq_fwd_f32(const float* gi, const float* whh_t, const float* b_hh, float* hs, float* save, int64_t B, int64_t L, int H,
                       void* stream);
int ur_gru_seq_bwd_f32(const float* dh_last, const float* save, const float* hs, const float* whh, float* dgi, float* dgh_all, int64_t B,
                       int64_t L, int H, void* stream);

/* ---- K8: fused target/negative gather + scores + bias/tau/clamp + loss + dLoss/dscore + dLoss/du, one pass over the rows.
 * replaces: forward_item_emb + InnerProductScorer + _predict_layer + _cal_loss and their autograd:
 *   unirec/model/base/recommender.py:55-59,66-96; unirec/model/modules.py:15-21,45-67; unirec/model/base/reco_abc.py:252-265
 * norm: softmax -> number of positive labels in the batch (device scalar from ur_count_positive_i32, or norm_host);
 *       bpr -> B*K.  loss_vec[b] is the per-sample loss term; ur_loss_finish_f32 reduces it and raises the NaN flag. */
int ur_score_loss_fwd_bwd_f32(const float* table, int d, const float* user_emb, const int64_t* item_id, int64_t B, int N,
                              const int32_t* label /*nullable*/, const float* item_bias /*nullable*/,
                              const float* user_bias /*nullable*/, const int64_t* user_id, float tau, float score_clip,
                              int loss_type, const float* norm_dev /*nullable*/, float norm_host, float* scores /*nullable*/,
                              float* loss_vec, float* dscore /*nullable*/, float* grad_user /*nullable*/, void* stream);
int ur_count_positive_i32(const int32_t* label, int64_t n, float* out, void* stream);
int ur_loss_finish_f32(const float* loss_vec, int64_t B, const float* denom_dev /*nullable*/, float denom_host, float* loss_out,
                       int32_t* nan_flag /*nullable*/, void* stream);

/* ---- K11+K12+K13: row-sparse gradient reduce fused with the optimizer (see csrc/sparse_opt.cu).
 * replaces: dense embedding grads + optim.Adam over whole tables + clip_grad_norm_, unirec/facility/trainer.py:134-152,346-349
 * head[V] must hold -1 between steps (apply restores it); n_uniq must be zeroed by the caller before the first link of a step.
 * u_begin_dev / u_end_dev: device-resident bounds of the slice of the unique-row list to update (default: all of it); rows are
 * appended in link order, so linking the history keys first makes [n_hist, n_uniq) the rows only the scorer touches -- those are
 * updated on a side stream while the encoder backward runs (small_ctas = 1: CTAs sized to co-reside with a GEMM CTA). */
/* keys are masked with key_mask first (packed ids carry the label in bit 31); world > 1: keys are GLOBAL ids of a row-sharded table,
 * only entries with key % world == rank are linked, under the local row key / world. */
int ur_rowlist_link(int32_t* head, const void* keys, int idx_bits, int64_t n, int64_t entry_offset, int32_t* next, int32_t* uniq,
                    int32_t* n_uniq, int64_t pad_id, int world, int rank, int64_t key_mask, void* stream);
int ur_rowlist_apply_f32(float* table, float* mom, float* var, int d, int32_t* head, const int32_t* next, const int32_t* uniq,
                         const int32_t* n_uniq, int64_t max_uniq, const float* src0, int64_t src0_group, const float* coef0,
                         int64_t coef0_group, int64_t n0, const float* src1, int64_t src1_group, const float* coef1,
                         int64_t coef1_group, int mode, float lr, float beta1, float beta2, float eps, float weight_decay,
                         const int32_t* step_dev, const float* grad_scale_dev, const int32_t* skip_flag, float* sqnorm_out,
                         const int32_t* u_begin_dev /*nullable*/, const int32_t* u_end_dev /*nullable*/, int small_ctas,
                         const void* src1_part_ptrs /*nullable: peer pointers, source 1 rows live in per-rank buffers*/,
                         int64_t src1_part_rows, void* stream);

/* ---- peer memory for the row-sharded tables (csrc/p2p.cu): CUDA IPC export / open of device buffers between the per-GPU processes of
 * one box.  replaces: the reference's DDP all-reduce of dense table gradients (trainer.py:67,346) -- rows are read over NVLink instead. */
int ur_ipc_export(const void* dev_ptr, void* handle_out /*64 host bytes*/, int64_t* offset_out /*host*/);
int ur_ipc_open(const void* handle /*64 host bytes*/, int64_t offset, int64_t* ptr_out /*host*/);
int ur_dense_opt_f32(float* param, const float* grad, float* mom, float* var, int64_t n, int mode, float lr, float beta1,
                     float beta2, float eps, float weight_decay, const int32_t* step_dev, const float* grad_scale_dev,
                     const int32_t* skip_flag, void* stream);
int ur_sqnorm_accum_f32(const float* grad, int64_t n, float* sqnorm, void* stream);
int ur_clip_coef_f32(const float* sqnorm, float max_norm, float* coef, void* stream);
int ur_step_advance(int32_t* step, const int32_t* skip_flag, void* stream);

/* ---- a15 / f1: device-side batch builder: (user, positive) pairs + CSR user history -> the batch contract
 * (item_id [B,1+K] positive first, label [B,1+K], item_seq [B,L] left-padded, item_seq_len [B]).
 * replaces: AddNegSamples.__call__ unirec/data/transform/addnegsamples.py:90-115 (uniform or popularity-alias draws, <=100 retries,
 *           reject the positive and the user's history, id 0 on failure); AddUserHistory.__call__ adduserhistory.py:32-73
 *           (mask_mode 1 = unorder, 2 = autoregressive, seq_last); SeqRecDataset._padding seqrecdataset.py:60-68.
 * hist_* / alias_* / item_seq are nullable (L = 0: no history columns). */
int ur_build_batch(const int64_t* user_id, const int64_t* pos_item, int64_t B, const int64_t* hist_ptr, const int32_t* hist_items,
                   const int32_t* hist_sorted, int64_t n_users, int64_t n_items, const float* alias_prob, const int32_t* alias_idx,
                   int K, int L, int mask_mode, int seq_last, int64_t seed, int64_t step, int64_t* item_id, int32_t* label,
                   int32_t* item_seq, int64_t* item_seq_len, void* stream);

/* ---- f3: one-vs-all ranking on the device (csrc/evalrank.cu): rank of the target among all items without a [B,V] score matrix;
 * with row-sharded tables each rank counts over its own rows (world / rank as below) and the caller sums tscore and counts across ranks.
 * replaces: Evaluator.evaluate_with_full_items unirec/facility/evaluation/evaluator_abc.py:190-278 (scores :232-241, history / target /
 *           padding masking with NINF = -9999 :249-257) + get_rank unirec/facility/evaluation/onepos.py:20-31.
 * score(s, g) = (<u_s, e_g> + item_bias[g] + user_bias[user_id[s]]) / tau, every dot accumulated in k order in one thread (bit-identical
 * between the three kernels).  Exact ties are not counted (the reference breaks them with +-1e-8 noise, onepos.py:116-120).
 *   ur_rank_target  tscore[s] = score(s, target[s]) if owned else 0
 *   ur_rank_count   counts[s] += #{owned g : g != 0, g != target[s], score(s, g) > tscore[s]}
 *   ur_rank_exclude counts[s] += sum over DISTINCT owned history items h (h != 0, h != target[s]) of [NINF > t] - [score(s,h) > t], plus
 *                   [NINF > t] for the target's own slot; history = CSR (hist_ptr by user id, slices sorted ascending) */
int ur_rank_target_f32(const float* table_local, int d, const float* user_emb, const int64_t* target, int64_t S,
                       const float* item_bias /*nullable*/, const float* user_bias /*nullable*/, const int64_t* user_id /*nullable*/,
                       float tau, int world, int rank, float* tscore, void* stream);
int ur_rank_count_f32(const float* table_local, int64_t n_local, int d, const float* user_emb, int64_t S, const int64_t* target,
                      const float* tscore, const float* item_bias /*nullable*/, const float* user_bias /*nullable*/,
                      const int64_t* user_id /*nullable*/, float tau, int world, int rank, int32_t* counts, void* stream);
int ur_rank_exclude_f32(const float* table_local, int d, const float* user_emb, int64_t S, const int64_t* target, const float* tscore,
                        const float* item_bias /*nullable*/, const float* user_bias /*nullable*/, const int64_t* user_id /*nullable*/,
                        float tau, int world, int rank, const int64_t* hist_ptr /*nullable*/, const int32_t* hist_sorted /*nullable*/,
                        int64_t n_hist_users, int32_t* counts, void* stream);

/* ---- Row-sharded tables (multi-GPU): rank r of `world` owns rows {id : id % world == r} at local index id / world.
 * These replace the reference's replicated tables + DDP all-reduce of dense [V,d] gradients
 * (unirec/model/base/reco_abc.py:167-170, unirec/facility/trainer.py:67,346) together with NCCL collectives issued by the host
 * (all-gather of ids / user vectors, reduce-scatter of per-sample partial states): see unirec_b200/sharding.py. */
int ur_shard_gather_rows_f32(const float* table_local, int d, const void* idx, int idx_bits, int64_t n, int world, int rank,
                             float* out, void* stream);
int ur_shard_localize(const void* idx, int idx_bits, int64_t n, int world, int rank, int64_t pad_id, int32_t* out, void* stream);
/* ---- owner-side scoring of the row-sharded target table (csrc/shard_ring.cu), "move queries, not rows".
 * ur_pack_ids_i32: out[e] = id | (label > 0) << 31 (label NULL: positive in column 0) -- the ONE id tensor all-gathered per step.
 * ur_score_partial_f32: per sample of any rank, one pass over the OWNED rows (cp.async.bulk ring, d = 128 / 256; generic fallback):
 *   state[s] = (max, sum exp, sum y*s, sum y, sum p*e [d], sum y*e [d]), z[s, j] = raw score of owned entries.
 * ur_score_merge_f32: home rank merges the W partial states of its B samples (states [W, B, 4+2d], one all-to-all) -> loss_vec, (lse, n_y),
 *   dLoss/du.  ur_score_dscore_f32: owner, dLoss/d(dot) of owned entries.  norm_dev = global number of positive labels.
 * replaces: forward_item_emb + InnerProductScorer + _predict_layer + softmax _cal_loss (recommender.py:55-96, reco_abc.py:260-265) under
 *   DDP's replicated tables (trainer.py:67,346). */
int ur_pack_ids_i32(const int64_t* item_id, const int32_t* label /*nullable*/, int64_t B, int N, int32_t* out, void* stream);
int ur_count_positive_packed(const int32_t* ids_packed, int64_t n, float* out, void* stream);
int ur_score_partial_f32(const float* table_local, int d, const float* user_emb, const int32_t* ids_packed, int64_t S, int N,
                         const float* item_bias /*nullable*/, const float* user_bias /*nullable*/, const int64_t* user_id, float tau,
                         float score_clip, int world, int rank, float* z, float* state, void* stream);
int ur_score_merge_f32(const float* states, int world, int64_t B, int d, float tau, const float* norm_dev, float* loss_vec,
                       float* lse_ny, float* grad_user, void* stream);
int ur_score_dscore_f32(const float* z, const int32_t* ids_packed, const float* lse_ny, int64_t S, int N, int world, int rank,
                        float tau, float score_clip, const float* norm_dev, float* dscore, void* stream);
/* BPR on a row-sharded table (modules.py:15-21, reco_abc.py:252-255): owners score their entries (zeros elsewhere, summed by a
 * reduce-scatter), the home rank forms loss and dLoss/d(dot) from the complete [B, N] scores (norm = S*K), owners form their part of
 * dLoss/du (summed by a reduce-scatter). */
int ur_shard_scores_f32(const float* table_local, int d, const float* user_emb, const int32_t* ids_packed, int64_t S, int N,
                        const float* item_bias /*nullable*/, const float* user_bias /*nullable*/, const int64_t* user_id, float tau,
                        int world, int rank, float* z, void* stream);
int ur_bpr_from_scores_f32(const float* z, int64_t B, int N, float tau, float score_clip, float norm, float* loss_vec, float* dscore,
                           float* scores /*nullable*/, void* stream);
int ur_shard_grad_user_f32(const float* table_local, int d, const int32_t* ids_packed, const float* dscore, int64_t S, int N, int world,
                           int rank, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIREC_B200_H_ */

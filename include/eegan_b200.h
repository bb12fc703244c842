/*
 * eegan_b200.h — C ABI of libeegan_b200.so: the B200 (sm_100a) word-region attention /
 * DAMSM loss hot path of qikizh/EE-GAN.
 *
 * The reference has no FFI layer: its operator API for this path is the set of Python
 * signatures in miscc/DAMSM_losses.py and sync_batchnorm/ (SURVEY.md §8b).  This header
 * is what a binding for those functions binds (ctypes stub: eegan_b200/_lib.py; the
 * reference-side shim is shown in INTEGRATION.md).  Each entry point cites the reference
 * code it replaces (file:line relative to the reference repo).
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless stated otherwise.
 *   - the caller owns and allocates every buffer, including workspaces (sizes from the
 *     *_workspace_bytes queries).  The library never allocates/frees device memory, never
 *     synchronises the device and enqueues only on the stream passed (a cudaStream_t
 *     passed as void*; NULL = legacy default stream).
 *   - return 0 on success, non-zero on error; eegan_last_error() returns a thread-local
 *     message.  No exceptions cross the boundary.
 *   - all floating point tensors are contiguous fp32.  There is no CPU fallback.
 *   - thread safety: every compute entry point may be called concurrently from several host
 *     threads on different streams / devices (nn.DataParallel calls modules from per-device
 *     threads); the device is the caller's current device, the stream the one passed.
 *     Per-call state lives in the caller's buffers.  What IS process-wide, and why:
 *       * the engine selectors eegan_set_contraction_engine / eegan_set_gag_engine (A/B
 *         validation knobs, atomics; the backward of a forward runs on autograd's thread,
 *         so a thread-local selector would split one call pair over two engines),
 *       * the bench-only stage profiler eegan_profile_* (mutex-guarded, off by default),
 *       * caches that only memoise idempotent driver queries: the per-device shared-memory
 *         opt-in and SM count, and a thread-local cache of encoded TMA descriptors.
 *     Environment variables are read once per process (EEGAN_ENGINE, EEGAN_PDL, EEGAN_GAG_*
 *     tuning knobs), never per launch; work-skipping timing switches exist only in builds
 *     made with -DEEGAN_DEBUG_SWITCHES and are absent from the shipped library.
 */
#ifndef EEGAN_B200_H_
#define EEGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEGAN_B200_ABI_VERSION 1

#define EEGAN_OK 0
#define EEGAN_ERR_INVALID 1 /* bad argument (shape / null pointer / unsupported size) */
#define EEGAN_ERR_CUDA 2    /* a CUDA runtime call or launch failed */
#define EEGAN_ERR_WORKSPACE 3

int eegan_abi_version(void);
const char* eegan_last_error(void);

/* ------------------------------------------------------------------------------------
 * DAMSM pair grid — words_similarity / words_loss, miscc/DAMSM_losses.py:168-231,272-342
 * (the caption loop :281-321 over func_attention :25-63 and cosine_similarity :17-23).
 *
 * img    [B_img, D, R]       region features (NCHW with H*W = R flattened), fp32
 * words  [B_cap, D, T_max]   word embeddings, fp32
 * cap_lens [B_cap] int32     valid words per caption (1 <= len <= T_max <= 32)
 * m      [B_img, B_cap]      OUT: log sum_t exp(gamma2 * cos_t)  (:315-317), *before* the
 *                            gamma3 scale and class mask (applied by eegan_pair_ce_*).
 * att    [B_cap, T_max, R]   OUT (optional, may be NULL): region attention of caption i on
 *                            image i + diag_offset (:301); rows t >= len are zero.
 * workspace                  scratch + the stash the backward consumes; must stay intact
 *                            (and img/words unchanged) until eegan_damsm_pair_bwd returns.
 * diag_offset                image index of caption 0's matching image (caption-row-sharded
 *                            multi-GPU runs pass rank * B_cap; single GPU passes 0).
 * Constraints: D % 4 == 0, D <= 1024, R <= 1024, T_max <= 32.
 * ---------------------------------------------------------------------------------- */
size_t eegan_damsm_pair_workspace_bytes(int B_img, int B_cap, int D, int R, int T_max);

int eegan_damsm_pair_fwd(const float* img, const float* words, const int32_t* cap_lens,
                         int B_img, int B_cap, int D, int R, int T_max,
                         float gamma1, float gamma2,
                         float* m, float* att, int diag_offset,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the grid.  dm [B_img, B_cap] = dL/dm (already includes gamma3, zero at masked
 * cells).  d_img [B_img, D, R] and d_words [B_cap, D, T_max] are OVERWRITTEN (either may be
 * NULL to skip it; train.py:172 detaches the words so only d_img is needed there). */
int eegan_damsm_pair_bwd(const float* img, const float* words, const int32_t* cap_lens,
                         int B_img, int B_cap, int D, int R, int T_max,
                         float gamma1, float gamma2,
                         const float* dm, float* d_img, float* d_words,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The same backward in two calls, so that a caller can start the reduce-scatter of d_img (caption-row-sharded
 * multi-GPU runs) while d_words is still being computed:
 *   phases = 1  per-column scalars, dU, dS and d_img (d_words ignored);
 *   phases = 2  d_words from the dS stash phase 1 left in the workspace (d_img ignored);
 *   phases = 3  both = eegan_damsm_pair_bwd.   Split phases need the default contraction engine. */
int eegan_damsm_pair_bwd_phased(const float* img, const float* words, const int32_t* cap_lens,
                                int B_img, int B_cap, int D, int R, int T_max,
                                float gamma1, float gamma2,
                                const float* dm, float* d_img, float* d_words, int phases,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * func_attention — miscc/DAMSM_losses.py:25-63, as a stand-alone op (sample b's query against
 * sample b's context; the pair grid is its all-pairs form).
 * query [B, D, T], context [B, D, R] -> u [B, D, T] (weightedContext), attn [B, T, R].
 * Backward accepts grads on both outputs (either may be NULL) and overwrites d_query,
 * d_context.  The workspace written by the forward must be passed to the backward.
 * ---------------------------------------------------------------------------------- */
size_t eegan_func_attention_workspace_bytes(int B, int D, int R, int T);
int eegan_func_attention_fwd(const float* query, const float* context, int B, int D, int R,
                             int T, float gamma1, float* u, float* attn,
                             void* workspace, size_t workspace_bytes, void* stream);
int eegan_func_attention_bwd(const float* query, const float* context, const float* attn,
                             const float* d_u, const float* d_attn, int B, int D, int R, int T,
                             float gamma1, float* d_query, float* d_context,
                             void* workspace, size_t workspace_bytes, void* stream);

/* cosine_similarity — miscc/DAMSM_losses.py:17-23, row-wise over [rows, D] operands.
 * out [rows]; norms [rows, 2] (|x1|, |x2|) kept for the backward. */
int eegan_cosine_rows_fwd(const float* x1, const float* x2, long long rows, int D, float eps,
                          float* out, float* norms, void* stream);
int eegan_cosine_rows_bwd(const float* x1, const float* x2, const float* out,
                          const float* norms, const float* g, long long rows, int D, float eps,
                          float* d_x1, float* d_x2, void* stream);

/* ------------------------------------------------------------------------------------
 * Two-way cross-entropy over a B x B score grid — the tail of words_loss (:331-338) and
 * sent_loss (:258-267).
 *
 * scores_in [B,B] row-major (row = image, col = caption/sentence); scale multiplies it
 * (gamma3 for the words grid, 1 for sentence scores that already carry gamma3).
 * class_ids [B] int64 or NULL: cell (a,b), a != b, with class_ids[a]==class_ids[b] is set
 * to -inf (:282-285,331-333).  labels [B] int64; a label outside [0, B) (torch's
 * CrossEntropyLoss raises) never reads out of bounds: it poisons loss01 with NaN and
 * contributes no one-hot term to the backward.
 * scores_out [B,B]  OUT: scaled + masked grid (what words_similarity returns).
 * loss01 [2]        OUT: loss0 = CE(scores, labels) over rows, loss1 = CE(scores^T, labels).
 * lse [2,B]         OUT: row / column log-sum-exp, consumed by the backward.
 * Backward: dscores_in[a,b] = scale * ( g[0]*(softmax_row - onehot)/B + g[1]*(softmax_col - onehot)/B ),
 * zero at -inf cells.  g_loss01 [2] is a DEVICE pointer (upstream grads of loss0/loss1).
 * ---------------------------------------------------------------------------------- */
int eegan_pair_ce_fwd(const float* scores_in, float scale, const int64_t* class_ids,
                      const int64_t* labels, int B,
                      float* scores_out, float* loss01, float* lse, void* stream);
int eegan_pair_ce_bwd(const float* scores_out, const float* lse, const int64_t* labels,
                      const float* g_loss01, float scale, int B,
                      float* dscores_in, void* stream);

/* ------------------------------------------------------------------------------------
 * Sentence scores — sent_similarity / sent_loss, miscc/DAMSM_losses.py:134-166,233-270.
 * scores[i,j] = gamma3 * <cnn_i, rnn_j> / max(|cnn_i| |rnn_j|, eps)   (:253-258)
 * cnn, rnn [B, D]; norms [2,B] OUT (|cnn_i|, |rnn_j|) for the backward.
 * Backward takes dscores [B,B] and writes d_cnn, d_rnn [B,D].
 * ---------------------------------------------------------------------------------- */
int eegan_sent_scores_fwd(const float* cnn, const float* rnn, int B, int D, float gamma3,
                          float eps, float* scores, float* norms, void* stream);
int eegan_sent_scores_bwd(const float* cnn, const float* rnn, const float* norms,
                          const float* dscores, int B, int D, float gamma3, float eps,
                          float* d_cnn, float* d_rnn, void* stream);

/* Forward engine of eegan_gag_fwd: 1 = tcgen05 kernel (x staged in tensor memory, softmax in the accumulator's threads, P
 * written back to TMEM for out = P value^T; 3xTF32, default where the shape allows: Q % 4 == 0, idf % 16 == 0, idf <= 256),
 * 0 = CUDA-core kernels; process-wide. */
int eegan_set_gag_engine(int engine);
int eegan_get_gag_engine(void);
/* Backward engine of eegan_gag_bwd_ws: 1 = one-pass tcgen05 kernel (gag_tc_bwd.cu: all four contractions of the backward
 * of miscc/DAMSM_losses.py:96-132 on the tensor cores, bf16 hi/lo operand pairs, d_out / x / attn read once; idf = 32 / 64 /
 * 128, d_out given; default), 0 = CUDA-core kernels (gag_bwd2.cu); process-wide, for A/B validation. */
int eegan_set_gag_bwd_engine(int engine);
int eegan_get_gag_bwd_engine(void);

/* ------------------------------------------------------------------------------------
 * R-precision scoring — test.py:306-336 (Tester.cal_sim_one_by_one), the evaluation-side consumer of
 * the sentence-score arithmetic.
 * cnn_code [B, D]: global code of each generated image; rnn_codes [B, R_val, D]: its R_val candidate
 * sentence codes, candidate 0 = the ground-truth caption (test.py:321).
 * scores [B, R_val] OUT (optional): <cnn_b, rnn_bk> / max(|cnn_b| |rnn_bk|, eps)   (test.py:323-327)
 * best [B] int32 OUT (optional): argmax_k scores[b][k], lowest index on ties (torch.argmax)
 * hit [B] uint8 OUT (optional): best[b] == 0   (test.py:329-330, R_hits)
 * ---------------------------------------------------------------------------------- */
int eegan_rprecision(const float* cnn_code, const float* rnn_codes, int B, int R_val, int D, float eps,
                     float* scores, int32_t* best, uint8_t* hit, void* stream);

/* ------------------------------------------------------------------------------------
 * GlobalAttentionGeneral.forward — miscc/DAMSM_losses.py:75-132.
 * x [B, idf, Q]; key, value [B, idf, T]; mask [B, T] uint8 (1 = padding) or NULL.
 * mask_mode 0 = reference quirk: row (b,q) uses mask[(b*Q+q) % B] (:114-118);
 *           1 = intended: row (b,q) uses mask[b].
 * out [B, idf, Q] (weightedContext), attn [B, T, Q].
 * Backward accepts grads on both outputs (either may be NULL = zero) and writes d_x
 * [B,idf,Q], d_key, d_value [B,idf,T] (overwritten).
 * Constraints: T <= 32, idf <= 512.
 * ---------------------------------------------------------------------------------- */
int eegan_gag_fwd(const float* x, const float* key, const float* value, const uint8_t* mask,
                  int mask_mode, int B, int idf, int Q, int T,
                  float* out, float* attn, void* stream);
int eegan_gag_bwd(const float* x, const float* key, const float* value, const float* attn,
                  const float* d_out, const float* d_attn, int B, int idf, int Q, int T,
                  float* d_x, float* d_key, float* d_value, void* stream);
/* Backward with a caller-owned workspace (eegan_gag_bwd_workspace_bytes, 0 = shape not covered).  Default
 * (eegan_set_gag_bwd_engine(1)) for idf = 32 / 64 / 128 with d_out given: the one-pass tcgen05 kernel of gag_tc_bwd.cu
 * (all four contractions on the tensor cores, bf16 hi/lo operand pairs: gradients ~1e-5 of their maximum from float64;
 * d_out, x, attn, d_attn read once).  Otherwise the register-tiled CUDA-core form (gag_bwd2.cu: per-pixel pass writing
 * d_x and ds, then d_key / d_value as per-channel sums over pixel chunks).  Either way the d_key / d_value partials are
 * added in a fixed order: deterministic, no atomics.  Falls back to eegan_gag_bwd when workspace is NULL or the shape is
 * not covered (idf % 32 != 0, Q % 4 != 0, unaligned rows). */
size_t eegan_gag_bwd_workspace_bytes(int B, int idf, int Q, int T);
int eegan_gag_bwd_ws(const float* x, const float* key, const float* value, const float* attn,
                     const float* d_out, const float* d_attn, int B, int idf, int Q, int T, float* d_x,
                     float* d_key, float* d_value, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * SyncBatchNorm — sync_batchnorm/batchnorm.py:48-78 (per-replica statistics :56-62,
 * _compute_mean_std :113-125, normalise :71-75).  The cross-replica reduction itself is an
 * NCCL all-reduce of `stats` issued by the host side between these calls.
 * x [N, C, HW].  stats [2*C]: sum_c, then square-sum_c (fp32; one pass over x).
 * finalize: count = total elements per channel over all replicas (host value), or, when
 *   count_dev != NULL, the device pair {count / 4096, count % 4096} that travelled through
 *   the all-reduce next to the statistics (no host sync); writes mean, inv_std
 *   = clamp(var_biased, eps)^-1/2 (clamp_mode 1, the N-replica formula :125) or
 *   1/sqrt(var_biased + eps) (clamp_mode 0, F.batch_norm, :50-53), and updates
 *   running_mean / running_var (unbiased, momentum) when non-NULL.
 * apply: y = (x - mean) * (inv_std * weight) + bias  (weight/bias may be NULL).
 * bwd_reduce: red [2*C] = sum_c dy, sum_c dy * xhat.   (all-reduced by the host side)
 * bwd_apply: dx = w*inv_std * (dy - red0/count - xhat * red1/count); d_weight = red1,
 *   d_bias = red0 are taken from the LOCAL `red` by the host.  With clamp_mode 1 a channel
 *   whose variance was clamped (inv_std == eps^-1/2) has no variance term.
 * ---------------------------------------------------------------------------------- */
int eegan_syncbn_stats(const float* x, int N, int C, int HW, float* stats, void* stream);
/* stats [2*C + 2]: as above, followed by the local element count N*HW as the exact fp32 pair
 * {count / 4096, count % 4096}: statistics and count cross the replicas in ONE all-reduce
 * (batchnorm.py:102 reduces sum, ssum and sum_size together) and finalize takes
 * count_dev = stats + 2*C. */
int eegan_syncbn_stats_counted(const float* x, int N, int C, int HW, float* stats, void* stream);
int eegan_syncbn_finalize(const float* stats, int C, double count, const float* count_dev,
                          float eps, float momentum,
                          int clamp_mode, float* mean, float* inv_std,
                          float* running_mean, float* running_var, void* stream);
int eegan_syncbn_apply(const float* x, const float* mean, const float* inv_std,
                       const float* weight, const float* bias, int N, int C, int HW,
                       float* y, void* stream);
int eegan_syncbn_bwd_reduce(const float* x, const float* dy, const float* mean,
                            const float* inv_std, int N, int C, int HW, float* red,
                            void* stream);
int eegan_syncbn_bwd_apply(const float* x, const float* dy, const float* mean,
                           const float* inv_std, const float* weight, const float* red,
                           double count, const float* count_dev, float eps, int clamp_mode,
                           int N, int C, int HW,
                           float* dx, void* stream);

/* One replica (no cross-rank reduction to wait for): forward / backward as ONE call each.  work [4*C] floats: statistics
 * scratch [2C], then mean [C] and inv_std [C] (kept by the caller for the backward).  Maps of at most 64K elements per
 * channel run as a single launch (one CTA per channel; 14 of Gen's 24 layers are 32x32 or smaller), larger ones as the
 * stats / finalize / apply sequence above.  F.batch_norm arithmetic (batchnorm.py:50-53).  red [2*C] OUT: sum dy (d_bias),
 * sum dy * xhat (d_weight). */
int eegan_syncbn_fwd_fused(const float* x, const float* weight, const float* bias, int N, int C, int HW,
                           float eps, float momentum, float* running_mean, float* running_var,
                           float* y, float* work, void* stream);
int eegan_syncbn_bwd_fused(const float* x, const float* dy, const float* work, const float* weight,
                           int N, int C, int HW, float eps, float* dx, float* red, void* stream);
int eegan_ssa_fwd_fused(const float* x, const float* gamma, const float* beta, const float* mask,
                        int N, int C, int HW, float eps, float momentum, float* running_mean,
                        float* running_var, float* y, float* work, void* stream);

/* Contraction engine of the pair grid's GEMM-shaped stages (process-wide setting; all are sm_100a
 * kernels of this library):
 *   3 = tcgen05 "3xFP16": operands stored pre-split as fp16 hi/lo pairs with per-tensor power-of-two
 *       scales, attention fused into the GEMM epilogues (pair_grid_h.cu, gemm_h.cu)
 *   2 = tcgen05 3xTF32, operands split in the kernel, attention fused into the GEMM epilogues
 *   1 = tcgen05 3xTF32 with separate softmax kernels;  0 = exact-fp32 CUDA-core FFMA (validation). */
int eegan_set_contraction_engine(int engine);
int eegan_get_contraction_engine(void);

/* ------------------------------------------------------------------------------------
 * affine_ssa — models.py:43-86: SyncBN(affine=False) + mask-gated modulation
 *   out[n,c,hw] = (gamma[n,c] * mask[n,hw] + 1) * xhat[n,c,hw] + beta[n,c] * mask[n,hw]     (:84-86)
 * with xhat = (x - mean[c]) * inv_std[c] from eegan_syncbn_stats / _finalize (models.py:69).
 * x, y, dy, dx [N,C,HW]; gamma, beta, dgamma, dbeta [N,C]; mask, dmask [N,HW]; mean, inv_std [C].
 * bwd_reduce writes red [2C] = {sum dxh, sum dxh*xhat} (dxh = dy*(gamma*mask+1); all-reduce it across
 * replicas before bwd_apply, as for eegan_syncbn_bwd_reduce) and the complete d_gamma, d_beta, d_mask.
 * bwd_apply: dx = inv_std * (dxh - red[c]/count - xhat * red[C+c]/count); red == NULL = statistics were
 * constants (eval mode): dx = inv_std * dxh.
 * ---------------------------------------------------------------------------------- */
int eegan_ssa_apply(const float* x, const float* mean, const float* inv_std, const float* gamma,
                    const float* beta, const float* mask, int N, int C, int HW, float* y, void* stream);
int eegan_ssa_bwd_reduce(const float* x, const float* dy, const float* mean, const float* inv_std,
                         const float* gamma, const float* beta, const float* mask, int N, int C, int HW,
                         float* red, float* dgamma, float* dbeta, float* dmask, void* stream);
int eegan_ssa_bwd_apply(const float* x, const float* dy, const float* mean, const float* inv_std,
                        const float* gamma, const float* mask, const float* red, double count,
                        const float* count_dev, float eps, int clamp_mode, int N, int C, int HW,
                        float* dx, void* stream);

/* ------------------------------------------------------------------------------------
 * The tensor-core contraction engine on its own (tests / microbenchmarks):
 *   C[z][m][n] = sum_k A[z](m,k) * B[z](n,k), fp32 in / fp32 out, computed as fp32-accurate
 *   3xTF32 on tcgen05 (hi/lo operand split, three MMAs per K-step, fp32 accumulators in TMEM).
 * a_kmajor: A is [M][K] with pitch lda, else [K][M]; b_kmajor: B is [N][K] with pitch ldb, else
 * [K][N].  Pitches and batch strides are in elements and must be multiples of 4; bases 16-byte
 * aligned (TMA).  C is [M][N] with pitch ldc (any alignment).
 * staging: 0 = both operands are read by the tensor core from shared memory; 1 = the A operand is
 * staged in tensor memory (needs a_kmajor = 0 and b_kmajor = 1; the form the pair-grid pipeline uses).
 * ---------------------------------------------------------------------------------- */
int eegan_gemm_tf32x3(const float* A, const float* B, float* C, int M, int N, int K,
                      int a_kmajor, int b_kmajor, long long lda, long long ldb, long long ldc,
                      long long bsA, long long bsB, long long bsC, int batch, int staging, void* stream);

/* ------------------------------------------------------------------------------------
 * CNN_ENCODER.emb_features — DAMSM.py:162, 229: conv1x1(768, nef), bias=False, on the
 * 17x17 Inception map; its output is the img_features argument of words_loss.
 *   y[b][co][r] = sum_ci w[co][ci] x[b][ci][r]        x [B,Cin,R], w [Cout,Cin], y [B,Cout,R], fp32
 * computed as batched fp32-accurate 3xTF32 GEMMs on tcgen05.  Cin, Cout multiples of 4.
 * fwd also writes xp = x re-pitched to rows of 4*ceil(R/4) floats (eegan_conv1x1_workspace_bytes(.., 0) bytes,
 * caller-owned): the backward takes xp instead of x.  bwd: dx and/or dw may be NULL; workspace
 * eegan_conv1x1_workspace_bytes(.., 1) bytes.  dw is reduced over the batch deterministically.
 * ---------------------------------------------------------------------------------- */
size_t eegan_conv1x1_workspace_bytes(int B, int Cin, int Cout, int R, int backward);
int eegan_conv1x1_fwd(const float* x, const float* w, int B, int Cin, int Cout, int R, float* y, float* xp,
                      size_t xp_bytes, void* stream);
int eegan_conv1x1_bwd(const float* xp, const float* w, const float* dy, int B, int Cin, int Cout, int R, float* dx,
                      float* dw, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * ATTR_Enhance.forward — models.py:146-168: self-attention over [sent ; attrs]
 *   combine [B,1+attr_num,D];  q,k,v = Linear(combine);  a = softmax(q k^T, -1) * norm_fact (scale after the
 *   softmax, :166);  out = a v  (= attn_attrs [B,1+attr_num,D]; attn_sent = out[:,0,:]).
 * sent [B,D], attrs [B,attr_num,D], W* [D,D] (nn.Linear weight: [out][in]), b* [D]; attr_num <= 7.
 * fwd: three launches (cat; the three projections as one batched small GEMM; per-sample attention).  qkv
 * [3, B*(1+attr_num), D], p [B,1+attr_num,1+attr_num] and combine [B*(1+attr_num), D] are the backward's stash
 * (p may be NULL for inference; qkv and combine are required).
 * bwd: d_attn_sent [B,D] and/or d_attn_attrs [B,1+attr_num,D] (NULL = zero); g [3, B*(1+attr_num), D] and dtok
 * [B*(1+attr_num), D] are scratch; d_sent, d_attrs, db* may be NULL, the three dW* are computed together (all or
 * none).  Weight gradients are plain fixed-order sums (no atomics).
 * ---------------------------------------------------------------------------------- */
int eegan_attr_enhance_fwd(const float* sent, const float* attrs, const float* Wq, const float* bq,
                           const float* Wk, const float* bk, const float* Wv, const float* bv, int B, int D,
                           int attr_num, float norm_fact, float* out, float* qkv, float* p, float* combine,
                           void* stream);
int eegan_attr_enhance_bwd(const float* d_attn_sent, const float* d_attn_attrs, const float* combine,
                           const float* qkv, const float* p, const float* Wq, const float* Wk, const float* Wv,
                           int B, int D, int attr_num, float norm_fact, float* g, float* dtok, float* d_sent,
                           float* d_attrs, float* dWq, float* dbq, float* dWk, float* dbk, float* dWv, float* dbv,
                           void* stream);
size_t eegan_attr_enhance_workspace_bytes(int B, int D, int attr_num);

/* The half-pair engine on its own (tests / microbenchmarks): C[z] = A[z] B[z]^T, fp32 in / out.
 * A is MN-major ([K][lda], M contiguous), B is K-major ([N][ldb]); lda, ldb and the batch strides are
 * multiples of 8 elements.  sa, sb: power-of-two scales the operands are stored with (x*s must stay
 * below 65504).  dual != 0 runs the two-accumulator form on the same operands twice (second copy stored
 * with other scales): the result is 2 A B^T.  workspace: (dual ? 2 : 1) * 2 * (bytes of A/2 + bytes of
 * B/2, each rounded up to 256) + 256 bytes. */
int eegan_gemm_f16x3(const float* A, const float* B, float* C, int M, int N, int K, long long lda,
                     long long ldb, long long ldc, long long bsA, long long bsB, long long bsC, int batch,
                     float sa, float sb, int dual, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Bench-only stage timing of the multi-kernel entry points (eegan_damsm_pair_fwd/_bwd).
 * Process-wide switch, off by default (the backward runs on autograd's thread).  While enabled, each call records a cudaEvent after every
 * stage on the caller's stream; eegan_profile_collect synchronises those events (the only
 * synchronising entry point of the library) and returns, per stage, the summed device time
 * in ms and the number of times the stage ran since the previous collect.
 * ---------------------------------------------------------------------------------- */
#define EEGAN_PROF_NSTAGES 11
int eegan_profile_enable(int on);
int eegan_profile_nstages(void);
const char* eegan_profile_stage_name(int stage);
int eegan_profile_collect(double* stage_ms /*[NSTAGES], host*/, int* stage_launches /*[NSTAGES], host*/);

#ifdef __cplusplus
}
#endif
#endif /* EEGAN_B200_H_ */

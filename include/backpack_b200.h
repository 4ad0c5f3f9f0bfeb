/* backpack_b200.h -- C ABI of libbackpack_b200.so
 *
 * The drop-in boundary for the Backpack hot path on B200 (sm_100a).  These entry points replace the four
 * pybind11 extensions the reference's Python op wrappers import (paths relative to the reference repository):
 *
 *   flash_attn_cuda.fwd                       csrc/flash_attn/fmha_api.cpp:189-325      -> bp_fmha_fwd
 *   dropout_layer_norm.dropout_add_ln_fwd     csrc/layer_norm/ln_api.cpp:83-251         -> bp_ln_residual_fwd[_dropout]
 *   fused_dense_lib.linear_gelu_forward       csrc/fused_dense_lib/fused_dense.cpp:88-142 -> bp_linear_bias_act_fwd
 *   rotary_emb.apply_rotary                   csrc/rotary/rotary.cpp:12-33              -> bp_rotary_qk_inplace
 *   flash_attn_cuda.bwd                       csrc/flash_attn/fmha_api.cpp:338-500      -> bp_fmha_bwd
 *   dropout_layer_norm.dropout_add_ln_bwd     csrc/layer_norm/ln_api.cpp:255-440        -> bp_ln_residual_bwd[_dropout]
 *   fused_dense_lib backward epilogues        csrc/fused_dense_lib/fused_dense_cuda.cu:559-787 -> bp_bias_act_bwd
 *   xentropy_cuda_lib.forward / .backward     csrc/xentropy/interface.cpp:57-58         -> bp_xentropy_fwd / _bwd
 *   (no reference kernel; eager PyTorch at    training/src/models/backpack.py:111-122,313) -> bp_sense_lse_fwd,
 *                                                                                            bp_sense_mix_fwd,
 *                                                                                            bp_sense_mix_table_fwd
 *   (autograd through the same eager lines in training)                                   -> bp_sense_softmax_bwd
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, strides (in ELEMENTS), a cudaStream_t passed as void*.
 *   - every function is asynchronous on `stream`, never allocates device memory and never synchronises;
 *     the caller owns every buffer (the reference allocates inside the extension, fmha_api.cpp:273-280).
 *   - return value 0 = success, negative = bp_status_t; bp_last_error() gives a thread-local message.
 *     (The reference raises C++ exceptions through TORCH_CHECK, fmha_api.cpp:210-252; the Python
 *     binding turns a non-zero status into RuntimeError with that message.)
 *   - the launch targets the CUDA device current on the calling thread.
 *   - dtype: activations are f16 or bf16 (fmha_api.cpp:215-221); statistics and residual streams f32.
 */
#ifndef BACKPACK_B200_H_
#define BACKPACK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_ABI_VERSION 1

typedef enum bp_dtype { BP_DTYPE_F16 = 0, BP_DTYPE_BF16 = 1, BP_DTYPE_F32 = 2 } bp_dtype_t;

typedef enum bp_status {
  BP_OK = 0,
  BP_ERR_INVALID_ARGUMENT = -1, /* shape / stride / dtype / alignment rejected */
  BP_ERR_UNSUPPORTED = -2,      /* valid in the reference but not implemented here */
  BP_ERR_CUDA = -3,             /* a CUDA runtime / driver call failed (message has the error string) */
  BP_ERR_ARCH = -4              /* device is not compute capability 10.x */
} bp_status_t;

typedef enum bp_activation { BP_ACT_NONE = 0, BP_ACT_GELU_TANH = 1 } bp_activation_t;

/* Library identification. */
int bp_abi_version(void);
const char* bp_last_error(void);
/* Compute capability check for the current device (0 when sm_100-class, BP_ERR_ARCH otherwise). */
int bp_check_device(void);

/* FlashAttention forward (replaces mha_fwd, csrc/flash_attn/fmha_api.cpp:189-325).
 *   q, k, v, out : (total_tokens, nheads, headdim) views, last-dim stride 1, row/head strides in elements
 *                  (packed qkv views are passed un-copied, as flash_attn_interface.py:59 does).
 *   softmax_lse  : (batch, nheads, lse_stride) f32, natural-log sum-exp of the scaled scores.
 *   cu_seqlens_* : (batch+1) int32 device arrays (fmha_api.cpp:230-233); sequence i occupies rows
 *                  [cu[i], cu[i+1]).  Causal masking is top-left aligned: key j visible to query i iff j<=i
 *                  (csrc/flash_attn/src/fmha/mask.h:70).
 *   headdim % 8 == 0 and headdim <= 128 (fmha_api.cpp:245).  No dropout (see bp_fmha_fwd_dropout).
 */
int bp_fmha_fwd(const void* q, const void* k, const void* v, void* out, float* softmax_lse,
                const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                int32_t batch, int32_t nheads, int32_t headdim,
                int32_t total_q, int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k,
                int64_t q_row_stride, int64_t q_head_stride,
                int64_t k_row_stride, int64_t k_head_stride,
                int64_t v_row_stride, int64_t v_head_stride,
                int64_t o_row_stride, int64_t o_head_stride,
                int32_t lse_stride, float softmax_scale, int32_t is_causal,
                int32_t dtype /* bp_dtype_t */, void* stream);

/* FlashAttention backward (replaces mha_bwd, csrc/flash_attn/fmha_api.cpp:338-500, as driven by
 * _flash_attn_backward, flash_attn/flash_attn_interface.py:31-47).  SURVEY.md section 8f rank 4.
 *   dout, q, out, dq : (total_q, nheads, headdim);  k, v, dk, dv : (total_k, nheads, headdim); last-dim stride 1.
 *   softmax_lse      : what bp_fmha_fwd returned.  No dropout (see bp_fmha_bwd_dropout).  dq / dk / dv are overwritten (not accumulated into);
 *                      the result is bit-wise reproducible (no atomics; the reference's determinism contract,
 *                      tests/test_flash_attn.py:727-793).
 *   strides          : HOST array of 16 element strides: {row, head} of dout, q, k, v, out, dq, dk, dv.
 *   workspace        : device scratch of at least bp_fmha_bwd_workspace_bytes(batch, nheads, max_seqlen_q) bytes,
 *                      16-byte aligned (per-row {lse, sum_d dout*out}; the reference returns the latter as softmax_d).
 * Three kernel launches on `stream`: row statistics, (dK, dV), dQ.
 */
int64_t bp_fmha_bwd_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_q);
int bp_fmha_bwd(const void* dout, const void* q, const void* k, const void* v, const void* out,
                const float* softmax_lse, void* dq, void* dk, void* dv,
                const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                int32_t batch, int32_t nheads, int32_t headdim,
                int32_t total_q, int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k,
                const int64_t* strides, int32_t lse_stride, float softmax_scale, int32_t is_causal,
                int32_t dtype /* bp_dtype_t */, void* workspace, int64_t workspace_bytes, void* stream);

/* Attention dropout (training): the same two operators with a Bernoulli keep mask on the attention probabilities,
 * out = ((D o P) / (1 - p)) V (csrc/flash_attn/src/fmha/softmax.h apply_dropout; fmha_api.cpp:189-204 p_dropout, gen).
 * The mask is counter-based -- keep(b, h, q, k) is a pure function of (seed, b * nheads + h, q, k), spelled out in
 * csrc/bp_common.cuh and restated in Python by flash_attn_interface.attention_dropout_mask -- so the backward
 * regenerates it from `seed` alone (the reference replays its Philox stream instead).  p is quantised to
 * round(256 p) / 256.  Workspaces: the forward needs bp_fmha_fwd_dropout_workspace_bytes, the backward
 * bp_fmha_bwd_dropout_workspace_bytes (both 16-byte aligned).  p_dropout = 0 is the plain operator.
 */
int64_t bp_fmha_fwd_dropout_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_k);
int bp_fmha_fwd_dropout(const void* q, const void* k, const void* v, void* out, float* softmax_lse,
                        const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                        int32_t batch, int32_t nheads, int32_t headdim,
                        int32_t total_q, int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k,
                        int64_t q_row_stride, int64_t q_head_stride,
                        int64_t k_row_stride, int64_t k_head_stride,
                        int64_t v_row_stride, int64_t v_head_stride,
                        int64_t o_row_stride, int64_t o_head_stride,
                        int32_t lse_stride, float softmax_scale, int32_t is_causal,
                        int32_t dtype /* bp_dtype_t */, float p_dropout, uint64_t seed,
                        void* workspace, int64_t workspace_bytes, void* stream);
int64_t bp_fmha_bwd_dropout_workspace_bytes(int32_t batch, int32_t nheads, int32_t max_seqlen_q, int32_t max_seqlen_k);
int bp_fmha_bwd_dropout(const void* dout, const void* q, const void* k, const void* v, const void* out,
                        const float* softmax_lse, void* dq, void* dk, void* dv,
                        const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                        int32_t batch, int32_t nheads, int32_t headdim,
                        int32_t total_q, int32_t total_k, int32_t max_seqlen_q, int32_t max_seqlen_k,
                        const int64_t* strides, int32_t lse_stride, float softmax_scale, int32_t is_causal,
                        int32_t dtype /* bp_dtype_t */, float p_dropout, uint64_t seed,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Backpack sense-mix, pass 1: per-sense causal softmax statistics.
 *   qk  : (batch, seqlen, 2, nv, dk) contiguous -- the output of ContextSelfAttn.Wqkv reshaped as
 *         training/src/models/backpack.py:111-116 (q = [:, :, 0], k = [:, :, 1]).
 *   lse : (batch, nv, seqlen) f32; lse[b,l,i] = log sum_{j<=i} exp(scale * q_li . k_lj).
 */
int bp_sense_lse_fwd(const void* qk, float* lse, int32_t batch, int32_t seqlen, int32_t nv, int32_t dk,
                     float softmax_scale, int32_t dtype, void* stream);

/* Backpack sense-mix, pass 2: out[b,i,:] = sum_l sum_{j<=i} softmax_j(scale q_li.k_lj) * content[b,l,j,:]
 * (backpack.py:117-122 + :313 fused; alpha (b,nv,s,s) is never materialised).
 *   content : element (b,l,j,c) at content + b*c_batch_stride + l*c_sense_stride + j*c_row_stride + c
 *             (the reference hands a transposed view of (b,s,nv,d), backpack.py:276).
 *   out     : (batch, seqlen, d) contiguous, same dtype.   lse: from bp_sense_lse_fwd.
 */
int bp_sense_mix_fwd(const void* qk, const void* content, const float* lse, void* out,
                     int32_t batch, int32_t seqlen, int32_t nv, int32_t dk, int32_t d,
                     int64_t c_batch_stride, int64_t c_sense_stride, int64_t c_row_stride,
                     float softmax_scale, int32_t dtype, void* stream);

/* Backpack sense-mix, pass 2, with the sense vectors gathered from a precomputed table inside the kernel:
 *   out[b,i,:] = sum_l sum_{j<=i} softmax_j(scale q_li.k_lj) * table[input_ids[b,j], l, :]
 * C_l(x) depends on the token id only (BackpackContentModule sees the word embedding without positions and has
 * an identity mixer: training/src/models/backpack.py:258, 125-143; the reference's analysis scripts rely on the
 * same fact, training/src/run_simlex.py:179-184), so for inference content_model(input_ids) is a row gather from a
 * (vocab, nv, d) table computed once with the same kernels.  The gather is done inside the kernel (16-byte cp.async
 * copies straight into the swizzled operand tiles): no (batch, seqlen, nv, d) tensor is ever written to or read from HBM.
 *   table     : (vocab, nv, d) contiguous, same dtype as qk.   input_ids : (batch, seqlen) int64 contiguous;
 *               ids are clamped to [0, vocab) (the caller validates them; the kernel cannot raise).
 *   lse       : from bp_sense_lse_fwd.   out : (batch, seqlen, d) contiguous.
 */
int bp_sense_mix_table_fwd(const void* qk, const void* table, const int64_t* input_ids, const float* lse, void* out,
                           int32_t batch, int32_t seqlen, int32_t nv, int32_t dk, int32_t d, int32_t vocab,
                           float softmax_scale, int32_t dtype, void* stream);

/* Residual add + LayerNorm forward without dropout (replaces dropout_add_ln_fwd with dropout_p = 0, no
 * rowscale/colscale/subset; csrc/layer_norm/ln_api.cpp:83-251, ln_fwd_kernels.cuh:98-188).
 *   x0 (rows, cols) x0_dtype; x1 (rows, cols) residual_dtype or NULL;
 *   x_out = x0 + x1 written in residual_dtype (may be NULL when not needed: prenorm=False);
 *   z = gamma * (x_out - mean) * rstd + beta in x0_dtype;  mu / rsigma (rows) f32, may be NULL.
 */
int bp_ln_residual_fwd(const void* x0, const void* x1, const void* gamma, const void* beta,
                       void* z, void* x_out, float* mu, float* rsigma,
                       int64_t rows, int32_t cols, float epsilon,
                       int32_t x0_dtype, int32_t residual_dtype, int32_t weight_dtype, void* stream);

/* out = act(x W^T + bias)  (replaces linear_gelu_forward, csrc/fused_dense_lib/fused_dense.cpp:88-142,
 * and the plain F.linear of FusedDense, flash_attn/ops/fused_dense.py:52).
 *   x (m, k) row-major, W (n, k) row-major (nn.Linear layout), bias (n) or NULL, out (m, n) row-major.
 *   k % 8 == 0, n % 8 == 0.
 */
int bp_linear_bias_act_fwd(const void* x, const void* w, const void* bias, void* out,
                           int64_t m, int32_t n, int32_t k, int32_t activation /* bp_activation_t */,
                           int32_t dtype, void* stream);

/* The same GEMM with the pre-activation as a second output: out = act(x W^T + bias), pre_out = x W^T + bias, both
 * (m, n) row-major in `dtype`, from one pass over the accumulators (replaces linear_gelu_forward with save_gelu_in =
 * true, csrc/fused_dense_lib/fused_dense.cpp:88-142, which training uses so that the backward need not recompute the
 * fc1 GEMM: flash_attn/ops/fused_dense.py:220-222).  m >= 256 (BP_ERR_UNSUPPORTED below); out != pre_out.
 */
int bp_linear_bias_act_aux_fwd(const void* x, const void* w, const void* bias, void* out, void* pre_out,
                               int64_t m, int32_t n, int32_t k, int32_t activation /* bp_activation_t */,
                               int32_t dtype, void* stream);

/* Tied LM head with the softmax statistics fused into the GEMM epilogue: logits = x W^T are never written to HBM
 * (6.6 GB per forward at Backpack-Small / batch 64 / seq 1024).  What evaluation and greedy decoding consume:
 *   lse[i]          = log sum_j exp(logits[i, j])          (the cross-entropy is lse[i] - target_logit[i])
 *   argmax[i]       = first j attaining max_j logits[i, j]  (torch.argmax order), max_logit[i] its value
 *   target_logit[i] = logits[i, targets[i]]                 (targets / target_logit may both be NULL)
 * over the columns j < n_valid (n_valid = n keeps the reference's padded vocabulary, backpack.py:285-288).
 * Replaces, for inference, lm_head (training/src/models/backpack.py:349) followed by the softmax cross-entropy of
 * csrc/xentropy/xentropy_kernel.cu:430-760 / flash_attn/losses/cross_entropy.py:19-129, and the
 * logits[:, -1].argmax of the generation loop (training/src/utils/generation.py:34-44).  The statistics are taken on
 * the fp32 accumulators (the reference rounds the logits to 16 bits first).
 *   x (m, k) row-major 16-bit, W (n, k) row-major (the embedding matrix), k % 8 == 0; targets (m) int64.
 */
int bp_lm_head_stats_fwd(const void* x, const void* w, const int64_t* targets, float* lse, int32_t* argmax,
                         float* max_logit, float* target_logit, int64_t m, int32_t n, int32_t k, int32_t n_valid,
                         int32_t dtype, void* stream);

/* Incremental decoding, attention of ONE new query per (batch, head) against the KV cache: the seqlen_q = 1 form of
 * MHA.forward with inference_params (flash_attn/modules/mha.py:356-380 _update_kv_cache, :432-440 inner_cross_attn with
 * causal=False -- the mask is top-left aligned, csrc/flash_attn/src/fmha/mask.h:70, so the new token sees every cached key).
 *   q (batch, nheads, headdim), out (batch, nheads, headdim), 16-bit, contiguous;
 *   kv_cache: the reference's (max_batch, max_seqlen, 2, nheads, headdim) cache, addressed as
 *     kv_cache + b * kv_batch_stride + j * kv_row_stride + which * kv_which_stride + h * headdim   (elements; which 0 = K, 1 = V)
 *   seqlen_k keys are attended to (the new token's K/V must already be in the cache), or seqlens_k[b] (device, int32)
 *   per batch element when not NULL.  headdim 64 or 128.
 */
int bp_decode_attn_fwd(const void* q, const void* kv_cache, void* out, const int32_t* seqlens_k, int32_t batch,
                       int32_t nheads, int32_t headdim, int32_t seqlen_k, int64_t kv_batch_stride, int64_t kv_row_stride,
                       int64_t kv_which_stride, float softmax_scale, int32_t dtype, void* stream);

/* Incremental decoding, Backpack sense-mix of the LAST position only:
 *   out[b, :] = sum_l sum_{j < len} softmax_j(scale * q[b,l,:] . k_cache[b,j,l,:]) * table[ids[b,j], l, :]
 * i.e. row i = len - 1 of training/src/models/backpack.py:116-122 + :313 with the sense vectors served from the
 * (vocab, nv, d) table of bp_sense_mix_table_fwd.  The reference has no incremental path for Backpacks (its generation
 * loop re-runs the whole forward per token, training/src/utils/generation.py:34-44, 62-72); SURVEY.md section 8 row F2.
 *   q (batch, nv, dk) contiguous; k_cache (max_batch, max_seqlen, nv, dk) with batch stride k_batch_stride (elements);
 *   ids (max_batch, max_seqlen) int64 with batch stride ids_batch_stride, clamped to [0, vocab); out (batch, d).
 *   len = seqlen, or seqlens[b] (device, int32) when not NULL.  dk % 8 == 0, d % 8 == 0.
 */
int bp_sense_mix_decode_fwd(const void* q, const void* k_cache, const int64_t* ids, const void* table, void* out,
                            const int32_t* seqlens, int32_t batch, int32_t seqlen, int32_t nv, int32_t dk, int32_t d,
                            int32_t vocab, int64_t k_batch_stride, int64_t ids_batch_stride, float softmax_scale,
                            int32_t dtype, void* stream);

/* residual += x W^T + bias, fp32 residual stream updated in place: the out_proj / fc2 GEMM of a pre-norm block
 * with the "dropout(0) + add" half of dropout_add_ln_fwd (flash_attn/modules/block.py:84-88, 101-105;
 * csrc/layer_norm/ln_fwd_kernels.cuh:98-131) moved into its epilogue, so the branch output is never written
 * to HBM in 16 bits and read back.  The fp32 accumulator is added un-rounded (the reference rounds the branch
 * to bf16/fp16 first), i.e. this is at least as close to exact arithmetic as the reference.
 *   x (m, k) row-major 16-bit, W (n, k) row-major, bias (n) or NULL, residual (m, n) row-major f32.
 *   k % 8 == 0, n % 8 == 0, m >= 256.
 */
int bp_linear_bias_residual_fwd(const void* x, const void* w, const void* bias, float* residual,
                                int64_t m, int32_t n, int32_t k, int32_t dtype, void* stream);

/* z = LayerNorm(x): the "LayerNorm" half of dropout_add_ln_fwd on an already-summed residual stream
 * (x (rows, cols) in x_dtype, typically f32; z in z_dtype, typically the 16-bit activation dtype; gamma / beta
 * in weight_dtype; mu / rsigma (rows) f32 or NULL).
 */
int bp_ln_fwd(const void* x, const void* gamma, const void* beta, void* z, float* mu, float* rsigma,
              int64_t rows, int32_t cols, float epsilon, int32_t x_dtype, int32_t z_dtype, int32_t weight_dtype,
              void* stream);

/* Backward of bp_ln_residual_fwd (replaces dropout_add_ln_bwd, csrc/layer_norm/ln_api.cpp:255-440 as driven by
 * _dropout_add_layer_norm_backward, flash_attn/ops/layer_norm.py:27-47; dropout 0, no rowscale / colscale).
 *   dz (rows, cols) in x0_dtype; x (rows, cols) in residual_dtype: the pre-norm sum x0 + x1 the forward wrote to
 *   x_out (x0 itself when the forward wrote none); dx_residual: gradient arriving through the residual output of a
 *   pre-norm block (residual_dtype) or NULL; gamma in weight_dtype.
 *   dx0 (x0_dtype), dx1 (residual_dtype, or NULL when the forward had no x1), dgamma / dbeta (weight_dtype).
 *   mu / rsigma (rows) f32: the statistics bp_ln_residual_fwd wrote, or both NULL (they are recomputed from x).
 *   workspace: bp_ln_bwd_workspace_bytes(cols) bytes, 16-byte aligned.
 * Two launches (row pass + column-sum finalisation); deterministic.
 */
int64_t bp_ln_bwd_workspace_bytes(int32_t cols);
int bp_ln_residual_bwd(const void* dz, const void* dx_residual, const void* x, const void* gamma,
                       const float* mu, const float* rsigma, void* dx0, void* dx1,
                       void* dgamma, void* dbeta, void* workspace, int64_t workspace_bytes, int64_t rows, int32_t cols,
                       float epsilon, int32_t x0_dtype, int32_t residual_dtype, int32_t weight_dtype, void* stream);

/* The same two functions with the reference's residual dropout inside the kernel (dropout_add_ln_fwd / _bwd with
 * dropout_p > 0, ln_fwd_kernels.cuh:96-131, ln_bwd_kernels.cuh): x = dropout(x0) / (1 - p) + x1 in the forward,
 * dx0 = keep ? dx / (1 - p) : 0 in the backward (dx1 = dx).  The keep mask is a pure function of (seed, row, col) --
 * one hashed word per row, one per column, bp_common.cuh -- so the forward stores no mask and the backward regenerates
 * it from the same seed.  dropout_p in [0, 1) is quantised to round(256 p) / 256 (at least 1 / 256 when p > 0);
 * dropout_p == 0 is exactly bp_ln_residual_fwd / _bwd.  rows < 2^32.
 */
int bp_ln_residual_fwd_dropout(const void* x0, const void* x1, const void* gamma, const void* beta,
                               void* z, void* x_out, float* mu, float* rsigma,
                               int64_t rows, int32_t cols, float epsilon,
                               int32_t x0_dtype, int32_t residual_dtype, int32_t weight_dtype,
                               float dropout_p, uint64_t seed, void* stream);
int bp_ln_residual_bwd_dropout(const void* dz, const void* dx_residual, const void* x, const void* gamma,
                               const float* mu, const float* rsigma, void* dx0, void* dx1,
                               void* dgamma, void* dbeta, void* workspace, int64_t workspace_bytes, int64_t rows,
                               int32_t cols, float epsilon, int32_t x0_dtype, int32_t residual_dtype,
                               int32_t weight_dtype, float dropout_p, uint64_t seed, void* stream);

/* Backward of the bias + activation epilogue of bp_linear_bias_act_fwd: what the reference gets from cuBLASLt's
 * BGRADB / DGELU_BGRAD epilogues (csrc/fused_dense_lib/fused_dense_cuda.cu:559-787; flash_attn/ops/fused_dense.py:
 * 57-80, 239-300).  The backward GEMMs themselves stay plain GEMMs.
 *   BP_ACT_GELU_TANH: dpre = dact * gelu_tanh'(pre) (all (m, n) row-major, 16-bit), dbias = sum_rows dpre (or NULL).
 *   BP_ACT_NONE     : dbias = sum_rows dact; pre / dpre ignored.
 *   workspace: bp_bias_act_bwd_workspace_bytes(n) bytes when dbias is requested.  n % 8 == 0.  Deterministic.
 */
int64_t bp_bias_act_bwd_workspace_bytes(int32_t n);
int bp_bias_act_bwd(const void* dact, const void* pre, void* dpre, void* dbias, void* workspace,
                    int64_t workspace_bytes, int64_t m, int32_t n, int32_t activation, int32_t dtype, void* stream);

/* Softmax cross-entropy over (rows, vocab) logits, forward and backward (replaces xentropy_cuda_lib.forward /
 * .backward, csrc/xentropy/interface.cpp:57-58, as driven by SoftmaxCrossEntropyLossFn,
 * flash_attn/losses/cross_entropy.py:19-109).  logits in f16 / bf16 / f32 with a row stride in elements; labels int64;
 * losses / lse / grad_losses (rows) f32.  loss = (1 - smoothing) (lse - x[y]) + smoothing (lse - sum(x) / C), 0 where
 * y == ignore_index; C = total_classes (-1: vocab).  The backward writes g (softmax - (1 - smoothing) onehot -
 * smoothing / C) into grad_logits, which MAY BE the logits buffer itself (the reference's inplace_backward).
 * The evaluation-only form that never materialises logits is bp_lm_head_stats_fwd.
 */
int bp_xentropy_fwd(const void* logits, const int64_t* labels, float* losses, float* lse, int64_t rows, int32_t vocab,
                    int64_t row_stride, float smoothing, int64_t ignore_index, int32_t total_classes, int32_t dtype,
                    void* stream);
int bp_xentropy_bwd(const float* grad_losses, const void* logits, const float* lse, const int64_t* labels,
                    void* grad_logits, int64_t rows, int32_t vocab, int64_t row_stride, int64_t grad_row_stride,
                    float smoothing, int64_t ignore_index, int32_t total_classes, int32_t dtype, void* stream);

/* Element-wise half of the sense-mix backward (the backward of `softmax(q k^T scale + causal mask)` inside
 * ContextSelfAttn.forward, training/src/models/backpack.py:116-122, which the reference leaves to autograd), in place:
 *   scores_probs   (rows, seqlen) 16-bit: raw scores q.k in, probabilities P = softmax_{j<=t}(scale * scores) out;
 *   dalpha_dscores (rows, seqlen) 16-bit: dL/dP in, scale * P o (dL/dP - rowsum(P o dL/dP)) out;
 * row r belongs to query t = r % seqlen (square causal matrices stacked along rows); entries right of the diagonal are
 * ignored on input and written as zeros.  seqlen % 8 == 0, seqlen <= 8192 (BP_ERR_UNSUPPORTED otherwise).
 */
int bp_sense_softmax_bwd(void* scores_probs, void* dalpha_dscores, int64_t rows, int32_t seqlen, float softmax_scale,
                         int32_t dtype, void* stream);

/* In-place rotary embedding on q and k of a packed qkv tensor (replaces apply_rotary as driven by
 * ApplyRotaryEmbQKV_.forward, flash_attn/layers/rotary.py:81-105; csrc/rotary/rotary_cuda.cu:5-41).
 *   qkv (batch, seqlen, 3, nheads, headdim) contiguous; cos/sin (seqlen, rotary_dim/2) in qkv's dtype;
 *   cos_k/sin_k: tables for k (XPos), or NULL to reuse cos/sin.  Non-interleaved (GPT-NeoX) pairing.
 */
int bp_rotary_qk_inplace(void* qkv, const void* cos, const void* sin, const void* cos_k, const void* sin_k,
                         int32_t batch, int32_t seqlen, int32_t nheads, int32_t headdim, int32_t rotary_dim,
                         int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BACKPACK_B200_H_ */

/* libb200timm.so -- C ABI of the B200 (sm_100a) kernels behind timm's ViT / ConvNeXt hot path.
 *
 * The reference (huggingface/pytorch-image-models, timm 1.0.28.dev0) is pure Python and has NO native
 * boundary: its hot path bottoms out in torch.nn.functional calls.  Each entry point below replaces
 * one such call site (cited as reference file:line, relative to the timm checkout); INTEGRATION.md
 * shows the ctypes binding + forward rebinding a timm maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator); the library
 *     never allocates, frees or synchronises; all work is enqueued on `stream` (a cudaStream_t).
 *   - return value: 0 (B200_OK) or an error code; b200_last_error() gives the message of the last
 *     failing call on the calling thread.  There is no CPU fallback: unsupported shapes are errors.
 *   - activations are bf16 unless noted; statistics / accumulators are fp32; parameters (bias,
 *     LayerNorm weight, LayerScale gamma) may be bf16 or fp32 (dtype enum).
 *   - re-entrant, stream-ordered, no global mutable state besides an immutable device-property cache.
 */
#ifndef B200TIMM_H_
#define B200TIMM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#define B200_OK 0
#define B200_ERR_INVALID_ARG 1
#define B200_ERR_UNSUPPORTED 2
#define B200_ERR_CUDA 3

#define B200_DTYPE_BF16 0
#define B200_DTYPE_F32 1
#define B200_DTYPE_U8 2    /* images / patches of the input pipeline only (b200_batch_patchify, b200_patches_u8) */

/* GEMM epilogues (applied to acc = A*B^T before the store). */
#define B200_EPI_BIAS 0       /* out = acc + bias                       (bias may be NULL)          */
#define B200_EPI_BIAS_GELU 1  /* out = gelu_erf(acc + bias)             timm/layers/mlp.py:48-49     */
#define B200_EPI_BIAS_RESIDUAL 2 /* out = aux + gamma * (acc + bias)    vision_transformer.py:212-213, layer_scale.py:27 */
#define B200_EPI_DGELU 3      /* out = (acc + bias) * gelu_erf'(aux)    backward of mlp.py:49        */
#define B200_EPI_BIAS_ROPE 4  /* out = rope(acc + bias) on the q / k columns of a fused QKV projection:
                                 attention.py:243-253 (AttentionRope), pos_embed_sincos.py:227-297 (interleaved
                                 `rot` layout, head_dim 64, one [sin | cos] table shared by batch and heads)       */

B200_API const char* b200_last_error(void);
/* 8 now (7 -> 8: head_dim 128 in b200_attn_fwd / b200_attn_bwd and their _bias variants, b200_attn_bwd_ws_floats_hd; b200_gemm_desc.ln_*, b200_gemm_ln_slots, b200_dwconv_nhwc_ln); 6 -> 7: b200_attn_fwd_bias / b200_attn_bwd_bias; 5 at the end of round 1 (4 -> 5: optimizer / loss / EMA / qk-norm / GLU / GRN / uint8 entry points, b200_gemm_desc.act). */
B200_API int b200_abi_version(void);
/* 0 if the current device is sm_100 and the tensor-map driver entry point resolves. */
B200_API int b200_device_check(void);

/* ------------------------------------------------------------------ GEMM (tcgen05 + TMA)
 * out[M,N] = epilogue( sum_k A[m,k] * B[n,k] )         bf16 operands, fp32 accumulation in TMEM.
 * Replaces F.linear at timm/layers/attention.py:115,137 and timm/layers/mlp.py:48,52 (forward,
 * a_mn=b_mn=0: A = activations [M,K], B = nn.Linear.weight [N,K]) and the autograd dgrad / wgrad
 * GEMMs of those calls (b_mn=1: B stored [K,N]; a_mn=1: A stored [K,M]).
 */
typedef struct b200_gemm_desc {
  const void* a;   /* bf16; a_mn=0: [M,K] row-major (lda = row stride in elements); a_mn=1: [K,M] */
  const void* b;   /* bf16; b_mn=0: [N,K] row-major (ldb);                           b_mn=1: [K,N] */
  void* out;       /* [M,N] row-major, ldo elements per row; bf16 or fp32 (out_dtype)             */
  const void* bias;  /* [N] or NULL; dtype bias_dtype                                             */
  const void* aux;   /* [M,N] bf16, ld = ldaux: residual (EPI_BIAS_RESIDUAL) / pre-activation (EPI_DGELU) */
  const void* gamma; /* [N] or NULL (LayerScale), dtype gamma_dtype; EPI_BIAS_RESIDUAL only        */
  void* out2;        /* optional second bf16 output [M,N] (ld = ldo2): the un-transformed acc + bias, saved for
                        backward -- the GELU pre-activation (EPI_BIAS_GELU) or the residual branch output before
                        LayerScale / DropPath (EPI_BIAS_RESIDUAL); `out` then applies the epilogue to those bf16 values */
  const float* row_scale; /* EPI_BIAS_RESIDUAL only, fp32 [ceil(M / rows_per_scale)] or NULL: per-sample DropPath factor
                        (keep-mask / keep_prob, timm/layers/drop.py:158-175): out = aux + row_scale*gamma*(acc+bias) */
  int64_t lda, ldb, ldo, ldaux, ldo2;
  int32_t M, N, K;
  int32_t a_mn, b_mn;
  int32_t epilogue;
  int32_t out_dtype;
  int32_t bias_dtype, gamma_dtype;
  int32_t accumulate; /* out += result instead of out = result (out_dtype f32 only)               */
  int32_t rows_per_scale; /* tokens per sample: row r uses row_scale[r / rows_per_scale]                       */
  int32_t split_k;    /* >1: K is split over CTAs, partial tiles are atomically added into fp32 out
                         (requires out_dtype=f32, epilogue=B200_EPI_BIAS with bias=NULL)           */
  /* EPI_BIAS_ROPE only (ABI >= 4).  rope: fp32 [rope_tokens - rope_prefix, 128] = [sin(64) | cos(64)], the `emb` of
   * apply_rot_embed_cat; row r of `out` is token r % rope_tokens of its sample; the first rope_prefix tokens
   * (cls / register tokens) and the columns >= rope_cols (the V third) pass through.  The value is rounded to bf16
   * before the rotation, as the reference rotates the bf16 output of nn.Linear in fp32 (`.type_as(v)` rounds again). */
  const float* rope;
  int32_t rope_tokens, rope_prefix, rope_cols;
  int32_t act;      /* activation applied by EPI_BIAS_GELU / differentiated by EPI_DGELU: 0 = exact (erf) GELU (nn.GELU()),
                       1 = tanh GELU (timm 'gelu_tanh': GELUTanh / nn.GELU(approximate='tanh'), activations.py), 2 = SiLU.
                       (this field was `reserved0` = 0 before ABI 5)                                                  */
  float* colsum;    /* ABI >= 6, EPI_DGELU only (bf16 output, M >= 256, N > 128), or NULL: fp32 [N] to which the column sums of
                       `out` are ADDED (red.global.add from the epilogue's staging tile) -- the bias gradient of the Linear
                       whose output gradient this GEMM produces (db1 = sum_rows dY1; the reference gets it from autograd of
                       F.linear, timm/layers/mlp.py:48).  The caller zero-initialises it.                             */
  /* ABI >= 8 -- LayerNorm folded into the GEMMs either side of it (inference): `x = x + ls(f(..))` followed by `norm(x)` and the
   * next Linear (timm/models/vision_transformer.py:212-213, timm/layers/norm.py:85-90) without a LayerNorm pass over HBM.
   * The statistics buffer is fp32 [S + 1][M][2] with S = b200_gemm_ln_slots(M, N) of the PRODUCER's output shape.
   * Producer (EPI_BIAS / EPI_BIAS_RESIDUAL, one bf16 output): ln_stats_out receives in slots 0 .. S-1, per row and slot (a
   * slot = the columns one epilogue warp stores), the sum and the sum of squares of the bf16 values written to `out`.
   * Consumer (EPI_BIAS / EPI_BIAS_GELU, bf16 output, a_mn = 0): `a` is the UN-normalised activation, `b` = weight * ln_weight
   * (column-wise, rounded to bf16), ln_stats the producer's buffer and ln_slots = S, ln_eps the norm's eps, ln_c1 = fp32 [N]
   * row sums of `b`, `bias` = fp32 [N] ln_bias . weight^T + the Linear's bias; the call first reduces the slots to
   * (rstd, -mean * rstd) per row in slot S of the buffer (mean / var over the K columns of the row, var = E[x^2] - mean^2 in
   * fp32 -- the one place the buffer is written through ln_stats), then the GEMM's epilogue evaluates
   * rstd * (acc - mean * ln_c1[n]) + bias[n].                                                                              */
  float* ln_stats_out;
  float* ln_stats;
  const float* ln_c1;
  int32_t ln_slots;
  float ln_eps;
} b200_gemm_desc;

B200_API int b200_gemm(const b200_gemm_desc* desc, void* stream);
/* Number S of partial-statistic slots a b200_gemm launch with ln_stats_out writes for an [M, N] output (ABI >= 8); the
 * buffer holds S + 1 slots (see b200_gemm_desc.ln_stats). */
B200_API int32_t b200_gemm_ln_slots(int32_t M, int32_t N);

/* In-place rotary embedding of the q and k thirds of a fused QKV activation qkv[B*N, 3*H*64] (bf16); replaces
 * apply_rot_embed_cat at timm/layers/attention.py:249-253 (pos_embed_sincos.py:227-297).  emb: fp32 [N - num_prefix, 128]
 * = [sin | cos].  half_layout: 0 = interleaved pairs (`rot`), 1 = `rope_rotate_half`.  inverse = 1 applies the
 * transposed map (the backward of the rotation, run on the q / k thirds of d(qkv)).
 * per_sample (ABI >= 6): 0 = one table for all samples; 1 = emb is [B, N - num_prefix, 128], sample b uses emb[b] (the
 * [B, 1, N, 2 * head_dim] rope of NaFlex / variable-grid batches, attention.py:243-253). */
B200_API int b200_rope_qk(void* qkv, const float* emb, int32_t B, int32_t N, int32_t H, int32_t num_prefix,
                          int32_t half_layout, int32_t inverse, int32_t per_sample, void* stream);

/* Debug hook (tools/gemm_trace.py, tools/attn_trace.py), not part of the reference-facing surface.  kernel 0 = b200_gemm:
 * with a non-NULL `device_buf` (>= 12 * #SMs uint64, zeroed by the caller) every following launch records per-CTA
 * %globaltimer stamps (entry, setup done, first operands landed, first / last accumulator ready, stores drained, exit)
 * and barrier-wait cycle totals.  kernel 1 = b200_attn_fwd (N <= 208 kernel): 16 int64 per CTA, cycles spent per softmax
 * phase (wait S, max pass, exp pass, wait O, epilogue) and per issuer wait.  NULL switches a trace off. */
B200_API void b200_set_trace(int32_t kernel, void* device_buf);

/* ------------------------------------------------------------------ LayerNorm   timm/layers/norm.py:85-90
 * y = (x - mean) * rstd * weight + bias over the last dim (cols).  x,y: [rows, cols]; x bf16 or f32,
 * y bf16 or f32; mean/rstd (fp32 [rows]) are written when non-NULL (saved for backward).          */
B200_API int b200_layernorm_fwd(const void* x, const void* weight, const void* bias, void* y, float* mean,
                       float* rstd, int64_t rows, int32_t cols, float eps, int32_t x_dtype,
                       int32_t y_dtype, int32_t param_dtype, void* stream);
/* dres (optional, dtype of x): gradient flowing through the residual connection around the norm; dx =
 * LN-backward(dy) + dres in one pass (Block.forward's `x + f(norm(x))`, vision_transformer.py:212-213).
 * dx (same dtype as x) and, accumulated over rows, dweight/dbias written as fp32 [cols] into
 * dw_partial/db_partial after an internal two-stage reduction (workspace: fp32 [ws_rows, 2*cols],
 * ws_rows = b200_layernorm_bwd_ws_rows(rows)).                                                       */
B200_API int b200_layernorm_bwd_ws_rows(int64_t rows, int32_t cols);
B200_API int b200_layernorm_bwd(const void* dy, const void* x, const void* weight, const float* mean,
                       const float* rstd, const void* dres, void* dx, float* dweight, float* dbias, float* workspace,
                       int64_t rows, int32_t cols, int32_t x_dtype, int32_t dy_dtype,
                       int32_t param_dtype, void* stream);

/* RMSNorm over the last dimension: y = x * rsqrt(mean(x^2) + eps) * weight.  Replaces rms_norm / RmsNorm.forward at
 * timm/layers/fast_norm.py:119-139, timm/layers/norm.py:202-248 (the LayerNorm kernels with the mean pinned to zero
 * and no bias).  rstd (fp32 [rows]) may be NULL in inference; the backward mirrors b200_layernorm_bwd (workspace of
 * b200_layernorm_bwd_ws_rows(rows, cols) * 2 * cols floats; dres = gradient of a residual branch around the norm). */
B200_API int b200_rmsnorm_fwd(const void* x, const void* weight, void* y, float* rstd, int64_t rows, int32_t cols, float eps,
                              int32_t x_dtype, int32_t y_dtype, int32_t param_dtype, void* stream);
B200_API int b200_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, const void* dres, void* dx,
                              float* dweight, float* workspace, int64_t rows, int32_t cols, int32_t x_dtype, int32_t dy_dtype,
                              int32_t param_dtype, void* stream);

/* ------------------------------------------------------------------ attention   timm/layers/attention.py:115-135
 * qkv: bf16 [B, N, 3, H, hd] (the QKV linear's output, viewed exactly as attention.py:115 reshapes it);
 * out: bf16 [B, N, H*hd] (== x.transpose(1,2).reshape(B,N,C) of attention.py:135);
 * lse: fp32 [B, H, N] natural-log-sum-exp of the scaled scores (NULL: not written);
 * kv_len: int32 [B] number of valid (prefix) keys per sample or NULL (all N) -- the NaFlex key-padding
 * mask of naflexvit.py:964-1029 given as lengths.  hd: 64, or 128 (ABI >= 8: any N in the forward, N <= 1120 in the
 * backward).  Other head widths (attention.py:85-93: `attn_head_dim` or dim // num_heads -- 32 of timm's test_vit models, 72
 * of SigLIP SO400M, 80 of ViT-H, 96) are served by zero-padding every head to the next supported width inside the qkv / proj
 * weights (b200timm/functional.py: head_pad), which leaves the scores and the output columns unchanged.                    */
B200_API int b200_attn_fwd(const void* qkv, void* out, float* lse, const int32_t* kv_len, int32_t B, int32_t N,
                  int32_t H, int32_t hd, float scale, void* stream);
/* Floats of workspace b200_attn_bwd needs behind its `delta` argument: B*H*N for the row sums delta[b,h,q] = sum_d
 * dO*O, plus -- for N > 208, where dQ is accumulated over key tiles in fp32 through TMA reduce-adds -- B*N*H*64. */
B200_API int64_t b200_attn_bwd_ws_floats(int32_t B, int32_t N, int32_t H);
/* The same for a head width hd (ABI >= 8): the fp32 dQ buffer B*N*H*hd is needed for N > 208 or hd = 128. */
B200_API int64_t b200_attn_bwd_ws_floats_hd(int32_t B, int32_t N, int32_t H, int32_t hd);

/* Backward: dqkv bf16 [B,N,3,H,hd] from dout bf16 [B,N,H*hd]; recomputes P from qkv and lse.
 * delta: fp32 workspace [B,H,N].  dbias (ABI >= 6): NULL, or a zero-initialised fp32 [3*H*hd] that receives the column sums
 * of dqkv over all B*N rows -- the gradient of the qkv projection's bias (autograd of F.linear at
 * timm/layers/attention.py:115), reduced from the tiles as they are stored; N <= 208 only (else B200_ERR_UNSUPPORTED). */
B200_API int b200_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                  void* dqkv, const int32_t* kv_len, int32_t B, int32_t N, int32_t H, int32_t hd,
                  float scale, float* dbias, void* stream);

/* Attention with an additive bias on the scaled scores (ABI >= 7): softmax(q k^T * scale + bias) v -- the `is_causal` and
 * general `attn_mask` cases of Attention.forward / AttentionRope.forward (timm/layers/attention.py:108-133, 255-277), whose
 * resolve_self_attn_mask (attention.py:21-40) turns a causal flag into a -inf upper triangle, a bool keep-mask into 0 / -inf
 * and passes float masks through.  bias: fp32 [batch or 1, heads or 1, N, bias_ld] (bias_ld >= N, a multiple of 4; 16-byte
 * aligned); bias_batch_stride / bias_head_stride in elements, 0 = the same bias for every sample / head.  -inf entries give
 * exactly zero probability; a row that is -inf everywhere yields NaN, as in the reference.  The bias receives no gradient.
 * All other arguments as b200_attn_fwd / b200_attn_bwd (key-padding lengths are expressed through the bias instead).     */
B200_API int b200_attn_fwd_bias(const void* qkv, void* out, float* lse, const float* bias, int64_t bias_batch_stride,
                  int64_t bias_head_stride, int32_t bias_ld, int32_t B, int32_t N, int32_t H, int32_t hd, float scale,
                  void* stream);
B200_API int b200_attn_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                  void* dqkv, const float* bias, int64_t bias_batch_stride, int64_t bias_head_stride, int32_t bias_ld,
                  int32_t B, int32_t N, int32_t H, int32_t hd, float scale, float* dbias, void* stream);

/* ------------------------------------------------------------------ patch embed   timm/layers/patch_embed.py:136-138
 * Gathers non-overlapping p x p patches of an NCHW bf16 image into rows of a [B*gh*gw, Kpad] bf16 matrix
 * (k = c*p*p + ky*p + kx, matching Conv2d.weight.flatten(1)); Kpad >= C*p*p, zero padded.              */
B200_API int b200_patchify(const void* img, void* patches, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                  int32_t Kpad, int32_t img_dtype, void* stream);
/* The input-pipeline tail folded into the patch gather: img is the uint8 NCHW batch the loader's
 * fast_collate produces; every pixel is normalised as PrefetchLoader does it (timm/data/loader.py:137-142:
 * .to(float32).sub_(mean * 255).div_(std * 255)) and rounded to bf16 once: patches = bf16((u8 - mean[c]) / std[c]).
 * mean / std: HOST arrays of C floats, already multiplied by 255 (C <= 8).                                           */
B200_API int b200_patchify_u8(const void* img, void* patches, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p,
                  int32_t Kpad, const float* mean, const float* std, void* stream);
/* pos_has_prefix=1: tokens[b, j, :] = (j < n_prefix ? prefix[j] : proj[b*G + j - n_prefix]) + pos[j]
 * pos_has_prefix=0: prefix tokens get no position, patch token i gets pos[i]  (no_embed_class=True)
 * (vision_transformer.py:1038-1075, _pos_embed).  prefix / pos: [n_prefix, D] / [.., D], dtype param_dtype. */
B200_API int b200_assemble_tokens(const void* proj, const void* prefix, const void* pos, void* tokens, int32_t B,
                         int32_t G, int32_t n_prefix, int32_t D, int32_t param_dtype,
                         int32_t pos_has_prefix, void* stream);

/* NaFlex token assembly (timm/models/naflexvit.py:857-960): tokens[b, n_prefix + i] = proj[b, i] + table[pos_off[b] + i]
 * for i < pos_len[b] (the sample's h*w grid, row-major; `table` holds the position embeddings already interpolated to
 * each distinct grid size, bf16 [T, D]) and proj[b, i] beyond it; tokens[b, j < n_prefix] = prefix[j].              */
B200_API int b200_naflex_assemble(const void* proj, const void* prefix, const void* table, const int32_t* pos_off,
                         const int32_t* pos_len, void* tokens, int32_t B, int32_t N, int32_t n_prefix,
                         int32_t D, int32_t prefix_dtype, void* stream);
/* Backward of b200_naflex_assemble (autograd of naflexvit.py:857-960): dproj (bf16 [B*N, D]) = the patch-token rows of
 * dtokens (bf16 [B, n_prefix+N, D]); dprefix (fp32 [n_prefix, D], may be NULL) += sum over b of the prefix rows;
 * dtable (fp32 [T, D], may be NULL) [pos_off[b] + i] += dtokens[b, n_prefix + i] for i < pos_len[b].  dprefix / dtable
 * are ACCUMULATED into (zero them first); the gradient of the position table through the bicubic resize is the caller's. */
B200_API int b200_naflex_assemble_bwd(const void* dtokens, void* dproj, float* dprefix, float* dtable,
                         const int32_t* pos_off, const int32_t* pos_len, int32_t B, int32_t N, int32_t n_prefix,
                         int32_t D, void* stream);

/* ------------------------------------------------------------------ q / k normalisation (row a4, qk_norm=True)
 * `q, k = self.q_norm(q), self.k_norm(k)` with norm_layer(head_dim) = LayerNorm or RmsNorm, timm/layers/attention.py:
 * 101-102,117.  qkv: bf16 [T, 3, H, 64] (T = B*N tokens), normalised IN PLACE on its q and k rows (v untouched).
 * Training also gives save (bf16 [T, 2, H, 64]: the raw q / k rows) and mean / rstd (fp32 [T * 2 * H], mean unused
 * for RMSNorm); inference passes NULL for all three.  Weights / biases: [64], param_dtype; biases NULL for RMSNorm.
 * Backward, in place on dqkv: d(q), d(k) w.r.t. the normalised rows become gradients w.r.t. the raw rows; dparams
 * (fp32 [4][64] = dq_weight, dq_bias, dk_weight, dk_bias) is ACCUMULATED into (zero it first).                     */
B200_API int b200_qk_norm_fwd(void* qkv, void* save, float* mean, float* rstd, const void* q_weight, const void* q_bias,
                      const void* k_weight, const void* k_bias, int32_t param_dtype, int64_t T, int32_t H,
                      int32_t head_dim, float eps, int32_t rms, void* stream);
B200_API int b200_qk_norm_bwd(void* dqkv, const void* save, const float* mean, const float* rstd, const void* q_weight,
                      const void* k_weight, int32_t param_dtype, float* dparams, int64_t T, int32_t H,
                      int32_t head_dim, int32_t rms, void* stream);

/* ------------------------------------------------------------------ Global Response Norm (ConvNeXt-V2)
 * GlobalResponseNorm inside GlobalResponseNormMlp, timm/layers/grn.py:17-49, mlp.py:251-295, on channels-last rows
 * h bf16 [B, HW, C] (C = hidden width): gx = ||h[b,:,c]||_2, nx = gx / (mean_c gx + eps), h' = h * (1 + weight * nx) + bias.
 * Forward, IN PLACE on h: gsq (fp32 [B, C], ZEROED by the caller) receives gx^2 and scale (fp32 [B, C]) the applied
 * 1 + weight * nx; both are kept for the backward.
 * Backward: dh2 = d loss / d h' (bf16 [B, HW, C]), pre = the fc1 pre-activation (h = GELU(pre) is recomputed);
 * dpre = (dh2 * scale + h * t) * GELU'(pre), i.e. the gradient w.r.t. the pre-activation (GRN and GELU backward in one
 * pass).  a_ws (fp32 [B, C], ZEROED) and t_ws (fp32 [B, C]) are workspaces; dweight (fp32 [C]) is ACCUMULATED into;
 * the bias gradient is the column sum of dh2 (b200_colsum).                                                          */
B200_API int b200_grn_fwd(void* h, float* gsq, float* scale, const void* weight, const void* bias, int32_t param_dtype,
                  int32_t B, int32_t HW, int32_t C, float eps, void* stream);
B200_API int b200_grn_bwd(const void* dh2, const void* pre, const float* gsq, const float* scale, const void* weight,
                  int32_t param_dtype, float* a_ws, float* t_ws, float* dweight, void* dpre, int32_t B, int32_t HW,
                  int32_t C, float eps, void* stream);

/* ------------------------------------------------------------------ gated MLPs
 * GluMlp / SwiGLUPacked / SwiGLU, timm/layers/mlp.py:57-165: h = fc1(x) is bf16 [M, 2H]; the gate is the first half of
 * each row (gate_first=1: SwiGLUPacked `act(x1) * x2`, and SwiGLU with [fc1_g; fc1_x] stacked) or the second
 * (gate_first=0: GluMlp's default `x1 * act(x2)`).  out[M, H] = act(gate) * value; act: 0 = SiLU, 1 = sigmoid.
 * Backward: dh[M, 2H] from dout[M, H] and the saved h.  H % 8 == 0.                                                  */
B200_API int b200_glu_fwd(const void* h, void* out, int64_t M, int32_t H, int32_t gate_first, int32_t act, void* stream);
B200_API int b200_glu_bwd(const void* h, const void* dout, void* dh, int64_t M, int32_t H, int32_t gate_first, int32_t act,
                  void* stream);

/* ------------------------------------------------------------------ ConvNeXt (channels-last activations [B,H,W,C] bf16)
 * Depthwise k x k convolution, stride 1, zero padding k/2, + bias: F.conv2d(groups=C) of timm/models/convnext.py:
 * 173-182,200 evaluated NHWC so the permutes of :205,208 disappear.  weight_t: the conv weight [C,1,k,k] transposed to
 * bf16 [k, k, C] (one 16-byte load per tap and 8 channels); k in {3,5,7}.
 * residual (optional, bf16 [B,H,W,C]) is added to the output: the backward uses the same kernel for
 * dx = dwconv(dy, spatially flipped weight) + shortcut gradient.                                                 */
B200_API int b200_dwconv_nhwc(const void* x, const void* weight_t, const void* bias, const void* residual, void* y,
                     int32_t B, int32_t H, int32_t W, int32_t C, int32_t ksize, int32_t bias_dtype, void* stream);
/* ABI >= 8.  The same convolution as the PRODUCER of a folded LayerNorm (b200_gemm_desc.ln_*): `y = conv_dw(x)` followed by
 * `norm(y)` and `fc1` of ConvNeXtBlock.forward (convnext.py:200-207) without a LayerNorm pass over HBM.  ln_stats_out: fp32
 * [C / 64 + 1][B*H*W][2]; slot c / 64 receives, per pixel, the sum and the sum of squares of the 64 bf16 channel values the
 * kernel stores; the consuming b200_gemm call gets ln_slots = C / 64.  Needs C % 64 == 0 and ksize == 7 (the TMA-staged
 * kernel), otherwise B200_ERR_UNSUPPORTED and the caller keeps the LayerNorm kernel.                                      */
B200_API int b200_dwconv_nhwc_ln(const void* x, const void* weight_t, const void* bias, void* y, float* ln_stats_out, int32_t B,
                        int32_t H, int32_t W, int32_t C, int32_t ksize, int32_t bias_dtype, void* stream);
/* dweight_t[ky,kx,c] += sum_{b,y,x} dy[b,y,x,c] * x[b,y+ky-k/2,x+kx-k/2,c]  (fp32 [k,k,C], zero-initialised by the caller). */
B200_API int b200_dwconv_nhwc_wgrad(const void* x, const void* dy, float* dweight_t, int32_t B, int32_t H, int32_t W,
                           int32_t C, int32_t ksize, void* stream);
/* inverse of b200_patchify_nhwc (backward of the downsample gather).                                              */
B200_API int b200_unpatchify_nhwc(const void* patches, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t p,
                         void* stream);
/* out[(b,oy,ox), (ky*p+kx)*C + c] = x[b, oy*p+ky, ox*p+kx, c]: the gather of the p x p / stride p downsample conv
 * (convnext.py:263-275); the GEMM then uses the conv weight permuted to [Cout, ky, kx, Cin].                   */
B200_API int b200_patchify_nhwc(const void* x, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t p,
                       void* stream);
/* out[b, c] = mean_hw x[b, hw, c]   (SelectAdaptivePool2d('avg'), timm/layers/classifier.py:211-213).            */
B200_API int b200_avgpool_nhwc(const void* x, void* out, int32_t B, int32_t HW, int32_t C, void* stream);

/* ------------------------------------------------------------------ elementwise / reductions
 * dst[i] (+)= src[i]: fp32 -> bf16/fp32 hand-off of split-K weight gradients into .grad (n % 4 == 0).        */
B200_API int b200_convert_f32(const float* src, void* dst, int64_t n, int32_t dst_dtype, int32_t accumulate,
                     void* stream);
/* out[r,c] = x[r,c] * col_scale[c] * row_scale[r / rows_per_scale] (bf16): gradient of a residual branch through
 * LayerScale (layer_scale.py:26-27) and DropPath (drop.py:158-175); either scale may be NULL.                  */
B200_API int b200_scale_rows_cols(const void* x, const void* col_scale, const float* row_scale, void* out, int64_t rows,
                         int32_t cols, int32_t rows_per_scale, int32_t col_scale_dtype, void* stream);
/* out[c] (+)= sum_r a[r,c] * b[r,c] * row_scale[r / rows_per_scale]: the LayerScale gamma gradient (fp32).        */
B200_API int b200_colsum_mul(const void* a, const void* b, const float* row_scale, float* out, int64_t rows,
                    int32_t cols, int32_t rows_per_scale, int32_t accumulate, void* stream);
/*
 * column sums of a bf16 [rows, cols] matrix into fp32 [cols] (bias gradients).                       */
B200_API int b200_colsum(const void* x, float* out, int64_t rows, int32_t cols, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------ optimizer step
 * The parameter update that follows the backward in timm's training loop: train.py:1262-1269 ->
 * torch.optim.SGD / torch.optim.AdamW / torch.optim.Adam as built by timm/optim/_optim_factory.py:355-404
 * ('sgd' = Nesterov, 'momentum', 'adamw', 'adam'), and the gradient clipping of timm/utils/clip_grad.py:18
 * (torch.nn.utils.clip_grad_norm_, 2-norm).  One call updates a whole parameter group in ONE launch.
 *
 * `tensors` is a HOST array (it travels by value in the kernel's parameter space, so nothing is staged through device
 * memory and gradient pointers may change every step); the pointers inside are device pointers.  param / grad have the
 * dtype given by `dtype` (bf16 or fp32), optimizer state is always fp32, arithmetic is fp32.  Hyper-parameters are
 * doubles because the reference's are Python floats: derived scalars (1 - beta, lr / (1 - beta1^t), ...) are formed in
 * double and rounded once, exactly as torch/optim does before handing them to its kernels.
 * grad_scale: optional device scalar multiplied into every gradient as it is read (the clip coefficient of
 * b200_clip_coef) -- the fused form of clip_grad_norm_ followed by step(); NULL = 1.                                  */
typedef struct b200_optim_tensor {
  void* param;        /* [numel] dtype; updated in place                                  */
  const void* grad;   /* [numel] dtype                                                     */
  float* state1;      /* SGD: momentum buffer (NULL iff momentum == 0); Adam: exp_avg      */
  float* state2;      /* Adam: exp_avg_sq; SGD: NULL                                       */
  int64_t numel;
} b200_optim_tensor;

/* torch/optim/sgd.py _single_tensor_sgd: g += weight_decay * p; buf = first_step ? g : momentum * buf + (1 - dampening) * g;
 * g = nesterov ? g + momentum * buf : buf; p -= lr * g.  first_step=1 for tensors whose momentum buffer has not been
 * initialised yet (torch clones the gradient into it).                                                                */
B200_API int b200_sgd_step(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, double lr, double momentum,
                         double dampening, double weight_decay, int32_t nesterov, int32_t first_step,
                         const float* grad_scale, void* stream);
/* torch/optim/adam.py _single_tensor_adam, step counted from 1.  decoupled=1 (AdamW): p *= 1 - lr * weight_decay;
 * decoupled=0 (Adam): g += weight_decay * p.  m = lerp(m, g, 1 - beta1); v = beta2 * v + (1 - beta2) * g * g;
 * p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps).                                                  */
B200_API int b200_adam_step(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, double lr, double beta1,
                         double beta2, double eps, double weight_decay, int32_t step, int32_t decoupled,
                         const float* grad_scale, void* stream);
/* *sumsq += sum over every gradient element of g^2 (fp32 device scalar; zero it first); only grad / numel are read. */
B200_API int b200_grad_sumsq(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, float* sumsq, void* stream);
/* torch/nn/utils/clip_grad.py: *total_norm = sqrt(*sumsq) (optional), *coef = min(1, max_norm / (total_norm + 1e-6)). */
B200_API int b200_clip_coef(const float* sumsq, double max_norm, float* coef, float* total_norm, void* stream);
/* g *= *scale in place over every gradient of the table (the unfused, in-place form of clip_grad_norm_). */
B200_API int b200_grad_scale(const b200_optim_tensor* tensors, int32_t n, int32_t dtype, const float* scale, void* stream);
/* Weight EMA, ModelEmaV3.apply_update_ (timm/utils/model_ema.py:213-231; called from train.py:1329 task.update_ema):
 * ema.lerp_(model, weight) with weight = 1 - decay for every tensor of the table, one launch.  Table entries:
 * param = the EMA tensor (ema_dtype, updated in place), grad = the model tensor (model_dtype); state1 / state2 unused.
 * The two dtypes may differ (an fp32 EMA of a bf16 model keeps updates of relative size 1e-4 that bf16 would drop).     */
B200_API int b200_ema_update(const b200_optim_tensor* tensors, int32_t n, int32_t ema_dtype, int32_t model_dtype,
                         double weight, void* stream);

/* ------------------------------------------------------------------ classification loss
 * nn.CrossEntropyLoss (train.py:855, benchmark.py:380; smoothing = 0), LabelSmoothingCrossEntropy
 * (timm/loss/cross_entropy.py:11-26: (1 - s) * nll + s * mean_c(-log_softmax)) with class indices `target` (int64 [B]),
 * or SoftTargetCrossEntropy (:29-36: sum_c(-t * log_softmax)) with a distribution `soft_target` [B, C] (bf16 / fp32);
 * exactly one of the two is non-NULL.  'mean' reduction over the B rows.  logits [B, C] bf16 or fp32, contiguous.
 * Forward writes loss_rows[B], lse[B] (log-sum-exp), tsum[B] (sum of the row's target weights), loss_mean[0] = the mean
 * and loss_mean[1] = its denominator (all fp32).  A row whose class index equals `ignore_index` (nn.CrossEntropyLoss's
 * argument; also any index outside [0, C)) contributes nothing and is not counted: the denominator is the number of
 * remaining rows, as in torch (all rows ignored -> NaN, as in torch); soft targets divide by B.
 * Backward: dlogits (dtype of logits) = (softmax * tsum - target distribution) * *grad_out / *denom.                  */
B200_API int b200_cross_entropy_fwd(const void* logits, int32_t logits_dtype, const int64_t* target, const void* soft_target,
                         int32_t soft_dtype, double smoothing, int64_t ignore_index, float* loss_rows, float* lse,
                         float* tsum, float* loss_mean, int32_t B, int32_t C, void* stream);
B200_API int b200_cross_entropy_bwd(const void* logits, int32_t logits_dtype, const int64_t* target, const void* soft_target,
                         int32_t soft_dtype, double smoothing, const float* lse, const float* tsum,
                         const float* grad_out, const float* denom, void* dlogits, int32_t B, int32_t C, void* stream);

/* ------------------------------------------------------------------ im2col-free patch embedding (ABI >= 6)
 * PatchEmbed.forward (timm/layers/patch_embed.py:117-142; the conv at :62,136 == a GEMM over non-overlapping patches) fused
 * with VisionTransformer._pos_embed's position add (vision_transformer.py:1038-1075), WITHOUT a patch matrix in memory:
 *     out[b, row_off + g, :] = patch(b, g) . weight^T + bias (+ pos[pos_off + g, :]),   g = py * (W / patch) + px
 * The GEMM's A operand is read straight from the bf16 NCHW image through a 5-D tensor map {kx %% 8, px, (b, c, py), kx / 8,
 * ky} (one TMA box = 64 / patch kernel rows of one channel for up to 128 patches, i.e. whole image rows from HBM, landing in
 * shared memory as the un-swizzled K-major core-matrix layout the MMA reads); weight is the
 * conv weight viewed [D, C * patch * patch] (bf16, K-major); pos: bf16 [pos_rows, D] or NULL; out: bf16 [B, out_tokens, D]
 * -- rows below row_off (class / register tokens) are left to the caller.  patch in {8, 16, 32, 64}; H, W multiples of it;
 * W / patch <= 128; D %% 8 == 0; otherwise B200_ERR_UNSUPPORTED (callers fall back to b200_patchify + b200_gemm).      */
B200_API int b200_patch_embed_fwd(const void* img, const void* weight, const void* bias, int32_t bias_dtype, const void* pos,
                                  int32_t pos_rows, int32_t pos_off, void* out, int32_t B, int32_t C, int32_t H, int32_t W,
                                  int32_t patch, int32_t D, int32_t out_tokens, int32_t row_off, void* stream);

/* ------------------------------------------------------------------ NaFlex input tail, ABI >= 6
 * b200_batch_patchify: batch_patchify (timm/models/naflexvit.py:159-201, pad=True) as the gather that builds the patch
 * projection's GEMM operand: NCHW image [B,C,H,W] (img_dtype uint8 / bf16 / f32) -> bf16 [B*ceil(H/p)*ceil(W/p), Kpad],
 * patches row-major over (y, x), flat layout P-P-C (channels_last = 1: k = (ky*p + kx)*C + c, the NaFlex default) or
 * C-P-P (0); pixels beyond H / W are zeros; columns K..Kpad are zero.  uint8 pixels are normalised
 * (v - mean[c]) / std[c] (mean / std: HOST arrays of C floats in 0..255 units, NaFlexPrefetchLoader's mean*255 / std*255,
 * timm/data/naflex_loader.py:75-80,161) in fp32 with one rounding to bf16; bf16 / f32 images are copied / cast.
 * b200_patches_u8: the same normalisation for patches the dataset already cut: uint8 [rows, K] -> bf16 [rows, Kpad]. */
B200_API int b200_batch_patchify(const void* img, int32_t img_dtype, void* patches, int32_t B, int32_t C, int32_t H, int32_t W,
                                 int32_t p, int32_t Kpad, int32_t channels_last, const float* mean, const float* std,
                                 void* stream);
B200_API int b200_patches_u8(const void* patches_u8, void* out, int64_t rows, int32_t K, int32_t Kpad, int32_t C,
                             int32_t channels_last, const float* mean, const float* std, void* stream);

/* dpre = dy * act'(pre) elementwise (bf16, n a multiple of 8): the activation backward of an Mlp with an inner norm between
 * the activation and fc2 (Mlp(norm_layer=...), timm/layers/mlp.py:44-53); act as in b200_gemm_desc.act.  ABI >= 6.      */
B200_API int b200_act_bwd(const void* dy, const void* pre, void* dpre, int64_t n, int32_t act, void* stream);

/* ------------------------------------------------------------------ data-parallel gradient exchange
 * The all-reduce(mean) of the parameter gradients that the reference gets from DistributedDataParallel
 * (timm/task/classification.py:64-65, train.py:1236-1269), fused into the kernel that finalises a block's gradients:
 * dst[i] += bf16(src[i] * scale) on EVERY rank of the node, through one multimem.red.add per 16 bytes on the NVLink
 * multicast address `dst_multicast` of the (symmetric, zero-initialised) flat bf16 gradient buffer.  src: fp32 (a block's
 * gradient workspace) or bf16, n elements (multiple of 8), 16-byte aligned; scale = 1 / world_size.  The caller orders
 * zeroing / pushes / reads with peer barriers (b200timm/grad_sync.py).  ABI >= 6.                                      */
B200_API int b200_grad_push(const void* src, int32_t src_dtype, void* dst_multicast, int64_t n, float scale, void* stream);
/* Replica consistency after the pushes: the adds land in arrival order and bf16 addition is not associative, so with more
 * than two ranks the copies may differ in the last bit.  Rank r re-publishes ITS slice of the buffer (`src_local`: the
 * slice in this rank's copy, `dst_multicast`: the same slice's multicast address, `bytes` a multiple of 16) to every rank
 * with multimem.st -- afterwards all copies are bit-identical, as after an all-reduce.                                  */
B200_API int b200_grad_publish(const void* src_local, void* dst_multicast, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200TIMM_H_ */

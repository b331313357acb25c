/* uvc_b200.h — C ABI of libuvc_sm100.so: the B200-native replacement for the device work on
 * UVC's data-parallel hot path.
 *
 * The reference (VITA-Group/UVC) has NO native interface: every op below replaces a span of
 * Python/ATen code, cited per entry point as `file:line` under /root/reference/UVC.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer (including workspaces); the library never allocates or frees
 *     device memory and never keeps a caller pointer after a call returns;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); there are no
 *     hidden host syncs, so every call can be captured into a CUDA graph;
 *   - return value: 0 on success, a negative uvc_status otherwise; uvc_last_error() gives the text;
 *     nothing throws or aborts across this boundary;
 *   - all matrices are fp32 row-major; tensor-core math is TF32 (fp32 storage, fp32 accumulate).
 */
#ifndef UVC_B200_H_
#define UVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVC_ABI_VERSION 9
#define UVC_MAX_DEPTH 32      /* most transformer blocks a uvc_vit_* call accepts */

#if defined(UVC_BUILD_DLL)
#define UVC_API __attribute__((visibility("default")))
#else
#define UVC_API
#endif

typedef enum {
  UVC_OK = 0,
  UVC_ERR_BAD_SHAPE = -1,    /* a dimension / stride / alignment the kernels do not support */
  UVC_ERR_BAD_ARG = -2,      /* null pointer, bad enum, inconsistent arguments */
  UVC_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed (text in uvc_last_error) */
  UVC_ERR_WORKSPACE = -4     /* workspace too small (see uvc_*_workspace_bytes) */
} uvc_status;

UVC_API int uvc_version(void);                 /* ABI version, bumped on any signature change */
UVC_API const char* uvc_last_error(void);      /* thread-local text of the last failure */
UVC_API int uvc_abi_sizeof(const char* struct_name);   /* sizeof() of an ABI struct, for binding self-checks */
UVC_API long long uvc_launch_count(void);      /* kernels this library has launched so far in this process */
/* Optional timing of every uvc GEMM launch with a CUDA-event pair on the launching stream (bench.py's roofline leg).
 * uvc_gemm_profile(1) clears the records and starts recording, (0) stops; _read synchronises the recorded events and
 * returns the summed kernel time, the summed algorithmic FLOPs (2*M*N*K*batch) and the number of launches. */
UVC_API int uvc_gemm_profile(int enable);
UVC_API int uvc_gemm_profile_read(double* total_ms, double* total_flops, long long* launches);
/* same, restricted to one kernel: kind 1 = gemm_tf32_kernel (128 x 128 tiles), 2 = gemm2_tf32_kernel (persistent CTA pairs), 0 = both */
UVC_API int uvc_gemm_profile_read_kind(int kind, double* total_ms, double* total_flops, long long* launches);

/* ------------------------------------------------------------------------------------------
 * Batched TF32 tensor-core GEMM (tcgen05.mma + TMEM accumulator + TMA operand staging)
 *   D[z] = epilogue( alpha * op(A[z]) . op(B[z])^T )        z = i2 * nb1 + i1
 * Replaces every nn.Linear / torch.matmul / autograd GEMM on the path:
 *   models/model_distilled.py:116,122 (fc1, fc2), :175 (qkv), :179,184 (QK^T, PV), :187 (proj),
 *   :149 (patch-embed conv as GEMM), :522 (head) and their autograd backward.
 *
 * Operand A is logically [M, K], operand B is logically [N, K].
 *   mn_major = 0: memory is [rows = M|N][cols = K], row stride ld           ("K-major")
 *   mn_major = 1: memory is [rows = K][cols = M|N], row stride ld           ("MN-major", i.e. transposed)
 * so forward (X.W^T), dgrad (dY.W) and wgrad (dY^T.X) all run without a transpose pass.
 * ld and batch strides are in elements; ld % 4 == 0 and ptr % 16 == 0 (TMA), batch strides % 4 == 0.
 */
typedef struct {
  const float* ptr;
  int64_t ld;
  int64_t bs1, bs2;     /* batch strides for i1, i2 (0 = operand shared across that batch index) */
  int32_t mn_major;
  int32_t _pad;
} uvc_operand;

enum {
  UVC_EPI_BIAS = 1,        /* v += bias[col] */
  UVC_EPI_GELU = 2,        /* aux[row,col] = gelu_erf'(v) (if aux != NULL: the derivative at the pre-activation, all the backward needs);
                              v = gelu_erf(v) */
  UVC_EPI_GELU_BWD = 4,    /* v *= aux[row,col]   (aux as written by UVC_EPI_GELU: the GELU backward is a multiply) */
  UVC_EPI_RESIDUAL = 8,    /* v += beta * R[row,col] */
  UVC_EPI_ATOMIC = 16,     /* D += v with red.global.add (required when splits > 1) */
  UVC_EPI_ROUND_TF32 = 32, /* round v to nearest TF32 before the store: for outputs that only feed other GEMMs (the
                              tensor core truncates its inputs, rounding here keeps the error unbiased) */
  UVC_EPI_COLSUM = 64,     /* colsum[col] += sum_rows v[row,col] (before TF32 rounding): the bias gradient of the Linear whose
                              output gradient this GEMM produces, fused so the tensor is not re-read (fp32 atomics) */
  UVC_GEMM_F16 = 128,      /* A and B hold fp16 values (ld / batch strides in fp16 elements, multiples of 8): tcgen05.mma kind::f16 with fp32
                              accumulation -- the 10 mantissa bits of TF32 at half the operand bytes.  K-major unbatched operands run on the
                              CTA-pair kernel; MN-major operands, split-K and batches on the 128 x 128 kernel (fp32 output only there). */
  UVC_EPI_BLEND = 512,     /* block gate blend fused behind the residual (models/model_distilled.py:493): with t = the value formed so far
                              (alpha acc + bias + beta R), D2 = t (optional) and v = blend_dev[1] * t + blend_dev[0] * R2.  CTA-pair kernel, plain epilogue. */
  UVC_EPI_AUX_F16 = 256    /* `aux` points to fp16 values (ldaux in fp16 elements, ldaux % 4 == 0, 8 B-aligned): the gelu' factors are in
                              [-0.13, 1.13] and the product they enter is rounded to TF32 anyway, so 11 significant bits lose nothing and
                              the GELU pair of epilogues moves half the aux bytes.  CTA-pair kernel only (unbatched operands). */
};

typedef struct {
  int32_t M, N, K;
  int32_t nb1, nb2;        /* batch counts (>= 1) */
  int32_t splits;          /* split-K factor (>= 1); > 1 requires UVC_EPI_ATOMIC */
  uvc_operand A, B;
  float* D; int64_t ldd, d_bs1, d_bs2;
  const float* bias;       /* [N] */
  const float* R; int64_t ldr, r_bs1, r_bs2;
  float* aux; int64_t ldaux, aux_bs1, aux_bs2;
  float alpha;             /* scales the accumulator before bias */
  float beta;              /* scales R */
  const float* alpha_dev;  /* optional device scalar multiplied into alpha (gate values) */
  const float* beta_dev;   /* optional device scalar multiplied into beta */
  int32_t flags;           /* UVC_EPI_* */
  float colsum_scale;      /* multiplies the UVC_EPI_COLSUM column sums before they are accumulated; 0 means 1 (how the loss scale of the
                              fp16 gradient tensors is taken back out of a fused bias gradient) */
  float* colsum;           /* [N], accumulated into when UVC_EPI_COLSUM is set */
  const float* R2; int64_t ldr2;   /* UVC_EPI_BLEND: the block input x */
  float* D2; int64_t ldd2;         /* UVC_EPI_BLEND: where the un-blended t goes (the backward's gate gradient needs it), or NULL */
  const float* blend_dev;          /* UVC_EPI_BLEND: device (d0, d1) */
  const float* alpha_dev2; /* a second optional device scalar multiplied into alpha (gate value x inverse loss scale) */
  const float* colsum_scale_dev;  /* optional device scalar multiplied into colsum_scale */
  void* D16; int64_t ldd16; /* optional fp16 copy of the output [M, ldd16] (round to nearest), for consumers that read it as an fp16 GEMM
                              operand; D may then be NULL.  Needs the unbatched CTA-pair kernel (N % 4 == 0, 16 B-aligned rows). */
} uvc_gemm_args;

UVC_API int uvc_gemm_tf32(const uvc_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-wise bandwidth-bound kernels (warp per row, float4 accesses).  `ld*` are row strides in
 * elements (multiples of 4).  Every function is asynchronous on `stream`.
 * `round_tf32` != 0 rounds the output to the nearest TF32 value (use it when the output only feeds tensor-core GEMMs).
 */

/* y = (x - mean) * rstd * gamma + beta per row; mean/rstd (optional) are saved for the backward.
 * Replaces nn.LayerNorm at models/model_distilled.py:199,204,288,507 (eps 1e-6) and
 * T2TViT/models/transformer_block.py:84,88 (eps 1e-5). */
UVC_API int uvc_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                              float* y, int64_t ldy, float* mean, float* rstd, int32_t M, int32_t C, int32_t round_tf32, void* stream);
/* dx = r1 + s2 * r2 + LN'(dy); dgamma += sum dy*xhat; dbeta += sum dy.  r1, r2 (same ld as dx), s2_dev,
 * dgamma/dbeta may be NULL.  The two residual inputs fuse the skip connection and the block-gate blend. */
UVC_API int uvc_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                              const float* gamma, const float* r1, const float* r2, const float* s2_dev,
                              float* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M, int32_t C, void* stream);
/* Same, and additionally cs_r1[col] += sum_rows (r1 + s2*r2)[row,col], cs_out[col] += sum_rows dx[row,col] (either may be NULL): the bias gradients of
 * the Linear layers on either side of the norm are column sums of tensors this kernel streams anyway (fc2.bias <- r1 = d(block output),
 * attn.proj.bias <- dx = d(x1) for norm2 of models/model_distilled.py:204,243). */
UVC_API int uvc_layernorm_bwd_cs(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                                 const float* gamma, const float* r1, const float* r2, const float* s2_dev,
                                 float* dx, int64_t lddx, float* dgamma, float* dbeta, float* cs_r1, float* cs_out, int32_t M, int32_t C, void* stream);
/* 16-bit operand storage variants (the engine's default precision mode, see uvc_vit_dims.operand_f16):
 *   uvc_layernorm_fwd_f16: y16 is fp16 [M, ldy] -- the A operand of the kind::f16 GEMM that follows the norm.
 *   uvc_layernorm_bwd_f16: dy16 is fp16 and carries the backward's loss scale (dy = dy_scale * dy16); dx (fp32, unscaled) as above and,
 *                          when dx16 != NULL, dx16 = fp16(dx16_scale * dx): the operand copy of the stream gradient for the next GEMMs.
 *   uvc_cvt_f16:           dst16 [rows, cols] = fp16(src) and / or dstT16 [cols, rows] = fp16(src^T) (weights: forward and dgrad operands). */
UVC_API int uvc_layernorm_fwd_f16(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y16, int64_t ldy,
                                  float* mean, float* rstd, int32_t M, int32_t C, void* stream);
UVC_API int uvc_layernorm_bwd_f16(const void* dy16, int64_t lddy, float dy_scale, const float* x, int64_t ldx, const float* mean, const float* rstd,
                                  const float* gamma, const float* r1, const float* r2, const float* s2_dev, float* dx, void* dx16, float dx16_scale,
                                  int64_t lddx, float* dgamma, float* dbeta, float* cs_r1, float* cs_out, int32_t M, int32_t C, void* stream);
UVC_API int uvc_cvt_f16(const float* src, void* dst16, void* dstT16, int32_t rows, int32_t cols, void* stream);
/* in-place row softmax over the first n (<= 256) columns of S[rows][ld] (models/model_distilled.py:180) */
UVC_API int uvc_softmax_fwd(float* S, int64_t ld, int64_t rows, int32_t n, int32_t round_tf32, void* stream);
/* dP <- scale * P .* (dP - rowsum(dP .* P)) */
UVC_API int uvc_softmax_bwd(const float* P, float* dP, int64_t ld, int64_t rows, int32_t n, float scale, int32_t round_tf32, void* stream);
/* out[col] += scale * sum_rows X[row, col]  (bias gradients); scale_dev may be NULL (= 1) */
UVC_API int uvc_colsum(const float* X, int64_t ld, int32_t M, int32_t N, const float* scale_dev, float* out, void* stream);
/* out = d[1] * t + d[0] * x   -- block gate blend, models/model_distilled.py:493 */
UVC_API int uvc_blend_fwd(const float* t, const float* x, const float* d, float* out, int64_t n, void* stream);
/* dots[0] += <g, x>, dots[1] += <g, t>  -- gradient of the blend weights */
UVC_API int uvc_blend_dots(const float* g, const float* t, const float* x, float* dots, int64_t n, void* stream);
/* im2col of the patch x patch / stride patch conv: out[(b, py, px), (c, ky, kx)]  (models/model_distilled.py:149) */
UVC_API int uvc_im2col16(const float* x, float* out, int32_t B, int32_t Cin, int32_t HW, int32_t P, int32_t round_tf32, void* stream);
/* tok[b,0,:] = cls + pos[0]; tok[b,1+p,:] = pe[b,p,:] * pscale[p] * tmask[b,p] + pos[1+p]
 * (models/model_distilled.py:434-471; pscale / tmask may be NULL) */
UVC_API int uvc_assemble_tokens(const float* pe, const float* cls, const float* pos, const float* pscale, const float* tmask,
                                float* tok, int32_t B, int32_t np, int32_t C, void* stream);
/* backward of uvc_assemble_tokens: dpe = g * scale; dscale[p] += ..., dtmask[b,p] = ..., dpos += sum_b g, dcls += sum_b g[:,0] */
UVC_API int uvc_assemble_tokens_bwd(const float* g, const float* pe, const float* pscale, const float* tmask, float* dpe,
                                    float* dscale, float* dtmask, float* dpos, float* dcls, int32_t B, int32_t np, int32_t C, void* stream);
/* Token (patch) slimming gate, mode 2 of --enable_patch_gating (models/model_distilled.py:446-456, gumbel_softmax :36-63, scatter :21-33),
 * one CTA per image, no host round trip:
 *   scores = pscale * (feat . v + c1) + gate_b ; l = log_softmax(scores) ; y = softmax((l + noise) / tau) ;
 *   hard = the k largest y (exact rank, ties to the lower index = torch.topk on CUDA) ; mask = (hard - y) + y ; mask[:, 0] = 1.
 * The score is an fp32 dot product over the rows the patch GEMM reads (feat = im2col rows [B*np, Kf] with v, c1 from uvc_token_gate_fold =
 * W_patch^T w_gate, b_patch . w_gate), or over the token embeddings themselves (T2T: feat = tokens, v = w_gate, c1 = NULL), so the kept-token
 * indices do not depend on tensor-core rounding.  noise [B, np] is the caller's Gumbel sample (torch's generator: the reference's stream).
 * ysoft / ls / scores (each [B, np], optional) are what the backward needs.
 *   uvc_token_gate_bwd:   dscores from dmask (straight-through: d mask / d y = 1, none through column 0).
 *   uvc_token_gate_apply: dx += dscores pscale (x) w_gate ; d_gate_w += sum dscores pscale x ; d_gate_b += sum dscores ; d_pscale[p] += ... */
UVC_API int uvc_token_gate_fold(const float* patch_w, const float* patch_b, const float* gate_w, int32_t C, int32_t Kp, float* v, float* c1, void* stream);
UVC_API int uvc_token_gate_fwd(const float* feat, int64_t ldf, int32_t Kf, const float* v, const float* c1, const float* gate_b, const float* pscale,
                               const float* noise, float tau, int32_t k, int32_t B, int32_t np, float* mask, float* ysoft, float* ls, float* scores,
                               void* stream);
UVC_API int uvc_token_gate_bwd(const float* dmask, const float* ysoft, const float* ls, float tau, int32_t B, int32_t np, float* dscores, void* stream);
UVC_API int uvc_token_gate_apply(const float* dscores, const float* x, const float* gate_w, const float* pscale, int32_t B, int32_t np, int32_t C,
                                 float* dx, float* d_gate_w, float* d_gate_b, float* d_pscale, void* stream);
/* dst[i] = round_to_nearest_tf32(src[i])  (weights are rounded once per forward into the workspace) */
UVC_API int uvc_round_tf32(const float* src, float* dst, int64_t n, void* stream);
/* y += s * (*s_dev) * x */
UVC_API int uvc_scale_add(float* y, const float* x, const float* s_dev, float s, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention core  softmax(Q K^T * scale) V  per (image, head)   (models/model_distilled.py:175-185)
 * qkv: [B*N, 3*H*d] exactly as nn.Linear(dim, 3*dim) writes it (q | k | v, head-major inside each).
 * P:   [B, H, N, ldp] attention probabilities, saved for the backward (ldp = uvc_attn_ldp(N)).  May be NULL when d == 64 and
 *      N <= 208 (inference / teacher forward): the fused kernel keeps scores and probabilities in tensor memory and only qkv -> ctx
 *      touches HBM.  ctx: [B*N, H*d] (already "transposed back", ready for the proj GEMM).
 */
UVC_API int32_t uvc_attn_ldp(int32_t N);
UVC_API int uvc_attention_fwd(const float* qkv, float* P, float* ctx, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream);
/* Training forward of the fused path (d == 64, N <= 208): instead of the probabilities it saves lse[B,H,N], the per-row log-sum-exp of the
 * scaled scores in the log2 domain (P = exp2(S * scale * log2(e) - lse)), 4 bytes per row instead of an 800-byte row of P. */
UVC_API int uvc_attention_fwd_lse(const float* qkv, float* lse, float* ctx, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream);
/* Fused backward with recomputation (autograd backward of models/model_distilled.py:175-185): dqkv [B*N, 3*H*d] from dctx, the forward's
 * ctx and lse; D_ws is [B,H,N] scratch (rowsum(dctx .* ctx)).  Scores, probabilities and dS live only in tensor memory.
 * dqkv_bias (optional, [3*H*d]) accumulates the column sums of dqkv, i.e. the gradient of attn.qkv.bias, in the same pass. */
UVC_API int uvc_attention_bwd_fused(const float* qkv, const float* lse, const float* ctx, const float* dctx, float* D_ws, float* dqkv,
                                    float* dqkv_bias, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream);
/* The same fused pair with fp16 operand storage (d == 64, N <= 208): qkv16 / ctx16 / dctx16 / dqkv16 hold fp16 values in the fp32 path's
 * layouts; kind::f16 MMAs, fp32 accumulation and softmax; Q / K / V double-buffered in shared memory; one staged copy of each operand
 * serves the score (K-major) and the output (MN-major) MMAs.  dctx16 may carry a loss scale (the backward is linear in it): dqkv16
 * then carries it too and dqkv_bias receives db_scale * column sums (pass 1 / scale). */
UVC_API int uvc_attention_fwd_f16(const void* qkv16, void* ctx16, float* lse, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream);
UVC_API int uvc_attention_bwd_f16(const void* qkv16, const float* lse, const void* ctx16, const void* dctx16, float* D_ws, void* dqkv16,
                                  float* dqkv_bias, float db_scale, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream);
/* dqkv [B*N, 3*H*d] from dctx [B*N, H*d]; dP is scratch of the same size as P */
UVC_API int uvc_attention_bwd(const float* qkv, const float* P, const float* dctx, float* dP, float* dqkv,
                              int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Loss: soft-target cross entropy + soft distillation KL, forward and d(loss)/d(logits) in one pass
 * (utils/losses.py:38-64 with timm SoftTargetCrossEntropy as base criterion, outputs_kd is outputs):
 *   loss = (1-alpha) * mean_b sum_c -y log_softmax(s) + alpha * T^2 / (B*NC) * sum exp(lt) (lt - ls),
 *   ls = log_softmax(s/T), lt = log_softmax(t/T).   t may be NULL (distillation 'none', alpha ignored).
 * loss_out[0] = loss, loss_out[1] = base, loss_out[2] = kd (zeroed by the call).  dlogits is scaled by grad_scale.
 */
UVC_API int uvc_distill_loss(const float* logits, const float* teacher_logits, const float* targets, int32_t B, int32_t NC,
                             float alpha, float T, float grad_scale, float* loss_out, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input pipeline on the device: timm batch-mode Mixup / CutMix + smoothed mixed one-hot targets in ONE launch
 * (joint_train.py:409,930-933; post_train.py:362,618-621).  lam and the CutMix box come from the host RNG as in timm.
 *   mixup : x[i] <- lam x[i] + (1-lam) x[B-1-i] in place;  cutmix: the box [yl,yh) x [xl,xh) of x[i] <- that of x[B-1-i] in place;
 *   targets[i,c] = lam * sm(y[i])[c] + (1-lam) * sm(y[B-1-i])[c],  sm = smoothing/NC + (1-smoothing) [c == y].   lam == 1: images untouched.
 * x: [B, C, H, W] fp32 (B even), y: [B] int64, targets: [B, num_classes] fp32 (may be NULL).  Parity unpinned by the reference (timm is not in
 * its tree): pinned against the plain-torch formula in tests/. */
UVC_API int uvc_mixup(float* x, const int64_t* y, float* targets, int32_t B, int32_t C, int32_t H, int32_t W, int32_t num_classes, float lam,
                      float smoothing, int32_t use_cutmix, int32_t yl, int32_t yh, int32_t xl, int32_t xh, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser: global-norm clip + AdamW   (joint_train.py:428-429, torch.optim.AdamW semantics)
 *   uvc_sqnorm_accum : acc[0] += sum g^2                       (call once per gradient buffer)
 *   uvc_clip_adamw   : coef = min(1, max_norm / (sqrt(acc[0]) + 1e-6)) (1 if max_norm <= 0); g *= coef;
 *                      p *= 1 - lr*wd; m,v updated; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
 * `step` is the 1-based step count; hyper-parameters are host scalars.  mask (optional, same size as p)
 * multiplies the update so masked weights stay exactly 0 (Stage 2, post_train.py:357-360).
 */
UVC_API int uvc_sqnorm_accum(const float* g, int64_t n, float* acc, void* stream);
/* Flat-arena variants with one option byte per element (arenas 16 B aligned, n % 4 == 0), so the WHOLE model stays one launch each also in
 * Stage 2 (post_train.py:357-379: re-masked weights, timm's decay / no-decay parameter groups) and with tenants that have no gradient:
 *   bit 0 keep   (0: pruned weight -- updated like the reference's AdamW does, then forced back to exactly 0 = the next step's `weight *= mask`),
 *   bit 1 decay  (decoupled weight decay applies), bit 2 active (0: no gradient this step -- frozen parameter or hard-skipped block: untouched
 *   and excluded from the clip norm, as a `.grad is None` parameter is in torch). */
UVC_API int uvc_sqnorm_accum_flags(const float* g, const uint8_t* flags, int64_t n, float* acc, void* stream);
UVC_API int uvc_clip_adamw_flags(float* p, float* g, float* m, float* v, const uint8_t* flags, int64_t n, const float* sqnorm_acc, float max_norm,
                                 float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);
UVC_API int uvc_clip_adamw(float* p, float* g, float* m, float* v, const float* mask, int64_t n, const float* sqnorm_acc, float max_norm,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-model engine: DistilledVisionTransformer forward / backward as one call each
 * (models/model_distilled.py:429-531).  Tensors are passed by pointer; `blocks` is a HOST array.
 * The same struct shape is used for parameters and for gradients.
 */
typedef struct {
  float *norm1_w, *norm1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *norm2_w, *norm2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} uvc_block_tensors;             /* qkv_b may be NULL (qkv_bias=False, T2TViT/models/transformer_block.py:50) */

typedef struct {
  float *patch_w, *patch_b;      /* [C, in_chans*patch*patch], [C] */
  float *cls_token, *pos_embed;  /* [C], [ntok, C] */
  float *norm_w, *norm_b;        /* final LayerNorm */
  float *head_w, *head_b;        /* [num_classes, C], [num_classes] */
  const uvc_block_tensors* blocks;   /* host array, L entries */
} uvc_vit_tensors;

typedef struct {
  int32_t B, img, patch, in_chans;
  int32_t C, H, Fh, L, num_classes;
  float ln_eps;
  int32_t operand_f16;           /* 0: fp32 operand storage, tcgen05 kind::tf32 everywhere (round 1's path).
                                    1: every tensor that only feeds GEMMs (LayerNorm outputs, qkv, attention context, gelu(fc1), the operand
                                       copies of the gradients) is stored as fp16 -- the same 10 mantissa bits -- and the block GEMMs / attention run
                                       kind::f16 with fp32 accumulation; residual stream, statistics, softmax, logits, loss and every parameter
                                       gradient stay fp32.  Needs head dim 64, <= 208 tokens, C % 8 == 0, Fh % 8 == 0.
                                    The workspace layout depends on it: use the same value for workspace_bytes / forward / backward. */
} uvc_vit_dims;

/* Stage-2 physical compaction for TRAINING.  The reference re-masks dense weights before every step (post_train.py:357-360) and
 * multiplies by the zeros; with a layout the engine gathers, per block, the live heads (64-wide q / k / v row groups of attn.qkv.weight and
 * the matching 64 input columns of attn.proj.weight) and the live neurons (rows of mlp.fc1.weight, columns of mlp.fc2.weight) into compact
 * fp16 operand copies while it converts the weights, runs every GEMM / attention kernel of the block at the compact widths
 * (3*64*n_heads, 64*n_heads, n_neurons) and scatters the compact weight / bias gradients back into the dense gradient tensors.
 * Entries outside the live lists receive no gradient ("clip norm under compaction").  operand_f16 only.
 *   head_idx   device [L, H]  int32: row l lists the n_heads[l]   live head numbers   (ascending; the rest of the row is ignored)
 *   neuron_idx device [L, Fh] int32: row l lists the n_neurons[l] live neuron numbers (ascending)
 * 1 <= n_heads[l] <= H;  64 <= n_neurons[l] <= Fh and n_neurons[l] % 64 == 0: a caller tops the lists up with pruned entries -- their masked
 * weights are zero, so they change nothing.  Hard-skipped blocks are not looked at. */
typedef struct {
  int32_t n_heads[UVC_MAX_DEPTH];
  int32_t n_neurons[UVC_MAX_DEPTH];
  const int32_t* head_idx;
  const int32_t* neuron_idx;
} uvc_vit_layout;

typedef struct {
  uvc_vit_dims dims;
  uvc_vit_tensors w;
  const float* x;                /* [B, in_chans, img, img] */
  const float* blend;            /* device [L,2] = (d0, d1) per block: x <- d1*blk(x) + d0*x; NULL = plain blocks */
  const uint8_t* skip_host;      /* host [L] or NULL; nonzero = block is not executed (hard skip, :496-500) */
  const float* patch_scale;      /* device [np] or NULL (patch gate mode 1, :434-444) */
  const float* token_mask;       /* device [B, np] or NULL (token gate, :446-456) */
  int32_t save_for_backward;     /* 1: keep activations in the workspace for uvc_vit_backward */
  int32_t enable_jumping;        /* final norm sees the sum of all block outputs (:503-506) */
  float* logits;                 /* [B, num_classes] */
  float* pe_out;                 /* optional [B*np, C]: raw patch embeddings (before gates), may be NULL */
  void* workspace; uint64_t workspace_bytes;
  const float* pe_in;            /* optional [B*np, C]: token embeddings computed by the caller; replaces x -> im2col -> patch GEMM.
                                    This is how the T2T-ViT backbone (T2TViT/models/t2t_vit.py:168-208: cls + sinusoid pos-embed, 14 Blocks,
                                    norm, head) runs behind the same entry point, fed by tokens_to_token (:46-105).  x, w.patch_* may be NULL. */
  const uvc_vit_layout* layout;  /* host struct or NULL (dense): Stage-2 compaction, see uvc_vit_layout */
  int32_t weights_converted;     /* 1: the workspace already holds the operand copies (fp16 / TF32-rounded) of exactly these weights from an earlier call
                                    with the same workspace -- the caller's promise for a FROZEN model (the distillation teacher, utils/losses.py:47-49);
                                    the per-forward weight conversion launch is skipped.  0: convert (always correct). */
} uvc_vit_forward_args;

typedef struct {
  uvc_vit_dims dims;
  uvc_vit_tensors w;
  uvc_vit_tensors g;             /* gradients, ACCUMULATED into (caller zeroes) */
  const float* dlogits;          /* [B, num_classes] */
  const float* blend;
  const uint8_t* skip_host;
  const float* patch_scale;
  const float* token_mask;
  int32_t enable_jumping;
  float grad_scale;              /* operand_f16 only: the power-of-two loss scale S the fp16 gradient operands carry (g16 = fp16(S g)); it is taken
                                    back out wherever an fp32 result is produced, so every output of this call is the true gradient.
                                    <= 0 (default): chosen on the device per call, the largest power of two with S * max|dlogits| <= 128 --
                                    every gradient is linear in dlogits, so the fp16 range use does not depend on how the loss was scaled. */
  float* d_blend;                /* [L,2], accumulated; NULL if blend == NULL */
  float* d_patch_scale;          /* [np] accumulated, or NULL */
  float* d_token_mask;           /* [B, np] written, or NULL */
  void* workspace; uint64_t workspace_bytes;   /* the workspace the forward ran with */
  float* d_pe;                   /* forward ran with pe_in: gradient w.r.t. pe_in, [B*np, C] WRITTEN (g.patch_* untouched); else NULL */
  const uvc_vit_layout* layout;  /* the layout the forward ran with, or NULL */
} uvc_vit_backward_args;

UVC_API uint64_t uvc_vit_workspace_bytes(const uvc_vit_dims* dims, int32_t save_for_backward);
UVC_API int uvc_vit_forward(const uvc_vit_forward_args* args, void* stream);
UVC_API int uvc_vit_backward(const uvc_vit_backward_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tokens-to-token front end of T2T-ViT (T2TViT/models/t2t_vit.py:46-105 with tokens_type = 'performer', token_performer.py:8-69):
 *   x [B, in_chans, img, img] -> Unfold 7x7/4/2 -> Token_performer(in_chans*49 -> 64) -> Unfold 3x3/2/1 -> Token_performer(576 -> 64)
 *     -> Unfold 3x3/2/1 -> Linear(576, C) -> tokens [B * (img/16)^2, C]      (token_dim 64, 32 random features, as T2T_module builds them).
 * One call per pass: soft split + LayerNorm fused into one gather kernel (the unfolded tensor is never materialised), tcgen05 GEMMs with fp16
 * operands for kqv / proj / mlp / project, the linear attention on CUDA cores with the random features recomputed in the backward.
 * tokens feed uvc_vit_forward as `pe_in`; its `d_pe` comes back here as d_tokens.  Gradients are ACCUMULATED into g (caller zeroes); `w` of a
 * Token_performer (the fixed random features) has no gradient.  dropout_p > 0 applies the reference's nn.Dropout(0.1) sites (training mode) with a
 * stateless hash of (seed, element) -- use the same seed in the forward and the backward call; the random stream is not torch's.
 */
typedef struct {
  const float *norm1_w, *norm1_b;        /* [dim] */
  const float *kqv_w, *kqv_b;            /* [192, dim], [192]  (k | q | v) */
  const float *w;                        /* [32, 64] random features */
  const float *proj_w, *proj_b;          /* [64, 64], [64] */
  const float *norm2_w, *norm2_b;        /* [64] */
  const float *mlp0_w, *mlp0_b, *mlp2_w, *mlp2_b;   /* [64, 64], [64] */
} uvc_performer_tensors;
typedef struct {
  uvc_performer_tensors attn1, attn2;    /* tokens_to_token.attention1 / attention2 */
  const float *project_w, *project_b;    /* [C, 576], [C] */
} uvc_t2t_tensors;
typedef struct {
  int32_t B, img, in_chans, C;
  float ln_eps;                          /* nn.LayerNorm default 1e-5 */
} uvc_t2t_dims;
typedef struct {
  uvc_t2t_dims dims;
  uvc_t2t_tensors w;
  const float* x;                        /* [B, in_chans, img, img] */
  float* tokens;                         /* [B * (img/16)^2, C] written */
  int32_t save_for_backward;
  float dropout_p; uint64_t seed;
  void* workspace; uint64_t workspace_bytes;
} uvc_t2t_forward_args;
typedef struct {
  uvc_t2t_dims dims;
  uvc_t2t_tensors w;
  uvc_t2t_tensors g;                     /* gradients, accumulated into (the `w` slots are ignored); pointers are written through */
  const float* x;
  const float* d_tokens;                 /* [B * (img/16)^2, C] */
  float dropout_p; uint64_t seed;
  float grad_scale;                      /* loss scale of the fp16 gradient operands; <= 0: picked on the device from max|d_tokens| */
  void* workspace; uint64_t workspace_bytes;     /* the workspace the forward ran with (save_for_backward = 1) */
} uvc_t2t_backward_args;
UVC_API uint64_t uvc_t2t_workspace_bytes(const uvc_t2t_dims* dims, int32_t save_for_backward);
UVC_API int uvc_t2t_forward(const uvc_t2t_forward_args* args, void* stream);
UVC_API int uvc_t2t_backward(const uvc_t2t_backward_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * ADMM primal-dual update (uvc_optimizer.py:37-144 + uvc_utils.py:54-73,177-269,315-471) on the device.
 * One argument block for all six entry points; each reads the fields it needs.
 *   W1[l] = blocks[l].attn.proj.weight [C, C] (prunable input columns, grouped in H heads of d),
 *   W3[l] = blocks[l].mlp.fc2.weight  [C, Fh] (prunable input columns), W2[l] = mlp.fc1.weight [Fh, C] (rows follow W3).
 * Call order of one step (uvc_optimizer):  scores -> prox -> scores -> primal -> [host: gate SGD every
 * `gating_interval` steps] -> dual.  A "k smallest" selection is rank < k (ties: lower index first).
 */
typedef struct {
  int32_t L, H, d, Fh;                 /* C = H * d */
  float* const* w1; float* const* w3;  /* HOST arrays [L] of device pointers to the weights */
  float* const* m1; float* const* m3; float* const* m2;   /* HOST arrays [L] of device pointers to the .mask buffers (uvc_admm_masks) */
  float *c1, *c2, *c3;                 /* device workspace: [L,C] column norms, [L,H] head norms, [L,Fh] neuron norms */
  int32_t *rank1, *rank2, *rank3;      /* device workspace: rank inside head [L,C], across heads [L,H], across neurons [L,Fh] */
  float *s, *r, *y, *p, *z;            /* ADMM variables: [L,2], [L,H], [L,2], [L,H], [1] */
  const float* gate;                   /* block_skip_gating [L,2] or NULL */
  const float* gate_grad;              /* its task-loss gradient [L,2] or NULL */
  const float* noise;                  /* Gumbel noise [L,2] for this evaluation of the resource model (use_gumbel) */
  const float* macs;                   /* [L,6] per-block MACs at batch 1, as fp32 (joint_train.py:1010-1012) */
  double embed_macs, full_flops;       /* patch-embed MACs; dense model FLOPs (resource_ub, uvc_optimizer.py:178-188) */
  double lr;                           /* prox: the weight optimiser's current learning rate */
  float budget, z_grad_clip, slr, rlr, ylr, plr, zlr, sl2wd, gating_weight, eps;
  int32_t use_gumbel, gumbel_hard, warmup;
  float gate_mult;                     /* (global_step % gating_interval): weight of this step in the gate-gradient buffer */
  float* gate_grad_acc;                /* [L,2] running sum of the weighted gate gradients, or NULL */
  float* out;                          /* [1]: resource (FLOPs fraction of the dense model) of this evaluation */
} uvc_admm_args;

UVC_API int uvc_admm_scores(const uvc_admm_args* args, void* stream);    /* uvc_utils.py:54-73 for every layer + ranks */
UVC_API int uvc_admm_prox(const uvc_admm_args* args, void* stream);      /* uvc_utils.py:315-345 */
UVC_API int uvc_admm_masks(const uvc_admm_args* args, void* stream);     /* uvc_utils.py:376-401 */
UVC_API int uvc_admm_primal(const uvc_admm_args* args, void* stream);    /* uvc_optimizer.py:46-123 */
UVC_API int uvc_admm_dual(const uvc_admm_args* args, void* stream);      /* uvc_optimizer.py:126-135 */
UVC_API int uvc_admm_resource(const uvc_admm_args* args, void* stream);  /* uvc_utils.py:409-471 */

#ifdef __cplusplus
}
#endif
#endif  /* UVC_B200_H_ */

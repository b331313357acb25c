// Whole-model engine: DistilledVisionTransformer forward and backward as ONE C-ABI call each, so the
// Python host issues a single call per pass and the ~60 kernel launches per block are enqueued from
// native code on the caller's stream (no per-op Python / autograd dispatch, CUDA-graph capturable).
//
// Replaces models/model_distilled.py:429-531 (forward_features + forward) and its autograd backward:
//   patch embed (im2col + tcgen05 GEMM)  -> token assembly (cls, pos, patch/token gates)
//   L x [ LN -> QKV GEMM -> attention -> proj GEMM (+residual) -> LN -> fc1 GEMM (+GELU) -> fc2 GEMM (+residual) -> gate blend ]
//   final LN on the cls rows -> head GEMM.
// All activations the backward needs live in a caller-owned workspace carved by a deterministic bump
// allocator (same carve in forward and backward).  HBM layout: every activation is a row-major
// [B*ntok, width] fp32 matrix, rows 16-byte aligned; attention probabilities are [B, H, ntok, ldp].
#include "kernels.h"
#include <cuda_fp16.h>

namespace uvc {

namespace {

struct Bump {
  char* base; size_t off, cap;
  explicit Bump(void* p, size_t c) : base(static_cast<char*>(p)), off(0), cap(c) {}
  float* f(size_t n) {
    size_t bytes = (n * sizeof(float) + 255) & ~size_t(255);
    char* r = base ? base + off : nullptr;
    off += bytes;
    return reinterpret_cast<float*>(r);
  }
};

struct LayerWs {
  float *mean1, *rstd1, *mean2, *rstd2;
  float *ln1, *qkv, *P, *lse, *ctx, *x1, *ln2, *hpre, *h, *t, *xout;   // fused attention saves lse [B,H,ntok] instead of P
};

struct WeightsR {   // TF32-rounded copies of every GEMM weight (refreshed at the start of each forward)
  float *patch_w, *head_w;
  float *qkv_w[UVC_MAX_DEPTH], *proj_w[UVC_MAX_DEPTH], *fc1_w[UVC_MAX_DEPTH], *fc2_w[UVC_MAX_DEPTH];
};

struct Ws {
  WeightsR wr;
  // persistent (saved for backward)
  float *cols, *pe, *tok, *mean_f, *rstd_f, *cls_ln, *accum;
  LayerWs layer[UVC_MAX_DEPTH];
  // scratch
  float *g_a, *g_b, *g_c, *dh, *dqkv, *dP, *Dv, *dcls_ln, *dpe, *dlog_pad;
  size_t bytes;
};

struct Dims {
  int B, np, ntok, C, H, d, Fh, L, NC, Kp, img, patch, cin;
  long long M;
};

int check_dims(const uvc_vit_dims& v, Dims* o) {
  UVC_REQUIRE(v.B > 0 && v.C > 0 && v.H > 0 && v.Fh > 0 && v.L > 0 && v.num_classes > 0, UVC_ERR_BAD_SHAPE, "vit: non-positive dimension");
  UVC_REQUIRE(v.L <= UVC_MAX_DEPTH, UVC_ERR_BAD_SHAPE, "vit: depth %d > UVC_MAX_DEPTH %d", v.L, UVC_MAX_DEPTH);
  UVC_REQUIRE(v.patch > 0 && v.img % v.patch == 0 && v.patch % 4 == 0, UVC_ERR_BAD_SHAPE, "vit: patch %d must divide img %d and be a multiple of 4", v.patch, v.img);
  UVC_REQUIRE(v.C % v.H == 0 && (v.C / v.H) % 4 == 0, UVC_ERR_BAD_SHAPE, "vit: C=%d must split into H=%d heads of a multiple of 4", v.C, v.H);
  UVC_REQUIRE(v.C % 4 == 0 && v.Fh % 4 == 0 && v.C <= 1024, UVC_ERR_BAD_SHAPE, "vit: C, Fh must be multiples of 4 and C <= 1024");
  o->B = v.B; o->img = v.img; o->patch = v.patch; o->cin = v.in_chans;
  const int g = v.img / v.patch;
  o->np = g * g; o->ntok = o->np + 1;
  UVC_REQUIRE(o->ntok <= 256, UVC_ERR_BAD_SHAPE, "vit: %d tokens > 256 unsupported", o->ntok);
  o->C = v.C; o->H = v.H; o->d = v.C / v.H; o->Fh = v.Fh; o->L = v.L; o->NC = v.num_classes;
  o->Kp = v.in_chans * v.patch * v.patch;
  o->M = (long long)v.B * o->ntok;
  UVC_REQUIRE(o->M < (1ll << 31), UVC_ERR_BAD_SHAPE, "vit: too many rows");
  if (v.operand_f16)
    UVC_REQUIRE(o->d == 64 && attn_f16_ok(o->ntok, o->d) && (v.C & 7) == 0 && (v.Fh & 7) == 0, UVC_ERR_BAD_SHAPE,
                "vit: fp16 operand storage needs head dim 64, <= 208 tokens, C and Fh multiples of 8 (got d=%d, tokens=%d, C=%d, Fh=%d)", o->d, o->ntok, v.C, v.Fh);
  return UVC_OK;
}

// carve the workspace; with base == NULL only sizes are computed
void carve(const Dims& D, bool save, void* base, size_t cap, Ws* w) {
  Bump b(base, cap);
  const size_t M = (size_t)D.M, C = D.C, Fh = D.Fh;
  const bool fused_attn = attn_fused_ok(D.ntok, D.d);       // fused tcgen05 attention: no probabilities in HBM, 4 B per row of statistics instead
  const size_t psz = fused_attn ? 0 : (size_t)D.B * D.H * D.ntok * attn_ldp(D.ntok);
  const size_t lsz = fused_attn ? (size_t)D.B * D.H * D.ntok : 0;
  w->wr.patch_w = b.f(C * (size_t)D.Kp); w->wr.head_w = b.f((size_t)D.NC * C);
  for (int l = 0; l < D.L; ++l) {
    w->wr.qkv_w[l] = b.f(3 * C * C); w->wr.proj_w[l] = b.f(C * C); w->wr.fc1_w[l] = b.f(Fh * C); w->wr.fc2_w[l] = b.f(C * Fh);
  }
  w->cols = b.f((size_t)D.B * D.np * D.Kp);
  w->pe = b.f((size_t)D.B * D.np * C);
  w->tok = b.f(M * C);
  w->mean_f = b.f(D.B); w->rstd_f = b.f(D.B);
  w->cls_ln = b.f((size_t)D.B * C);
  w->accum = b.f(M * C);
  if (save) {
    for (int l = 0; l < D.L; ++l) {
      LayerWs& L = w->layer[l];
      L.mean1 = b.f(M); L.rstd1 = b.f(M); L.mean2 = b.f(M); L.rstd2 = b.f(M);
      L.ln1 = b.f(M * C); L.qkv = b.f(M * 3 * C); L.P = psz ? b.f(psz) : nullptr; L.lse = lsz ? b.f(lsz) : nullptr; L.ctx = b.f(M * C);
      L.x1 = b.f(M * C); L.ln2 = b.f(M * C); L.hpre = b.f((M * Fh + 1) / 2); L.h = b.f(M * Fh);   // hpre: gelu' as fp16
      L.t = b.f(M * C); L.xout = b.f(M * C);
    }
    w->g_a = b.f(M * C); w->g_b = b.f(M * C); w->g_c = b.f(M * C);
    w->dh = b.f(M * Fh); w->dqkv = b.f(M * 3 * C); w->dP = psz ? b.f(psz) : nullptr; w->Dv = lsz ? b.f(lsz) : nullptr;
    w->dcls_ln = b.f((size_t)D.B * C); w->dpe = b.f((size_t)D.B * D.np * C);
    w->dlog_pad = b.f((size_t)D.B * ((D.NC + 3) / 4 * 4));
  } else {
    // inference: every layer reuses one set of buffers; the residual stream ping-pongs between t and xout
    LayerWs L0;
    L0.mean1 = L0.rstd1 = L0.mean2 = L0.rstd2 = nullptr;
    L0.ln1 = b.f(M * C); L0.qkv = b.f(M * 3 * C); L0.P = psz ? b.f(psz) : nullptr; L0.lse = nullptr; L0.ctx = b.f(M * C);
    L0.x1 = b.f(M * C); L0.ln2 = L0.ln1; L0.hpre = nullptr; L0.h = b.f(M * Fh);
    L0.t = b.f(M * C); L0.xout = b.f(M * C);
    float* ping = L0.xout; float* pong = b.f(M * C);
    for (int l = 0; l < D.L; ++l) { w->layer[l] = L0; w->layer[l].xout = (l & 1) ? pong : ping; }
    w->g_a = w->g_b = w->g_c = w->dh = w->dqkv = w->dP = w->Dv = w->dcls_ln = w->dpe = w->dlog_pad = nullptr;
  }
  w->bytes = b.off;
}

#define UVC_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

int linear_fwd(const float* X, long long ldx, const float* W, const float* bias, float* Y, long long ldy, int M, int N, int K, cudaStream_t st,
               int extra_flags = 0, float* aux = nullptr, const float* R = nullptr, long long ldr = 0) {
  uvc_gemm_args a = gemm_args(M, N, K, op_k(X, ldx), op_k(W, K), Y, ldy);
  a.flags = extra_flags;
  if (bias) { a.bias = bias; a.flags |= UVC_EPI_BIAS; }
  if (aux) { a.aux = aux; a.ldaux = ldy; }
  if (R) { a.R = R; a.ldr = ldr; a.flags |= UVC_EPI_RESIDUAL; }
  return gemm_tf32(a, st);
}
// dX[M,K] = dY[M,N] W[N,K]   (W read MN-major)
int linear_dgrad(const float* dY, long long lddy, const float* W, float* dX, long long lddx, int M, int N, int K, cudaStream_t st,
                 int extra_flags = 0, float* aux = nullptr, long long ldaux = 0, float* colsum_out = nullptr, const float* scale_dev = nullptr) {
  uvc_gemm_args a = gemm_args(M, K, N, op_k(dY, lddy), op_mn(W, K), dX, lddx);
  a.flags = extra_flags;
  a.alpha_dev = scale_dev;              // dY is scaled by a device scalar (the block gate) without materialising the product
  if (aux) { a.aux = aux; a.ldaux = ldaux; }
  if (colsum_out) { a.colsum = colsum_out; a.flags |= UVC_EPI_COLSUM; }   // bias gradient of the layer below, summed in the epilogue
  return gemm_tf32(a, st);
}
// dW[N,K] += dY[M,N]^T X[M,K]   (split-K, atomic accumulate) ; db[N] += colsum(dY)
int linear_wgrad(const float* dY, long long lddy, const float* X, long long ldx, float* dW, float* db, int M, int N, int K, cudaStream_t st,
                 const float* scale_dev = nullptr) {
  uvc_gemm_args a = gemm_args(N, K, M, op_mn(dY, lddy), op_mn(X, ldx), dW, K);
  a.flags = UVC_EPI_ATOMIC;
  a.alpha_dev = scale_dev;
  a.splits = wgrad_splits(N, K, M);
  UVC_TRY(gemm_tf32(a, st));
  if (db) UVC_TRY(colsum(dY, lddy, M, N, nullptr, db, st));
  return UVC_OK;
}

// weights -> TF32-rounded copies in the workspace (one launch per <= 96 tensors)
int round_weights(const uvc_vit_tensors& p, const Dims& D, const WeightsR& wr, cudaStream_t st) {
  const float* src[2 + 4 * UVC_MAX_DEPTH]; float* dst[2 + 4 * UVC_MAX_DEPTH]; long long n[2 + 4 * UVC_MAX_DEPTH];
  int k = 0;
  const long long C = D.C, Fh = D.Fh;
  if (p.patch_w) { src[k] = p.patch_w; dst[k] = wr.patch_w; n[k++] = C * D.Kp; }   // absent when the caller supplies the token embeddings
  src[k] = p.head_w; dst[k] = wr.head_w; n[k++] = (long long)D.NC * C;
  for (int l = 0; l < D.L; ++l) {
    const uvc_block_tensors& b = p.blocks[l];
    src[k] = b.qkv_w; dst[k] = wr.qkv_w[l]; n[k++] = 3 * C * C;
    src[k] = b.proj_w; dst[k] = wr.proj_w[l]; n[k++] = C * C;
    src[k] = b.fc1_w; dst[k] = wr.fc1_w[l]; n[k++] = Fh * C;
    src[k] = b.fc2_w; dst[k] = wr.fc2_w[l]; n[k++] = C * Fh;
  }
  return round_tf32_segs(src, dst, n, k, st);
}

// TMA operands need 16-byte row pitches: a class count that is not a multiple of 4 (cifar10: 10) gets a zero-padded copy of dlogits
int pad_dlogits(const float** dlog, long long* ld, float* pad, int B, int NC, cudaStream_t st) {
  if ((NC & 3) == 0 && (reinterpret_cast<uintptr_t>(*dlog) & 15) == 0) return UVC_OK;
  const int NCp = (NC + 3) / 4 * 4;
  cudaError_t e = cudaMemsetAsync(pad, 0, (size_t)B * NCp * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(pad, (size_t)NCp * sizeof(float), *dlog, (size_t)NC * sizeof(float), (size_t)NC * sizeof(float), B, cudaMemcpyDeviceToDevice, st);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_backward: padding dlogits: %s", cudaGetErrorString(e));
  *dlog = pad; *ld = NCp;
  return UVC_OK;
}

int check_tensors(const uvc_vit_tensors& w, int L, const char* what, bool need_patch) {
  UVC_REQUIRE((!need_patch || (w.patch_w && w.patch_b)) && w.cls_token && w.pos_embed && w.norm_w && w.norm_b && w.head_w && w.head_b && w.blocks, UVC_ERR_BAD_ARG,
              "vit: %s has a NULL tensor", what);
  for (int l = 0; l < L; ++l) {
    const uvc_block_tensors& b = w.blocks[l];
    UVC_REQUIRE(b.norm1_w && b.norm1_b && b.qkv_w && b.proj_w && b.proj_b && b.norm2_w && b.norm2_b && b.fc1_w && b.fc1_b && b.fc2_w && b.fc2_b,
                UVC_ERR_BAD_ARG, "vit: %s block %d has a NULL tensor", what, l);
  }
  return UVC_OK;
}

}  // namespace

unsigned long long vit_workspace_bytes_f16(const Dims& D, bool save);

unsigned long long vit_workspace_bytes(const uvc_vit_dims& dims, int save) {
  Dims D;
  if (check_dims(dims, &D)) return 0;
  if (dims.operand_f16) return vit_workspace_bytes_f16(D, save != 0);
  Ws w;
  carve(D, save != 0, nullptr, 0, &w);
  return w.bytes;
}

int vit_forward_f16(const uvc_vit_forward_args& a, const Dims& D, cudaStream_t st);
int vit_backward_f16(const uvc_vit_backward_args& a, const Dims& D, cudaStream_t st);

int vit_forward(const uvc_vit_forward_args& a, cudaStream_t st) {
  Dims D;
  UVC_TRY(check_dims(a.dims, &D));
  if (a.dims.operand_f16) return vit_forward_f16(a, D, st);
  UVC_REQUIRE(!a.layout, UVC_ERR_BAD_ARG, "vit_forward: a compaction layout needs operand_f16 = 1");
  UVC_TRY(check_tensors(a.w, D.L, "w", a.pe_in == nullptr));
  UVC_REQUIRE((a.x || a.pe_in) && a.logits && a.workspace, UVC_ERR_BAD_ARG, "vit_forward: NULL x / logits / workspace");
  const bool save = a.save_for_backward != 0;
  Ws w;
  carve(D, save, a.workspace, a.workspace_bytes, &w);
  UVC_REQUIRE(w.bytes <= a.workspace_bytes, UVC_ERR_WORKSPACE, "vit_forward: workspace %llu bytes < required %llu",
              (unsigned long long)a.workspace_bytes, (unsigned long long)w.bytes);
  const int M = (int)D.M, C = D.C, Fh = D.Fh;
  const float eps = a.dims.ln_eps;
  const float scale = 1.0f / sqrtf((float)D.d);

  // patch embed: 16x16/16 conv == GEMM over im2col rows
  if (!a.weights_converted) UVC_TRY(round_weights(a.w, D, w.wr, st));
  const float* pe = a.pe_in;
  if (!pe) {
    UVC_TRY(im2col16(a.x, w.cols, D.B, D.cin, D.img, D.patch, st, 1));
    float* pe_w = a.pe_out ? a.pe_out : w.pe;
    UVC_TRY(linear_fwd(w.cols, D.Kp, w.wr.patch_w, a.w.patch_b, pe_w, C, D.B * D.np, C, D.Kp, st));
    pe = pe_w;
  } else if (a.pe_out && a.pe_out != pe) {
    cudaError_t e = cudaMemcpyAsync(a.pe_out, pe, (size_t)D.B * D.np * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_forward: memcpy pe_out: %s", cudaGetErrorString(e));
  }
  UVC_TRY(assemble_tokens(pe, a.w.cls_token, a.w.pos_embed, a.patch_scale, a.token_mask, w.tok, D.B, D.np, C, st));
  if (pe != w.pe && save) {   // the gate gradients of the token assembly read the raw embeddings
    cudaError_t e = cudaMemcpyAsync(w.pe, pe, (size_t)D.B * D.np * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_forward: memcpy pe: %s", cudaGetErrorString(e));
  }
  if (a.enable_jumping) {
    cudaError_t e = cudaMemsetAsync(w.accum, 0, (size_t)M * C * sizeof(float), st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_forward: memset accum: %s", cudaGetErrorString(e));
  }

  const float* x = w.tok;
  for (int l = 0; l < D.L; ++l) {
    const bool skipped = a.skip_host && a.skip_host[l];
    if (!skipped) {
      const uvc_block_tensors& p = a.w.blocks[l];
      LayerWs& L = w.layer[l];
      UVC_TRY(layernorm_fwd(x, C, p.norm1_w, p.norm1_b, eps, L.ln1, C, L.mean1, L.rstd1, M, C, st, 1));
      UVC_TRY(linear_fwd(L.ln1, C, w.wr.qkv_w[l], p.qkv_b, L.qkv, 3 * C, M, 3 * C, C, st, UVC_EPI_ROUND_TF32));
      UVC_TRY(attention_fwd(L.qkv, L.P, L.ctx, D.B, D.H, D.ntok, D.d, scale, st, save && L.P != nullptr, save ? L.lse : nullptr));   // fused path: probabilities never reach HBM
      UVC_TRY(linear_fwd(L.ctx, C, w.wr.proj_w[l], p.proj_b, L.x1, C, M, C, C, st, 0, nullptr, x, C));     // x1 = x + proj(ctx)
      UVC_TRY(layernorm_fwd(L.x1, C, p.norm2_w, p.norm2_b, eps, L.ln2, C, L.mean2, L.rstd2, M, C, st, 1));
      UVC_TRY(linear_fwd(L.ln2, C, w.wr.fc1_w[l], p.fc1_b, L.h, Fh, M, Fh, C, st, UVC_EPI_GELU | UVC_EPI_ROUND_TF32 | UVC_EPI_AUX_F16, L.hpre));   // h = gelu(fc1); hpre holds gelu'(fc1) (fp16) for the backward
      if (a.blend) {
        UVC_TRY(linear_fwd(L.h, Fh, w.wr.fc2_w[l], p.fc2_b, L.t, C, M, C, Fh, st, 0, nullptr, L.x1, C));    // t = x1 + fc2(h)
        UVC_TRY(blend_fwd(L.t, x, a.blend + 2 * l, L.xout, (long long)M * C, st));                            // x <- d1 t + d0 x
      } else {
        UVC_TRY(linear_fwd(L.h, Fh, w.wr.fc2_w[l], p.fc2_b, L.xout, C, M, C, Fh, st, 0, nullptr, L.x1, C));
      }
      x = L.xout;
    }
    if (a.enable_jumping) UVC_TRY(scale_add(w.accum, x, nullptr, 1.0f, (long long)M * C, st));
  }
  const float* xf = a.enable_jumping ? w.accum : x;
  // final LayerNorm only on the cls rows (row stride ntok*C), then the classifier head
  UVC_TRY(layernorm_fwd(xf, (long long)D.ntok * C, a.w.norm_w, a.w.norm_b, eps, w.cls_ln, C, w.mean_f, w.rstd_f, D.B, C, st, 1));
  UVC_TRY(linear_fwd(w.cls_ln, C, w.wr.head_w, a.w.head_b, a.logits, D.NC, D.B, D.NC, C, st));
  return UVC_OK;
}

int vit_backward(const uvc_vit_backward_args& a, cudaStream_t st) {
  Dims D;
  UVC_TRY(check_dims(a.dims, &D));
  if (a.dims.operand_f16) return vit_backward_f16(a, D, st);
  UVC_REQUIRE(!a.layout, UVC_ERR_BAD_ARG, "vit_backward: a compaction layout needs operand_f16 = 1");
  UVC_TRY(check_tensors(a.w, D.L, "w", a.d_pe == nullptr));
  UVC_TRY(check_tensors(a.g, D.L, "g", a.d_pe == nullptr));
  UVC_REQUIRE(a.dlogits && a.workspace, UVC_ERR_BAD_ARG, "vit_backward: NULL dlogits / workspace");
  UVC_REQUIRE(!a.blend || a.d_blend, UVC_ERR_BAD_ARG, "vit_backward: blend given without d_blend");
  Ws w;
  carve(D, true, a.workspace, a.workspace_bytes, &w);
  UVC_REQUIRE(w.bytes <= a.workspace_bytes, UVC_ERR_WORKSPACE, "vit_backward: workspace %llu bytes < required %llu",
              (unsigned long long)a.workspace_bytes, (unsigned long long)w.bytes);
  const int M = (int)D.M, C = D.C, Fh = D.Fh;
  const float scale = 1.0f / sqrtf((float)D.d);
  const size_t xbytes = (size_t)M * C * sizeof(float);

  // residual-stream input of every block (skipped blocks forward their input)
  const float* xin[UVC_MAX_DEPTH + 1];
  xin[0] = w.tok;
  for (int l = 0; l < D.L; ++l) xin[l + 1] = (a.skip_host && a.skip_host[l]) ? xin[l] : w.layer[l].xout;
  const float* xf = a.enable_jumping ? w.accum : xin[D.L];

  // head: dW += dlogits^T cls_ln ; db += colsum ; dcls_ln = dlogits W
  const float* dlog = a.dlogits; long long ldl = D.NC;
  UVC_TRY(pad_dlogits(&dlog, &ldl, w.dlog_pad, D.B, D.NC, st));
  UVC_TRY(linear_wgrad(dlog, ldl, w.cls_ln, C, a.g.head_w, a.g.head_b, D.B, D.NC, C, st));
  UVC_TRY(linear_dgrad(dlog, ldl, w.wr.head_w, w.dcls_ln, C, D.B, D.NC, C, st));
  // final LN backward on the cls rows; every other row of the stream gradient is zero
  float* g = w.g_a;         // gradient wrt the current residual stream
  float* g_jump = nullptr;  // with jumping connections the final-norm gradient reaches every block output
  cudaError_t e = cudaMemsetAsync(g, 0, xbytes, st);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_backward: memset: %s", cudaGetErrorString(e));
  UVC_TRY(layernorm_bwd(w.dcls_ln, C, xf, (long long)D.ntok * C, w.mean_f, w.rstd_f, a.w.norm_w, nullptr, nullptr, nullptr, g,
                        (long long)D.ntok * C, a.g.norm_w, a.g.norm_b, D.B, C, st));
  float* spare1 = w.g_b;
  float* spare2 = w.g_c;
  if (a.enable_jumping) {
    // keep the final-norm gradient in accum's storage (accum itself is no longer needed after the LN backward)
    g_jump = w.accum;
    e = cudaMemcpyAsync(g_jump, g, xbytes, cudaMemcpyDeviceToDevice, st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_backward: memcpy: %s", cudaGetErrorString(e));
  }

  for (int l = D.L - 1; l >= 0; --l) {
    const bool skipped = a.skip_host && a.skip_host[l];
    if (!skipped) {
      const uvc_block_tensors& p = a.w.blocks[l];
      const uvc_block_tensors& gp = a.g.blocks[l];
      const LayerWs& L = w.layer[l];
      const float* x = xin[l];
      const float* d = a.blend ? a.blend + 2 * l : nullptr;
      // gate blend: x_out = d1 t + d0 x  ->  dt = d1 g, dx += d0 g, dd0 = <g,x>, dd1 = <g,t>.  dt is never materialised: the two fc2 GEMMs take d1
      // as a device-scalar alpha and the LN2 backward adds d1 * g as its scaled residual input.
      const float* dt = g;
      const float* d1 = d ? d + 1 : nullptr;
      if (d) UVC_TRY(blend_dots(g, L.t, x, a.d_blend + 2 * l, (long long)M * C, st));
      // ---- MLP:  t = x1 + fc2(gelu(fc1(ln2)))
      // Bias gradients are column sums of tensors other kernels stream anyway, so they ride along instead of costing a pass each:
      //   fc1_b <- epilogue of the fc2 dgrad GEMM (sums dhpre), fc2_b / proj_b <- the LN2 backward (sums its residual input dt / its output dx1).
      UVC_TRY(linear_wgrad(dt, C, L.h, Fh, gp.fc2_w, nullptr, M, C, Fh, st, d1));
      UVC_TRY(linear_dgrad(dt, C, w.wr.fc2_w[l], w.dh, Fh, M, C, Fh, st, UVC_EPI_GELU_BWD | UVC_EPI_ROUND_TF32 | UVC_EPI_AUX_F16, L.hpre, Fh, gp.fc1_b, d1));     // dhpre
      UVC_TRY(linear_wgrad(w.dh, Fh, L.ln2, C, gp.fc1_w, nullptr, M, Fh, C, st));
      UVC_TRY(linear_dgrad(w.dh, Fh, w.wr.fc1_w[l], spare2, C, M, Fh, C, st));                                          // dln2
      // dx1 = dt + LN2'(dln2)
      UVC_TRY(layernorm_bwd(spare2, C, L.x1, C, L.mean2, L.rstd2, p.norm2_w, d ? nullptr : dt, d ? g : nullptr, d1, spare2, C, gp.norm2_w, gp.norm2_b,
                            M, C, st, gp.fc2_b, gp.proj_b));
      float* dx1 = spare2;
      // ---- attention:  x1 = x + proj(ctx)
      UVC_TRY(linear_wgrad(dx1, C, L.ctx, C, gp.proj_w, nullptr, M, C, C, st));
      float* dctx = spare1;
      UVC_TRY(linear_dgrad(dx1, C, w.wr.proj_w[l], dctx, C, M, C, C, st, UVC_EPI_ROUND_TF32));
      // fused path: S, P recomputed in tensor memory; its epilogues also sum the columns of dqkv (the qkv bias gradient)
      if (L.lse) UVC_TRY(attention_bwd_fused(L.qkv, L.lse, L.ctx, dctx, w.Dv, w.dqkv, D.B, D.H, D.ntok, scale, st, gp.qkv_b));
      else UVC_TRY(attention_bwd(L.qkv, L.P, dctx, w.dP, w.dqkv, D.B, D.H, D.ntok, D.d, scale, st));
      UVC_TRY(linear_wgrad(w.dqkv, 3 * C, L.ln1, C, gp.qkv_w, L.lse ? nullptr : gp.qkv_b, M, 3 * C, C, st));
      UVC_TRY(linear_dgrad(w.dqkv, 3 * C, w.wr.qkv_w[l], spare1, C, M, 3 * C, C, st));                                  // dln1
      // dx = dx1 + LN1'(dln1) + d0 g        (written over spare1)
      UVC_TRY(layernorm_bwd(spare1, C, x, C, L.mean1, L.rstd1, p.norm1_w, dx1, d ? g : nullptr, d, spare1, C, gp.norm1_w, gp.norm1_b, M, C, st));
      // rotate buffers: new stream gradient is spare1
      float* old = g; g = spare1; spare1 = old;
    }
    if (g_jump && l > 0) UVC_TRY(scale_add(g, g_jump, nullptr, 1.0f, (long long)M * C, st));   // output of block l-1 also feeds accum
  }

  // token assembly + patch embed
  const int rows = D.B * D.np;
  float* dpe = a.d_pe ? a.d_pe : w.dpe;
  UVC_TRY(assemble_tokens_bwd(g, w.pe, a.patch_scale, a.token_mask, dpe, a.patch_scale ? a.d_patch_scale : nullptr,
                              a.token_mask ? a.d_token_mask : nullptr, a.g.pos_embed, a.g.cls_token, D.B, D.np, C, st));
  if (!a.d_pe) UVC_TRY(linear_wgrad(w.dpe, C, w.cols, D.Kp, a.g.patch_w, a.g.patch_b, rows, C, D.Kp, st));
  return UVC_OK;
}

// ====================================================================================================================
// 16-bit operand storage (uvc_vit_dims.operand_f16): the same model, but every tensor that only feeds GEMMs lives in HBM as fp16 -- the 10
// mantissa bits the TF32 path keeps, at half the bytes and twice the tensor-core rate (tcgen05.mma kind::f16, fp32 accumulation).
//   forward : LayerNorm writes fp16; qkv, gelu(fc1) and the attention context are written ONLY as fp16 by their producers; the residual
//             stream (tok, x1, block outputs), LayerNorm statistics, softmax, logits and the loss stay fp32.
//   backward: the fp32 stream gradient g is carried unscaled; its GEMM-operand copy is g16 = fp16(S g) with the loss scale S (a power of two),
//             so gradients of 1e-6 stay normal fp16 numbers.  Everything downstream of g16 is linear in it: the fp16 intermediates (dh, dln,
//             dctx, dqkv) carry S, and S is taken back out where fp32 results are produced -- alpha = 1/S in the weight-gradient GEMMs (exact,
//             a power of two), 1/S on the fused bias-gradient column sums, and on load in the LayerNorm backward.
//   weights : converted once per forward to fp16 [out, in] (forward B operand) and, for training, fp16 [in, out] (data-gradient B operand, so
//             only the weight-gradient GEMMs read 16-bit operands MN-major).
// The patch-embed and head GEMMs (1.5 % of the FLOPs, K = 768 image pixels / B rows) stay on the TF32 path.
// ====================================================================================================================
namespace {

typedef __half h16;

struct Layer16 {
  float *mean1, *rstd1, *mean2, *rstd2, *lse, *x1, *t, *xout;
  h16 *ln1, *qkv, *ctx, *ln2, *hpre, *h;
};
struct Ws16 {
  float *patch_w, *head_w;
  h16 *qkv_w[UVC_MAX_DEPTH], *proj_w[UVC_MAX_DEPTH], *fc1_w[UVC_MAX_DEPTH], *fc2_w[UVC_MAX_DEPTH];
  h16 *qkv_wT[UVC_MAX_DEPTH], *proj_wT[UVC_MAX_DEPTH], *fc1_wT[UVC_MAX_DEPTH], *fc2_wT[UVC_MAX_DEPTH];
  float *cols, *pe, *tok, *mean_f, *rstd_f, *cls_ln, *accum;
  Layer16 layer[UVC_MAX_DEPTH];
  float *g_a, *g_b, *g_c, *Dv, *dcls_ln, *dpe, *dlog_pad, *scales;
  h16 *g16, *dx1_16, *dln16, *dctx16, *dh16, *dqkv16;
  // Stage-2 compaction (uvc_vit_layout): gathered qkv / fc1 biases per block, and one block's compact weight / bias gradients
  float *qkv_bc[UVC_MAX_DEPTH], *fc1_bc[UVC_MAX_DEPTH];
  float *cg_qkv_w, *cg_qkv_b, *cg_proj_w, *cg_fc1_w, *cg_fc1_b, *cg_fc2_w, *cg_fc2_b; size_t cg_floats;
  size_t bytes;
};

// widths of block l: dense, or the live widths of a compacted block
struct BlockShape { int Hl, Cl, Ql, Fl; const int* hidx; const int* nidx; };
inline BlockShape block_shape(const Dims& D, const uvc_vit_layout* lay, int l) {
  BlockShape s;
  s.Hl = lay ? lay->n_heads[l] : D.H; s.Cl = s.Hl * D.d; s.Ql = 3 * s.Cl; s.Fl = lay ? lay->n_neurons[l] : D.Fh;
  s.hidx = lay ? lay->head_idx + (size_t)l * D.H : nullptr; s.nidx = lay ? lay->neuron_idx + (size_t)l * D.Fh : nullptr;
  return s;
}
int check_layout(const uvc_vit_layout* lay, const Dims& D, const uint8_t* skip_host) {
  if (!lay) return UVC_OK;
  UVC_REQUIRE(lay->head_idx && lay->neuron_idx, UVC_ERR_BAD_ARG, "vit: layout without index arrays");
  for (int l = 0; l < D.L; ++l) {
    if (skip_host && skip_host[l]) continue;
    UVC_REQUIRE(lay->n_heads[l] >= 1 && lay->n_heads[l] <= D.H, UVC_ERR_BAD_SHAPE, "vit: layout block %d keeps %d heads (1..%d)", l, lay->n_heads[l], D.H);
    UVC_REQUIRE(lay->n_neurons[l] >= 64 && lay->n_neurons[l] <= D.Fh && lay->n_neurons[l] % 64 == 0, UVC_ERR_BAD_SHAPE,
                "vit: layout block %d keeps %d neurons (64..%d, a multiple of 64)", l, lay->n_neurons[l], D.Fh);
  }
  return UVC_OK;
}

struct Bump16 : Bump {
  using Bump::Bump;
  h16* h(size_t n) { return reinterpret_cast<h16*>(f((n + 1) / 2)); }
};

void carve16(const Dims& D, bool save, void* base, size_t cap, Ws16* w) {
  Bump16 b(base, cap);
  const size_t M = (size_t)D.M, C = D.C, Fh = D.Fh;
  const size_t lsz = (size_t)D.B * D.H * D.ntok;
  w->patch_w = b.f(C * (size_t)D.Kp); w->head_w = b.f((size_t)D.NC * C);
  for (int l = 0; l < D.L; ++l) {
    w->qkv_w[l] = b.h(3 * C * C); w->proj_w[l] = b.h(C * C); w->fc1_w[l] = b.h(Fh * C); w->fc2_w[l] = b.h(C * Fh);
    if (save) { w->qkv_wT[l] = b.h(3 * C * C); w->proj_wT[l] = b.h(C * C); w->fc1_wT[l] = b.h(Fh * C); w->fc2_wT[l] = b.h(C * Fh); }
    else w->qkv_wT[l] = w->proj_wT[l] = w->fc1_wT[l] = w->fc2_wT[l] = nullptr;
  }
  for (int l = 0; l < D.L; ++l) { w->qkv_bc[l] = b.f(3 * C); w->fc1_bc[l] = b.f(Fh); }
  w->cols = b.f((size_t)D.B * D.np * D.Kp);
  w->pe = b.f((size_t)D.B * D.np * C);
  w->tok = b.f(M * C);
  w->mean_f = b.f(D.B); w->rstd_f = b.f(D.B);
  w->cls_ln = b.f((size_t)D.B * C);
  w->accum = b.f(M * C);
  w->cg_qkv_w = w->cg_qkv_b = w->cg_proj_w = w->cg_fc1_w = w->cg_fc1_b = w->cg_fc2_w = w->cg_fc2_b = nullptr; w->cg_floats = 0;
  if (save) {
    // one contiguous run (zeroed by one memset per block)
    const size_t cg0 = b.off;
    w->cg_qkv_w = b.f(3 * C * C); w->cg_qkv_b = b.f(3 * C); w->cg_proj_w = b.f(C * C); w->cg_fc1_w = b.f(Fh * C); w->cg_fc1_b = b.f(Fh);
    w->cg_fc2_w = b.f(C * Fh); w->cg_fc2_b = b.f(C);
    w->cg_floats = (b.off - cg0) / sizeof(float);
    for (int l = 0; l < D.L; ++l) {
      Layer16& L = w->layer[l];
      L.mean1 = b.f(M); L.rstd1 = b.f(M); L.mean2 = b.f(M); L.rstd2 = b.f(M); L.lse = b.f(lsz);
      L.ln1 = b.h(M * C); L.qkv = b.h(M * 3 * C); L.ctx = b.h(M * C); L.x1 = b.f(M * C);
      L.ln2 = b.h(M * C); L.hpre = b.h(M * Fh); L.h = b.h(M * Fh);
      L.t = b.f(M * C); L.xout = b.f(M * C);
    }
    w->g_a = b.f(M * C); w->g_b = b.f(M * C); w->g_c = b.f(M * C);
    w->Dv = b.f(lsz); w->dcls_ln = b.f((size_t)D.B * C); w->dpe = b.f((size_t)D.B * D.np * C);
    w->dlog_pad = b.f((size_t)D.B * ((D.NC + 3) / 4 * 4));
    w->scales = b.f(4);
    w->g16 = b.h(M * C); w->dx1_16 = b.h(M * C); w->dln16 = b.h(M * C); w->dctx16 = b.h(M * C); w->dh16 = b.h(M * Fh); w->dqkv16 = b.h(M * 3 * C);
  } else {
    Layer16 L0;
    L0.mean1 = L0.rstd1 = L0.mean2 = L0.rstd2 = L0.lse = nullptr;
    L0.ln1 = b.h(M * C); L0.qkv = b.h(M * 3 * C); L0.ctx = b.h(M * C); L0.x1 = b.f(M * C); L0.ln2 = L0.ln1; L0.hpre = nullptr; L0.h = b.h(M * Fh);
    L0.t = b.f(M * C); L0.xout = b.f(M * C);
    float* ping = L0.xout; float* pong = b.f(M * C);
    for (int l = 0; l < D.L; ++l) { w->layer[l] = L0; w->layer[l].xout = (l & 1) ? pong : ping; }
    w->g_a = w->g_b = w->g_c = w->Dv = w->dcls_ln = w->dpe = w->dlog_pad = w->scales = nullptr;
    w->g16 = w->dx1_16 = w->dln16 = w->dctx16 = w->dh16 = w->dqkv16 = nullptr;
  }
  w->bytes = b.off;
}

inline uvc_operand op16_k(const h16* p, long long ld) { return uvc_operand{reinterpret_cast<const float*>(p), ld, 0, 0, 0, 0}; }
inline uvc_operand op16_mn(const h16* p, long long ld) { return uvc_operand{reinterpret_cast<const float*>(p), ld, 0, 0, 1, 0}; }

// Y = epilogue(X16 W16^T): fp32 output D (with optional fp32 residual R) and / or fp16 output D16
int linear16(const h16* X, long long ldx, const h16* W, const float* bias, float* Dout, void* D16, long long ldd, int M, int N, int K, cudaStream_t st,
             int extra_flags = 0, void* aux16 = nullptr, const float* R = nullptr, long long ldr = 0, const float* alpha_dev = nullptr,
             float* colsum_out = nullptr, const float* colsum_scale_dev = nullptr, const float* blend_dev = nullptr, const float* R2 = nullptr,
             float* D2 = nullptr) {
  uvc_gemm_args a = gemm_args(M, N, K, op16_k(X, ldx), op16_k(W, K), Dout, ldd);
  a.flags = extra_flags | UVC_GEMM_F16;
  a.D16 = D16; a.ldd16 = ldd;
  a.alpha_dev = alpha_dev;
  if (bias) { a.bias = bias; a.flags |= UVC_EPI_BIAS; }
  if (aux16) { a.aux = static_cast<float*>(aux16); a.ldaux = ldd; a.flags |= UVC_EPI_AUX_F16; }
  if (R) { a.R = R; a.ldr = ldr; a.flags |= UVC_EPI_RESIDUAL; }
  if (colsum_out) { a.colsum = colsum_out; a.colsum_scale_dev = colsum_scale_dev; a.flags |= UVC_EPI_COLSUM; }
  if (blend_dev) { a.blend_dev = blend_dev; a.R2 = R2; a.ldr2 = ldd; a.D2 = D2; a.ldd2 = ldd; a.flags |= UVC_EPI_BLEND; }
  return gemm_tf32(a, st);
}
// dW[N,K] += (*inv_scale_dev) * dY16[M,N]^T X16[M,K]   (both operands MN-major fp16, split-K with fp32 atomics)
int linear_wgrad16(const h16* dY, long long lddy, const h16* X, long long ldx, float* dW, int M, int N, int K, const float* inv_scale_dev, cudaStream_t st,
                   const float* scale_dev = nullptr) {
  uvc_gemm_args a = gemm_args(N, K, M, op16_mn(dY, lddy), op16_mn(X, ldx), dW, K);
  a.flags = UVC_EPI_ATOMIC | UVC_GEMM_F16;
  a.alpha_dev = scale_dev;
  a.alpha_dev2 = inv_scale_dev;
  a.splits = wgrad_splits(N, K, M, 64);
  return gemm_tf32(a, st);
}

int convert_weights16(const uvc_vit_tensors& p, const Dims& D, const Ws16& w, bool save, cudaStream_t st, const uvc_vit_layout* lay = nullptr,
                      const uint8_t* skip_host = nullptr) {
  if (lay) {
    // compacted blocks: gather the live rows / columns while converting (compact.cu); biases of the compacted outputs ride along
    GatherSeg seg[6 * UVC_MAX_DEPTH];
    int k = 0;
    const int C = D.C, d = D.d;
    for (int l = 0; l < D.L; ++l) {
      if (skip_host && skip_host[l]) continue;
      const uvc_block_tensors& b = p.blocks[l];
      const BlockShape s = block_shape(D, lay, l);
      const AxisMap none{nullptr, 0, 0, 0};
      const AxisMap qrows{s.hidx, s.Cl, d, C}, hcols{s.hidx, s.Cl, d, 0}, nmap{s.nidx, s.Fl, 1, 0};
      seg[k++] = GatherSeg{b.qkv_w, C, w.qkv_w[l], save ? w.qkv_wT[l] : nullptr, nullptr, s.Ql, C, qrows, none};
      seg[k++] = GatherSeg{b.proj_w, C, w.proj_w[l], save ? w.proj_wT[l] : nullptr, nullptr, C, s.Cl, none, hcols};
      seg[k++] = GatherSeg{b.fc1_w, C, w.fc1_w[l], save ? w.fc1_wT[l] : nullptr, nullptr, s.Fl, C, nmap, none};
      seg[k++] = GatherSeg{b.fc2_w, D.Fh, w.fc2_w[l], save ? w.fc2_wT[l] : nullptr, nullptr, C, s.Fl, none, nmap};
      if (b.qkv_b) seg[k++] = GatherSeg{b.qkv_b, 3 * C, nullptr, nullptr, w.qkv_bc[l], 1, s.Ql, none, AxisMap{s.hidx, s.Cl, d, C}};
      seg[k++] = GatherSeg{b.fc1_b, D.Fh, nullptr, nullptr, w.fc1_bc[l], 1, s.Fl, none, nmap};
    }
    UVC_TRY(gather_cvt(seg, k, st));
  } else {
    const float* src[4 * UVC_MAX_DEPTH]; void* dst[4 * UVC_MAX_DEPTH]; void* dstT[4 * UVC_MAX_DEPTH]; int rows[4 * UVC_MAX_DEPTH], cols[4 * UVC_MAX_DEPTH];
    int k = 0;
    for (int l = 0; l < D.L; ++l) {
      const uvc_block_tensors& b = p.blocks[l];
      src[k] = b.qkv_w; dst[k] = w.qkv_w[l]; dstT[k] = w.qkv_wT[l]; rows[k] = 3 * D.C; cols[k++] = D.C;
      src[k] = b.proj_w; dst[k] = w.proj_w[l]; dstT[k] = w.proj_wT[l]; rows[k] = D.C; cols[k++] = D.C;
      src[k] = b.fc1_w; dst[k] = w.fc1_w[l]; dstT[k] = w.fc1_wT[l]; rows[k] = D.Fh; cols[k++] = D.C;
      src[k] = b.fc2_w; dst[k] = w.fc2_w[l]; dstT[k] = w.fc2_wT[l]; rows[k] = D.C; cols[k++] = D.Fh;
    }
    UVC_TRY(cvt_f16_segs(src, dst, save ? dstT : nullptr, rows, cols, k, st));
  }
  // patch embed / head stay TF32
  const float* s2[2]; float* d2[2]; long long n2[2]; int m = 0;
  if (p.patch_w) { s2[m] = p.patch_w; d2[m] = w.patch_w; n2[m++] = (long long)D.C * D.Kp; }
  s2[m] = p.head_w; d2[m] = w.head_w; n2[m++] = (long long)D.NC * D.C;
  return round_tf32_segs(s2, d2, n2, m, st);
}

}  // namespace

unsigned long long vit_workspace_bytes_f16(const Dims& D, bool save) {
  Ws16 w;
  carve16(D, save, nullptr, 0, &w);
  return w.bytes;
}

int vit_forward_f16(const uvc_vit_forward_args& a, const Dims& D, cudaStream_t st) {
  UVC_TRY(check_tensors(a.w, D.L, "w", a.pe_in == nullptr));
  UVC_REQUIRE((a.x || a.pe_in) && a.logits && a.workspace, UVC_ERR_BAD_ARG, "vit_forward: NULL x / logits / workspace");
  const bool save = a.save_for_backward != 0;
  Ws16 w;
  carve16(D, save, a.workspace, a.workspace_bytes, &w);
  UVC_REQUIRE(w.bytes <= a.workspace_bytes, UVC_ERR_WORKSPACE, "vit_forward: workspace %llu bytes < required %llu",
              (unsigned long long)a.workspace_bytes, (unsigned long long)w.bytes);
  const int M = (int)D.M, C = D.C;
  const float eps = a.dims.ln_eps;
  const float scale = 1.0f / sqrtf((float)D.d);

  const uvc_vit_layout* lay = a.layout;
  UVC_TRY(check_layout(lay, D, a.skip_host));
  if (!a.weights_converted) UVC_TRY(convert_weights16(a.w, D, w, save, st, lay, a.skip_host));
  const float* pe = a.pe_in;
  if (!pe) {
    UVC_TRY(im2col16(a.x, w.cols, D.B, D.cin, D.img, D.patch, st, 1));
    float* pe_w = a.pe_out ? a.pe_out : w.pe;
    UVC_TRY(linear_fwd(w.cols, D.Kp, w.patch_w, a.w.patch_b, pe_w, C, D.B * D.np, C, D.Kp, st));
    pe = pe_w;
  } else if (a.pe_out && a.pe_out != pe) {
    cudaError_t e = cudaMemcpyAsync(a.pe_out, pe, (size_t)D.B * D.np * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_forward: memcpy pe_out: %s", cudaGetErrorString(e));
  }
  UVC_TRY(assemble_tokens(pe, a.w.cls_token, a.w.pos_embed, a.patch_scale, a.token_mask, w.tok, D.B, D.np, C, st));
  if (pe != w.pe && save) {
    cudaError_t e = cudaMemcpyAsync(w.pe, pe, (size_t)D.B * D.np * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_forward: memcpy pe: %s", cudaGetErrorString(e));
  }
  if (a.enable_jumping) {
    cudaError_t e = cudaMemsetAsync(w.accum, 0, (size_t)M * C * sizeof(float), st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_forward: memset accum: %s", cudaGetErrorString(e));
  }

  const float* x = w.tok;
  for (int l = 0; l < D.L; ++l) {
    const bool skipped = a.skip_host && a.skip_host[l];
    if (!skipped) {
      const uvc_block_tensors& p = a.w.blocks[l];
      Layer16& L = w.layer[l];
      // widths of this block: dense, or (Stage-2 layout) its live heads / neurons -- the operand copies were gathered to these widths above
      const BlockShape s = block_shape(D, lay, l);
      const float* qkv_b = (lay && p.qkv_b) ? w.qkv_bc[l] : p.qkv_b;
      const float* fc1_b = lay ? w.fc1_bc[l] : p.fc1_b;
      UVC_TRY(layernorm_fwd(x, C, p.norm1_w, p.norm1_b, eps, nullptr, C, L.mean1, L.rstd1, M, C, st, 0, L.ln1));
      UVC_TRY(linear16(L.ln1, C, w.qkv_w[l], qkv_b, nullptr, L.qkv, s.Ql, M, s.Ql, C, st));
      UVC_TRY(attention_fwd_f16(L.qkv, L.ctx, save ? L.lse : nullptr, D.B, s.Hl, D.ntok, scale, st));
      UVC_TRY(linear16(L.ctx, s.Cl, w.proj_w[l], p.proj_b, L.x1, nullptr, C, M, C, s.Cl, st, 0, nullptr, x, C));        // x1 = x + proj(ctx)
      UVC_TRY(layernorm_fwd(L.x1, C, p.norm2_w, p.norm2_b, eps, nullptr, C, L.mean2, L.rstd2, M, C, st, 0, L.ln2));
      UVC_TRY(linear16(L.ln2, C, w.fc1_w[l], fc1_b, nullptr, L.h, s.Fl, M, s.Fl, C, st, UVC_EPI_GELU, L.hpre));         // h = gelu(fc1) (fp16); hpre = gelu'(fc1) (fp16)
      if (a.blend) {
        // t = x1 + fc2(h) and the gate blend x <- d1 t + d0 x in ONE epilogue (t is kept only for the backward's gate gradient)
        UVC_TRY(linear16(L.h, s.Fl, w.fc2_w[l], p.fc2_b, L.xout, nullptr, C, M, C, s.Fl, st, 0, nullptr, L.x1, C, nullptr, nullptr, nullptr, a.blend + 2 * l, x,
                         save ? L.t : nullptr));
      } else {
        UVC_TRY(linear16(L.h, s.Fl, w.fc2_w[l], p.fc2_b, L.xout, nullptr, C, M, C, s.Fl, st, 0, nullptr, L.x1, C));
      }
      x = L.xout;
    }
    if (a.enable_jumping) UVC_TRY(scale_add(w.accum, x, nullptr, 1.0f, (long long)M * C, st));
  }
  const float* xf = a.enable_jumping ? w.accum : x;
  UVC_TRY(layernorm_fwd(xf, (long long)D.ntok * C, a.w.norm_w, a.w.norm_b, eps, w.cls_ln, C, w.mean_f, w.rstd_f, D.B, C, st, 1));
  UVC_TRY(linear_fwd(w.cls_ln, C, w.head_w, a.w.head_b, a.logits, D.NC, D.B, D.NC, C, st));
  return UVC_OK;
}

int vit_backward_f16(const uvc_vit_backward_args& a, const Dims& D, cudaStream_t st) {
  UVC_TRY(check_tensors(a.w, D.L, "w", a.d_pe == nullptr));
  UVC_TRY(check_tensors(a.g, D.L, "g", a.d_pe == nullptr));
  UVC_REQUIRE(a.dlogits && a.workspace, UVC_ERR_BAD_ARG, "vit_backward: NULL dlogits / workspace");
  UVC_REQUIRE(!a.blend || a.d_blend, UVC_ERR_BAD_ARG, "vit_backward: blend given without d_blend");
  Ws16 w;
  carve16(D, true, a.workspace, a.workspace_bytes, &w);
  UVC_REQUIRE(w.bytes <= a.workspace_bytes, UVC_ERR_WORKSPACE, "vit_backward: workspace %llu bytes < required %llu",
              (unsigned long long)a.workspace_bytes, (unsigned long long)w.bytes);
  const int M = (int)D.M, C = D.C;
  const float scale = 1.0f / sqrtf((float)D.d);
  const uvc_vit_layout* lay = a.layout;
  UVC_TRY(check_layout(lay, D, a.skip_host));
  const size_t xbytes = (size_t)M * C * sizeof(float);
  // loss scale of the fp16 gradient operands: device scalars {S, 1/S} (fixed by the caller, or the largest power of two with S max|dlogits| <= 128)
  const float* Sd = w.scales; const float* invSd = w.scales + 1;
  UVC_TRY(grad_scale(a.dlogits, (long long)D.B * D.NC, 128.0f, a.grad_scale, w.scales, st));

  const float* xin[UVC_MAX_DEPTH + 1];
  xin[0] = w.tok;
  for (int l = 0; l < D.L; ++l) xin[l + 1] = (a.skip_host && a.skip_host[l]) ? xin[l] : w.layer[l].xout;
  const float* xf = a.enable_jumping ? w.accum : xin[D.L];

  // head (TF32, unscaled): dW += dlogits^T cls_ln ; db += colsum ; dcls_ln = dlogits W
  const float* dlog = a.dlogits; long long ldl = D.NC;
  UVC_TRY(pad_dlogits(&dlog, &ldl, w.dlog_pad, D.B, D.NC, st));
  UVC_TRY(linear_wgrad(dlog, ldl, w.cls_ln, C, a.g.head_w, a.g.head_b, D.B, D.NC, C, st));
  UVC_TRY(linear_dgrad(dlog, ldl, w.head_w, w.dcls_ln, C, D.B, D.NC, C, st));
  // final LN backward on the cls rows; every other row of the stream gradient (and of its fp16 operand copy) is zero
  float* g = w.g_a;
  float* g_jump = nullptr;
  cudaError_t e = cudaMemsetAsync(g, 0, xbytes, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(w.g16, 0, xbytes / 2, st);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_backward: memset: %s", cudaGetErrorString(e));
  UVC_TRY(layernorm_bwd(w.dcls_ln, C, xf, (long long)D.ntok * C, w.mean_f, w.rstd_f, a.w.norm_w, nullptr, nullptr, nullptr, g,
                        (long long)D.ntok * C, a.g.norm_w, a.g.norm_b, D.B, C, st, nullptr, nullptr, nullptr, 1.0f, w.g16, 1.0f, w.scales));
  float* spare1 = w.g_b;
  float* spare2 = w.g_c;
  if (a.enable_jumping) {
    g_jump = w.accum;
    e = cudaMemcpyAsync(g_jump, g, xbytes, cudaMemcpyDeviceToDevice, st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_backward: memcpy: %s", cudaGetErrorString(e));
  }

  for (int l = D.L - 1; l >= 0; --l) {
    const bool skipped = a.skip_host && a.skip_host[l];
    if (!skipped) {
      const uvc_block_tensors& p = a.w.blocks[l];
      const uvc_block_tensors& gp = a.g.blocks[l];
      const Layer16& L = w.layer[l];
      const float* x = xin[l];
      const float* d = a.blend ? a.blend + 2 * l : nullptr;
      const float* d1 = d ? d + 1 : nullptr;
      // Stage-2 layout: the block ran at its live widths; its weight / bias gradients are formed compact in scratch (zeroed here, the split-K
      // GEMMs accumulate) and scattered into the dense gradient tensors at the end of the block.
      const BlockShape s = block_shape(D, lay, l);
      float* g_qkv_w = lay ? w.cg_qkv_w : gp.qkv_w; float* g_qkv_b = lay ? (gp.qkv_b ? w.cg_qkv_b : nullptr) : gp.qkv_b;
      float* g_proj_w = lay ? w.cg_proj_w : gp.proj_w;
      float* g_fc1_w = lay ? w.cg_fc1_w : gp.fc1_w; float* g_fc1_b = lay ? w.cg_fc1_b : gp.fc1_b;
      float* g_fc2_w = lay ? w.cg_fc2_w : gp.fc2_w;
      if (lay) {
        e = cudaMemsetAsync(w.cg_qkv_w, 0, w.cg_floats * sizeof(float), st);
        UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "vit_backward: memset (compact gradients): %s", cudaGetErrorString(e));
      }
      // ---- MLP:  t = x1 + fc2(gelu(fc1(ln2))); dt = d1 g is never materialised (d1 rides as a device-scalar alpha).  The gate gradients
      // dd1 = <g, t>, dd0 = <g, x> are folded into the two LayerNorm backward kernels below, which stream g (and x) anyway.
      UVC_TRY(linear_wgrad16(w.g16, C, L.h, s.Fl, g_fc2_w, M, C, s.Fl, invSd, st, d1));
      UVC_TRY(linear16(w.g16, C, w.fc2_wT[l], nullptr, nullptr, w.dh16, s.Fl, M, s.Fl, C, st, UVC_EPI_GELU_BWD, L.hpre, nullptr, 0, d1, g_fc1_b, invSd));   // dhpre (x S)
      UVC_TRY(linear_wgrad16(w.dh16, s.Fl, L.ln2, C, g_fc1_w, M, s.Fl, C, invSd, st));
      UVC_TRY(linear16(w.dh16, s.Fl, w.fc1_wT[l], nullptr, nullptr, w.dln16, C, M, C, s.Fl, st));                                                      // dln2 (x S)
      // dx1 = dt + LN2'(dln2); fc2.bias / proj.bias gradients ride along as column sums
      UVC_TRY(layernorm_bwd(nullptr, C, L.x1, C, L.mean2, L.rstd2, p.norm2_w, d ? nullptr : g, d ? g : nullptr, d1, spare2, C, gp.norm2_w, gp.norm2_b,
                            M, C, st, lay ? w.cg_fc2_b : gp.fc2_b, gp.proj_b, w.dln16, 1.0f, w.dx1_16, 1.0f, w.scales, d ? L.t : nullptr,
                            d ? a.d_blend + 2 * l : nullptr, 0));
      // compacted block: the masked fc2 columns still have the reference's (closed-form) gradient; this call's d fc2.bias joins the dense one
      if (lay) UVC_TRY(pruned_fc2_grad(gp.fc2_w, D.Fh, gp.fc2_b, w.cg_fc2_b, p.fc1_b, s.nidx + s.Fl, D.Fh - s.Fl, C, st));
      float* dx1 = spare2;
      // ---- attention:  x1 = x + proj(ctx)
      UVC_TRY(linear_wgrad16(w.dx1_16, C, L.ctx, s.Cl, g_proj_w, M, C, s.Cl, invSd, st));
      UVC_TRY(linear16(w.dx1_16, C, w.proj_wT[l], nullptr, nullptr, w.dctx16, s.Cl, M, s.Cl, C, st));                                                  // dctx (x S)
      UVC_TRY(attention_bwd_f16(L.qkv, L.lse, L.ctx, w.dctx16, w.Dv, w.dqkv16, D.B, s.Hl, D.ntok, scale, st, g_qkv_b, 1.0f, invSd));
      UVC_TRY(linear_wgrad16(w.dqkv16, s.Ql, L.ln1, C, g_qkv_w, M, s.Ql, C, invSd, st));
      UVC_TRY(linear16(w.dqkv16, s.Ql, w.qkv_wT[l], nullptr, nullptr, w.dln16, C, M, C, s.Ql, st));                                                    // dln1 (x S)
      // dx = dx1 + LN1'(dln1) + d0 g   (written over spare1); its fp16 operand copy replaces g16 (last read by the fc2 GEMMs above)
      UVC_TRY(layernorm_bwd(nullptr, C, x, C, L.mean1, L.rstd1, p.norm1_w, dx1, d ? g : nullptr, d, spare1, C, gp.norm1_w, gp.norm1_b, M, C, st,
                            nullptr, nullptr, w.dln16, 1.0f, w.g16, 1.0f, w.scales, nullptr, d ? a.d_blend + 2 * l : nullptr, d ? 1 : 0));
      if (lay) {
        const AxisMap none{nullptr, 0, 0, 0};
        const AxisMap qrows{s.hidx, s.Cl, D.d, C}, hcols{s.hidx, s.Cl, D.d, 0}, nmap{s.nidx, s.Fl, 1, 0};
        ScatterSeg seg[6]; int k = 0;
        seg[k++] = ScatterSeg{w.cg_qkv_w, gp.qkv_w, C, s.Ql, C, qrows, none};
        seg[k++] = ScatterSeg{w.cg_proj_w, gp.proj_w, C, C, s.Cl, none, hcols};
        seg[k++] = ScatterSeg{w.cg_fc1_w, gp.fc1_w, C, s.Fl, C, nmap, none};
        seg[k++] = ScatterSeg{w.cg_fc2_w, gp.fc2_w, D.Fh, C, s.Fl, none, nmap};
        if (gp.qkv_b) seg[k++] = ScatterSeg{w.cg_qkv_b, gp.qkv_b, 3 * C, 1, s.Ql, none, qrows};
        seg[k++] = ScatterSeg{w.cg_fc1_b, gp.fc1_b, D.Fh, 1, s.Fl, none, nmap};
        UVC_TRY(scatter_add(seg, k, st));
      }
      float* old = g; g = spare1; spare1 = old;
    }
    if (g_jump && l > 0) {
      UVC_TRY(scale_add(g, g_jump, nullptr, 1.0f, (long long)M * C, st));
      UVC_TRY(scale_to_f16(w.g16, g, 1.0f, (long long)M * C, st, Sd));
    }
  }

  const int rows = D.B * D.np;
  float* dpe = a.d_pe ? a.d_pe : w.dpe;
  UVC_TRY(assemble_tokens_bwd(g, w.pe, a.patch_scale, a.token_mask, dpe, a.patch_scale ? a.d_patch_scale : nullptr,
                              a.token_mask ? a.d_token_mask : nullptr, a.g.pos_embed, a.g.cls_token, D.B, D.np, C, st));
  if (!a.d_pe) UVC_TRY(linear_wgrad(w.dpe, C, w.cols, D.Kp, a.g.patch_w, a.g.patch_b, rows, C, D.Kp, st));
  return UVC_OK;
}

}  // namespace uvc

extern "C" uint64_t uvc_vit_workspace_bytes(const uvc_vit_dims* dims, int32_t save_for_backward) {
  if (!dims) { uvc::set_error("uvc_vit_workspace_bytes: dims is NULL"); return 0; }
  return uvc::vit_workspace_bytes(*dims, save_for_backward);
}
extern "C" int uvc_vit_forward(const uvc_vit_forward_args* args, void* stream) {
  if (!args) { uvc::set_error("uvc_vit_forward: args is NULL"); return UVC_ERR_BAD_ARG; }
  return uvc::vit_forward(*args, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_vit_backward(const uvc_vit_backward_args* args, void* stream) {
  if (!args) { uvc::set_error("uvc_vit_backward: args is NULL"); return UVC_ERR_BAD_ARG; }
  return uvc::vit_backward(*args, static_cast<cudaStream_t>(stream));
}

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
  float *g_a, *g_b, *g_c, *dh, *dqkv, *dP, *Dv, *dcls_ln, *dpe;
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
  } else {
    // inference: every layer reuses one set of buffers; the residual stream ping-pongs between t and xout
    LayerWs L0;
    L0.mean1 = L0.rstd1 = L0.mean2 = L0.rstd2 = nullptr;
    L0.ln1 = b.f(M * C); L0.qkv = b.f(M * 3 * C); L0.P = psz ? b.f(psz) : nullptr; L0.lse = nullptr; L0.ctx = b.f(M * C);
    L0.x1 = b.f(M * C); L0.ln2 = L0.ln1; L0.hpre = nullptr; L0.h = b.f(M * Fh);
    L0.t = b.f(M * C); L0.xout = b.f(M * C);
    float* ping = L0.xout; float* pong = b.f(M * C);
    for (int l = 0; l < D.L; ++l) { w->layer[l] = L0; w->layer[l].xout = (l & 1) ? pong : ping; }
    w->g_a = w->g_b = w->g_c = w->dh = w->dqkv = w->dP = w->Dv = w->dcls_ln = w->dpe = nullptr;
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

unsigned long long vit_workspace_bytes(const uvc_vit_dims& dims, int save) {
  Dims D;
  if (check_dims(dims, &D)) return 0;
  Ws w;
  carve(D, save != 0, nullptr, 0, &w);
  return w.bytes;
}

int vit_forward(const uvc_vit_forward_args& a, cudaStream_t st) {
  Dims D;
  UVC_TRY(check_dims(a.dims, &D));
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
  UVC_TRY(round_weights(a.w, D, w.wr, st));
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
  UVC_TRY(linear_wgrad(a.dlogits, D.NC, w.cls_ln, C, a.g.head_w, a.g.head_b, D.B, D.NC, C, st));
  UVC_TRY(linear_dgrad(a.dlogits, D.NC, w.wr.head_w, w.dcls_ln, C, D.B, D.NC, C, st));
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

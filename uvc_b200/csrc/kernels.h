// Internal (C++) launch functions shared between the .cu files of libuvc_sm100.so.  Each one validates
// its arguments, enqueues on `st` and returns a uvc_status; the extern "C" entry points in
// include/uvc_b200.h are thin wrappers over these, and vit_engine.cu composes them into the model.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace uvc {

int gemm_tf32(const uvc_gemm_args& a, cudaStream_t st);
int encode_tmap_4d(CUtensorMap* tm, const float* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                   const unsigned int box[4], bool atom32, const char* name);

int encode_tmap_4d_f16(CUtensorMap* tm, const void* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                       const unsigned int box[4], const char* name);

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, float* y, long long ldy, float* mean,
                  float* rstd, int M, int C, cudaStream_t st, int round_out = 0, void* y16 = nullptr);   // y16: write fp16 there instead of fp32 to y
int layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* mean, const float* rstd, const float* gamma,
                  const float* r1, const float* r2, const float* s2_dev, float* dx, long long lddx, float* dgamma, float* dbeta, int M, int C,
                  cudaStream_t st, float* cs_r1 = nullptr, float* cs_out = nullptr,    // cs_*: fused column sums of r1 / of dx (bias gradients)
                  const void* dy16 = nullptr, float dy_scale = 1.0f,                    // dy16: dy as fp16 (row stride lddy), multiplied by dy_scale on load
                  void* dx16 = nullptr, float out_scale = 1.0f,                         // dx16: also write fp16(out_scale * dx), row stride lddx
                  const float* scales_dev = nullptr,                                    // device {S, 1/S}: out_scale *= S, dy_scale *= 1/S
                  const float* dot_t = nullptr, float* dots = nullptr, int dot_x = 0);  // gate gradients: dots[0] += <r2, x> (dot_x), dots[1] += <r2, dot_t>
int softmax_fwd(float* S, long long ld, long long rows, int n, cudaStream_t st, int round_out = 0);
int softmax_bwd(const float* P, float* dP, long long ld, long long rows, int n, float scale, cudaStream_t st, int round_out = 0);
int colsum(const float* X, long long ld, int M, int N, const float* scale_dev, float* out, cudaStream_t st);
int blend_fwd(const float* t, const float* x, const float* d, float* out, long long n, cudaStream_t st);
int blend_dots(const float* g, const float* t, const float* x, float* dots, long long n, cudaStream_t st);
int im2col16(const float* x, float* out, int B, int Cin, int HW, int P, cudaStream_t st, int round_out = 0);
constexpr int kMaxRoundSegs = 96;
int round_tf32_segs(const float* const* src, float* const* dst, const long long* n, int nseg, cudaStream_t st);
// fp32 -> fp16 copies (dst, [rows, cols]) and optional transposed copies (dstT, [cols, rows]) of up to nseg weight matrices
int cvt_f16_segs(const float* const* src, void* const* dst, void* const* dstT, const int* rows, const int* cols, int nseg, cudaStream_t st);
int assemble_tokens(const float* pe, const float* cls, const float* pos, const float* pscale, const float* tmask, float* tok, int B, int np, int C,
                    cudaStream_t st);
int assemble_tokens_bwd(const float* g, const float* pe, const float* pscale, const float* tmask, float* dpe, float* dscale, float* dtmask,
                        float* dpos, float* dcls, int B, int np, int C, cudaStream_t st);
int scale_add(float* y, const float* x, const float* s_dev, float s, long long n, cudaStream_t st);
int scale_to_f16(void* dst16, const float* src, float s, long long n, cudaStream_t st, const float* s_dev = nullptr);   // dst16 = fp16(s * (*s_dev) * src)
int grad_scale(const float* dlogits, long long n, float target, float fixed, float* scales, cudaStream_t st);   // scales = {S, 1/S} (see rowwise.cu)

int attn_ldp(int N);
int attention_fwd(const float* qkv, float* P, float* ctx, int B, int H, int N, int d, float scale, cudaStream_t st, bool need_P = true, float* lse = nullptr);
bool attn_fused_ok(int N, int d);     // the fused tcgen05 attention kernels cover d == 64, N <= 208
int attention_bwd_fused(const float* qkv, const float* lse, const float* ctx, const float* dctx, float* Dv, float* dqkv, int B, int H, int N, float scale,
                        cudaStream_t st, float* dqkv_bias = nullptr);   // dqkv_bias [3*H*64]: += column sums of dqkv (the qkv bias gradient)
int attention_bwd(const float* qkv, const float* P, const float* dctx, float* dP, float* dqkv, int B, int H, int N, int d, float scale,
                  cudaStream_t st);

// fp16-operand attention (attention_f16.cu): qkv16 [B*N, 3*H*64] fp16 in, ctx16 [B*N, H*64] fp16 out, lse [B,H,N] fp32 (may be NULL)
bool attn_f16_ok(int N, int d);
int attention_fwd_f16(const void* qkv16, void* ctx16, float* lse, int B, int H, int N, float scale, cudaStream_t st);
// dqkv16 [B*N, 3*H*64] fp16 from dctx16 (fp16, may carry a loss scale: everything downstream is linear in it); Dv: [B,H,N] fp32 scratch;
// dqkv_bias (optional, fp32 [3*H*64]) += db_scale * column sums of dqkv
int attention_bwd_f16(const void* qkv16, const float* lse, const void* ctx16, const void* dctx16, float* Dv, void* dqkv16, int B, int H, int N,
                      float scale, cudaStream_t st, float* dqkv_bias = nullptr, float db_scale = 1.0f, const float* db_scale_dev = nullptr);

// Stage-2 compaction (compact.cu).  AxisMap sends a compact row / column number to its dense one: the axis is `nsect` sections of
// `per` = n_live * group compact entries; entry w of a section is member (w % group) of live group idx[w / group]; dense sections are
// sect_stride apart.  (heads: group 64, qkv rows: 3 sections C apart; neurons: group 1.)  idx == NULL: identity.
struct AxisMap { const int* idx; int per, group, sect_stride; };
struct GatherSeg {            // dst[r, c] = src[rmap(r) * src_ld + cmap(c)] for a compact [rows, cols] matrix; any of the three outputs may be NULL
  const float* src; long long src_ld;
  __half* dst16; __half* dstT16; float* dst32;
  int rows, cols;
  AxisMap rmap, cmap;
};
struct ScatterSeg {           // dst[rmap(r) * dst_ld + cmap(c)] += src[r, c]
  const float* src; float* dst; long long dst_ld;
  int rows, cols;
  AxisMap rmap, cmap;
};
int gather_cvt(const GatherSeg* segs, int nseg, cudaStream_t st);
int scatter_add(const ScatterSeg* segs, int nseg, cudaStream_t st);
// dW[:, dead[j]] += gelu(fc1_b[dead[j]]) * db_call[:] ; db += db_call   (closed-form gradient of the masked fc2 columns, see compact.cu)
int pruned_fc2_grad(float* dW, long long ldw, float* db, const float* db_call, const float* fc1_b, const int* dead, int n_dead, int C, cudaStream_t st);

// timm batch-mode mixup / cutmix + mixed smoothed targets (input_pipeline.cu)
int mixup_batch(float* x, const long long* y, float* targets, int B, int C, int H, int W, int NC, float lam, float smoothing, int use_cutmix, int yl, int yh,
                int xl, int xh, cudaStream_t st);
// token slimming gate (token_gate.cu)
int token_gate_fold(const float* patch_w, const float* patch_b, const float* gate_w, int C, int Kp, float* v, float* c1, cudaStream_t st);
int token_gate_fwd(const float* F, long long ldf, int Kf, const float* v, const float* c1, const float* gate_b, const float* pscale, const float* noise,
                   float tau, int k, int B, int np, float* mask, float* ysoft, float* ls, float* scores, cudaStream_t st);
int token_gate_bwd(const float* dmask, const float* ysoft, const float* ls, float tau, int B, int np, float* dscores, cudaStream_t st);
int token_gate_apply(const float* dscores, const float* x, const float* gate_w, const float* pscale, int B, int np, int C, float* dx, float* d_gate_w,
                     float* d_gate_b, float* d_pscale, cudaStream_t st);

// ---- helpers to describe GEMM operands tersely
inline uvc_operand op_k(const float* p, long long ld, long long bs1 = 0, long long bs2 = 0) { return uvc_operand{p, ld, bs1, bs2, 0, 0}; }
inline uvc_operand op_mn(const float* p, long long ld, long long bs1 = 0, long long bs2 = 0) { return uvc_operand{p, ld, bs1, bs2, 1, 0}; }
inline uvc_gemm_args gemm_args(int M, int N, int K, uvc_operand A, uvc_operand B, float* D, long long ldd) {
  uvc_gemm_args a{};
  a.M = M; a.N = N; a.K = K; a.nb1 = 1; a.nb2 = 1; a.splits = 1;
  a.A = A; a.B = B; a.D = D; a.ldd = ldd;
  a.alpha = 1.0f; a.beta = 1.0f;
  return a;
}
// split-K factor for a weight-gradient GEMM (few output tiles, very long K): fill ~2 waves of 148 SMs
inline int wgrad_splits(int M, int N, int K, int bk = 32) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  int s = (2 * 148) / tiles;            // floor: tiles * s CTAs must fit ONE wave of 2 CTAs per SM (a 297th CTA costs a whole second wave)
  const int nkb = (K + bk - 1) / bk;
  if (s > nkb / 4) s = nkb / 4;
  if (s < 1) s = 1;
  return s;
}

}  // namespace uvc

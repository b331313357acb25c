// Attention core softmax(Q K^T * scale) V per (image, head), forward and backward
// (models/model_distilled.py:175-185 and its autograd backward).
//
// Round-1 composition: batched tcgen05 TF32 GEMMs (gemm_tf32.cu) reading Q/K/V straight out of the
// [B*N, 3*H*d] qkv buffer through strided TMA descriptors (no permute/copy kernels), a warp-per-row
// softmax, and the PV GEMM writing ctx directly in [B*N, H*d] layout.  P is kept ([B,H,N,ldp]) for the
// backward, as the reference's autograd does.
#include "kernels.h"

namespace uvc {

int attn_ldp(int N) { return (N + 3) / 4 * 4; }

static int check_attn(int B, int H, int N, int d) {
  UVC_REQUIRE(B > 0 && H > 0 && N > 0 && d > 0, UVC_ERR_BAD_SHAPE, "attention: bad dims B=%d H=%d N=%d d=%d", B, H, N, d);
  UVC_REQUIRE(N <= 256, UVC_ERR_BAD_SHAPE, "attention: N=%d tokens > 256 is not supported", N);
  UVC_REQUIRE((d & 3) == 0, UVC_ERR_BAD_SHAPE, "attention: head dim %d must be a multiple of 4", d);
  return UVC_OK;
}

int attention_fwd(const float* qkv, float* P, float* ctx, int B, int H, int N, int d, float scale, cudaStream_t st) {
  int rc = check_attn(B, H, N, d);
  if (rc) return rc;
  const long long C = (long long)H * d, ld3 = 3 * C, ldp = attn_ldp(N);
  const float* q = qkv; const float* k = qkv + C; const float* v = qkv + 2 * C;
  // S[b,h] = scale * Q K^T
  uvc_gemm_args a = gemm_args(N, N, d, op_k(q, ld3, d, (long long)N * ld3), op_k(k, ld3, d, (long long)N * ld3), P, ldp);
  a.nb1 = H; a.nb2 = B; a.d_bs1 = (long long)N * ldp; a.d_bs2 = (long long)H * N * ldp; a.alpha = scale;
  if ((rc = gemm_tf32(a, st))) return rc;
  if ((rc = softmax_fwd(P, ldp, (long long)B * H * N, N, st, 1))) return rc;     // P only feeds GEMMs: rounded to TF32
  // ctx[b, :, h*d:(h+1)*d] = P[b,h] V[b,h]        (V is [k=N rows][n=d cols] in memory -> MN-major B operand)
  uvc_gemm_args c = gemm_args(N, d, N, op_k(P, ldp, (long long)N * ldp, (long long)H * N * ldp), op_mn(v, ld3, d, (long long)N * ld3), ctx, C);
  c.nb1 = H; c.nb2 = B; c.d_bs1 = d; c.d_bs2 = (long long)N * C;
  c.flags = UVC_EPI_ROUND_TF32;
  return gemm_tf32(c, st);
}

int attention_bwd(const float* qkv, const float* P, const float* dctx, float* dP, float* dqkv, int B, int H, int N, int d, float scale,
                  cudaStream_t st) {
  int rc = check_attn(B, H, N, d);
  if (rc) return rc;
  const long long C = (long long)H * d, ld3 = 3 * C, ldp = attn_ldp(N);
  const long long pb1 = (long long)N * ldp, pb2 = (long long)H * N * ldp, qb2 = (long long)N * ld3, cb2 = (long long)N * C;
  const float* q = qkv; const float* k = qkv + C; const float* v = qkv + 2 * C;
  float* dq = dqkv; float* dk = dqkv + C; float* dv = dqkv + 2 * C;
  // dP = dctx V^T
  uvc_gemm_args a = gemm_args(N, N, d, op_k(dctx, C, d, cb2), op_k(v, ld3, d, qb2), dP, ldp);
  a.nb1 = H; a.nb2 = B; a.d_bs1 = pb1; a.d_bs2 = pb2;
  if ((rc = gemm_tf32(a, st))) return rc;
  // dV = P^T dctx    (A = P^T: memory rows index K -> MN-major; B = dctx^T likewise)
  uvc_gemm_args b = gemm_args(N, d, N, op_mn(P, ldp, pb1, pb2), op_mn(dctx, C, d, cb2), dv, ld3);
  b.nb1 = H; b.nb2 = B; b.d_bs1 = d; b.d_bs2 = qb2; b.flags = UVC_EPI_ROUND_TF32;
  if ((rc = gemm_tf32(b, st))) return rc;
  // dS = scale * P .* (dP - rowsum(dP .* P))
  if ((rc = softmax_bwd(P, dP, ldp, (long long)B * H * N, N, scale, st, 1))) return rc;
  // dQ = dS K
  uvc_gemm_args c = gemm_args(N, d, N, op_k(dP, ldp, pb1, pb2), op_mn(k, ld3, d, qb2), dq, ld3);
  c.nb1 = H; c.nb2 = B; c.d_bs1 = d; c.d_bs2 = qb2; c.flags = UVC_EPI_ROUND_TF32;
  if ((rc = gemm_tf32(c, st))) return rc;
  // dK = dS^T Q
  uvc_gemm_args e = gemm_args(N, d, N, op_mn(dP, ldp, pb1, pb2), op_mn(q, ld3, d, qb2), dk, ld3);
  e.nb1 = H; e.nb2 = B; e.d_bs1 = d; e.d_bs2 = qb2; e.flags = UVC_EPI_ROUND_TF32;
  return gemm_tf32(e, st);
}

}  // namespace uvc

extern "C" int32_t uvc_attn_ldp(int32_t N) { return uvc::attn_ldp(N); }
extern "C" int uvc_attention_fwd(const float* qkv, float* P, float* ctx, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv && P && ctx, UVC_ERR_BAD_ARG, "uvc_attention_fwd: NULL pointer");
  return uvc::attention_fwd(qkv, P, ctx, B, H, N, d, scale, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_attention_bwd(const float* qkv, const float* P, const float* dctx, float* dP, float* dqkv, int32_t B, int32_t H, int32_t N,
                                 int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv && P && dctx && dP && dqkv, UVC_ERR_BAD_ARG, "uvc_attention_bwd: NULL pointer");
  return uvc::attention_bwd(qkv, P, dctx, dP, dqkv, B, H, N, d, scale, static_cast<cudaStream_t>(stream));
}

// Token (patch) slimming gate, mode 2 of --enable_patch_gating: reference models/model_distilled.py:446-456 with gumbel_softmax :36-63
// and scatter :21-33.
//
//   scores[b,p] = Linear(C,1)(x[b,p,:])                      x = patch embeddings (times the mode-1 patch gate if that is on too)
//   l           = log_softmax(scores[b,:])
//   y           = softmax((l + gumbel[b,:]) / tau)            gumbel noise is supplied by the caller (torch's generator, the reference's stream)
//   hard        = one-hot of the k largest y                  (k = int(ratio * 196); reference: topk + a .tolist() host round trip of B*k indices)
//   mask        = (hard - y) + y ;  mask[:,0] = 1             straight-through value in the reference's float arithmetic
//
// One CTA per image, no host round trip, bit-exact selection arithmetic in fp32:
//   * the score is NOT taken from the tensor-core patch embedding (10-bit operands: ~1e-3 relative, enough to flip a top-k boundary in a batch);
//     it is an fp32 CUDA-core dot product straight from the rows the patch GEMM reads:  scores = pscale * (F v + c1) + b_gate with
//     v = W_patch^T w_gate, c1 = b_patch . w_gate folded once per forward (DeiT: F = im2col rows, 768 wide), or F = tokens, v = w_gate (T2T);
//   * top-k by exact rank inside the CTA: rank(p) = #{j : y_j > y_p or (y_j == y_p and j < p)}, hard = rank < k -- ties go to the lower index,
//     which is what torch.topk does on CUDA (radix select, then ties in index order) where the reference runs this op.
// HBM-bound: B * np * Kf * 4 bytes read once (DeiT-Small, B = 128: 77 MB, ~12 us).
#include "kernels.h"

namespace uvc {

namespace {

constexpr int kMaxNp = 256;

// v[k] = sum_c W[c, k] * wg[c] ; c1 = sum_c b[c] * wg[c]          (W: [C, Kp] row-major)
__global__ void __launch_bounds__(256) token_gate_fold_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ wg,
                                                              int C, int Kp, float* __restrict__ v, float* __restrict__ c1) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < Kp) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(__ldg(W + (long long)c * Kp + k), __ldg(wg + c), a);
    v[k] = a;
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    float a = 0.f;
    for (int c = threadIdx.x; c < C; c += 32) a = fmaf(__ldg(b + c), __ldg(wg + c), a);
    a = warp_sum(a);
    if (threadIdx.x == 0) *c1 = a;
  }
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) token_gate_fwd_kernel(const float* __restrict__ F, long long ldf, int Kf, const float* __restrict__ v,
                                                             const float* __restrict__ c1, const float* __restrict__ gate_b,
                                                             const float* __restrict__ pscale, const float* __restrict__ noise, float tau, int k,
                                                             int np, float* __restrict__ mask, float* __restrict__ ysoft, float* __restrict__ ls_out,
                                                             float* __restrict__ scores_out) {
  __shared__ float sc[kMaxNp];
  __shared__ float ys[kMaxNp];
  __shared__ float red[8];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, p = threadIdx.x;
  const float cc = c1 ? __ldg(c1) : 0.f, bg = gate_b ? __ldg(gate_b) : 0.f;
  const int nv = Kf >> 2;
  const float4* v4 = reinterpret_cast<const float4*>(v);
  for (int r = warp; r < np; r += 8) {
    const float4* f4 = reinterpret_cast<const float4*>(F + ((long long)b * np + r) * ldf);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int i = lane; i < nv; i += 32) {
      const float4 x = f4[i], w = __ldg(v4 + i);
      a0 = fmaf(x.x, w.x, a0); a1 = fmaf(x.y, w.y, a1); a2 = fmaf(x.z, w.z, a2); a3 = fmaf(x.w, w.w, a3);
    }
    const float dot = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) sc[r] = (pscale ? __ldg(pscale + r) : 1.0f) * (dot + cc) + bg;
  }
  __syncthreads();
  const bool on = p < np;
  const float s = on ? sc[p] : -INFINITY;
  // l = log_softmax(scores)
  const float m1 = block_reduce_max(s, red);
  const float se = block_reduce_sum(on ? expf(s - m1) : 0.f, red);
  const float l = s - m1 - logf(se);
  // y = softmax((l + gumbel) / tau)
  const float z = on ? (l + noise[(long long)b * np + p]) / tau : -INFINITY;
  const float m2 = block_reduce_max(z, red);
  const float e = on ? expf(z - m2) : 0.f;
  const float tot = block_reduce_sum(e, red);
  const float y = e / tot;
  if (on) ys[p] = y;
  __syncthreads();
  if (on) {
    int rank = 0;
    for (int j = 0; j < np; ++j) {
      const float yj = ys[j];
      rank += (yj > y || (yj == y && j < p)) ? 1 : 0;
    }
    const float hard = rank < k ? 1.0f : 0.0f;
    float mk = (hard - y) + y;                 // the straight-through value exactly as the reference forms it (y_hard - y_soft.detach() + y_soft)
    if (p == 0) mk = 1.0f;                     // token_mask[:, 0] = 1.
    const long long o = (long long)b * np + p;
    mask[o] = mk;
    if (ysoft) ysoft[o] = y;
    if (ls_out) ls_out[o] = l;
    if (scores_out) scores_out[o] = s;
  }
}

// d(scores) from d(mask): mask = y (+ constants), mask[:,0] overwritten  ->  dy = dmask (0 at p = 0);  dz = y (dy - <dy, y>);  dl = dz / tau;
// dscores = dl - exp(l) sum(dl)
__global__ void __launch_bounds__(256) token_gate_bwd_kernel(const float* __restrict__ dmask, const float* __restrict__ ysoft, const float* __restrict__ ls,
                                                             float tau, int np, float* __restrict__ dscores) {
  __shared__ float red[8];
  const int b = blockIdx.x, p = threadIdx.x;
  const bool on = p < np;
  const long long o = (long long)b * np + p;
  const float y = on ? ysoft[o] : 0.f;
  const float dy = (on && p > 0) ? dmask[o] : 0.f;
  const float dot = block_reduce_sum(dy * y, red);
  const float dl = y * (dy - dot) / tau;
  const float sdl = block_reduce_sum(dl, red);
  if (on) dscores[o] = dl - expf(ls[o]) * sdl;
}

// scores = pscale[p] * (x[b,p,:] . wg) + ...  ->  dx[b,p,:] += ds * pscale[p] * wg ;  dwg += ds * pscale[p] * x[b,p,:] ;  dpscale[p] += ds * (x . wg) ;
// dbg += ds.  One warp per token row; the per-lane dwg partials are combined through shared memory, one atomic per column and block.
template <int NV>
__global__ void __launch_bounds__(256) token_gate_apply_kernel(const float* __restrict__ ds, const float* __restrict__ x, const float* __restrict__ wg,
                                                               const float* __restrict__ pscale, int rows, int np, int C, int rows_per_block,
                                                               float* __restrict__ dx, float* __restrict__ dwg, float* __restrict__ dbg,
                                                               float* __restrict__ dpscale) {
  __shared__ float red[8][32 * 4 + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 2;
  float4 acc[NV], w[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i] = make_float4(0, 0, 0, 0);
    const int c = lane + i * 32;
    w[i] = c < nv ? __ldg(reinterpret_cast<const float4*>(wg) + c) : make_float4(0, 0, 0, 0);
  }
  float sb = 0.f;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int r = r0 + warp; r < r1; r += 8) {
    const int p = r % np;
    const float d = ds[r], ps = pscale ? __ldg(pscale + p) : 1.0f;
    const float dps = d * ps;
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)r * C);
    float4* dr = dx ? reinterpret_cast<float4*>(dx + (long long)r * C) : nullptr;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const float4 xv = xr[c];
        dot += (xv.x * w[i].x + xv.y * w[i].y) + (xv.z * w[i].z + xv.w * w[i].w);
        acc[i].x = fmaf(dps, xv.x, acc[i].x); acc[i].y = fmaf(dps, xv.y, acc[i].y); acc[i].z = fmaf(dps, xv.z, acc[i].z); acc[i].w = fmaf(dps, xv.w, acc[i].w);
        if (dr) {
          float4 g = dr[c];
          g.x = fmaf(dps, w[i].x, g.x); g.y = fmaf(dps, w[i].y, g.y); g.z = fmaf(dps, w[i].z, g.z); g.w = fmaf(dps, w[i].w, g.w);
          dr[c] = g;
        }
      }
    }
    if (dpscale) {
      dot = warp_sum(dot);
      if (lane == 0) atomicAdd(dpscale + p, d * dot);
    }
    if (lane == 0) sb += d;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 >= nv) break;
    __syncthreads();
    float* q = &red[warp][lane * 4];
    q[0] = acc[i].x; q[1] = acc[i].y; q[2] = acc[i].z; q[3] = acc[i].w;
    __syncthreads();
    if (threadIdx.x < 128) {
      const int col = i * 128 + threadIdx.x;
      if (col < C) {
        float a = 0.f;
        for (int wv = 0; wv < 8; ++wv) a += red[wv][threadIdx.x];
        atomicAdd(dwg + col, a);
      }
    }
  }
  if (dbg) {
    __syncthreads();
    if (lane == 0) red[warp][0] = sb;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f;
      for (int wv = 0; wv < 8; ++wv) a += red[wv][0];
      atomicAdd(dbg, a);
    }
  }
}

}  // namespace

int token_gate_fold(const float* patch_w, const float* patch_b, const float* gate_w, int C, int Kp, float* v, float* c1, cudaStream_t st) {
  token_gate_fold_kernel<<<(Kp + 255) / 256, 256, 0, st>>>(patch_w, patch_b, gate_w, C, Kp, v, c1);
  return check_launch("token_gate_fold");
}

int token_gate_fwd(const float* F, long long ldf, int Kf, const float* v, const float* c1, const float* gate_b, const float* pscale, const float* noise,
                   float tau, int k, int B, int np, float* mask, float* ysoft, float* ls, float* scores, cudaStream_t st) {
  UVC_REQUIRE(np >= 1 && np <= kMaxNp, UVC_ERR_BAD_SHAPE, "token_gate: %d tokens per image unsupported (1..%d)", np, kMaxNp);
  UVC_REQUIRE(Kf > 0 && (Kf & 3) == 0 && (ldf & 3) == 0 && (reinterpret_cast<uintptr_t>(F) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
              UVC_ERR_BAD_SHAPE, "token_gate: feature rows must be 16 B aligned with a width that is a multiple of 4");
  UVC_REQUIRE(tau > 0.f && k >= 0 && k <= np, UVC_ERR_BAD_ARG, "token_gate: need tau > 0 and 0 <= k <= np (tau=%g, k=%d, np=%d)", (double)tau, k, np);
  if (B <= 0) return UVC_OK;
  token_gate_fwd_kernel<<<B, 256, 0, st>>>(F, ldf, Kf, v, c1, gate_b, pscale, noise, tau, k, np, mask, ysoft, ls, scores);
  return check_launch("token_gate_fwd");
}

int token_gate_bwd(const float* dmask, const float* ysoft, const float* ls, float tau, int B, int np, float* dscores, cudaStream_t st) {
  UVC_REQUIRE(np >= 1 && np <= kMaxNp && tau > 0.f, UVC_ERR_BAD_SHAPE, "token_gate_bwd: bad np / tau");
  if (B <= 0) return UVC_OK;
  token_gate_bwd_kernel<<<B, 256, 0, st>>>(dmask, ysoft, ls, tau, np, dscores);
  return check_launch("token_gate_bwd");
}

int token_gate_apply(const float* dscores, const float* x, const float* gate_w, const float* pscale, int B, int np, int C, float* dx, float* d_gate_w,
                     float* d_gate_b, float* d_pscale, cudaStream_t st) {
  UVC_REQUIRE(C > 0 && (C & 3) == 0 && C <= 1024, UVC_ERR_BAD_SHAPE, "token_gate_apply: C=%d must be a multiple of 4 and <= 1024", C);
  const int rows = B * np;
  if (rows <= 0) return UVC_OK;
  int blocks = 148 * 4;
  int rpb = (rows + blocks - 1) / blocks;
  if (rpb < 8) rpb = 8;
  blocks = (rows + rpb - 1) / rpb;
  if (C <= 256) token_gate_apply_kernel<2><<<blocks, 256, 0, st>>>(dscores, x, gate_w, pscale, rows, np, C, rpb, dx, d_gate_w, d_gate_b, d_pscale);
  else if (C <= 384) token_gate_apply_kernel<3><<<blocks, 256, 0, st>>>(dscores, x, gate_w, pscale, rows, np, C, rpb, dx, d_gate_w, d_gate_b, d_pscale);
  else if (C <= 768) token_gate_apply_kernel<6><<<blocks, 256, 0, st>>>(dscores, x, gate_w, pscale, rows, np, C, rpb, dx, d_gate_w, d_gate_b, d_pscale);
  else token_gate_apply_kernel<8><<<blocks, 256, 0, st>>>(dscores, x, gate_w, pscale, rows, np, C, rpb, dx, d_gate_w, d_gate_b, d_pscale);
  return check_launch("token_gate_apply");
}

}  // namespace uvc

#define UVC_ST static_cast<cudaStream_t>(stream)
extern "C" {
int uvc_token_gate_fold(const float* patch_w, const float* patch_b, const float* gate_w, int32_t C, int32_t Kp, float* v, float* c1, void* stream) {
  UVC_REQUIRE(patch_w && patch_b && gate_w && v && c1, UVC_ERR_BAD_ARG, "uvc_token_gate_fold: NULL pointer");
  return uvc::token_gate_fold(patch_w, patch_b, gate_w, C, Kp, v, c1, UVC_ST);
}
int uvc_token_gate_fwd(const float* feat, int64_t ldf, int32_t Kf, const float* v, const float* c1, const float* gate_b, const float* pscale,
                       const float* noise, float tau, int32_t k, int32_t B, int32_t np, float* mask, float* ysoft, float* ls, float* scores, void* stream) {
  UVC_REQUIRE(feat && v && noise && mask, UVC_ERR_BAD_ARG, "uvc_token_gate_fwd: NULL pointer");
  return uvc::token_gate_fwd(feat, ldf, Kf, v, c1, gate_b, pscale, noise, tau, k, B, np, mask, ysoft, ls, scores, UVC_ST);
}
int uvc_token_gate_bwd(const float* dmask, const float* ysoft, const float* ls, float tau, int32_t B, int32_t np, float* dscores, void* stream) {
  UVC_REQUIRE(dmask && ysoft && ls && dscores, UVC_ERR_BAD_ARG, "uvc_token_gate_bwd: NULL pointer");
  return uvc::token_gate_bwd(dmask, ysoft, ls, tau, B, np, dscores, UVC_ST);
}
int uvc_token_gate_apply(const float* dscores, const float* x, const float* gate_w, const float* pscale, int32_t B, int32_t np, int32_t C, float* dx,
                         float* d_gate_w, float* d_gate_b, float* d_pscale, void* stream) {
  UVC_REQUIRE(dscores && x && gate_w && d_gate_w, UVC_ERR_BAD_ARG, "uvc_token_gate_apply: NULL pointer");
  return uvc::token_gate_apply(dscores, x, gate_w, pscale, B, np, C, dx, d_gate_w, d_gate_b, d_pscale, UVC_ST);
}
}  // extern "C"

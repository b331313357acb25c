// Tokens-to-token front end of T2T-ViT-14 as ONE C-ABI call per pass.
// Reference: UVC/T2TViT/models/t2t_vit.py:46-105 (T2T_module, tokens_type='performer') and token_performer.py:8-69:
//
//   x [B,3,H,W] --Unfold 7x7/4 pad 2--> [B, T0, 147] --Token_performer(147 -> 64)--> [B, T0, 64] = image [B,64,H/4,W/4]
//     --Unfold 3x3/2 pad 1--> [B, T1, 576] --Token_performer(576 -> 64)--> [B, T1, 64] --Unfold 3x3/2 pad 1--> [B, T2, 576] --Linear--> [B, T2, C]
//   Token_performer: xn = LN(x); k,q,v = split(kqv(xn)); kp = prm_exp(k), qp = prm_exp(q) (exp(w^T x - |x|^2/2)/sqrt(m), m = 32);
//     y = (qp kptv^T) / (qp . sum_t kp + 1e-8), kptv = v^T kp per image; y = v + drop(proj(y)); out = y + drop(mlp(LN2(y)))
//
// The front end is ~6 % of the model's FLOPs but was 65 % of the T2T step while it ran as ~150 eager torch launches (im2col / col2im, fp32 SIMT
// GEMMs, LayerNorm over 147- and 576-wide rows).  It is memory- and launch-bound work, laid out here for HBM:
//   * unfold_ln_kernel    -- soft split + LayerNorm in one pass: a warp gathers its token's k x k x Cin patch (NCHW image or the previous stage's
//                            token-major [B, T, 64] = NHWC), normalises it in shared memory and writes the fp16 A operand of the kqv GEMM.  The
//                            unfolded fp32 tensor (236 MB for B = 128) is never materialised.
//   * kqv / proj / mlp / project: the tcgen05 GEMMs of gemm_tf32.cu (fp16 operands, fp32 accumulation, fused bias / GELU / residual epilogues).
//   * performer_kv / performer_out -- the linear attention on CUDA cores (64 x 32 contractions per token): pass 1 reduces kptv [64,32] and
//                            sum_t kp [32] per image, pass 2 forms y.  The random features kp / qp are recomputed, never stored.
//   * backward: performer_bwd_a (dq, per-image d kptv / d ksum) and performer_bwd_b (dk, dv), unfold_ln_bwd_kernel (LayerNorm backward + the
//     fold (col2im) as vector atomics into the previous stage's token gradient), weight gradients on the split-K GEMM.
// fp16 gradient operands carry a power-of-two loss scale chosen on the device from max|d tokens| (as the block engine does).
#include "kernels.h"

#include <cuda_fp16.h>

namespace uvc {

namespace {

typedef __half h16;
constexpr int kE = 64;        // token_dim (T2T_module token_dim = 64)
constexpr int kF = 32;        // random features m = emb * kernel_ratio
constexpr int kKqv = 3 * kE;

#define UVC_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// ------------------------------------------------------------------------------------------------ soft split (+ LayerNorm)
struct Geom {
  int nchw;              // 1: src is the image [B, Cin, Hs, Ws]; 0: src is token-major [B, Hs*Ws, 64]
  int Cin, Hs, Ws, k, stride, pad, Ho, Wo;
  int dim, Kp;           // dim = Cin * k * k (nn.Unfold order: c * k*k + ky * k + kx); Kp = dim rounded up to 8 (fp16 row stride)
};

// NCHW gather table (one per block, in shared memory): element e = c k^2 + ky k + kx reads src at (c Hs + ky) Ws + kx relative to the patch
// origin; the low 16 bits carry (ky, kx) for the border test.  Saves two integer divisions per gathered element (the kernel was issue-bound).
__device__ __forceinline__ void build_tab(const Geom& g, int* tab) {
  if (!g.nchw) return;
  const int k2 = g.k * g.k;
  for (int e = threadIdx.x; e < g.dim; e += blockDim.x) {
    const int c = e / k2, r = e - c * k2, ky = r / g.k, kx = r - ky * g.k;
    tab[e] = (((c * g.Hs + ky) * g.Ws + kx) << 8) | (ky << 4) | kx;        // k <= 7: 4 bits each; offset < 2^23
  }
}
// fills rb[0..dim) with the patch of output token (b, oy, ox)
__device__ __forceinline__ void gather_patch(const float* __restrict__ src, const Geom& g, const int* tab, int b, int oy, int ox, int lane, float* rb) {
  const int k2 = g.k * g.k;
  const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
  if (g.nchw) {
    const float* base = src + (long long)b * g.Cin * g.Hs * g.Ws + (long long)iy0 * g.Ws + ix0;
    for (int e = lane; e < g.dim; e += 32) {
      const int t = tab[e];
      const int iy = iy0 + ((t >> 4) & 15), ix = ix0 + (t & 15);
      rb[e] = (iy >= 0 && iy < g.Hs && ix >= 0 && ix < g.Ws) ? __ldg(base + (t >> 8)) : 0.f;
    }
  } else {               // 64 channels per pixel: one float2 per lane, coalesced 256 B per pixel
    for (int p = 0; p < k2; ++p) {
      const int ky = p / g.k, kx = p - ky * g.k;
      const int iy = iy0 + ky, ix = ix0 + kx;
      float2 v = make_float2(0.f, 0.f);
      if (iy >= 0 && iy < g.Hs && ix >= 0 && ix < g.Ws)
        v = __ldg(reinterpret_cast<const float2*>(src + (((long long)b * g.Hs + iy) * g.Ws + ix) * kE) + lane);
      rb[(2 * lane) * k2 + p] = v.x; rb[(2 * lane + 1) * k2 + p] = v.y;
    }
  }
}

template <bool LN>
__global__ void __launch_bounds__(256) unfold_ln_kernel(const float* __restrict__ src, const Geom g, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, h16* __restrict__ out, int ldo, int split,
                                                        float* __restrict__ mean, float* __restrict__ rstd, int M) {
  extern __shared__ float rowbuf[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* rb = rowbuf + warp * g.Kp;
  int* tab = reinterpret_cast<int*>(rowbuf + 8 * g.Kp);
  build_tab(g, tab);
  __syncthreads();
  const int per_img = g.Ho * g.Wo;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const int b = row / per_img, t = row - b * per_img, oy = t / g.Wo, ox = t - oy * g.Wo;
    gather_patch(src, g, tab, b, oy, ox, lane, rb);
    __syncwarp();
    float mu = 0.f, rs = 1.f;
    if (LN) {
      float s = 0.f;
      for (int e = lane; e < g.dim; e += 32) s += rb[e];
      mu = warp_sum(s) / g.dim;
      float v = 0.f;
      for (int e = lane; e < g.dim; e += 32) { const float d = rb[e] - mu; v += d * d; }
      rs = rsqrtf(warp_sum(v) / g.dim + eps);
      if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
    }
    // split: the row is written as [hi | lo] with lo = fp16(value - hi), i.e. ~21 significant bits for the kqv GEMM (see t2t_forward)
    __half2* o2 = reinterpret_cast<__half2*>(out + (long long)row * ldo);
    for (int e = 2 * lane; e < g.Kp; e += 64) {
      float a = 0.f, c = 0.f;
      if (e < g.dim) a = LN ? (rb[e] - mu) * rs * __ldg(gamma + e) + __ldg(beta + e) : rb[e];
      if (e + 1 < g.dim) c = LN ? (rb[e + 1] - mu) * rs * __ldg(gamma + e + 1) + __ldg(beta + e + 1) : rb[e + 1];
      const __half2 hi = __floats2half2_rn(a, c);
      o2[e >> 1] = hi;
      if (split) o2[(g.Kp + e) >> 1] = __floats2half2_rn(a - __low2float(hi), c - __high2float(hi));
    }
    __syncwarp();
  }
}

// Backward of the soft split (+ LayerNorm): dy16 [M, Kp] is the gradient w.r.t. the (normalised) unfolded rows, times the loss scale S.
// LN: dx = rstd (g - mean(g) - xhat mean(g xhat)), g = dy gamma; dgamma += dy xhat, dbeta += dy  (xhat recomputed from the gathered patch).
// dsrc (token-major [B, Hs*Ws, 64], zeroed by the caller; NULL for the image stage) receives the fold: every patch element is added to the
// pixel it was read from (float2 vector atomics; a pixel is touched by <= 4 patches).
template <bool LN, int kMaxPerLane>       // kMaxPerLane = ceil(dim / 32): 5 for the 147-wide image stage (40 fewer registers, twice the resident warps), 18 for 576
__global__ void __launch_bounds__(256, (kMaxPerLane <= 5 ? 4 : 2)) unfold_ln_bwd_kernel(const h16* __restrict__ dy16, const float* __restrict__ src, const Geom g,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ scales,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dsrc, int M) {
  extern __shared__ float rowbuf[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* rb = rowbuf + warp * g.Kp;                       // patch, then dx
  float* red = rowbuf + 8 * g.Kp;                         // [2 * dim] block-level dgamma / dbeta
  int* tab = reinterpret_cast<int*>(red + 2 * g.dim);
  if (LN) build_tab(g, tab);
  const float invS = scales[1];
  float ag[kMaxPerLane], ab[kMaxPerLane];
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) ag[i] = ab[i] = 0.f;
  if (LN) { for (int i = threadIdx.x; i < 2 * g.dim; i += blockDim.x) red[i] = 0.f; }
  __syncthreads();
  const int per_img = g.Ho * g.Wo, k2 = g.k * g.k;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const int b = row / per_img, t = row - b * per_img, oy = t / g.Wo, ox = t - oy * g.Wo;
    const h16* dyr = dy16 + (long long)row * g.Kp;
    if (LN) {
      gather_patch(src, g, tab, b, oy, ox, lane, rb);
      __syncwarp();
      const float mu = mean[row], rs = rstd[row];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxPerLane; ++i) {
        const int e = lane + 32 * i;
        if (e < g.dim) {
          const float dy = __half2float(dyr[e]) * invS, xh = (rb[e] - mu) * rs, gg = dy * __ldg(gamma + e);
          ag[i] += dy * xh; ab[i] += dy;
          s1 += gg; s2 += gg * xh;
        }
      }
      s1 = warp_sum(s1) / g.dim; s2 = warp_sum(s2) / g.dim;
      if (dsrc) {
        for (int e = lane; e < g.dim; e += 32) {
          const float dy = __half2float(dyr[e]) * invS, xh = (rb[e] - mu) * rs;
          rb[e] = rs * (dy * __ldg(gamma + e) - s1 - xh * s2);
        }
      }
    } else {
      for (int e = lane; e < g.dim; e += 32) rb[e] = __half2float(dyr[e]) * invS;
    }
    __syncwarp();
    if (dsrc) {
      const int iy0 = oy * g.stride - g.pad, ix0 = ox * g.stride - g.pad;
      for (int p = 0; p < k2; ++p) {
        const int ky = p / g.k, kx = p - ky * g.k;
        const int iy = iy0 + ky, ix = ix0 + kx;
        if (iy >= 0 && iy < g.Hs && ix >= 0 && ix < g.Ws)
          atomicAdd(reinterpret_cast<float2*>(dsrc + (((long long)b * g.Hs + iy) * g.Ws + ix) * kE) + lane,
                    make_float2(rb[(2 * lane) * k2 + p], rb[(2 * lane + 1) * k2 + p]));
      }
    }
    __syncwarp();
  }
  if (LN) {
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int e = lane + 32 * i;
      if (e < g.dim) { atomicAdd(red + e, ag[i]); atomicAdd(red + g.dim + e, ab[i]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.dim; i += blockDim.x) { atomicAdd(dgamma + i, red[i]); atomicAdd(dbeta + i, red[g.dim + i]); }
  }
}

// ------------------------------------------------------------------------------------------------ performer attention (token_performer.py:31-52)
// kqv: [B*T, 192] fp32 (k | q | v, nn.Linear(dim, 3*emb) order).  ws: per image [64*32 + 32] = kptv[e][j] then ksum[j].
constexpr int kWsPerImg = kE * kF + kF;
constexpr int kTokPerWarp = 16;

// These kernels are bound by the shared-memory pipe (ncu: l1tex throughput 73-83 %, profiles/r02_ncu_t2t_frontend.txt): every FMA of a
// 64 x 32 contraction reads one DISTINCT word per lane (this lane's element of w / kptv: 128 B per warp instruction) plus one broadcast word.
// So each warp works on TWO tokens at a time -- every distinct read feeds two FMAs -- and the per-token vectors are read as 128-bit
// broadcasts (one wavefront for four values): ~60-140 shared-memory wavefronts per token instead of ~200-330.
struct __align__(16) WarpScratch { float a[2][kE]; float b[2][kE]; float c[2][kE]; };

__device__ __forceinline__ void load_w(const float* __restrict__ w, float (*w_s)[kE + 1]) {
  for (int i = threadIdx.x; i < kF * kE; i += blockDim.x) w_s[i / kE][i % kE] = __ldg(w + i);
}
// random features of two tokens for this lane's feature j = lane: exp(w_j . x - |x|^2 / 2) / sqrt(m); xa / xb hold the two 64-vectors
__device__ __forceinline__ void prm_feature2(const float (*w_s)[kE + 1], const float* xa, const float* xb, int lane, float& fa, float& fb) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int e = 0; e < kE; e += 4) {
    const float4 A = *reinterpret_cast<const float4*>(xa + e), B = *reinterpret_cast<const float4*>(xb + e);
    const float w0 = w_s[lane][e], w1 = w_s[lane][e + 1], w2 = w_s[lane][e + 2], w3 = w_s[lane][e + 3];
    a0 = fmaf(w0, A.x, a0); a1 = fmaf(w1, A.y, a1); a0 = fmaf(w2, A.z, a0); a1 = fmaf(w3, A.w, a1);
    b0 = fmaf(w0, B.x, b0); b1 = fmaf(w1, B.y, b1); b0 = fmaf(w2, B.z, b0); b1 = fmaf(w3, B.w, b1);
  }
  float na = xa[lane] * xa[lane] + xa[lane + 32] * xa[lane + 32], nb = xb[lane] * xb[lane] + xb[lane + 32] * xb[lane + 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { na += __shfl_xor_sync(0xffffffffu, na, o); nb += __shfl_xor_sync(0xffffffffu, nb, o); }
  fa = expf((a0 + a1) - 0.5f * na) * 0.17677669529663687f;     // 1 / sqrt(32)
  fb = expf((b0 + b1) - 0.5f * nb) * 0.17677669529663687f;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
}
// out_e = sum_j u[j] w_s[j][e] for e = lane, lane + 32 and two tokens (u vectors in shared memory)
__device__ __forceinline__ void wT_times2(const float (*w_s)[kE + 1], const float* ua, const float* ub, int lane, float& a0, float& a1, float& b0, float& b1) {
#pragma unroll
  for (int j = 0; j < kF; j += 4) {
    const float4 A = *reinterpret_cast<const float4*>(ua + j), B = *reinterpret_cast<const float4*>(ub + j);
    float w0 = w_s[j][lane], w1 = w_s[j][lane + 32];
    a0 = fmaf(A.x, w0, a0); a1 = fmaf(A.x, w1, a1); b0 = fmaf(B.x, w0, b0); b1 = fmaf(B.x, w1, b1);
    w0 = w_s[j + 1][lane]; w1 = w_s[j + 1][lane + 32];
    a0 = fmaf(A.y, w0, a0); a1 = fmaf(A.y, w1, a1); b0 = fmaf(B.y, w0, b0); b1 = fmaf(B.y, w1, b1);
    w0 = w_s[j + 2][lane]; w1 = w_s[j + 2][lane + 32];
    a0 = fmaf(A.z, w0, a0); a1 = fmaf(A.z, w1, a1); b0 = fmaf(B.z, w0, b0); b1 = fmaf(B.z, w1, b1);
    w0 = w_s[j + 3][lane]; w1 = w_s[j + 3][lane + 32];
    a0 = fmaf(A.w, w0, a0); a1 = fmaf(A.w, w1, a1); b0 = fmaf(B.w, w0, b0); b1 = fmaf(B.w, w1, b1);
  }
}

__global__ void __launch_bounds__(256) performer_kv_kernel(const float* __restrict__ kqv, const float* __restrict__ w, float* __restrict__ ws, int T) {
  __shared__ float w_s[kF][kE + 1];
  __shared__ float red[kWsPerImg];
  __shared__ WarpScratch sc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  load_w(w, w_s);
  for (int i = threadIdx.x; i < kWsPerImg; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[kE];
#pragma unroll
  for (int e = 0; e < kE; ++e) acc[e] = 0.f;
  float ks = 0.f;
  const int t0 = (blockIdx.x * 8 + warp) * kTokPerWarp, t1 = min(T, t0 + kTokPerWarp);
  for (int t = t0; t < t1; t += 2) {
    const bool two = t + 1 < t1;
    const float* ra = kqv + ((long long)b * T + t) * kKqv;
    const float* rb = two ? ra + kKqv : ra;
    const float2 ka = *reinterpret_cast<const float2*>(ra + 2 * lane), va = *reinterpret_cast<const float2*>(ra + 2 * kE + 2 * lane);
    const float2 kb = *reinterpret_cast<const float2*>(rb + 2 * lane), vb = *reinterpret_cast<const float2*>(rb + 2 * kE + 2 * lane);
    __syncwarp();
    *reinterpret_cast<float2*>(sc[warp].a[0] + 2 * lane) = ka; *reinterpret_cast<float2*>(sc[warp].a[1] + 2 * lane) = kb;
    *reinterpret_cast<float2*>(sc[warp].b[0] + 2 * lane) = va; *reinterpret_cast<float2*>(sc[warp].b[1] + 2 * lane) = vb;
    __syncwarp();
    float kpa, kpb;
    prm_feature2(w_s, sc[warp].a[0], sc[warp].a[1], lane, kpa, kpb);
    if (!two) kpb = 0.f;
    ks += kpa + kpb;
#pragma unroll
    for (int e = 0; e < kE; e += 4) {
      const float4 A = *reinterpret_cast<const float4*>(sc[warp].b[0] + e), B = *reinterpret_cast<const float4*>(sc[warp].b[1] + e);
      acc[e] = fmaf(A.x, kpa, fmaf(B.x, kpb, acc[e])); acc[e + 1] = fmaf(A.y, kpa, fmaf(B.y, kpb, acc[e + 1]));
      acc[e + 2] = fmaf(A.z, kpa, fmaf(B.z, kpb, acc[e + 2])); acc[e + 3] = fmaf(A.w, kpa, fmaf(B.w, kpb, acc[e + 3]));
    }
  }
#pragma unroll
  for (int e = 0; e < kE; ++e) atomicAdd(&red[e * kF + lane], acc[e]);
  atomicAdd(&red[kE * kF + lane], ks);
  __syncthreads();
  for (int i = threadIdx.x; i < kWsPerImg; i += blockDim.x) atomicAdd(ws + (long long)b * kWsPerImg + i, red[i]);
}

__global__ void __launch_bounds__(256) performer_out_kernel(const float* __restrict__ kqv, const float* __restrict__ w, const float* __restrict__ ws,
                                                            h16* __restrict__ y16, int T, float eps) {
  __shared__ float w_s[kF][kE + 1];
  __shared__ WarpScratch sc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  load_w(w, w_s);
  const float* wsb = ws + (long long)b * kWsPerImg;
  float kr0[kF], kr1[kF];                               // rows `lane`, `lane + 32` of this image's kptv
#pragma unroll
  for (int j = 0; j < kF; j += 4) {
    const float4 a = *reinterpret_cast<const float4*>(wsb + lane * kF + j), c = *reinterpret_cast<const float4*>(wsb + (lane + 32) * kF + j);
    kr0[j] = a.x; kr0[j + 1] = a.y; kr0[j + 2] = a.z; kr0[j + 3] = a.w; kr1[j] = c.x; kr1[j + 1] = c.y; kr1[j + 2] = c.z; kr1[j + 3] = c.w;
  }
  const float ksum = wsb[kE * kF + lane];
  __syncthreads();
  const int t0 = (blockIdx.x * 8 + warp) * kTokPerWarp, t1 = min(T, t0 + kTokPerWarp);
  for (int t = t0; t < t1; t += 2) {
    const bool two = t + 1 < t1;
    const long long r = (long long)b * T + t;
    const float2 qa = *reinterpret_cast<const float2*>(kqv + r * kKqv + kE + 2 * lane);
    const float2 qb = *reinterpret_cast<const float2*>(kqv + (two ? r + 1 : r) * kKqv + kE + 2 * lane);
    __syncwarp();
    *reinterpret_cast<float2*>(sc[warp].a[0] + 2 * lane) = qa; *reinterpret_cast<float2*>(sc[warp].a[1] + 2 * lane) = qb;
    __syncwarp();
    float qpa, qpb;
    prm_feature2(w_s, sc[warp].a[0], sc[warp].a[1], lane, qpa, qpb);
    float dena = qpa * ksum, denb = qpb * ksum;
    warp_sum2(dena, denb);
    sc[warp].b[0][lane] = qpa; sc[warp].b[1][lane] = qpb;
    __syncwarp();
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
    for (int j = 0; j < kF; j += 4) {
      const float4 A = *reinterpret_cast<const float4*>(sc[warp].b[0] + j), B = *reinterpret_cast<const float4*>(sc[warp].b[1] + j);
      a0 = fmaf(kr0[j], A.x, a0); a1 = fmaf(kr1[j], A.x, a1); b0 = fmaf(kr0[j], B.x, b0); b1 = fmaf(kr1[j], B.x, b1);
      a0 = fmaf(kr0[j + 1], A.y, a0); a1 = fmaf(kr1[j + 1], A.y, a1); b0 = fmaf(kr0[j + 1], B.y, b0); b1 = fmaf(kr1[j + 1], B.y, b1);
      a0 = fmaf(kr0[j + 2], A.z, a0); a1 = fmaf(kr1[j + 2], A.z, a1); b0 = fmaf(kr0[j + 2], B.z, b0); b1 = fmaf(kr1[j + 2], B.z, b1);
      a0 = fmaf(kr0[j + 3], A.w, a0); a1 = fmaf(kr1[j + 3], A.w, a1); b0 = fmaf(kr0[j + 3], B.w, b0); b1 = fmaf(kr1[j + 3], B.w, b1);
    }
    dena += eps; denb += eps;
    y16[r * kE + lane] = __float2half_rn(a0 / dena); y16[r * kE + lane + 32] = __float2half_rn(a1 / dena);
    if (two) { y16[(r + 1) * kE + lane] = __float2half_rn(b0 / denb); y16[(r + 1) * kE + lane + 32] = __float2half_rn(b1 / denb); }
  }
}

// Backward, pass A (per token, needs the forward's per-image kptv / ksum): with den = qp . ksum + eps, y = num / den,
//   dnum = dy / den ; dden = -(dy . y) / den ; dqp = kptv^T dnum + dden ksum ; d kptv += dnum (x) qp ; d ksum += dden qp ;
//   du = dqp qp ; dq = w^T du - (sum du) q      (u_j = w_j . q - |q|^2 / 2).
// dy16 carries the loss scale S; everything written here (dq16, d kptv, d ksum) carries it too, the bias-gradient sums take it back out.
__global__ void __launch_bounds__(256) performer_bwd_a_kernel(const float* __restrict__ kqv, const float* __restrict__ w, const float* __restrict__ ws,
                                                              const h16* __restrict__ y16, const h16* __restrict__ dy16, float* __restrict__ dws,
                                                              h16* __restrict__ dkqv16, float* __restrict__ dbias, const float* __restrict__ scales,
                                                              int T, float eps) {
  __shared__ float w_s[kF][kE + 1];
  __shared__ float kptv_s[kE][kF + 1];
  __shared__ float red[kWsPerImg];
  __shared__ float cs_s[kE];
  __shared__ WarpScratch sc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  load_w(w, w_s);
  const float* wsb = ws + (long long)b * kWsPerImg;
  for (int i = threadIdx.x; i < kE * kF; i += blockDim.x) kptv_s[i / kF][i % kF] = wsb[i];
  for (int i = threadIdx.x; i < kWsPerImg; i += blockDim.x) red[i] = 0.f;
  if (threadIdx.x < kE) cs_s[threadIdx.x] = 0.f;
  __syncthreads();
  const float ksum = wsb[kE * kF + lane];
  float acc[kE];
#pragma unroll
  for (int e = 0; e < kE; ++e) acc[e] = 0.f;
  float dks = 0.f, cs0 = 0.f, cs1 = 0.f;
  const int t0 = (blockIdx.x * 8 + warp) * kTokPerWarp, t1 = min(T, t0 + kTokPerWarp);
  for (int t = t0; t < t1; t += 2) {
    const bool two = t + 1 < t1;
    const long long ra = (long long)b * T + t, rb = two ? ra + 1 : ra;
    const float2 qa = *reinterpret_cast<const float2*>(kqv + ra * kKqv + kE + 2 * lane), qb = *reinterpret_cast<const float2*>(kqv + rb * kKqv + kE + 2 * lane);
    const float ya0 = __half2float(y16[ra * kE + lane]), ya1 = __half2float(y16[ra * kE + lane + 32]);
    const float yb0 = __half2float(y16[rb * kE + lane]), yb1 = __half2float(y16[rb * kE + lane + 32]);
    const float da0 = __half2float(dy16[ra * kE + lane]), da1 = __half2float(dy16[ra * kE + lane + 32]);
    const float db0 = two ? __half2float(dy16[rb * kE + lane]) : 0.f, db1 = two ? __half2float(dy16[rb * kE + lane + 32]) : 0.f;
    __syncwarp();
    *reinterpret_cast<float2*>(sc[warp].a[0] + 2 * lane) = qa; *reinterpret_cast<float2*>(sc[warp].a[1] + 2 * lane) = qb;
    __syncwarp();
    float qpa, qpb;
    prm_feature2(w_s, sc[warp].a[0], sc[warp].a[1], lane, qpa, qpb);
    float dena = qpa * ksum, denb = qpb * ksum;
    warp_sum2(dena, denb);
    dena += eps; denb += eps;
    float dda = da0 * ya0 + da1 * ya1, ddb = db0 * yb0 + db1 * yb1;
    warp_sum2(dda, ddb);
    dda = -dda / dena; ddb = -ddb / denb;                        // dden of the two tokens (zero for a missing second token: its dy is zero)
    sc[warp].b[0][lane] = da0 / dena; sc[warp].b[0][lane + 32] = da1 / dena; sc[warp].b[1][lane] = db0 / denb; sc[warp].b[1][lane + 32] = db1 / denb;
    __syncwarp();
    float dqa = dda * ksum, dqb = ddb * ksum, dqa1 = 0.f, dqb1 = 0.f;
#pragma unroll
    for (int e = 0; e < kE; e += 4) {
      const float4 A = *reinterpret_cast<const float4*>(sc[warp].b[0] + e), B = *reinterpret_cast<const float4*>(sc[warp].b[1] + e);
      const float k0 = kptv_s[e][lane], k1 = kptv_s[e + 1][lane], k2 = kptv_s[e + 2][lane], k3 = kptv_s[e + 3][lane];
      dqa = fmaf(A.x, k0, dqa); dqa1 = fmaf(A.y, k1, dqa1); dqa = fmaf(A.z, k2, dqa); dqa1 = fmaf(A.w, k3, dqa1);
      dqb = fmaf(B.x, k0, dqb); dqb1 = fmaf(B.y, k1, dqb1); dqb = fmaf(B.z, k2, dqb); dqb1 = fmaf(B.w, k3, dqb1);
      acc[e] = fmaf(A.x, qpa, fmaf(B.x, qpb, acc[e])); acc[e + 1] = fmaf(A.y, qpa, fmaf(B.y, qpb, acc[e + 1]));
      acc[e + 2] = fmaf(A.z, qpa, fmaf(B.z, qpb, acc[e + 2])); acc[e + 3] = fmaf(A.w, qpa, fmaf(B.w, qpb, acc[e + 3]));
    }
    dqa += dqa1; dqb += dqb1;
    dks = fmaf(dda, qpa, fmaf(ddb, qpb, dks));
    float dua = dqa * qpa, dub = dqb * qpb;
    sc[warp].c[0][lane] = dua; sc[warp].c[1][lane] = dub;
    warp_sum2(dua, dub);                                         // sum_j du
    __syncwarp();
    float a0 = -dua * sc[warp].a[0][lane], a1 = -dua * sc[warp].a[0][lane + 32], b0 = -dub * sc[warp].a[1][lane], b1 = -dub * sc[warp].a[1][lane + 32];
    wT_times2(w_s, sc[warp].c[0], sc[warp].c[1], lane, a0, a1, b0, b1);
    dkqv16[ra * kKqv + kE + lane] = __float2half_rn(a0); dkqv16[ra * kKqv + kE + lane + 32] = __float2half_rn(a1);
    cs0 += a0; cs1 += a1;
    if (two) { dkqv16[rb * kKqv + kE + lane] = __float2half_rn(b0); dkqv16[rb * kKqv + kE + lane + 32] = __float2half_rn(b1); cs0 += b0; cs1 += b1; }
  }
#pragma unroll
  for (int e = 0; e < kE; ++e) atomicAdd(&red[e * kF + lane], acc[e]);
  atomicAdd(&red[kE * kF + lane], dks);
  atomicAdd(&cs_s[lane], cs0); atomicAdd(&cs_s[lane + 32], cs1);
  __syncthreads();
  for (int i = threadIdx.x; i < kWsPerImg; i += blockDim.x) atomicAdd(dws + (long long)b * kWsPerImg + i, red[i]);
  if (dbias && threadIdx.x < kE) atomicAdd(dbias + kE + threadIdx.x, cs_s[threadIdx.x] * scales[1]);
}

// Backward, pass B (after pass A has reduced d kptv / d ksum over the whole image):
//   dv = d kptv kp + S dres   (dres: the fp32 gradient of the skip connection y = v + proj(.), unscaled) ;
//   dkp = d kptv^T v + d ksum ; du = dkp kp ; dk = w^T du - (sum du) k.
__global__ void __launch_bounds__(256) performer_bwd_b_kernel(const float* __restrict__ kqv, const float* __restrict__ w, const float* __restrict__ dws,
                                                              const float* __restrict__ dres, h16* __restrict__ dkqv16, float* __restrict__ dbias,
                                                              const float* __restrict__ scales, int T) {
  __shared__ float w_s[kF][kE + 1];
  __shared__ float dkptv_s[kE][kF + 1];
  __shared__ float cs_s[2 * kE];
  __shared__ WarpScratch sc[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
  load_w(w, w_s);
  const float* dwsb = dws + (long long)b * kWsPerImg;
  for (int i = threadIdx.x; i < kE * kF; i += blockDim.x) dkptv_s[i / kF][i % kF] = dwsb[i];
  if (threadIdx.x < 2 * kE) cs_s[threadIdx.x] = 0.f;
  float kr0[kF], kr1[kF];                               // rows `lane`, `lane + 32` of this image's d kptv
#pragma unroll
  for (int j = 0; j < kF; j += 4) {
    const float4 a = *reinterpret_cast<const float4*>(dwsb + lane * kF + j), c = *reinterpret_cast<const float4*>(dwsb + (lane + 32) * kF + j);
    kr0[j] = a.x; kr0[j + 1] = a.y; kr0[j + 2] = a.z; kr0[j + 3] = a.w; kr1[j] = c.x; kr1[j + 1] = c.y; kr1[j + 2] = c.z; kr1[j + 3] = c.w;
  }
  const float dksum = dwsb[kE * kF + lane];
  __syncthreads();
  const float S = scales[0];
  float ck0 = 0.f, ck1 = 0.f, cv0 = 0.f, cv1 = 0.f;
  const int t0 = (blockIdx.x * 8 + warp) * kTokPerWarp, t1 = min(T, t0 + kTokPerWarp);
  for (int t = t0; t < t1; t += 2) {
    const bool two = t + 1 < t1;
    const long long ra = (long long)b * T + t, rb = two ? ra + 1 : ra;
    const float2 ka = *reinterpret_cast<const float2*>(kqv + ra * kKqv + 2 * lane), va = *reinterpret_cast<const float2*>(kqv + ra * kKqv + 2 * kE + 2 * lane);
    const float2 kb = *reinterpret_cast<const float2*>(kqv + rb * kKqv + 2 * lane), vb = *reinterpret_cast<const float2*>(kqv + rb * kKqv + 2 * kE + 2 * lane);
    float dva0 = S * dres[ra * kE + lane], dva1 = S * dres[ra * kE + lane + 32], dvb0 = S * dres[rb * kE + lane], dvb1 = S * dres[rb * kE + lane + 32];
    __syncwarp();
    *reinterpret_cast<float2*>(sc[warp].a[0] + 2 * lane) = ka; *reinterpret_cast<float2*>(sc[warp].a[1] + 2 * lane) = kb;
    *reinterpret_cast<float2*>(sc[warp].b[0] + 2 * lane) = va; *reinterpret_cast<float2*>(sc[warp].b[1] + 2 * lane) = vb;
    __syncwarp();
    float kpa, kpb;
    prm_feature2(w_s, sc[warp].a[0], sc[warp].a[1], lane, kpa, kpb);
    sc[warp].c[0][lane] = kpa; sc[warp].c[1][lane] = kpb;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kF; j += 4) {
      const float4 A = *reinterpret_cast<const float4*>(sc[warp].c[0] + j), B = *reinterpret_cast<const float4*>(sc[warp].c[1] + j);
      dva0 = fmaf(kr0[j], A.x, dva0); dva1 = fmaf(kr1[j], A.x, dva1); dvb0 = fmaf(kr0[j], B.x, dvb0); dvb1 = fmaf(kr1[j], B.x, dvb1);
      dva0 = fmaf(kr0[j + 1], A.y, dva0); dva1 = fmaf(kr1[j + 1], A.y, dva1); dvb0 = fmaf(kr0[j + 1], B.y, dvb0); dvb1 = fmaf(kr1[j + 1], B.y, dvb1);
      dva0 = fmaf(kr0[j + 2], A.z, dva0); dva1 = fmaf(kr1[j + 2], A.z, dva1); dvb0 = fmaf(kr0[j + 2], B.z, dvb0); dvb1 = fmaf(kr1[j + 2], B.z, dvb1);
      dva0 = fmaf(kr0[j + 3], A.w, dva0); dva1 = fmaf(kr1[j + 3], A.w, dva1); dvb0 = fmaf(kr0[j + 3], B.w, dvb0); dvb1 = fmaf(kr1[j + 3], B.w, dvb1);
    }
    float dka = dksum, dkb = dksum, dka1 = 0.f, dkb1 = 0.f;
#pragma unroll
    for (int e = 0; e < kE; e += 4) {
      const float4 A = *reinterpret_cast<const float4*>(sc[warp].b[0] + e), B = *reinterpret_cast<const float4*>(sc[warp].b[1] + e);
      const float k0 = dkptv_s[e][lane], k1 = dkptv_s[e + 1][lane], k2 = dkptv_s[e + 2][lane], k3 = dkptv_s[e + 3][lane];
      dka = fmaf(k0, A.x, dka); dka1 = fmaf(k1, A.y, dka1); dka = fmaf(k2, A.z, dka); dka1 = fmaf(k3, A.w, dka1);
      dkb = fmaf(k0, B.x, dkb); dkb1 = fmaf(k1, B.y, dkb1); dkb = fmaf(k2, B.z, dkb); dkb1 = fmaf(k3, B.w, dkb1);
    }
    float dua = (dka + dka1) * kpa, dub = (dkb + dkb1) * kpb;
    __syncwarp();
    sc[warp].c[0][lane] = dua; sc[warp].c[1][lane] = dub;
    warp_sum2(dua, dub);
    __syncwarp();
    float a0 = -dua * sc[warp].a[0][lane], a1 = -dua * sc[warp].a[0][lane + 32], b0 = -dub * sc[warp].a[1][lane], b1 = -dub * sc[warp].a[1][lane + 32];
    wT_times2(w_s, sc[warp].c[0], sc[warp].c[1], lane, a0, a1, b0, b1);
    h16* oa = dkqv16 + ra * kKqv;
    oa[lane] = __float2half_rn(a0); oa[lane + 32] = __float2half_rn(a1); oa[2 * kE + lane] = __float2half_rn(dva0); oa[2 * kE + lane + 32] = __float2half_rn(dva1);
    ck0 += a0; ck1 += a1; cv0 += dva0; cv1 += dva1;
    if (two) {
      h16* ob = dkqv16 + rb * kKqv;
      ob[lane] = __float2half_rn(b0); ob[lane + 32] = __float2half_rn(b1); ob[2 * kE + lane] = __float2half_rn(dvb0); ob[2 * kE + lane + 32] = __float2half_rn(dvb1);
      ck0 += b0; ck1 += b1; cv0 += dvb0; cv1 += dvb1;
    }
  }
  atomicAdd(&cs_s[lane], ck0); atomicAdd(&cs_s[lane + 32], ck1); atomicAdd(&cs_s[kE + lane], cv0); atomicAdd(&cs_s[kE + lane + 32], cv1);
  __syncthreads();
  if (dbias && threadIdx.x < 2 * kE) {
    const int col = threadIdx.x < kE ? threadIdx.x : kE + threadIdx.x;        // k -> [0, 64), v -> [128, 192)
    atomicAdd(dbias + col, cs_s[threadIdx.x] * scales[1]);
  }
}

// ------------------------------------------------------------------------------------------------ dropout (training mode only)
// Token_performer applies nn.Dropout(0.1) to proj(y) and to the MLP output (token_performer.py:20,28,51,66).  The keep decision of element i
// is a stateless hash of (seed, i) -- the same in forward and backward, nothing stored.  The reference's CUDA Philox stream cannot be reproduced
// bit for bit on another device anyway; the distribution (Bernoulli(1 - p), scaled by 1 / (1 - p)) is the same.
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned long long i, float p) {
  unsigned long long z = seed + i * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;      // splitmix64 finaliser
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}
// out = res + keep * z / (1 - p)         (z, out: [M, 64] contiguous; res row stride ldr)
__global__ void __launch_bounds__(256) dropout_add_kernel(const float* __restrict__ z, const float* __restrict__ res, long long ldr, float* __restrict__ out,
                                                          long long n, unsigned long long seed, float p) {
  const float ik = 1.0f / (1.0f - p);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / kE; const int c = (int)(i - r * kE);
    out[i] = res[r * ldr + c] + (keep_elem(seed, (unsigned long long)i, p) ? z[i] * ik : 0.f);
  }
}
// g16 = fp16(S * keep * d / (1 - p)) and db[c] += sum_rows keep * d / (1 - p)        (d: [M, 64])
__global__ void __launch_bounds__(256) dropout_grad_kernel(const float* __restrict__ d, h16* __restrict__ g16, float* __restrict__ db, long long M,
                                                           unsigned long long seed, float p, const float* __restrict__ scales) {
  __shared__ float cs_s[kE];
  if (threadIdx.x < kE) cs_s[threadIdx.x] = 0.f;
  __syncthreads();
  const float ik = p > 0.f ? 1.0f / (1.0f - p) : 1.0f, S = scales[0];
  const int c = threadIdx.x & 63, sub = threadIdx.x >> 6;      // 4 rows per block iteration
  float cs = 0.f;
  for (long long r = (long long)blockIdx.x * 4 + sub; r < M; r += (long long)gridDim.x * 4) {
    const long long i = r * kE + c;
    const float v = (p > 0.f && !keep_elem(seed, (unsigned long long)i, p)) ? 0.f : d[i] * ik;
    g16[i] = __float2half_rn(S * v);
    cs += v;
  }
  atomicAdd(&cs_s[c], cs);
  __syncthreads();
  if (db && threadIdx.x < kE) atomicAdd(db + threadIdx.x, cs_s[threadIdx.x]);
}

// loss scale of the fp16 gradient operands from max|d tokens| (the engine's grad_scale_kernel is one block: fine for [B, 1000] logits, not for
// the 9.6 M token gradients here): grid-wide atomic max on the bit pattern (non-negative floats order like unsigned integers), then one thread
// picks the largest power of two S with S * max <= target.  scales = {S, 1/S, max bits}.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ out) {
  float mx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) mx = fmaxf(mx, fabsf(x[i]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(mx));
}
__global__ void pick_scale_kernel(float* scales, float target, float fixed) {
  float S = fixed;
  if (!(fixed > 0.f)) {
    const float mx = __uint_as_float(reinterpret_cast<const unsigned*>(scales)[2]);
    int e = 0;
    if (mx > 0.f && mx < INFINITY) { frexpf(target / mx, &e); e -= 1; }
    e = max(-24, min(24, e));
    S = ldexpf(1.0f, e);
  }
  scales[0] = S; scales[1] = 1.0f / S;
}

// fp32 [rows, cols] -> fp16 [rows, ld] (zero padded) and optionally the transpose fp16 [cols_pad, rows] (rows cols..cols_pad zero)
__global__ void __launch_bounds__(256) cvt_pad_kernel(const float* __restrict__ src, int rows, int cols, h16* __restrict__ dst, int ld, h16* __restrict__ dstT,
                                                      int cols_pad) {
  const long long n = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i - (long long)r * ld);
    const float v = c < cols ? src[(long long)r * cols + c] : 0.f;
    dst[i] = __float2half_rn(v);
    if (dstT && c < cols_pad) dstT[(long long)c * rows + r] = __float2half_rn(v);
  }
}
// kqv weight [rows, cols] -> hi2 [rows, 2 ld] = [hi | hi], lo [rows, ld] = fp16(w - hi), hiT [ld, rows] (zero padded to ld columns)
__global__ void __launch_bounds__(256) cvt_split_kernel(const float* __restrict__ src, int rows, int cols, int ld, h16* __restrict__ hi2, h16* __restrict__ lo,
                                                        h16* __restrict__ hiT) {
  const long long n = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i - (long long)r * ld);
    const float v = c < cols ? src[(long long)r * cols + c] : 0.f;
    const h16 h = __float2half_rn(v);
    hi2[(long long)r * 2 * ld + c] = h; hi2[(long long)r * 2 * ld + ld + c] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
    if (hiT) hiT[(long long)c * rows + r] = h;
  }
}
// dst[r * cols + c] += scale * src[r * ld + c]    (strips the padding of a weight gradient formed at the padded width)
__global__ void __launch_bounds__(256) add_strip_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int rows, int cols) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    dst[i] += src[(long long)r * ld + c];
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct Bump {
  char* base; size_t off;
  explicit Bump(void* p) : base(static_cast<char*>(p)), off(0) {}
  void* raw(size_t bytes) { bytes = (bytes + 255) & ~size_t(255); char* r = base ? base + off : nullptr; off += bytes; return r; }
  float* f(size_t n) { return static_cast<float*>(raw(n * 4)); }
  h16* h(size_t n) { return static_cast<h16*>(raw(n * 2)); }
};

struct StageWs {
  Geom g; int M, T;
  h16 *kqv_w16, *kqv_wlo16, *kqv_wT16, *proj_w16, *proj_wT16, *m0_w16, *m0_wT16, *m2_w16, *m2_wT16;
  h16 *A16, *y16, *ln2_16, *h16_, *hpre16;
  float *mean1, *rstd1, *kqv, *ws, *z, *y2, *mean2, *rstd2, *out;
  // backward
  float *dout, *dy2, *dws, *dkqv_w;
  h16 *g16, *dh16, *dln2_16, *dy2_16, *dy16, *dkqv16, *dxn16;
};
struct T2TWs {
  StageWs st[2];
  Geom g2; int M2;
  h16 *pr_w16, *pr_wT16, *A2_16, *dtok16, *dA2_16;
  float* scales;
  size_t bytes;
};

Geom make_geom(int nchw, int Cin, int Hs, int k, int stride, int pad) {
  Geom g; g.nchw = nchw; g.Cin = Cin; g.Hs = g.Ws = Hs; g.k = k; g.stride = stride; g.pad = pad;
  g.Ho = g.Wo = (Hs + 2 * pad - k) / stride + 1;
  g.dim = Cin * k * k; g.Kp = (g.dim + 7) / 8 * 8;
  return g;
}

void carve_t2t(int B, int img, int cin, int C, bool save, void* base, T2TWs* w) {
  Bump b(base);
  Geom g0 = make_geom(1, cin, img, 7, 4, 2);
  Geom g1 = make_geom(0, kE, g0.Ho, 3, 2, 1);
  Geom g2 = make_geom(0, kE, g1.Ho, 3, 2, 1);
  w->g2 = g2; w->M2 = B * g2.Ho * g2.Wo;
  const Geom gs[2] = {g0, g1};
  for (int s = 0; s < 2; ++s) {
    StageWs& S = w->st[s];
    S.g = gs[s]; S.T = S.g.Ho * S.g.Wo; S.M = B * S.T;
    const size_t M = S.M, Kp = S.g.Kp;
    S.kqv_w16 = b.h(kKqv * 2 * Kp); S.kqv_wlo16 = b.h(kKqv * Kp); S.kqv_wT16 = b.h(Kp * kKqv);
    S.proj_w16 = b.h(kE * kE); S.proj_wT16 = b.h(kE * kE); S.m0_w16 = b.h(kE * kE); S.m0_wT16 = b.h(kE * kE); S.m2_w16 = b.h(kE * kE); S.m2_wT16 = b.h(kE * kE);
    S.A16 = b.h(M * 2 * Kp); S.mean1 = b.f(M); S.rstd1 = b.f(M);      // [hi | lo] halves, row stride 2 Kp
    S.kqv = b.f(M * kKqv); S.ws = b.f((size_t)B * kWsPerImg);
    S.y16 = b.h(M * kE); S.z = b.f(M * kE); S.y2 = b.f(M * kE); S.mean2 = b.f(M); S.rstd2 = b.f(M);
    S.ln2_16 = b.h(M * kE); S.h16_ = b.h(M * kE); S.hpre16 = b.h(M * kE); S.out = b.f(M * kE);
    if (save) {
      S.dout = b.f(M * kE); S.dy2 = b.f(M * kE); S.dws = b.f((size_t)B * kWsPerImg); S.dkqv_w = b.f(kKqv * Kp);
      S.g16 = b.h(M * kE); S.dh16 = b.h(M * kE); S.dln2_16 = b.h(M * kE); S.dy2_16 = b.h(M * kE); S.dy16 = b.h(M * kE);
      S.dkqv16 = b.h(M * kKqv); S.dxn16 = b.h(M * Kp);
    } else {
      S.dout = S.dy2 = S.dws = S.dkqv_w = nullptr; S.g16 = S.dh16 = S.dln2_16 = S.dy2_16 = S.dy16 = S.dkqv16 = S.dxn16 = nullptr;
    }
  }
  const size_t M2 = w->M2, K2 = g2.Kp;
  w->pr_w16 = b.h((size_t)C * K2); w->pr_wT16 = b.h(K2 * (size_t)C);
  w->A2_16 = b.h(M2 * K2);
  if (save) { w->dtok16 = b.h(M2 * (size_t)C); w->dA2_16 = b.h(M2 * K2); w->scales = b.f(4); }
  else { w->dtok16 = w->dA2_16 = nullptr; w->scales = nullptr; }
  w->bytes = b.off;
}

inline uvc_operand o16k(const h16* p, long long ld) { return uvc_operand{reinterpret_cast<const float*>(p), ld, 0, 0, 0, 0}; }
inline uvc_operand o16mn(const h16* p, long long ld) { return uvc_operand{reinterpret_cast<const float*>(p), ld, 0, 0, 1, 0}; }

// Y = epi(X16 [M,K] W16[N,K]^T)
int lin16(const h16* X, long long ldx, const h16* W, long long ldw, const float* bias, float* D, void* D16, long long ldd, int M, int N, int K, cudaStream_t st,
          int flags = 0, void* aux16 = nullptr, const float* R = nullptr, long long ldr = 0, float* colsum = nullptr, const float* colsum_scale_dev = nullptr) {
  uvc_gemm_args a = gemm_args(M, N, K, o16k(X, ldx), o16k(W, ldw), D, ldd);
  a.flags = flags | UVC_GEMM_F16;
  a.D16 = D16; a.ldd16 = ldd;
  if (bias) { a.bias = bias; a.flags |= UVC_EPI_BIAS; }
  if (aux16) { a.aux = static_cast<float*>(aux16); a.ldaux = ldd; a.flags |= UVC_EPI_AUX_F16; }
  if (R) { a.R = R; a.ldr = ldr; a.flags |= UVC_EPI_RESIDUAL; }
  if (colsum) { a.colsum = colsum; a.colsum_scale_dev = colsum_scale_dev; a.flags |= UVC_EPI_COLSUM; }
  return gemm_tf32(a, st);
}
// dW[N, K] (row stride ldw) += (1/S) dY16[M, N]^T X16[M, K]
int wgrad16(const h16* dY, long long lddy, const h16* X, long long ldx, float* dW, long long ldw, int M, int N, int K, const float* invS, cudaStream_t st) {
  uvc_gemm_args a = gemm_args(N, K, M, o16mn(dY, lddy), o16mn(X, ldx), dW, ldw);
  a.flags = UVC_EPI_ATOMIC | UVC_GEMM_F16;
  a.alpha_dev = invS;
  a.splits = wgrad_splits(N, K, M, 64);
  return gemm_tf32(a, st);
}
inline int blocks_for(long long n, int per_block, int cap = 148 * 8) { long long b = (n + per_block - 1) / per_block; return (int)(b < 1 ? 1 : (b > cap ? cap : b)); }

int cvt_pad(const float* src, int rows, int cols, h16* dst, int ld, h16* dstT, cudaStream_t st) {
  cvt_pad_kernel<<<blocks_for((long long)rows * ld, 256), 256, 0, st>>>(src, rows, cols, dst, ld, dstT, ld);
  return check_launch("t2t cvt_pad");
}
int unfold_ln(const float* src, const Geom& g, const float* gamma, const float* beta, float eps, h16* out, int ldo, int split, float* mean, float* rstd, int M,
              cudaStream_t st) {
  const size_t smem = (8 * (size_t)g.Kp + (g.nchw ? g.dim : 0)) * sizeof(float);
  const int grid = blocks_for(M, 8 * 4, 148 * 6);
  if (gamma) unfold_ln_kernel<true><<<grid, 256, smem, st>>>(src, g, gamma, beta, eps, out, ldo, split, mean, rstd, M);
  else unfold_ln_kernel<false><<<grid, 256, smem, st>>>(src, g, nullptr, nullptr, eps, out, ldo, split, nullptr, nullptr, M);
  return check_launch("t2t unfold_ln");
}
int unfold_ln_bwd(const h16* dy16, const float* src, const Geom& g, const float* gamma, const float* mean, const float* rstd, const float* scales,
                  float* dgamma, float* dbeta, float* dsrc, int M, cudaStream_t st) {
  const size_t smem = (8 * (size_t)g.Kp + 3 * (size_t)g.dim) * sizeof(float);
  const int grid = blocks_for(M, 8 * 16, 148 * (g.dim <= 160 ? 8 : 4));
  if (gamma && g.dim <= 160) unfold_ln_bwd_kernel<true, 5><<<grid, 256, smem, st>>>(dy16, src, g, gamma, mean, rstd, scales, dgamma, dbeta, dsrc, M);
  else if (gamma) unfold_ln_bwd_kernel<true, 18><<<grid, 256, smem, st>>>(dy16, src, g, gamma, mean, rstd, scales, dgamma, dbeta, dsrc, M);
  else unfold_ln_bwd_kernel<false, 1><<<grid, 256, smem, st>>>(dy16, src, g, nullptr, nullptr, nullptr, scales, nullptr, nullptr, dsrc, M);
  return check_launch("t2t unfold_ln_bwd");
}

int check_perf(const uvc_performer_tensors& p, const char* what) {
  UVC_REQUIRE(p.norm1_w && p.norm1_b && p.kqv_w && p.kqv_b && p.w && p.proj_w && p.proj_b && p.norm2_w && p.norm2_b && p.mlp0_w && p.mlp0_b && p.mlp2_w &&
              p.mlp2_b, UVC_ERR_BAD_ARG, "t2t: NULL tensor in %s", what);
  return UVC_OK;
}
int check_t2t(const uvc_t2t_dims& d) {
  UVC_REQUIRE(d.B > 0 && d.img >= 32 && d.img % 16 == 0 && d.in_chans > 0 && d.in_chans * 49 <= 576 && d.C > 0 && d.C % 8 == 0, UVC_ERR_BAD_SHAPE,
              "t2t: unsupported dims B=%d img=%d in_chans=%d C=%d", d.B, d.img, d.in_chans, d.C);
  UVC_REQUIRE((long long)d.B * (d.img / 4) * (d.img / 4) < (1ll << 30), UVC_ERR_BAD_SHAPE, "t2t: too many tokens");
  return UVC_OK;
}

}  // namespace

unsigned long long t2t_workspace_bytes(const uvc_t2t_dims& d, bool save) {
  T2TWs w;
  carve_t2t(d.B, d.img, d.in_chans, d.C, save, nullptr, &w);
  return w.bytes;
}

int t2t_forward(const uvc_t2t_forward_args& a, cudaStream_t st) {
  UVC_TRY(check_t2t(a.dims));
  UVC_TRY(check_perf(a.w.attn1, "attention1")); UVC_TRY(check_perf(a.w.attn2, "attention2"));
  UVC_REQUIRE(a.w.project_w && a.w.project_b && a.x && a.tokens && a.workspace, UVC_ERR_BAD_ARG, "t2t_forward: NULL pointer");
  const bool save = a.save_for_backward != 0;
  T2TWs w;
  carve_t2t(a.dims.B, a.dims.img, a.dims.in_chans, a.dims.C, save, a.workspace, &w);
  UVC_REQUIRE(w.bytes <= a.workspace_bytes, UVC_ERR_WORKSPACE, "t2t_forward: workspace %llu bytes < required %llu", (unsigned long long)a.workspace_bytes,
              (unsigned long long)w.bytes);
  const float eps = a.dims.ln_eps, p = a.dropout_p;
  UVC_REQUIRE(p >= 0.f && p < 1.f, UVC_ERR_BAD_ARG, "t2t_forward: dropout_p must be in [0, 1)");
  const int B = a.dims.B, C = a.dims.C;
  const float* src = a.x;
  for (int s = 0; s < 2; ++s) {
    const uvc_performer_tensors& P = s ? a.w.attn2 : a.w.attn1;
    StageWs& S = w.st[s];
    const int M = S.M, Kp = S.g.Kp, dim = S.g.dim;
    cvt_split_kernel<<<blocks_for((long long)kKqv * Kp, 256), 256, 0, st>>>(P.kqv_w, kKqv, dim, Kp, S.kqv_w16, S.kqv_wlo16, save ? S.kqv_wT16 : nullptr);
    UVC_TRY(check_launch("t2t cvt_split"));
    UVC_TRY(cvt_pad(P.proj_w, kE, kE, S.proj_w16, kE, save ? S.proj_wT16 : nullptr, st));
    UVC_TRY(cvt_pad(P.mlp0_w, kE, kE, S.m0_w16, kE, save ? S.m0_wT16 : nullptr, st));
    UVC_TRY(cvt_pad(P.mlp2_w, kE, kE, S.m2_w16, kE, save ? S.m2_wT16 : nullptr, st));
    UVC_TRY(unfold_ln(src, S.g, P.norm1_w, P.norm1_b, eps, S.A16, 2 * Kp, 1, S.mean1, S.rstd1, M, st));
    // k | q | v in fp32 (v is the skip connection).  k and q enter exp(w . x - |x|^2 / 2): an operand error of 2^-11 there becomes ~0.5 % in the
    // random features, so this one GEMM runs at split precision, x W^T ~ (x_hi + x_lo) W_hi^T + x_hi W_lo^T (two launches, fp32 accumulation):
    // [x_hi | x_lo] against [W_hi | W_hi] over K = 2 Kp, then x_hi against W_lo accumulated in place.  3x a 0.1 ms GEMM.
    UVC_TRY(lin16(S.A16, 2 * Kp, S.kqv_w16, 2 * Kp, P.kqv_b, S.kqv, nullptr, kKqv, M, kKqv, 2 * Kp, st));
    UVC_TRY(lin16(S.A16, 2 * Kp, S.kqv_wlo16, Kp, nullptr, S.kqv, nullptr, kKqv, M, kKqv, Kp, st, 0, nullptr, S.kqv, kKqv));
    cudaError_t e = cudaMemsetAsync(S.ws, 0, (size_t)B * kWsPerImg * sizeof(float), st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "t2t_forward: memset: %s", cudaGetErrorString(e));
    const dim3 pg((S.T + 8 * kTokPerWarp - 1) / (8 * kTokPerWarp), B);
    performer_kv_kernel<<<pg, 256, 0, st>>>(S.kqv, P.w, S.ws, S.T);
    UVC_TRY(check_launch("t2t performer_kv"));
    performer_out_kernel<<<pg, 256, 0, st>>>(S.kqv, P.w, S.ws, S.y16, S.T, 1e-8f);
    UVC_TRY(check_launch("t2t performer_out"));
    const float* v = S.kqv + 2 * kE;
    if (p > 0.f) {      // y2 = v + drop(proj(y));  out = y2 + drop(mlp(LN2(y2)))
      UVC_TRY(lin16(S.y16, kE, S.proj_w16, kE, P.proj_b, S.z, nullptr, kE, M, kE, kE, st));
      dropout_add_kernel<<<blocks_for((long long)M * kE, 256), 256, 0, st>>>(S.z, v, kKqv, S.y2, (long long)M * kE, a.seed + 2 * s, p);
      UVC_TRY(check_launch("t2t dropout_add"));
    } else {
      UVC_TRY(lin16(S.y16, kE, S.proj_w16, kE, P.proj_b, S.y2, nullptr, kE, M, kE, kE, st, 0, nullptr, v, kKqv));
    }
    UVC_TRY(layernorm_fwd(S.y2, kE, P.norm2_w, P.norm2_b, eps, nullptr, kE, S.mean2, S.rstd2, M, kE, st, 0, S.ln2_16));
    UVC_TRY(lin16(S.ln2_16, kE, S.m0_w16, kE, P.mlp0_b, nullptr, S.h16_, kE, M, kE, kE, st, UVC_EPI_GELU, save ? S.hpre16 : nullptr));
    if (p > 0.f) {
      UVC_TRY(lin16(S.h16_, kE, S.m2_w16, kE, P.mlp2_b, S.z, nullptr, kE, M, kE, kE, st));
      dropout_add_kernel<<<blocks_for((long long)M * kE, 256), 256, 0, st>>>(S.z, S.y2, kE, S.out, (long long)M * kE, a.seed + 2 * s + 1, p);
      UVC_TRY(check_launch("t2t dropout_add"));
    } else {
      UVC_TRY(lin16(S.h16_, kE, S.m2_w16, kE, P.mlp2_b, S.out, nullptr, kE, M, kE, kE, st, 0, nullptr, S.y2, kE));
    }
    src = S.out;
  }
  UVC_TRY(cvt_pad(a.w.project_w, C, w.g2.dim, w.pr_w16, w.g2.Kp, save ? w.pr_wT16 : nullptr, st));
  UVC_TRY(unfold_ln(src, w.g2, nullptr, nullptr, eps, w.A2_16, w.g2.Kp, 0, nullptr, nullptr, w.M2, st));
  UVC_TRY(lin16(w.A2_16, w.g2.Kp, w.pr_w16, w.g2.Kp, a.w.project_b, a.tokens, nullptr, C, w.M2, C, w.g2.Kp, st));
  return UVC_OK;
}

int t2t_backward(const uvc_t2t_backward_args& a, cudaStream_t st) {
  UVC_TRY(check_t2t(a.dims));
  UVC_TRY(check_perf(a.w.attn1, "attention1")); UVC_TRY(check_perf(a.w.attn2, "attention2"));
  UVC_TRY(check_perf(a.g.attn1, "grad attention1")); UVC_TRY(check_perf(a.g.attn2, "grad attention2"));
  UVC_REQUIRE(a.w.project_w && a.g.project_w && a.g.project_b && a.x && a.d_tokens && a.workspace, UVC_ERR_BAD_ARG, "t2t_backward: NULL pointer");
  T2TWs w;
  carve_t2t(a.dims.B, a.dims.img, a.dims.in_chans, a.dims.C, true, a.workspace, &w);
  UVC_REQUIRE(w.bytes <= a.workspace_bytes, UVC_ERR_WORKSPACE, "t2t_backward: workspace %llu bytes < required %llu", (unsigned long long)a.workspace_bytes,
              (unsigned long long)w.bytes);
  const float p = a.dropout_p;
  const int B = a.dims.B, C = a.dims.C;
  const float* Sd = w.scales; const float* invS = w.scales + 1;
  cudaError_t e;
  e = cudaMemsetAsync(w.scales, 0, 4 * sizeof(float), st);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "t2t_backward: memset: %s", cudaGetErrorString(e));
  absmax_kernel<<<148 * 4, 256, 0, st>>>(a.d_tokens, (long long)w.M2 * C, reinterpret_cast<unsigned*>(w.scales) + 2);
  UVC_TRY(check_launch("t2t absmax"));
  pick_scale_kernel<<<1, 1, 0, st>>>(w.scales, 128.0f, a.grad_scale);
  UVC_TRY(check_launch("t2t pick_scale"));
  UVC_TRY(scale_to_f16(w.dtok16, a.d_tokens, 1.0f, (long long)w.M2 * C, st, Sd));
  // project (t2t_vit.py:105): dW += dtok^T A2 ; db += colsum(dtok) ; dA2 = dtok W
  UVC_TRY(wgrad16(w.dtok16, C, w.A2_16, w.g2.Kp, const_cast<float*>(a.g.project_w), w.g2.Kp, w.M2, C, w.g2.Kp, invS, st));
  UVC_TRY(colsum(a.d_tokens, C, w.M2, C, nullptr, const_cast<float*>(a.g.project_b), st));
  UVC_TRY(lin16(w.dtok16, C, w.pr_wT16, C, nullptr, nullptr, w.dA2_16, w.g2.Kp, w.M2, w.g2.Kp, C, st));
  // fold the last soft split into the gradient of stage 1's output
  e = cudaMemsetAsync(w.st[1].dout, 0, (size_t)w.st[1].M * kE * sizeof(float), st);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "t2t_backward: memset: %s", cudaGetErrorString(e));
  UVC_TRY(unfold_ln_bwd(w.dA2_16, nullptr, w.g2, nullptr, nullptr, nullptr, w.scales, nullptr, nullptr, w.st[1].dout, w.M2, st));
  for (int s = 1; s >= 0; --s) {
    const uvc_performer_tensors& P = s ? a.w.attn2 : a.w.attn1;
    const uvc_performer_tensors& G = s ? a.g.attn2 : a.g.attn1;
    StageWs& S = w.st[s];
    const int M = S.M, Kp = S.g.Kp, dim = S.g.dim;
    const float* src = s ? w.st[0].out : a.x;
    float* g_m2b = const_cast<float*>(G.mlp2_b); float* g_pb = const_cast<float*>(G.proj_b);
    // ---- out = y2 + drop(mlp2(gelu(mlp0(LN2(y2)))))
    dropout_grad_kernel<<<blocks_for(M, 4, 148 * 8), 256, 0, st>>>(S.dout, S.g16, g_m2b, M, a.seed + 2 * s + 1, p, w.scales);     // g16 = S drop'(dout); db2
    UVC_TRY(check_launch("t2t dropout_grad"));
    UVC_TRY(wgrad16(S.g16, kE, S.h16_, kE, const_cast<float*>(G.mlp2_w), kE, M, kE, kE, invS, st));
    UVC_TRY(lin16(S.g16, kE, S.m2_wT16, kE, nullptr, nullptr, S.dh16, kE, M, kE, kE, st, UVC_EPI_GELU_BWD, S.hpre16, nullptr, 0, const_cast<float*>(G.mlp0_b), invS));
    UVC_TRY(wgrad16(S.dh16, kE, S.ln2_16, kE, const_cast<float*>(G.mlp0_w), kE, M, kE, kE, invS, st));
    UVC_TRY(lin16(S.dh16, kE, S.m0_wT16, kE, nullptr, nullptr, S.dln2_16, kE, M, kE, kE, st));
    // dy2 = dout + LN2'(dln2)
    UVC_TRY(layernorm_bwd(nullptr, kE, S.y2, kE, S.mean2, S.rstd2, P.norm2_w, S.dout, nullptr, nullptr, S.dy2, kE, const_cast<float*>(G.norm2_w),
                          const_cast<float*>(G.norm2_b), M, kE, st, nullptr, nullptr, S.dln2_16, 1.0f, nullptr, 1.0f, w.scales));
    // ---- y2 = v + drop(proj(y))
    dropout_grad_kernel<<<blocks_for(M, 4, 148 * 8), 256, 0, st>>>(S.dy2, S.dy2_16, g_pb, M, a.seed + 2 * s, p, w.scales);
    UVC_TRY(check_launch("t2t dropout_grad"));
    UVC_TRY(wgrad16(S.dy2_16, kE, S.y16, kE, const_cast<float*>(G.proj_w), kE, M, kE, kE, invS, st));
    UVC_TRY(lin16(S.dy2_16, kE, S.proj_wT16, kE, nullptr, nullptr, S.dy16, kE, M, kE, kE, st));
    // ---- performer attention
    e = cudaMemsetAsync(S.dws, 0, (size_t)B * kWsPerImg * sizeof(float), st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "t2t_backward: memset: %s", cudaGetErrorString(e));
    const dim3 pg((S.T + 8 * kTokPerWarp - 1) / (8 * kTokPerWarp), B);
    float* g_kqvb = const_cast<float*>(G.kqv_b);
    performer_bwd_a_kernel<<<pg, 256, 0, st>>>(S.kqv, P.w, S.ws, S.y16, S.dy16, S.dws, S.dkqv16, g_kqvb, w.scales, S.T, 1e-8f);
    UVC_TRY(check_launch("t2t performer_bwd_a"));
    performer_bwd_b_kernel<<<pg, 256, 0, st>>>(S.kqv, P.w, S.dws, S.dy2, S.dkqv16, g_kqvb, w.scales, S.T);
    UVC_TRY(check_launch("t2t performer_bwd_b"));
    // ---- kqv = LN1(unfold(src)) W^T + b : the weight gradient is formed at the padded width and stripped into the [192, dim] tensor
    e = cudaMemsetAsync(S.dkqv_w, 0, (size_t)kKqv * Kp * sizeof(float), st);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "t2t_backward: memset: %s", cudaGetErrorString(e));
    UVC_TRY(wgrad16(S.dkqv16, kKqv, S.A16, 2 * Kp, S.dkqv_w, Kp, M, kKqv, Kp, invS, st));        // the hi half of the split rows
    add_strip_kernel<<<blocks_for((long long)kKqv * dim, 256), 256, 0, st>>>(S.dkqv_w, Kp, const_cast<float*>(G.kqv_w), kKqv, dim);
    UVC_TRY(check_launch("t2t add_strip"));
    UVC_TRY(lin16(S.dkqv16, kKqv, S.kqv_wT16, kKqv, nullptr, nullptr, S.dxn16, Kp, M, Kp, kKqv, st));
    float* dsrc = nullptr;
    if (s == 1) {
      dsrc = w.st[0].dout;
      e = cudaMemsetAsync(dsrc, 0, (size_t)w.st[0].M * kE * sizeof(float), st);
      UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "t2t_backward: memset: %s", cudaGetErrorString(e));
    }
    UVC_TRY(unfold_ln_bwd(S.dxn16, src, S.g, P.norm1_w, S.mean1, S.rstd1, w.scales, const_cast<float*>(G.norm1_w), const_cast<float*>(G.norm1_b), dsrc, M, st));
  }
  return UVC_OK;
}

}  // namespace uvc

extern "C" uint64_t uvc_t2t_workspace_bytes(const uvc_t2t_dims* dims, int32_t save_for_backward) {
  if (!dims) { uvc::set_error("uvc_t2t_workspace_bytes: dims is NULL"); return 0; }
  if (uvc::check_t2t(*dims)) return 0;
  return uvc::t2t_workspace_bytes(*dims, save_for_backward != 0);
}
extern "C" int uvc_t2t_forward(const uvc_t2t_forward_args* args, void* stream) {
  if (!args) { uvc::set_error("uvc_t2t_forward: args is NULL"); return UVC_ERR_BAD_ARG; }
  return uvc::t2t_forward(*args, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_t2t_backward(const uvc_t2t_backward_args* args, void* stream) {
  if (!args) { uvc::set_error("uvc_t2t_backward: args is NULL"); return UVC_ERR_BAD_ARG; }
  return uvc::t2t_backward(*args, static_cast<cudaStream_t>(stream));
}

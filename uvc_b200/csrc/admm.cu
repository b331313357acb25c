// ADMM primal-dual update of UVC on the device (replaces the Python loop nest of uvc_optimizer.py:37-144 and
// uvc_utils.py:54-73,177-269,315-471, which costs ~65k host syncs / ~1.2 s per step on DeiT-Base):
//
//   uvc_admm_scores   column sq-norms of every W1 = attn.proj.weight and W3 = mlp.fc2.weight (the one
//                     HBM-bound pass: 4*L*(C*C + C*Fh) bytes read), head sums, and the RANK of every
//                     column / head / neuron inside its group.  A selection "k smallest" is then rank < k,
//                     the (k+1)-th smallest value is the element with rank == k: no sort, no host sync.
//   uvc_admm_prox     proximal shrink of the selected columns   (uvc_utils.py:315-345)
//   uvc_admm_masks    mask rewrite from the same selections     (uvc_utils.py:376-401)
//   uvc_admm_primal   closed-form gradients of sloss1 / rloss1 / resource wrt s, r, gate; projected SGD on s, r
//                     (uvc_optimizer.py:46-123); also returns the current resource
//   uvc_admm_dual     dual ascent on y, p, z + clamp            (uvc_optimizer.py:126-135)
//   uvc_admm_resource the FLOPs model alone                     (uvc_utils.py:409-471; epoch-end prints)
//
// Ties: ranks break ties by lower index first (torch.topk's tie order on CPU is unspecified).
#include "kernels.h"

namespace uvc {

namespace {

constexpr int kMaxL = UVC_MAX_DEPTH;
constexpr int kMaxH = 32;
constexpr int kMaxFh = 4096;

struct WPtrs { float* w1[kMaxL]; float* w3[kMaxL]; float* w2[kMaxL]; };

// ------------------------------------------------------------------------------------------ scores
// grid (ceil(cols/32), L, 2), block (32, 8): column sq-norms, coalesced 128 B row segments
__global__ void __launch_bounds__(256) colnorm_kernel(const __grid_constant__ WPtrs P, int C, int Fh, float* __restrict__ c1, float* __restrict__ c3) {
  __shared__ float part[8][33];
  const int which = blockIdx.z, l = blockIdx.y;
  const int cols = which ? Fh : C;
  if ((int)blockIdx.x * 32 >= cols) return;
  const float* W = which ? P.w3[l] : P.w1[l];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f;
  if (col < cols) {
    int row = threadIdx.y;
    for (; row + 8 < C; row += 16) {
      const float v0 = W[(long long)row * cols + col], v1 = W[(long long)(row + 8) * cols + col];
      a0 += v0 * v0; a1 += v1 * v1;
    }
    for (; row < C; row += 8) { const float v = W[(long long)row * cols + col]; a0 += v * v; }
  }
  part[threadIdx.y][threadIdx.x] = a0 + a1;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
    (which ? c3 : c1)[(long long)l * cols + col] = s;
  }
}

// ranks ("k smallest" == rank < k, ties -> lower index first, as torch.topk(largest=False) on CPU).  grid (L, 1 + ceil(Fh/256)):
//   blockIdx.y == 0 : head sums (fixed order: deterministic), ranks inside each head (c1) and across heads (c2)
//   blockIdx.y >= 1 : 128 neurons of the layer each, ranked against all Fh (the O(Fh^2) part, spread over 12 blocks per layer instead of one)
constexpr int kRankPerBlock = 128;     // neurons ranked per block: 12 blocks per layer for Fh = 1536 (the grid then covers all SMs)
__global__ void __launch_bounds__(256) rank_kernel(int H, int d, int Fh, const float* __restrict__ c1, float* __restrict__ c2,
                                                   const float* __restrict__ c3, int* __restrict__ rank1, int* __restrict__ rank2,
                                                   int* __restrict__ rank3) {
  __shared__ __align__(16) float sv[kMaxFh + 4];
  __shared__ float sh[kMaxH];
  const int l = blockIdx.x, C = H * d, tid = threadIdx.x;
  if (blockIdx.y == 0) {
    for (int i = tid; i < C; i += blockDim.x) sv[i] = c1[(long long)l * C + i];
    __syncthreads();
    for (int i = tid; i < C; i += blockDim.x) {
      const int h = i / d, j = i - h * d;
      const float v = sv[i];
      int rk = 0;
      for (int m = 0; m < d; ++m) { const float u = sv[h * d + m]; rk += (u < v) || (u == v && m < j); }
      rank1[(long long)l * C + i] = rk;
    }
    if (tid < H) {        // head sums in a fixed order (deterministic)
      float s = 0.f;
      for (int m = 0; m < d; ++m) s += sv[tid * d + m];
      sh[tid] = s;
      c2[l * H + tid] = s;
    }
    __syncthreads();
    if (tid < H) {
      const float v = sh[tid];
      int rk = 0;
      for (int m = 0; m < H; ++m) rk += (sh[m] < v) || (sh[m] == v && m < tid);
      rank2[l * H + tid] = rk;
    }
    return;
  }
  for (int i = tid; i < Fh; i += blockDim.x) sv[i] = c3[(long long)l * Fh + i];
  for (int i = Fh + tid; i < ((Fh + 3) & ~3); i += blockDim.x) sv[i] = INFINITY;          // pad to a multiple of 4: +inf is never "smaller"
  __syncthreads();
  const int i = (blockIdx.y - 1) * kRankPerBlock + tid;
  if (tid < kRankPerBlock && i < Fh) {
    // rank = #{m : u_m < v} + #{m < i : u_m == v}: 128-bit broadcast reads, branch-free counting (two independent counters)
    const float v = sv[i];
    int rk0 = 0, rk1 = 0;
    const float4* sv4 = reinterpret_cast<const float4*>(sv);
    const int n4 = (Fh + 3) >> 2;
#pragma unroll 4
    for (int m4 = 0; m4 < n4; ++m4) {
      const float4 u = sv4[m4];
      const int m = m4 * 4;
      rk0 += (int)(u.x < v) + (int)((u.x == v) & (m < i)) + (int)(u.y < v) + (int)((u.y == v) & (m + 1 < i));
      rk1 += (int)(u.z < v) + (int)((u.z == v) & (m + 2 < i)) + (int)(u.w < v) + (int)((u.w == v) & (m + 3 < i));
    }
    rank3[(long long)l * Fh + i] = rk0 + rk1;
  }
}

// prox shrink (uvc_utils.py:315-345): W[:, col] *= 1/(1+2 lr p) for the bottom-ceil(r) columns of each head, then *= 1/(1+2 lr y) for the
// columns of the bottom-ceil(s0) heads (W1) / the bottom-ceil(s1) neurons (W3).  grid (ceil(cols/128), L, 2), block (32, 8): a thread owns
// 4 adjacent columns (one 128-bit access per row) and every 8th row; threads whose 4 columns are all unselected exit without touching memory.
constexpr int kProxRowSplit = 4;
__global__ void __launch_bounds__(256) prox_kernel(const __grid_constant__ WPtrs P, int H, int d, int Fh, const int* __restrict__ rank1,
                                                   const int* __restrict__ rank2, const int* __restrict__ rank3, const float* __restrict__ s,
                                                   const float* __restrict__ r, const float* __restrict__ y, const float* __restrict__ p, double lr) {
  // blockIdx.x = column group * kProxRowSplit + row slice: four times the blocks of the first version (288 blocks, 15 % occupancy, 1 TB/s)
  const int which = blockIdx.z, l = blockIdx.y, C = H * d;
  const int cols = which ? Fh : C;
  const int rs = blockIdx.x % kProxRowSplit;
  const int col0 = (blockIdx.x / kProxRowSplit) * 128 + threadIdx.x * 4;
  if (col0 >= cols) return;
  float fa[4], fb[4];
  bool any = false;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int col = col0 + e;
    fa[e] = fb[e] = 1.0f;
    if (col >= cols) continue;
    if (which == 0) {
      const int h = col / d;
      const int R = (int)ceilf(r[l * H + h]), S0 = (int)ceilf(s[l * 2 + 0]);
      if (rank1[(long long)l * C + col] < R) { fa[e] = 1.0f / (float)(1.0 + 2.0 * lr * (double)p[l * H + h]); any = true; }
      if (rank2[l * H + h] < S0) { fb[e] = 1.0f / (float)(1.0 + 2.0 * lr * (double)y[l * 2 + 0]); any = true; }
    } else {
      const int S1 = (int)ceilf(s[l * 2 + 1]);
      if (rank3[(long long)l * Fh + col] < S1) { fa[e] = 1.0f / (float)(1.0 + 2.0 * lr * (double)y[l * 2 + 1]); any = true; }
    }
  }
  if (!any) return;
  float* W = (which ? P.w3[l] : P.w1[l]) + col0;
  // the reference multiplies selected columns by the first factor, then by the second: keep the two roundings (x * 1.0f is exact)
  const int rows_per = (C + kProxRowSplit - 1) / kProxRowSplit, row_lo = rs * rows_per, row_hi = min(C, row_lo + rows_per);
#pragma unroll 4
  for (int row = row_lo + threadIdx.y; row < row_hi; row += 8) {
    float4* q = reinterpret_cast<float4*>(W + (long long)row * cols);
    float4 v = *q;
    v.x = (v.x * fa[0]) * fb[0]; v.y = (v.y * fa[1]) * fb[1]; v.z = (v.z * fa[2]) * fb[2]; v.w = (v.w * fa[3]) * fb[3];
    *q = v;
  }
}

// ------------------------------------------------------------------------------------------ prox / masks
// grid (ceil(cols/128), L, 2), block 128: one thread per column, looping over rows (coalesced across the block)
// mode 0: W[:, col] *= 1/(1+2 lr p) then *= 1/(1+2 lr y)  for selected columns (prox)
// mode 1: mask rewrite: W1/W3 masks = 1 except selected columns = 0; W2 mask rows selected = 0 (never reset, as the reference)
__global__ void __launch_bounds__(128) prox_mask_kernel(const __grid_constant__ WPtrs P, int mode, int H, int d, int Fh, const int* __restrict__ rank1,
                                                        const int* __restrict__ rank2, const int* __restrict__ rank3, const float* __restrict__ s,
                                                        const float* __restrict__ r, const float* __restrict__ y, const float* __restrict__ p,
                                                        double lr) {
  const int which = blockIdx.z, l = blockIdx.y, C = H * d;
  const int cols = which ? Fh : C;
  const int col = blockIdx.x * 128 + threadIdx.x;
  if (col >= cols) return;
  float* W = which ? P.w3[l] : P.w1[l];
  bool sel_a = false, sel_b = false;
  float inv_a = 1.f, inv_b = 1.f;
  if (which == 0) {
    const int h = col / d;
    const int R = (int)ceilf(r[l * H + h]), S0 = (int)ceilf(s[l * 2 + 0]);
    sel_a = rank1[(long long)l * C + col] < R;
    sel_b = rank2[l * H + h] < S0;
    if (mode == 0) {
      inv_a = 1.0f / (float)(1.0 + 2.0 * lr * (double)p[l * H + h]);
      inv_b = 1.0f / (float)(1.0 + 2.0 * lr * (double)y[l * 2 + 0]);
    }
  } else {
    const int S1 = (int)ceilf(s[l * 2 + 1]);
    sel_a = rank3[(long long)l * Fh + col] < S1;
    if (mode == 0) inv_a = 1.0f / (float)(1.0 + 2.0 * lr * (double)y[l * 2 + 1]);
  }
  if (mode == 0) {
    if (!sel_a && !sel_b) return;
    for (int row = 0; row < C; ++row) {
      float v = W[(long long)row * cols + col];
      if (sel_a) v *= inv_a;
      if (sel_b) v *= inv_b;
      W[(long long)row * cols + col] = v;
    }
  } else {
    const float m = (sel_a || sel_b) ? 0.f : 1.f;
    for (int row = 0; row < C; ++row) W[(long long)row * cols + col] = m;
    if (which == 1 && sel_a && P.w2[l]) {       // fc1 rows of the pruned neurons: mask[col, :] = 0
      float* M2 = P.w2[l] + (long long)col * C;
      for (int k = 0; k < C; ++k) M2[k] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------ scalar algebra
struct AdmmK {
  int L, H, d, Fh;
  const float *c1, *c2, *c3; const int *rank1, *rank2, *rank3;
  float *s, *r, *y, *p, *z;
  const float* gate; const float* gate_grad; const float* noise;
  const float* macs;         // [L,6] as fp32 (torch.Tensor(total_macs))
  float embed_macs; double full_flops;
  float budget, z_grad_clip, slr, rlr, ylr, plr, zlr, sl2wd, gating_weight, eps;
  int use_gumbel, gumbel_hard, warmup;
  float* gate_grad_acc;      // [L,2] running sum of G * (step mod interval)
  float gate_mult;           // (global_step mod interval)
  float* out;                // [0] = flops (resource / full), written by every kernel that evaluates the model
};

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float in01(float v) { return (v >= 0.f && v <= 1.f) ? 1.f : 0.f; }

// FLOPs model (uvc_utils.py:409-471) for ceiled S, R held in shared memory; thread l handles layer l, thread 0 reduces.
// Writes per-layer terms for the gradient: g (gate value), dg0/dg1 (d g / d gate), rho0, rho1, rhor and the clamp indicators.
struct LayerTerms { float g, dg0, dg1, rho0, rho1, rhor, ok0, ok1, okr, w01, w23, w45; };

__device__ float flops_model(const AdmmK& k, const float* S, const float* R, const float* noise, LayerTerms* T, float* red) {
  const int l = threadIdx.x;
  const int C = k.H * k.d;
  if (l < k.L) {
    const float S0 = S[l * 2 + 0], S1 = S[l * 2 + 1];
    const float v0 = ((float)k.H - S0) / (float)k.H, v1 = ((float)k.Fh - S1) / (float)k.Fh;
    float a = (float)C;                       // r_ub.sum(1)
    a -= S0 * (float)k.d;
    const int iS0 = (int)ceilf(S0);
    for (int h = 0; h < k.H; ++h) if (!(k.rank2[l * k.H + h] < iS0)) a -= R[l * k.H + h];
    const float vr = a / (float)C;
    float g = 1.f, dg0 = 0.f, dg1 = 0.f;
    if (k.gate) {
      const float g0 = k.gate[l * 2 + 0], g1 = k.gate[l * 2 + 1];
      if (k.use_gumbel) {
        const float u0 = (g0 + noise[l * 2 + 0]) / 0.5f, u1 = (g1 + noise[l * 2 + 1]) / 0.5f;
        const float mx = fmaxf(u0, u1);
        const float e0 = expf(u0 - mx), e1 = expf(u1 - mx);
        const float soft = e1 / (e0 + e1);
        g = soft;
        if (k.gumbel_hard) g = (u1 > u0) ? 1.f : 0.f;     // straight-through value; gradient is the soft one
        dg1 = soft * (1.f - soft) / 0.5f; dg0 = -dg1;
      } else {
        const float t = g1 * g1;
        g = t / (t + k.eps);
        dg1 = 2.f * g1 * k.eps / ((t + k.eps) * (t + k.eps)); dg0 = 0.f;
      }
    }
    LayerTerms t;
    t.g = g; t.dg0 = dg0; t.dg1 = dg1;
    t.rho0 = clamp01(v0); t.rho1 = clamp01(v1); t.rhor = clamp01(vr);
    t.ok0 = in01(v0); t.ok1 = in01(v1); t.okr = in01(vr);
    const float* m = k.macs + l * 6;
    t.w01 = m[0] + m[1]; t.w23 = m[2] + m[3]; t.w45 = m[4] + m[5];
    T[l] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // same association as the reference: embed + sum_l(tm0*rho0) + sum_l(tm1*rho0) + sum_l(tm2*rhor) + ...
    float acc = k.embed_macs;
    for (int j = 0; j < 6; ++j) {
      float sj = 0.f;
      for (int l2 = 0; l2 < k.L; ++l2) {
        const float rho = (j < 2) ? T[l2].rho0 : (j < 4 ? T[l2].rhor : T[l2].rho1);
        sj += (k.macs[l2 * 6 + j] * T[l2].g) * rho;
      }
      acc += sj;
    }
    red[0] = (float)((double)(acc * 2.0f) / k.full_flops);
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ float block_absmax(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[1 + (threadIdx.x >> 5)] = v;
  __syncthreads();
  float m = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[1 + w]);
  __syncthreads();
  return m;
}

// single block of 1024 threads (the scans over L*Fh ranks are latency-bound: more threads = fewer serial rounds)
__global__ void __launch_bounds__(1024) admm_primal_kernel(const AdmmK k) {
  __shared__ float S[kMaxL * 2], R[kMaxL * kMaxH], gs[kMaxL * 2], gr[kMaxL * kMaxH], kth2[kMaxL], kth3[kMaxL], kth1[kMaxL * kMaxH];
  __shared__ LayerTerms T[kMaxL];
  __shared__ float red[40];
  const int tid = threadIdx.x, L = k.L, H = k.H, d = k.d, Fh = k.Fh, C = H * d;
  for (int i = tid; i < L * 2; i += blockDim.x) S[i] = ceilf(k.s[i]);
  for (int i = tid; i < L * H; i += blockDim.x) R[i] = ceilf(k.r[i]);
  __syncthreads();
  const float flops = flops_model(k, S, R, k.noise, T, red);
  if (tid == 0) k.out[0] = flops;
  if (k.warmup) return;                       // uvc_optimizer.py:52-58
  // (k+1)-th smallest group norms (LeastSsum backward, uvc_utils.py:79-90): element whose rank == min(k, n-1)
  for (int i = tid; i < L * H; i += blockDim.x) {
    const int l = i / H;
    if (k.rank2[i] == min((int)S[l * 2 + 0], H - 1)) kth2[l] = k.c2[i];
  }
  for (int i = tid; i < L * Fh; i += blockDim.x) {
    const int l = i / Fh;
    if (k.rank3[i] == min((int)S[l * 2 + 1], Fh - 1)) kth3[l] = k.c3[i];
  }
  for (int i = tid; i < L * C; i += blockDim.x) {
    const int l = i / C, h = (i - l * C) / d;
    if (k.rank1[i] == min((int)R[l * H + h], d - 1)) kth1[l * H + h] = k.c1[i];
  }
  __syncthreads();
  const float z = k.z[0];
  const float sr = flops - k.budget;
  const float pass = (sr >= -k.z_grad_clip && sr <= k.z_grad_clip) ? 1.f : 0.f;     // clamp() passes gradient inside [min, max]
  const float two_over_full = (float)(2.0 / k.full_flops);
  // gradients
  for (int i = tid; i < L * 2; i += blockDim.x) {
    const int l = i >> 1, j = i & 1;
    const LayerTerms& t = T[l];
    const float ub = j ? (float)Fh : (float)H;
    const float g1 = k.y[i] * (j ? kth3[l] : kth2[l]) + k.sl2wd * (k.s[i] / ub);
    float g2;
    if (j == 0) g2 = two_over_full * t.g * (t.w01 * (-1.f / (float)H) * t.ok0 + t.w23 * (-(float)d / (float)C) * t.okr);
    else g2 = two_over_full * t.g * (t.w45 * (-1.f / (float)Fh) * t.ok1);
    gs[i] = g1 + z * (g2 * pass);
  }
  for (int i = tid; i < L * H; i += blockDim.x) {
    const int l = i / H;
    const LayerTerms& t = T[l];
    const float g1 = k.p[i] * kth1[i] + k.sl2wd * (k.r[i] / (float)d);
    const bool pruned = k.rank2[i] < (int)S[l * 2 + 0];
    const float g2 = pruned ? 0.f : two_over_full * t.g * t.w23 * (-1.f / (float)C) * t.okr;
    gr[i] = g1 + z * (g2 * pass);
  }
  // gate: G = gate.grad + z * gating_weight * d flops / d gate ; accumulate G * (step mod interval)  (uvc_optimizer.py:89-91)
  if (k.gate && k.gate_grad_acc) {
    for (int i = tid; i < L * 2; i += blockDim.x) {
      const int l = i >> 1, j = i & 1;
      const LayerTerms& t = T[l];
      const float inner = t.w01 * t.rho0 + t.w23 * t.rhor + t.w45 * t.rho1;
      const float gres = two_over_full * inner * (j ? t.dg1 : t.dg0) * pass;
      const float G = (k.gate_grad ? k.gate_grad[i] : 0.f) + z * k.gating_weight * gres;
      k.gate_grad_acc[i] += G * k.gate_mult;
    }
  }
  __syncthreads();
  // ---- s: bound handling, inf-norm clip, SGD, clamp   (uvc_optimizer.py:100-110)
  {
    float mx = 0.f;
    bool over[1] = {false};
    float smax_i = 0.f, gi = 0.f, si = 0.f;
    const int i = tid;
    if (i < L * 2) {
      const float ub = (i & 1) ? (float)Fh : (float)H;
      smax_i = fmaxf(ub - 1.f - 1e-8f, 0.f);
      si = k.s[i]; gi = gs[i];
      over[0] = si >= smax_i;
      if (over[0]) gi = fmaxf(gi, 0.f);
      if (si <= 0.f) gi = fminf(gi, 0.f);
      mx = fabsf(gi);
    }
    const float total = block_absmax(mx, red);
    const float coef = fminf(1.0f / (total + 1e-6f), 1.0f);
    if (i < L * 2) {
      float v = si - k.slr * (gi * coef);
      v = fmaxf(v, 0.f);
      if (over[0]) v = smax_i;
      k.s[i] = v;
    }
  }
  // ---- r  (uvc_optimizer.py:113-123); L*H can exceed the block: strided in two passes
  {
    const float rmax = fmaxf((float)d - 1.f - 1e-8f, 0.f);
    float mx = 0.f;
    for (int i = tid; i < L * H; i += blockDim.x) {
      const float ri = k.r[i];
      float gi = gr[i];
      if (ri >= rmax) gi = fmaxf(gi, 0.f);
      if (ri <= 0.f) gi = fminf(gi, 0.f);
      gr[i] = gi;
      mx = fmaxf(mx, fabsf(gi));
    }
    const float total = block_absmax(mx, red);
    const float coef = fminf(1.0f / (total + 1e-6f), 1.0f);
    for (int i = tid; i < L * H; i += blockDim.x) {
      const float ri = k.r[i];
      float v = ri - k.rlr * (gr[i] * coef);
      v = fmaxf(v, 0.f);
      if (ri >= rmax) v = rmax;
      k.r[i] = v;
    }
  }
}

// dual ascent with the NEW s, r (uvc_optimizer.py:126-135, uvc_utils.py:231-269,403-406)
__global__ void __launch_bounds__(1024) admm_dual_kernel(const AdmmK k) {
  __shared__ float S[kMaxL * 2], R[kMaxL * kMaxH], acc2[kMaxL], acc3[kMaxL], acc1[kMaxL * kMaxH];
  __shared__ LayerTerms T[kMaxL];
  __shared__ float red[40];
  const int tid = threadIdx.x, L = k.L, H = k.H, d = k.d, Fh = k.Fh;
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = tid; i < L * 2; i += blockDim.x) S[i] = ceilf(k.s[i]);
  for (int i = tid; i < L * H; i += blockDim.x) R[i] = ceilf(k.r[i]);
  __syncthreads();
  // sums of the k smallest norms: one warp per group, fixed lane-strided order + shuffle tree (deterministic)
  for (int l = warp; l < L; l += nw) {
    float a = 0.f;
    const int K0 = (int)S[l * 2 + 0];
    for (int h = lane; h < H; h += 32) if (k.rank2[l * H + h] < K0) a += k.c2[l * H + h];
    a = warp_sum(a);
    float b = 0.f;
    const int K1 = (int)S[l * 2 + 1];
    for (int n = lane; n < Fh; n += 32) if (k.rank3[(long long)l * Fh + n] < K1) b += k.c3[(long long)l * Fh + n];
    b = warp_sum(b);
    if (lane == 0) { acc2[l] = a; acc3[l] = b; }
  }
  for (int g = warp; g < L * H; g += nw) {
    float a = 0.f;
    const int K = (int)R[g];
    for (int j = lane; j < d; j += 32) if (k.rank1[(long long)g * d + j] < K) a += k.c1[(long long)g * d + j];
    a = warp_sum(a);
    if (lane == 0) acc1[g] = a;
  }
  __syncthreads();
  const float flops = flops_model(k, S, R, k.noise, T, red);
  for (int i = tid; i < L * 2; i += blockDim.x) k.y[i] = fmaxf(k.y[i] + k.ylr * ((i & 1) ? acc3[i >> 1] : acc2[i >> 1]), 0.f);
  for (int i = tid; i < L * H; i += blockDim.x) k.p[i] = fmaxf(k.p[i] + k.plr * acc1[i], 0.f);
  if (tid == 0) {
    k.z[0] = fmaxf(k.z[0] + k.zlr * (flops - k.budget), 0.f);
    k.out[0] = flops;
  }
}

__global__ void __launch_bounds__(256) admm_resource_kernel(const AdmmK k) {
  __shared__ float S[kMaxL * 2], R[kMaxL * kMaxH];
  __shared__ LayerTerms T[kMaxL];
  __shared__ float red[40];
  for (int i = threadIdx.x; i < k.L * 2; i += blockDim.x) S[i] = ceilf(k.s[i]);
  for (int i = threadIdx.x; i < k.L * k.H; i += blockDim.x) R[i] = ceilf(k.r[i]);
  __syncthreads();
  const float flops = flops_model(k, S, R, k.noise, T, red);
  if (threadIdx.x == 0) k.out[0] = flops;
}

int check_state(const uvc_admm_args& a, bool need_w) {
  UVC_REQUIRE(a.L > 0 && a.L <= kMaxL && a.H > 0 && a.H <= kMaxH && a.d > 0 && a.Fh > 0 && a.Fh <= kMaxFh && a.H * a.d <= kMaxFh, UVC_ERR_BAD_SHAPE,
              "admm: unsupported dims L=%d H=%d d=%d Fh=%d (limits L<=%d H<=%d Fh<=%d)", a.L, a.H, a.d, a.Fh, kMaxL, kMaxH, kMaxFh);
  UVC_REQUIRE(a.c1 && a.c2 && a.c3 && a.rank1 && a.rank2 && a.rank3, UVC_ERR_BAD_ARG, "admm: NULL score / rank workspace");
  if (need_w) {
    UVC_REQUIRE(a.w1 && a.w3, UVC_ERR_BAD_ARG, "admm: NULL weight pointer table");
    for (int l = 0; l < a.L; ++l) UVC_REQUIRE(a.w1[l] && a.w3[l], UVC_ERR_BAD_ARG, "admm: NULL weight pointer for layer %d", l);
  }
  return UVC_OK;
}

WPtrs make_ptrs(const uvc_admm_args& a, bool masks) {
  WPtrs P;
  for (int l = 0; l < kMaxL; ++l) {
    P.w1[l] = l < a.L ? (masks ? a.m1[l] : a.w1[l]) : nullptr;
    P.w3[l] = l < a.L ? (masks ? a.m3[l] : a.w3[l]) : nullptr;
    P.w2[l] = (l < a.L && masks && a.m2) ? a.m2[l] : nullptr;
  }
  return P;
}

AdmmK make_k(const uvc_admm_args& a) {
  AdmmK k;
  k.L = a.L; k.H = a.H; k.d = a.d; k.Fh = a.Fh;
  k.c1 = a.c1; k.c2 = a.c2; k.c3 = a.c3; k.rank1 = a.rank1; k.rank2 = a.rank2; k.rank3 = a.rank3;
  k.s = a.s; k.r = a.r; k.y = a.y; k.p = a.p; k.z = a.z;
  k.gate = a.gate; k.gate_grad = a.gate_grad; k.noise = a.noise; k.macs = a.macs;
  k.embed_macs = (float)a.embed_macs; k.full_flops = a.full_flops;
  k.budget = a.budget; k.z_grad_clip = a.z_grad_clip; k.slr = a.slr; k.rlr = a.rlr; k.ylr = a.ylr; k.plr = a.plr; k.zlr = a.zlr;
  k.sl2wd = a.sl2wd; k.gating_weight = a.gating_weight; k.eps = a.eps;
  k.use_gumbel = a.use_gumbel; k.gumbel_hard = a.gumbel_hard; k.warmup = a.warmup;
  k.gate_grad_acc = a.gate_grad_acc; k.gate_mult = a.gate_mult; k.out = a.out;
  return k;
}

int check_scalar(const uvc_admm_args& a) {
  UVC_REQUIRE(a.s && a.r && a.macs && a.out, UVC_ERR_BAD_ARG, "admm: NULL s / r / macs / out");
  UVC_REQUIRE(!a.gate || !a.use_gumbel || a.noise, UVC_ERR_BAD_ARG, "admm: use_gumbel needs the [L,2] Gumbel noise");
  UVC_REQUIRE(a.full_flops > 0, UVC_ERR_BAD_ARG, "admm: full_flops must be > 0");
  return UVC_OK;
}

}  // namespace

int admm_scores(const uvc_admm_args& a, cudaStream_t st) {
  int rc = check_state(a, true);
  if (rc) return rc;
  const int C = a.H * a.d;
  const int mx = a.Fh > C ? a.Fh : C;
  colnorm_kernel<<<dim3((mx + 31) / 32, a.L, 2), dim3(32, 8), 0, st>>>(make_ptrs(a, false), C, a.Fh, a.c1, a.c3);
  if ((rc = check_launch("admm colnorm"))) return rc;
  rank_kernel<<<dim3(a.L, 1 + (a.Fh + kRankPerBlock - 1) / kRankPerBlock), 256, 0, st>>>(a.H, a.d, a.Fh, a.c1, a.c2, a.c3, a.rank1, a.rank2, a.rank3);
  return check_launch("admm rank");
}

int admm_prox(const uvc_admm_args& a, cudaStream_t st) {
  int rc = check_state(a, true);
  if (rc) return rc;
  UVC_REQUIRE(a.s && a.r && a.y && a.p, UVC_ERR_BAD_ARG, "admm_prox: NULL s / r / y / p");
  const int C = a.H * a.d;
  const int mx = a.Fh > C ? a.Fh : C;
  bool vec_ok = (C % 4 == 0) && (a.Fh % 4 == 0);
  for (int l = 0; l < a.L && vec_ok; ++l) vec_ok = ((reinterpret_cast<uintptr_t>(a.w1[l]) | reinterpret_cast<uintptr_t>(a.w3[l])) & 15) == 0;
  if (vec_ok)
    prox_kernel<<<dim3(((mx + 127) / 128) * kProxRowSplit, a.L, 2), dim3(32, 8), 0, st>>>(make_ptrs(a, false), a.H, a.d, a.Fh, a.rank1, a.rank2, a.rank3, a.s, a.r, a.y, a.p, a.lr);
  else
    prox_mask_kernel<<<dim3((mx + 127) / 128, a.L, 2), 128, 0, st>>>(make_ptrs(a, false), 0, a.H, a.d, a.Fh, a.rank1, a.rank2, a.rank3, a.s, a.r, a.y,
                                                                    a.p, a.lr);
  return check_launch("admm prox");
}

int admm_masks(const uvc_admm_args& a, cudaStream_t st) {
  int rc = check_state(a, false);
  if (rc) return rc;
  UVC_REQUIRE(a.s && a.r && a.m1 && a.m3, UVC_ERR_BAD_ARG, "admm_masks: NULL s / r / mask tables");
  for (int l = 0; l < a.L; ++l) UVC_REQUIRE(a.m1[l] && a.m3[l], UVC_ERR_BAD_ARG, "admm_masks: NULL mask pointer for layer %d", l);
  const int C = a.H * a.d;
  const int mx = a.Fh > C ? a.Fh : C;
  prox_mask_kernel<<<dim3((mx + 127) / 128, a.L, 2), 128, 0, st>>>(make_ptrs(a, true), 1, a.H, a.d, a.Fh, a.rank1, a.rank2, a.rank3, a.s, a.r, nullptr,
                                                                  nullptr, 0.0);
  return check_launch("admm masks");
}

int admm_primal(const uvc_admm_args& a, cudaStream_t st) {
  int rc = check_state(a, false);
  if (rc) return rc;
  if ((rc = check_scalar(a))) return rc;
  UVC_REQUIRE(a.warmup || (a.y && a.p && a.z), UVC_ERR_BAD_ARG, "admm_primal: NULL y / p / z");
  UVC_REQUIRE(a.L * 2 <= 256, UVC_ERR_BAD_SHAPE, "admm_primal: too many layers");
  admm_primal_kernel<<<1, 1024, 0, st>>>(make_k(a));
  return check_launch("admm primal");
}

int admm_dual(const uvc_admm_args& a, cudaStream_t st) {
  int rc = check_state(a, false);
  if (rc) return rc;
  if ((rc = check_scalar(a))) return rc;
  UVC_REQUIRE(a.y && a.p && a.z, UVC_ERR_BAD_ARG, "admm_dual: NULL y / p / z");
  admm_dual_kernel<<<1, 1024, 0, st>>>(make_k(a));
  return check_launch("admm dual");
}

int admm_resource(const uvc_admm_args& a, cudaStream_t st) {
  int rc = check_state(a, false);
  if (rc) return rc;
  if ((rc = check_scalar(a))) return rc;
  admm_resource_kernel<<<1, 256, 0, st>>>(make_k(a));
  return check_launch("admm resource");
}

}  // namespace uvc

#define UVC_ADMM_ENTRY(name, fn)                                                         \
  extern "C" int name(const uvc_admm_args* args, void* stream) {                         \
    if (!args) { uvc::set_error(#name ": args is NULL"); return UVC_ERR_BAD_ARG; }       \
    return uvc::fn(*args, static_cast<cudaStream_t>(stream));                            \
  }
UVC_ADMM_ENTRY(uvc_admm_scores, admm_scores)
UVC_ADMM_ENTRY(uvc_admm_prox, admm_prox)
UVC_ADMM_ENTRY(uvc_admm_masks, admm_masks)
UVC_ADMM_ENTRY(uvc_admm_primal, admm_primal)
UVC_ADMM_ENTRY(uvc_admm_dual, admm_dual)
UVC_ADMM_ENTRY(uvc_admm_resource, admm_resource)

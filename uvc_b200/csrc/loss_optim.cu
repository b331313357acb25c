// Loss and optimiser kernels (HBM-bound):
//   * soft-target cross entropy + soft-distillation KL, forward value and d(loss)/d(logits) in one pass
//     (utils/losses.py:38-64, timm SoftTargetCrossEntropy; feeds the classifier-head backward GEMM);
//   * global gradient sq-norm, and clip + AdamW in a single read-modify-write sweep over the flat
//     parameter / gradient / moment arenas (joint_train.py:428-429).
#include "kernels.h"

namespace uvc {

// ------------------------------------------------------------------------------------------ loss
// one 256-thread block per image row; NC classes are strided over the block.
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < nw; ++w) r = is_max ? fmaxf(r, sh[w]) : r + sh[w];
  return r;
}

__global__ void __launch_bounds__(256) distill_loss_kernel(const float* __restrict__ logits, const float* __restrict__ teacher,
                                                           const float* __restrict__ targets, int B, int NC, float alpha, float T,
                                                           float grad_scale, float* __restrict__ loss_out, float* __restrict__ dlogits) {
  __shared__ float sh[8];
  const int b = blockIdx.x;
  const float* s = logits + (long long)b * NC;
  const float* y = targets + (long long)b * NC;
  const float* t = teacher ? teacher + (long long)b * NC : nullptr;
  const float invT = 1.0f / T;
  const bool kd = (t != nullptr);
  const float wbase = kd ? (1.0f - alpha) : 1.0f;

  // pass 1: maxima
  float ms = -INFINITY, mt = -INFINITY;
  for (int c = threadIdx.x; c < NC; c += blockDim.x) { ms = fmaxf(ms, s[c]); if (kd) mt = fmaxf(mt, t[c]); }
  ms = block_reduce(ms, sh, true);
  if (kd) mt = block_reduce(mt, sh, true);
  // pass 2: partition functions (at temperature 1 for the base loss and at T for the KD term), sum of targets
  float z1 = 0.f, zs = 0.f, zt = 0.f, ysum = 0.f;
  for (int c = threadIdx.x; c < NC; c += blockDim.x) {
    const float sv = s[c] - ms;
    z1 += expf(sv);
    ysum += y[c];
    if (kd) { zs += expf(sv * invT); zt += expf((t[c] - mt) * invT); }
  }
  z1 = block_reduce(z1, sh, false);
  ysum = block_reduce(ysum, sh, false);
  if (kd) { zs = block_reduce(zs, sh, false); zt = block_reduce(zt, sh, false); }
  const float lz1 = logf(z1), lzs = kd ? logf(zs) : 0.f, lzt = kd ? logf(zt) : 0.f;
  // pass 3: loss terms and gradient
  const float gb = wbase / (float)B;                          // d base / d s = gb * (softmax(s) * sum(y) - y)
  const float gk = kd ? alpha * T / ((float)B * (float)NC) : 0.f;   // d kd / d s = gk * (p_s(T) - p_t(T))
  float base = 0.f, kl = 0.f;
  for (int c = threadIdx.x; c < NC; c += blockDim.x) {
    const float sv = s[c] - ms;
    const float ls1 = sv - lz1;
    const float yv = y[c];
    base -= yv * ls1;
    float g = gb * (expf(ls1) * ysum - yv);
    if (kd) {
      const float ls = sv * invT - lzs, lt = (t[c] - mt) * invT - lzt;
      const float pt = expf(lt);
      kl += pt * (lt - ls);
      g += gk * (expf(ls) - pt);
    }
    if (dlogits) dlogits[(long long)b * NC + c] = g * grad_scale;
  }
  base = block_reduce(base, sh, false);
  if (kd) kl = block_reduce(kl, sh, false);
  if (threadIdx.x == 0) {
    const float lb = base / (float)B;
    const float lk = kd ? kl * T * T / ((float)B * (float)NC) : 0.f;
    atomicAdd(loss_out + 1, lb);
    atomicAdd(loss_out + 2, lk);
    atomicAdd(loss_out + 0, wbase * lb + (kd ? alpha * lk : 0.f));
  }
}

int distill_loss(const float* logits, const float* teacher, const float* targets, int B, int NC, float alpha, float T, float grad_scale,
                 float* loss_out, float* dlogits, cudaStream_t st) {
  UVC_REQUIRE(B > 0 && NC > 0, UVC_ERR_BAD_SHAPE, "distill_loss: bad B=%d NC=%d", B, NC);
  UVC_REQUIRE(T > 0.f, UVC_ERR_BAD_ARG, "distill_loss: temperature must be > 0");
  cudaError_t e = cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), st);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "distill_loss: memset failed: %s", cudaGetErrorString(e));
  distill_loss_kernel<<<B, 256, 0, st>>>(logits, teacher, targets, B, NC, alpha, T, grad_scale, loss_out, dlogits);
  return check_launch("distill_loss");
}

// ------------------------------------------------------------------------------------------ optimiser
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ acc) {
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const bool aligned = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if (aligned) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = g4[i];
      s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
  }
  __shared__ float sh[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    atomicAdd(acc, t);
  }
}

struct AdamHyper { float max_norm, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt; };

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float mk, float coef, const AdamHyper& h) {
  g *= coef;
  p *= 1.0f - h.lr * h.wd;
  m = m + (g - m) * (1.0f - h.beta1);                  // torch: exp_avg.lerp_(grad, 1 - beta1)
  v = v * h.beta2 + (1.0f - h.beta2) * g * g;
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p -= (h.lr / h.bc1) * (m / denom);
  p *= mk;
}

__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                         const float* __restrict__ mask, long long n, const float* __restrict__ acc, AdamHyper h) {
  float coef = 1.0f;
  if (h.max_norm > 0.f && acc) {
    const float nrm = sqrtf(__ldg(acc));
    coef = fminf(h.max_norm / (nrm + 1e-6f), 1.0f);
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
  const long long n4 = aligned ? (n >> 2) : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float4 kv = mask ? reinterpret_cast<const float4*>(mask)[i] : make_float4(1.f, 1.f, 1.f, 1.f);
    adam_one(pv.x, gv.x, mv.x, vv.x, kv.x, coef, h); adam_one(pv.y, gv.y, mv.y, vv.y, kv.y, coef, h);
    adam_one(pv.z, gv.z, mv.z, vv.z, kv.z, coef, h); adam_one(pv.w, gv.w, mv.w, vv.w, kv.w, coef, h);
    reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(g)[i] = gv;
    reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pv = p[i], gv = g[i], mv = m[i], vv = v[i];
    adam_one(pv, gv, mv, vv, mask ? mask[i] : 1.0f, coef, h);
    p[i] = pv; g[i] = gv; m[i] = mv; v[i] = vv;
  }
}

// Per-element option bytes for the flat-arena sweep (one launch for the whole model also in Stage 2 and with frozen tenants):
//   bit 0  keep   : 0 = the weight is pruned (mask 0): it is updated like any other (moments included) and then forced back to exactly 0
//   bit 1  decay  : decoupled weight decay applies (timm's param groups: matrices yes, biases / norms / tokens no)
//   bit 2  active : 0 = no gradient this step (frozen parameter, hard-skipped block): p, m, v untouched, excluded from the clip norm
__global__ void __launch_bounds__(256) sqnorm_flags_kernel(const float* __restrict__ g, const uint8_t* __restrict__ flags, long long n, float* __restrict__ acc) {
  float s = 0.f;
  const long long n4 = n >> 2;        // arenas are 16-byte aligned and padded to 4 elements
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    const uchar4 f = reinterpret_cast<const uchar4*>(flags)[i];
    s += ((f.x & 4) ? v.x * v.x : 0.f) + ((f.y & 4) ? v.y * v.y : 0.f) + ((f.z & 4) ? v.z * v.z : 0.f) + ((f.w & 4) ? v.w * v.w : 0.f);
  }
  __shared__ float sh[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    atomicAdd(acc, t);
  }
}

__device__ __forceinline__ void adam_one_f(float& p, float& g, float& m, float& v, unsigned f, float coef, const AdamHyper& h) {
  if (!(f & 4u)) return;
  g *= coef;
  if (f & 2u) p *= 1.0f - h.lr * h.wd;
  m = m + (g - m) * (1.0f - h.beta1);
  v = v * h.beta2 + (1.0f - h.beta2) * g * g;
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p -= (h.lr / h.bc1) * (m / denom);
  if (!(f & 1u)) p = 0.0f;
}

__global__ void __launch_bounds__(256) clip_adamw_flags_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                               const uint8_t* __restrict__ flags, long long n, const float* __restrict__ acc, AdamHyper h) {
  float coef = 1.0f;
  if (h.max_norm > 0.f && acc) {
    const float nrm = sqrtf(__ldg(acc));
    coef = fminf(h.max_norm / (nrm + 1e-6f), 1.0f);
  }
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uchar4 f = reinterpret_cast<const uchar4*>(flags)[i];
    if (!((f.x | f.y | f.z | f.w) & 4)) continue;          // whole vector inactive (a skipped block's weights): no traffic beyond the flag bytes
    float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    adam_one_f(pv.x, gv.x, mv.x, vv.x, f.x, coef, h); adam_one_f(pv.y, gv.y, mv.y, vv.y, f.y, coef, h);
    adam_one_f(pv.z, gv.z, mv.z, vv.z, f.z, coef, h); adam_one_f(pv.w, gv.w, mv.w, vv.w, f.w, coef, h);
    reinterpret_cast<float4*>(p)[i] = pv; reinterpret_cast<float4*>(g)[i] = gv;
    reinterpret_cast<float4*>(m)[i] = mv; reinterpret_cast<float4*>(v)[i] = vv;
  }
}

static inline int blocks_for(long long n, int per_thread_elems) {
  long long b = (n / per_thread_elems + 255) / 256;
  const long long cap = 148 * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int sqnorm_accum(const float* g, long long n, float* acc, cudaStream_t st) {
  if (n <= 0) return UVC_OK;
  sqnorm_kernel<<<blocks_for(n, 4), 256, 0, st>>>(g, n, acc);
  return check_launch("sqnorm_accum");
}

int clip_adamw(float* p, float* g, float* m, float* v, const float* mask, long long n, const float* acc, float max_norm, float lr, float beta1,
               float beta2, float eps, float wd, int step, cudaStream_t st) {
  UVC_REQUIRE(step >= 1, UVC_ERR_BAD_ARG, "clip_adamw: step must be >= 1 (got %d)", step);
  if (n <= 0) return UVC_OK;
  AdamHyper h;
  h.max_norm = max_norm; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = wd;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  clip_adamw_kernel<<<blocks_for(n, 4), 256, 0, st>>>(p, g, m, v, mask, n, acc, h);
  return check_launch("clip_adamw");
}

int sqnorm_accum_flags(const float* g, const uint8_t* flags, long long n, float* acc, cudaStream_t st) {
  UVC_REQUIRE((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(flags)) & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0,
              UVC_ERR_BAD_SHAPE, "sqnorm_accum_flags: arenas must be 16 B aligned with a length that is a multiple of 4");
  if (n <= 0) return UVC_OK;
  sqnorm_flags_kernel<<<blocks_for(n, 4), 256, 0, st>>>(g, flags, n, acc);
  return check_launch("sqnorm_accum_flags");
}

int clip_adamw_flags(float* p, float* g, float* m, float* v, const uint8_t* flags, long long n, const float* acc, float max_norm, float lr, float beta1,
                     float beta2, float eps, float wd, int step, cudaStream_t st) {
  UVC_REQUIRE(step >= 1, UVC_ERR_BAD_ARG, "clip_adamw_flags: step must be >= 1 (got %d)", step);
  UVC_REQUIRE((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(flags) & 3) == 0, UVC_ERR_BAD_SHAPE, "clip_adamw_flags: arenas must be 16 B aligned with a length that is a multiple of 4");
  if (n <= 0) return UVC_OK;
  AdamHyper h;
  h.max_norm = max_norm; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.wd = wd;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  clip_adamw_flags_kernel<<<blocks_for(n, 4), 256, 0, st>>>(p, g, m, v, flags, n, acc, h);
  return check_launch("clip_adamw_flags");
}

}  // namespace uvc

extern "C" int uvc_sqnorm_accum_flags(const float* g, const uint8_t* flags, int64_t n, float* acc, void* stream) {
  UVC_REQUIRE(g && flags && acc, UVC_ERR_BAD_ARG, "uvc_sqnorm_accum_flags: NULL pointer");
  return uvc::sqnorm_accum_flags(g, flags, n, acc, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_clip_adamw_flags(float* p, float* g, float* m, float* v, const uint8_t* flags, int64_t n, const float* sqnorm_acc, float max_norm, float lr,
                                    float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  UVC_REQUIRE(p && g && m && v && flags, UVC_ERR_BAD_ARG, "uvc_clip_adamw_flags: NULL pointer");
  return uvc::clip_adamw_flags(p, g, m, v, flags, n, sqnorm_acc, max_norm, lr, beta1, beta2, eps, weight_decay, step, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_distill_loss(const float* logits, const float* teacher_logits, const float* targets, int32_t B, int32_t NC, float alpha, float T,
                                float grad_scale, float* loss_out, float* dlogits, void* stream) {
  UVC_REQUIRE(logits && targets && loss_out, UVC_ERR_BAD_ARG, "uvc_distill_loss: NULL pointer");
  return uvc::distill_loss(logits, teacher_logits, targets, B, NC, alpha, T, grad_scale, loss_out, dlogits, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_sqnorm_accum(const float* g, int64_t n, float* acc, void* stream) {
  UVC_REQUIRE(g && acc, UVC_ERR_BAD_ARG, "uvc_sqnorm_accum: NULL pointer");
  return uvc::sqnorm_accum(g, n, acc, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_clip_adamw(float* p, float* g, float* m, float* v, const float* mask, int64_t n, const float* sqnorm_acc, float max_norm, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  UVC_REQUIRE(p && g && m && v, UVC_ERR_BAD_ARG, "uvc_clip_adamw: NULL pointer");
  return uvc::clip_adamw(p, g, m, v, mask, n, sqnorm_acc, max_norm, lr, beta1, beta2, eps, weight_decay, step, static_cast<cudaStream_t>(stream));
}

// On-device input pipeline of the training step: timm's batch-mode Mixup / CutMix plus the smoothed, mixed one-hot targets as ONE launch
// (reference call sites joint_train.py:409,930-933 and post_train.py:362,618-621; timm.data.Mixup(mode='batch'), restated from its public
// semantics -- timm is an un-vendored dependency of the reference).  lambda and the CutMix box are drawn on the HOST from numpy's RNG exactly
// as timm does; this kernel applies them:
//   mixup :  x[i] <- lam x[i] + (1 - lam) x[B-1-i]                       (in place; a thread owns both members of a pair)
//   cutmix:  x[i][:, yl:yh, xl:xh] <- x[B-1-i][:, yl:yh, xl:xh]          (in place, only the box is touched)
//   target[i, c] = lam * smooth_onehot(y[i])[c] + (1 - lam) * smooth_onehot(y[B-1-i])[c],  smooth_onehot = off + (on - off) [c == y]
// HBM-bound: mixup moves 2 x 4 B per pixel value (77 MB read + 77 MB written for 128 x 3 x 224 x 224), the targets 4 B per class.
#include "kernels.h"

namespace uvc {

namespace {

__global__ void __launch_bounds__(256) mixup_kernel(float* __restrict__ x, const long long* __restrict__ y, float* __restrict__ tgt, int B, int Cc, int Hh,
                                                    int Ww, int NC, float lam, float on, float off, int use_cutmix, int yl, int yh, int xl, int xh,
                                                    int do_mix) {
  if (blockIdx.y == 1) {                                   // targets
    const long long total = (long long)B * NC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int b = (int)(i / NC), c = (int)(i % NC);
      const float a = (y[b] == c) ? on : off, f = (y[B - 1 - b] == c) ? on : off;
      tgt[i] = a * lam + f * (1.0f - lam);
    }
    return;
  }
  if (!do_mix) return;
  const long long img = (long long)Cc * Hh * Ww;
  const int pairs = B >> 1;
  if (!use_cutmix) {
    const long long img4 = img >> 2, total = (long long)pairs * img4;      // img % 4 == 0 (checked by the launcher)
    const float mu = 1.0f - lam;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long p = i / img4, e = i % img4;
      float4* pa = reinterpret_cast<float4*>(x + p * img) + e;
      float4* pb = reinterpret_cast<float4*>(x + (long long)(B - 1 - p) * img) + e;
      const float4 a = *pa, b = *pb;
      // timm: x_flipped = x.flip(0).mul_(1 - lam); x.mul_(lam).add_(x_flipped)  -- two roundings of the products, one of the sum
      *pa = make_float4(a.x * lam + b.x * mu, a.y * lam + b.y * mu, a.z * lam + b.z * mu, a.w * lam + b.w * mu);
      *pb = make_float4(b.x * lam + a.x * mu, b.y * lam + a.y * mu, b.z * lam + a.z * mu, b.w * lam + a.w * mu);
    }
  } else {
    const int bh = yh - yl, bw = xh - xl;
    if (bh <= 0 || bw <= 0) return;
    const long long box = (long long)Cc * bh * bw, total = (long long)pairs * box;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      long long t = i;
      const int xx = (int)(t % bw); t /= bw;
      const int yy = (int)(t % bh); t /= bh;
      const int c = (int)(t % Cc); t /= Cc;
      const long long o = ((long long)c * Hh + (yl + yy)) * Ww + xl + xx;
      float* pa = x + t * img + o;
      float* pb = x + (long long)(B - 1 - t) * img + o;
      const float a = *pa, b = *pb;
      *pa = b; *pb = a;
    }
  }
}

}  // namespace

int mixup_batch(float* x, const long long* y, float* targets, int B, int C, int H, int W, int NC, float lam, float smoothing, int use_cutmix, int yl, int yh,
                int xl, int xh, cudaStream_t st) {
  UVC_REQUIRE(B > 0 && (B & 1) == 0, UVC_ERR_BAD_SHAPE, "mixup: batch size %d must be even and positive", B);
  UVC_REQUIRE(((long long)C * H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, UVC_ERR_BAD_SHAPE, "mixup: images must be 16 B aligned with C*H*W %% 4 == 0");
  UVC_REQUIRE(!use_cutmix || (0 <= yl && yl <= yh && yh <= H && 0 <= xl && xl <= xh && xh <= W), UVC_ERR_BAD_ARG, "mixup: cutmix box out of range");
  const float off = smoothing / (float)NC, on = 1.0f - smoothing + off;
  const int do_mix = lam != 1.0f;
  mixup_kernel<<<dim3(148 * 8, targets ? 2 : 1), 256, 0, st>>>(x, y, targets, B, C, H, W, NC, lam, on, off, use_cutmix, yl, yh, xl, xh, do_mix);
  return check_launch("mixup");
}

}  // namespace uvc

extern "C" int uvc_mixup(float* x, const int64_t* y, float* targets, int32_t B, int32_t C, int32_t H, int32_t W, int32_t num_classes, float lam,
                         float smoothing, int32_t use_cutmix, int32_t yl, int32_t yh, int32_t xl, int32_t xh, void* stream) {
  UVC_REQUIRE(x && (!targets || y), UVC_ERR_BAD_ARG, "uvc_mixup: NULL pointer");
  return uvc::mixup_batch(x, reinterpret_cast<const long long*>(y), targets, B, C, H, W, num_classes, lam, smoothing, use_cutmix, yl, yh, xl, xh,
                          static_cast<cudaStream_t>(stream));
}

// Stage-2 physical compaction, device side (reference: post_train.py:357-360 re-masks dense weights before every step and the
// model multiplies by the zeros; the hard block skip is models/model_distilled.py:496-500).
//
// The engine keeps the reference's dense parameter / gradient tensors (state dicts, the optimiser and the ADMM code see what they always saw)
// and compacts on the fly, where the bytes are touched anyway:
//   * gather_cvt_kernel  -- the per-forward fp32 -> fp16 weight conversion reads only the live rows / columns (uvc_vit_layout) and writes
//                           compact operand copies [live_out, live_in] (+ their transposes for the data-gradient GEMMs); the live entries of
//                           qkv.bias / fc1.bias are gathered in the same launch (fp32).
//   * scatter_add_kernel -- after a block's backward, the compact weight / bias gradients (fp32 scratch, split-K atomics land there) are
//                           added into the dense gradient tensors at the live positions.  The map is injective: no atomics.
// Both are HBM-bound sweeps over <= the dense weight bytes of one block (Base: 28 MB fp32), a few microseconds.
#include "kernels.h"

namespace uvc {

namespace {

__device__ __forceinline__ int amap(const AxisMap& m, int i) {
  if (!m.idx) return i;
  const int s = i / m.per, w = i - s * m.per;
  const int g = w / m.group;
  return s * m.sect_stride + __ldg(m.idx + g) * m.group + (w - g * m.group);
}

constexpr int kSegsPerLaunch = 24;
struct GatherSegs { GatherSeg s[kSegsPerLaunch]; };
struct ScatterSegs { ScatterSeg s[kSegsPerLaunch]; };

// 32 x 32 tiles through shared memory so that the transposed fp16 copy is written coalesced as well (as cvt_f16_segs_kernel)
__global__ void __launch_bounds__(256) gather_cvt_kernel(const __grid_constant__ GatherSegs segs) {
  __shared__ float tile[32][33];
  const GatherSeg& g = segs.s[blockIdx.y];
  const int rows = g.rows, cols = g.cols;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tc = (cols + 31) / 32, tr = (rows + 31) / 32;
  for (int t = blockIdx.x; t < tc * tr; t += gridDim.x) {
    const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
    const int c = c0 + tx;
    const int sc = c < cols ? amap(g.cmap, c) : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + i * 8;
      float v = 0.f;
      if (r < rows && c < cols) {
        v = g.src[(long long)amap(g.rmap, r) * g.src_ld + sc];
        if (g.dst16) g.dst16[(long long)r * cols + c] = __float2half_rn(v);
        if (g.dst32) g.dst32[(long long)r * cols + c] = v;
      }
      tile[ty + i * 8][tx] = v;
    }
    if (g.dstT16) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cc = c0 + ty + i * 8, r = r0 + tx;
        if (r < rows && cc < cols) g.dstT16[(long long)cc * rows + r] = __float2half_rn(tile[tx][ty + i * 8]);
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) scatter_add_kernel(const __grid_constant__ ScatterSegs segs) {
  const ScatterSeg& g = segs.s[blockIdx.y];
  const long long n = (long long)g.rows * g.cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / g.cols), c = (int)(i - (long long)r * g.cols);
    float* d = g.dst + (long long)amap(g.rmap, r) * g.dst_ld + amap(g.cmap, c);
    *d += g.src[i];
  }
}

// Gradient of the masked fc2 columns, in closed form.  A pruned neuron n has a zeroed fc1 row (W2 mask) and a zeroed fc2 column (W3 mask,
// uvc_utils.py:376-401), so its activation is the constant gelu(fc1.bias[n]) on every token and autograd gives the reference
//   d fc2.weight[:, n] = gelu(fc1.bias[n]) * colsum(dY) = gelu(fc1.bias[n]) * d fc2.bias      (d fc1.weight[n, :] = 0, d fc1.bias[n] = 0):
// a gradient the reference's clip_grad_norm_ counts although the weight is re-masked to zero.  The compacted block never forms that
// activation; this kernel adds the rank-1 term for the neurons outside the live list and folds this call's d fc2.bias (kept apart in
// scratch so that gradient accumulation over several backward calls stays correct) into the dense bias gradient.
__global__ void __launch_bounds__(256) pruned_fc2_grad_kernel(float* __restrict__ dW, long long ldw, float* __restrict__ db, const float* __restrict__ db_call,
                                                              const float* __restrict__ fc1_b, const int* __restrict__ dead, int n_dead, int C) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = blockIdx.y * 32, c1 = min(C, c0 + 32);
  if (blockIdx.x == 0 && threadIdx.x < 32 && c0 + (int)threadIdx.x < c1) db[c0 + threadIdx.x] += db_call[c0 + threadIdx.x];
  if (j >= n_dead) return;
  const int n = __ldg(dead + j);
  const float b = __ldg(fc1_b + n);
  const float ge = 0.5f * b * (1.0f + erff(b * 0.70710678118654752f));
  for (int c = c0; c < c1; ++c) dW[(long long)c * ldw + n] += ge * __ldg(db_call + c);
}

}  // namespace

int pruned_fc2_grad(float* dW, long long ldw, float* db, const float* db_call, const float* fc1_b, const int* dead, int n_dead, int C, cudaStream_t st) {
  UVC_REQUIRE(dW && db && db_call && fc1_b && (dead || n_dead == 0) && C > 0 && n_dead >= 0, UVC_ERR_BAD_ARG, "pruned_fc2_grad: bad arguments");
  const int bx = n_dead > 0 ? (n_dead + 255) / 256 : 1;
  pruned_fc2_grad_kernel<<<dim3(bx, (C + 31) / 32), 256, 0, st>>>(dW, ldw, db, db_call, fc1_b, dead, n_dead, C);
  return check_launch("pruned_fc2_grad");
}

int gather_cvt(const GatherSeg* segs, int nseg, cudaStream_t st) {
  for (int base = 0; base < nseg; base += kSegsPerLaunch) {
    GatherSegs k;
    const int n = nseg - base < kSegsPerLaunch ? nseg - base : kSegsPerLaunch;
    long long mx = 0;
    for (int i = 0; i < n; ++i) {
      k.s[i] = segs[base + i];
      const GatherSeg& g = k.s[i];
      UVC_REQUIRE(g.src && g.rows >= 0 && g.cols >= 0 && g.src_ld > 0, UVC_ERR_BAD_ARG, "gather_cvt: bad segment %d", base + i);
      UVC_REQUIRE((!g.rmap.idx || (g.rmap.per > 0 && g.rmap.group > 0)) && (!g.cmap.idx || (g.cmap.per > 0 && g.cmap.group > 0)), UVC_ERR_BAD_ARG,
                  "gather_cvt: bad index map in segment %d", base + i);
      const long long tiles = (long long)((g.rows + 31) / 32) * ((g.cols + 31) / 32);
      if (tiles > mx) mx = tiles;
    }
    if (mx == 0) continue;
    gather_cvt_kernel<<<dim3((unsigned)(mx < 96 ? mx : 96), n), 256, 0, st>>>(k);
    int rc = check_launch("gather_cvt");
    if (rc) return rc;
  }
  return UVC_OK;
}

int scatter_add(const ScatterSeg* segs, int nseg, cudaStream_t st) {
  for (int base = 0; base < nseg; base += kSegsPerLaunch) {
    ScatterSegs k;
    const int n = nseg - base < kSegsPerLaunch ? nseg - base : kSegsPerLaunch;
    long long mx = 0;
    for (int i = 0; i < n; ++i) {
      k.s[i] = segs[base + i];
      const ScatterSeg& g = k.s[i];
      UVC_REQUIRE(g.src && g.dst && g.rows >= 0 && g.cols >= 0 && g.dst_ld > 0, UVC_ERR_BAD_ARG, "scatter_add: bad segment %d", base + i);
      const long long e = (long long)g.rows * g.cols;
      if (e > mx) mx = e;
    }
    if (mx == 0) continue;
    long long blocks = (mx + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148) blocks = 148;
    scatter_add_kernel<<<dim3((unsigned)blocks, n), 256, 0, st>>>(k);
    int rc = check_launch("scatter_add");
    if (rc) return rc;
  }
  return UVC_OK;
}

}  // namespace uvc

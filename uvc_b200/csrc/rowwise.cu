// Bandwidth-bound row-wise kernels of the ViT block: LayerNorm fwd/bwd, softmax fwd/bwd over
// attention rows, bias-gradient column sums, gate blend + gate-gradient dot products, im2col for the
// 16x16/16 patch-embed conv, token assembly (cls + pos-embed + patch gate).  All are warp-per-row or
// thread-per-column with float4 accesses; none of them reshapes work to reach the tensor cores.
//
// Reference spans: models/model_distilled.py:199,204,288,507 (LayerNorm), :179-181 (softmax),
// :433-471 (patch embed, gates, cls/pos), :477-503 (block gate blend).
#include <cstdlib>
#include <cuda_fp16.h>
#include "kernels.h"

namespace uvc {

constexpr int kMaxVec = 8;   // float4 per lane -> C <= 1024

// The LayerNorm kernels are templated on NV = float4 per lane (2: C <= 256, 3: C <= 384, 6: C <= 768, 8: C <= 1024) so the per-row
// register arrays are sized for the model at hand: with a fixed 8 the backward needed ~200 registers (one 256-thread block per SM) and ran
// at a quarter of HBM bandwidth on DeiT-Small.
#define UVC_LN_DISPATCH(C, CALL)                         \
  do {                                                   \
    if ((C) <= 256) { constexpr int NV = 2; CALL; }      \
    else if ((C) <= 384) { constexpr int NV = 3; CALL; } \
    else if ((C) <= 768) { constexpr int NV = 6; CALL; } \
    else { constexpr int NV = 8; CALL; }                 \
  } while (0)

// ------------------------------------------------------------------------------------------ LayerNorm fwd
// one warp per row; two-pass moments in registers (mean, then centred variance) like ATen's RowwiseMoments.
// Y16: the output is written as fp16 (the operand storage format of the kind::f16 GEMM that consumes it) instead of fp32
template <int NV, bool Y16>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float* __restrict__ y, long long ldy,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int C, int rnd) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * ldx);
  const int nv = C >> 2;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) { v[i] = xr[c]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  float4* yr = reinterpret_cast<float4*>(y + (long long)row * ldy);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv) {
      const float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (Y16) {
        reinterpret_cast<uint2*>(reinterpret_cast<__half*>(y) + (long long)row * ldy)[c] = pack_half4(o.x, o.y, o.z, o.w);
      } else {
        if (rnd) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        yr[c] = o;
      }
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm bwd
// dx[row] = r1[row] + s2 * r2[row] + rstd * (g - mean(g) - xhat * mean(g * xhat)),   g = dy * gamma
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (register partials per lane -> smem -> atomics)
// Optional fused bias gradients of the neighbouring Linears (they are column sums of tensors this kernel touches anyway):
//   cs_r1[col] += sum_rows (r1 + s2 * r2)[row, col]      cs_out[col] += sum_rows dx[row, col]
// DY16: dy holds fp16 values that carry the loss scale (dy_scale = 1 / scale takes it out on load).  dx16 (optional): an fp16 copy of
// out_scale * dx for the GEMMs that consume the stream gradient as an operand; the fp32 dx stays unscaled.
template <int NV, bool CS, bool DY16>
#ifndef UVC_LN_BWD_MINB
#define UVC_LN_BWD_MINB 3
#endif
__global__ void __launch_bounds__(256, (NV <= 3 ? UVC_LN_BWD_MINB : 2)) layernorm_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ r1, const float* __restrict__ r2,
                                                            const float* __restrict__ s2_dev, float* __restrict__ dx, long long lddx,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ cs_r1,
                                                            float* __restrict__ cs_out, int M, int C, int rows_per_block,
                                                            float dy_scale, __half* __restrict__ dx16, float out_scale,
                                                            const float* __restrict__ scales_dev, const float* __restrict__ dot_t, float* __restrict__ dots,
                                                            int dot_x) {
  __shared__ float red[4][8][32 * 4 + 4];
  pdl_launch_dependents();
  pdl_wait();
  if (scales_dev) { out_scale *= __ldg(scales_dev); dy_scale *= __ldg(scales_dev + 1); }     // {S, 1/S} chosen on the device (see grad_scale_kernel)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nv = C >> 2;
  const float s2 = (r2 && s2_dev) ? __ldg(s2_dev) : 1.0f;
  float4 ag[NV], ab[NV], a1[CS ? NV : 1], ao[CS ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  if (CS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { a1[i] = make_float4(0, 0, 0, 0); ao[i] = make_float4(0, 0, 0, 0); }
  }
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(M, row0 + rows_per_block);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  // block-gate gradients ride along (models/model_distilled.py:493: x_out = d1 t + d0 x -> dd0 = <g, x>, dd1 = <g, t>): r2 is g, x is this
  // norm's input (norm1: the block input) and dot_t the un-blended block output t (norm2), all streamed here anyway except t
  float acc_dx = 0.f, acc_dt = 0.f;
  for (int row = row0 + warp; row < row1; row += nwarps) {
    const float4* dyr = reinterpret_cast<const float4*>(dy + (long long)row * lddy);
    const uint2* dyr16 = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(dy) + (long long)row * lddy);
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * ldx);
    const float4* r1r = r1 ? reinterpret_cast<const float4*>(r1 + (long long)row * lddx) : nullptr;
    const float4* r2r = r2 ? reinterpret_cast<const float4*>(r2 + (long long)row * lddx) : nullptr;
    const float4* dtr = dot_t ? reinterpret_cast<const float4*>(dot_t + (long long)row * lddx) : nullptr;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], gg[NV], res[NV];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {                      // every load of the row is issued before the first reduction
      const int c = lane + i * 32;
      res[i] = make_float4(0, 0, 0, 0);
      if (c < nv) {
        float4 d;
        if (DY16) {
          const uint2 u = dyr16[c];
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
          d = make_float4(lo.x * dy_scale, lo.y * dy_scale, hi.x * dy_scale, hi.y * dy_scale);
        } else {
          d = dyr[c];
        }
        const float4 xv = xr[c], gm = __ldg(g4 + c);
        if (r1r) res[i] = r1r[c];
        if (r2r) {
          const float4 a = r2r[c];
          res[i].x += s2 * a.x; res[i].y += s2 * a.y; res[i].z += s2 * a.z; res[i].w += s2 * a.w;
          if (dot_x) acc_dx += (a.x * xv.x + a.y * xv.y) + (a.z * xv.z + a.w * xv.w);
          if (dtr) { const float4 tv = dtr[c]; acc_dt += (a.x * tv.x + a.y * tv.y) + (a.z * tv.z + a.w * tv.w); }
        }
        if (CS) { a1[i].x += res[i].x; a1[i].y += res[i].y; a1[i].z += res[i].z; a1[i].w += res[i].w; }
        xh[i].x = (xv.x - mu) * rs; xh[i].y = (xv.y - mu) * rs; xh[i].z = (xv.z - mu) * rs; xh[i].w = (xv.w - mu) * rs;
        gg[i].x = d.x * gm.x; gg[i].y = d.y * gm.y; gg[i].z = d.z * gm.z; gg[i].w = d.w * gm.w;
        sg += (gg[i].x + gg[i].y) + (gg[i].z + gg[i].w);
        sgx += (gg[i].x * xh[i].x + gg[i].y * xh[i].y) + (gg[i].z * xh[i].z + gg[i].w * xh[i].w);
        ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
        ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      }
    }
    const float mg = warp_sum(sg) / (float)C, mgx = warp_sum(sgx) / (float)C;
    float4* dxr = reinterpret_cast<float4*>(dx + (long long)row * lddx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        float4 o;
        o.x = res[i].x + rs * (gg[i].x - mg - xh[i].x * mgx); o.y = res[i].y + rs * (gg[i].y - mg - xh[i].y * mgx);
        o.z = res[i].z + rs * (gg[i].z - mg - xh[i].z * mgx); o.w = res[i].w + rs * (gg[i].w - mg - xh[i].w * mgx);
        if (CS) { ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w; }
        dxr[c] = o;
        if (dx16) reinterpret_cast<uint2*>(dx16 + (long long)row * lddx)[c] = pack_half4(o.x * out_scale, o.y * out_scale, o.z * out_scale, o.w * out_scale);
      }
    }
  }
  if (dots) {
    if (dot_x) { acc_dx = warp_sum(acc_dx); if (lane == 0) atomicAdd(dots, acc_dx); }
    if (dot_t) { acc_dt = warp_sum(acc_dt); if (lane == 0) atomicAdd(dots + 1, acc_dt); }
  }
  if (!dgamma && !(CS && (cs_r1 || cs_out))) return;
  // cross-warp reduction of the per-lane column partials, one 128-column slab (i) at a time
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 >= nv) break;
    __syncthreads();
    float* rg = &red[0][warp][lane * 4];
    float* rb = &red[1][warp][lane * 4];
    rg[0] = ag[i].x; rg[1] = ag[i].y; rg[2] = ag[i].z; rg[3] = ag[i].w;
    rb[0] = ab[i].x; rb[1] = ab[i].y; rb[2] = ab[i].z; rb[3] = ab[i].w;
    if (CS) {
      float* q1 = &red[2][warp][lane * 4];
      float* qo = &red[3][warp][lane * 4];
      q1[0] = a1[i].x; q1[1] = a1[i].y; q1[2] = a1[i].z; q1[3] = a1[i].w;
      qo[0] = ao[i].x; qo[1] = ao[i].y; qo[2] = ao[i].z; qo[3] = ao[i].w;
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int col = i * 128 + threadIdx.x;
      if (col < C) {
        float sgm = 0.f, sbt = 0.f, s1 = 0.f, so = 0.f;
        for (int w = 0; w < nwarps; ++w) {
          sgm += red[0][w][threadIdx.x]; sbt += red[1][w][threadIdx.x];
          if (CS) { s1 += red[2][w][threadIdx.x]; so += red[3][w][threadIdx.x]; }
        }
        if (dgamma) { atomicAdd(dgamma + col, sgm); atomicAdd(dbeta + col, sbt); }
        if (CS) {
          if (cs_r1) atomicAdd(cs_r1 + col, s1);
          if (cs_out) atomicAdd(cs_out + col, so);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm bwd, streamed through shared memory
// The register-resident kernel above keeps one row (4-5 tensors x 1.5 KB) in registers per warp; at the 80 registers that give it 24 warps per
// SM the compiler cannot hoist all of a row's loads, so a row costs ~3 dependent round trips and the kernel runs at 2.9-4.1 TB/s (0.45-0.63 of
// the HBM peak, ncu: nothing saturated, 36 % occupancy).  Here the loads cost no registers: one thread issues 1-D bulk-asynchronous copies
// (cp.async.bulk, completion on an mbarrier) of EIGHT consecutive rows of every input tensor (6-12 KB each) into a two-stage shared-memory
// ring, so a whole stage (30-54 KB per block, two blocks per SM) is in flight while the eight warps work on the previous one from shared memory.
// Same arithmetic, same outputs, same fused column sums / gate-gradient dots as layernorm_bwd_kernel<NV, CS, true>; requires contiguous rows.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int NV, bool CS>
__global__ void __launch_bounds__(256, (NV <= 3 ? 2 : 1)) layernorm_bwd_stream_kernel(const __half* __restrict__ dy16, const float* __restrict__ x, const float* __restrict__ mean,
                                                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                      const float* __restrict__ r1, const float* __restrict__ r2, const float* __restrict__ s2_dev,
                                                                      float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                      float* __restrict__ cs_r1, float* __restrict__ cs_out, int M, int C, int rows_per_block,
                                                                      float dy_scale, __half* __restrict__ dx16, float out_scale,
                                                                      const float* __restrict__ scales_dev, const float* __restrict__ dot_t,
                                                                      float* __restrict__ dots, int dot_x) {
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ float red[4][8][32 * 4 + 4];
  __shared__ __align__(8) uint64_t bars[2];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 2;
  // stage layout: dy16 [8, C] fp16 | x [8, C] | r1 | r2 | t (the optional ones only when present)
  const uint32_t b16 = 8u * C * 2u, b32 = 8u * C * 4u;
  const uint32_t off_x = b16, off_r1 = off_x + b32, off_r2 = off_r1 + (r1 ? b32 : 0u), off_t = off_r2 + (r2 ? b32 : 0u);
  const uint32_t stage_bytes = off_t + (dot_t ? b32 : 0u);
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_barrier_init(); }
  __syncthreads();
  pdl_wait();
  if (scales_dev) { out_scale *= __ldg(scales_dev); dy_scale *= __ldg(scales_dev + 1); }
  const float s2 = (r2 && s2_dev) ? __ldg(s2_dev) : 1.0f;
  const int row0 = blockIdx.x * rows_per_block;
  const int row1 = min(M, row0 + rows_per_block);
  const int nst = (row1 - row0 + 7) / 8;
  auto issue = [&](int st) {          // one thread: expect the stage's bytes, then one bulk copy per tensor
    const int ra = row0 + st * 8, n = min(8, row1 - ra);
    const uint32_t base = ring_u32 + (uint32_t)(st & 1) * stage_bytes, bar = bar0 + 8u * (st & 1);
    const uint32_t n16 = (uint32_t)n * C * 2u, n32 = (uint32_t)n * C * 4u;
    mbar_expect_tx(bar, n16 + n32 + (r1 ? n32 : 0u) + (r2 ? n32 : 0u) + (dot_t ? n32 : 0u));
    bulk_g2s(base, dy16 + (long long)ra * C, n16, bar);
    bulk_g2s(base + off_x, x + (long long)ra * C, n32, bar);
    if (r1) bulk_g2s(base + off_r1, r1 + (long long)ra * C, n32, bar);
    if (r2) bulk_g2s(base + off_r2, r2 + (long long)ra * C, n32, bar);
    if (dot_t) bulk_g2s(base + off_t, dot_t + (long long)ra * C, n32, bar);
  };
  if (threadIdx.x == 0 && nst > 0) issue(0);
  float4 ag[NV], ab[NV], a1[CS ? NV : 1], ao[CS ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  if (CS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { a1[i] = make_float4(0, 0, 0, 0); ao[i] = make_float4(0, 0, 0, 0); }
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float acc_dx = 0.f, acc_dt = 0.f;
  for (int st = 0; st < nst; ++st) {
    __syncthreads();                                   // every warp is done with stage st - 1: its buffer may be refilled
    if (threadIdx.x == 0 && st + 1 < nst) issue(st + 1);
    mbar_wait(bar0 + 8u * (st & 1), (uint32_t)(st >> 1) & 1u);
    const int row = row0 + st * 8 + warp;
    if (row >= row1) continue;
    const unsigned char* base = ring + (size_t)(st & 1) * stage_bytes;
    const uint2* dyr = reinterpret_cast<const uint2*>(base) + (size_t)warp * nv;
    const float4* xr = reinterpret_cast<const float4*>(base + off_x) + (size_t)warp * nv;
    const float4* r1r = reinterpret_cast<const float4*>(base + off_r1) + (size_t)warp * nv;
    const float4* r2r = reinterpret_cast<const float4*>(base + off_r2) + (size_t)warp * nv;
    const float4* dtr = reinterpret_cast<const float4*>(base + off_t) + (size_t)warp * nv;
    const float mu = mean[row], rs = rstd[row];
    // two passes over the staged row (shared memory is cheap to re-read; keeping xhat / g / residual in registers between the passes cost
    // 12 * NV registers and spilled for C = 768): pass 1 reduces, pass 2 recomputes and writes
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const uint2 u = dyr[c];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        const float4 d = make_float4(lo.x * dy_scale, lo.y * dy_scale, hi.x * dy_scale, hi.y * dy_scale);
        const float4 xv = xr[c], gm = __ldg(g4 + c);
        if (CS || dot_x || dot_t) {
          float4 res = make_float4(0, 0, 0, 0);
          if (r1) res = r1r[c];
          if (r2) {
            const float4 a = r2r[c];
            res.x += s2 * a.x; res.y += s2 * a.y; res.z += s2 * a.z; res.w += s2 * a.w;
            if (dot_x) acc_dx += (a.x * xv.x + a.y * xv.y) + (a.z * xv.z + a.w * xv.w);
            if (dot_t) { const float4 tv = dtr[c]; acc_dt += (a.x * tv.x + a.y * tv.y) + (a.z * tv.z + a.w * tv.w); }
          }
          if (CS) { a1[i].x += res.x; a1[i].y += res.y; a1[i].z += res.z; a1[i].w += res.w; }
        }
        const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        const float4 gg = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        sg += (gg.x + gg.y) + (gg.z + gg.w);
        sgx += (gg.x * xh.x + gg.y * xh.y) + (gg.z * xh.z + gg.w * xh.w);
        ag[i].x += d.x * xh.x; ag[i].y += d.y * xh.y; ag[i].z += d.z * xh.z; ag[i].w += d.w * xh.w;
        ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      }
    }
    const float mg = warp_sum(sg) / (float)C, mgx = warp_sum(sgx) / (float)C;
    float4* dxr = reinterpret_cast<float4*>(dx + (long long)row * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv) {
        const uint2 u = dyr[c];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        const float4 xv = xr[c], gm = __ldg(g4 + c);
        float4 res = make_float4(0, 0, 0, 0);
        if (r1) res = r1r[c];
        if (r2) { const float4 a = r2r[c]; res.x += s2 * a.x; res.y += s2 * a.y; res.z += s2 * a.z; res.w += s2 * a.w; }
        float4 o;
        o.x = res.x + rs * ((lo.x * dy_scale) * gm.x - mg - ((xv.x - mu) * rs) * mgx); o.y = res.y + rs * ((lo.y * dy_scale) * gm.y - mg - ((xv.y - mu) * rs) * mgx);
        o.z = res.z + rs * ((hi.x * dy_scale) * gm.z - mg - ((xv.z - mu) * rs) * mgx); o.w = res.w + rs * ((hi.y * dy_scale) * gm.w - mg - ((xv.w - mu) * rs) * mgx);
        if (CS) { ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w; }
        dxr[c] = o;
        if (dx16) reinterpret_cast<uint2*>(dx16 + (long long)row * C)[c] = pack_half4(o.x * out_scale, o.y * out_scale, o.z * out_scale, o.w * out_scale);
      }
    }
  }
  if (dots) {
    if (dot_x) { acc_dx = warp_sum(acc_dx); if (lane == 0) atomicAdd(dots, acc_dx); }
    if (dot_t) { acc_dt = warp_sum(acc_dt); if (lane == 0) atomicAdd(dots + 1, acc_dt); }
  }
  if (!dgamma && !(CS && (cs_r1 || cs_out))) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 >= nv) break;
    __syncthreads();
    float* rg = &red[0][warp][lane * 4];
    float* rb = &red[1][warp][lane * 4];
    rg[0] = ag[i].x; rg[1] = ag[i].y; rg[2] = ag[i].z; rg[3] = ag[i].w;
    rb[0] = ab[i].x; rb[1] = ab[i].y; rb[2] = ab[i].z; rb[3] = ab[i].w;
    if (CS) {
      float* q1 = &red[2][warp][lane * 4];
      float* qo = &red[3][warp][lane * 4];
      q1[0] = a1[i].x; q1[1] = a1[i].y; q1[2] = a1[i].z; q1[3] = a1[i].w;
      qo[0] = ao[i].x; qo[1] = ao[i].y; qo[2] = ao[i].z; qo[3] = ao[i].w;
    }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int col = i * 128 + threadIdx.x;
      if (col < C) {
        float sgm = 0.f, sbt = 0.f, s1 = 0.f, so = 0.f;
        for (int w = 0; w < 8; ++w) {
          sgm += red[0][w][threadIdx.x]; sbt += red[1][w][threadIdx.x];
          if (CS) { s1 += red[2][w][threadIdx.x]; so += red[3][w][threadIdx.x]; }
        }
        if (dgamma) { atomicAdd(dgamma + col, sgm); atomicAdd(dbeta + col, sbt); }
        if (CS) {
          if (cs_r1) atomicAdd(cs_r1 + col, s1);
          if (cs_out) atomicAdd(cs_out + col, so);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ softmax over attention rows
// in place on S[rows][ld], n valid columns (<= 256); one warp per row
__global__ void __launch_bounds__(256) softmax_fwd_kernel(float* __restrict__ S, long long ld, long long rows, int n, int rnd) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* s = S + row * ld;
  float v[8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int c = lane + i * 32; v[i] = (c < n) ? s[c] : -INFINITY; mx = fmaxf(mx, v[i]); }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int c = lane + i * 32; v[i] = (c < n) ? expf(v[i] - mx) : 0.f; sum += v[i]; }
  const float inv = 1.0f / warp_sum(sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int c = lane + i * 32; if (c < n) s[c] = rnd ? round_tf32(v[i] * inv) : v[i] * inv; }
}

// dS = scale * P .* (dP - sum_j dP_j P_j), written over dP
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, long long ld, long long rows, int n, float scale, int rnd) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = P + row * ld;
  float* d = dP + row * ld;
  float pv[8], dv[8];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    pv[i] = (c < n) ? p[c] : 0.f; dv[i] = (c < n) ? d[c] : 0.f;
    dot += pv[i] * dv[i];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    if (c < n) { const float o = scale * pv[i] * (dv[i] - dot); d[c] = rnd ? round_tf32(o) : o; }
  }
}

// ------------------------------------------------------------------------------------------ column sums (bias grads)
// out[col] += scale * sum_rows X[row, col].  A block of 256 threads covers 128 columns (32 float4 lanes) x 8 row phases; each thread
// keeps 4 independent 128-bit loads in flight, the 8 phases are combined through shared memory and one atomicAdd per column and block
// lands in `out`.  N % 4 == 0 takes this path; ragged N falls back to the scalar kernel below.
__global__ void __launch_bounds__(256) colsum4_kernel(const float* __restrict__ X, long long ld, int M, int N, const float* __restrict__ scale_dev,
                                                      float* __restrict__ out, int rows_per_block) {
  __shared__ float4 red[8][32];
  const int cl = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cl) * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float4 a = make_float4(0, 0, 0, 0), b = a, c = a, d = a;
  if (col < N) {
    const float* base = X + col;
    int r = r0 + ph;
    for (; r + 24 < r1; r += 32) {
      const float4 v0 = *reinterpret_cast<const float4*>(base + (long long)r * ld), v1 = *reinterpret_cast<const float4*>(base + (long long)(r + 8) * ld);
      const float4 v2 = *reinterpret_cast<const float4*>(base + (long long)(r + 16) * ld), v3 = *reinterpret_cast<const float4*>(base + (long long)(r + 24) * ld);
      a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w; b.x += v1.x; b.y += v1.y; b.z += v1.z; b.w += v1.w;
      c.x += v2.x; c.y += v2.y; c.z += v2.z; c.w += v2.w; d.x += v3.x; d.y += v3.y; d.z += v3.z; d.w += v3.w;
    }
    for (; r < r1; r += 8) { const float4 v0 = *reinterpret_cast<const float4*>(base + (long long)r * ld); a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w; }
  }
  red[ph][cl] = make_float4((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w));
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c4 = threadIdx.x >> 2, e = threadIdx.x & 3;
    const int oc = (blockIdx.x * 32 + c4) * 4 + e;
    if (oc < N) {
      float sum = 0.f;
#pragma unroll
      for (int p = 0; p < 8; ++p) sum += reinterpret_cast<const float*>(&red[p][c4])[e];
      atomicAdd(out + oc, (scale_dev ? __ldg(scale_dev) : 1.0f) * sum);
    }
  }
}
// scalar fallback: thread per column, grid.y row chunks, atomics to combine
__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ X, long long ld, int M, int N, const float* __restrict__ scale_dev,
                                                     float* __restrict__ out, int rows_per_block) {
  const int col = blockIdx.x * 128 + threadIdx.x;
  if (col >= N) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    a0 += X[(long long)r * ld + col]; a1 += X[(long long)(r + 1) * ld + col];
    a2 += X[(long long)(r + 2) * ld + col]; a3 += X[(long long)(r + 3) * ld + col];
  }
  for (; r < r1; ++r) a0 += X[(long long)r * ld + col];
  const float sc = scale_dev ? __ldg(scale_dev) : 1.0f;
  atomicAdd(out + col, sc * ((a0 + a1) + (a2 + a3)));
}

// ------------------------------------------------------------------------------------------ block gate blend
// out = d[1] * t + d[0] * x           (models/model_distilled.py:493)
__global__ void __launch_bounds__(256) blend_fwd_kernel(const float4* __restrict__ t, const float4* __restrict__ x, const float* __restrict__ d,
                                                        float4* __restrict__ out, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  const float d0 = __ldg(d), d1 = __ldg(d + 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = t[i], b = x[i];
    out[i] = make_float4(d1 * a.x + d0 * b.x, d1 * a.y + d0 * b.y, d1 * a.z + d0 * b.z, d1 * a.w + d0 * b.w);
  }
}
// dots[0] += <g, x>, dots[1] += <g, t>   (gradients of the loss wrt the blend weights d0, d1)
__global__ void __launch_bounds__(256) blend_dots_kernel(const float4* __restrict__ g, const float4* __restrict__ t, const float4* __restrict__ x,
                                                         float* __restrict__ dots, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  float sx = 0.f, st = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 gg = g[i], a = t[i], b = x[i];
    sx += (gg.x * b.x + gg.y * b.y) + (gg.z * b.z + gg.w * b.w);
    st += (gg.x * a.x + gg.y * a.y) + (gg.z * a.z + gg.w * a.w);
  }
  __shared__ float rs[2][8];
  sx = warp_sum(sx); st = warp_sum(st);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { rs[0][warp] = sx; rs[1][warp] = st; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += rs[0][w]; b += rs[1][w]; }
    atomicAdd(dots, a); atomicAdd(dots + 1, b);
  }
}

// ------------------------------------------------------------------------------------------ patch embed glue
// im2col of the 16x16 stride-16 conv: out[(b*196 + py*14 + px), c*256 + ky*16 + kx] = x[b, c, py*16+ky, px*16+kx]
__global__ void __launch_bounds__(256) im2col16_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int Cin, int HW, int P, int rnd) {
  const int G = HW / P;                       // patches per side
  const long long total4 = (long long)B * G * G * Cin * P * (P / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic (the host checks total4 < 2^31): the 64-bit divisions made this copy issue-bound (70 % issue-active at 2.8 TB/s)
    unsigned t = (unsigned)i;
    const unsigned P4 = (unsigned)P / 4;
    const int kx4 = (int)(t % P4); t /= P4;
    const int ky = (int)(t % (unsigned)P); t /= (unsigned)P;
    const int c = (int)(t % (unsigned)Cin); t /= (unsigned)Cin;
    const int px = (int)(t % (unsigned)G); t /= (unsigned)G;
    const int py = (int)(t % (unsigned)G); t /= (unsigned)G;
    const int b = (int)t;
    float4 v = *reinterpret_cast<const float4*>(x + (((long long)b * Cin + c) * HW + (py * P + ky)) * HW + px * P + kx4 * 4);
    if (rnd) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}
// scatter-add inverse (not needed: the input image needs no gradient)

// tok[b,0,:] = cls + pos[0];  tok[b,1+p,:] = pe[b*np+p,:] * pscale[p] * tmask[b,p] + pos[1+p]
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const float* __restrict__ pe, const float* __restrict__ cls, const float* __restrict__ pos,
                                                              const float* __restrict__ pscale, const float* __restrict__ tmask,
                                                              float* __restrict__ tok, int B, int np, int C) {
  const int nv = C >> 2;
  const long long total = (long long)B * (np + 1) * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;                 // 32-bit index arithmetic (the host checks total < 2^31)
    const int c = (int)(iu % (unsigned)nv);
    const unsigned r = iu / (unsigned)nv;
    const int n = (int)(r % (unsigned)(np + 1));
    const int b = (int)(r / (unsigned)(np + 1));
    const float4 po = __ldg(reinterpret_cast<const float4*>(pos) + (long long)n * nv + c);
    float4 o;
    if (n == 0) {
      const float4 cl = __ldg(reinterpret_cast<const float4*>(cls) + c);
      o = make_float4(cl.x + po.x, cl.y + po.y, cl.z + po.z, cl.w + po.w);
    } else {
      const float4 v = reinterpret_cast<const float4*>(pe)[((long long)b * np + (n - 1)) * nv + c];
      float s = 1.0f;
      if (pscale) s *= __ldg(pscale + (n - 1));
      if (tmask) s *= __ldg(tmask + (long long)b * np + (n - 1));
      o = make_float4(v.x * s + po.x, v.y * s + po.y, v.z * s + po.z, v.w * s + po.w);
    }
    reinterpret_cast<float4*>(tok)[i] = o;
  }
}
// backward of assemble: d_pe = g[b,1+p,:] * s ; dscale[p] += sum_{b,c} g * pe ; (dpos, dcls via colsum-style kernel below)
__global__ void __launch_bounds__(128) assemble_tokens_bwd_kernel(const float* __restrict__ g, const float* __restrict__ pe, const float* __restrict__ pscale,
                                                                 const float* __restrict__ tmask, float* __restrict__ dpe, float* __restrict__ dscale,
                                                                 float* __restrict__ dtmask, int B, int np, int C) {
  // one warp per (b, p) row
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * np) return;
  const int p = (int)(row % np);
  const int b = (int)(row / np);
  const int nv = C >> 2;
  const float4* gr = reinterpret_cast<const float4*>(g) + ((long long)b * (np + 1) + 1 + p) * nv;
  const float4* pr = reinterpret_cast<const float4*>(pe) + row * nv;
  float4* dr = reinterpret_cast<float4*>(dpe) + row * nv;
  float s = 1.0f, sm = 1.0f, sp = 1.0f;
  if (pscale) { sp = __ldg(pscale + p); s *= sp; }
  if (tmask) { sm = __ldg(tmask + row); s *= sm; }
  float dot = 0.f;
  const bool need_dot = (dscale != nullptr) || (dtmask != nullptr);
  for (int c = lane; c < nv; c += 32) {
    const float4 gg = gr[c];
    if (need_dot) { const float4 v = pr[c]; dot += (gg.x * v.x + gg.y * v.y) + (gg.z * v.z + gg.w * v.w); }
    dr[c] = make_float4(gg.x * s, gg.y * s, gg.z * s, gg.w * s);
  }
  if (need_dot) {
    dot = warp_sum(dot);
    if (lane == 0) {
      if (dscale) atomicAdd(dscale + p, dot * sm);
      if (dtmask) dtmask[row] = dot * sp;
    }
  }
}
// dpos[n,:] += sum_b g[b,n,:] ; dcls[:] += sum_b g[b,0,:]
// grid (ntok, batch slices): a thread owns 4 adjacent columns and sums its slice of the batch with four independent accumulators (the first version
// walked all B images serially in one thread per column: 56 us for 39 MB; this one is bound by the read, ~10 us)
__global__ void __launch_bounds__(256) pos_cls_grad_kernel(const float* __restrict__ g, float* __restrict__ dpos, float* __restrict__ dcls, int B, int ntok, int C) {
  const int n = blockIdx.x, nv = C >> 2;
  const int per = (B + gridDim.y - 1) / gridDim.y, b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  for (int c = threadIdx.x; c < nv; c += blockDim.x) {
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    const float4* base = reinterpret_cast<const float4*>(g) + (long long)n * nv + c;
    const long long stride = (long long)ntok * nv;
    int b = b0;
    for (; b + 3 < b1; b += 4) {
      const float4 v0 = base[b * stride], v1 = base[(b + 1) * stride], v2 = base[(b + 2) * stride], v3 = base[(b + 3) * stride];
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w; a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w; a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; b < b1; ++b) { const float4 v = base[b * stride]; a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w; }
    const float sx = (a0.x + a1.x) + (a2.x + a3.x), sy = (a0.y + a1.y) + (a2.y + a3.y), sz = (a0.z + a1.z) + (a2.z + a3.z), sw = (a0.w + a1.w) + (a2.w + a3.w);
    float* o = dpos + (long long)n * C + 4 * c;
    atomicAdd(o, sx); atomicAdd(o + 1, sy); atomicAdd(o + 2, sz); atomicAdd(o + 3, sw);
    if (n == 0 && dcls) { atomicAdd(dcls + 4 * c, sx); atomicAdd(dcls + 4 * c + 1, sy); atomicAdd(dcls + 4 * c + 2, sz); atomicAdd(dcls + 4 * c + 3, sw); }
  }
}

// ------------------------------------------------------------------------------------------ generic small helpers
__global__ void scale_add_kernel(float4* __restrict__ y, const float4* __restrict__ x, const float* __restrict__ s_dev, float s, long long n4) {
  const float sc = s * (s_dev ? __ldg(s_dev) : 1.0f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = x[i]; float4 b = y[i];
    b.x += sc * a.x; b.y += sc * a.y; b.z += sc * a.z; b.w += sc * a.w;
    y[i] = b;
  }
}

__global__ void scale_to_f16_kernel(uint2* __restrict__ dst, const float4* __restrict__ src, float s, const float* __restrict__ s_dev, long long n4) {
  if (s_dev) s *= __ldg(s_dev);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = src[i];
    dst[i] = pack_half4(a.x * s, a.y * s, a.z * s, a.w * s);
  }
}

// Loss scale of the fp16 gradient operands, chosen ON THE DEVICE from the gradient that enters the backward: the largest power of two S with
// S * max|dlogits| <= target.  Every gradient is linear in dlogits, so this makes the fp16 range use independent of how the caller scaled the loss
// (mean over 128 images, sum over a batch, an arbitrary autograd head, ...).  scales = {S, 1/S}; a fixed S (> 0) bypasses the reduction.
__global__ void __launch_bounds__(1024) grad_scale_kernel(const float* __restrict__ dlogits, long long n, float target, float fixed, float* __restrict__ scales) {
  __shared__ float red[32];
  float S = fixed;
  if (!(fixed > 0.f)) {
    float mx = 0.f;
    // NaN / inf entries: fmaxf drops NaN, inf gives S = 0 -> clamped below.  One block (the result is one scalar): 128-bit loads, four in flight
    long long i0 = 0;
    if ((reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) {
      const float4* d4 = reinterpret_cast<const float4*>(dlogits);
      const long long n4 = n >> 2;
      long long i = threadIdx.x;
      for (; i + 3 * blockDim.x < n4; i += 4 * blockDim.x) {
        const float4 a = d4[i], b = d4[i + blockDim.x], c = d4[i + 2 * blockDim.x], d = d4[i + 3 * blockDim.x];
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))), fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))), fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)))));
      }
      for (; i < n4; i += blockDim.x) { const float4 a = d4[i]; mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)))); }
      i0 = n4 << 2;
    }
    for (long long i = i0 + threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(dlogits[i]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
      mx = warp_max(threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f);
      int e = 0;
      if (mx > 0.f && mx < INFINITY) { frexpf(target / mx, &e); e -= 1; }       // 2^e <= target / mx < 2^(e+1)
      e = max(-24, min(24, e));
      S = ldexpf(1.0f, e);
    }
  }
  if (threadIdx.x == 0) { scales[0] = S; scales[1] = 1.0f / S; }
}

// dst = rna_tf32(src) over up to kMaxSeg tensors in one launch (all GEMM weights of the model)
struct RoundSegs { const float* src[kMaxRoundSegs]; float* dst[kMaxRoundSegs]; long long n4[kMaxRoundSegs]; int nseg; };
__global__ void __launch_bounds__(256) round_segs_kernel(const __grid_constant__ RoundSegs segs) {
  const int seg = blockIdx.y;
  const float4* s = reinterpret_cast<const float4*>(segs.src[seg]);
  float4* d = reinterpret_cast<float4*>(segs.dst[seg]);
  const long long n4 = segs.n4[seg];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = s[i];
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    d[i] = v;
  }
}

// fp32 weights -> fp16 operand copies, all GEMM weights of the model in one launch: dst [rows, cols] (K-major B operand of the forward
// GEMM) and, when dstT != NULL, the transpose [cols, rows] (K-major B operand of the data-gradient GEMM, so no MN-major 16-bit path is
// needed there).  32 x 32 tiles through shared memory; both stores are coalesced.
struct CvtSegs { const float* src[kMaxRoundSegs]; __half* dst[kMaxRoundSegs]; __half* dstT[kMaxRoundSegs]; int rows[kMaxRoundSegs]; int cols[kMaxRoundSegs]; int nseg; int vec; };
__global__ void __launch_bounds__(256) cvt_f16_segs_kernel(const __grid_constant__ CvtSegs segs) {
  __shared__ float tile[32][33];
  const int seg = blockIdx.y;
  const float* __restrict__ src = segs.src[seg];
  __half* __restrict__ dst = segs.dst[seg];
  __half* __restrict__ dstT = segs.dstT[seg];
  const int rows = segs.rows[seg], cols = segs.cols[seg];
  const int tc = (cols + 31) / 32, tr = (rows + 31) / 32;
  if (segs.vec) {
    // every matrix has rows, cols % 4 == 0 and 16-byte aligned rows: one 128-bit load and one 64-bit store per thread for the straight copy, four
    // values down a column per 64-bit store for the transposed one (the scalar version below ran at 2 TB/s: 66 us for the 22 M DeiT-Small weights)
    const int r_in = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;           // load: 32 rows x 8 float4
    const int c_out = threadIdx.x >> 3, r4 = (threadIdx.x & 7) * 4;          // transposed store: 32 columns x 8 groups of 4 rows
    for (int t = blockIdx.x; t < tc * tr; t += gridDim.x) {
      const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
      const int r = r0 + r_in, c = c0 + c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows && c < cols) {
        v = *reinterpret_cast<const float4*>(src + (long long)r * cols + c);
        if (dst) *reinterpret_cast<uint2*>(dst + (long long)r * cols + c) = pack_half4(v.x, v.y, v.z, v.w);
      }
      if (dstT) {
        tile[r_in][c4] = v.x; tile[r_in][c4 + 1] = v.y; tile[r_in][c4 + 2] = v.z; tile[r_in][c4 + 3] = v.w;
        __syncthreads();
        const int cc = c0 + c_out, rr = r0 + r4;
        if (cc < cols && rr < rows)
          *reinterpret_cast<uint2*>(dstT + (long long)cc * rows + rr) = pack_half4(tile[r4][c_out], tile[r4 + 1][c_out], tile[r4 + 2][c_out], tile[r4 + 3][c_out]);
        __syncthreads();
      }
    }
    return;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < tc * tr; t += gridDim.x) {
    const int r0 = (t / tc) * 32, c0 = (t % tc) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + i * 8, c = c0 + tx;
      float v = 0.f;
      if (r < rows && c < cols) {
        v = src[(long long)r * cols + c];
        if (dst) dst[(long long)r * cols + c] = __float2half_rn(v);
      }
      tile[ty + i * 8][tx] = v;
    }
    if (dstT) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, r = r0 + tx;
        if (r < rows && c < cols) dstT[(long long)c * rows + r] = __float2half_rn(tile[tx][ty + i * 8]);
      }
      __syncthreads();
    }
  }
}

static inline int grid_for(long long n, int threads, int max_blocks = 148 * 8) {
  long long b = (n + threads - 1) / threads;
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  return (int)b;
}

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, float* y, long long ldy, float* mean,
                  float* rstd, int M, int C, cudaStream_t st, int rnd, void* y16) {
  UVC_REQUIRE(C > 0 && (C & 3) == 0 && C <= kMaxVec * 128, UVC_ERR_BAD_SHAPE, "layernorm: C=%d must be a multiple of 4 and <= %d", C, kMaxVec * 128);
  UVC_REQUIRE((ldx & 3) == 0 && (ldy & 3) == 0, UVC_ERR_BAD_SHAPE, "layernorm: row strides must be multiples of 4");
  if (M <= 0) return UVC_OK;
  if (y16) UVC_LN_DISPATCH(C, launch_pdl(layernorm_fwd_kernel<NV, true>, dim3((M + 7) / 8), dim3(256), 0, st, x, ldx, gamma, beta, eps, static_cast<float*>(y16), ldy, mean, rstd, M, C, 0));
  else UVC_LN_DISPATCH(C, launch_pdl(layernorm_fwd_kernel<NV, false>, dim3((M + 7) / 8), dim3(256), 0, st, x, ldx, gamma, beta, eps, y, ldy, mean, rstd, M, C, rnd));
  return check_launch("layernorm_fwd");
}

int layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* mean, const float* rstd, const float* gamma,
                  const float* r1, const float* r2, const float* s2_dev, float* dx, long long lddx, float* dgamma, float* dbeta, int M, int C,
                  cudaStream_t st, float* cs_r1, float* cs_out, const void* dy16, float dy_scale, void* dx16, float out_scale, const float* scales_dev,
                  const float* dot_t, float* dots, int dot_x) {
  UVC_REQUIRE(!(dot_t || dot_x) || (dots && r2), UVC_ERR_BAD_ARG, "layernorm_bwd: the gate-gradient dot products need r2 and an output pointer");
  UVC_REQUIRE(C > 0 && (C & 3) == 0 && C <= kMaxVec * 128, UVC_ERR_BAD_SHAPE, "layernorm_bwd: C=%d unsupported", C);
  UVC_REQUIRE((ldx & 3) == 0 && (lddy & 3) == 0 && (lddx & 3) == 0, UVC_ERR_BAD_SHAPE, "layernorm_bwd: row strides must be multiples of 4");
  UVC_REQUIRE(!cs_r1 || r1 || r2, UVC_ERR_BAD_ARG, "layernorm_bwd: cs_r1 without a residual input");
  if (M <= 0) return UVC_OK;
  // one wave of 3 resident blocks per SM (the kernel is compiled for 80 registers): each warp has one row (4 x 1.5 KB) in flight, so the
  // bytes in flight per SM, not the column reductions, set the rate (measured: no gain from dropping the atomics, +8..15 % from 16 -> 24 warps)
  int blocks = 148 * (C <= 384 ? UVC_LN_BWD_MINB : 2);           // resident blocks per SM the kernel is compiled for (wider rows need more registers)
  int rpb = (M + blocks - 1) / blocks;
  if (rpb < 8) rpb = 8;
  blocks = (M + rpb - 1) / rpb;
  __half* h16 = static_cast<__half*>(dx16);
  // fp16 dy with contiguous rows (the engine's block loop): the streamed kernel (bulk-asynchronous staging, see layernorm_bwd_stream_kernel)
  const bool stream_ok = !getenv("UVC_LN_BWD_REG");          // bring-up / tests: force the register-resident kernel
  if (dy16 && stream_ok && lddy == C && ldx == C && lddx == C && (C & 7) == 0 && M >= 1024 && C >= 128 &&
      ((reinterpret_cast<uintptr_t>(dy16) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r1) | reinterpret_cast<uintptr_t>(r2) |
        reinterpret_cast<uintptr_t>(dot_t)) & 15) == 0) {
    const int ntens32 = 1 + (r1 ? 1 : 0) + (r2 ? 1 : 0) + (dot_t ? 1 : 0);
    const size_t smem = 2 * (size_t)(8 * C * 2 + ntens32 * 8 * C * 4);
    if (smem + 18 * 1024 <= 220 * 1024) {                       // two blocks per SM up to C = 384, one block for wider rows
      int nb = 148 * 2;
      int rpb2 = ((M + nb - 1) / nb + 7) / 8 * 8;
      nb = (M + rpb2 - 1) / rpb2;
      const __half* d16h = static_cast<const __half*>(dy16);
      __half* h16s = static_cast<__half*>(dx16);
      cudaError_t e = cudaSuccess;
      if (cs_r1 || cs_out) {
        UVC_LN_DISPATCH(C, { e = cudaFuncSetAttribute(layernorm_bwd_stream_kernel<NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                             launch_pdl(layernorm_bwd_stream_kernel<NV, true>, dim3(nb), dim3(256), smem, st, d16h, x, mean, rstd, gamma, r1, r2, s2_dev, dx, dgamma, dbeta,
                                        cs_r1, cs_out, M, C, rpb2, dy_scale, h16s, out_scale, scales_dev, dot_t, dots, dot_x); });
      } else {
        UVC_LN_DISPATCH(C, { e = cudaFuncSetAttribute(layernorm_bwd_stream_kernel<NV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                             launch_pdl(layernorm_bwd_stream_kernel<NV, false>, dim3(nb), dim3(256), smem, st, d16h, x, mean, rstd, gamma, r1, r2, s2_dev, dx, dgamma, dbeta,
                                        nullptr, nullptr, M, C, rpb2, dy_scale, h16s, out_scale, scales_dev, dot_t, dots, dot_x); });
      }
      UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(layernorm_bwd_stream smem=%zu): %s", smem, cudaGetErrorString(e));
      return check_launch("layernorm_bwd_stream");
    }
  }
  if (dy16) {
    const float* d16 = static_cast<const float*>(dy16);
    if (cs_r1 || cs_out)
      UVC_LN_DISPATCH(C, launch_pdl(layernorm_bwd_kernel<NV, true, true>, dim3(blocks), dim3(256), 0, st, d16, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, cs_r1, cs_out, M, C, rpb, dy_scale, h16, out_scale, scales_dev, dot_t, dots, dot_x));
    else
      UVC_LN_DISPATCH(C, launch_pdl(layernorm_bwd_kernel<NV, false, true>, dim3(blocks), dim3(256), 0, st, d16, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, nullptr, nullptr, M, C, rpb, dy_scale, h16, out_scale, scales_dev, dot_t, dots, dot_x));
  } else if (cs_r1 || cs_out)
    UVC_LN_DISPATCH(C, launch_pdl(layernorm_bwd_kernel<NV, true, false>, dim3(blocks), dim3(256), 0, st, dy, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, cs_r1, cs_out, M, C, rpb, 1.0f, h16, out_scale, scales_dev, dot_t, dots, dot_x));
  else
    UVC_LN_DISPATCH(C, launch_pdl(layernorm_bwd_kernel<NV, false, false>, dim3(blocks), dim3(256), 0, st, dy, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, nullptr, nullptr, M, C, rpb, 1.0f, h16, out_scale, scales_dev, dot_t, dots, dot_x));
  return check_launch("layernorm_bwd");
}

int softmax_fwd(float* S, long long ld, long long rows, int n, cudaStream_t st, int rnd) {
  UVC_REQUIRE(n > 0 && n <= 256, UVC_ERR_BAD_SHAPE, "softmax: n=%d must be in [1,256]", n);
  if (rows <= 0) return UVC_OK;
  softmax_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(S, ld, rows, n, rnd);
  return check_launch("softmax_fwd");
}
int softmax_bwd(const float* P, float* dP, long long ld, long long rows, int n, float scale, cudaStream_t st, int rnd) {
  UVC_REQUIRE(n > 0 && n <= 256, UVC_ERR_BAD_SHAPE, "softmax_bwd: n=%d must be in [1,256]", n);
  if (rows <= 0) return UVC_OK;
  softmax_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(P, dP, ld, rows, n, scale, rnd);
  return check_launch("softmax_bwd");
}
int colsum(const float* X, long long ld, int M, int N, const float* scale_dev, float* out, cudaStream_t st) {
  if (M <= 0 || N <= 0) return UVC_OK;
  if ((N & 3) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
    const int cb = (N + 127) / 128;
    int chunks = 148 * 4 / cb;
    if (chunks < 1) chunks = 1;
    int rpb = (M + chunks - 1) / chunks;
    if (rpb < 64) rpb = 64;
    chunks = (M + rpb - 1) / rpb;
    colsum4_kernel<<<dim3(cb, chunks), 256, 0, st>>>(X, ld, M, N, scale_dev, out, rpb);
    return check_launch("colsum");
  }
  int chunks = 148 * 4 / ((N + 127) / 128);
  if (chunks < 1) chunks = 1;
  int rpb = (M + chunks - 1) / chunks;
  if (rpb < 16) rpb = 16;
  chunks = (M + rpb - 1) / rpb;
  colsum_kernel<<<dim3((N + 127) / 128, chunks), 128, 0, st>>>(X, ld, M, N, scale_dev, out, rpb);
  return check_launch("colsum");
}
int blend_fwd(const float* t, const float* x, const float* d, float* out, long long n, cudaStream_t st) {
  UVC_REQUIRE((n & 3) == 0, UVC_ERR_BAD_SHAPE, "blend: element count must be a multiple of 4");
  launch_pdl(blend_fwd_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, st, reinterpret_cast<const float4*>(t), reinterpret_cast<const float4*>(x), d, reinterpret_cast<float4*>(out), n / 4);
  return check_launch("blend_fwd");
}
int blend_dots(const float* g, const float* t, const float* x, float* dots, long long n, cudaStream_t st) {
  UVC_REQUIRE((n & 3) == 0, UVC_ERR_BAD_SHAPE, "blend_dots: element count must be a multiple of 4");
  launch_pdl(blend_dots_kernel, dim3(grid_for(n / 4, 256, 148 * 4)), dim3(256), 0, st, reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(t), reinterpret_cast<const float4*>(x), dots, n / 4);
  return check_launch("blend_dots");
}
int im2col16(const float* x, float* out, int B, int Cin, int HW, int P, cudaStream_t st, int rnd) {
  UVC_REQUIRE(P % 4 == 0 && HW % P == 0, UVC_ERR_BAD_SHAPE, "im2col: patch %d must divide image %d and be a multiple of 4", P, HW);
  const long long total4 = (long long)B * Cin * HW * HW / 4;
  UVC_REQUIRE(total4 < (1ll << 31), UVC_ERR_BAD_SHAPE, "im2col: batch too large (%lld elements)", total4 * 4);
  im2col16_kernel<<<grid_for(total4, 256, 148 * 16), 256, 0, st>>>(x, out, B, Cin, HW, P, rnd);
  return check_launch("im2col16");
}
int assemble_tokens(const float* pe, const float* cls, const float* pos, const float* pscale, const float* tmask, float* tok, int B, int np, int C,
                    cudaStream_t st) {
  UVC_REQUIRE((C & 3) == 0, UVC_ERR_BAD_SHAPE, "assemble_tokens: C must be a multiple of 4");
  UVC_REQUIRE((long long)B * (np + 1) * (C / 4) < (1ll << 31), UVC_ERR_BAD_SHAPE, "assemble_tokens: too many elements");
  assemble_tokens_kernel<<<grid_for((long long)B * (np + 1) * (C / 4), 256, 148 * 16), 256, 0, st>>>(pe, cls, pos, pscale, tmask, tok, B, np, C);
  return check_launch("assemble_tokens");
}
int assemble_tokens_bwd(const float* g, const float* pe, const float* pscale, const float* tmask, float* dpe, float* dscale, float* dtmask,
                        float* dpos, float* dcls, int B, int np, int C, cudaStream_t st) {
  UVC_REQUIRE((C & 3) == 0, UVC_ERR_BAD_SHAPE, "assemble_tokens_bwd: C must be a multiple of 4");
  const long long rows = (long long)B * np;
  assemble_tokens_bwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(g, pe, pscale, tmask, dpe, dscale, dtmask, B, np, C);
  int rc = check_launch("assemble_tokens_bwd");
  if (rc) return rc;
  if (dpos) {
    pos_cls_grad_kernel<<<dim3(np + 1, B >= 32 ? 8 : 1), (C / 4 <= 128 ? 128 : 256), 0, st>>>(g, dpos, dcls, B, np + 1, C);
    rc = check_launch("pos_cls_grad");
  }
  return rc;
}
int round_tf32_segs(const float* const* src, float* const* dst, const long long* n, int nseg, cudaStream_t st) {
  for (int base = 0; base < nseg; base += kMaxRoundSegs) {
    RoundSegs segs;
    segs.nseg = (nseg - base < kMaxRoundSegs) ? nseg - base : kMaxRoundSegs;
    long long mx = 0;
    for (int i = 0; i < segs.nseg; ++i) {
      UVC_REQUIRE((n[base + i] & 3) == 0, UVC_ERR_BAD_SHAPE, "round_tf32: element count must be a multiple of 4");
      segs.src[i] = src[base + i]; segs.dst[i] = dst[base + i]; segs.n4[i] = n[base + i] / 4;
      if (segs.n4[i] > mx) mx = segs.n4[i];
    }
    if (mx == 0) continue;
    round_segs_kernel<<<dim3(grid_for(mx, 256, 32), segs.nseg), 256, 0, st>>>(segs);
    int rc = check_launch("round_tf32");
    if (rc) return rc;
  }
  return UVC_OK;
}
int scale_to_f16(void* dst16, const float* src, float s, long long n, cudaStream_t st, const float* s_dev) {
  UVC_REQUIRE((n & 3) == 0, UVC_ERR_BAD_SHAPE, "scale_to_f16: element count must be a multiple of 4");
  scale_to_f16_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(static_cast<uint2*>(dst16), reinterpret_cast<const float4*>(src), s, s_dev, n / 4);
  return check_launch("scale_to_f16");
}
int grad_scale(const float* dlogits, long long n, float target, float fixed, float* scales, cudaStream_t st) {
  grad_scale_kernel<<<1, 1024, 0, st>>>(dlogits, n, target, fixed, scales);
  return check_launch("grad_scale");
}
int cvt_f16_segs(const float* const* src, void* const* dst, void* const* dstT, const int* rows, const int* cols, int nseg, cudaStream_t st) {
  for (int base = 0; base < nseg; base += kMaxRoundSegs) {
    CvtSegs segs;
    segs.nseg = (nseg - base < kMaxRoundSegs) ? nseg - base : kMaxRoundSegs;
    long long mx = 0;
    segs.vec = 1;
    for (int i = 0; i < segs.nseg; ++i) {
      segs.src[i] = src[base + i]; segs.dst[i] = static_cast<__half*>(dst[base + i]); segs.dstT[i] = dstT ? static_cast<__half*>(dstT[base + i]) : nullptr;
      segs.rows[i] = rows[base + i]; segs.cols[i] = cols[base + i];
      if ((rows[base + i] & 3) || (cols[base + i] & 3) || (reinterpret_cast<uintptr_t>(segs.src[i]) & 15) || (reinterpret_cast<uintptr_t>(segs.dst[i]) & 7) ||
          (reinterpret_cast<uintptr_t>(segs.dstT[i]) & 7)) segs.vec = 0;
      const long long tiles = (long long)((rows[base + i] + 31) / 32) * ((cols[base + i] + 31) / 32);
      if (tiles > mx) mx = tiles;
    }
    if (mx == 0) continue;
    cvt_f16_segs_kernel<<<dim3((unsigned)(mx < 64 ? mx : 64), segs.nseg), 256, 0, st>>>(segs);
    int rc = check_launch("cvt_f16_segs");
    if (rc) return rc;
  }
  return UVC_OK;
}
int scale_add(float* y, const float* x, const float* s_dev, float s, long long n, cudaStream_t st) {
  UVC_REQUIRE((n & 3) == 0, UVC_ERR_BAD_SHAPE, "scale_add: element count must be a multiple of 4");
  scale_add_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(x), s_dev, s, n / 4);
  return check_launch("scale_add");
}

}  // namespace uvc

// ------------------------------------------------------------------------------------------ C ABI
#define UVC_ST static_cast<cudaStream_t>(stream)
extern "C" {
int uvc_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, float* y, int64_t ldy, float* mean,
                      float* rstd, int32_t M, int32_t C, int32_t round_tf32, void* stream) {
  UVC_REQUIRE(x && gamma && beta && y, UVC_ERR_BAD_ARG, "uvc_layernorm_fwd: NULL pointer");
  return uvc::layernorm_fwd(x, ldx, gamma, beta, eps, y, ldy, mean, rstd, M, C, UVC_ST, round_tf32);
}
int uvc_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                      const float* r1, const float* r2, const float* s2_dev, float* dx, int64_t lddx, float* dgamma, float* dbeta, int32_t M,
                      int32_t C, void* stream) {
  UVC_REQUIRE(dy && x && mean && rstd && gamma && dx, UVC_ERR_BAD_ARG, "uvc_layernorm_bwd: NULL pointer");
  UVC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), UVC_ERR_BAD_ARG, "uvc_layernorm_bwd: dgamma and dbeta must both be given or both NULL");
  return uvc::layernorm_bwd(dy, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, M, C, UVC_ST);
}
int uvc_layernorm_bwd_cs(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                         const float* r1, const float* r2, const float* s2_dev, float* dx, int64_t lddx, float* dgamma, float* dbeta, float* cs_r1,
                         float* cs_out, int32_t M, int32_t C, void* stream) {
  UVC_REQUIRE(dy && x && mean && rstd && gamma && dx, UVC_ERR_BAD_ARG, "uvc_layernorm_bwd_cs: NULL pointer");
  UVC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), UVC_ERR_BAD_ARG, "uvc_layernorm_bwd_cs: dgamma and dbeta must both be given or both NULL");
  return uvc::layernorm_bwd(dy, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, M, C, UVC_ST, cs_r1, cs_out);
}
int uvc_layernorm_fwd_f16(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y16, int64_t ldy, float* mean,
                          float* rstd, int32_t M, int32_t C, void* stream) {
  UVC_REQUIRE(x && gamma && beta && y16, UVC_ERR_BAD_ARG, "uvc_layernorm_fwd_f16: NULL pointer");
  return uvc::layernorm_fwd(x, ldx, gamma, beta, eps, nullptr, ldy, mean, rstd, M, C, UVC_ST, 0, y16);
}
int uvc_layernorm_bwd_f16(const void* dy16, int64_t lddy, float dy_scale, const float* x, int64_t ldx, const float* mean, const float* rstd,
                          const float* gamma, const float* r1, const float* r2, const float* s2_dev, float* dx, void* dx16, float dx16_scale,
                          int64_t lddx, float* dgamma, float* dbeta, float* cs_r1, float* cs_out, int32_t M, int32_t C, void* stream) {
  UVC_REQUIRE(dy16 && x && mean && rstd && gamma && dx, UVC_ERR_BAD_ARG, "uvc_layernorm_bwd_f16: NULL pointer");
  UVC_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), UVC_ERR_BAD_ARG, "uvc_layernorm_bwd_f16: dgamma and dbeta must both be given or both NULL");
  return uvc::layernorm_bwd(nullptr, lddy, x, ldx, mean, rstd, gamma, r1, r2, s2_dev, dx, lddx, dgamma, dbeta, M, C, UVC_ST, cs_r1, cs_out, dy16, dy_scale,
                            dx16, dx16_scale, nullptr);
}
int uvc_cvt_f16(const float* src, void* dst16, void* dstT16, int32_t rows, int32_t cols, void* stream) {
  UVC_REQUIRE(src && (dst16 || dstT16), UVC_ERR_BAD_ARG, "uvc_cvt_f16: NULL pointer");
  return uvc::cvt_f16_segs(&src, &dst16, dstT16 ? &dstT16 : nullptr, &rows, &cols, 1, UVC_ST);
}
int uvc_softmax_fwd(float* S, int64_t ld, int64_t rows, int32_t n, int32_t round_tf32, void* stream) {
  UVC_REQUIRE(S, UVC_ERR_BAD_ARG, "uvc_softmax_fwd: NULL pointer");
  return uvc::softmax_fwd(S, ld, rows, n, UVC_ST, round_tf32);
}
int uvc_softmax_bwd(const float* P, float* dP, int64_t ld, int64_t rows, int32_t n, float scale, int32_t round_tf32, void* stream) {
  UVC_REQUIRE(P && dP, UVC_ERR_BAD_ARG, "uvc_softmax_bwd: NULL pointer");
  return uvc::softmax_bwd(P, dP, ld, rows, n, scale, UVC_ST, round_tf32);
}
int uvc_colsum(const float* X, int64_t ld, int32_t M, int32_t N, const float* scale_dev, float* out, void* stream) {
  UVC_REQUIRE(X && out, UVC_ERR_BAD_ARG, "uvc_colsum: NULL pointer");
  return uvc::colsum(X, ld, M, N, scale_dev, out, UVC_ST);
}
int uvc_blend_fwd(const float* t, const float* x, const float* d, float* out, int64_t n, void* stream) {
  UVC_REQUIRE(t && x && d && out, UVC_ERR_BAD_ARG, "uvc_blend_fwd: NULL pointer");
  return uvc::blend_fwd(t, x, d, out, n, UVC_ST);
}
int uvc_blend_dots(const float* g, const float* t, const float* x, float* dots, int64_t n, void* stream) {
  UVC_REQUIRE(g && t && x && dots, UVC_ERR_BAD_ARG, "uvc_blend_dots: NULL pointer");
  return uvc::blend_dots(g, t, x, dots, n, UVC_ST);
}
int uvc_im2col16(const float* x, float* out, int32_t B, int32_t Cin, int32_t HW, int32_t P, int32_t round_tf32, void* stream) {
  UVC_REQUIRE(x && out, UVC_ERR_BAD_ARG, "uvc_im2col16: NULL pointer");
  return uvc::im2col16(x, out, B, Cin, HW, P, UVC_ST, round_tf32);
}
int uvc_assemble_tokens(const float* pe, const float* cls, const float* pos, const float* pscale, const float* tmask, float* tok, int32_t B,
                        int32_t np, int32_t C, void* stream) {
  UVC_REQUIRE(pe && cls && pos && tok, UVC_ERR_BAD_ARG, "uvc_assemble_tokens: NULL pointer");
  return uvc::assemble_tokens(pe, cls, pos, pscale, tmask, tok, B, np, C, UVC_ST);
}
int uvc_assemble_tokens_bwd(const float* g, const float* pe, const float* pscale, const float* tmask, float* dpe, float* dscale, float* dtmask,
                            float* dpos, float* dcls, int32_t B, int32_t np, int32_t C, void* stream) {
  UVC_REQUIRE(g && pe && dpe, UVC_ERR_BAD_ARG, "uvc_assemble_tokens_bwd: NULL pointer");
  return uvc::assemble_tokens_bwd(g, pe, pscale, tmask, dpe, dscale, dtmask, dpos, dcls, B, np, C, UVC_ST);
}
int uvc_round_tf32(const float* src, float* dst, int64_t n, void* stream) {
  UVC_REQUIRE(src && dst, UVC_ERR_BAD_ARG, "uvc_round_tf32: NULL pointer");
  const long long nn = n;
  return uvc::round_tf32_segs(&src, &dst, &nn, 1, UVC_ST);
}
int uvc_scale_add(float* y, const float* x, const float* s_dev, float s, int64_t n, void* stream) {
  UVC_REQUIRE(y && x, UVC_ERR_BAD_ARG, "uvc_scale_add: NULL pointer");
  return uvc::scale_add(y, x, s_dev, s, n, UVC_ST);
}
}  // extern "C"

// Attention core softmax(Q K^T * scale) V per (image, head) with 16-bit OPERAND STORAGE (fp16 in HBM and shared memory, tcgen05.mma
// kind::f16, fp32 accumulation in tensor memory, fp32 softmax) -- forward and recompute backward.
// Replaces models/model_distilled.py:175-185 and its autograd backward, like attention.cu; this file is the path the engine uses when
// the activations that feed GEMMs are stored as fp16 (same 10-bit mantissa as the TF32 path, half the bytes, twice the MMA rate).
//
// What 16-bit staging buys over the fp32-staged kernels of attention.cu (measured limits of those: every operand single-buffered because
// Q 64 + K 52 + V 52 KB filled shared memory; MN-major 32-bit operands need a second, differently swizzled copy):
//   * a head's Q (2 x 16 KB), K and V (26 KB each) are 84 KB, so the whole operand set is DOUBLE-BUFFERED: the TMA warp loads head i + 1
//     while head i computes, and no load latency is exposed between heads;
//   * 16-bit operands use the standard 128 B swizzle for K-major AND MN-major UMMA descriptors, so ONE staged copy of K serves S = Q K^T
//     (K-major B) and dQ = dS K (MN-major B); the backward's restaging loads are gone;
//   * P / dS are handed to the output MMAs as packed fp16 pairs in tensor memory (32-bit column c of the A operand = k 2c | 2c + 1).
#include "kernels.h"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace uvc {

namespace {

constexpr int kNK = 208;                       // key rows staged / score columns (N <= 208)
constexpr int kThreadsA = 320;                 // warp 0 TMA, warp 1 MMA + TMEM, warps 2-9 softmax / elementwise / epilogue
constexpr int kTileBytes = 128 * 128;          // one 128-row operand tile: 128 rows x 64 fp16
constexpr int kKVBytes = kNK * 128;            // one 208-row operand: 208 rows x 64 fp16 (26 KB, a multiple of 1024)


// ====================================================================================================================
// Forward.  One persistent CTA per SM walks (image, head) pairs.
//   S_t = Q_t K^T      128 x 208 x 64, A/B K-major from shared memory, D in TMEM columns [208 t, 208 t + 208)
//   P_t = exp2(...)    four softmax warps per query tile (thread = query row): row max, exp2, row sum in fp32; the un-normalised probabilities
//                      go back to TMEM IN PLACE as packed fp16 pairs (columns [208 t, 208 t + 104)): chunk c is read from columns [32 c, 32 c + 32)
//                      before its packed form is written to [16 c, 16 c + 16), so a thread only ever overwrites columns it has already consumed
//   O_t = P_t V        128 x 64 x 208, A from TMEM (packed fp16), B = V MN-major (same staged bytes a K-major descriptor would read)
//   ctx = O_t / sum    fp16, 32 x 64 blocks through swizzled shared memory and one TMA store per warp
// ====================================================================================================================
constexpr int kFStage = 2 * kTileBytes + 2 * kKVBytes;      // Q0 | Q1 | K | V = 84 KB
constexpr int kFStg = 8 * 4096;
constexpr int kFSmem = 2 * kFStage + kFStg + 1024;

struct alignas(64) Attn16FwdParams {
  CUtensorMap tmQ, tmK, tmV, tmO;
  float* lse;
  int B, H, N, ntiles;
  float scale_log2e;
};

__global__ void __launch_bounds__(kThreadsA, 1) attn16_fwd_kernel(const __grid_constant__ Attn16FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int s, int t) { return smem_base + (uint32_t)(s * kFStage + t * kTileBytes); };
  auto sK = [&](int s) { return smem_base + (uint32_t)(s * kFStage + 2 * kTileBytes); };
  auto sV = [&](int s) { return smem_base + (uint32_t)(s * kFStage + 2 * kTileBytes + kKVBytes); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 16 + 8u * s; };
  auto o_full = [&](int t) { return bar0 + 32 + 8u * t; };
  auto o_empty = [&](int t) { return bar0 + 48 + 8u * t; };
  auto s_full = [&](int t) { return bar0 + 64 + 8u * t; };
  auto p_ready = [&](int t) { return bar0 + 80 + 8u * t; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV); tma_prefetch_desc(&p.tmO);
    for (int s = 0; s < 2; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(o_full(t), 1); mbar_init(o_empty(t), 4); mbar_init(s_full(t), 1); mbar_init(p_ready(t), 4); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                                          // on-chip setup above overlaps the previous kernel's tail
  const int nheads = p.B * p.H;
  const int ntiles = p.ntiles;
  constexpr uint32_t kO = 2 * kNK;               // O accumulator columns [416, 480)

  if (warp == 0) {
    // ===================== TMA producer: head i + 1 is loaded while head i computes =====================
    uint32_t it = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++it) {
      const int b = hd / p.H, h = hd % p.H;
      const int s = it & 1;
      mbar_wait(empty(s), ((it >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full(s), (uint32_t)(ntiles * kTileBytes + 2 * kKVBytes));
        for (int t = 0; t < ntiles; ++t) tma_load_4d(sQ(s, t), &p.tmQ, full(s), 0, t * 128, h, b);
        tma_load_4d(sK(s), &p.tmK, full(s), 0, 0, h, b);
        tma_load_4d(sV(s), &p.tmV, full(s), 0, 0, h, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptors: D = F32 (bit 4), A = B = F16 (format 0), B MN-major (bit 16) for the output MMA, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(kNK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);              // S = Q K^T
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // O = P V
    const uint32_t d_hi = umma_desc_hi(1024, 2);      // 128 B swizzle, 8-row groups 1024 B apart: K-major and MN-major 16-bit operands alike
    // Issue order (the tensor pipe executes in order):  S(0,0) S(0,1) | O(h,0) S(h+1,0) O(h,1) S(h+1,1) | ...
    auto issue_scores = [&](int s, int t) {
      if (elect_one()) {
        const uint32_t q_lo = umma_desc_lo(sQ(s, t), 16), k_lo = umma_desc_lo(sK(s), 16);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)       // four K = 16 slices of 32 bytes along the 128-byte rows
          umma_f16_lh(tmem_base + t * kNK, q_lo + k4 * 2, d_hi, k_lo + k4 * 2, d_hi, idesc1, k4 ? 1u : 0u);
        umma_commit(s_full(t));
      }
      __syncwarp();
    };
    uint32_t it = 0;
    if ((int)blockIdx.x < nheads) {
      mbar_wait(full(0), 0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_scores(0, t);
    }
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = it & 1u;
      const bool more = hd + (int)gridDim.x < nheads;
      const uint32_t v_lo = umma_desc_lo(sV(s), kKVBytes);
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(p_ready(t), ph);                  // P_t is in TMEM and group t no longer reads S_t
        // the single O accumulator: its previous user must have read it out (previous tile of this head, or the last tile of the previous head)
        if (t > 0) mbar_wait(o_empty(t - 1), ph);
        else if (it > 0) mbar_wait(o_empty(ntiles - 1), ph ^ 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kNK / 16; ++ks)     // 13 K = 16 slices: 8 packed TMEM columns of P, 16 key rows (2048 B) of V
            umma_f16_ts(tmem_base + kO, tmem_base + t * kNK + ks * 8, v_lo + ks * 128, d_hi, idesc2, ks ? 1u : 0u);
          umma_commit(o_full(t));
        }
        __syncwarp();
        if (more) {
          if (t == 0) { mbar_wait(full(s ^ 1), ((it + 1) >> 1) & 1u); tc_fence_after(); }
          issue_scores(s ^ 1, t);
        }
      }
      if (elect_one()) umma_commit(empty(s));       // every MMA that reads this head's staged operands has been issued
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue =====================
    const int ew = warp - 2;
    const int t = ew >> 2;                           // query tile of this warp's group
    const int q = warp & 3;                          // TMEM lane quadrant
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t s_addr = lane_addr + t * kNK;
    const int row = t * 128 + q * 32 + lane;         // query row inside the head
    const int N = p.N;
    const uint32_t stage_blk = smem_base + 2 * kFStage + (uint32_t)ew * 4096u;
    const uint32_t stage_row = stage_blk + (uint32_t)lane * 128u;
    if (t < ntiles) {
      uint32_t it = 0;
      for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++it) {
        const int b = hd / p.H, h = hd % p.H;
        mbar_wait(s_full(t), it & 1u);
        tc_fence_after();
        // pass 1: row maximum over the valid key columns (columns >= N hold Q . 0 = 0 from the zero-filled key rows)
        const int full_chunks = min(6, N >> 5);
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < full_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
          float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(r[j])); m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(r[j + 2])); m3 = fmaxf(m3, __uint_as_float(r[j + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        }
#pragma unroll 1
        for (int c = full_chunks; c < 6; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) if (c * 32 + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        {
          uint32_t r[16];
          tmem_ld_32x16(s_addr + 192, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) if (192 + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        // pass 2: e = exp2((s - max) * scale * log2 e), un-normalised, packed to fp16 pairs in place as the A operand of the PV MMA.
        // The row sum uses the unrounded e (zero-mean 2^-12 / sqrt(N) relative difference to the sum of the rounded values).
        const float mxs = mx * p.scale_log2e;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll 1
        for (int c = 0; c < full_chunks; ++c) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -mxs));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -mxs));
            const float e2 = ex2_approx(fmaf(__uint_as_float(r[j + 2]), p.scale_log2e, -mxs));
            const float e3 = ex2_approx(fmaf(__uint_as_float(r[j + 3]), p.scale_log2e, -mxs));
            sum0 += e0; sum1 += e1; sum2 += e2; sum3 += e3;
            pk[j >> 1] = pack_half2(e0, e1); pk[(j >> 1) + 1] = pack_half2(e2, e3);
          }
          tmem_st_32x16(s_addr + c * 16, pk);
        }
#pragma unroll 1
        for (int c = full_chunks; c < 6; ++c) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -mxs));
            float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -mxs));
            if (c * 32 + j >= N) e0 = 0.f;
            if (c * 32 + j + 1 >= N) e1 = 0.f;
            sum0 += e0; sum1 += e1;
            pk[j >> 1] = pack_half2(e0, e1);
          }
          tmem_st_32x16(s_addr + c * 16, pk);
        }
        {
          uint32_t r[16], pk[8];
          tmem_ld_32x16(s_addr + 192, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -mxs));
            float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -mxs));
            if (192 + j >= N) e0 = 0.f;
            if (192 + j + 1 >= N) e1 = 0.f;
            sum2 += e0; sum3 += e1;
            pk[j >> 1] = pack_half2(e0, e1);
          }
          tmem_st_32x8(s_addr + 96, pk);
        }
        const float sum = (sum0 + sum1) + (sum2 + sum3);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready(t));
        const float inv = 1.0f / sum;
        if (p.lse && row < N) p.lse[((long long)b * p.H + h) * N + row] = mxs + log2f(sum);   // log2-domain log-sum-exp: P = exp2(S scale log2e - lse)
        // epilogue: O_t / rowsum -> ctx (fp16)
        mbar_wait(o_full(t), it & 1u);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(lane_addr + kO, o0);
        tmem_ld_32x32(lane_addr + kO + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty(t));
        if (t * 128 + q * 32 < N) {                  // warp-uniform; rows >= N inside the block are clipped by the tensor map
          if (lane == 0) tma_store_wait_read0();     // the previous store out of this block has read it
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g) {              // 16-byte granule g (8 fp16) of row `lane` sits at granule g ^ (row & 7): the 128-byte TMA swizzle
            const uint32_t* o = (g < 4) ? o0 : o1;
            const int j = (g & 3) * 8;
            const uint32_t v0 = pack_half2(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
            const uint32_t v1 = pack_half2(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
            const uint32_t v2 = pack_half2(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
            const uint32_t v3 = pack_half2(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (uint32_t)((g ^ (lane & 7)) << 4)), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&p.tmO, stage_blk, 0, t * 128 + q * 32, h, b);
            tma_store_commit();
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ====================================================================================================================
// Backward with recomputation: two kernels per layer, persistent over (image, head, 128-row tile).
//
//   phase 1 (query-major, produces dQ)          phase 2 (key-major, produces dK and dV)
//     X = Q_t K^T          (scores)               X = K_t' Q^T          (scores, transposed)
//     Y = dO_t V^T         (dP)                   Y = V_t' dO^T         (dP, transposed)
//     Y <- dS = scale * P .* (dP - D)             X <- P^T,  Y <- dS^T
//     dQ_t = dS K                                 dV_t' = P^T dO,  dK_t' = dS^T Q
//
// P = exp2(S * scale * log2e - lse2) from the forward's log-sum-exp, D = rowsum(dO .* O) from a small pre-kernel.  The B operands (K, V in
// phase 1; Q, dO in phase 2) are staged ONCE per head and serve the score MMAs (K-major) and the output MMAs (MN-major view of the same
// bytes).  Both the per-tile A operands and the per-head B operands are double-buffered.
// Eight elementwise warps, two per TMEM lane quadrant, split the 208 score columns [0, 112) / [112, 208).  P / dS return to TMEM as packed
// fp16 pairs; each warp writes them only over columns it has itself already read: the first warp ascending into [0, 56), the second walks
// its chunks downwards and fills [160, 208) -- the output MMAs pick the K = 16 slices up from those two places.
// TMEM: X [0, 208)  Y [208, 416)  accumulator [416, 480); phase 2 keeps its second accumulator (dK) in X's free columns [64, 128).
// ====================================================================================================================
constexpr int kBStageA = 2 * kTileBytes;        // two 128-row A tiles
constexpr int kBStageB = 2 * kKVBytes;          // two 208-row B operands
constexpr int kBSmem = 2 * kBStageA + 2 * kBStageB + 1024;

struct alignas(64) Attn16BwdParams {
  CUtensorMap tmA0, tmA1;      // A tiles (128-row boxes):   phase 1: Q, dO    phase 2: K, V
  CUtensorMap tmB0, tmB1;      // B operands (208-row boxes): phase 1: K, V     phase 2: Q, dO
  const float* lse; const float* Dv;    // [B, H, N]
  __half* out0; __half* out1;  // phase 1: dq, -   phase 2: dv, dk   (column 0 of the head-0 slice inside dqkv16, row stride ldo)
  float* db0; float* db1;      // optional bias gradients of the same slices (fp32, atomics), or NULL
  long long ldo;
  int B, H, N, ntiles;
  float scale, scale_log2e, db_scale;
  const float* db_scale_dev;   // optional device scalar multiplied into db_scale (the inverse loss scale chosen on the device)
};

__device__ __forceinline__ void warp_colsum32f(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float send = upper ? v[k] : v[k + s];
      const float keep = upper ? v[k + s] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

template <int PHASE>
__global__ void __launch_bounds__(kThreadsA, 1) attn16_bwd_kernel(const __grid_constant__ Attn16BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[256];
  __shared__ __align__(16) float s_D[256];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sA = [&](int s, int i) { return smem_base + (uint32_t)(s * kBStageA + i * kTileBytes); };
  auto sB = [&](int s, int i) { return smem_base + (uint32_t)(2 * kBStageA + s * kBStageB + i * kKVBytes); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 16 + 8u * s; };
  auto b_full = [&](int s) { return bar0 + 32 + 8u * s; };
  auto b_empty = [&](int s) { return bar0 + 48 + 8u * s; };
  const uint32_t sc_full = bar0 + 64, el_done = bar0 + 72, acc_full = bar0 + 80, acc_empty = bar0 + 88;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0); tma_prefetch_desc(&p.tmA1); tma_prefetch_desc(&p.tmB0); tma_prefetch_desc(&p.tmB1);
    for (int s = 0; s < 2; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(sc_full, 1); mbar_init(el_done, 8); mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                                          // on-chip setup above overlaps the previous kernel's tail
  const int nheads = p.B * p.H, ntiles = p.ntiles;
  constexpr uint32_t kX = 0, kY = kNK, kAcc = 2 * kNK, kAcc2 = 64;     // TMEM columns (kAcc2: inside X, free once the elementwise pass has read it)

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t ia = 0, ib = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++ib) {
      const int b = hd / p.H, h = hd % p.H;
      const int sb = ib & 1;
      mbar_wait(b_empty(sb), ((ib >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(b_full(sb), 2 * kKVBytes);
        tma_load_4d(sB(sb, 0), &p.tmB0, b_full(sb), 0, 0, h, b);
        tma_load_4d(sB(sb, 1), &p.tmB1, b_full(sb), 0, 0, h, b);
      }
      __syncwarp();
      for (int t = 0; t < ntiles; ++t, ++ia) {
        const int sa = ia & 1;
        mbar_wait(a_empty(sa), ((ia >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(a_full(sa), 2 * kTileBytes);
          tma_load_4d(sA(sa, 0), &p.tmA0, a_full(sa), 0, t * 128, h, b);
          tma_load_4d(sA(sa, 1), &p.tmA1, a_full(sa), 0, t * 128, h, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(kNK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);              // 128 x 208, K-major A and B
    constexpr uint32_t idesc2 = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // 128 x 64, A from TMEM, B MN-major
    const uint32_t d_hi = umma_desc_hi(1024, 2);
    // packed TMEM column of K = 16 slice ks of an A operand written by the elementwise warps: slices 0-6 at [0, 56), slices 7-12 at [160, 208)
    auto a_col = [](int ks) { return (uint32_t)(ks < 7 ? ks * 8 : 160 + (ks - 7) * 8); };
    uint32_t ia = 0, ib = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++ib) {
      const int sb = ib & 1;
      mbar_wait(b_full(sb), (ib >> 1) & 1u);
      const uint32_t b0_lo = umma_desc_lo(sB(sb, 0), 16), b1_lo = umma_desc_lo(sB(sb, 1), 16);                 // K-major view (LBO unused)
      const uint32_t m0_lo = umma_desc_lo(sB(sb, 0), kKVBytes), m1_lo = umma_desc_lo(sB(sb, 1), kKVBytes);     // MN-major view of the same bytes
      for (int t = 0; t < ntiles; ++t, ++ia) {
        const int sa = ia & 1;
        const uint32_t ph = ia & 1u;
        mbar_wait(a_full(sa), (ia >> 1) & 1u);
        mbar_wait(acc_empty, ph ^ 1u);               // previous tile's epilogue has drained X / Y / accumulators
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0_lo = umma_desc_lo(sA(sa, 0), 16), a1_lo = umma_desc_lo(sA(sa, 1), 16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) umma_f16_lh(tmem_base + kX, a0_lo + k4 * 2, d_hi, b0_lo + k4 * 2, d_hi, idesc1, k4 ? 1u : 0u);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) umma_f16_lh(tmem_base + kY, a1_lo + k4 * 2, d_hi, b1_lo + k4 * 2, d_hi, idesc1, k4 ? 1u : 0u);
          umma_commit(sc_full);
          umma_commit(a_empty(sa));
        }
        __syncwarp();
        mbar_wait(el_done, ph);                      // dS (and P^T) are in TMEM, X / Y have been read
        tc_fence_after();
        if (elect_one()) {
          if (PHASE == 1) {
#pragma unroll
            for (int ks = 0; ks < kNK / 16; ++ks) umma_f16_ts(tmem_base + kAcc, tmem_base + kY + a_col(ks), m0_lo + ks * 128, d_hi, idesc2, ks ? 1u : 0u);   // dQ = dS K
          } else {
#pragma unroll
            for (int ks = 0; ks < kNK / 16; ++ks) umma_f16_ts(tmem_base + kAcc, tmem_base + kX + a_col(ks), m1_lo + ks * 128, d_hi, idesc2, ks ? 1u : 0u);   // dV = P^T dO
#pragma unroll
            for (int ks = 0; ks < kNK / 16; ++ks) umma_f16_ts(tmem_base + kAcc2, tmem_base + kY + a_col(ks), m0_lo + ks * 128, d_hi, idesc2, ks ? 1u : 0u);  // dK = dS^T Q
          }
          umma_commit(acc_full);
          if (t == ntiles - 1) umma_commit(b_empty(sb));
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== elementwise pass + epilogue (8 warps) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                          // TMEM lane quadrant
    const int hh = ew >> 2;                          // column range: hh = 0 -> [0, 112), hh = 1 -> [112, 208)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int N = p.N;
    const float db_scale = p.db_scale * ((p.db0 && p.db_scale_dev) ? __ldg(p.db_scale_dev) : 1.0f);
    const int tid2 = threadIdx.x - 64;
    uint32_t ia = 0;
    // The per-row statistics are fetched one step AHEAD (phase 1: next tile's row values into registers; phase 2: the next head's 208 column
    // values, one per thread), so their DRAM latency is off the critical path of every tile.
    float nx_lse = INFINITY, nx_D = 0.f;
    auto fetch = [&](int hd2, int t2) {
      nx_lse = INFINITY; nx_D = 0.f;
      if (hd2 >= nheads) return;
      const long long sr = (long long)hd2 * N;     // (b*H + h) * N with hd2 = b*H + h
      const int idx = (PHASE == 1) ? t2 * 128 + q * 32 + lane : tid2;
      if (idx < N) { nx_lse = __ldg(p.lse + sr + idx); nx_D = __ldg(p.Dv + sr + idx); }
    };
    fetch(blockIdx.x, 0);
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x) {
      const int b = hd / p.H, h = hd % p.H;
      if (PHASE == 2) {                              // per-query statistics of this head, indexed by score column
        named_bar_sync(1, 256);
        s_lse[tid2] = nx_lse;
        s_D[tid2] = nx_D * p.scale;
        named_bar_sync(1, 256);
        fetch(hd + gridDim.x, 0);
      }
      for (int t = 0; t < ntiles; ++t, ++ia) {
        const int row = t * 128 + q * 32 + lane;     // phase 1: query row ; phase 2: key row
        float lse_r = INFINITY, D_r = 0.f;
        if (PHASE == 1) {
          lse_r = nx_lse; D_r = nx_D * p.scale;     // dS = scale P (dP - D) = P (scale dP - scale D)
          if (t + 1 < ntiles) fetch(hd, t + 1); else fetch(hd + gridDim.x, 0);
        }
        mbar_wait(sc_full, ia & 1u);
        tc_fence_after();
        // No validity test on the score columns: columns >= N multiply zero-filled rows of K (phase 1), and rows / columns >= N carry
        // lse = +inf, i.e. P = 0 (phase 2 key rows >= N produce values that are never stored).
        auto pd = [&](uint32_t x, uint32_t y, int col, float& pr, float& ds) {
          if (PHASE == 1) {
            pr = ex2_approx(fmaf(__uint_as_float(x), p.scale_log2e, -lse_r));
            ds = pr * fmaf(__uint_as_float(y), p.scale, -D_r);
          } else {
            pr = ex2_approx(fmaf(__uint_as_float(x), p.scale_log2e, -s_lse[col]));
            ds = pr * fmaf(__uint_as_float(y), p.scale, -s_D[col]);
          }
        };
        auto chunk32 = [&](int src, int dst) {
          uint32_t rx[32], ry[32], px[16], py[16];
          tmem_ld_32x32(lane_addr + kX + src, rx);
          tmem_ld_32x32(lane_addr + kY + src, ry);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0, p1, d0, d1;
            pd(rx[j], ry[j], src + j, p0, d0);
            pd(rx[j + 1], ry[j + 1], src + j + 1, p1, d1);
            if (PHASE == 2) px[j >> 1] = pack_half2(p0, p1);
            py[j >> 1] = pack_half2(d0, d1);
          }
          if (PHASE == 2) tmem_st_32x16(lane_addr + kX + dst, px);
          tmem_st_32x16(lane_addr + kY + dst, py);
        };
        if (hh == 0) {
#pragma unroll 1
          for (int c = 0; c < 3; ++c) chunk32(c * 32, c * 16);
          uint32_t rx[16], ry[16], px[8], py[8];
          tmem_ld_32x16(lane_addr + kX + 96, rx);
          tmem_ld_32x16(lane_addr + kY + 96, ry);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float p0, p1, d0, d1;
            pd(rx[j], ry[j], 96 + j, p0, d0);
            pd(rx[j + 1], ry[j + 1], 96 + j + 1, p1, d1);
            if (PHASE == 2) px[j >> 1] = pack_half2(p0, p1);
            py[j >> 1] = pack_half2(d0, d1);
          }
          if (PHASE == 2) tmem_st_32x8(lane_addr + kX + 48, px);
          tmem_st_32x8(lane_addr + kY + 48, py);
        } else {
#pragma unroll 1
          for (int c = 2; c >= 0; --c) chunk32(112 + c * 32, 160 + c * 16);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(el_done);
        // epilogue: this warp's 32 rows x its 32-column half of the 64-wide outputs
        mbar_wait(acc_full, ia & 1u);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(lane_addr + kAcc + hh * 32, o0);
        if (PHASE == 2) tmem_ld_32x32(lane_addr + kAcc2 + hh * 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        if (p.db0) {
          // qkv bias gradient = column sums of dq / dk / dv: a 32 x 32 transpose-reduce over the warp (31 shuffles) leaves column `lane` in each lane
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = row < N ? __uint_as_float(o0[j]) : 0.f;
          warp_colsum32f(v, lane);
          atomicAdd(p.db0 + h * 64 + hh * 32 + lane, v[0] * db_scale);
          if (PHASE == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = row < N ? __uint_as_float(o1[j]) : 0.f;
            warp_colsum32f(v, lane);
            atomicAdd(p.db1 + h * 64 + hh * 32 + lane, v[0] * db_scale);
          }
        }
        if (row < N) {
          const long long off = ((long long)b * N + row) * p.ldo + h * 64 + hh * 32;      // 64-byte aligned: ldo and the head offsets are multiples of 32 halves
          auto pk = [](uint32_t a, uint32_t c) { return pack_half2(__uint_as_float(a), __uint_as_float(c)); };
#pragma unroll
          for (int j = 0; j < 32; j += 16)
            st_global_v8u(p.out0 + off + j, pk(o0[j], o0[j + 1]), pk(o0[j + 2], o0[j + 3]), pk(o0[j + 4], o0[j + 5]), pk(o0[j + 6], o0[j + 7]),
                          pk(o0[j + 8], o0[j + 9]), pk(o0[j + 10], o0[j + 11]), pk(o0[j + 12], o0[j + 13]), pk(o0[j + 14], o0[j + 15]));
          if (PHASE == 2) {
#pragma unroll
            for (int j = 0; j < 32; j += 16)
              st_global_v8u(p.out1 + off + j, pk(o1[j], o1[j + 1]), pk(o1[j + 2], o1[j + 3]), pk(o1[j + 4], o1[j + 5]), pk(o1[j + 6], o1[j + 7]),
                            pk(o1[j + 8], o1[j + 9]), pk(o1[j + 10], o1[j + 11]), pk(o1[j + 12], o1[j + 13]), pk(o1[j + 14], o1[j + 15]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[b,h,row] = sum_j dO[row, h*64 + j] * O[row, h*64 + j]   (one warp per token row; 8 lanes x 8 halves cover one head)
__global__ void __launch_bounds__(256) attn16_rowdot_kernel(const __half* __restrict__ dO, const __half* __restrict__ O, float* __restrict__ Dv, int B, int H, int N) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), n = (int)(row % N);
  const int C8 = H * 8;
  const uint4* a = reinterpret_cast<const uint4*>(dO) + row * C8;
  const uint4* o = reinterpret_cast<const uint4*>(O) + row * C8;
  for (int c0 = 0; c0 < C8; c0 += 32) {
    const int c = c0 + lane;
    float v = 0.f;
    if (c < C8) {
      const uint4 x = a[c], y = o[c];
      const __half2* xh = reinterpret_cast<const __half2*>(&x);
      const __half2* yh = reinterpret_cast<const __half2*>(&y);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 xf = __half22float2(xh[i]), yf = __half22float2(yh[i]); v = fmaf(xf.x, yf.x, fmaf(xf.y, yf.y, v)); }
    }
#pragma unroll
    for (int s2 = 4; s2 > 0; s2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s2);
    if ((lane & 7) == 0 && c < C8) Dv[((long long)b * H + (c >> 3)) * N + n] = v;
  }
}

// 4-D view (head-dim, token, head, image) of a [B*N, ld] fp16 activation matrix whose columns are grouped in heads of 64
int attn16_tmap(CUtensorMap* tm, const void* base, long long ld, int B, int H, int N, unsigned box_rows, const char* name) {
  const unsigned long long dims[4] = {64, (unsigned long long)N, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long strides[3] = {(unsigned long long)ld * 2, 64 * 2, (unsigned long long)N * ld * 2};
  const unsigned int box[4] = {64, box_rows, 1, 1};
  return encode_tmap_4d_f16(tm, base, dims, strides, box, name);
}

int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace

bool attn_f16_ok(int N, int d) { return d == 64 && N >= 1 && N <= kNK; }

int attention_fwd_f16(const void* qkv16, void* ctx16, float* lse, int B, int H, int N, float scale, cudaStream_t st) {
  UVC_REQUIRE(B > 0 && H > 0 && attn_f16_ok(N, 64), UVC_ERR_BAD_SHAPE, "attention_fwd_f16: bad dims B=%d H=%d N=%d (N <= %d)", B, H, N, kNK);
  UVC_REQUIRE((reinterpret_cast<uintptr_t>(qkv16) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx16) & 15) == 0, UVC_ERR_BAD_SHAPE, "attention_fwd_f16: buffers must be 16 B aligned");
  const long long C = (long long)H * 64, ld3 = 3 * C;
  const __half* base = static_cast<const __half*>(qkv16);
  Attn16FwdParams kp;
  int rc;
  if ((rc = attn16_tmap(&kp.tmQ, base, ld3, B, H, N, 128, "attn16 Q"))) return rc;
  if ((rc = attn16_tmap(&kp.tmK, base + C, ld3, B, H, N, kNK, "attn16 K"))) return rc;
  if ((rc = attn16_tmap(&kp.tmV, base + 2 * C, ld3, B, H, N, kNK, "attn16 V"))) return rc;
  if ((rc = attn16_tmap(&kp.tmO, ctx16, C, B, H, N, 32, "attn16 ctx"))) return rc;
  kp.lse = lse; kp.B = B; kp.H = H; kp.N = N; kp.ntiles = (N + 127) / 128;
  kp.scale_log2e = scale * 1.4426950408889634f;
  cudaError_t e = cudaFuncSetAttribute(attn16_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(attn16_fwd smem=%d): %s", kFSmem, cudaGetErrorString(e));
  const int sms = sm_count();
  const int grid = B * H < sms ? B * H : sms;
  launch_pdl(attn16_fwd_kernel, dim3(grid), dim3(kThreadsA), kFSmem, st, kp);
  return check_launch("attn16_fwd_kernel");
}

int attention_bwd_f16(const void* qkv16, const float* lse, const void* ctx16, const void* dctx16, float* Dv, void* dqkv16, int B, int H, int N,
                      float scale, cudaStream_t st, float* dqkv_bias, float db_scale, const float* db_scale_dev) {
  UVC_REQUIRE(B > 0 && H > 0 && attn_f16_ok(N, 64), UVC_ERR_BAD_SHAPE, "attention_bwd_f16: bad dims B=%d H=%d N=%d (N <= %d)", B, H, N, kNK);
  const long long C = (long long)H * 64, ld3 = 3 * C;
  const __half* q = static_cast<const __half*>(qkv16);
  const __half* k = q + C; const __half* v = q + 2 * C;
  __half* dq = static_cast<__half*>(dqkv16);
  launch_pdl(attn16_rowdot_kernel, dim3((unsigned)(((long long)B * N + 7) / 8)), dim3(256), 0, st, static_cast<const __half*>(dctx16), static_cast<const __half*>(ctx16), Dv, B, H, N);
  int rc = check_launch("attn16_rowdot_kernel");
  if (rc) return rc;
  const int sms = sm_count();
  const int grid = B * H < sms ? B * H : sms;
  cudaError_t e = cudaFuncSetAttribute(attn16_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn16_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem);
  UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(attn16_bwd): %s", cudaGetErrorString(e));
  Attn16BwdParams kp;
  kp.lse = lse; kp.Dv = Dv; kp.ldo = ld3; kp.B = B; kp.H = H; kp.N = N; kp.ntiles = (N + 127) / 128;
  kp.scale = scale; kp.scale_log2e = scale * 1.4426950408889634f; kp.db_scale = db_scale; kp.db_scale_dev = db_scale_dev;
  // phase 1: dQ
  if ((rc = attn16_tmap(&kp.tmA0, q, ld3, B, H, N, 128, "attn16 bwd Q tile"))) return rc;
  if ((rc = attn16_tmap(&kp.tmA1, dctx16, C, B, H, N, 128, "attn16 bwd dO tile"))) return rc;
  if ((rc = attn16_tmap(&kp.tmB0, k, ld3, B, H, N, kNK, "attn16 bwd K"))) return rc;
  if ((rc = attn16_tmap(&kp.tmB1, v, ld3, B, H, N, kNK, "attn16 bwd V"))) return rc;
  kp.out0 = dq; kp.out1 = nullptr;
  kp.db0 = dqkv_bias; kp.db1 = nullptr;
  launch_pdl(attn16_bwd_kernel<1>, dim3(grid), dim3(kThreadsA), kBSmem, st, kp);
  if ((rc = check_launch("attn16_bwd_kernel<1>"))) return rc;
  // phase 2: dK, dV
  if ((rc = attn16_tmap(&kp.tmA0, k, ld3, B, H, N, 128, "attn16 bwd K tile"))) return rc;
  if ((rc = attn16_tmap(&kp.tmA1, v, ld3, B, H, N, 128, "attn16 bwd V tile"))) return rc;
  if ((rc = attn16_tmap(&kp.tmB0, q, ld3, B, H, N, kNK, "attn16 bwd Q"))) return rc;
  if ((rc = attn16_tmap(&kp.tmB1, dctx16, C, B, H, N, kNK, "attn16 bwd dO"))) return rc;
  kp.out0 = dq + 2 * C; kp.out1 = dq + C;
  kp.db0 = dqkv_bias ? dqkv_bias + 2 * C : nullptr; kp.db1 = dqkv_bias ? dqkv_bias + C : nullptr;
  launch_pdl(attn16_bwd_kernel<2>, dim3(grid), dim3(kThreadsA), kBSmem, st, kp);
  return check_launch("attn16_bwd_kernel<2>");
}

}  // namespace uvc

extern "C" int uvc_attention_fwd_f16(const void* qkv16, void* ctx16, float* lse, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv16 && ctx16, UVC_ERR_BAD_ARG, "uvc_attention_fwd_f16: NULL pointer");
  UVC_REQUIRE(uvc::attn_f16_ok(N, d), UVC_ERR_BAD_SHAPE, "uvc_attention_fwd_f16: needs d == 64 and N <= 208 (got d=%d, N=%d)", d, N);
  return uvc::attention_fwd_f16(qkv16, ctx16, lse, B, H, N, scale, static_cast<cudaStream_t>(stream));
}
extern "C" int uvc_attention_bwd_f16(const void* qkv16, const float* lse, const void* ctx16, const void* dctx16, float* D_ws, void* dqkv16,
                                     float* dqkv_bias, float db_scale, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv16 && lse && ctx16 && dctx16 && D_ws && dqkv16, UVC_ERR_BAD_ARG, "uvc_attention_bwd_f16: NULL pointer");
  UVC_REQUIRE(uvc::attn_f16_ok(N, d), UVC_ERR_BAD_SHAPE, "uvc_attention_bwd_f16: needs d == 64 and N <= 208 (got d=%d, N=%d)", d, N);
  return uvc::attention_bwd_f16(qkv16, lse, ctx16, dctx16, D_ws, dqkv16, B, H, N, scale, static_cast<cudaStream_t>(stream), dqkv_bias, db_scale, nullptr);
}

// Attention core softmax(Q K^T * scale) V per (image, head), forward and backward
// (models/model_distilled.py:175-185 and its autograd backward).
//
// Round-1 composition: batched tcgen05 TF32 GEMMs (gemm_tf32.cu) reading Q/K/V straight out of the
// [B*N, 3*H*d] qkv buffer through strided TMA descriptors (no permute/copy kernels), a warp-per-row
// softmax, and the PV GEMM writing ctx directly in [B*N, H*d] layout.  P is kept ([B,H,N,ldp]) for the
// backward, as the reference's autograd does.
#include "kernels.h"
#include <stdlib.h>

namespace uvc {

int attn_ldp(int N) { return (N + 3) / 4 * 4; }

// ====================================================================================================================
// Fused forward: one persistent CTA per SM walks (image, head) pairs; scores and probabilities never leave the SM.
//
//   S_t = Q_t K^T        tcgen05.mma kind::tf32, A = Q tile (128 query rows) and B = K (208 key rows, zero-filled past N) from
//                        TMA-staged swizzled shared memory, D = 128 lanes x 208 columns of TMEM; both query tiles (t = 0, 1) are
//                        issued back to back into their own column ranges
//   P_t = exp(S_t - max) eight softmax warps (four per query tile; warp w owns TMEM lanes 32 (w % 4) ..): each thread owns one query row,
//                        reads it with tcgen05.ld, writes the un-normalised TF32-rounded probabilities back IN PLACE with tcgen05.st
//   O_t = P_t V          tcgen05.mma with the A operand read straight from TMEM (the P just written), B = V (MN-major) from shared memory
//   ctx = O_t / rowsum   epilogue by the same warps: tcgen05.ld, scale, TF32 rounding, then 32 x 32 blocks through swizzled shared memory and a
//                        TMA store per warp (a thread owns a row: direct 128-bit stores hit 32 different lines per instruction and kept the
//                        LSU busy for ~1.6 us per head)
//
// With save_P the normalised probabilities are also written to HBM for the (still GEMM-composed) backward; the teacher / inference
// forward skips that and moves only qkv in and ctx out (the algorithmic minimum: 4 * B*N * 4C bytes).
// TMEM columns: S0 [0,208)  S1 [208,416)  O [416,480) (one accumulator, tiles take turns).  Shared memory: Q 64 KB, K 52 KB, V 52 KB, 32 KB ctx staging.
// All three operands are single-buffered (fp32 staging leaves no room for more), so the TMA warp also prefetches the NEXT head's boxes into L2
// when it loads the current one: the load that follows a freed buffer then pays L2 latency, not DRAM latency.
// ====================================================================================================================
// Bit pattern for a TMEM-sourced TF32 operand: the tensor core truncates the low 13 mantissa bits, so adding half a TF32 ulp first makes that
// truncation a round-to-nearest (2 instructions instead of cvt.rna's 3; the passes that call this are issue-bound).  The signed max keeps the
// hardware's canonical NaN (0x7fffffff, where the add would wrap into the sign bit) a NaN, so a diverged run still shows up downstream.
__device__ __forceinline__ uint32_t tf32_up(float x) {
  const uint32_t u = __float_as_uint(x), v = u + 0x1000u;      // unsigned add: no signed-overflow assumptions for the compiler to exploit
  return (uint32_t)max((int)v, (int)u);
}

constexpr int kANK = 208;                     // key rows staged / score columns (N <= 208)
constexpr int kAThreads = 320;                // warp 0 TMA, warp 1 MMA + TMEM, warps 2-9 softmax / epilogue
constexpr int kAQBytes = 2 * 2 * 128 * 128;   // 2 query tiles x 2 k-blocks x 128 rows x 128 B
constexpr int kAKBytes = 2 * kANK * 128;      // 2 k-blocks x 208 rows x 128 B
constexpr int kAVBytes = 2 * kANK * 128;      // 2 groups of 32 head-dims x 208 tokens x 128 B
constexpr int kAStage = 8 * 4096;             // per softmax warp: 32 rows x 32 columns of ctx staged for a TMA store
constexpr int kASmem = kAQBytes + kAKBytes + kAVBytes + kAStage + 1024;

struct alignas(64) AttnFwdParams {
  CUtensorMap tmQ, tmK, tmV, tmO;
  float* P; float* ctx; float* lse;
  long long ldp;
  int B, H, N, C, ntiles, save_P;
  float scale_log2e;
};

#ifdef UVC_ATTN_TRACE     // bring-up only (UVC_NVCC_EXTRA=-DUVC_ATTN_TRACE): per-warp event clocks of CTA 0 for heads 0..3
__device__ long long g_attn_trace[10 * 4 * 8];
#define ATR(ev) do { if (blockIdx.x == 0 && lane == 0 && it < 4) g_attn_trace[(warp * 4 + it) * 8 + (ev)] = clock64(); } while (0)
#else
#define ATR(ev) do { } while (0)
#endif

__global__ void __launch_bounds__(kAThreads, 1) attn_fwd_kernel(const __grid_constant__ AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sK = sQ + kAQBytes, sV = sK + kAKBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  // Every barrier is waited on by parties that see EACH of its phases (a parity wait is only meaningful for the current or the
  // immediately preceding phase), hence one o_full / o_empty per query tile although the tiles share one O accumulator.
  const uint32_t qk_full = bar0, v_full = bar0 + 8, qk_empty = bar0 + 16, v_empty = bar0 + 24;
  auto o_full = [&](int t) { return bar0 + 32 + 8u * t; };
  auto o_empty = [&](int t) { return bar0 + 48 + 8u * t; };
  auto s_full = [&](int t) { return bar0 + 64 + 8u * t; };
  auto p_ready = [&](int t) { return bar0 + 80 + 8u * t; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV); tma_prefetch_desc(&p.tmO);
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(qk_empty, 1); mbar_init(v_empty, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(o_full(t), 1); mbar_init(o_empty(t), 4); mbar_init(s_full(t), 1); mbar_init(p_ready(t), 4); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int nheads = p.B * p.H;
  const int ntiles = p.ntiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++it) {
      const int b = hd / p.H, h = hd % p.H;
      const int hn = hd + (int)gridDim.x, bn = hn / p.H, hh = hn % p.H;
      mbar_wait(qk_empty, (it & 1u) ^ 1u);
      ATR(0);
      if (elect_one()) {
        mbar_expect_tx(qk_full, (uint32_t)(ntiles * 2 * 128 * 128 + kAKBytes));
        for (int t = 0; t < ntiles; ++t)
          for (int kb = 0; kb < 2; ++kb) tma_load_4d(sQ + t * 32768 + kb * 16384, &p.tmQ, qk_full, kb * 32, t * 128, h, b);
        for (int kb = 0; kb < 2; ++kb) tma_load_4d(sK + kb * (kANK * 128), &p.tmK, qk_full, kb * 32, 0, h, b);
        if (hn < nheads) {                           // next head of this CTA: pull its Q / K boxes into L2 a full period ahead
          for (int c = 0; c < 2; ++c) {
            for (int t = 0; t < ntiles; ++t) tma_prefetch_l2_4d(&p.tmQ, c * 32, t * 128, hh, bn);
            tma_prefetch_l2_4d(&p.tmK, c * 32, 0, hh, bn);
          }
        }
      }
      __syncwarp();
      mbar_wait(v_empty, (it & 1u) ^ 1u);
      ATR(1);
      if (elect_one()) {
        mbar_expect_tx(v_full, kAVBytes);
        for (int c = 0; c < 2; ++c) tma_load_4d(sV + c * (kANK * 128), &p.tmV, v_full, c * 32, 0, h, b);
        if (hn < nheads)
          for (int c = 0; c < 2; ++c) tma_prefetch_l2_4d(&p.tmV, c * 32, 0, hh, bn);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kANK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);              // S = Q K^T
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // O = P V (B MN-major)
    const uint32_t k_hi = umma_desc_hi(1024, 2), v_hi = umma_desc_hi(512, 1);
    const uint32_t q_lo0 = umma_desc_lo(sQ, 16), k_lo0 = umma_desc_lo(sK, 16), v_lo0 = umma_desc_lo(sV, kANK * 128);
    // Issue order (the tensor pipe executes in order):  S(0,0) S(0,1) | O(h,0) S(h+1,0) O(h,1) S(h+1,1) | ...
    // The score MMA of the NEXT head for tile t goes out right behind the output MMA that consumes P_t, so group t finds its next scores ready
    // when it returns from the epilogue instead of waiting behind the other tile's output MMA (the two groups end up half a period apart).
    // Hazards: S_t/P_t is rewritten only after the output MMA that reads it (same pipe, in order) and after group t's last TMEM read of it
    // (p_ready[t]); the single O accumulator is handed from tile to tile through o_empty[].
    auto issue_scores = [&](int t) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t kb = kk >> 2, k4 = kk & 3;
          umma_tf32_lh(tmem_base + t * kANK, q_lo0 + ((t * 32768 + kb * 16384) >> 4) + k4 * 2, k_hi, k_lo0 + ((kb * (kANK * 128)) >> 4) + k4 * 2, k_hi, idesc1, kk ? 1u : 0u);
        }
        umma_commit(s_full(t));
      }
      __syncwarp();
    };
    uint32_t it = 0;
    if ((int)blockIdx.x < nheads) {
      mbar_wait(qk_full, 0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_scores(t);
      if (elect_one()) umma_commit(qk_empty);       // Q and K staging may be refilled for the next head
      __syncwarp();
    }
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++it) {
      const uint32_t ph = it & 1u;
      const bool more = hd + (int)gridDim.x < nheads;
      mbar_wait(v_full, ph);
      ATR(0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(p_ready(t), ph);                  // P_t is in TMEM and group t no longer reads S_t
        ATR(1 + 3 * t);
        // the single O accumulator: its previous user must have read it out (previous tile of this head, or the last tile of the previous head)
        if (t > 0) mbar_wait(o_empty(t - 1), ph);
        else if (it > 0) mbar_wait(o_empty(ntiles - 1), ph ^ 1u);
        ATR(2 + 3 * t);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 2
          for (int ks = 0; ks < kANK / 8; ++ks)
            umma_tf32_ts(tmem_base + 416, tmem_base + t * kANK + ks * 8, v_lo0 + ks * 64, v_hi, idesc2, ks ? 1u : 0u);
          umma_commit(o_full(t));
        }
        __syncwarp();
        if (more) {
          if (t == 0) { mbar_wait(qk_full, ph ^ 1u); tc_fence_after(); }
          ATR(3 + 3 * t);
          issue_scores(t);
          if (t == ntiles - 1) {
            if (elect_one()) umma_commit(qk_empty);
            __syncwarp();
          }
        }
      }
      if (elect_one()) umma_commit(v_empty);
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue =====================
    const int ew = warp - 2;
    const int t = ew >> 2;                           // query tile of this warp's group
    const int q = warp & 3;                          // TMEM lane quadrant
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t s_addr = lane_addr + t * kANK;
    const int row = t * 128 + q * 32 + lane;         // query row inside the head
    const int N = p.N;
    const uint32_t stage = sV + kAVBytes + (uint32_t)ew * 4096u + (uint32_t)lane * 128u;   // this thread's 128-byte row of the warp's block
    if (t < ntiles) {
      uint32_t it = 0;
      for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x, ++it) {
        const int b = hd / p.H, h = hd % p.H;
        mbar_wait(s_full(t), it & 1u);
        ATR(0);
        tc_fence_after();
        // pass 1: row maximum over the valid key columns (columns >= N hold Q . 0 = 0 from the zero-filled key rows).  Chunks that lie
        // entirely below N (all six for N >= 192) run without the per-column validity test: this loop and the next are issue-bound.
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
        // pass 2: e = exp((s - max) * scale), un-normalised, written back in place as the A operand of the PV MMA.  The tensor core
        // truncates the low 13 mantissa bits of a TF32 operand, so adding half a TF32 ulp to the bit pattern (e is in [0, 1]: no overflow)
        // makes that truncation a round-to-nearest: one integer add per element instead of the three-instruction cvt.rna sequence.
        // The row sum uses the unrounded e: it differs from the sum of the rounded values by a zero-mean 2^-12 / sqrt(N) relative error.
        ATR(1);
        const float mxs = mx * p.scale_log2e;
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll 1
        for (int c = 0; c < full_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -mxs));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[j + 1]), p.scale_log2e, -mxs));
            const float e2 = ex2_approx(fmaf(__uint_as_float(r[j + 2]), p.scale_log2e, -mxs));
            const float e3 = ex2_approx(fmaf(__uint_as_float(r[j + 3]), p.scale_log2e, -mxs));
            sum0 += e0; sum1 += e1; sum2 += e2; sum3 += e3;
            r[j] = tf32_up(e0); r[j + 1] = tf32_up(e1);
            r[j + 2] = tf32_up(e2); r[j + 3] = tf32_up(e3);
          }
          tmem_st_32x32(s_addr + c * 32, r);
        }
#pragma unroll 1
        for (int c = full_chunks; c < 6; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float e = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -mxs));
            if (c * 32 + j >= N) e = 0.f;
            sum0 += e;
            r[j] = (c * 32 + j >= N) ? 0u : tf32_up(e);
          }
          tmem_st_32x32(s_addr + c * 32, r);
        }
        {
          uint32_t r[16];
          tmem_ld_32x16(s_addr + 192, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e = ex2_approx(fmaf(__uint_as_float(r[j]), p.scale_log2e, -mxs));
            if (192 + j >= N) e = 0.f;
            sum1 += e;
            r[j] = (192 + j >= N) ? 0u : tf32_up(e);
          }
          tmem_st_32x16(s_addr + 192, r);
        }
        const float sum = (sum0 + sum1) + (sum2 + sum3);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        // p_ready also releases S_t for the next head's score MMA: with save_P the third pass below still reads it, so the arrive moves behind it
        if (!p.save_P && lane == 0) mbar_arrive(p_ready(t));
        ATR(2);
        const float inv = 1.0f / sum;
        if (p.lse && row < N) p.lse[((long long)b * p.H + h) * N + row] = mxs + log2f(sum);   // log2-domain log-sum-exp: P = exp2(S scale log2e - lse)
        // optional pass 3: normalised probabilities to HBM for the backward (TMEM loads are warp-collective: only the stores are predicated)
        if (p.save_P) {      // (the TMEM copy carries the rounding increment in its low 13 bits: mask them, as the tensor core does)
          float* prow = p.P + (((long long)b * p.H + h) * N + row) * p.ldp;
          const bool rok = row < N;
#pragma unroll 1
          for (int c = 0; c < 6; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(s_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (rok && c * 32 + j < N)    // ldp % 4 == 0 and ldp >= N: a float4 starting below N stays inside the row's padding
                *reinterpret_cast<float4*>(prow + c * 32 + j) = make_float4(round_tf32(__uint_as_float(r[j] & 0xffffe000u) * inv), round_tf32(__uint_as_float(r[j + 1] & 0xffffe000u) * inv),
                                                                             round_tf32(__uint_as_float(r[j + 2] & 0xffffe000u) * inv), round_tf32(__uint_as_float(r[j + 3] & 0xffffe000u) * inv));
            }
          }
          uint32_t r[16];
          tmem_ld_32x16(s_addr + 192, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            if (rok && 192 + j < N)
              *reinterpret_cast<float4*>(prow + 192 + j) = make_float4(round_tf32(__uint_as_float(r[j] & 0xffffe000u) * inv), round_tf32(__uint_as_float(r[j + 1] & 0xffffe000u) * inv),
                                                                       round_tf32(__uint_as_float(r[j + 2] & 0xffffe000u) * inv), round_tf32(__uint_as_float(r[j + 3] & 0xffffe000u) * inv));
          }
        }
        if (p.save_P) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_ready(t));
        }
        // epilogue: O_t / rowsum -> ctx
        mbar_wait(o_full(t), it & 1u);
        ATR(3);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(lane_addr + 416, o0);
        tmem_ld_32x32(lane_addr + 448, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty(t));
        ATR(4);
        if (t * 128 + q * 32 < N) {                  // warp-uniform; rows >= N inside the block are clipped by the tensor map
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t* o = half ? o1 : o0;
            if (lane == 0) tma_store_wait_read0();   // the previous store out of this block has read it
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {            // 16-byte granule g of row `lane` sits at granule g ^ (row & 7): the 128-byte TMA swizzle
              const uint32_t v0 = __float_as_uint(round_tf32(__uint_as_float(o[4 * g]) * inv));
              const uint32_t v1 = __float_as_uint(round_tf32(__uint_as_float(o[4 * g + 1]) * inv));
              const uint32_t v2 = __float_as_uint(round_tf32(__uint_as_float(o[4 * g + 2]) * inv));
              const uint32_t v3 = __float_as_uint(round_tf32(__uint_as_float(o[4 * g + 3]) * inv));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)((g ^ (lane & 7)) << 4)), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&p.tmO, stage, half * 32, t * 128 + q * 32, h, b);   // lane 0's row address is the block base
              tma_store_commit();
            }
          }
        }
        ATR(5);
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


static int check_attn(int B, int H, int N, int d) {
  UVC_REQUIRE(B > 0 && H > 0 && N > 0 && d > 0, UVC_ERR_BAD_SHAPE, "attention: bad dims B=%d H=%d N=%d d=%d", B, H, N, d);
  UVC_REQUIRE(N <= 256, UVC_ERR_BAD_SHAPE, "attention: N=%d tokens > 256 is not supported", N);
  UVC_REQUIRE((d & 3) == 0, UVC_ERR_BAD_SHAPE, "attention: head dim %d must be a multiple of 4", d);
  return UVC_OK;
}


// ====================================================================================================================
// Fused backward with recomputation (no saved probabilities): two kernels per layer, both persistent over (image, head, 128-row tile).
//
//   phase 1 (query-major, produces dQ)          phase 2 (key-major, produces dK and dV)
//     X = Q_t K^T          (scores)               X = K_t' Q^T          (scores, transposed)
//     Y = dO_t V^T         (dP)                   Y = V_t' dO^T         (dP, transposed)
//     Y <- dS = scale * P .* (dP - D)             X <- P^T,  Y <- dS^T
//     dQ_t = dS K          (A operand = Y in TMEM)  dV_t' = P^T dO (A = X in TMEM),  dK_t' = dS^T Q (A = Y in TMEM)
//
// with P = exp2(S * scale * log2e - lse2) recomputed from the forward's per-row log-sum-exp (lse2, 4 B per row instead of the 800 B row of P)
// and D = rowsum(dO .* O) from a small pre-kernel.  The transposed pass exists because the A operand of a TMEM-sourced MMA always has its
// rows on the TMEM lanes: dK and dV need key-major probabilities, which are cheaper to recompute (one 128 x 208 x 64 MMA) than to transpose.
// X and Y are 208 TMEM columns each; the output accumulators take 64 more (phase 2 parks dK in the first 64 columns of X once the dV MMA
// has consumed P^T).  Eight elementwise warps split each 32-lane quadrant's 208 columns in two halves: the backward needs no row reduction.
// Shared memory: two 32 KB A tiles + two (phase 1: three) 52 KB B operands; phase 2 reloads its B buffers with the MN-major (32 B swizzle atom)
// view of dO and Q for the output MMAs while the elementwise pass runs (the tensor core accepts 32-bit MN-major operands only in that layout,
// and K-major operands only in the 16 B atom layout, so one staged copy cannot serve both roles).
// ====================================================================================================================
// v[j] of lane l = element (row l, column j) of a 32 x 32 block; on return v[0] of lane c is the sum of column c over the 32 rows
__device__ __forceinline__ void warp_colsum32(float (&v)[32], int lane) {
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

constexpr int kBABytes = 2 * 128 * 128;       // one A tile: 2 k-blocks x 128 rows x 128 B
constexpr int kBBBytes = 2 * kANK * 128;      // one B operand: 2 x 208 rows x 128 B

struct alignas(64) AttnBwdParams {
  CUtensorMap tmA0, tmA1;      // A tiles (128-row boxes, K-major):           phase 1: Q, dO    phase 2: K, V
  CUtensorMap tmB0, tmB1;      // score B operands (208-row boxes, K-major):  phase 1: K, V     phase 2: Q, dO
  CUtensorMap tmC0, tmC1;      // output B operands (MN-major, 32 B atoms):   phase 1: K, -     phase 2: Q, dO
  const float* lse; const float* Dv;    // [B, H, N]
  float* out0; float* out1;    // phase 1: dq, -   phase 2: dv, dk   (pointers to column 0 of the head-0 slice inside dqkv, row stride ldo)
  float* db0; float* db1;      // optional bias gradients of the same slices (column sums over all rows, accumulated with atomics), or NULL
  long long ldo;
  int B, H, N, ntiles;
  float scale, scale_log2e;
};

#ifdef UVC_ATTN_TRACE
#define BTR(ev) do { if (blockIdx.x == 0 && lane == 0 && ia >= 2 && ia < 6) g_attn_trace[(warp * 4 + (ia - 2)) * 8 + (ev)] = clock64(); } while (0)
#else
#define BTR(ev) do { } while (0)
#endif

template <int PHASE>
__global__ void __launch_bounds__(kAThreads, 1) attn_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[12];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[256];
  __shared__ __align__(16) float s_D[256];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA0 = smem_base, sA1 = sA0 + kBABytes, sB0 = sA1 + kBABytes, sB1 = sB0 + kBBBytes, sB2 = sB1 + kBBBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(&bars[0]);
  const uint32_t a_full = bar0, a_empty = bar0 + 8, bs_full = bar0 + 16, bs_empty = bar0 + 24, bo_full = bar0 + 32, bo_empty = bar0 + 40,
                 sc_full = bar0 + 48, el_done = bar0 + 56, acc_full = bar0 + 64, acc_empty = bar0 + 72, mid = bar0 + 80;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0); tma_prefetch_desc(&p.tmA1); tma_prefetch_desc(&p.tmB0); tma_prefetch_desc(&p.tmB1); tma_prefetch_desc(&p.tmC0);
    if (PHASE == 2) tma_prefetch_desc(&p.tmC1);
    mbar_init(a_full, 1); mbar_init(a_empty, 1); mbar_init(bs_full, 1); mbar_init(bs_empty, 1); mbar_init(bo_full, 1); mbar_init(bo_empty, 1);
    mbar_init(sc_full, 1); mbar_init(el_done, 8); mbar_init(acc_full, 1); mbar_init(acc_empty, 8); mbar_init(mid, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int nheads = p.B * p.H, ntiles = p.ntiles;
  constexpr uint32_t kX = 0, kY = kANK, kAcc = 2 * kANK;     // TMEM columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t ia = 0, ibs = 0, ibo = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x) {
      const int b = hd / p.H, h = hd % p.H;
      for (int t = 0; t < ntiles; ++t, ++ia) {
        mbar_wait(a_empty, (ia & 1u) ^ 1u);
        BTR(0);
        if (elect_one()) {
          mbar_expect_tx(a_full, 2 * kBABytes);
          for (int kb = 0; kb < 2; ++kb) {
            tma_load_4d(sA0 + kb * 16384, &p.tmA0, a_full, kb * 32, t * 128, h, b);
            tma_load_4d(sA1 + kb * 16384, &p.tmA1, a_full, kb * 32, t * 128, h, b);
          }
        }
        __syncwarp();
        if (PHASE == 2 || t == 0) {
          mbar_wait(bs_empty, (ibs & 1u) ^ 1u);                    // score MMAs that read the previous contents are done
          if (PHASE == 2) mbar_wait(bo_empty, (ibo & 1u) ^ 1u);    // same memory was last used by the output MMAs of the previous tile
          BTR(1);
          if (elect_one()) {
            mbar_expect_tx(bs_full, 2 * kBBBytes);
            for (int kb = 0; kb < 2; ++kb) {
              tma_load_4d(sB0 + kb * (kANK * 128), &p.tmB0, bs_full, kb * 32, 0, h, b);
              tma_load_4d(sB1 + kb * (kANK * 128), &p.tmB1, bs_full, kb * 32, 0, h, b);
            }
          }
          __syncwarp();
          ++ibs;
        }
        if (PHASE == 1) {
          if (t == 0) {
            mbar_wait(bo_empty, (ibo & 1u) ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(bo_full, kBBBytes);
              for (int c = 0; c < 2; ++c) tma_load_4d(sB2 + c * (kANK * 128), &p.tmC0, bo_full, c * 32, 0, h, b);
            }
            __syncwarp();
            ++ibo;
          }
        } else {
          mbar_wait(bs_empty, (ibs - 1u) & 1u);                    // this tile's score MMAs have consumed the K-major copies: restage MN-major
          BTR(2);
          if (elect_one()) {
            mbar_expect_tx(bo_full, 2 * kBBBytes);
            for (int c = 0; c < 2; ++c) {
              tma_load_4d(sB0 + c * (kANK * 128), &p.tmC0, bo_full, c * 32, 0, h, b);
              tma_load_4d(sB1 + c * (kANK * 128), &p.tmC1, bo_full, c * 32, 0, h, b);
            }
          }
          __syncwarp();
          ++ibo;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kANK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);              // 128 x 208, K-major A and B
    constexpr uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // 128 x 64, B MN-major
    const uint32_t k_hi = umma_desc_hi(1024, 2), m_hi = umma_desc_hi(512, 1);
    const uint32_t a0_lo = umma_desc_lo(sA0, 16), a1_lo = umma_desc_lo(sA1, 16), b0_lo = umma_desc_lo(sB0, 16), b1_lo = umma_desc_lo(sB1, 16);
    const uint32_t m0_lo = umma_desc_lo(sB0, kANK * 128), m1_lo = umma_desc_lo(sB1, kANK * 128), m2_lo = umma_desc_lo(sB2, kANK * 128);
    uint32_t ia = 0, ibs = 0, ibo = 0;
    for (int hd = blockIdx.x; hd < nheads; hd += gridDim.x) {
      for (int t = 0; t < ntiles; ++t, ++ia) {
        const uint32_t ph = ia & 1u;
        mbar_wait(a_full, ph);
        BTR(0);
        if (PHASE == 2 || t == 0) { mbar_wait(bs_full, ibs & 1u); ++ibs; }
        BTR(1);
        mbar_wait(acc_empty, ph ^ 1u);               // previous tile's epilogue has drained X / Y / accumulators
        BTR(2);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t kb = kk >> 2, k4 = kk & 3;
            umma_tf32_lh(tmem_base + kX, a0_lo + ((kb * 16384) >> 4) + k4 * 2, k_hi, b0_lo + ((kb * (kANK * 128)) >> 4) + k4 * 2, k_hi, idesc1, kk ? 1u : 0u);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t kb = kk >> 2, k4 = kk & 3;
            umma_tf32_lh(tmem_base + kY, a1_lo + ((kb * 16384) >> 4) + k4 * 2, k_hi, b1_lo + ((kb * (kANK * 128)) >> 4) + k4 * 2, k_hi, idesc1, kk ? 1u : 0u);
          }
          umma_commit(sc_full);
          umma_commit(a_empty);
          if (PHASE == 2 || t == ntiles - 1) umma_commit(bs_empty);
        }
        __syncwarp();
        mbar_wait(el_done, ph);                      // dS (and P^T) are in TMEM
        BTR(3);
        if (PHASE == 2 || t == 0) { mbar_wait(bo_full, ibo & 1u); ++ibo; }
        BTR(4);
        tc_fence_after();
        if (PHASE == 1) {
          if (elect_one()) {
#pragma unroll 2
            for (int ks = 0; ks < kANK / 8; ++ks) umma_tf32_ts(tmem_base + kAcc, tmem_base + kY + ks * 8, m2_lo + ks * 64, m_hi, idesc2, ks ? 1u : 0u);   // dQ = dS K
            umma_commit(acc_full);
            if (t == ntiles - 1) umma_commit(bo_empty);
          }
          __syncwarp();
        } else {
          if (elect_one()) {
#pragma unroll 2
            for (int ks = 0; ks < kANK / 8; ++ks) umma_tf32_ts(tmem_base + kAcc, tmem_base + kX + ks * 8, m1_lo + ks * 64, m_hi, idesc2, ks ? 1u : 0u);   // dV = P^T dO
            umma_commit(mid);
          }
          __syncwarp();
          mbar_wait(mid, ph);                        // P^T consumed: its first 64 columns become the dK accumulator
          BTR(5);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll 2
            for (int ks = 0; ks < kANK / 8; ++ks) umma_tf32_ts(tmem_base + kX, tmem_base + kY + ks * 8, m0_lo + ks * 64, m_hi, idesc2, ks ? 1u : 0u);     // dK = dS^T Q
            umma_commit(acc_full);
            umma_commit(bo_empty);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== elementwise pass + epilogue (8 warps) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                          // TMEM lane quadrant
    const int hh = ew >> 2;                          // column half: [104 hh, 104 hh + 104)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int N = p.N;
    const int tid2 = threadIdx.x - 64;
    uint32_t ia = 0;
    // The per-row statistics are fetched one step AHEAD (phase 1: next tile's row values into registers; phase 2: the next head's 208 column
    // values, one per thread): issued right before use, their DRAM latency sat on the critical path of every tile (10-15 % of the kernel).
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
        BTR(0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col0 = hh * 104 + c * 32;
          if (c < 3) {
            uint32_t rx[32], ry[32];
            tmem_ld_32x32(lane_addr + kX + col0, rx);
            tmem_ld_32x32(lane_addr + kY + col0, ry);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              // P and dS feed TMEM-sourced MMAs, which truncate to TF32: + half a TF32 ulp on the bit pattern turns that into round-to-nearest
              // (one integer add instead of cvt.rna's three instructions; this loop is issue-bound).  No validity test on the score columns:
              // columns >= N multiply zero-filled rows of K (phase 1), and rows / columns >= N carry lse = +inf, i.e. P = 0.
              float pr, ds;
              if (PHASE == 1) {
                pr = ex2_approx(fmaf(__uint_as_float(rx[j]), p.scale_log2e, -lse_r));
                ds = pr * fmaf(__uint_as_float(ry[j]), p.scale, -D_r);
              } else {
                pr = ex2_approx(fmaf(__uint_as_float(rx[j]), p.scale_log2e, -s_lse[col0 + j]));
                ds = pr * fmaf(__uint_as_float(ry[j]), p.scale, -s_D[col0 + j]);
                rx[j] = tf32_up(pr);
              }
              ry[j] = tf32_up(ds);
            }
            if (PHASE == 2) tmem_st_32x32(lane_addr + kX + col0, rx);
            tmem_st_32x32(lane_addr + kY + col0, ry);
          } else {
            uint32_t rx[8], ry[8];
            tmem_ld_32x8(lane_addr + kX + col0, rx);
            tmem_ld_32x8(lane_addr + kY + col0, ry);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // P and dS feed TMEM-sourced MMAs, which truncate to TF32: + half a TF32 ulp on the bit pattern turns that into round-to-nearest
              // (one integer add instead of cvt.rna's three instructions; this loop is issue-bound).  No validity test on the score columns:
              // columns >= N multiply zero-filled rows of K (phase 1), and rows / columns >= N carry lse = +inf, i.e. P = 0.
              float pr, ds;
              if (PHASE == 1) {
                pr = ex2_approx(fmaf(__uint_as_float(rx[j]), p.scale_log2e, -lse_r));
                ds = pr * fmaf(__uint_as_float(ry[j]), p.scale, -D_r);
              } else {
                pr = ex2_approx(fmaf(__uint_as_float(rx[j]), p.scale_log2e, -s_lse[col0 + j]));
                ds = pr * fmaf(__uint_as_float(ry[j]), p.scale, -s_D[col0 + j]);
                rx[j] = tf32_up(pr);
              }
              ry[j] = tf32_up(ds);
            }
            if (PHASE == 2) tmem_st_32x8(lane_addr + kX + col0, rx);
            tmem_st_32x8(lane_addr + kY + col0, ry);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(el_done);
        BTR(1);
        // epilogue: this warp's 32 rows x its 32-column half of the 64-wide outputs
        mbar_wait(acc_full, ia & 1u);
        BTR(2);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(lane_addr + kAcc + hh * 32, o0);
        if (PHASE == 2) tmem_ld_32x32(lane_addr + kX + hh * 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        BTR(3);
        if (p.db0) {
          // qkv bias gradient = column sums of dq / dk / dv: a 32 x 32 transpose-reduce over the warp (31 shuffles) leaves column `lane` in each lane
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = row < N ? __uint_as_float(o0[j]) : 0.f;
          warp_colsum32(v, lane);
          atomicAdd(p.db0 + h * 64 + hh * 32 + lane, v[0]);
          if (PHASE == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = row < N ? __uint_as_float(o1[j]) : 0.f;
            warp_colsum32(v, lane);
            atomicAdd(p.db1 + h * 64 + hh * 32 + lane, v[0]);
          }
        }
        BTR(4);
        if (row < N) {
          const long long off = ((long long)b * N + row) * p.ldo + h * 64 + hh * 32;      // 128-byte aligned: ldo and the head offsets are multiples of 32 floats
          auto rnd = [](uint32_t u) { return round_tf32(__uint_as_float(u)); };    // cvt.rna: keeps NaN / inf (the integer shortcut used for
                                                                                    // the TMEM operands would turn a NaN into -0 here and hide a diverged run)
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            st_global_v8(p.out0 + off + j, rnd(o0[j]), rnd(o0[j + 1]), rnd(o0[j + 2]), rnd(o0[j + 3]), rnd(o0[j + 4]), rnd(o0[j + 5]), rnd(o0[j + 6]), rnd(o0[j + 7]));
          if (PHASE == 2) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              st_global_v8(p.out1 + off + j, rnd(o1[j]), rnd(o1[j + 1]), rnd(o1[j + 2]), rnd(o1[j + 3]), rnd(o1[j + 4]), rnd(o1[j + 5]), rnd(o1[j + 6]), rnd(o1[j + 7]));
          }
        }
        BTR(5);
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

// D[b,h,row] = sum_j dO[row, h*64 + j] * O[row, h*64 + j]   (one warp per token row; 16 lanes x float4 cover one head)
__global__ void __launch_bounds__(256) attn_rowdot_kernel(const float* __restrict__ dO, const float* __restrict__ O, float* __restrict__ Dv, int B, int H, int N) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), n = (int)(row % N);
  const int C4 = H * 16;
  const float4* a = reinterpret_cast<const float4*>(dO) + row * C4;
  const float4* o = reinterpret_cast<const float4*>(O) + row * C4;
  for (int c0 = 0; c0 < C4; c0 += 32) {
    const int c = c0 + lane;
    float v = 0.f;
    if (c < C4) { const float4 x = a[c], y = o[c]; v = (x.x * y.x + x.y * y.y) + (x.z * y.z + x.w * y.w); }
#pragma unroll
    for (int s2 = 8; s2 > 0; s2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s2);
    if ((lane & 15) == 0 && c < C4) Dv[((long long)b * H + (c >> 4)) * N + n] = v;
  }
}

static int attn_fused_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("UVC_ATTN_FUSED"); mode = e ? atoi(e) : 1; }
  return mode;
}

// fused forward (d == 64, N <= 208): P may be NULL (inference / teacher: probabilities are never materialised)
static int attention_fwd_fused(const float* qkv, float* P, float* ctx, float* lse, int B, int H, int N, float scale, cudaStream_t st) {
  const long long C = (long long)H * 64, ld3 = 3 * C;
  AttnFwdParams kp;
  const unsigned long long dims[4] = {64, (unsigned long long)N, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long strides[3] = {(unsigned long long)ld3 * 4, 64 * 4, (unsigned long long)N * ld3 * 4};
  const unsigned int boxq[4] = {32, 128, 1, 1}, boxk[4] = {32, (unsigned)kANK, 1, 1};
  int rc;
  if ((rc = encode_tmap_4d(&kp.tmQ, qkv, dims, strides, boxq, false, "attn Q"))) return rc;
  if ((rc = encode_tmap_4d(&kp.tmK, qkv + C, dims, strides, boxk, false, "attn K"))) return rc;
  if ((rc = encode_tmap_4d(&kp.tmV, qkv + 2 * C, dims, strides, boxk, true, "attn V"))) return rc;
  const unsigned long long ostrides[3] = {(unsigned long long)C * 4, 64 * 4, (unsigned long long)N * C * 4};
  const unsigned int boxo[4] = {32, 32, 1, 1};
  if ((rc = encode_tmap_4d(&kp.tmO, ctx, dims, ostrides, boxo, false, "attn ctx"))) return rc;
  kp.P = P; kp.ctx = ctx; kp.lse = lse; kp.ldp = attn_ldp(N);
  kp.B = B; kp.H = H; kp.N = N; kp.C = (int)C; kp.ntiles = (N + 127) / 128; kp.save_P = P != nullptr;
  kp.scale_log2e = scale * 1.4426950408889634f;
  static unsigned long long attr_devs = 0;       // devices this kernel's attribute has been set on
  if (first_on_device(&attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kASmem);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(attn_fwd smem=%d): %s", kASmem, cudaGetErrorString(e));
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = B * H < sms ? B * H : sms;
  attn_fwd_kernel<<<grid, kAThreads, kASmem, st>>>(kp);
  return check_launch("attn_fwd_kernel");
}

#ifdef UVC_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int uvc_attn_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(long long) * 10 * 4 * 8);
}
#endif

bool attn_fused_ok(int N, int d) { return attn_fused_mode() != 0 && d == 64 && N <= kANK; }

int attention_fwd(const float* qkv, float* P, float* ctx, int B, int H, int N, int d, float scale, cudaStream_t st, bool need_P, float* lse) {
  int rc = check_attn(B, H, N, d);
  if (rc) return rc;
  if (attn_fused_ok(N, d) && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0 &&
      (!P || (reinterpret_cast<uintptr_t>(P) & 15) == 0))
    return attention_fwd_fused(qkv, need_P ? P : nullptr, ctx, lse, B, H, N, scale, st);
  UVC_REQUIRE(lse == nullptr, UVC_ERR_BAD_SHAPE, "attention_fwd: the log-sum-exp output needs the fused kernel (d == 64, N <= %d)", kANK);
  UVC_REQUIRE(P != nullptr, UVC_ERR_BAD_ARG, "attention_fwd: the GEMM-composed path needs a P buffer");
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


// 4-D view (head-dim, token, head, image) of a [B*N, ld] activation matrix whose columns are grouped in heads of 64
static int attn_tmap(CUtensorMap* tm, const float* base, long long ld, int B, int H, int N, unsigned box_rows, bool atom32, const char* name) {
  const unsigned long long dims[4] = {64, (unsigned long long)N, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long strides[3] = {(unsigned long long)ld * 4, 64 * 4, (unsigned long long)N * ld * 4};
  const unsigned int box[4] = {32, box_rows, 1, 1};
  return encode_tmap_4d(tm, base, dims, strides, box, atom32, name);
}

// fused backward with recomputation: needs the forward's lse [B,H,N], ctx and a [B,H,N] scratch for D = rowsum(dctx .* ctx)
int attention_bwd_fused(const float* qkv, const float* lse, const float* ctx, const float* dctx, float* Dv, float* dqkv, int B, int H, int N, float scale,
                        cudaStream_t st, float* dqkv_bias) {
  UVC_REQUIRE(N <= kANK, UVC_ERR_BAD_SHAPE, "attention_bwd_fused: N=%d > %d", N, kANK);
  const long long C = (long long)H * 64, ld3 = 3 * C;
  const float* q = qkv; const float* k = qkv + C; const float* v = qkv + 2 * C;
  attn_rowdot_kernel<<<(unsigned)(((long long)B * N + 7) / 8), 256, 0, st>>>(dctx, ctx, Dv, B, H, N);
  int rc = check_launch("attn_rowdot_kernel");
  if (rc) return rc;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = B * H < sms ? B * H : sms;
  constexpr int smem1 = 2 * kBABytes + 3 * kBBBytes + 1024, smem2 = 2 * kBABytes + 2 * kBBBytes + 1024;
  static unsigned long long attr_devs = 0;       // devices this kernel's attribute has been set on
  if (first_on_device(&attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(attn_bwd): %s", cudaGetErrorString(e));
  }
  AttnBwdParams kp;
  kp.lse = lse; kp.Dv = Dv; kp.ldo = ld3; kp.B = B; kp.H = H; kp.N = N; kp.ntiles = (N + 127) / 128;
  kp.scale = scale; kp.scale_log2e = scale * 1.4426950408889634f;
  // phase 1: dQ
  if ((rc = attn_tmap(&kp.tmA0, q, ld3, B, H, N, 128, false, "attn bwd Q tile"))) return rc;
  if ((rc = attn_tmap(&kp.tmA1, dctx, C, B, H, N, 128, false, "attn bwd dO tile"))) return rc;
  if ((rc = attn_tmap(&kp.tmB0, k, ld3, B, H, N, kANK, false, "attn bwd K"))) return rc;
  if ((rc = attn_tmap(&kp.tmB1, v, ld3, B, H, N, kANK, false, "attn bwd V"))) return rc;
  if ((rc = attn_tmap(&kp.tmC0, k, ld3, B, H, N, kANK, true, "attn bwd K (MN)"))) return rc;
  kp.tmC1 = kp.tmC0;
  kp.out0 = dqkv; kp.out1 = nullptr;
  kp.db0 = dqkv_bias; kp.db1 = nullptr;
  attn_bwd_kernel<1><<<grid, kAThreads, smem1, st>>>(kp);
  if ((rc = check_launch("attn_bwd_kernel<1>"))) return rc;
#ifdef UVC_ATTN_TRACE
  if (getenv("UVC_TRACE_PHASE1")) return UVC_OK;
#endif
  // phase 2: dK, dV
  if ((rc = attn_tmap(&kp.tmA0, k, ld3, B, H, N, 128, false, "attn bwd K tile"))) return rc;
  if ((rc = attn_tmap(&kp.tmA1, v, ld3, B, H, N, 128, false, "attn bwd V tile"))) return rc;
  if ((rc = attn_tmap(&kp.tmB0, q, ld3, B, H, N, kANK, false, "attn bwd Q"))) return rc;
  if ((rc = attn_tmap(&kp.tmB1, dctx, C, B, H, N, kANK, false, "attn bwd dO"))) return rc;
  if ((rc = attn_tmap(&kp.tmC0, q, ld3, B, H, N, kANK, true, "attn bwd Q (MN)"))) return rc;
  if ((rc = attn_tmap(&kp.tmC1, dctx, C, B, H, N, kANK, true, "attn bwd dO (MN)"))) return rc;
  kp.out0 = dqkv + 2 * C; kp.out1 = dqkv + C;
  kp.db0 = dqkv_bias ? dqkv_bias + 2 * C : nullptr; kp.db1 = dqkv_bias ? dqkv_bias + C : nullptr;
  attn_bwd_kernel<2><<<grid, kAThreads, smem2, st>>>(kp);
  return check_launch("attn_bwd_kernel<2>");
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
  UVC_REQUIRE(qkv && ctx, UVC_ERR_BAD_ARG, "uvc_attention_fwd: NULL pointer");
  return uvc::attention_fwd(qkv, P, ctx, B, H, N, d, scale, static_cast<cudaStream_t>(stream), P != nullptr);
}
extern "C" int uvc_attention_fwd_lse(const float* qkv, float* lse, float* ctx, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv && lse && ctx, UVC_ERR_BAD_ARG, "uvc_attention_fwd_lse: NULL pointer");
  return uvc::attention_fwd(qkv, nullptr, ctx, B, H, N, d, scale, static_cast<cudaStream_t>(stream), false, lse);
}
extern "C" int uvc_attention_bwd_fused(const float* qkv, const float* lse, const float* ctx, const float* dctx, float* D_ws, float* dqkv,
                                       float* dqkv_bias, int32_t B, int32_t H, int32_t N, int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv && lse && ctx && dctx && D_ws && dqkv, UVC_ERR_BAD_ARG, "uvc_attention_bwd_fused: NULL pointer");
  UVC_REQUIRE(uvc::attn_fused_ok(N, d), UVC_ERR_BAD_SHAPE, "uvc_attention_bwd_fused: needs d == 64 and N <= 208 (got d=%d, N=%d)", d, N);
  return uvc::attention_bwd_fused(qkv, lse, ctx, dctx, D_ws, dqkv, B, H, N, scale, static_cast<cudaStream_t>(stream), dqkv_bias);
}
extern "C" int uvc_attention_bwd(const float* qkv, const float* P, const float* dctx, float* dP, float* dqkv, int32_t B, int32_t H, int32_t N,
                                 int32_t d, float scale, void* stream) {
  UVC_REQUIRE(qkv && P && dctx && dP && dqkv, UVC_ERR_BAD_ARG, "uvc_attention_bwd: NULL pointer");
  return uvc::attention_bwd(qkv, P, dctx, dP, dqkv, B, H, N, d, scale, static_cast<cudaStream_t>(stream));
}

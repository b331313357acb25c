// Batched TF32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulator in TMEM, operands staged
// by TMA into 128B-swizzled shared memory).  One CTA computes one 128 x BN output tile:
//
//   warp 0   : TMA producer   (one elected lane; ring of STAGES {A,B} k-blocks of 32 fp32 = 128 B)
//   warp 1   : TMEM allocator + MMA issuer (one lane issues 4 x tcgen05.mma kind::tf32, K=8 each, per stage)
//   warps 2-5: epilogue       (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> HBM)
//
// Operands may be K-major (rows index M|N) or MN-major (rows index K), so forward, dgrad and wgrad
// GEMMs of every nn.Linear on the path need no transpose pass:
//   K-major : TMA box 32(k) x rows,  SWIZZLE_128B,          UMMA layout SWIZZLE_128B   (SBO 1024 B)
//   MN-major: TMA box 32(mn) x 32(k), SWIZZLE_128B_ATOM_32B, UMMA layout 128B_BASE32B  (SBO 512 B, LBO 4096 B)
// Two CTAs are resident per SM (96 KB smem + 128 TMEM columns each) so one tile's epilogue overlaps the
// other's main loop.  Split-K (red.global.add) covers the long-K / few-tile weight-gradient GEMMs.
//
// Replaces: models/model_distilled.py:116,122,149,175,179,184,187,522 and their autograd backward.
#include "kernels.h"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>

namespace uvc {

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements per k-block = one 128 B swizzle row
constexpr int kThreads = 192;

struct alignas(64) GemmKParams {
  CUtensorMap tmA, tmB;
  float* D; long long ldd, d_bs1, d_bs2;
  const float* bias;
  const float* R; long long ldr, r_bs1, r_bs2;
  const float* R2; long long ldr2;      // v2, UVC_EPI_BLEND: second residual (the block input x)
  float* D2; long long ldd2;            // v2, UVC_EPI_BLEND: optional copy of the un-blended value t
  const float* blend_dev;               // v2, UVC_EPI_BLEND: device (d0, d1)
  float* aux; long long ldaux, aux_bs1, aux_bs2;
  const float* alpha_dev; const float* beta_dev; const float* alpha_dev2; const float* colsum_scale_dev;
  float* colsum;
  void* D16; long long ldd16;           // optional fp16 copy of the output (v2 kernel), row stride in fp16 elements
  float alpha, beta, colsum_scale;
  int M, N, K, nb1, nb2, splits, flags;
  int a_mn, b_mn;
  int a_use1, a_use2, b_use1, b_use2;   // operand varies with batch index i1 / i2 (else coordinate 0)
  int f16;                              // v2: A and B are fp16 in memory (K-major), MMA kind::f16
  int aux16;                            // v2: aux holds fp16 values (UVC_EPI_AUX_F16)
  int a_grp, b_grp;                     // v2: MN-major operand described by a grouped tensor map (one TMA op per stage)
  int m_tiles, n_tiles, units;          // v2 (persistent CTA-pair kernel): 256-row x BN-column tiles, units = tiles * splits
};

template <int BN, int STAGES>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + slack for the 1024 B round-up
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, (BN <= 128 ? 2 : 1))
gemm_tf32_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = GemmCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int zs = blockIdx.z;
  const int split = zs % p.splits;
  const int z = zs / p.splits;
  const int i1 = z % p.nb1, i2 = z / p.nb1;

  const int bke = p.f16 ? 2 * BK : BK;                 // elements per 128-byte k-block row: 32 fp32 or 64 fp16
  const int nkb_total = (p.K + bke - 1) / bke;
  const int kb0 = (int)(((long long)nkb_total * split) / p.splits);
  const int kb1 = (int)(((long long)nkb_total * (split + 1)) / p.splits);
  const int nkb = kb1 - kb0;

  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[STAGES + s]); };
  const uint32_t tmem_full_bar = smem_u32(&bars[2 * STAGES]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), BN);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                                          // everything above is on-chip setup; operands / outputs belong to the stream order

  if (warp == 0) {
    // ===================== TMA producer =====================
    // the whole warp walks the loop (uniform control flow keeps addresses in uniform registers); one elected lane issues
    const int za1 = p.a_use1 ? i1 : 0, za2 = p.a_use2 ? i2 : 0;
    const int zb1 = p.b_use1 ? i1 : 0, zb2 = p.b_use2 ? i2 : 0;
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
        const int k0 = (kb0 + it) * bke;
        const uint32_t sA = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + Cfg::A_BYTES;
        // MN-major staging: fp32 = groups of 32 (mn) x 32 (k) in the 32 B-atom swizzle (4 KB each); fp16 = groups of 64 (mn) x 64 (k) in the
        // standard 128 B swizzle (8 KB each).  Either way one group is one TMA box and the UMMA descriptor's LBO is the group pitch.
        if (!p.a_mn) {
          tma_load_4d(sA, &p.tmA, full_bar(s), k0, m0, za1, za2);
        } else if (p.f16) {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) tma_load_4d(sA + c * 8192, &p.tmA, full_bar(s), m0 + c * 64, k0, za1, za2);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) tma_load_4d(sA + c * 4096, &p.tmA, full_bar(s), m0 + c * 32, k0, za1, za2);
        }
        if (!p.b_mn) {
          tma_load_4d(sB, &p.tmB, full_bar(s), k0, n0, zb1, zb2);
        } else if (p.f16) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_4d(sB + c * 8192, &p.tmB, full_bar(s), n0 + c * 64, k0, zb1, zb2);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_4d(sB + c * 4096, &p.tmB, full_bar(s), n0 + c * 32, k0, zb1, zb2);
        }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=F32, A=B=TF32, majors, N>>3, M>>4  (cute/arch/mma_sm100_desc.hpp bit layout).  The smem descriptors are
    // split into a constant high word and a low word advanced by one add per k-slice / stage (see umma_desc_lo).
    // fp16 operands (kind::f16, K = 16 per instruction): K-major staging is byte-identical to the fp32 one (128 B rows, 16 B atoms, one K slice =
    // 32 B); MN-major fp16 uses the standard 128 B swizzle with SBO = 1024 B (8 k-rows), LBO = 8192 B (64 mn) and one K slice = 16 rows = 2048 B.
    const bool f16 = p.f16 != 0;
    const uint32_t fmt = f16 ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t mn_hi = f16 ? umma_desc_hi(1024, 2) : umma_desc_hi(512, 1);
    const uint32_t mn_k = f16 ? 128u : 64u, mn_lbo = f16 ? 8192u : 4096u;
    const uint32_t a_hi = p.a_mn ? mn_hi : umma_desc_hi(1024, 2);
    const uint32_t b_hi = p.b_mn ? mn_hi : umma_desc_hi(1024, 2);
    const uint32_t a_k = p.a_mn ? mn_k : 2u, b_k = p.b_mn ? mn_k : 2u;
    const uint32_t a_lo0 = umma_desc_lo(smem_base, p.a_mn ? mn_lbo : 16u);
    const uint32_t b_lo0 = umma_desc_lo(smem_base + Cfg::A_BYTES, p.b_mn ? mn_lbo : 16u);
    int s = 0; uint32_t ph = 0, acc = 0;
    for (int it = 0; it < nkb; ++it) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
#pragma unroll
        for (int k4 = 0; k4 < BK / 8; ++k4) {
          if (f16) umma_f16_lh(tmem_base, a_lo + k4 * a_k, a_hi, b_lo + k4 * b_k, b_hi, idesc, (k4 > 0) ? 1u : acc);
          else umma_tf32_lh(tmem_base, a_lo + k4 * a_k, a_hi, b_lo + k4 * b_k, b_hi, idesc, (k4 > 0) ? 1u : acc);
        }
        umma_commit(empty_bar(s));     // frees this smem stage once the MMAs above have read it
      }
      __syncwarp();
      acc = 1u;
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);      // accumulator complete
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = m0 + q * 32 + lane;
    const int flags = p.flags;
    float alpha = p.alpha, beta = p.beta;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    if (p.alpha_dev2) alpha *= __ldg(p.alpha_dev2);
    if (p.beta_dev) beta *= __ldg(p.beta_dev);
    const float cs_scale = p.colsum_scale * (p.colsum_scale_dev ? __ldg(p.colsum_scale_dev) : 1.0f);
    const bool first_split = (split == 0);

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();

    float* Drow = p.D + (long long)i1 * p.d_bs1 + (long long)i2 * p.d_bs2 + (long long)row * p.ldd;
    const float* Rrow = p.R ? p.R + (long long)i1 * p.r_bs1 + (long long)i2 * p.r_bs2 + (long long)row * p.ldr : nullptr;
    float* Xrow = p.aux ? p.aux + (long long)i1 * p.aux_bs1 + (long long)i2 * p.aux_bs2 + (long long)row * p.ldaux : nullptr;
    const bool vec_ok = ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) &&
                        ((p.d_bs1 & 3) == 0) && ((p.d_bs2 & 3) == 0) &&
                        (!Rrow || (((p.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.R) & 15) == 0) && ((p.r_bs1 & 3) == 0) && ((p.r_bs2 & 3) == 0))) &&
                        (!Xrow || (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && ((p.aux_bs1 & 3) == 0) && ((p.aux_bs2 & 3) == 0)));

#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int col0 = n0 + c0;
      if (col0 >= p.N) break;                      // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (nkb == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (p.colsum && !(flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD | UVC_EPI_RESIDUAL))) {
        // fused bias gradient (fallback kernel: lane = row, so a column sum is a warp reduction per column)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = (row < p.M) ? __uint_as_float(r[j]) * alpha : 0.f;
          if ((flags & UVC_EPI_BIAS) && first_split && col0 + j < p.N) t += (row < p.M) ? __ldg(p.bias + col0 + j) : 0.f;
          t = warp_sum(t);
          if (lane == 0 && col0 + j < p.N) atomicAdd(p.colsum + col0 + j, t * cs_scale);
        }
      }
      if (row < p.M) {
        const bool full = (col0 + 32 <= p.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
        if ((flags & UVC_EPI_BIAS) && first_split) {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (full || col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
        }
        if (flags & UVC_EPI_GELU) {
          float dg[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) gelu_both(v[j], v[j], dg[j]);
          if (Xrow) {                                // aux receives gelu'(pre-activation): all the backward needs
            if (full && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(Xrow + col0 + j) = make_float4(dg[j], dg[j + 1], dg[j + 2], dg[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) Xrow[col0 + j] = dg[j];
            }
          }
        }
        if (flags & UVC_EPI_GELU_BWD) {
          if (full && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 u = *reinterpret_cast<const float4*>(Xrow + col0 + j);
              v[j] *= u.x; v[j + 1] *= u.y; v[j + 2] *= u.z; v[j + 3] *= u.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] *= Xrow[col0 + j];
          }
        }
        if ((flags & UVC_EPI_RESIDUAL) && first_split) {
          if (full && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 u = *reinterpret_cast<const float4*>(Rrow + col0 + j);
              v[j] += beta * u.x; v[j + 1] += beta * u.y; v[j + 2] += beta * u.z; v[j + 3] += beta * u.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += beta * Rrow[col0 + j];
          }
        }
        if (flags & UVC_EPI_ROUND_TF32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
        }
        if (flags & UVC_EPI_ATOMIC) {
          if (full && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add_v4(Drow + col0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) atomicAdd(Drow + col0 + j, v[j]);
          }
        } else {
          if (full && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(Drow + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) Drow[col0 + j] = v[j];
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}


// ====================================================================================================================
// v2: persistent CTA-PAIR kernel (tcgen05 cta_group::2).  A cluster of two CTAs (one per SM of a TPC) owns one
// 256 x BN output tile at a time: CTA r stages rows [128 r, 128 r + 128) of A and rows [BN/2 r, BN/2 r + BN/2) of B, so
// every operand byte is fetched from L2 once per PAIR (half the shared-memory fill traffic of the 128 x 128 kernel,
// which on fp32 operands is what bounds it), the leader CTA's elected thread issues 256 x BN x 8 MMAs, and each CTA's
// TMEM receives its own 128 accumulator rows.  The grid is one pair per two SMs and loops over work units
// (tile x split-K slice, N fastest so concurrently running pairs share the same A rows in L2); the accumulator is
// double-buffered in TMEM (2 x BN columns), so the epilogue of unit i overlaps the main loop of unit i + 1.
//
//   warp 0    : TMA producer (both CTAs; completion bytes land on the LEADER's full barrier)
//   warp 1    : TMEM allocator (both CTAs) + MMA issuer (leader only; commits are multicast to both CTAs' barriers)
//   warps 2-9 : epilogue, two warps per TMEM lane quadrant (each takes half of the BN columns):
//               tcgen05.ld 32 lanes x 32 columns -> XOR-swizzled 4 KB shared staging (transpose) -> every global access of
//               the fused epilogue (bias / GELU (+aux) / GELU' / residual / TF32 rounding / store or red.add) is a fully
//               coalesced 128-bit access: 8 lanes cover 128 contiguous bytes of one output row.
// ====================================================================================================================
// epilogue warps: 8 (two per TMEM lane quadrant, each half of the BN columns); the GELU / GELU' epilogues at BN = 192 take 12 (three per
// quadrant, 64 columns each): they are issue-bound on the transcendental math with only two warps per SM sub-partition
#ifndef UVC_GELU_EW
#define UVC_GELU_EW 12
#endif
__host__ __device__ constexpr int epi_warps2(int BN, int MODE) { return (BN == 192 && (MODE == 1 || MODE == 2)) ? UVC_GELU_EW : 8; }
__host__ __device__ constexpr int threads2(int BN, int MODE) { return 64 + 32 * epi_warps2(BN, MODE); }

template <int BN, int STAGES, int MODE = 0>
struct Gemm2Cfg {
  static constexpr int EW = epi_warps2(BN, MODE);
  static constexpr int WPQ = EW / 4;                  // warps per lane quadrant
  static constexpr int COLS_PER_WARP = BN / WPQ;      // 64, 96 or 128
  static constexpr int HALF_N = BN / 2;
  static constexpr int A_BYTES = 128 * BK * 4;
  static constexpr int B_BYTES = HALF_N * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = EW * 4096;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
};

enum { kEpiPlain = 0, kEpiGelu = 1, kEpiGeluBwd = 2, kEpiBlend = 3 };   // kEpiBlend: plain + residual + the fused block-gate blend (its own instantiation: 32 more live registers)

template <int BN, int STAGES, bool A_MN, bool B_MN, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads2(BN, MODE), 1)
gemm2_tf32_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = Gemm2Cfg<BN, STAGES, MODE>;
  constexpr int HALF_N = Cfg::HALF_N;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  const uint32_t bar0 = smem_u32(&bars[0]);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * Cfg::EW); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(&tmem_slot), Cfg::TMEM_COLS);
  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                                          // everything above is on-chip setup; operands / outputs belong to the stream order

  const int tiles = p.m_tiles * p.n_tiles;
  const int bke = p.f16 ? 2 * BK : BK;                 // elements per 128-byte k-block row: 32 fp32 or 64 fp16
  const int nkb_total = (p.K + bke - 1) / bke;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // The whole warp walks the loop (uniform control flow), one elected lane issues; a k-block costs a wait, an expect_tx and two TMA issues.
    {
      const uint32_t full0_leader = mapa_cluster(full_bar(0), 0);
      int s = 0; uint32_t ph = 0;
      for (int u = pair; u < p.units; u += npairs) {
        const int tile = u % tiles, split = u / tiles;
        const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
        const int m0 = mt * 256 + (int)rank * 128;
        const int n0 = nt * BN + (int)rank * HALF_N;
        const int kb0 = (int)(((long long)nkb_total * split) / p.splits);
        const int kb1 = (int)(((long long)nkb_total * (split + 1)) / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
            const uint32_t fb = full0_leader + 8u * s;
            const int k0 = kb * bke;
            const uint32_t sA = smem_base + s * Cfg::STAGE_BYTES;
            const uint32_t sB = sA + Cfg::A_BYTES;
            if constexpr (!A_MN) {
              tma_load_4d_2sm(sA, &p.tmA, fb, k0, m0, 0, 0);
            } else {
              if (p.a_grp) {
                tma_load_4d_2sm(sA, &p.tmA, fb, 0, k0, m0 >> 5, 0);            // one box = 4 groups of 32 (mn) x 32 (k)
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_4d_2sm(sA + c * 4096, &p.tmA, fb, m0 + c * 32, k0, 0, 0);
              }
            }
            if constexpr (!B_MN) {
              tma_load_4d_2sm(sB, &p.tmB, fb, k0, n0, 0, 0);
            } else {
              if (p.b_grp) {
                tma_load_4d_2sm(sB, &p.tmB, fb, 0, k0, n0 >> 5, 0);
              } else {
#pragma unroll
                for (int c = 0; c < HALF_N / 32; ++c) tma_load_4d_2sm(sB + c * 4096, &p.tmB, fb, n0 + c * 32, k0, 0, 0);
              }
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // A single thread feeds the tensor cores of both SMs; one 256 x BN x 8 MMA takes BN/2 cycles, so the loop body must
    // stay far below 2*BN cycles per k-block: descriptors are split into a constant high word and a low word advanced by
    // one add (stage: STAGE_BYTES/16, k-slice: 2 (K-major, 32 B) or 64 (MN-major, 8 rows of 128 B)).
    if (rank == 0) {
      // the whole warp walks the loop (warp-uniform control flow keeps addresses in uniform registers); one elected lane issues
      // instruction descriptor: accumulator F32 (bits 4-5 = 1), operand formats at bits 7-9 / 10-12 (F16 = 0, TF32 = 2), majors, N >> 3, M >> 4
      const uint32_t fmt = p.f16 ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const bool f16 = p.f16 != 0;
      constexpr uint32_t a_hi = A_MN ? ((512u >> 4) | (1u << 14) | (1u << 29)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
      constexpr uint32_t b_hi = B_MN ? ((512u >> 4) | (1u << 14) | (1u << 29)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
      constexpr uint32_t a_k = A_MN ? 64u : 2u, b_k = B_MN ? 64u : 2u;
      const uint32_t a_lo0 = umma_desc_lo(smem_base, A_MN ? 4096u : 16u);
      const uint32_t b_lo0 = umma_desc_lo(smem_base + Cfg::A_BYTES, B_MN ? 4096u : 16u);
      int s = 0; uint32_t ph = 0, ac = 0;
      for (int u = pair; u < p.units; u += npairs, ++ac) {
        const int split = u / tiles;
        const int kb0 = (int)(((long long)nkb_total * split) / p.splits);
        const int kb1 = (int)(((long long)nkb_total * (split + 1)) / p.splits);
        const uint32_t a = ac & 1u, aph = (ac >> 1) & 1u;
        mbar_wait(tempty_bar(a), aph ^ 1u);            // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)s * (Cfg::STAGE_BYTES >> 4);
#pragma unroll
            for (int k4 = 0; k4 < BK / 8; ++k4) {       // four K-slices of 32 bytes per stage row: K = 8 (tf32) or 16 (fp16) each
              if (f16) umma_f16_2sm_lh(d_tmem, a_lo + k4 * a_k, a_hi, b_lo + k4 * b_k, b_hi, idesc, (k4 > 0) ? 1u : acc);
              else umma_tf32_2sm_lh(d_tmem, a_lo + k4 * a_k, a_hi, b_lo + k4 * b_k, b_hi, idesc, (k4 > 0) ? 1u : acc);
            }
            umma_commit_2sm(empty_bar(s), 3);          // frees this smem stage in BOTH CTAs
          }
          __syncwarp();
          acc = 1u;
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit_2sm(tfull_bar(a), 3);   // accumulator stage complete: wake both CTAs' epilogues
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs) =====================
    // MODE is a template parameter so the GELU / GELU' code only exists in the instantiations that need it: with every mode
    // inlined the loop body was 2.5 K instructions (40 KB) and the eight epilogue warps were instruction-fetch bound.
    const int ew = warp - 2;
    const int q = warp & 3;                            // TMEM lane quadrant this warp may access
    const int part = ew >> 2;                          // which slice of the BN columns (Cfg::COLS_PER_WARP each)
    const uint32_t stg = stg_base + ew * 4096;
    const int flags = p.flags;
    float alpha = p.alpha, beta = p.beta;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    if (p.alpha_dev2) alpha *= __ldg(p.alpha_dev2);
    if (p.beta_dev) beta *= __ldg(p.beta_dev);
    const float cs_scale = p.colsum_scale * (p.colsum_scale_dev ? __ldg(p.colsum_scale_dev) : 1.0f);
    const int rl = lane >> 3, cc = lane & 7;
    const uint32_t st_row = stg + lane * 128;                                        // staging: this lane's TMEM row
    const uint32_t ld_even = stg + rl * 128 + (((uint32_t)cc ^ (uint32_t)rl) << 4);  // rows i*4 + rl, i even: (row & 7) = rl
    const uint32_t ld_odd = stg + rl * 128 + (((uint32_t)cc ^ (uint32_t)(rl + 4)) << 4);
    const uint32_t tempty_leader = mapa_cluster(tempty_bar(0), 0);
    const bool do_round = (flags & UVC_EPI_ROUND_TF32) != 0, do_atomic = (flags & UVC_EPI_ATOMIC) != 0;
    const bool do_blend = (flags & UVC_EPI_BLEND) != 0;
    const float bl_d0 = do_blend ? __ldg(p.blend_dev) : 0.f, bl_d1 = do_blend ? __ldg(p.blend_dev + 1) : 1.f;
    uint32_t ac = 0;
    for (int u = pair; u < p.units; u += npairs, ++ac) {
      const int tile = u % tiles, split = u / tiles;
      const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      const int row_base = mt * 256 + (int)rank * 128 + q * 32;
      const bool first_split = (split == 0);
      const bool do_bias = (flags & UVC_EPI_BIAS) && first_split, do_res = (flags & UVC_EPI_RESIDUAL) && first_split;
      const uint32_t a = ac & 1u, aph = (ac >> 1) & 1u;
      mbar_wait(tfull_bar(a), aph);
      tc_fence_after();
      if (row_base < p.M) {
        const int rows_left = p.M - row_base - rl;     // row i*4 + rl is valid iff i*4 < rows_left
        const long long roff = (long long)(row_base + rl);
#pragma unroll 1
        for (int c = 0; c < Cfg::COLS_PER_WARP / 32; ++c) {
          const int col_t = part * Cfg::COLS_PER_WARP + c * 32;
          const int col0 = nt * BN + col_t;
          if (col0 >= p.N) break;                      // warp-uniform
          const int gcol = col0 + cc * 4;
          const bool colok = gcol < p.N;               // N % 4 == 0: a float4 is entirely in or out
          // Per-row pointers advance by a constant step; every kernel-uniform option (which outputs exist, aux width, residual, rounding, ...) is
          // tested once per chunk, outside the unrolled row loops, and chunks that lie entirely inside the matrix (all but the last row / column
          // tile) run without per-access predicates.  The epilogue warps are issue-bound: before this restructuring the loop spent ~60
          // instructions per output element, two thirds of them address arithmetic, null checks and bounds predicates.
          float* const dptr = p.D + roff * p.ldd + gcol; const long long dstep = 4 * p.ldd;
          __half* const d16ptr = reinterpret_cast<__half*>(p.D16) + roff * p.ldd16 + gcol; const long long d16step = 4 * p.ldd16;
          const float* const rptr = p.R + roff * p.ldr + gcol; const long long rstep = 4 * p.ldr;
          const float* const r2ptr = p.R2 + roff * p.ldr2 + gcol; const long long r2step = 4 * p.ldr2;
          float* const d2ptr = p.D2 + roff * p.ldd2 + gcol; const long long d2step = 4 * p.ldd2;
          float* const xptr = p.aux + roff * p.ldaux + gcol;
          __half* const xptr16 = reinterpret_cast<__half*>(p.aux) + roff * p.ldaux + gcol;     // UVC_EPI_AUX_F16 view of the same argument
          const long long xstep = 4 * p.ldaux;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + (uint32_t)col_t;
          auto chunk = [&](auto interior_c) {
            constexpr bool INTR = decltype(interior_c)::value;
            auto ok = [&](int i) { return INTR || (colok && i * 4 < rows_left); };
            // (a) operand loads that do not depend on the accumulator go first: their latency overlaps the TMEM read and the transpose
            float4 rr[8];
            if (MODE == kEpiGeluBwd) {                 // gelu'(pre-activation) factors
              if (p.aux16) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  uint2 u = make_uint2(0u, 0u);
                  if (ok(i)) u = *reinterpret_cast<const uint2*>(xptr16 + i * xstep);
                  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                  rr[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) rr[i] = ok(i) ? *reinterpret_cast<const float4*>(xptr + i * xstep) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            } else if (do_res) {
#pragma unroll
              for (int i = 0; i < 8; ++i) rr[i] = ok(i) ? *reinterpret_cast<const float4*>(rptr + i * rstep) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 rr2[MODE == kEpiBlend ? 8 : 1];
            if (MODE == kEpiBlend) {
#pragma unroll
              for (int i = 0; i < 8; ++i) rr2[i] = ok(i) ? *reinterpret_cast<const float4*>(r2ptr + i * r2step) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (do_bias && colok) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
            // (b) accumulator chunk: TMEM -> registers -> XOR-swizzled staging tile (lane = row) -> registers in the coalesced layout
            {
              uint32_t r[32];
              tmem_ld_32x32(taddr, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t addr = st_row + (((uint32_t)j ^ ((uint32_t)lane & 7u)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
              }
            }
            __syncwarp();
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t addr = ((i & 1) ? ld_odd : ld_even) + i * 512;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(addr) : "memory");
              v[i].x = fmaf(v[i].x, alpha, b4.x); v[i].y = fmaf(v[i].y, alpha, b4.y); v[i].z = fmaf(v[i].z, alpha, b4.z); v[i].w = fmaf(v[i].w, alpha, b4.w);
            }
            // (c) the fused epilogue math
            if (MODE == kEpiGelu) {
              if (p.aux) {                             // aux receives gelu'(pre-activation): the same exponential gives both, and the backward epilogue becomes a multiply
                if (p.aux16) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float4 dg;
                    gelu_pair<true>(v[i].x, v[i].y, v[i].x, v[i].y, dg.x, dg.y); gelu_pair<true>(v[i].z, v[i].w, v[i].z, v[i].w, dg.z, dg.w);
                    if (ok(i)) *reinterpret_cast<uint2*>(xptr16 + i * xstep) = pack_half4(dg.x, dg.y, dg.z, dg.w);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float4 dg;
                    gelu_pair<true>(v[i].x, v[i].y, v[i].x, v[i].y, dg.x, dg.y); gelu_pair<true>(v[i].z, v[i].w, v[i].z, v[i].w, dg.z, dg.w);
                    if (ok(i)) *reinterpret_cast<float4*>(xptr + i * xstep) = dg;
                  }
                }
              } else {                                 // inference / teacher forward: no derivative
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float u0, u1;
                  gelu_pair<false>(v[i].x, v[i].y, v[i].x, v[i].y, u0, u1); gelu_pair<false>(v[i].z, v[i].w, v[i].z, v[i].w, u0, u1);
                }
              }
            }
            if (MODE == kEpiGeluBwd) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { v[i].x *= rr[i].x; v[i].y *= rr[i].y; v[i].z *= rr[i].z; v[i].w *= rr[i].w; }
            } else if (do_res) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[i].x = fmaf(beta, rr[i].x, v[i].x); v[i].y = fmaf(beta, rr[i].y, v[i].y); v[i].z = fmaf(beta, rr[i].z, v[i].z); v[i].w = fmaf(beta, rr[i].w, v[i].w);
              }
            }
            if (MODE == kEpiBlend) {                   // block gate (models/model_distilled.py:493): x_out = d1 t + d0 x, t = the value formed so far
              if (p.D2) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (ok(i)) *reinterpret_cast<float4*>(d2ptr + i * d2step) = v[i];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[i].x = fmaf(bl_d0, rr2[i].x, bl_d1 * v[i].x); v[i].y = fmaf(bl_d0, rr2[i].y, bl_d1 * v[i].y);
                v[i].z = fmaf(bl_d0, rr2[i].z, bl_d1 * v[i].z); v[i].w = fmaf(bl_d0, rr2[i].w, bl_d1 * v[i].w);
              }
            }
            if (p.colsum) {                            // fused bias gradient: column sums of this warp's 32 rows (before any rounding)
              float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 8; ++i) if (ok(i)) { cs.x += v[i].x; cs.y += v[i].y; cs.z += v[i].z; cs.w += v[i].w; }
              // lanes with the same cc hold the same 4 columns: fold the 4 row phases, one red per 4 columns
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
              if (rl == 0 && colok) red_add_v4(p.colsum + gcol, cs.x * cs_scale, cs.y * cs_scale, cs.z * cs_scale, cs.w * cs_scale);
            }
            if (do_round) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { v[i].x = round_tf32(v[i].x); v[i].y = round_tf32(v[i].y); v[i].z = round_tf32(v[i].z); v[i].w = round_tf32(v[i].w); }
            }
            // (d) stores
            if (p.D16) {
#pragma unroll
              for (int i = 0; i < 8; ++i) if (ok(i)) *reinterpret_cast<uint2*>(d16ptr + i * d16step) = pack_half4(v[i].x, v[i].y, v[i].z, v[i].w);
            }
            if (p.D) {
              if (do_atomic) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (ok(i)) red_add_v4(dptr + i * dstep, v[i].x, v[i].y, v[i].z, v[i].w);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (ok(i)) *reinterpret_cast<float4*>(dptr + i * dstep) = v[i];
              }
            }
          };
          if (row_base + 32 <= p.M && col0 + 32 <= p.N) chunk(std::true_type{}); else chunk(std::false_type{});
          __syncwarp();                                // staging tile is rewritten by the next chunk
        }
      }
      // this warp's TMEM reads of the stage are complete (tcgen05.wait::ld above): hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + a * 8);
    }
  }

  tc_fence_before();
  cluster_sync_all();                                  // nobody exits (or frees TMEM) while the pair still works
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(sym);
  return fn;
}

// generic 4-D fp32 tiled tensor map (used by the fused attention kernels): 128 B swizzle, 16 B atoms (K-major operands) or 32 B atoms (MN-major)
int encode_tmap_4d(CUtensorMap* tm, const float* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                   const unsigned int box[4], bool atom32, const char* name) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  UVC_REQUIRE(enc != nullptr, UVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UVC_REQUIRE(r == CUDA_SUCCESS, UVC_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", name, (int)r);
  return UVC_OK;
}

// generic 4-D fp16 tiled tensor map, 128 B swizzle (16 B atoms: one staged copy serves K-major and MN-major 16-bit UMMA operands alike)
int encode_tmap_4d_f16(CUtensorMap* tm, const void* base, const unsigned long long dims[4], const unsigned long long strides_bytes[3],
                       const unsigned int box[4], const char* name) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  UVC_REQUIRE(enc != nullptr, UVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UVC_REQUIRE(r == CUDA_SUCCESS, UVC_ERR_CUDA, "cuTensorMapEncodeTiled(%s, fp16) failed with CUresult %d", name, (int)r);
  return UVC_OK;
}

// rows_mn: logical M or N extent; box_rows: tile rows for the K-major box
static int make_tmap(CUtensorMap* tm, const uvc_operand& op, int rows_mn, int K, int nb1, int nb2, int box_rows, const char* name) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  UVC_REQUIRE(enc != nullptr, UVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  UVC_REQUIRE(op.ptr != nullptr, UVC_ERR_BAD_ARG, "gemm operand %s is NULL", name);
  UVC_REQUIRE((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, UVC_ERR_BAD_SHAPE, "gemm operand %s not 16B aligned", name);
  UVC_REQUIRE(op.ld > 0 && (op.ld & 3) == 0, UVC_ERR_BAD_SHAPE, "gemm operand %s: ld=%lld must be a positive multiple of 4", name, (long long)op.ld);
  UVC_REQUIRE((op.bs1 & 3) == 0 && (op.bs2 & 3) == 0 && op.bs1 >= 0 && op.bs2 >= 0, UVC_ERR_BAD_SHAPE, "gemm operand %s: batch strides must be non-negative multiples of 4", name);
  const cuuint64_t cols = op.mn_major ? (cuuint64_t)rows_mn : (cuuint64_t)K;
  const cuuint64_t rows = op.mn_major ? (cuuint64_t)K : (cuuint64_t)rows_mn;
  UVC_REQUIRE((long long)cols <= op.ld, UVC_ERR_BAD_SHAPE, "gemm operand %s: ld=%lld smaller than its %llu columns", name, (long long)op.ld, (unsigned long long)cols);
  const cuuint64_t n1 = (op.bs1 != 0) ? (cuuint64_t)nb1 : 1, n2 = (op.bs2 != 0) ? (cuuint64_t)nb2 : 1;
  cuuint64_t dims[4] = {cols, rows, n1, n2};
  const cuuint64_t row_bytes = (cuuint64_t)op.ld * 4;
  cuuint64_t strides[3] = {row_bytes, op.bs1 ? (cuuint64_t)op.bs1 * 4 : row_bytes * rows, op.bs2 ? (cuuint64_t)op.bs2 * 4 : row_bytes * rows};
  cuuint32_t box[4] = {32, op.mn_major ? 32u : (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   op.mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UVC_REQUIRE(r == CUDA_SUCCESS, UVC_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu x %llu x %llu x %llu, ld %lld)",
              name, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3], (long long)op.ld);
  return UVC_OK;
}

// fp16 operand.  K-major [rows][K]: boxes of 64 elements (128 B) x box_rows, 128 B swizzle -- byte-for-byte the staging layout of the fp32
// path.  MN-major [K rows][MN cols]: boxes of 64 (mn) x 64 (k), 128 B swizzle = one 8 KB group of the canonical MN-major UMMA layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute/atom/mma_traits_sm100.hpp), SBO = 1024 B, LBO = 8192 B.
static int make_tmap_f16(CUtensorMap* tm, const uvc_operand& op, int rows_mn, int K, int nb1, int nb2, int box_rows, const char* name) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  UVC_REQUIRE(enc != nullptr, UVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  UVC_REQUIRE(op.ptr != nullptr, UVC_ERR_BAD_ARG, "gemm (fp16 operands): %s is NULL", name);
  UVC_REQUIRE((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0 && op.ld > 0 && (op.ld & 7) == 0, UVC_ERR_BAD_SHAPE,
              "gemm (fp16 operands): %s needs a 16 B-aligned base and ld=%lld a positive multiple of 8", name, (long long)op.ld);
  UVC_REQUIRE((op.bs1 & 7) == 0 && (op.bs2 & 7) == 0 && op.bs1 >= 0 && op.bs2 >= 0, UVC_ERR_BAD_SHAPE, "gemm (fp16 operands): %s batch strides must be non-negative multiples of 8", name);
  const cuuint64_t cols = op.mn_major ? (cuuint64_t)rows_mn : (cuuint64_t)K;
  const cuuint64_t rows = op.mn_major ? (cuuint64_t)K : (cuuint64_t)rows_mn;
  UVC_REQUIRE((long long)cols <= op.ld, UVC_ERR_BAD_SHAPE, "gemm (fp16 operands): %s ld=%lld smaller than its %llu columns", name, (long long)op.ld, (unsigned long long)cols);
  const cuuint64_t n1 = (op.bs1 != 0) ? (cuuint64_t)nb1 : 1, n2 = (op.bs2 != 0) ? (cuuint64_t)nb2 : 1;
  cuuint64_t dims[4] = {cols, rows, n1, n2};
  const cuuint64_t row_bytes = (cuuint64_t)op.ld * 2;
  cuuint64_t strides[3] = {row_bytes, op.bs1 ? (cuuint64_t)op.bs1 * 2 : row_bytes * rows, op.bs2 ? (cuuint64_t)op.bs2 * 2 : row_bytes * rows};
  cuuint32_t box[4] = {64, op.mn_major ? 64u : (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<float*>(op.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UVC_REQUIRE(r == CUDA_SUCCESS, UVC_ERR_CUDA, "cuTensorMapEncodeTiled(%s, fp16) failed with CUresult %d", name, (int)r);
  return UVC_OK;
}

// MN-major operand [K rows][MN cols] viewed as (32 mn, K, MN/32 groups): ONE TMA box of 32 x 32 x `groups` lands `groups` consecutive
// 4 KB swizzle atoms [g][k][mn] in shared memory -- the layout the MN-major UMMA descriptor walks (LBO = 4096 B between groups).
// Returns 1 when the view is not expressible (MN % 32 != 0, or the driver rejects the overlapping strides): caller falls back to 32 x 32 boxes.
static int make_tmap_grouped(CUtensorMap* tm, const uvc_operand& op, int rows_mn, int K, int groups) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc || !op.mn_major || (rows_mn & 31) || op.ld <= 0 || (op.ld & 3) || (reinterpret_cast<uintptr_t>(op.ptr) & 15)) return 1;
  cuuint64_t dims[4] = {32, (cuuint64_t)K, (cuuint64_t)(rows_mn / 32), 1};
  cuuint64_t strides[3] = {(cuuint64_t)op.ld * 4, 128, (cuuint64_t)op.ld * 4 * (cuuint64_t)K};
  cuuint32_t box[4] = {32, 32, (cuuint32_t)groups, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(op.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

template <int BN, int STAGES>
static int launch(const GemmKParams& kp, dim3 grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN, STAGES>;
  static unsigned long long attr_devs = 0;       // devices this kernel's attribute has been set on
  if (first_on_device(&attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(gemm smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
  }
  launch_pdl(gemm_tf32_kernel<BN, STAGES>, grid, dim3(kThreads), Cfg::SMEM_BYTES, st, kp);
  return check_launch("gemm_tf32_kernel");
}


template <int BN, int STAGES, bool A_MN, bool B_MN, int MODE>
static int launch2k(const GemmKParams& kp, int pairs, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN, STAGES, MODE>;
  static unsigned long long attr_devs = 0;       // devices this kernel's attribute has been set on
  if (first_on_device(&attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tf32_kernel<BN, STAGES, A_MN, B_MN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(gemm2 smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
  }
  launch_pdl(gemm2_tf32_kernel<BN, STAGES, A_MN, B_MN, MODE>, dim3(2 * pairs), dim3(threads2(BN, MODE)), Cfg::SMEM_BYTES, st, kp);
  return check_launch("gemm2_tf32_kernel");
}
template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch2m(const GemmKParams& kp, int pairs, cudaStream_t st) {
  if (kp.flags & UVC_EPI_GELU) return launch2k<BN, STAGES, A_MN, B_MN, kEpiGelu>(kp, pairs, st);
  if (kp.flags & UVC_EPI_GELU_BWD) return launch2k<BN, STAGES, A_MN, B_MN, kEpiGeluBwd>(kp, pairs, st);
  if constexpr (!A_MN && !B_MN) { if (kp.flags & UVC_EPI_BLEND) return launch2k<BN, STAGES, A_MN, B_MN, kEpiBlend>(kp, pairs, st); }
  return launch2k<BN, STAGES, A_MN, B_MN, kEpiPlain>(kp, pairs, st);
}
template <int BN, int STAGES>
static int launch2(const GemmKParams& kp, int pairs, cudaStream_t st) {
  if (kp.a_mn) return kp.b_mn ? launch2m<BN, STAGES, true, true>(kp, pairs, st) : launch2m<BN, STAGES, true, false>(kp, pairs, st);
  return kp.b_mn ? launch2m<BN, STAGES, false, true>(kp, pairs, st) : launch2m<BN, STAGES, false, false>(kp, pairs, st);
}

static int sm_pairs() {
  static int pairs[64] = {0};            // per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 74;
  int& p = pairs[dev & 63];
  if (p == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2) sms = 148;
    p = sms / 2;
  }
  return p;
}

// 0 = always the 128 x 128 kernel, 1 = pick per shape (default), 2 = CTA-pair kernel whenever it is legal
static int gemm_v2_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("UVC_GEMM_V2");
    mode = e ? atoi(e) : 1;
  }
  return mode;
}
static int gemm_v2_force_bn() {
  static int bn = -1;
  if (bn < 0) {
    const char* e = getenv("UVC_GEMM_V2_BN");
    bn = e ? atoi(e) : 0;
  }
  return bn;
}

// the CTA-pair kernel needs unbatched operands and 16 B-aligned rows everywhere its epilogue touches
static bool v2_legal(const uvc_gemm_args& a) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (a.nb1 != 1 || a.nb2 != 1 || a.K < 1) return false;
  if ((a.flags & UVC_EPI_GELU) && (a.flags & UVC_EPI_GELU_BWD)) return false;
  if ((a.flags & UVC_EPI_GELU_BWD) && (a.flags & UVC_EPI_RESIDUAL)) return false;
  if (a.N & 3) return false;
  if (a.D && ((a.ldd & 3) || !al16(a.D))) return false;
  if (a.D16 && ((a.ldd16 & 3) || (reinterpret_cast<uintptr_t>(a.D16) & 7))) return false;
  if (!a.D && !a.D16) return false;
  if ((a.flags & UVC_EPI_BIAS) && !al16(a.bias)) return false;
  if ((a.flags & UVC_EPI_COLSUM) && !al16(a.colsum)) return false;
  if ((a.flags & UVC_EPI_RESIDUAL) && ((a.ldr & 3) || !al16(a.R))) return false;
  if (a.flags & UVC_EPI_BLEND) {
    if (a.A.mn_major || a.B.mn_major || !a.R2 || !a.blend_dev || (a.ldr2 & 3) || !al16(a.R2) || (a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD | UVC_EPI_ATOMIC)) || a.splits != 1) return false;
    if (a.D2 && ((a.ldd2 & 3) || !al16(a.D2))) return false;
  }
  if ((a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD)) && a.aux) {
    if (a.flags & UVC_EPI_AUX_F16) { if ((a.ldaux & 3) || (reinterpret_cast<uintptr_t>(a.aux) & 7)) return false; }
    else if ((a.ldaux & 3) || !al16(a.aux)) return false;
  }
  return true;
}

// tile width minimising (waves over the SM pairs) x (tile cost ~ BN); ties go to the wider tile (fewer operand re-reads)
static int v2_pick_bn(int M, int N, int splits, int pairs) {
  const int mt = (M + 255) / 256;
  int best = 0; long long best_cost = 0;
  const int cand[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    const long long units = (long long)mt * ((N + bn - 1) / bn) * splits;
    const long long cost = ((units + pairs - 1) / pairs) * (bn + 24);     // + fixed per-unit overhead (pipeline fill, barriers)
    if (!best || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

int gemm_tf32(const uvc_gemm_args& a, cudaStream_t st) {
  UVC_REQUIRE(a.M > 0 && a.N > 0 && a.K >= 0, UVC_ERR_BAD_SHAPE, "gemm: bad M,N,K = %d,%d,%d", a.M, a.N, a.K);
  UVC_REQUIRE(a.nb1 >= 1 && a.nb2 >= 1 && a.splits >= 1, UVC_ERR_BAD_ARG, "gemm: nb1, nb2, splits must be >= 1");
  UVC_REQUIRE(a.D != nullptr || a.D16 != nullptr, UVC_ERR_BAD_ARG, "gemm: D and D16 are both NULL");
  UVC_REQUIRE(a.splits == 1 || (a.flags & UVC_EPI_ATOMIC), UVC_ERR_BAD_ARG, "gemm: splits > 1 requires UVC_EPI_ATOMIC");
  UVC_REQUIRE(!(a.flags & UVC_EPI_ATOMIC) || !(a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD)), UVC_ERR_BAD_ARG, "gemm: GELU epilogues cannot be combined with split-K accumulation");
  UVC_REQUIRE(!(a.flags & UVC_EPI_ATOMIC) || !(a.flags & UVC_EPI_ROUND_TF32), UVC_ERR_BAD_ARG, "gemm: UVC_EPI_ROUND_TF32 cannot be combined with atomic accumulation");
  UVC_REQUIRE(!(a.flags & UVC_EPI_BIAS) || a.bias, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_BIAS without bias");
  UVC_REQUIRE(!(a.flags & UVC_EPI_RESIDUAL) || a.R, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_RESIDUAL without R");
  UVC_REQUIRE(!(a.flags & UVC_EPI_GELU_BWD) || a.aux, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_GELU_BWD without aux");
  UVC_REQUIRE(!(a.flags & UVC_EPI_COLSUM) || a.colsum, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_COLSUM without colsum");
  const bool f16 = (a.flags & UVC_GEMM_F16) != 0;
  const int nkb = f16 ? (a.K + 2 * BK - 1) / (2 * BK) : (a.K + BK - 1) / BK;
  int splits = a.splits;
  const bool colsum_simple = !(a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD | UVC_EPI_RESIDUAL));
  if (splits > nkb) splits = nkb > 0 ? nkb : 1;

  GemmKParams kp;
  constexpr int BN = 128;
  const int mode = gemm_v2_mode();
  const int pairs = sm_pairs();
  int bn2 = 0;
  const bool aux16 = (a.flags & UVC_EPI_AUX_F16) && (a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD)) && a.aux;
  // fp16 operands: K-major unbatched problems run on the CTA-pair kernel; MN-major operands (the weight gradients: both operands are read
  // transposed), split-K and batched problems on the 128 x 128 kernel.
  const bool f16_v1 = f16 && (a.A.mn_major || a.B.mn_major || a.splits > 1 || a.nb1 != 1 || a.nb2 != 1);
  const bool need_v2 = ((a.flags & UVC_EPI_COLSUM) && !colsum_simple) || (f16 && !f16_v1) || a.D16 || !a.D || aux16 || (a.flags & UVC_EPI_BLEND);   // features only the CTA-pair kernel has
  UVC_REQUIRE(!(f16_v1 && need_v2), UVC_ERR_BAD_ARG, "gemm: fp16 MN-major / split-K / batched operands cannot be combined with D16, fp16 aux or COLSUM under a GELU / residual epilogue");
  UVC_REQUIRE(!need_v2 || v2_legal(a), UVC_ERR_BAD_ARG, "gemm: fp16 operands / D16 / UVC_EPI_COLSUM with GELU or residual epilogues need unbatched, 16 B-aligned operands and N % 4 == 0");
  // Split-K weight gradients (few output tiles, K = all tokens) stay on the 128 x 128 kernel: its tiles fit the C-multiple weight shapes without
  // padding and tiles x splits fills one wave of 2 CTAs per SM; measured on the four DeiT-Small shapes it is 0-30 % faster than the CTA-pair
  // kernel there (qkv_w 47 vs 69 us, fc1_w 55 vs 61 us; tests/bringup/wgrad_perf.py).
  const bool splitk_wgrad = a.splits > 1 && (a.flags & UVC_EPI_ATOMIC) && mode != 2 && !need_v2;
  if ((mode > 0 || need_v2) && v2_legal(a) && !splitk_wgrad && !f16_v1 && (mode == 2 || need_v2 || (a.M >= 512 && a.N >= 128))) {
    // split-K (caller allows it by passing splits > 1 with UVC_EPI_ATOMIC): the persistent kernel wants ~2 units per SM pair
    if (splits > 1 && !getenv("UVC_GEMM_V2_KEEP_SPLITS")) {
      const int bn0 = gemm_v2_force_bn() ? gemm_v2_force_bn() : 128;
      const int tiles2 = ((a.M + 255) / 256) * ((a.N + bn0 - 1) / bn0);
      int sp = (2 * pairs) / tiles2;               // floor: one unit over two full waves costs a third wave
      if (sp > nkb / 8) sp = nkb / 8;              // keep >= 8 k-blocks per slice
      splits = sp < 1 ? 1 : sp;
    }
    bn2 = gemm_v2_force_bn() ? gemm_v2_force_bn() : (splits > 1 ? 128 : v2_pick_bn(a.M, a.N, splits, pairs));
    UVC_REQUIRE(bn2 == 128 || bn2 == 192 || bn2 == 256, UVC_ERR_BAD_ARG, "gemm: UVC_GEMM_V2_BN must be 128, 192 or 256");
  }
  kp.a_grp = kp.b_grp = 0;
  int rc = UVC_OK;
  kp.f16 = f16 ? 1 : 0;
  kp.aux16 = aux16 ? 1 : 0;
  kp.D16 = a.D16; kp.ldd16 = a.ldd16;
  if (f16) {
    if ((rc = make_tmap_f16(&kp.tmA, a.A, a.M, a.K, a.nb1, a.nb2, BM, "A"))) return rc;
    if ((rc = make_tmap_f16(&kp.tmB, a.B, a.N, a.K, a.nb1, a.nb2, bn2 ? bn2 / 2 : BN, "B"))) return rc;
  } else {
    if (bn2 && a.A.mn_major && make_tmap_grouped(&kp.tmA, a.A, a.M, a.K, 4) == 0) kp.a_grp = 1;
    else if ((rc = make_tmap(&kp.tmA, a.A, a.M, a.K, a.nb1, a.nb2, BM, "A"))) return rc;
    if (bn2 && a.B.mn_major && make_tmap_grouped(&kp.tmB, a.B, a.N, a.K, bn2 / 64) == 0) kp.b_grp = 1;
    else if ((rc = make_tmap(&kp.tmB, a.B, a.N, a.K, a.nb1, a.nb2, bn2 ? bn2 / 2 : BN, "B"))) return rc;
  }
  kp.D = a.D; kp.ldd = a.ldd; kp.d_bs1 = a.d_bs1; kp.d_bs2 = a.d_bs2;
  kp.bias = a.bias;
  kp.R = (a.flags & UVC_EPI_RESIDUAL) ? a.R : nullptr; kp.ldr = a.ldr; kp.r_bs1 = a.r_bs1; kp.r_bs2 = a.r_bs2;
  const bool blend = (a.flags & UVC_EPI_BLEND) != 0;
  kp.R2 = blend ? a.R2 : nullptr; kp.ldr2 = a.ldr2; kp.D2 = blend ? a.D2 : nullptr; kp.ldd2 = a.ldd2; kp.blend_dev = blend ? a.blend_dev : nullptr;
  kp.aux = (a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD)) ? a.aux : nullptr; kp.ldaux = a.ldaux; kp.aux_bs1 = a.aux_bs1; kp.aux_bs2 = a.aux_bs2;
  kp.alpha_dev = a.alpha_dev; kp.beta_dev = a.beta_dev; kp.alpha_dev2 = a.alpha_dev2; kp.colsum_scale_dev = a.colsum_scale_dev;
  kp.colsum = (a.flags & UVC_EPI_COLSUM) ? a.colsum : nullptr;
  kp.alpha = a.alpha; kp.beta = a.beta; kp.colsum_scale = (a.colsum_scale != 0.0f) ? a.colsum_scale : 1.0f;
  kp.M = a.M; kp.N = a.N; kp.K = a.K; kp.nb1 = a.nb1; kp.nb2 = a.nb2; kp.splits = splits; kp.flags = a.flags;
  kp.a_mn = a.A.mn_major ? 1 : 0; kp.b_mn = a.B.mn_major ? 1 : 0;
  kp.a_use1 = a.A.bs1 != 0; kp.a_use2 = a.A.bs2 != 0; kp.b_use1 = a.B.bs1 != 0; kp.b_use2 = a.B.bs2 != 0;

  if (bn2) {
    kp.m_tiles = (a.M + 255) / 256; kp.n_tiles = (a.N + bn2 - 1) / bn2;
    const long long units = (long long)kp.m_tiles * kp.n_tiles * splits;
    UVC_REQUIRE(units < (1ll << 30), UVC_ERR_BAD_SHAPE, "gemm: too many work units (%lld)", units);
    kp.units = (int)units;
    const int np = units < pairs ? (int)units : pairs;
    const bool prof2 = prof_enabled();
    if (prof2) prof_begin(st, 2.0 * a.M * a.N * (double)a.K, f16 ? 3 : 2);
    if (bn2 == 256) rc = launch2<256, 6>(kp, np, st);
    else if (bn2 == 192) rc = launch2<192, 6>(kp, np, st);
    else rc = launch2<128, 8>(kp, np, st);
    if (prof2) prof_end(st);
    return rc;
  }
  kp.m_tiles = kp.n_tiles = kp.units = 0;
  const long long gz = (long long)a.nb1 * a.nb2 * splits;
  const long long gy = (a.M + BM - 1) / BM;
  UVC_REQUIRE(gz <= 65535 && gy <= 65535, UVC_ERR_BAD_SHAPE, "gemm: grid too large (m tiles %lld, batch*splits %lld)", gy, gz);
  dim3 grid((a.N + BN - 1) / BN, (unsigned)gy, (unsigned)gz);
  const bool prof = prof_enabled();
  if (prof) prof_begin(st, 2.0 * a.M * a.N * (double)a.K * a.nb1 * a.nb2, f16 ? 4 : 1);
  rc = launch<BN, 3>(kp, grid, st);
  if (prof) prof_end(st);
  return rc;
}

}  // namespace uvc

extern "C" int uvc_gemm_tf32(const uvc_gemm_args* args, void* stream) {
  if (!args) { uvc::set_error("uvc_gemm_tf32: args is NULL"); return UVC_ERR_BAD_ARG; }
  return uvc::gemm_tf32(*args, static_cast<cudaStream_t>(stream));
}

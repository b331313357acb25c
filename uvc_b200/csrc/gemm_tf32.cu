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

namespace uvc {

constexpr int BM = 128;
constexpr int BK = 32;                 // fp32 elements per k-block = one 128 B swizzle row
constexpr int kThreads = 192;

struct alignas(64) GemmKParams {
  CUtensorMap tmA, tmB;
  float* D; long long ldd, d_bs1, d_bs2;
  const float* bias;
  const float* R; long long ldr, r_bs1, r_bs2;
  float* aux; long long ldaux, aux_bs1, aux_bs2;
  const float* alpha_dev; const float* beta_dev;
  float alpha, beta;
  int M, N, K, nb1, nb2, splits, flags;
  int a_mn, b_mn;
  int a_use1, a_use2, b_use1, b_use2;   // operand varies with batch index i1 / i2 (else coordinate 0)
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

  const int nkb_total = (p.K + BK - 1) / BK;
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int za1 = p.a_use1 ? i1 : 0, za2 = p.a_use2 ? i2 : 0;
      const int zb1 = p.b_use1 ? i1 : 0, zb2 = p.b_use2 ? i2 : 0;
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
        const int k0 = (kb0 + it) * BK;
        const uint32_t sA = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + Cfg::A_BYTES;
        if (!p.a_mn) {
          tma_load_4d(sA, &p.tmA, full_bar(s), k0, m0, za1, za2);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) tma_load_4d(sA + c * 4096, &p.tmA, full_bar(s), m0 + c * 32, k0, za1, za2);
        }
        if (!p.b_mn) {
          if (BN <= 256) {
            // a TMA box dimension is limited to 256 rows
            tma_load_4d(sB, &p.tmB, full_bar(s), k0, n0, zb1, zb2);
          }
        } else {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_4d(sB + c * 4096, &p.tmB, full_bar(s), n0 + c * 32, k0, zb1, zb2);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, majors, N>>3, M>>4  (cute/arch/mma_sm100_desc.hpp bit layout)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sA = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < BK / 8; ++k4) {
          const uint64_t adesc = p.a_mn ? umma_smem_desc(sA + k4 * 1024, 4096, 512, 1)
                                        : umma_smem_desc(sA + k4 * 32, 16, 1024, 2);
          const uint64_t bdesc = p.b_mn ? umma_smem_desc(sB + k4 * 1024, 4096, 512, 1)
                                        : umma_smem_desc(sB + k4 * 32, 16, 1024, 2);
          umma_tf32(tmem_base, adesc, bdesc, idesc, (it | k4) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));     // frees this smem stage once the MMAs above have read it
      }
      umma_commit(tmem_full_bar);      // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = m0 + q * 32 + lane;
    const int flags = p.flags;
    float alpha = p.alpha, beta = p.beta;
    if (p.alpha_dev) alpha *= __ldg(p.alpha_dev);
    if (p.beta_dev) beta *= __ldg(p.beta_dev);
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
          if (Xrow) {
            if (full && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(Xrow + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) Xrow[col0 + j] = v[j];
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
        }
        if (flags & UVC_EPI_GELU_BWD) {
          if (full && vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 u = *reinterpret_cast<const float4*>(Xrow + col0 + j);
              v[j] *= gelu_grad_f(u.x); v[j + 1] *= gelu_grad_f(u.y); v[j + 2] *= gelu_grad_f(u.z); v[j + 3] *= gelu_grad_f(u.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] *= gelu_grad_f(Xrow[col0 + j]);
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

template <int BN, int STAGES>
static int launch(const GemmKParams& kp, dim3 grid, cudaStream_t st) {
  using Cfg = GemmCfg<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    UVC_REQUIRE(e == cudaSuccess, UVC_ERR_CUDA, "cudaFuncSetAttribute(gemm smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  gemm_tf32_kernel<BN, STAGES><<<grid, kThreads, Cfg::SMEM_BYTES, st>>>(kp);
  return check_launch("gemm_tf32_kernel");
}

int gemm_tf32(const uvc_gemm_args& a, cudaStream_t st) {
  UVC_REQUIRE(a.M > 0 && a.N > 0 && a.K >= 0, UVC_ERR_BAD_SHAPE, "gemm: bad M,N,K = %d,%d,%d", a.M, a.N, a.K);
  UVC_REQUIRE(a.nb1 >= 1 && a.nb2 >= 1 && a.splits >= 1, UVC_ERR_BAD_ARG, "gemm: nb1, nb2, splits must be >= 1");
  UVC_REQUIRE(a.D != nullptr, UVC_ERR_BAD_ARG, "gemm: D is NULL");
  UVC_REQUIRE(a.splits == 1 || (a.flags & UVC_EPI_ATOMIC), UVC_ERR_BAD_ARG, "gemm: splits > 1 requires UVC_EPI_ATOMIC");
  UVC_REQUIRE(!(a.flags & UVC_EPI_ATOMIC) || !(a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD)), UVC_ERR_BAD_ARG, "gemm: GELU epilogues cannot be combined with split-K accumulation");
  UVC_REQUIRE(!(a.flags & UVC_EPI_ATOMIC) || !(a.flags & UVC_EPI_ROUND_TF32), UVC_ERR_BAD_ARG, "gemm: UVC_EPI_ROUND_TF32 cannot be combined with atomic accumulation");
  UVC_REQUIRE(!(a.flags & UVC_EPI_BIAS) || a.bias, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_BIAS without bias");
  UVC_REQUIRE(!(a.flags & UVC_EPI_RESIDUAL) || a.R, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_RESIDUAL without R");
  UVC_REQUIRE(!(a.flags & UVC_EPI_GELU_BWD) || a.aux, UVC_ERR_BAD_ARG, "gemm: UVC_EPI_GELU_BWD without aux");
  const int nkb = (a.K + BK - 1) / BK;
  int splits = a.splits;
  if (splits > nkb) splits = nkb > 0 ? nkb : 1;

  GemmKParams kp;
  constexpr int BN = 128;
  int rc = make_tmap(&kp.tmA, a.A, a.M, a.K, a.nb1, a.nb2, BM, "A");
  if (rc) return rc;
  rc = make_tmap(&kp.tmB, a.B, a.N, a.K, a.nb1, a.nb2, BN, "B");
  if (rc) return rc;
  kp.D = a.D; kp.ldd = a.ldd; kp.d_bs1 = a.d_bs1; kp.d_bs2 = a.d_bs2;
  kp.bias = a.bias;
  kp.R = (a.flags & UVC_EPI_RESIDUAL) ? a.R : nullptr; kp.ldr = a.ldr; kp.r_bs1 = a.r_bs1; kp.r_bs2 = a.r_bs2;
  kp.aux = (a.flags & (UVC_EPI_GELU | UVC_EPI_GELU_BWD)) ? a.aux : nullptr; kp.ldaux = a.ldaux; kp.aux_bs1 = a.aux_bs1; kp.aux_bs2 = a.aux_bs2;
  kp.alpha_dev = a.alpha_dev; kp.beta_dev = a.beta_dev;
  kp.alpha = a.alpha; kp.beta = a.beta;
  kp.M = a.M; kp.N = a.N; kp.K = a.K; kp.nb1 = a.nb1; kp.nb2 = a.nb2; kp.splits = splits; kp.flags = a.flags;
  kp.a_mn = a.A.mn_major ? 1 : 0; kp.b_mn = a.B.mn_major ? 1 : 0;
  kp.a_use1 = a.A.bs1 != 0; kp.a_use2 = a.A.bs2 != 0; kp.b_use1 = a.B.bs1 != 0; kp.b_use2 = a.B.bs2 != 0;

  const long long gz = (long long)a.nb1 * a.nb2 * splits;
  const long long gy = (a.M + BM - 1) / BM;
  UVC_REQUIRE(gz <= 65535 && gy <= 65535, UVC_ERR_BAD_SHAPE, "gemm: grid too large (m tiles %lld, batch*splits %lld)", gy, gz);
  dim3 grid((a.N + BN - 1) / BN, (unsigned)gy, (unsigned)gz);
  const bool prof = prof_enabled();
  if (prof) prof_begin(st, 2.0 * a.M * a.N * (double)a.K * a.nb1 * a.nb2);
  rc = launch<BN, 3>(kp, grid, st);
  if (prof) prof_end(st);
  return rc;
}

}  // namespace uvc

extern "C" int uvc_gemm_tf32(const uvc_gemm_args* args, void* stream) {
  if (!args) { uvc::set_error("uvc_gemm_tf32: args is NULL"); return UVC_ERR_BAD_ARG; }
  return uvc::gemm_tf32(*args, static_cast<cudaStream_t>(stream));
}

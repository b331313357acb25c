/* uvc_b200.h — C ABI of libuvc_sm100.so: the B200-native replacement for the device work on
 * UVC's data-parallel hot path.
 *
 * The reference (VITA-Group/UVC) has NO native interface: every op below replaces a span of
 * Python/ATen code, cited per entry point as `file:line` under /root/reference/UVC.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer (including workspaces); the library never allocates or frees
 *     device memory and never keeps a caller pointer after a call returns;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); there are no
 *     hidden host syncs, so every call can be captured into a CUDA graph;
 *   - return value: 0 on success, a negative uvc_status otherwise; uvc_last_error() gives the text;
 *     nothing throws or aborts across this boundary;
 *   - all matrices are fp32 row-major; tensor-core math is TF32 (fp32 storage, fp32 accumulate).
 */
#ifndef UVC_B200_H_
#define UVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVC_ABI_VERSION 1

#if defined(UVC_BUILD_DLL)
#define UVC_API __attribute__((visibility("default")))
#else
#define UVC_API
#endif

typedef enum {
  UVC_OK = 0,
  UVC_ERR_BAD_SHAPE = -1,    /* a dimension / stride / alignment the kernels do not support */
  UVC_ERR_BAD_ARG = -2,      /* null pointer, bad enum, inconsistent arguments */
  UVC_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed (text in uvc_last_error) */
  UVC_ERR_WORKSPACE = -4     /* workspace too small (see uvc_*_workspace_bytes) */
} uvc_status;

UVC_API int uvc_version(void);                 /* ABI version, bumped on any signature change */
UVC_API const char* uvc_last_error(void);      /* thread-local text of the last failure */
UVC_API int uvc_abi_sizeof(const char* struct_name);   /* sizeof() of an ABI struct, for binding self-checks */

/* ------------------------------------------------------------------------------------------
 * Batched TF32 tensor-core GEMM (tcgen05.mma + TMEM accumulator + TMA operand staging)
 *   D[z] = epilogue( alpha * op(A[z]) . op(B[z])^T )        z = i2 * nb1 + i1
 * Replaces every nn.Linear / torch.matmul / autograd GEMM on the path:
 *   models/model_distilled.py:116,122 (fc1, fc2), :175 (qkv), :179,184 (QK^T, PV), :187 (proj),
 *   :149 (patch-embed conv as GEMM), :522 (head) and their autograd backward.
 *
 * Operand A is logically [M, K], operand B is logically [N, K].
 *   mn_major = 0: memory is [rows = M|N][cols = K], row stride ld           ("K-major")
 *   mn_major = 1: memory is [rows = K][cols = M|N], row stride ld           ("MN-major", i.e. transposed)
 * so forward (X.W^T), dgrad (dY.W) and wgrad (dY^T.X) all run without a transpose pass.
 * ld and batch strides are in elements; ld % 4 == 0 and ptr % 16 == 0 (TMA), batch strides % 4 == 0.
 */
typedef struct {
  const float* ptr;
  int64_t ld;
  int64_t bs1, bs2;     /* batch strides for i1, i2 (0 = operand shared across that batch index) */
  int32_t mn_major;
  int32_t _pad;
} uvc_operand;

enum {
  UVC_EPI_BIAS = 1,        /* v += bias[col] */
  UVC_EPI_GELU = 2,        /* aux[row,col] = v (pre-activation, if aux != NULL); v = gelu_erf(v) */
  UVC_EPI_GELU_BWD = 4,    /* v *= gelu'(aux[row,col]) */
  UVC_EPI_RESIDUAL = 8,    /* v += beta * R[row,col] */
  UVC_EPI_ATOMIC = 16      /* D += v with red.global.add (required when splits > 1) */
};

typedef struct {
  int32_t M, N, K;
  int32_t nb1, nb2;        /* batch counts (>= 1) */
  int32_t splits;          /* split-K factor (>= 1); > 1 requires UVC_EPI_ATOMIC */
  uvc_operand A, B;
  float* D; int64_t ldd, d_bs1, d_bs2;
  const float* bias;       /* [N] */
  const float* R; int64_t ldr, r_bs1, r_bs2;
  float* aux; int64_t ldaux, aux_bs1, aux_bs2;
  float alpha;             /* scales the accumulator before bias */
  float beta;              /* scales R */
  const float* alpha_dev;  /* optional device scalar multiplied into alpha (gate values) */
  const float* beta_dev;   /* optional device scalar multiplied into beta */
  int32_t flags;           /* UVC_EPI_* */
  int32_t _pad;
} uvc_gemm_args;

UVC_API int uvc_gemm_tf32(const uvc_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* UVC_B200_H_ */

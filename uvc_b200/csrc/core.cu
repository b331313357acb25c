// Error plumbing + ABI introspection for libuvc_sm100.so.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace uvc {
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UVC_ERR_CUDA;
  }
  return UVC_OK;
}
}  // namespace uvc

extern "C" int uvc_version(void) { return UVC_ABI_VERSION; }
extern "C" const char* uvc_last_error(void) { return uvc::g_err; }

extern "C" int uvc_abi_sizeof(const char* name) {
  if (!name) return -1;
#define UVC_SZ(T) if (strcmp(name, #T) == 0) return (int)sizeof(T)
  UVC_SZ(uvc_operand);
  UVC_SZ(uvc_gemm_args);
  UVC_SZ(uvc_block_tensors);
  UVC_SZ(uvc_vit_tensors);
  UVC_SZ(uvc_vit_dims);
  UVC_SZ(uvc_vit_forward_args);
  UVC_SZ(uvc_vit_backward_args);
#undef UVC_SZ
  return -1;
}

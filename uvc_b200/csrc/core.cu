// Error plumbing + ABI introspection for libuvc_sm100.so.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <vector>
#include <mutex>

namespace uvc {
static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("UVC_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on != 0;
}

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UVC_ERR_CUDA;
  }
  return UVC_OK;
}

// ---- optional per-launch timing of the GEMM kernel with CUDA events on the launching stream (bench.py's roofline leg)
static bool g_prof_on = false;
struct ProfRec { cudaEvent_t e0, e1; double flops; int kind; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
bool prof_enabled() { return g_prof_on; }
void prof_begin(cudaStream_t st, double flops, int kind) {
  ProfRec r; r.flops = flops; r.kind = kind;
  cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, st);
}
}  // namespace uvc

extern "C" long long uvc_launch_count(void) { return uvc::launch_count(); }
extern "C" int uvc_gemm_profile(int enable) {
  std::lock_guard<std::mutex> lk(uvc::g_prof_mu);
  for (auto& r : uvc::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  uvc::g_prof.clear();
  uvc::g_prof_on = enable != 0;
  return UVC_OK;
}
extern "C" int uvc_gemm_profile_read_kind(int kind, double* total_ms, double* total_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(uvc::g_prof_mu);
  double ms = 0.0, fl = 0.0;
  long long n = 0;
  for (auto& r : uvc::g_prof) {
    if (kind && r.kind != kind) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) { uvc::set_error("uvc_gemm_profile_read_kind: event sync failed"); return UVC_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms += t; fl += r.flops; ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = n;
  return UVC_OK;
}
extern "C" int uvc_gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(uvc::g_prof_mu);
  double ms = 0.0, fl = 0.0;
  for (auto& r : uvc::g_prof) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) { uvc::set_error("uvc_gemm_profile_read: event sync failed"); return UVC_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms += t; fl += r.flops;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = (long long)uvc::g_prof.size();
  return UVC_OK;
}

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
  UVC_SZ(uvc_vit_layout);
  UVC_SZ(uvc_t2t_tensors);
  UVC_SZ(uvc_t2t_forward_args);
  UVC_SZ(uvc_t2t_backward_args);
  UVC_SZ(uvc_vit_forward_args);
  UVC_SZ(uvc_vit_backward_args);
  UVC_SZ(uvc_admm_args);
#undef UVC_SZ
  return -1;
}

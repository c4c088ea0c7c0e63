// Library runtime: error string, device query, version.
#include "common.cuh"
#include <mutex>

namespace {
thread_local char g_err[512] = "";
int g_sms = 0;
}

void vq_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int vq_sm_count() {
  if (g_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sms = n;
    else
      g_sms = 148;
  }
  return g_sms;
}

VQ_API const char* vqcpc_last_error(void) { return g_err; }
VQ_API int vqcpc_abi_version(void) { return 1; }

// sm_count / compute capability of the current device; fails on anything that is not sm_100 (B200)
VQ_API int vqcpc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  VQ_CUDA_CALL(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VQ_CUDA_CALL(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return VQ_OK;
}

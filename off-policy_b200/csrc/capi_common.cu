// Error reporting, launch accounting and device queries shared by every entry point.
#include <stdarg.h>
#include <stdio.h>

#include "mx_internal.h"

static thread_local char g_err[512] = "";
long long g_mx_launches = 0;

void mx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mx_last_error(void) { return g_err; }
extern "C" int mx_abi_version(void) { return MX_ABI_VERSION; }
extern "C" int mx_is_cuda_build(void) { return MX_EMU ? 0 : 1; }
extern "C" int64_t mx_launch_count(void) { return g_mx_launches; }

#if !MX_EMU
int mx_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
int mx_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    mx_set_error("CUDA error after %s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}
#endif

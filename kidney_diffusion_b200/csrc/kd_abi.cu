// Error plumbing, version and device check for libkidney_b200.
#include <stdarg.h>
#include <stdlib.h>

#include "kd_common.cuh"

static thread_local char g_last_error[1024] = "";

void kd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

// programmatic dependent launch between consecutive kernels of a step; KD_NO_PDL=1 in the environment switches it off (A/B timing)
int g_kd_pdl = getenv("KD_NO_PDL") ? 0 : 1;

int kd_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

extern "C" int kd_version(void) { return 100; }

extern "C" const char* kd_last_error(void) { return g_last_error; }

extern "C" int kd_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  KD_CUDA(cudaGetDevice(&dev));
  KD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  KD_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10 || minor != 0)
    KD_FAIL(KD_ERR_ARCH, "libkidney_b200 is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)", dev, major, minor);
  return KD_OK;
}

// Version / error plumbing of libfyc_sm100a.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void fyc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fyc_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}

extern "C" int32_t fyc_version(void) { return FYC_VERSION; }
extern "C" const char* fyc_last_error(void) { return g_err; }

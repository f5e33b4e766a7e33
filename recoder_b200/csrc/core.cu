// Error reporting and device queries shared by all entry points.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void rcd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int rcd_num_sms() {
  static int sms = 0;
  if (sms > 0) return sms;
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  sms = v;
  return sms;
}

RCD_EXPORT int rcd_abi_version(void) { return RCD_ABI_VERSION; }
RCD_EXPORT const char* rcd_last_error(void) { return g_err; }
RCD_EXPORT int rcd_device_sms(void) {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    rcd_set_error("rcd_device_sms: no CUDA device");
    return RCD_ERR_CUDA;
  }
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    rcd_set_error("rcd_device_sms: attribute query failed");
    return RCD_ERR_CUDA;
  }
  return v;
}

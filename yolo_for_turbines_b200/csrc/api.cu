// Error reporting + device query of the C-ABI (include/yolo_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void yb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* yolo_last_error(void) { return g_err; }
extern "C" int yolo_version(void) { return 100; }  // 0.1.0

extern "C" int yolo_device_info(int device, int* cc_major, int* cc_minor, int* sm_count) {
  cudaDeviceProp prop;
  YB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    yb_set_error("device %d is sm_%d%d; libyolo_b200 holds sm_100a code only", device, prop.major,
                 prop.minor);
    return YB_ERR_UNSUPPORTED;
  }
  return YB_OK;
}

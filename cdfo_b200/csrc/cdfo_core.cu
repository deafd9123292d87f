// Error plumbing and device probe of the C ABI.
#include "cdfo_common.cuh"

namespace cdfo {

static thread_local char g_err[512] = {0};

char *last_error_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace cdfo

extern "C" const char *cdfo_last_error(void) { return cdfo::last_error_buf(); }

extern "C" int cdfo_version(void) { return 1000; }

extern "C" int cdfo_device_ok(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cdfo::fail(CDFO_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
  return prop.major == 10 ? 1 : 0;
}

// ABI bookkeeping: version, thread-local error string, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace recad {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached_dev = dev;
    cached = n;
  }
  return cached;
}

}  // namespace recad

extern "C" {

int recad_abi_version(void) { return RECAD_ABI_VERSION; }

const char* recad_last_error(void) { return recad::g_err.c_str(); }

int recad_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  RECAD_CUDA_CHECK(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  RECAD_CUDA_CHECK(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  RECAD_CUDA_CHECK(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  RECAD_CUDA_CHECK(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = a;
  if (major) *major = b;
  if (minor) *minor = c;
  RECAD_REQUIRE(b == 10, RECAD_ERR_UNSUPPORTED, "recad_b200 is built for sm_100a only; device is sm_%d%d", b, c);
  return RECAD_OK;
}

}  // extern "C"

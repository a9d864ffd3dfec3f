// Error plumbing and library-level entry points of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace asb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return AS_OK;
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return AS_ERR_CUDA;
}

int check_arch() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = AS_ERR_ARCH;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaDeviceGetAttribute");
  cached_dev = dev;
  if (major != 10) {
    set_error("artspeech_b200 requires an sm_100 (B200) device, found compute capability %d.x",
              major);
    cached_rc = AS_ERR_ARCH;
  } else {
    cached_rc = AS_OK;
  }
  return cached_rc;
}

}  // namespace asb

namespace asb { int g_sm_limit = 0; }

extern "C" int as_version(void) { return 100; }
extern "C" int as_set_sm_limit(int32_t n) {
  const int prev = asb::g_sm_limit;
  asb::g_sm_limit = n > 0 ? n : 0;
  return prev;
}
extern "C" const char* as_last_error(void) { return asb::g_err; }

// Library-wide entry points of the C ABI: error string, version, device check.
#include "util.h"

namespace dlc {
char* last_error_buf() {
  static thread_local char buf[kErrBufLen] = {0};
  return buf;
}
}  // namespace dlc

using namespace dlc;

extern "C" const char* dlc_last_error(void) { return last_error_buf(); }
extern "C" int dlc_version(void) { return 100; }

extern "C" int dlc_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  DLC_CUDA(cudaGetDevice(&dev));
  DLC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DLC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return fail(DLC_EUNSUPPORTED, "dlc: device %d has compute capability %d.%d; this library is sm_100a only", dev,
                major, minor);
  return DLC_OK;
}

extern "C" int dlc_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

// Host-side helpers shared by the C-ABI translation units: error reporting and argument checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dlc.h"

namespace dlc {

char* last_error_buf();  // thread-local, defined in capi.cu
constexpr int kErrBufLen = 512;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), kErrBufLen, fmt, ap);
  va_end(ap);
  return code;
}

#define DLC_CHECK_ARG(cond)                                                                   \
  do {                                                                                        \
    if (!(cond)) return ::dlc::fail(DLC_EINVAL, "%s: invalid argument: %s", __func__, #cond); \
  } while (0)

#define DLC_CUDA(expr)                                                                                          \
  do {                                                                                                          \
    cudaError_t e_ = (expr);                                                                                    \
    if (e_ != cudaSuccess)                                                                                      \
      return ::dlc::fail(DLC_ECUDA, "%s: %s failed: %s", __func__, #expr, cudaGetErrorString(e_));              \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace dlc

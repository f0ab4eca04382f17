// a6 / a10 tails: score or distance matrix -> 8-bit image.
//   similarity (src/sdav/create_similarity_matrix.py:41-45):  move = 0 - min(M); divide = max(M) + move;
//                                                              img = 255 * ((M + move) / divide)
//   distance   (src/cnn_vtl/create_distance_matrix.py:40):     img = 255 - M / max(M) * 255
// both followed by cv2.imwrite (:48 / :41), which converts the float64 array with saturate_cast<uchar>: round half
// to even, clamp to 0..255. The reference's matrices are int64 (np.full([n, n], -1), :31): a float score is
// truncated toward zero when stored (`truncate_int`). All arithmetic in float64, in the reference's operation order.
// Two launches: block-wise min / max (order-independent, hence deterministic), then the map.
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "util.h"

namespace dlc {

constexpr int kImgBlocks = 592;   // 4 per SM
constexpr int kImgThreads = 256;

template <typename T>
__device__ __forceinline__ double img_value(const T* m, int64_t i, int truncate_int) {
  double v = static_cast<double>(m[i]);
  if (truncate_int) v = trunc(v);
  return v;
}

// non-finite entries (a matched pair of identical patches scores +inf, SimilarityCalculator.py:48) take no part in
// the range: in the reference they become INT64_MIN on the int64 store and wreck the normalisation; here they
// saturate (+inf -> 255, -inf / NaN -> 0) and the finite entries keep a meaningful range. Documented divergence.
template <typename T>
__global__ void __launch_bounds__(kImgThreads)
matrix_minmax_kernel(const T* __restrict__ m, int64_t n, int truncate_int, double* __restrict__ part) {
  __shared__ double s_min[kImgThreads / 32], s_max[kImgThreads / 32];
  double lo = INFINITY, hi = -INFINITY;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double v = img_value(m, i, truncate_int);
    if (isfinite(v)) {
      lo = fmin(lo, v);
      hi = fmax(hi, v);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) {
    s_min[threadIdx.x >> 5] = lo;
    s_max[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kImgThreads / 32; ++w) {
      lo = fmin(lo, s_min[w]);
      hi = fmax(hi, s_max[w]);
    }
    part[2 * blockIdx.x] = lo;
    part[2 * blockIdx.x + 1] = hi;
  }
}

__device__ __forceinline__ uint8_t saturate_u8(double x) {
  if (!(x == x)) return 0;                 // NaN (0 / 0 of a constant matrix): cvRound gives INT_MIN -> 0
  const double r = rint(x);                // cvRound: round half to even
  return r <= 0.0 ? 0 : (r >= 255.0 ? 255 : static_cast<uint8_t>(r));
}

template <typename T>
__global__ void __launch_bounds__(kImgThreads)
matrix_image_kernel(const T* __restrict__ m, int64_t n, int truncate_int, int mode, const double* __restrict__ part,
                    int n_part, uint8_t* __restrict__ out) {
  __shared__ double s_min, s_max;
  if (threadIdx.x < 32) {
    double lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < n_part; i += 32) {
      lo = fmin(lo, part[2 * i]);
      hi = fmax(hi, part[2 * i + 1]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (threadIdx.x == 0) {
      s_min = lo;
      s_max = hi;
    }
  }
  __syncthreads();
  const double vmin = s_min, vmax = s_max;
  const double move = 0.0 - vmin, divide = vmax + move;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double v = img_value(m, i, truncate_int);
    double img;
    if (mode == DLC_IMG_SIMILARITY)
      img = 255.0 * ((v + move) / divide);
    else
      img = 255.0 - v / vmax * 255.0;
    // +inf is the best similarity (white) / the worst distance (black); -inf the opposite; NaN black
    if (!isfinite(v)) img = (v > 0.0) == (mode == DLC_IMG_SIMILARITY) ? 255.0 : 0.0;
    if (v != v) img = 0.0;
    out[i] = saturate_u8(img);
  }
}

template <typename T>
static int run_image(const T* m, int64_t n, int truncate_int, int mode, uint8_t* out, double* part, cudaStream_t s) {
  const int blocks = static_cast<int>(std::min<int64_t>(kImgBlocks, ceil_div64(n, kImgThreads)));
  matrix_minmax_kernel<T><<<blocks, kImgThreads, 0, s>>>(m, n, truncate_int, part);
  matrix_image_kernel<T><<<blocks, kImgThreads, 0, s>>>(m, n, truncate_int, mode, part, blocks, out);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

}  // namespace dlc

using namespace dlc;

extern "C" size_t dlc_matrix_image_workspace_bytes(void) { return sizeof(double) * 2 * kImgBlocks; }

extern "C" int dlc_matrix_image(const void* m_dev, int dtype, int rows, int cols, int mode, int truncate_int,
                                uint8_t* out_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(rows >= 0 && cols >= 0);
  DLC_CHECK_ARG(dtype == DLC_F32 || dtype == DLC_I32);
  DLC_CHECK_ARG(mode == DLC_IMG_SIMILARITY || mode == DLC_IMG_DISTANCE);
  const int64_t n = static_cast<int64_t>(rows) * cols;
  if (n == 0) return DLC_OK;
  DLC_CHECK_ARG(m_dev && out_dev && ws_dev);
  if (ws_bytes < dlc_matrix_image_workspace_bytes())
    return fail(DLC_ENOMEM, "dlc_matrix_image: workspace of %zu bytes needed", dlc_matrix_image_workspace_bytes());
  double* part = static_cast<double*>(ws_dev);
  if (dtype == DLC_F32)
    return run_image(static_cast<const float*>(m_dev), n, truncate_int, mode, out_dev, part, as_stream(stream));
  return run_image(static_cast<const int32_t*>(m_dev), n, 0, mode, out_dev, part, as_stream(stream));
}

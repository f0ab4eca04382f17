// a1: patch gather + normalise. Replaces get_vectorized_patches_from_key_points, get_1d_boundaries /
// get_2d_boundaries and the `/ 255.0` of CvInputParser.parse (src/sdav/input/CvInputParser.py:100-123, 49-97, 27).
// Pure gather, bytes-bound: one CTA per patch, 16-byte vectorised stores of the fp16 hi/lo operand planes that
// the first encoder layer consumes through TMA (K padded 1681 -> 1728 with zeros).
#include "ptx.cuh"
#include "util.h"

namespace dlc {

// Lower bound of a `patch`-wide window centred on c, shifted to lie inside [0, L) exactly as get_1d_boundaries does
// (CvInputParser.py:69-86): shift forward when it starts below 0, shift back when it ends beyond L-1.
__host__ __device__ inline int window_lo(int c, int L, int patch) {
  const int half = patch / 2;
  const int lo = c - half;
  const int hi = c + half;
  const int fwd = lo < 0 ? -lo : 0;
  const int aux = hi - L + 1;
  const int back = aux > 0 ? aux : 0;
  return lo - back + fwd;
}

struct PatchOrigin {
  int r0, c0;
};
__device__ __forceinline__ PatchOrigin patch_origin(const float* xy, int H, int W, int patch, int swap_xy_quirk) {
  // Python's round() on the keypoint coordinates = round half to even (CvInputParser.py:111).
  const int x = static_cast<int>(rintf(xy[0]));
  const int y = static_cast<int>(rintf(xy[1]));
  PatchOrigin o;
  if (swap_xy_quirk) {  // reference: axis 0 (rows, length H) is driven by x, axis 1 (cols, length W) by y
    o.r0 = window_lo(x, H, patch);
    o.c0 = window_lo(y, W, patch);
  } else {
    o.r0 = window_lo(y, H, patch);
    o.c0 = window_lo(x, W, patch);
  }
  return o;
}

// (hi, lo) fp16 pair of v / 255 for the 256 pixel values, computed once per process on the device.
__device__ uint32_t g_pixel_lut[256];
__global__ void pixel_lut_kernel() {
  const int i = threadIdx.x;
  __half h, l;
  split_f64(static_cast<double>(i) / 255.0, h, l);
  g_pixel_lut[i] = static_cast<uint32_t>(__half_as_ushort(h)) | (static_cast<uint32_t>(__half_as_ushort(l)) << 16);
}

// kPatchesPerCta patches per CTA (one warp-pair each); a thread produces 8 consecutive K entries = one 16-byte store
// per plane, walking (patch row, patch column) incrementally instead of dividing per element.
constexpr int kPatchesPerCta = 4;
constexpr int kThreadsPerPatch = 64;
__global__ void __launch_bounds__(kPatchesPerCta* kThreadsPerPatch)
patch_gather_planes_kernel(const uint8_t* __restrict__ img, int H, int W, const float* __restrict__ xy, int P,
                           int64_t n_patches, int patch, int swap_xy_quirk, __half* __restrict__ out_hi,
                           __half* __restrict__ out_lo, int ld, int raw) {
  __shared__ uint32_t lut[256];
  // raw: the plane holds the pixel value itself (0..255, exact in fp16; the encoder folds the 1/255 into layer 0)
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    lut[i] = raw ? static_cast<uint32_t>(__half_as_ushort(__int2half_rn(i))) : g_pixel_lut[i];
  __syncthreads();
  const int64_t pidx = static_cast<int64_t>(blockIdx.x) * kPatchesPerCta + threadIdx.x / kThreadsPerPatch;  // b * P + p
  if (pidx >= n_patches) return;
  const int t = threadIdx.x % kThreadsPerPatch;
  const int b = static_cast<int>(pidx / P);
  const PatchOrigin o = patch_origin(xy + pidx * 2, H, W, patch, swap_xy_quirk);
  const uint8_t* src = img + static_cast<int64_t>(b) * H * W + static_cast<int64_t>(o.r0) * W + o.c0;
  const int n = patch * patch;
  const int chunks = ld >> 3;
  for (int c = t; c < chunks; c += kThreadsPerPatch) {
    int e = c * 8;
    int pr = e / patch;
    int pc = e - pr * patch;
    uint32_t v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q, ++e) {
      v[q] = e < n ? lut[src[pr * W + pc]] : 0u;  // (0, 0) for the K padding
      if (++pc == patch) {
        pc = 0;
        ++pr;
      }
    }
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      hh[q] = (v[2 * q] & 0xFFFFu) | (v[2 * q + 1] << 16);
      ll[q] = (v[2 * q] >> 16) | (v[2 * q + 1] & 0xFFFF0000u);
    }
    const int64_t off = pidx * ld + c * 8;
    *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

__global__ void __launch_bounds__(128)
patch_gather_f64_kernel(const uint8_t* __restrict__ img, int H, int W, const float* __restrict__ xy, int P, int patch,
                        int swap_xy_quirk, double* __restrict__ out) {
  const int64_t pidx = blockIdx.x;
  const int b = static_cast<int>(pidx / P);
  const PatchOrigin o = patch_origin(xy + pidx * 2, H, W, patch, swap_xy_quirk);
  const uint8_t* src = img + static_cast<int64_t>(b) * H * W;
  const int n = patch * patch;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int pr = e / patch;
    const int pc = e - pr * patch;
    out[pidx * n + e] = static_cast<double>(src[static_cast<int64_t>(o.r0 + pr) * W + (o.c0 + pc)]) / 255.0;
  }
}

}  // namespace dlc

using namespace dlc;

static int check_patch_args(const void* img, int B, int H, int W, const void* xy, int P, int patch) {
  DLC_CHECK_ARG(B >= 0 && P > 0);
  DLC_CHECK_ARG((img && xy) || B == 0);
  DLC_CHECK_ARG(patch > 0 && (patch & 1) == 1);  // CvInputParser.py:64-65: patch size must be odd
  DLC_CHECK_ARG(H >= patch && W >= patch);
  return DLC_OK;
}

extern "C" int dlc_patch_gather(const uint8_t* img_dev, int B, int H, int W, const float* xy_dev, int P, int patch,
                                int swap_xy_quirk, void* out_hi_dev, void* out_lo_dev, int ld, void* stream) {
  if (int rc = check_patch_args(img_dev, B, H, W, xy_dev, P, patch)) return rc;
  DLC_CHECK_ARG(out_hi_dev || B == 0);
  DLC_CHECK_ARG(ld >= patch * patch && ld % 8 == 0);
  if (B == 0) return DLC_OK;
  static bool lut_ready[64] = {false};  // per device; built (and waited for) once, so later calls on any stream see it
  int dev = 0;
  DLC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !lut_ready[dev]) {
    pixel_lut_kernel<<<1, 256, 0, as_stream(stream)>>>();
    DLC_CUDA(cudaGetLastError());
    DLC_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (dev >= 0 && dev < 64) lut_ready[dev] = true;
  }
  const int64_t n_patches = static_cast<int64_t>(B) * P;
  const int grid = static_cast<int>((n_patches + kPatchesPerCta - 1) / kPatchesPerCta);
  patch_gather_planes_kernel<<<grid, kPatchesPerCta * kThreadsPerPatch, 0, as_stream(stream)>>>(
      img_dev, H, W, xy_dev, P, n_patches, patch, swap_xy_quirk, static_cast<__half*>(out_hi_dev),
      static_cast<__half*>(out_lo_dev), ld, 0);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_patch_gather_u8(const uint8_t* img_dev, int B, int H, int W, const float* xy_dev, int P, int patch,
                                   int swap_xy_quirk, void* out_dev, int ld, void* stream) {
  if (int rc = check_patch_args(img_dev, B, H, W, xy_dev, P, patch)) return rc;
  DLC_CHECK_ARG(out_dev || B == 0);
  DLC_CHECK_ARG(ld >= patch * patch && ld % 8 == 0);
  if (B == 0) return DLC_OK;
  const int64_t n_patches = static_cast<int64_t>(B) * P;
  const int grid = static_cast<int>((n_patches + kPatchesPerCta - 1) / kPatchesPerCta);
  patch_gather_planes_kernel<<<grid, kPatchesPerCta * kThreadsPerPatch, 0, as_stream(stream)>>>(
      img_dev, H, W, xy_dev, P, n_patches, patch, swap_xy_quirk, static_cast<__half*>(out_dev), nullptr, ld, 1);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_patch_gather_f64(const uint8_t* img_dev, int B, int H, int W, const float* xy_dev, int P,
                                    int patch, int swap_xy_quirk, double* out_dev, void* stream) {
  if (int rc = check_patch_args(img_dev, B, H, W, xy_dev, P, patch)) return rc;
  DLC_CHECK_ARG(out_dev || B == 0);
  if (B == 0) return DLC_OK;
  patch_gather_f64_kernel<<<B * P, 128, 0, as_stream(stream)>>>(img_dev, H, W, xy_dev, P, patch, swap_xy_quirk,
                                                               out_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

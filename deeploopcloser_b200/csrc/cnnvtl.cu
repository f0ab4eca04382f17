// a7: building blocks of the cnn_vtl AlexNet conv head (src/cnn_vtl/network/cnn_vtl.py:28-128).
// Convolutions run as im2col operand planes + the tcgen05 contraction of planes.cu (bias + ReLU fused, NHWC in and
// out, so the GEMM output [N*OH*OW, Cout] IS the NHWC activation and also the flattened per-image descriptor
// segment of cnn_vtl.py:96-106). Max-pooling and the min/max -> int8 -> column-gather tail (cnn_vtl.py:109-128) are
// bytes-bound SIMT kernels; the tail touches the full 546,944-wide descriptor once (for min/max) and then only the
// ~2,243 kept columns.
#include <math.h>

#include <algorithm>

#include "ptx.cuh"
#include "util.h"

namespace dlc {

// x planes [N*H*W, ld_in] -> im2col planes [N*OH*OW, ld]; column = (kh*KW + kw)*C + c; 8 columns per thread.
__global__ void __launch_bounds__(256)
im2col_kernel(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo, int N, int H, int W, int C, int ld_in,
              int KH, int KW, int stride, int pad_t, int pad_l, int OH, int OW, __half* __restrict__ o_hi,
              __half* __restrict__ o_lo, int ld) {
  const int chunks = ld >> 3;
  const int K = KH * KW * C;
  const int64_t total = static_cast<int64_t>(N) * OH * OW * chunks;
  const bool vec = (C & 7) == 0;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(t % chunks);
    const int64_t orow = t / chunks;
    const int ow = static_cast<int>(orow % OW);
    const int oh = static_cast<int>((orow / OW) % OH);
    const int n = static_cast<int>(orow / (static_cast<int64_t>(OW) * OH));
    const int col0 = ch << 3;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
    if (vec) {
      if (col0 < K) {
        const int kk = col0 / C, c = col0 - kk * C;
        const int kh = kk / KW, kw = kk - kh * KW;
        const int ih = oh * stride - pad_t + kh, iw = ow * stride - pad_l + kw;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
          const int64_t src = ((static_cast<int64_t>(n) * H + ih) * W + iw) * ld_in + c;
          vh = *reinterpret_cast<const uint4*>(x_hi + src);
          if (x_lo) vl = *reinterpret_cast<const uint4*>(x_lo + src);
        }
      }
    } else {
      __align__(16) __half h[8];
      __align__(16) __half l[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        h[q] = __float2half_rn(0.f);
        l[q] = __float2half_rn(0.f);
        const int col = col0 + q;
        if (col < K) {
          const int kk = col / C, c = col - kk * C;
          const int kh = kk / KW, kw = kk - kh * KW;
          const int ih = oh * stride - pad_t + kh, iw = ow * stride - pad_l + kw;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            const int64_t src = ((static_cast<int64_t>(n) * H + ih) * W + iw) * ld_in + c;
            h[q] = x_hi[src];
            if (x_lo) l[q] = x_lo[src];
          }
        }
      }
      vh = *reinterpret_cast<const uint4*>(h);
      vl = *reinterpret_cast<const uint4*>(l);
    }
    const int64_t o = orow * ld + col0;
    *reinterpret_cast<uint4*>(o_hi + o) = vh;
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + o) = vl;
  }
}

// NHWC float32 max-pool (VALID) -> planes; 8 channels per thread
__global__ void __launch_bounds__(256)
maxpool_kernel(const float* __restrict__ x, int N, int H, int W, int C, int window, int stride, int OH, int OW,
               __half* __restrict__ o_hi, __half* __restrict__ o_lo, int ld) {
  const int chunks = ld >> 3;
  const int64_t total = static_cast<int64_t>(N) * OH * OW * chunks;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(t % chunks);
    const int64_t orow = t / chunks;
    const int ow = static_cast<int>(orow % OW);
    const int oh = static_cast<int>((orow / OW) % OH);
    const int n = static_cast<int>(orow / (static_cast<int64_t>(OW) * OH));
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = (ch << 3) + q;
      float m = 0.0f;
      if (c < C) {
        m = -INFINITY;
        for (int dy = 0; dy < window; ++dy)
          for (int dx = 0; dx < window; ++dx) {
            const int ih = oh * stride + dy, iw = ow * stride + dx;
            m = fmaxf(m, x[((static_cast<int64_t>(n) * H + ih) * W + iw) * C + c]);
          }
      }
      split_f32(m, h[q], l[q]);
    }
    const int64_t o = orow * ld + (ch << 3);
    *reinterpret_cast<uint4*>(o_hi + o) = *reinterpret_cast<const uint4*>(h);
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + o) = *reinterpret_cast<const uint4*>(l);
  }
}

// ---------------- descriptor tail: per-image min/max, then quantise only the kept columns ----------------
constexpr int kMaxSeg = 8;
struct SegTable {
  const float* ptr[kMaxSeg];
  int64_t size[kMaxSeg];   // floats per image in this segment
  int64_t start[kMaxSeg];  // first column of the segment in the concatenated descriptor
  int n_seg;
};

__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int N, int* mm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N) {
    mm[2 * t] = float_to_ordered(INFINITY);
    mm[2 * t + 1] = float_to_ordered(-INFINITY);
  }
}
// grid (slabs, n_seg, N)
__global__ void __launch_bounds__(256) minmax_kernel(SegTable st, int* mm) {
  const int seg = blockIdx.y, n = blockIdx.z;
  const int64_t size = st.size[seg];
  const float* p = st.ptr[seg] + static_cast<int64_t>(n) * size;
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < size;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = p[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 2 * n, float_to_ordered(lo));
    atomicMax(mm + 2 * n + 1, float_to_ordered(hi));
  }
}
// q = int8(trunc((d - min) * (255 / (max - min)))), arithmetic in float64 like the reference graph
// (cnn_vtl.py:109-116). The float64 -> int8 cast of values >= 128 is out of range; the x86 TensorFlow build
// converts through int32 and keeps the low byte (two's-complement wrap), which is what is reproduced here.
__global__ void quantise_gather_kernel(SegTable st, int N, const int64_t* __restrict__ keep, int M,
                                       const int* __restrict__ mm, int8_t* __restrict__ out) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= static_cast<int64_t>(N) * M) return;
  const int n = static_cast<int>(t / M), m = static_cast<int>(t % M);
  const int64_t col = keep[m];
  int seg = 0;
  while (seg + 1 < st.n_seg && col >= st.start[seg + 1]) ++seg;
  const double d = static_cast<double>(st.ptr[seg][static_cast<int64_t>(n) * st.size[seg] + (col - st.start[seg])]);
  const double lo = static_cast<double>(ordered_to_float(mm[2 * n]));
  const double hi = static_cast<double>(ordered_to_float(mm[2 * n + 1]));
  const double scaled = (d - lo) * (255.0 / (hi - lo));
  int q = 0;
  if (scaled == scaled && fabs(scaled) < 2147483648.0) q = static_cast<int>(scaled);  // trunc toward zero
  else q = static_cast<int>(0x80000000u);                                              // cvttsd2si "indefinite"
  out[t] = static_cast<int8_t>(q & 0xff);
}

}  // namespace dlc

using namespace dlc;

extern "C" int dlc_im2col_planes(const void* x_hi_dev, const void* x_lo_dev, int N, int H, int W, int C, int ld_in,
                                 int KH, int KW, int stride, int pad_t, int pad_l, int OH, int OW, void* out_hi_dev,
                                 void* out_lo_dev, int ld, void* stream) {
  DLC_CHECK_ARG(x_hi_dev && out_hi_dev);
  DLC_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && ld_in >= C);
  DLC_CHECK_ARG(KH > 0 && KW > 0 && stride > 0 && pad_t >= 0 && pad_l >= 0 && OH > 0 && OW > 0);
  DLC_CHECK_ARG(ld >= KH * KW * C && ld % 8 == 0);
  DLC_CHECK_ARG((C & 7) != 0 || (ld_in & 7) == 0);
  DLC_CHECK_ARG((OH - 1) * stride - pad_t + KH - 1 < H + KH && (OW - 1) * stride - pad_l + KW - 1 < W + KW);
  const int64_t total = static_cast<int64_t>(N) * OH * OW * (ld / 8);
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 32));
  im2col_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __half*>(x_hi_dev),
                                                     static_cast<const __half*>(x_lo_dev), N, H, W, C, ld_in, KH, KW,
                                                     stride, pad_t, pad_l, OH, OW, static_cast<__half*>(out_hi_dev),
                                                     static_cast<__half*>(out_lo_dev), ld);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_maxpool_planes(const float* x_dev, int N, int H, int W, int C, int window, int stride, int OH,
                                  int OW, void* out_hi_dev, void* out_lo_dev, int ld, void* stream) {
  DLC_CHECK_ARG(x_dev && out_hi_dev);
  DLC_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && window > 0 && stride > 0 && OH > 0 && OW > 0);
  DLC_CHECK_ARG((OH - 1) * stride + window <= H && (OW - 1) * stride + window <= W);
  DLC_CHECK_ARG(ld >= C && ld % 8 == 0);
  const int64_t total = static_cast<int64_t>(N) * OH * OW * (ld / 8);
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 32));
  maxpool_kernel<<<grid, 256, 0, as_stream(stream)>>>(x_dev, N, H, W, C, window, stride, OH, OW,
                                                      static_cast<__half*>(out_hi_dev),
                                                      static_cast<__half*>(out_lo_dev), ld);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_cnnvtl_quantise(const float* const* seg_ptrs_host, const int64_t* seg_sizes_host, int n_seg, int N,
                                   const int64_t* keep_cols_dev, int M, float* minmax_dev, int8_t* out_dev,
                                   void* stream) {
  DLC_CHECK_ARG(seg_ptrs_host && seg_sizes_host && keep_cols_dev && minmax_dev && out_dev);
  DLC_CHECK_ARG(n_seg >= 1 && n_seg <= kMaxSeg && N > 0 && M > 0);
  SegTable st{};
  st.n_seg = n_seg;
  int64_t start = 0;
  for (int i = 0; i < n_seg; ++i) {
    DLC_CHECK_ARG(seg_ptrs_host[i] && seg_sizes_host[i] > 0);
    st.ptr[i] = seg_ptrs_host[i];
    st.size[i] = seg_sizes_host[i];
    st.start[i] = start;
    start += seg_sizes_host[i];
  }
  cudaStream_t s = as_stream(stream);
  int* mm = reinterpret_cast<int*>(minmax_dev);
  minmax_init_kernel<<<ceil_div(N, 256), 256, 0, s>>>(N, mm);
  minmax_kernel<<<dim3(16, n_seg, N), 256, 0, s>>>(st, mm);
  const int64_t total = static_cast<int64_t>(N) * M;
  quantise_gather_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(st, N, keep_cols_dev, M, mm, out_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

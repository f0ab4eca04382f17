// a7: the cnn_vtl AlexNet conv head (src/cnn_vtl/network/cnn_vtl.py:28-128).
//
// Fused path (dlc_cnnvtl_* handle): every convolution is ONE tcgen05 kernel. conv2..conv5 (stride 1) are implicit
// GEMMs - the A operand is the NHWC activation itself, read tap by tap with im2col-mode TMA, so no im2col matrix
// ever exists in memory; conv1 (11x11 / 4 on 3 channels, 6-byte pixels: below TMA's 16-byte granule) reads a patch
// matrix written straight from the uint8 image. Bias + ReLU, the re-split of the activation into the next layer's
// fp16 planes, the per-image min/max of cnn_vtl.py:109-111 and the gather of the ~2,243 kept descriptor columns
// (:119-128) all happen in the GEMM epilogue: the 546,944-wide float descriptor of :96-106 is never materialised.
//
// Building blocks (dlc_im2col_planes / dlc_maxpool_planes / dlc_cnnvtl_quantise) are the explicit-im2col
// formulation of the same head; the GPU tests use them to cross-check the implicit path.
#include <math.h>

#include <algorithm>

#include <vector>

#include "bias_act.cuh"
#include "gemm_pair_sm100.cuh"
#include "util.h"

namespace dlc {

extern std::atomic<int> g_promote_k;  // planes.cu: K elements accumulated in TMEM before promotion to fp32 registers
extern std::atomic<int> g_cta_pair;   // planes.cu: CTA-pair kernels (0 never, 1 when the GPU is filled, 2 always)
int gemm_debug_flags();               // planes.cu

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

// ---------------- fused path kernels ----------------
template <typename T>
__device__ __forceinline__ void split_any(T x, __half& hi, __half& lo) {
  if (sizeof(T) == 8) split_f64(static_cast<double>(x), hi, lo);
  else split_f32(static_cast<float>(x), hi, lo);
}

// conv1 (11x11 / 4 VALID on 3 channels: 6-byte pixels, below TMA's 16-byte granule) is run as a 3x3 / 1 VALID
// convolution over the space-to-depth(4) image: s2d pixel (sh, sw) holds the 4 x 4 x 3 = 48 values of image rows
// 4sh..4sh+3, columns 4sw..4sw+3 in (dy, dx, c) order (96 bytes of a 128-byte pixel pitch), and filter tap (th, tw)
// channel (dy, dx, c) is W1[4th+dy, 4tw+dx, c] (zero where an index reaches 11). One thread = one (pixel, dy):
// 12 contiguous source elements -> 24 contiguous bytes per plane.
template <typename T>
__global__ void __launch_bounds__(256)
s2d4_planes_kernel(const T* __restrict__ x, int N, int H, int W, int SH, int SW, __half* __restrict__ o_hi,
                   __half* __restrict__ o_lo, int ld) {
  const int64_t total = static_cast<int64_t>(N) * SH * SW * 4;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int dy = static_cast<int>(t & 3);
    const int64_t spix = t >> 2;
    const int sw = static_cast<int>(spix % SW);
    const int sh = static_cast<int>((spix / SW) % SH);
    const int n = static_cast<int>(spix / (static_cast<int64_t>(SW) * SH));
    const int r = 4 * sh + dy;
    __align__(8) __half h[12];
    __align__(8) __half l[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
      h[q] = __float2half_rn(0.f);
      l[q] = __float2half_rn(0.f);
    }
    if (r < H) {
      const T* src = x + ((static_cast<int64_t>(n) * H + r) * W + 4 * sw) * 3;
      const int valid = min(4, W - 4 * sw) * 3;
#pragma unroll
      for (int q = 0; q < 12; ++q)
        if (q < valid) split_any(src[q], h[q], l[q]);
    }
    const int64_t o = spix * ld + dy * 12;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      reinterpret_cast<uint2*>(o_hi + o)[q] = reinterpret_cast<const uint2*>(h)[q];
      if (o_lo) reinterpret_cast<uint2*>(o_lo + o)[q] = reinterpret_cast<const uint2*>(l)[q];
    }
  }
}

// VALID max-pool on NHWC planes (value = hi + lo) -> planes; 8 channels per thread, only the C valid channels.
__global__ void __launch_bounds__(256)
maxpool_planes_kernel(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo, int N, int H, int W, int C,
                      int ld, int window, int stride, int OH, int OW, __half* __restrict__ o_hi,
                      __half* __restrict__ o_lo) {
  const int chunks = C >> 3;
  const int64_t total = static_cast<int64_t>(N) * OH * OW * chunks;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(t % chunks);
    const int64_t orow = t / chunks;
    const int ow = static_cast<int>(orow % OW);
    const int oh = static_cast<int>((orow / OW) % OH);
    const int n = static_cast<int>(orow / (static_cast<int64_t>(OW) * OH));
    float m[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) m[q] = -INFINITY;
    for (int dy = 0; dy < window; ++dy)
      for (int dx = 0; dx < window; ++dx) {
        const int64_t src = ((static_cast<int64_t>(n) * H + oh * stride + dy) * W + ow * stride + dx) * ld + (ch << 3);
        const uint4 vh = *reinterpret_cast<const uint4*>(x_hi + src);
        const __half* ph = reinterpret_cast<const __half*>(&vh);
        if (x_lo) {
          const uint4 vl = *reinterpret_cast<const uint4*>(x_lo + src);
          const __half* pl = reinterpret_cast<const __half*>(&vl);
#pragma unroll
          for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], __half2float(ph[q]) + __half2float(pl[q]));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], __half2float(ph[q]));
        }
      }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) split_f32(m[q], h[q], l[q]);
    const int64_t o = orow * ld + (ch << 3);
    *reinterpret_cast<uint4*>(o_hi + o) = *reinterpret_cast<const uint4*>(h);
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + o) = *reinterpret_cast<const uint4*>(l);
  }
}

// HWIO float64 filter [KH*KW, Cin, Cout] -> K-major planes [Cout, ld]: column = tap * c_pad + c (c_pad = Cin rounded
// up to the K block so that a K block never straddles two taps), zeros elsewhere.
__global__ void __launch_bounds__(256)
pack_conv_weight_kernel(const double* __restrict__ w, int taps, int cin, int cout, int c_pad, __half* __restrict__ hi,
                        __half* __restrict__ lo, int ld) {
  const int64_t total = static_cast<int64_t>(cout) * ld;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(t % cout);  // fastest across threads -> coalesced reads of w
    const int k = static_cast<int>(t / cout);
    const int tap = k / c_pad, c = k - tap * c_pad;
    __half h = __float2half_rn(0.f), l = __float2half_rn(0.f);
    if (tap < taps && c < cin) split_f64(w[(static_cast<int64_t>(tap) * cin + c) * cout + j], h, l);
    hi[static_cast<int64_t>(j) * ld + k] = h;
    if (lo) lo[static_cast<int64_t>(j) * ld + k] = l;
  }
}

// Descriptor tail of the fused head: gather the kept columns from the conv outputs' hi/lo planes (value = hi + lo,
// float32-exact to ~22 bits), scale with the per-image min/max the epilogues reduced, cast to int8 with the same
// arithmetic as quantise_gather_kernel (cnn_vtl.py:109-128).
struct PlaneSegTable {
  const __half* hi[kMaxSeg];
  const __half* lo[kMaxSeg];
  int64_t start[kMaxSeg];  // first column of the layer in the concatenated descriptor
  int cout[kMaxSeg];
  int ld[kMaxSeg];
  int pix[kMaxSeg];        // output pixels per image
  int n_seg;
};
__global__ void quantise_gather_planes_kernel(PlaneSegTable st, int N, const int64_t* __restrict__ keep, int M,
                                              const int* __restrict__ mm, int8_t* __restrict__ out) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (t >= static_cast<int64_t>(N) * M) return;
  const int n = static_cast<int>(t / M), m = static_cast<int>(t % M);
  const int64_t col = keep[m];
  int seg = 0;
  while (seg + 1 < st.n_seg && col >= st.start[seg + 1]) ++seg;
  const int64_t rel = col - st.start[seg];
  const int pixel = static_cast<int>(rel / st.cout[seg]), ch = static_cast<int>(rel % st.cout[seg]);
  const int64_t off = (static_cast<int64_t>(n) * st.pix[seg] + pixel) * st.ld[seg] + ch;
  const double d = static_cast<double>(__half2float(st.hi[seg][off]) + __half2float(st.lo[seg][off]));
  const double lo = static_cast<double>(ordered_to_float(mm[2 * n]));
  const double hi = static_cast<double>(ordered_to_float(mm[2 * n + 1]));
  const double scaled = (d - lo) * (255.0 / (hi - lo));
  int q = 0;
  if (scaled == scaled && fabs(scaled) < 2147483648.0) q = static_cast<int>(scaled);
  else q = static_cast<int>(0x80000000u);
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

// ------------------------------------------------------------------------------------------------------------
// Fused conv head behind a handle
// ------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kConvLayers = 5;
struct ConvSpec {
  int kh, kw, cin, cout, stride;
  bool same, relu, pool_after;
};
// cnn_vtl.py:33-93: ungrouped AlexNet convs, no LRN, conv5 linear, 3x3/2 max-pool after conv1 and conv2
constexpr ConvSpec kSpec[kConvLayers] = {
    {11, 11, 3, 96, 4, false, true, true},   {5, 5, 96, 256, 1, true, true, true}, {3, 3, 256, 384, 1, true, true, false},
    {3, 3, 384, 384, 1, true, true, false}, {3, 3, 384, 256, 1, true, false, false},
};

struct ConvGeo {
  int H, W, OH, OW;            // input / output spatial size
  int pad_t, pad_l, pad_b, pad_r;
  int in_ld;                   // pixel pitch (elements) of the input activation planes (implicit layers)
  int bk;                      // K block of this layer's kernel
  int c_pad, k_ld;             // channels per tap padded to bk; total K of the weight planes
  int out_ld;                  // pixel pitch of the output activation planes
  int PH, PW;                  // pooled size (pool_after)
  int vkh, vkw, vcin;          // the filter the kernel sees: the spec's, or 3 x 3 x 48 for the space-to-depth conv1
};

void same_pad(int size, int k, int stride, int* out, int* before, int* after) {
  *out = (size + stride - 1) / stride;
  int total = (*out - 1) * stride + k - size;
  if (total < 0) total = 0;
  *before = total / 2;  // TF puts the extra pixel after
  *after = total - total / 2;
}

}  // namespace

struct dlc_cnnvtl {
  int H = 0, W = 0, precision = DLC_PREC_FP16X2;
  ConvGeo geo[kConvLayers];
  void* w_hi[kConvLayers] = {nullptr};
  void* w_lo[kConvLayers] = {nullptr};
  float* bias[kConvLayers] = {nullptr};
  bool is_set[kConvLayers] = {false};
  int64_t seg_size[kConvLayers];   // descriptor columns contributed by each layer (OH*OW*Cout)
  int64_t seg_start[kConvLayers];
  int64_t total_cols = 0;
  int64_t* keep_cols = nullptr;    // device, strictly increasing
  int M = 0;
};

extern "C" int dlc_cnnvtl_create(dlc_cnnvtl** h, int H, int W, int precision) {
  DLC_CHECK_ARG(h);
  DLC_CHECK_ARG(precision == DLC_PREC_FP16 || precision == DLC_PREC_FP16X2);
  DLC_CHECK_ARG(H >= 11 && W >= 11 && H <= 8192 && W <= 8192);
  if (int rc = dlc_device_check()) return rc;
  dlc_cnnvtl* c = new dlc_cnnvtl();
  c->H = H;
  c->W = W;
  c->precision = precision;
  int hh = H, ww = W, in_ld = 0;
  int64_t start = 0;
  for (int l = 0; l < kConvLayers; ++l) {
    const ConvSpec& s = kSpec[l];
    ConvGeo& g = c->geo[l];
    g.H = hh;
    g.W = ww;
    g.in_ld = in_ld;
    if (s.same) {
      same_pad(hh, s.kh, s.stride, &g.OH, &g.pad_t, &g.pad_b);
      same_pad(ww, s.kw, s.stride, &g.OW, &g.pad_l, &g.pad_r);
    } else {
      g.OH = (hh - s.kh) / s.stride + 1;
      g.OW = (ww - s.kw) / s.stride + 1;
      g.pad_t = g.pad_l = g.pad_b = g.pad_r = 0;
    }
    g.vkh = s.kh;
    g.vkw = s.kw;
    g.vcin = s.cin;
    if (l == 0) {  // space-to-depth(4): 3x3 VALID over [OH + 2, OW + 2] pixels of 48 channels
      if (s.stride != 4 || s.kh > 12 || s.kw > 12 || s.cin != 3) {
        delete c;
        return fail(DLC_EUNSUPPORTED, "dlc_cnnvtl_create: conv1 must be <=12x12 / 4 on 3 channels");
      }
      g.H = g.OH + 2;
      g.W = g.OW + 2;
      g.in_ld = 64;
      g.vkh = g.vkw = 3;
      g.vcin = 48;
    }
    // implicit GEMM: a K block is bk channels of one tap
    g.bk = (precision == DLC_PREC_FP16X2 || g.vcin % 64 != 0) ? 32 : 64;
    // conv1: 48 channels + 16 zero-filled = one 64-wide block per tap. Its K loop is short and its accumulator
    // narrow, so the per-K-block issue overhead of the producer / MMA threads matters: use the wide block.
    if (l == 0) g.bk = 64;
    g.c_pad = (g.vcin + g.bk - 1) / g.bk * g.bk;
    g.k_ld = g.vkh * g.vkw * g.c_pad;
    g.out_ld = dlc_plane_ld(s.cout);
    c->seg_size[l] = static_cast<int64_t>(g.OH) * g.OW * s.cout;
    c->seg_start[l] = start;
    start += c->seg_size[l];
    hh = g.OH;
    ww = g.OW;
    g.PH = g.PW = 0;
    if (s.pool_after) {
      if (hh < 3 || ww < 3) {
        delete c;
        return fail(DLC_EINVAL, "dlc_cnnvtl_create: %dx%d input is too small for the conv head", H, W);
      }
      g.PH = hh = (hh - 3) / 2 + 1;
      g.PW = ww = (ww - 3) / 2 + 1;
    }
    in_ld = g.out_ld;
  }
  c->total_cols = start;
  for (int l = 0; l < kConvLayers; ++l) {
    const size_t plane = static_cast<size_t>(kSpec[l].cout) * c->geo[l].k_ld * 2;
    cudaError_t e = cudaMalloc(&c->w_hi[l], plane);
    if (e == cudaSuccess && precision == DLC_PREC_FP16X2) e = cudaMalloc(&c->w_lo[l], plane);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&c->bias[l]), sizeof(float) * kSpec[l].cout);
    if (e != cudaSuccess) {
      dlc_cnnvtl_destroy(c);
      return fail(DLC_ENOMEM, "dlc_cnnvtl_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
  }
  *h = c;
  return DLC_OK;
}

extern "C" int dlc_cnnvtl_destroy(dlc_cnnvtl* h) {
  if (!h) return DLC_OK;
  for (int l = 0; l < kConvLayers; ++l) {
    if (h->w_hi[l]) cudaFree(h->w_hi[l]);
    if (h->w_lo[l]) cudaFree(h->w_lo[l]);
    if (h->bias[l]) cudaFree(h->bias[l]);
  }
  if (h->keep_cols) cudaFree(h->keep_cols);
  delete h;
  return DLC_OK;
}

extern "C" int64_t dlc_cnnvtl_descriptor_len(const dlc_cnnvtl* h) { return h ? h->total_cols : 0; }

extern "C" int dlc_cnnvtl_set_conv(dlc_cnnvtl* h, int layer, const double* w_host, const double* b_host) {
  DLC_CHECK_ARG(h && w_host && b_host);
  DLC_CHECK_ARG(layer >= 0 && layer < kConvLayers);
  const ConvSpec& s = kSpec[layer];
  const ConvGeo& g = h->geo[layer];
  const int taps = g.vkh * g.vkw;
  const size_t wcount = static_cast<size_t>(taps) * g.vcin * s.cout;
  const size_t wbytes = sizeof(double) * wcount;
  std::vector<double> w_s2d;
  const double* w_src = w_host;
  if (layer == 0) {  // W1[kh, kw, c, :] -> tap (kh / 4, kw / 4), channel ((kh % 4) * 4 + kw % 4) * 3 + c
    w_s2d.assign(wcount, 0.0);
    for (int kh = 0; kh < s.kh; ++kh)
      for (int kw = 0; kw < s.kw; ++kw)
        for (int c = 0; c < s.cin; ++c) {
          const size_t dst = ((static_cast<size_t>(kh / 4) * 3 + kw / 4) * 48 + ((kh % 4) * 4 + kw % 4) * 3 + c) * s.cout;
          const size_t src = ((static_cast<size_t>(kh) * s.kw + kw) * s.cin + c) * s.cout;
          for (int j = 0; j < s.cout; ++j) w_s2d[dst + j] = w_host[src + j];
        }
    w_src = w_s2d.data();
  }
  double* w_dev = nullptr;
  DLC_CUDA(cudaMalloc(reinterpret_cast<void**>(&w_dev), wbytes));
  int rc = DLC_OK;
  cudaError_t e = cudaMemcpy(w_dev, w_src, wbytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_cnnvtl_set_conv: H2D copy failed: %s", cudaGetErrorString(e));
  if (rc == DLC_OK) {
    const int64_t total = static_cast<int64_t>(s.cout) * g.k_ld;
    pack_conv_weight_kernel<<<static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16)), 256>>>(
        w_dev, taps, g.vcin, s.cout, g.c_pad, static_cast<__half*>(h->w_hi[layer]),
        static_cast<__half*>(h->w_lo[layer]), g.k_ld);
    std::vector<float> b(s.cout);
    for (int i = 0; i < s.cout; ++i) b[i] = static_cast<float>(b_host[i]);
    e = cudaMemcpy(h->bias[layer], b.data(), sizeof(float) * b.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_cnnvtl_set_conv: %s", cudaGetErrorString(e));
  }
  cudaFree(w_dev);
  if (rc == DLC_OK) h->is_set[layer] = true;
  return rc;
}

extern "C" int dlc_cnnvtl_set_keep_cols(dlc_cnnvtl* h, const int64_t* keep_cols_host, int M) {
  DLC_CHECK_ARG(h && keep_cols_host && M > 0);
  for (int i = 0; i < M; ++i) {
    const int64_t c = keep_cols_host[i];
    if (c < 0 || c >= h->total_cols)
      return fail(DLC_EINVAL, "dlc_cnnvtl_set_keep_cols: column %lld out of range", static_cast<long long>(c));
    if (i > 0 && c <= keep_cols_host[i - 1])
      return fail(DLC_EINVAL, "dlc_cnnvtl_set_keep_cols: columns must be strictly increasing");
  }
  if (h->keep_cols) cudaFree(h->keep_cols);
  h->keep_cols = nullptr;
  DLC_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->keep_cols), sizeof(int64_t) * M));
  DLC_CUDA(cudaMemcpy(h->keep_cols, keep_cols_host, sizeof(int64_t) * M, cudaMemcpyHostToDevice));
  h->M = M;
  return DLC_OK;
}

namespace {

// Workspace carve-up for n images (every buffer 256-byte aligned).
struct ConvWs {
  size_t a1[2];                 // space-to-depth image planes (conv1's input)
  size_t act[kConvLayers][2];   // conv output planes, hi and lo (the descriptor tail gathers from them)
  size_t pool[kConvLayers][2];  // pooled planes (after conv1, conv2)
  size_t mm, total;
};

ConvWs carve(const dlc_cnnvtl* h, int n) {
  ConvWs w{};
  size_t off = 0;
  const int planes = h->precision == DLC_PREC_FP16X2 ? 2 : 1;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += align_up(bytes, 256);
    return at;
  };
  // TMA boxes of the last M tile may start inside the buffer and run past its logical end only in the zero-filled
  // out-of-bounds sense (the tensor maps carry the true extents), so no slack rows are needed.
  for (int p = 0; p < 2; ++p)
    w.a1[p] = p < planes ? take(static_cast<size_t>(n) * h->geo[0].H * h->geo[0].W * h->geo[0].in_ld * 2) : 0;
  for (int l = 0; l < kConvLayers; ++l) {
    const ConvGeo& g = h->geo[l];
    for (int p = 0; p < 2; ++p) {
      w.act[l][p] = take(static_cast<size_t>(n) * g.OH * g.OW * g.out_ld * 2);
      w.pool[l][p] = (p < planes && kSpec[l].pool_after) ? take(static_cast<size_t>(n) * g.PH * g.PW * g.out_ld * 2) : 0;
    }
  }
  w.mm = take(static_cast<size_t>(n) * 8);
  w.total = off;
  return w;
}

template <class Policy>
int run_conv(const dlc_cnnvtl* h, int l, int n, const void* a_hi, const void* a_lo, BiasActParams p,
             cudaStream_t stream) {
  constexpr int BK = Policy::Cfg::BK;
  constexpr bool split = Policy::Cfg::NPROD == 3;
  const ConvSpec& s = kSpec[l];
  const ConvGeo& g = h->geo[l];
  CUtensorMap ta0, ta1, tb0, tb1;
  bool ok = make_tmap_im2col_nhwc(&ta0, a_hi, 0, n, g.H, g.W, g.vcin, g.in_ld, g.vkh, g.vkw, g.pad_t, g.pad_l, g.pad_b,
                             g.pad_r, BK);
  ta1 = ta0;
  if (ok && split && a_lo)
    ok = make_tmap_im2col_nhwc(&ta1, a_lo, 0, n, g.H, g.W, g.vcin, g.in_ld, g.vkh, g.vkw, g.pad_t, g.pad_l, g.pad_b,
                               g.pad_r, BK);
  // CTA pairs (three-product policies): every conv layer is bound by L2 -> SM operand traffic on a single CTA
  // (conv1: the whole 221 KB weight set per 128-pixel tile = 8.2 GB per 1063 frames at the ~12 TB/s the L2 delivers;
  // ncu: tensor pipe 46 %); a pair stages each half of the weight tile once for 256 pixels.
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int pair_mode = g_cta_pair.load();
  const bool pairs = split && pair_mode && p.n_tile % 32 == 0 && (pair_tiles >= sm_count() / 2 || pair_mode == 2);
  const int b_rows = pairs ? p.n_tile / 2 : p.n_tile;
  if (ok) ok = make_tmap_k_major(&tb0, h->w_hi[l], 0, g.k_ld, s.cout, g.k_ld, BK, b_rows);
  tb1 = tb0;
  if (ok && split) ok = make_tmap_k_major(&tb1, h->w_lo[l], 0, g.k_ld, s.cout, g.k_ld, BK, b_rows);
  if (!ok) return fail(DLC_ECUDA, "dlc_cnnvtl_forward: tensor map encoding failed for conv%d", l + 1);
  p.k_blocks = g.k_ld / BK;
  // a short K range (conv1: 576) is accumulated in one TMEM pass, which also enables the alternate-tile epilogue
  p.kc = g.k_ld <= 1024 ? p.k_blocks : std::max(1, g_promote_k.load() / BK);
  if (!attach_plane_store_maps(p)) return fail(DLC_ECUDA, "dlc_cnnvtl_forward: tensor map encoding failed (outputs)");
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  cudaError_t e;
  if constexpr (split) {
    if (pairs) e = launch_gemm_pair<Policy>(ta0, ta1, tb0, tb1, p, std::min(pair_tiles, sm_count() / 2), stream);
    else e = launch_gemm<Policy>(ta0, ta1, tb0, tb1, p, grid, stream);
  } else {
    e = launch_gemm<Policy>(ta0, ta1, tb0, tb1, p, grid, stream);
  }
  if (e != cudaSuccess)
    return fail(DLC_ECUDA, "dlc_cnnvtl_forward: conv%d launch failed: %s", l + 1, cudaGetErrorString(e));
  return DLC_OK;
}

}  // namespace

extern "C" size_t dlc_cnnvtl_workspace_bytes(const dlc_cnnvtl* h, int n) {
  if (!h || n <= 0) return 0;
  return carve(h, n).total + 256;
}

extern "C" int dlc_cnnvtl_forward(dlc_cnnvtl* h, const void* x_dev, int x_dtype, int n, int8_t* out_dev,
                                  float* const* seg_f32_dev_host, void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(h && x_dev && ws_dev);
  DLC_CHECK_ARG(n > 0);
  DLC_CHECK_ARG(x_dtype == DLC_U8 || x_dtype == DLC_F32 || x_dtype == DLC_F64);
  DLC_CHECK_ARG(out_dev || seg_f32_dev_host);
  for (int l = 0; l < kConvLayers; ++l)
    if (!h->is_set[l]) return fail(DLC_EINVAL, "dlc_cnnvtl_forward: conv%d has no weights (dlc_cnnvtl_set_conv)", l + 1);
  if (out_dev && !h->keep_cols) return fail(DLC_EINVAL, "dlc_cnnvtl_forward: no kept columns (dlc_cnnvtl_set_keep_cols)");
  if (static_cast<int64_t>(n) * h->geo[0].OH * h->geo[0].OW > 0x7fffffff / 2)
    return fail(DLC_EINVAL, "dlc_cnnvtl_forward: too many images in one call (%d); split the batch", n);
  const ConvWs w = carve(h, n);
  if (ws_bytes < w.total)
    return fail(DLC_ENOMEM, "dlc_cnnvtl_forward: workspace of %zu bytes needed, %zu given", w.total + 256, ws_bytes);
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0);
  cudaStream_t s = as_stream(stream);
  const bool split = h->precision == DLC_PREC_FP16X2;
  char* base = static_cast<char*>(ws_dev);
  auto at = [&](size_t off) { return static_cast<void*>(base + off); };
  int* mm = reinterpret_cast<int*>(at(w.mm));

  minmax_init_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, mm);
  // space-to-depth image planes for conv1; 8-bit pixels are exact in fp16, so their residual plane is skipped
  const bool a_lo_zero = x_dtype == DLC_U8;
  {
    const ConvGeo& g = h->geo[0];
    const int64_t total = static_cast<int64_t>(n) * g.H * g.W * 4;
    const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 32));
    __half* o_hi = static_cast<__half*>(at(w.a1[0]));
    __half* o_lo = (split && !a_lo_zero) ? static_cast<__half*>(at(w.a1[1])) : nullptr;
    if (x_dtype == DLC_U8)
      s2d4_planes_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(x_dev), n, h->H, h->W, g.H, g.W,
                                                       o_hi, o_lo, g.in_ld);
    else if (x_dtype == DLC_F32)
      s2d4_planes_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x_dev), n, h->H, h->W, g.H, g.W, o_hi,
                                                     o_lo, g.in_ld);
    else
      s2d4_planes_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(x_dev), n, h->H, h->W, g.H, g.W, o_hi,
                                                      o_lo, g.in_ld);
    DLC_CUDA(cudaGetLastError());
  }
  const void* in_hi = at(w.a1[0]);
  const void* in_lo = (split && !a_lo_zero) ? at(w.a1[1]) : nullptr;
  for (int l = 0; l < kConvLayers; ++l) {
    const ConvSpec& sp = kSpec[l];
    const ConvGeo& g = h->geo[l];
    const bool last = l + 1 == kConvLayers;
    BiasActParams p{};
    int n_tile = 0;
    for (int cand = 256; cand >= 32; cand -= 32)
      if (sp.cout % cand == 0) {
        n_tile = cand;
        break;
      }
    p.n_tile = n_tile;
    p.ab_fmt = 0;
    p.M = n * g.OH * g.OW;
    p.N = sp.cout;
    p.m_tiles = ceil_div(p.M, kTileM);
    p.n_tiles = sp.cout / n_tile;
    p.bias = h->bias[l];
    p.act = sp.relu ? DLC_ACT_RELU : DLC_ACT_NONE;
    p.out_f32 = seg_f32_dev_host ? seg_f32_dev_host[l] : nullptr;
    p.out_ld = sp.cout;
    p.out_hi = at(w.act[l][0]);
    p.out_lo = at(w.act[l][1]);  // written in both precision modes: the descriptor tail reads hi + lo
    p.out_plane_ld = g.out_ld;
    p.cv_implicit = 1;
    p.cv_a_lo_zero = (l == 0 && a_lo_zero) ? 1 : 0;
    p.cv_ohw = g.OH * g.OW;
    p.cv_ow = g.OW;
    p.cv_pad_t = g.pad_t;
    p.cv_pad_l = g.pad_l;
    p.cv_kw = g.vkw;
    p.cv_cblocks = g.c_pad / g.bk;
    p.mm = out_dev ? mm : nullptr;
    p.dbg = gemm_debug_flags();
    int rc;
    if (split && g.bk == 64) rc = run_conv<BiasActPolicy<64, 3, true>>(h, l, n, in_hi, in_lo, p, s);
    else if (split) rc = run_conv<BiasActPolicy<32, 3, true>>(h, l, n, in_hi, in_lo, p, s);
    else if (g.bk == 32) rc = run_conv<BiasActPolicy<32, 1, true>>(h, l, n, in_hi, in_lo, p, s);
    else rc = run_conv<BiasActPolicy<64, 1, true>>(h, l, n, in_hi, in_lo, p, s);
    if (rc != DLC_OK) return rc;
    in_hi = p.out_hi;
    in_lo = split ? p.out_lo : nullptr;
    if (sp.pool_after) {
      __half* o_hi = static_cast<__half*>(at(w.pool[l][0]));
      __half* o_lo = split ? static_cast<__half*>(at(w.pool[l][1])) : nullptr;
      const int64_t total = static_cast<int64_t>(n) * g.PH * g.PW * (sp.cout / 8);
      const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 32));
      maxpool_planes_kernel<<<grid, 256, 0, s>>>(static_cast<const __half*>(in_hi), static_cast<const __half*>(in_lo),
                                                 n, g.OH, g.OW, sp.cout, g.out_ld, 3, 2, g.PH, g.PW, o_hi, o_lo);
      DLC_CUDA(cudaGetLastError());
      in_hi = o_hi;
      in_lo = o_lo;
    }
  }
  if (out_dev) {
    PlaneSegTable st{};
    st.n_seg = kConvLayers;
    for (int l = 0; l < kConvLayers; ++l) {
      st.hi[l] = static_cast<const __half*>(at(w.act[l][0]));
      st.lo[l] = static_cast<const __half*>(at(w.act[l][1]));
      st.start[l] = h->seg_start[l];
      st.cout[l] = kSpec[l].cout;
      st.ld[l] = h->geo[l].out_ld;
      st.pix[l] = h->geo[l].OH * h->geo[l].OW;
    }
    const int64_t total = static_cast<int64_t>(n) * h->M;
    quantise_gather_planes_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(st, n, h->keep_cols, h->M, mm,
                                                                                        out_dev);
    DLC_CUDA(cudaGetLastError());
  }
  return DLC_OK;
}

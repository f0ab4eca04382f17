// a9/a10: Hamming matrix of the int8 cnn_vtl descriptors. Replaces DistanceCalculator.calculate_distance /
// _bitwise_diff (src/cnn_vtl/similarity/DistanceCalculator.py:4-12) and the N x N Python loop around it
// (src/cnn_vtl/create_distance_matrix.py:31-36). Integer work, exact.
//
// Reference semantics: per element bin(a ^ b).count('1') on numpy int8 values. The XOR of two int8 is an int8;
// bin() of a negative number is '-0b' + bin(|x|), so the count is popcount(|int8(a ^ b)|) (|-128| = 128 -> 1).
// With 4 descriptor bytes per 32-bit word that is __popc(__vabs4(a ^ b)).
#include "ptx.cuh"
#include "util.h"

namespace dlc {

constexpr int kHamTile = 64;   // output tile edge per CTA
constexpr int kHamKWords = 32; // 32 words = 128 descriptor bytes per smem step

// desc words: [N, Mw] (M padded with zeros to a multiple of 4 by the wrapper copy kernel)
template <bool QUIRK>
__global__ void __launch_bounds__(256)
hamming_kernel(const uint32_t* __restrict__ dw, int N, int Mw, int32_t* __restrict__ D) {
  __shared__ uint32_t sa[kHamTile][kHamKWords + 1];
  __shared__ uint32_t sb[kHamTile][kHamKWords + 1];
  const int i0 = blockIdx.y * kHamTile, j0 = blockIdx.x * kHamTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  int acc[4][4] = {};
  for (int k0 = 0; k0 < Mw; k0 += kHamKWords) {
    for (int t = threadIdx.x; t < kHamTile * kHamKWords; t += 256) {
      const int r = t / kHamKWords, c = t % kHamKWords;
      const int k = k0 + c;
      sa[r][c] = (i0 + r < N && k < Mw) ? dw[static_cast<int64_t>(i0 + r) * Mw + k] : 0u;
      sb[r][c] = (j0 + r < N && k < Mw) ? dw[static_cast<int64_t>(j0 + r) * Mw + k] : 0u;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < kHamKWords; ++c) {
      uint32_t a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = sa[ty + 16 * q][c];
        b[q] = sb[tx + 16 * q][c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t x = a[u] ^ b[v];
          acc[u][v] += __popc(QUIRK ? __vabs4(x) : x);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
      if (i < N && j < N) D[static_cast<int64_t>(i) * N + j] = acc[u][v];
    }
}

// int8 [N, M] -> words [N, Mw] with zero tail (a ^ b = 0 contributes no bits under either semantics)
__global__ void hamming_pack_kernel(const int8_t* __restrict__ d, int N, int M, int Mw, uint32_t* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(N) * Mw;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(t / Mw), w = static_cast<int>(t % Mw);
    uint32_t x = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = w * 4 + q;
      if (c < M) x |= static_cast<uint32_t>(static_cast<uint8_t>(d[static_cast<int64_t>(r) * M + c])) << (8 * q);
    }
    out[t] = x;
  }
}

}  // namespace dlc

using namespace dlc;

extern "C" size_t dlc_hamming_workspace_bytes(int N, int M) {
  if (N <= 0 || M <= 0) return 0;
  return sizeof(uint32_t) * static_cast<size_t>(N) * ((M + 3) / 4);
}

extern "C" int dlc_hamming_matrix(const int8_t* desc_dev, int N, int M, int signed_bin_quirk, int32_t* D_dev,
                                  void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(N >= 0 && M > 0);
  if (N == 0) return DLC_OK;
  DLC_CHECK_ARG(desc_dev && D_dev);
  if (!ws_dev || ws_bytes < dlc_hamming_workspace_bytes(N, M) || (reinterpret_cast<uintptr_t>(ws_dev) & 3))
    return fail(DLC_ENOMEM, "dlc_hamming_matrix: 4-byte aligned workspace of %zu bytes needed, %zu given",
                dlc_hamming_workspace_bytes(N, M), ws_bytes);
  cudaStream_t s = as_stream(stream);
  const int Mw = (M + 3) / 4;
  uint32_t* words = static_cast<uint32_t*>(ws_dev);
  const int64_t total = static_cast<int64_t>(N) * Mw;
  hamming_pack_kernel<<<static_cast<int>(std::min<int64_t>((total + 255) / 256, 4096)), 256, 0, s>>>(desc_dev, N, M,
                                                                                                      Mw, words);
  dim3 grid(ceil_div(N, kHamTile), ceil_div(N, kHamTile));
  if (signed_bin_quirk) hamming_kernel<true><<<grid, 256, 0, s>>>(words, N, Mw, D_dev);
  else hamming_kernel<false><<<grid, 256, 0, s>>>(words, N, Mw, D_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

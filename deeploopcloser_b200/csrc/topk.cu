// Row-wise candidate selection: per-row top-k of a dense score matrix (loop candidates from the SDAV score
// matrix) and the deterministic k-way merge of partial top-k lists (per-CTA partials of the matcher, per-rank
// partials of the sharded matcher). New capability (north star); nearest reference analogue is the first-minimum
// np.argmin of src/sdav/similarity/SimilarityCalculator.py:33-35.
// Order: best score first; ties -> lowest reported index (so results do not depend on how the work was split).
// HBM/L2-bound integer/compare work: one warp per row, k selection passes with warp-shuffle arg-reduction.
#include <math.h>

#include <algorithm>

#include "topk.h"
#include "util.h"

namespace dlc {

struct Cand {
  float s;
  int64_t i;
};

// strict "a is better than b"
template <bool LARGEST>
__device__ __forceinline__ bool better(float as, int64_t ai, float bs, int64_t bi) {
  if (LARGEST) return as > bs || (as == bs && ai < bi);
  return as < bs || (as == bs && ai < bi);
}

template <bool LARGEST>
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ scores, const int64_t* __restrict__ cand_idx, int rows, int cols, int ld,
                 int k, int exclude_band, const float* __restrict__ row_add, float scale,
                 float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int row = warp;
  const float* srow = scores + static_cast<int64_t>(row) * ld;
  const int64_t* irow = cand_idx ? cand_idx + static_cast<int64_t>(row) * ld : nullptr;
  const float worst = LARGEST ? -INFINITY : INFINITY;
  float prev_s = LARGEST ? INFINITY : -INFINITY;
  int64_t prev_i = -2;  // (prev_s, prev_i) is better than every real candidate before the first pass
  bool have_prev = false;
  for (int sel = 0; sel < k; ++sel) {
    float bs = worst;
    int64_t bi = INT64_MAX;
    bool found = false;
    for (int c = lane; c < cols; c += 32) {
      if (exclude_band >= 0) {
        const int d = c - row;
        if ((d < 0 ? -d : d) <= exclude_band) continue;
      }
      const float s = srow[c];
      if (s != s) continue;  // NaN never selected
      const int64_t id = irow ? irow[c] : static_cast<int64_t>(c);
      if (id < 0) continue;  // padding entry of a partial list
      if (have_prev && !better<LARGEST>(prev_s, prev_i, s, id)) continue;  // already emitted (or equal to it)
      if (!found || better<LARGEST>(s, id, bs, bi)) {
        bs = s;
        bi = id;
        found = true;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      const int of = __shfl_xor_sync(0xffffffffu, static_cast<int>(found), off);
      if (of && (!found || better<LARGEST>(os, oi, bs, bi))) {
        bs = os;
        bi = oi;
        found = true;
      }
    }
    if (lane == 0) {
      const int64_t o = static_cast<int64_t>(row) * k + sel;
      if (found) {
        out_scores[o] = (row_add ? row_add[row] : 0.0f) + scale * bs;
        out_idx[o] = bi;
      } else {
        out_scores[o] = worst;
        out_idx[o] = -1;
      }
    }
    if (!found) {  // pad the rest of the row
      if (lane == 0)
        for (int t = sel + 1; t < k; ++t) {
          out_scores[static_cast<int64_t>(row) * k + t] = worst;
          out_idx[static_cast<int64_t>(row) * k + t] = -1;
        }
      break;
    }
    prev_s = bs;
    prev_i = bi;
    have_prev = true;
  }
}

// Same selection for rows of at most 32 * kRegCols columns without an index map: the row is read ONCE into registers
// (lane l holds columns l, l + 32, ...) and the k selection passes run on registers.
constexpr int kRegCols = 40;
template <bool LARGEST>
__global__ void __launch_bounds__(128)
topk_rows_reg_kernel(const float* __restrict__ scores, int rows, int cols, int ld, int k, int exclude_band,
                     float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* srow = scores + static_cast<int64_t>(row) * ld;
  const float worst = LARGEST ? -INFINITY : INFINITY;
  float v[kRegCols];
#pragma unroll
  for (int t = 0; t < kRegCols; ++t) {
    const int c = lane + 32 * t;
    float s = worst;
    bool ok = c < cols;
    if (ok && exclude_band >= 0) {
      const int d = c - row;
      ok = (d < 0 ? -d : d) > exclude_band;
    }
    if (ok) s = srow[c];
    // entries that can never be selected are parked at NaN: excluded, out of range, or NaN in the input
    v[t] = (ok && s == s) ? s : __int_as_float(0x7fc00000);
  }
  for (int sel = 0; sel < k; ++sel) {
    float bs = worst;
    int bi = 0x7fffffff;
    bool found = false;
#pragma unroll
    for (int t = 0; t < kRegCols; ++t) {
      const float s = v[t];
      if (s == s && (!found || (LARGEST ? s > bs : s < bs))) {  // columns ascend with t: strict keeps the lowest
        bs = s;
        bi = lane + 32 * t;
        found = true;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      const int of = __shfl_xor_sync(0xffffffffu, static_cast<int>(found), off);
      if (of && (!found || better<LARGEST>(os, oi, bs, bi))) {
        bs = os;
        bi = oi;
        found = true;
      }
    }
    const int64_t o = static_cast<int64_t>(row) * k + sel;
    if (!found) {
      if (lane == 0)
        for (int t = sel; t < k; ++t) {
          out_scores[static_cast<int64_t>(row) * k + t] = worst;
          out_idx[static_cast<int64_t>(row) * k + t] = -1;
        }
      break;
    }
    if (lane == 0) {
      out_scores[o] = bs;
      out_idx[o] = bi;
    }
    // retire the winner in the lane that holds it
    if ((bi & 31) == lane) {
      const int tt = bi >> 5;
#pragma unroll
      for (int t = 0; t < kRegCols; ++t)
        if (t == tt) v[t] = __int_as_float(0x7fc00000);
    }
  }
}

// Merge of per-CTA partial lists (the matcher: up to 148 lists of 16 / 32 entries per query): ONE CTA per query row
// reads the row's candidates once into registers, then k selection rounds run on-chip (thread-local best -> warp
// shuffle -> 8-way shared-memory reduction). The warp-per-row kernel above re-reads the candidates from L2 in each of
// its k passes, which left a small query batch (32 rows = 32 warps on the whole GPU) latency-bound: ~0.1 ms of a
// 1.5 ms database sweep.
constexpr int kMergeThreads = 256;
constexpr int kMergePerThread = 20;   // up to 5120 candidates per row
template <bool LARGEST>
__global__ void __launch_bounds__(kMergeThreads)
merge_partials_kernel(const float* __restrict__ scores, const int64_t* __restrict__ cand_idx, int cols, int ld, int k,
                      const float* __restrict__ row_add, float scale, float* __restrict__ out_scores,
                      int64_t* __restrict__ out_idx) {
  __shared__ float s_s[kMergeThreads / 32];
  __shared__ int64_t s_i[kMergeThreads / 32];
  __shared__ float w_s;
  __shared__ int64_t w_i;
  const int row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* srow = scores + static_cast<int64_t>(row) * ld;
  const int64_t* irow = cand_idx + static_cast<int64_t>(row) * ld;
  const float worst = LARGEST ? -INFINITY : INFINITY;
  float sc[kMergePerThread];
  int64_t ix[kMergePerThread];
#pragma unroll
  for (int t = 0; t < kMergePerThread; ++t) {
    const int c = threadIdx.x + kMergeThreads * t;
    sc[t] = worst;
    ix[t] = -1;
    if (c < cols) {
      const float s = srow[c];
      const int64_t id = irow[c];
      if (s == s && id >= 0) {   // NaN / padding never selected
        sc[t] = s;
        ix[t] = id;
      }
    }
  }
  for (int sel = 0; sel < k; ++sel) {
    float bs = worst;
    int64_t bi = -1;
#pragma unroll
    for (int t = 0; t < kMergePerThread; ++t)
      if (ix[t] >= 0 && (bi < 0 || better<LARGEST>(sc[t], ix[t], bs, bi))) {
        bs = sc[t];
        bi = ix[t];
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oi >= 0 && (bi < 0 || better<LARGEST>(os, oi, bs, bi))) {
        bs = os;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_s[warp] = bs;
      s_i[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fs = s_s[0];
      int64_t fi = s_i[0];
      for (int w = 1; w < kMergeThreads / 32; ++w)
        if (s_i[w] >= 0 && (fi < 0 || better<LARGEST>(s_s[w], s_i[w], fs, fi))) {
          fs = s_s[w];
          fi = s_i[w];
        }
      w_s = fs;
      w_i = fi;
      const int64_t o = static_cast<int64_t>(row) * k + sel;
      out_scores[o] = fi >= 0 ? (row_add ? row_add[row] : 0.0f) + scale * fs : worst;
      out_idx[o] = fi;
    }
    __syncthreads();
    const int64_t win = w_i;
    if (win < 0) {   // fewer than k candidates: pad the rest of the row
      if (threadIdx.x == 0)
        for (int t = sel + 1; t < k; ++t) {
          out_scores[static_cast<int64_t>(row) * k + t] = worst;
          out_idx[static_cast<int64_t>(row) * k + t] = -1;
        }
      return;
    }
#pragma unroll
    for (int t = 0; t < kMergePerThread; ++t)
      if (ix[t] == win) ix[t] = -1;   // reported indices are unique: the winner leaves the pool
  }
}

int merge_partials_impl(const float* scores, const int64_t* cand_idx, int rows, int cols, int ld, int k, int largest,
                        const float* row_add, float scale, float* out_scores, int64_t* out_idx, cudaStream_t stream) {
  if (rows == 0) return DLC_OK;
  if (cols > kMergeThreads * kMergePerThread)
    return topk_rows_impl(scores, cand_idx, rows, cols, ld, k, largest, -1, row_add, scale, out_scores, out_idx, stream);
  if (largest)
    merge_partials_kernel<true><<<rows, kMergeThreads, 0, stream>>>(scores, cand_idx, cols, ld, k, row_add, scale,
                                                                    out_scores, out_idx);
  else
    merge_partials_kernel<false><<<rows, kMergeThreads, 0, stream>>>(scores, cand_idx, cols, ld, k, row_add, scale,
                                                                     out_scores, out_idx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_match: merge launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

int topk_rows_impl(const float* scores, const int64_t* cand_idx, int rows, int cols, int ld, int k, int largest,
                   int exclude_band, const float* row_add, float scale, float* out_scores, int64_t* out_idx,
                   cudaStream_t stream) {
  if (rows == 0) return DLC_OK;
  if (!cand_idx && !row_add && scale == 1.0f && cols <= 32 * kRegCols) {  // the loop-candidate case: S is [N, N]
    const int grid = ceil_div(rows, 4);
    if (largest)
      topk_rows_reg_kernel<true><<<grid, 128, 0, stream>>>(scores, rows, cols, ld, k, exclude_band, out_scores, out_idx);
    else
      topk_rows_reg_kernel<false><<<grid, 128, 0, stream>>>(scores, rows, cols, ld, k, exclude_band, out_scores, out_idx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_topk_rows: launch failed: %s", cudaGetErrorString(e));
    return DLC_OK;
  }
  const int block = 256;
  const int grid = ceil_div(rows, block / 32);
  if (largest)
    topk_rows_kernel<true><<<grid, block, 0, stream>>>(scores, cand_idx, rows, cols, ld, k, exclude_band, row_add,
                                                       scale, out_scores, out_idx);
  else
    topk_rows_kernel<false><<<grid, block, 0, stream>>>(scores, cand_idx, rows, cols, ld, k, exclude_band, row_add,
                                                        scale, out_scores, out_idx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_topk_rows: launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

}  // namespace dlc

using namespace dlc;

extern "C" int dlc_topk_rows(const float* scores_dev, const int64_t* cand_idx_dev, int rows, int cols, int ld, int k,
                             int largest, int exclude_band, float* out_scores_dev, int64_t* out_idx_dev,
                             void* stream) {
  DLC_CHECK_ARG(scores_dev && out_scores_dev && out_idx_dev);
  DLC_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols);
  DLC_CHECK_ARG(k >= 1);
  return topk_rows_impl(scores_dev, cand_idx_dev, rows, cols, ld, k, largest, exclude_band, nullptr, 1.0f,
                        out_scores_dev, out_idx_dev, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
// Frame-level descriptor for the global matcher: mean over the `group_rows` patch descriptors of a frame
// ([groups*group_rows, cols] -> [groups, cols]). New definition (north star config 5), flagged in DESIGN.md: the
// reference never forms a single descriptor per frame. Bytes-bound, coalesced over columns.
namespace dlc {
__global__ void __launch_bounds__(256)
mean_pool_rows_kernel(const float* __restrict__ x, int groups, int group_rows, int cols, float* __restrict__ out) {
  const int g = blockIdx.y;
  const float inv = 1.0f / static_cast<float>(group_rows);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
    const float* p = x + static_cast<int64_t>(g) * group_rows * cols + c;
    float acc = 0.0f;
    for (int r = 0; r < group_rows; ++r) acc += p[static_cast<int64_t>(r) * cols];
    out[static_cast<int64_t>(g) * cols + c] = acc * inv;
  }
  (void)groups;
}
}  // namespace dlc

extern "C" int dlc_mean_pool_rows(const float* x_dev, int groups, int group_rows, int cols, float* out_dev,
                                  void* stream) {
  DLC_CHECK_ARG(x_dev && out_dev);
  DLC_CHECK_ARG(groups >= 0 && groups <= 65535 && group_rows >= 1 && cols >= 1);
  if (groups == 0) return DLC_OK;
  dim3 grid(std::min(ceil_div(cols, 256), 16), groups);
  mean_pool_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(x_dev, groups, group_rows, cols, out_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

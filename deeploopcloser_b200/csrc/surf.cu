// f3: the keypoint detector in front of the patch gather - get_top_n_key_points
// (src/sdav/input/CvInputParser.py:36-46: cv2.xfeatures2d.SURF_create().detect, sort by -response, first n).
// The arithmetic of that call lives in opencv-contrib 3.4.2's non-free xfeatures2d module, which is neither in the
// reference tree nor in this image: PARITY UNPINNED. This file implements the published fast-Hessian detector
// (Bay et al., CVIU 2008) with that implementation's documented constants; oracle/surf.py is the same algorithm in
// NumPy and the two agree bit for bit (every float operation below is an explicitly rounded intrinsic, no FMA
// contraction, same operation order as the oracle).
//
// One batch of frames per call:
//   integral image (row scan, then column scan by strips)                      int32 [B, H+1, W+1]
//   per octave, ONE kernel: box-filter Hessian determinant of its n_layers + 2 layers on a 32 x 32-sample tile (+ halo)
//               into shared memory - the layers never exist in HBM - then the strict 3x3x3 maxima of the middle
//               layers above the threshold, quadratic interpolation, orientation-window test -> appended to a
//               per-frame candidate list
//   per frame : n best candidates by (response desc, detection order asc) -> xy [B, n, 2]
// Cost model: 32 integral-image lookups (L1 hits) per sample and layer, 5 x 1.33 layer-samples per pixel: the path is
// bound by load instructions, not by HBM (the compulsory traffic is 1 B read + 8 B written/read per pixel).
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>

#include "util.h"

namespace dlc {

constexpr int kSurfMaxLayers = 6;     // n_layers + 2 per octave
constexpr int kSurfMaxOctaves = 5;
constexpr int kSurfCap = 16384;       // candidates kept per frame
constexpr int kOriRadius = 6;

struct SurfBox {
  int x1, y1, x2, y2;
  float w;
};
struct SurfLayer {
  int size, ni, nj, margin;
  SurfBox dx[3], dy[3], dxy[4];
  // integral-image BYTE offsets (4 * (y * pitch + x)) of the shared box corners, filled by surf_plan; 64-bit so that
  // a lookup address is one 64-bit add of the sample's pointer and a parameter-bank operand
  int64_t off_dx[8];    // [top | bottom][x1_0, x2_0, x2_1, x2_2]
  int64_t off_dy[8];    // [left | right][y1_0, y2_0, y2_1, y2_2]
  int64_t off_dxy[16];  // [y1_0, y2_0, y1_2, y2_2][x1_0, x2_0, x1_1, x2_1]
};
struct SurfOctave {
  int step, rows, cols, n;            // n layers of rows x cols samples
  SurfLayer layer[kSurfMaxLayers];
};

// ---------------------------------------------------------------- integral image
__global__ void __launch_bounds__(256)
surf_integral_rows_kernel(const uint8_t* __restrict__ img, int B, int H, int W, int32_t* __restrict__ sum) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * H) return;
  const int b = warp / H, r = warp - b * H;
  const uint8_t* src = img + (static_cast<int64_t>(b) * H + r) * W;
  int32_t* frame = sum + static_cast<int64_t>(b) * (H + 1) * (W + 1);
  int32_t* dst = frame + static_cast<int64_t>(r + 1) * (W + 1);
  if (r == 0)
    for (int c = lane; c <= W; c += 32) frame[c] = 0;
  if (lane == 0) dst[0] = 0;
  int carry = 0;
  for (int c0 = 0; c0 < W; c0 += 32) {
    const int c = c0 + lane;
    int v = c < W ? src[c] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, off);
      if (lane >= off) v += t;
    }
    if (c < W) dst[c + 1] = carry + v;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
}
// strip of 32 columns x all rows per CTA; warp w owns a block of consecutive rows, lane = column
__global__ void __launch_bounds__(1024)
surf_integral_cols_kernel(int B, int H, int W, int32_t* __restrict__ sum) {
  __shared__ int s_tot[32][33];
  const int strips = (W + 31) / 32;
  const int b = blockIdx.x / strips, strip = blockIdx.x - b * strips;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = 1 + strip * 32 + lane;
  const int per = (H + 31) / 32;
  const int r0 = 1 + w * per, r1 = min(H + 1, r0 + per);
  int32_t* frame = sum + static_cast<int64_t>(b) * (H + 1) * (W + 1);
  int acc = 0;
  if (c <= W)
    for (int r = r0; r < r1; ++r) acc += frame[static_cast<int64_t>(r) * (W + 1) + c];
  s_tot[w][lane] = acc;
  __syncthreads();
  int off = 0;
  for (int k = 0; k < w; ++k) off += s_tot[k][lane];
  if (c <= W)
    for (int r = r0; r < r1; ++r) {
      const int64_t i = static_cast<int64_t>(r) * (W + 1) + c;
      off += frame[i];
      frame[i] = off;
    }
}

// ---------------------------------------------------------------- Hessian determinant
// Box sums of one pattern from the integral image. The three Dxx boxes share their rows and abut in x (and the Dyy
// boxes transposed), the four Dxy boxes lie on a 4 x 4 grid of corners: 8 + 8 + 16 = 32 loads per sample instead of
// 40 (surf_plan checks the structure). Every box sum is formed exactly in int32, then float(box) * w is added in
// float64 in box order and rounded to float32 - the oracle's arithmetic.
__device__ __forceinline__ float haar_acc(const int (&box)[4], const float (&w)[4], int n) {
  double d = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n) d = __dadd_rn(d, static_cast<double>(__fmul_rn(__int2float_rn(box[k]), w[k])));
  return __double2float_rn(d);
}
__device__ __forceinline__ float det_at(const char* __restrict__ org, const SurfLayer& L) {
#define SURF_AT(off) __ldg(reinterpret_cast<const int32_t*>(org + (off)))
  int box[4];
  float w[4];
  {
    int top[4], bot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      top[q] = SURF_AT(L.off_dx[q]);
      bot[q] = SURF_AT(L.off_dx[4 + q]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      box[k] = (bot[k + 1] - bot[k]) - (top[k + 1] - top[k]);
      w[k] = L.dx[k].w;
    }
  }
  const float dx = haar_acc(box, w, 3);
  {
    int lft[4], rgt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      lft[q] = SURF_AT(L.off_dy[q]);
      rgt[q] = SURF_AT(L.off_dy[4 + q]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      box[k] = (rgt[k + 1] - rgt[k]) - (lft[k + 1] - lft[k]);
      w[k] = L.dy[k].w;
    }
  }
  const float dy = haar_acc(box, w, 3);
  {
    int g[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) g[r][c] = SURF_AT(L.off_dxy[r * 4 + c]);
    box[0] = (g[1][1] - g[1][0]) - (g[0][1] - g[0][0]);
    box[1] = (g[1][3] - g[1][2]) - (g[0][3] - g[0][2]);
    box[2] = (g[3][1] - g[3][0]) - (g[2][1] - g[2][0]);
    box[3] = (g[3][3] - g[3][2]) - (g[2][3] - g[2][2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = L.dxy[k].w;
  }
  const float dxy = haar_acc(box, w, 4);
#undef SURF_AT
  return __fsub_rn(__fmul_rn(dx, dy), __fmul_rn(__fmul_rn(0.81f, dxy), dxy));
}

// ---------------------------------------------------------------- maxima
// Gaussian elimination with partial pivoting, float64, the oracle's operation order (solve3 in oracle/surf.py)
__device__ __forceinline__ bool solve3(double (&M)[3][4], double (&x)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int p = k;
#pragma unroll
    for (int r = k + 1; r < 3; ++r)
      if (fabs(M[r][k]) > fabs(M[p][k])) p = r;
    if (M[p][k] == 0.0) return false;
    if (p != k) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double t = M[k][c];
        M[k][c] = M[p][c];
        M[p][c] = t;
      }
    }
#pragma unroll
    for (int r = k + 1; r < 3; ++r) {
      const double f = __ddiv_rn(M[r][k], M[k][k]);
#pragma unroll
      for (int c = k; c < 4; ++c) M[r][c] = __dsub_rn(M[r][c], __dmul_rn(f, M[k][c]));
    }
  }
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    double s = M[k][3];
#pragma unroll
    for (int c = k + 1; c < 3; ++c) s = __dsub_rn(s, __dmul_rn(M[k][c], x[c]));
    x[k] = __ddiv_rn(s, M[k][k]);
  }
  return true;
}

__device__ __forceinline__ int round_half_even(float v) { return __float2int_rn(v); }

// the orientation stage of the reference implementation drops a keypoint when no gradient sample of its radius-6s
// disc fits inside the image
__device__ bool orientation_samplable(float x, float y, float size, int H, int W) {
  const float s = __fdiv_rn(__fmul_rn(size, 1.2f), 9.0f);
  const int grad = 2 * round_half_even(__fmul_rn(2.0f, s));
  if (H + 1 < grad || W + 1 < grad) return false;
  const float half = __fdiv_rn(static_cast<float>(grad - 1), 2.0f);
  for (int i = -kOriRadius; i <= kOriRadius; ++i)
    for (int j = -kOriRadius; j <= kOriRadius; ++j)
      if (i * i + j * j <= kOriRadius * kOriRadius) {
        const int px = round_half_even(__fsub_rn(__fadd_rn(x, __fmul_rn(static_cast<float>(j), s)), half));
        const int py = round_half_even(__fsub_rn(__fadd_rn(y, __fmul_rn(static_cast<float>(i), s)), half));
        if (py >= 0 && py < H + 1 - grad && px >= 0 && px < W + 1 - grad) return true;
      }
  return false;
}

constexpr int kSurfTile = 32;
constexpr int kSurfHalo = kSurfTile + 2;
// Phase 2 of an octave kernel: the determinant layers of the tile (+ halo) are in shared memory; find the strict
// 3x3x3 maxima of the middle layers, interpolate, test the orientation window, append to the frame's candidate list.
__device__ __forceinline__ void surf_tile_maxima(const float* __restrict__ s_det, const SurfOctave& oc, int i0, int j0,
                                                 int b, int octave, float thr, int H, int W, float4* __restrict__ cand,
                                                 unsigned long long* __restrict__ keys, int* __restrict__ count) {
  constexpr int kCells = kSurfHalo * kSurfHalo;
  for (int e = threadIdx.x; e < (oc.n - 2) * kSurfTile * kSurfTile; e += blockDim.x) {
    const int l = 1 + e / (kSurfTile * kSurfTile);             // middle layers 1 .. n - 2
    const int rc = e % (kSurfTile * kSurfTile);
    const int r = 1 + rc / kSurfTile, c = 1 + rc % kSurfTile;
    const int i = i0 + r, j = j0 + c;
    const SurfLayer& L = oc.layer[l];
    const int size = L.size;
    const int margin = (oc.layer[l + 1].size / 2) / oc.step + 1;
    if (i < margin || i >= oc.rows - margin || j < margin || j >= oc.cols - margin) continue;
    const float* d1 = s_det + l * kCells + r * kSurfHalo + c;
    const float val0 = d1[0];
    if (!(val0 > thr)) continue;
    // strict maximum of the 3x3x3 neighbourhood: the 8 neighbours of the own layer first (a random sample fails
    // there 8 times out of 9), the other two layers and the interpolation stencil only for the survivors
    bool is_max = true;
#pragma unroll
    for (int di = -1; di <= 1; ++di)
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
        if (di != 0 || dj != 0) is_max = is_max && (val0 > d1[di * kSurfHalo + dj]);
    if (!is_max) continue;
    float N9[3][9];
#pragma unroll
    for (int dl = 0; dl < 3; ++dl)
#pragma unroll
      for (int di = 0; di < 3; ++di)
#pragma unroll
        for (int dj = 0; dj < 3; ++dj)
          N9[dl][di * 3 + dj] = d1[(dl - 1) * kCells + (di - 1) * kSurfHalo + (dj - 1)];
#pragma unroll
    for (int q = 0; q < 9; ++q) is_max = is_max && (val0 > N9[0][q]) && (val0 > N9[2][q]);
    if (!is_max) continue;
    const int sum_i = oc.step * (i - (size / 2) / oc.step);
    const int sum_j = oc.step * (j - (size / 2) / oc.step);
    const float cy = __fadd_rn(static_cast<float>(sum_i), __fmul_rn(static_cast<float>(size - 1), 0.5f));
    const float cx = __fadd_rn(static_cast<float>(sum_j), __fmul_rn(static_cast<float>(size - 1), 0.5f));
    const int ds = size - oc.layer[l - 1].size;
    // negative first derivatives and the Hessian of the 3x3x3 neighbourhood (float32, as Vec3f / Matx33f)
    const float bx = __fdiv_rn(-__fsub_rn(N9[1][5], N9[1][3]), 2.0f);
    const float by = __fdiv_rn(-__fsub_rn(N9[1][7], N9[1][1]), 2.0f);
    const float bs = __fdiv_rn(-__fsub_rn(N9[2][4], N9[0][4]), 2.0f);
    const float axx = __fadd_rn(__fsub_rn(N9[1][3], __fmul_rn(2.0f, N9[1][4])), N9[1][5]);
    const float ayy = __fadd_rn(__fsub_rn(N9[1][1], __fmul_rn(2.0f, N9[1][4])), N9[1][7]);
    const float ass = __fadd_rn(__fsub_rn(N9[0][4], __fmul_rn(2.0f, N9[1][4])), N9[2][4]);
    const float axy = __fdiv_rn(__fadd_rn(__fsub_rn(__fsub_rn(N9[1][8], N9[1][6]), N9[1][2]), N9[1][0]), 4.0f);
    const float axs = __fdiv_rn(__fadd_rn(__fsub_rn(__fsub_rn(N9[2][5], N9[2][3]), N9[0][5]), N9[0][3]), 4.0f);
    const float ays = __fdiv_rn(__fadd_rn(__fsub_rn(__fsub_rn(N9[2][7], N9[2][1]), N9[0][7]), N9[0][1]), 4.0f);
    double M[3][4] = {{axx, axy, axs, bx}, {axy, ayy, ays, by}, {axs, ays, ass, bs}};
    double xd[3];
    if (!solve3(M, xd)) continue;
    const float x0 = __double2float_rn(xd[0]), x1 = __double2float_rn(xd[1]), x2 = __double2float_rn(xd[2]);
    if (!((x0 != 0.0f || x1 != 0.0f || x2 != 0.0f) && fabsf(x0) <= 1.0f && fabsf(x1) <= 1.0f && fabsf(x2) <= 1.0f)) continue;
    const float px = __fadd_rn(cx, __fmul_rn(x0, static_cast<float>(oc.step)));
    const float py = __fadd_rn(cy, __fmul_rn(x1, static_cast<float>(oc.step)));
    const float ksize = rintf(__fadd_rn(static_cast<float>(size), __fmul_rn(x2, static_cast<float>(ds))));
    if (!orientation_samplable(px, py, ksize, H, W)) continue;
    const int slot = atomicAdd(count + b, 1);
    if (slot >= kSurfCap) continue;
    // order: response descending, then detection order (octave, layer, row, column) ascending
    const uint32_t order = (static_cast<uint32_t>(octave) << 29) | (static_cast<uint32_t>(l) << 26) |
                           (static_cast<uint32_t>(i) << 13) | static_cast<uint32_t>(j);
    cand[static_cast<int64_t>(b) * kSurfCap + slot] = make_float4(px, py, ksize, val0);
    keys[static_cast<int64_t>(b) * kSurfCap + slot] =
        (static_cast<unsigned long long>(__float_as_uint(val0)) << 32) | static_cast<unsigned long long>(~order);
  }
}

// One octave, fused: a CTA evaluates the determinant of all layers on a tile of kSurfTile x kSurfTile samples plus a
// one-sample halo into shared memory (the layers never go to HBM), then finds the maxima of the middle layers there.
__global__ void __launch_bounds__(256)
surf_octave_kernel(const int32_t* __restrict__ sum, int H, int W, const __grid_constant__ SurfOctave oc, int octave,
                   float thr, float4* __restrict__ cand, unsigned long long* __restrict__ keys, int* __restrict__ count) {
  extern __shared__ float s_det[];   // [oc.n][kSurfHalo][kSurfHalo]
  const int b = blockIdx.z;
  const int tiles_x = (oc.cols + kSurfTile - 1) / kSurfTile;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int i0 = ty * kSurfTile - 1, j0 = tx * kSurfTile - 1;   // sample coordinates of the halo's corner
  const int pitch = W + 1;
  const int32_t* frame = sum + static_cast<int64_t>(b) * (H + 1) * pitch;
  constexpr int kCells = kSurfHalo * kSurfHalo;
  // layer outermost and unrolled: the layer's offsets and weights are then direct parameter-bank operands
#pragma unroll
  for (int l = 0; l < kSurfMaxLayers; ++l) {
    if (l < oc.n) {
      const SurfLayer& L = oc.layer[l];
      for (int rc = threadIdx.x; rc < kCells; rc += blockDim.x) {
        const int r = rc / kSurfHalo, c = rc - r * kSurfHalo;
        const int i = i0 + r - L.margin, j = j0 + c - L.margin;  // filter origin in samples
        float v = 0.0f;                                          // the filter does not fit here
        if (i >= 0 && i < L.ni && j >= 0 && j < L.nj)
          v = det_at(reinterpret_cast<const char*>(frame + (i * pitch + j) * oc.step), L);
        s_det[l * kCells + rc] = v;
      }
    }
  }
  __syncthreads();
  surf_tile_maxima(s_det, oc, i0, j0, b, octave, thr, H, W, cand, keys, count);
}

// ---------------------------------------------------------------- octaves 0 and 1 of the default pyramid
// The generic kernel above is bound by instruction issue: every integral-image lookup costs a 64-bit address add (two
// instructions) and a global load. With the default pyramid (3 layers per octave: filter sizes (9 + 6 l) << octave)
// the box corners are compile-time constants, so this kernel first copies the integral-image region under the tile
// (+ halo + filter extent) into shared memory and then reads every corner with ONE instruction (LDS [cell + imm]).
// For step 2 the region is stored as 2 x 2 parity planes (pixel (y, x) -> plane (y & 1, x & 1), position (y >> 1,
// x >> 1)): the 32 lanes of a warp are consecutive SAMPLES, i.e. every second pixel, and a box corner has a fixed
// parity, so the lanes still read consecutive words. Box sums stay exact int32 and the float operations are the same
// intrinsics in the same order as det_at(): identical bits (tests/test_gpu_kernels.py::test_surf_detect_equals_oracle).
// dlc_surf_detect uses it when the runtime plan equals the compile-time tables (surf_fast_ok), else the generic kernel.
constexpr int kFastLayers = 5;                       // n_layers = 3 (+ 2)
constexpr int kFastMaxMargin = 16;                   // margin of the widest layer (size 33 << o): (33 / 2)
constexpr int kFastPD = 67;                          // plane dimension: ((33 + 33) * step + 1 pixels) / step, rounded up
__host__ __device__ constexpr int fast_cr(int x, int size) { return (2 * x * size + 9) / 18; }   // cvRound(size / 9.f * x): never a tie
__host__ __device__ constexpr int fast_size(int step, int l) { return (9 + 6 * l) * step; }
__host__ __device__ constexpr int fast_margin(int step, int l) { return (fast_size(step, l) / 2) / step; }
// word offset (relative to the cell's base word) of the integral-image pixel (dy, dx) of layer l's filter window
template <int STEP, int LAYER>
__device__ __forceinline__ constexpr int fast_off(int dy, int dx) {
  constexpr int add = kFastMaxMargin - fast_margin(STEP, LAYER);
  return (((dy % STEP) * STEP + (dx % STEP)) * kFastPD + (add + dy / STEP)) * kFastPD + (add + dx / STEP);
}
template <int STEP, int LAYER>
__device__ __forceinline__ float det_fast(const int32_t* __restrict__ cell, const SurfLayer& L) {
  constexpr int S = fast_size(STEP, LAYER);
  constexpr int e0 = fast_cr(0, S), e3 = fast_cr(3, S), e6 = fast_cr(6, S), e9 = fast_cr(9, S);   // lobe edges
  constexpr int b2 = fast_cr(2, S), b7 = fast_cr(7, S);                                           // band edges
  constexpr int g1 = fast_cr(1, S), g4 = fast_cr(4, S), g5 = fast_cr(5, S), g8 = fast_cr(8, S);   // Dxy grid
#define SURF_F(dy, dx) cell[fast_off<STEP, LAYER>(dy, dx)]
  int box[4];
  float w[4];
  {
    // int32 arithmetic is exact (and wraps like the reference's): regrouped as column differences, 7 subtractions
    const int d[4] = {SURF_F(b7, e0) - SURF_F(b2, e0), SURF_F(b7, e3) - SURF_F(b2, e3), SURF_F(b7, e6) - SURF_F(b2, e6),
                      SURF_F(b7, e9) - SURF_F(b2, e9)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      box[k] = d[k + 1] - d[k];
      w[k] = L.dx[k].w;
    }
  }
  const float dx = haar_acc(box, w, 3);
  {
    const int d[4] = {SURF_F(e0, b7) - SURF_F(e0, b2), SURF_F(e3, b7) - SURF_F(e3, b2), SURF_F(e6, b7) - SURF_F(e6, b2),
                      SURF_F(e9, b7) - SURF_F(e9, b2)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      box[k] = d[k + 1] - d[k];
      w[k] = L.dy[k].w;
    }
  }
  const float dy = haar_acc(box, w, 3);
  {
    const int g[4][4] = {{SURF_F(g1, g1), SURF_F(g1, g4), SURF_F(g1, g5), SURF_F(g1, g8)},
                         {SURF_F(g4, g1), SURF_F(g4, g4), SURF_F(g4, g5), SURF_F(g4, g8)},
                         {SURF_F(g5, g1), SURF_F(g5, g4), SURF_F(g5, g5), SURF_F(g5, g8)},
                         {SURF_F(g8, g1), SURF_F(g8, g4), SURF_F(g8, g5), SURF_F(g8, g8)}};
    box[0] = (g[1][1] - g[1][0]) - (g[0][1] - g[0][0]);
    box[1] = (g[1][3] - g[1][2]) - (g[0][3] - g[0][2]);
    box[2] = (g[3][1] - g[3][0]) - (g[2][1] - g[2][0]);
    box[3] = (g[3][3] - g[3][2]) - (g[2][3] - g[2][2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = L.dxy[k].w;
  }
  const float dxy = haar_acc(box, w, 4);
#undef SURF_F
  return __fsub_rn(__fmul_rn(dx, dy), __fmul_rn(__fmul_rn(0.81f, dxy), dxy));
}
template <int STEP, int LAYER>
__device__ __forceinline__ void det_fast_cell(const int32_t* __restrict__ cell, const SurfOctave& oc, int i_halo, int j_halo,
                                              float* __restrict__ out) {
  const SurfLayer& L = oc.layer[LAYER];
  const int i = i_halo - fast_margin(STEP, LAYER), j = j_halo - fast_margin(STEP, LAYER);   // filter origin (samples)
  float v = 0.0f;
  if (i >= 0 && i < L.ni && j >= 0 && j < L.nj) v = det_fast<STEP, LAYER>(cell, L);
  out[LAYER * kSurfHalo * kSurfHalo] = v;
}

template <int STEP>
__global__ void __launch_bounds__(256, 2)
surf_octave_fast_kernel(const int32_t* __restrict__ sum, int H, int W, const __grid_constant__ SurfOctave oc, int octave,
                        float thr, float4* __restrict__ cand, unsigned long long* __restrict__ keys,
                        int* __restrict__ count) {
  extern __shared__ float s_det[];   // [kFastLayers][kSurfHalo][kSurfHalo], then the integral-image planes
  constexpr int kCells = kSurfHalo * kSurfHalo;
  constexpr int kLog = STEP == 1 ? 0 : 1;
  static_assert(STEP == 1 || STEP == 2, "octaves 0 and 1");
  int32_t* s_int = reinterpret_cast<int32_t*>(s_det + kFastLayers * kCells);   // [STEP * STEP][kFastPD][kFastPD]
  const int b = blockIdx.z;
  const int tiles_x = (oc.cols + kSurfTile - 1) / kSurfTile;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int i0 = ty * kSurfTile - 1, j0 = tx * kSurfTile - 1;   // sample coordinates of the halo's corner
  const int pitch = W + 1;
  const int32_t* frame = sum + static_cast<int64_t>(b) * (H + 1) * pitch;
  // ---- the region: pixels [yb, yb + E) x [xb, xb + E), E = 66 * STEP + 1 (zeros outside the integral image)
  constexpr int E = 66 * STEP + 1;
  const int yb = (i0 - kFastMaxMargin) * STEP, xb = (j0 - kFastMaxMargin) * STEP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // two rows x all column groups per pass: 2 * kColIters independent loads in flight per thread
  constexpr int kColIters = (E + 31) / 32;
  for (int yr0 = 2 * warp; yr0 < E; yr0 += 16) {
    int v[2][kColIters];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int yr = yr0 + dy, y = yb + yr;
      const bool y_ok = yr < E && y >= 0 && y <= H;
      const int32_t* src = frame + static_cast<int64_t>(y) * pitch + xb;
#pragma unroll
      for (int k = 0; k < kColIters; ++k) {
        const int xr = k * 32 + lane, x = xb + xr;
        v[dy][k] = (y_ok && xr < E && x >= 0 && x <= W) ? __ldg(src + xr) : 0;
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int yr = yr0 + dy;
      int32_t* dst = s_int + (((yr & (STEP - 1)) * STEP) * kFastPD + (yr >> kLog)) * kFastPD;
#pragma unroll
      for (int k = 0; k < kColIters; ++k) {
        const int xr = k * 32 + lane;
        if (yr < E && xr < E) dst[(xr & (STEP - 1)) * (kFastPD * kFastPD) + (xr >> kLog)] = v[dy][k];
      }
    }
  }
  __syncthreads();
  // ---- determinants of the five layers on the 34 x 34 halo grid. A warp pass is one halo row, lanes = columns 0..31
  // (consecutive words: no bank conflict whatever the pitch); the last two columns are done with lanes = rows
  // (word stride kFastPD = 67 = 3 mod 32: conflict-free as well). 34 + 3 passes over 8 warps, 5 each at most.
  for (int it = 0; it < 5; ++it) {
    const int idx = warp + 8 * it;
    int r, c;
    if (idx < kSurfHalo) {
      r = idx;
      c = lane;
    } else {
      const int q = (idx - kSurfHalo) * 32 + lane;
      if (idx >= kSurfHalo + 3 || q >= 2 * kSurfHalo) continue;
      r = q % kSurfHalo;
      c = 32 + q / kSurfHalo;
    }
    const int32_t* cell = s_int + r * kFastPD + c;
    float* out = s_det + r * kSurfHalo + c;
    det_fast_cell<STEP, 0>(cell, oc, i0 + r, j0 + c, out);
    det_fast_cell<STEP, 1>(cell, oc, i0 + r, j0 + c, out);
    det_fast_cell<STEP, 2>(cell, oc, i0 + r, j0 + c, out);
    det_fast_cell<STEP, 3>(cell, oc, i0 + r, j0 + c, out);
    det_fast_cell<STEP, 4>(cell, oc, i0 + r, j0 + c, out);
  }
  __syncthreads();
  surf_tile_maxima(s_det, oc, i0, j0, b, octave, thr, H, W, cand, keys, count);
}
constexpr size_t surf_fast_smem(int step) {
  return sizeof(float) * kFastLayers * kSurfHalo * kSurfHalo + sizeof(int32_t) * step * step * kFastPD * kFastPD;
}

// ---------------------------------------------------------------- n best per frame
__global__ void __launch_bounds__(256)
surf_top_kernel(const float4* __restrict__ cand, unsigned long long* __restrict__ keys, const int* __restrict__ count,
                int H, int W, int top_n, float* __restrict__ xy, float* __restrict__ info, int* __restrict__ found) {
  __shared__ unsigned long long s_key[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  const int total = count[b];
  const int m = min(total, kSurfCap);
  unsigned long long* k = keys + static_cast<int64_t>(b) * kSurfCap;
  const float4* c = cand + static_cast<int64_t>(b) * kSurfCap;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int n = 0; n < top_n; ++n) {
    unsigned long long best = 0ull;   // valid keys are > 0 (response > threshold > 0)
    int bi = -1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const unsigned long long v = k[i];
      if (v > best) {
        best = v;
        bi = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > best) {   // keys are distinct (the order field is unique), no tie to break
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_key[w] = best;
      s_idx[w] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < 8; ++q)
        if (s_key[q] > best) {
          best = s_key[q];
          bi = s_idx[q];
        }
      float* o = xy + (static_cast<int64_t>(b) * top_n + n) * 2;
      float* oi = info ? info + (static_cast<int64_t>(b) * top_n + n) * 2 : nullptr;
      if (bi >= 0) {
        const float4 v = c[bi];
        o[0] = v.x;
        o[1] = v.y;
        if (oi) {
          oi[0] = v.z;
          oi[1] = v.w;
        }
        k[bi] = 0ull;  // taken
      } else {         // fewer than top_n keypoints: the image centre, size / response 0 (found[b] tells)
        o[0] = 0.5f * static_cast<float>(W - 1);
        o[1] = 0.5f * static_cast<float>(H - 1);
        if (oi) {
          oi[0] = 0.0f;
          oi[1] = 0.0f;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) found[b] = total;
}

// ---------------------------------------------------------------- host side
static inline int cv_round(float v) { return static_cast<int>(nearbyint(static_cast<double>(v))); }
static void resize_pattern(const int (*src)[5], int n, int size, SurfBox* dst) {
  const float ratio = static_cast<float>(size) / 9.0f;
  for (int k = 0; k < n; ++k) {
    const int x1 = cv_round(ratio * src[k][0]), y1 = cv_round(ratio * src[k][1]);
    const int x2 = cv_round(ratio * src[k][2]), y2 = cv_round(ratio * src[k][3]);
    dst[k] = SurfBox{x1, y1, x2, y2, src[k][4] / (static_cast<float>(x2 - x1) * static_cast<float>(y2 - y1))};
  }
}

struct SurfPlan {
  int n_octaves;
  SurfOctave oc[kSurfMaxOctaves];
  bool structure_ok;
  size_t off_sum, off_cand, off_keys, off_count, total;
};
static SurfPlan surf_plan(int B, int H, int W, int n_octaves, int n_layers) {
  static const int dx_s[3][5] = {{0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1}};
  static const int dy_s[3][5] = {{2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1}};
  static const int dxy_s[4][5] = {{1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1}};
  SurfPlan p{};
  p.n_octaves = n_octaves;
  p.structure_ok = true;
  for (int o = 0; o < n_octaves; ++o) {
    SurfOctave& oc = p.oc[o];
    oc.step = 1 << o;
    oc.rows = H / oc.step;
    oc.cols = W / oc.step;
    oc.n = n_layers + 2;
    for (int l = 0; l < oc.n; ++l) {
      SurfLayer& L = oc.layer[l];
      L.size = (9 + 6 * l) << o;
      const bool fits = L.size <= H && L.size <= W;
      L.ni = fits ? 1 + (H - L.size) / oc.step : 0;
      L.nj = fits ? 1 + (W - L.size) / oc.step : 0;
      L.margin = (L.size / 2) / oc.step;
      resize_pattern(dx_s, 3, L.size, L.dx);
      resize_pattern(dy_s, 3, L.size, L.dy);
      resize_pattern(dxy_s, 4, L.size, L.dxy);
      {
        const int pitch = W + 1;
        const int xs[4] = {L.dx[0].x1, L.dx[0].x2, L.dx[1].x2, L.dx[2].x2};
        const int ys[4] = {L.dy[0].y1, L.dy[0].y2, L.dy[1].y2, L.dy[2].y2};
        const int gx[4] = {L.dxy[0].x1, L.dxy[0].x2, L.dxy[1].x1, L.dxy[1].x2};
        const int gy[4] = {L.dxy[0].y1, L.dxy[0].y2, L.dxy[2].y1, L.dxy[2].y2};
        for (int q = 0; q < 4; ++q) {
          L.off_dx[q] = 4ll * (L.dx[0].y1 * pitch + xs[q]);
          L.off_dx[4 + q] = 4ll * (L.dx[0].y2 * pitch + xs[q]);
          L.off_dy[q] = 4ll * (ys[q] * pitch + L.dy[0].x1);
          L.off_dy[4 + q] = 4ll * (ys[q] * pitch + L.dy[0].x2);
          for (int c = 0; c < 4; ++c) L.off_dxy[q * 4 + c] = 4ll * (gy[q] * pitch + gx[c]);
        }
      }
      // the shared-corner evaluation of det_at() relies on this (it follows from scaling equal coordinates equally)
      for (int k = 0; k < 3; ++k) {
        p.structure_ok = p.structure_ok && L.dx[k].y1 == L.dx[0].y1 && L.dx[k].y2 == L.dx[0].y2 &&
                         L.dy[k].x1 == L.dy[0].x1 && L.dy[k].x2 == L.dy[0].x2 &&
                         (k == 0 || (L.dx[k].x1 == L.dx[k - 1].x2 && L.dy[k].y1 == L.dy[k - 1].y2));
      }
      p.structure_ok = p.structure_ok && L.dxy[2].x1 == L.dxy[0].x1 && L.dxy[2].x2 == L.dxy[0].x2 &&
                       L.dxy[3].x1 == L.dxy[1].x1 && L.dxy[3].x2 == L.dxy[1].x2 && L.dxy[1].y1 == L.dxy[0].y1 &&
                       L.dxy[1].y2 == L.dxy[0].y2 && L.dxy[3].y1 == L.dxy[2].y1 && L.dxy[3].y2 == L.dxy[2].y2;
    }
  }
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  p.off_sum = take(sizeof(int32_t) * static_cast<size_t>(B) * (H + 1) * (W + 1));
  p.off_cand = take(sizeof(float4) * static_cast<size_t>(B) * kSurfCap);
  p.off_keys = take(sizeof(unsigned long long) * static_cast<size_t>(B) * kSurfCap);
  p.off_count = take(sizeof(int) * static_cast<size_t>(B));
  p.total = o;
  return p;
}

// The compile-time tables of surf_octave_fast_kernel describe this octave of the runtime plan?
static bool surf_fast_ok(const SurfOctave& oc, int octave) {
  if (octave > 1 || oc.n != kFastLayers || oc.step != (1 << octave)) return false;
  for (int l = 0; l < kFastLayers; ++l) {
    const SurfLayer& L = oc.layer[l];
    const int S = fast_size(oc.step, l);
    if (L.size != S || L.margin != fast_margin(oc.step, l) || L.margin > kFastMaxMargin) return false;
    const int e[4] = {fast_cr(0, S), fast_cr(3, S), fast_cr(6, S), fast_cr(9, S)};
    const int b2 = fast_cr(2, S), b7 = fast_cr(7, S);
    const int g[4] = {fast_cr(1, S), fast_cr(4, S), fast_cr(5, S), fast_cr(8, S)};
    for (int k = 0; k < 3; ++k) {
      if (L.dx[k].x1 != e[k] || L.dx[k].x2 != e[k + 1] || L.dx[k].y1 != b2 || L.dx[k].y2 != b7) return false;
      if (L.dy[k].y1 != e[k] || L.dy[k].y2 != e[k + 1] || L.dy[k].x1 != b2 || L.dy[k].x2 != b7) return false;
    }
    for (int k = 0; k < 4; ++k) {
      const int gx = (k & 1) * 2, gy = (k >> 1) * 2;
      if (L.dxy[k].x1 != g[gx] || L.dxy[k].x2 != g[gx + 1] || L.dxy[k].y1 != g[gy] || L.dxy[k].y2 != g[gy + 1]) return false;
    }
    if (e[3] != S) return false;
  }
  return true;
}
std::atomic<int> g_surf_fast{1};  // dlc_debug_set key 10: 0 = always the generic octave kernel (developer A/B, tests)

}  // namespace dlc

using namespace dlc;

static int check_surf_args(int B, int H, int W, int n_octaves, int n_layers) {
  DLC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && H < 8192 && W < 8192);  // row / column fit the 13-bit order fields
  DLC_CHECK_ARG(n_octaves >= 1 && n_octaves <= kSurfMaxOctaves);
  DLC_CHECK_ARG(n_layers >= 1 && n_layers + 2 <= kSurfMaxLayers);
  return DLC_OK;
}

extern "C" size_t dlc_surf_workspace_bytes(int B, int H, int W, int n_octaves, int n_layers) {
  if (B <= 0 || check_surf_args(B, H, W, n_octaves, n_layers) != DLC_OK) return 0;
  return surf_plan(B, H, W, n_octaves, n_layers).total;
}

extern "C" int dlc_surf_detect(const uint8_t* img_dev, int B, int H, int W, float hessian_threshold, int n_octaves,
                               int n_layers, int top_n, float* xy_dev, float* info_dev, int32_t* found_dev,
                               void* ws_dev, size_t ws_bytes, void* stream) {
  if (int rc = check_surf_args(B, H, W, n_octaves, n_layers)) return rc;
  DLC_CHECK_ARG(top_n >= 1 && hessian_threshold > 0.0f);
  if (B == 0) return DLC_OK;
  DLC_CHECK_ARG(img_dev && xy_dev && found_dev && ws_dev);
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0);
  const SurfPlan p = surf_plan(B, H, W, n_octaves, n_layers);
  if (!p.structure_ok) return fail(DLC_EUNSUPPORTED, "dlc_surf_detect: unexpected box-filter layout");
  if (ws_bytes < p.total)
    return fail(DLC_ENOMEM, "dlc_surf_detect: workspace of %zu bytes needed, %zu given", p.total, ws_bytes);
  cudaStream_t s = as_stream(stream);
  char* ws = static_cast<char*>(ws_dev);
  int32_t* sum = reinterpret_cast<int32_t*>(ws + p.off_sum);
  float4* cand = reinterpret_cast<float4*>(ws + p.off_cand);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + p.off_keys);
  int* count = reinterpret_cast<int*>(ws + p.off_count);
  DLC_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * B, s));
  surf_integral_rows_kernel<<<ceil_div(B * H, 8), 256, 0, s>>>(img_dev, B, H, W, sum);
  surf_integral_cols_kernel<<<B * ceil_div(W, 32), 1024, 0, s>>>(B, H, W, sum);
  for (int o = 0; o < n_octaves; ++o) {
    const SurfOctave& oc = p.oc[o];
    const int cells = oc.rows * oc.cols;
    if (cells == 0) continue;
    const int tiles = ceil_div(oc.rows, kSurfTile) * ceil_div(oc.cols, kSurfTile);
    if (g_surf_fast.load() && surf_fast_ok(oc, o)) {   // compile-time box tables + shared-memory lookups
      if (o == 0) {
        surf_octave_fast_kernel<1><<<dim3(tiles, 1, B), 256, surf_fast_smem(1), s>>>(sum, H, W, oc, o, hessian_threshold,
                                                                                      cand, keys, count);
      } else {
        static std::atomic<bool> attr_set[64] = {};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        if (!attr_set[dev].load(std::memory_order_acquire)) {
          DLC_CUDA(cudaFuncSetAttribute(surf_octave_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(surf_fast_smem(2))));
          attr_set[dev].store(true, std::memory_order_release);
        }
        surf_octave_fast_kernel<2><<<dim3(tiles, 1, B), 256, surf_fast_smem(2), s>>>(sum, H, W, oc, o, hessian_threshold,
                                                                                      cand, keys, count);
      }
      continue;
    }
    surf_octave_kernel<<<dim3(tiles, 1, B), 256, sizeof(float) * oc.n * kSurfHalo * kSurfHalo, s>>>(
        sum, H, W, oc, o, hessian_threshold, cand, keys, count);
  }
  surf_top_kernel<<<B, 256, 0, s>>>(cand, keys, count, H, W, top_n, xy_dev, info_dev, found_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

// Training step of the denoising autoencoders: the elementwise / reduction kernels around the
// tcgen05 contractions of planes.cu. Replaces the TensorFlow graph pieces of
//   SDAV._define_model / _define_loss_for_layer / _define_optimizer (src/sdav/network/SDAV.py:120-186, 223-226) and
//   DA._define_fitting_model / _define_loss / _define_optimizer / _corrupt_tensor
//   (src/sdav/network/DenoisingAutoencoderVariant.py:103-158, 182-202).
// One layer's step is: corrupt -> GEMM(+bias+sigmoid) -> decoder GEMM(+bias+sigmoid) -> softmax-cross-entropy
// gradient -> GEMM (gradient into the hidden layer) -> hidden gradient (sparsity + consecutive-frame terms, sigmoid
// derivative) -> column sums (bias gradients) -> transposes + ONE GEMM for the tied-weight gradient
// (x~^T dzh + dzy^T h, the two contractions concatenated along K) -> SGD on the float64 master weights -> re-pack
// of the operand planes. Every GEMM is dlc_gemm_planes (fp16 hi/lo split, 3 products); the host sequences the
// calls (deeploopcloser_b200/training.py), as the reference's Python sequences TensorFlow ops.
#include <math.h>

#include <algorithm>

#include "ptx.cuh"
#include "util.h"

namespace dlc {

// out = x * keep[r % mask_rows] + add[r % mask_rows] -> float32 and/or hi/lo planes (columns >= C written as zero)
__global__ void __launch_bounds__(256)
corrupt_kernel(const float* __restrict__ x, const float* __restrict__ keep, const float* __restrict__ add, int R, int C,
               int mask_rows, float* __restrict__ out_f32, __half* __restrict__ hi, __half* __restrict__ lo, int ld) {
  const int chunks = ld >> 3;
  const int64_t total = static_cast<int64_t>(R) * chunks;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(t / chunks);
    const int c0 = static_cast<int>(t % chunks) << 3;
    const int mr = r % mask_rows;
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float v = 0.0f;
      if (c < C) {
        v = x[static_cast<int64_t>(r) * C + c];
        if (keep) v *= keep[static_cast<int64_t>(mr) * C + c];
        if (add) v += add[static_cast<int64_t>(mr) * C + c];
        if (out_f32) out_f32[static_cast<int64_t>(r) * C + c] = v;
      }
      split_f32(v, h[j], l[j]);
    }
    if (hi) {
      const int64_t o = static_cast<int64_t>(r) * ld + c0;
      *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
      if (lo) *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += sh[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, sh[i]);
  return t;
}

// One block per row. cd = mean_rows( -sum_j L_j * log_softmax(y)_j )   (softmax_cross_entropy_with_logits_v2)
//   dy_j = (softmax(y)_j - L_j) / R  - TensorFlow's REGISTERED gradient (the op's backprop output is softmax - labels,
//          xent_op.h, scaled by the incoming gradient, nn_grad.py [TF1-doc]): what optimizer.minimize follows in the
//          reference although the labels (patch rows) sum to hundreds, not one;
//          exact != 0: the mathematical derivative (softmax(y)_j * sum(L) - L_j) / R instead;
//   dzy = dy * y * (1 - y)  (y is a sigmoid output),
//   dlabel_j = -log_softmax(y)_j / R  (the _v2 op back-propagates into its labels).
__global__ void __launch_bounds__(256)
xent_grad_kernel(const float* __restrict__ y, const float* __restrict__ labels, int R, int C,
                 float* __restrict__ dzy_f32, __half* __restrict__ hi, __half* __restrict__ lo, int ld,
                 float* __restrict__ dlabel, double* __restrict__ loss, int exact) {
  __shared__ double shd[8];
  __shared__ float shf[8];
  const int r = blockIdx.x;
  const float* yr = y + static_cast<int64_t>(r) * C;
  const float* lr = labels + static_cast<int64_t>(r) * C;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, yr[c]);
  m = block_max(m, shf);
  double se = 0.0, sl = 0.0, sly = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double yv = yr[c], lv = lr[c];
    se += exp(yv - static_cast<double>(m));
    sl += lv;
    sly += lv * yv;
  }
  se = block_sum(se, shd);
  sl = block_sum(sl, shd);
  sly = block_sum(sly, shd);
  const double lse = static_cast<double>(m) + log(se);
  const double inv_r = 1.0 / R;
  if (threadIdx.x == 0) atomicAdd(loss, -(sly - sl * lse) * inv_r);
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float g = 0.0f;
    if (c < C) {
      const double yv = yr[c];
      const double ls = yv - lse;
      const double dy = (exp(ls) * (exact ? sl : 1.0) - static_cast<double>(lr[c])) * inv_r;
      g = static_cast<float>(dy * yv * (1.0 - yv));
      if (dzy_f32) dzy_f32[static_cast<int64_t>(r) * C + c] = g;
      if (dlabel) dlabel[static_cast<int64_t>(r) * C + c] = static_cast<float>(-ls * inv_r);
    }
    if (hi) {
      __half h, l;
      split_f32(g, h, l);
      hi[static_cast<int64_t>(r) * ld + c] = h;
      if (lo) lo[static_cast<int64_t>(r) * ld + c] = l;
    }
  }
}

// norms[b] = || h[b] - h[b+1] ||_F over the P x C block of each frame; one block per consecutive pair.
// Also accumulates cc = mean_b norms[b] into loss[0] scaled by `coef` (= consecutive_penalty / (B - 1)).
__global__ void __launch_bounds__(256)
frame_diff_norm_kernel(const float* __restrict__ h, int P, int C, double* __restrict__ norms, double coef,
                       double* __restrict__ loss) {
  __shared__ double shd[8];
  const int b = blockIdx.x;
  const int64_t n = static_cast<int64_t>(P) * C;
  const float* a = h + b * n;
  const float* c = a + n;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = static_cast<double>(a[i]) - static_cast<double>(c[i]);
    s += d * d;
  }
  s = block_sum(s, shd);
  if (threadIdx.x == 0) {
    const double nr = sqrt(s);
    norms[b] = nr;
    if (loss) atomicAdd(loss, coef * nr);
  }
}

// dzh = (dh_rec + dh_up + cs_coef * sign(h - s) + cc_coef * dcc) * h * (1 - h)   -> float32 + planes
//   dcc[b] = (h[b] - h[b+1]) / norm[b]  -  (h[b-1] - h[b]) / norm[b-1]
// cs loss: loss += cs_coef * sum |h - s|   (cs_coef already holds sparse_penalty / count).
__global__ void __launch_bounds__(256)
hidden_grad_kernel(const float* __restrict__ h, const float* __restrict__ dh_rec, const float* __restrict__ dh_up,
                   const double* __restrict__ norms, int B, int P, int C, float sparse_level, double cs_coef,
                   double cc_coef, float* __restrict__ dzh_f32, __half* __restrict__ hi, __half* __restrict__ lo,
                   int ld, double* __restrict__ loss) {
  __shared__ double shd[8];
  const int R = B * P;
  const int64_t total = static_cast<int64_t>(R) * ld;
  const int64_t frame = static_cast<int64_t>(P) * C;
  double abs_sum = 0.0;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(t / ld);
    const int c = static_cast<int>(t % ld);
    float out = 0.0f;
    if (c < C) {
      const int64_t i = static_cast<int64_t>(r) * C + c;
      const double hv = h[i];
      double g = 0.0;
      if (dh_rec) g += dh_rec[i];
      if (dh_up) g += dh_up[i];
      if (cs_coef != 0.0) {
        const double d = hv - static_cast<double>(sparse_level);
        g += cs_coef * (d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0));
        abs_sum += fabs(d);
      }
      if (cc_coef != 0.0) {
        const int b = r / P;
        double dcc = 0.0;
        if (b + 1 < B) dcc += (hv - static_cast<double>(h[i + frame])) / norms[b];
        if (b > 0) dcc -= (static_cast<double>(h[i - frame]) - hv) / norms[b - 1];
        g += cc_coef * dcc;
      }
      out = static_cast<float>(g * hv * (1.0 - hv));
      if (dzh_f32) dzh_f32[i] = out;
    }
    if (hi) {
      __half hh, ll;
      split_f32(out, hh, ll);
      hi[static_cast<int64_t>(r) * ld + c] = hh;
      if (lo) lo[static_cast<int64_t>(r) * ld + c] = ll;
    }
  }
  if (loss && cs_coef != 0.0) {
    abs_sum = block_sum(abs_sum, shd);
    if (threadIdx.x == 0) atomicAdd(loss, cs_coef * abs_sum);
  }
}

// out[c] = sum_r a[r, c]  (float64 accumulation; one thread per column, rows strided so reads stay coalesced)
__global__ void __launch_bounds__(128) colsum_kernel(const float* __restrict__ a, int R, int C, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int r = 0; r < R; ++r) s += a[static_cast<int64_t>(r) * C + c];
  out[c] = s;
}

// a [R, C] float32 -> transposed planes [C, ld]: plane[c][col_off + r] = a[r][c]; 32 x 32 tiles through smem.
__global__ void __launch_bounds__(256)
transpose_planes_kernel(const float* __restrict__ a, int R, int C, __half* __restrict__ hi, __half* __restrict__ lo,
                        int ld, int col_off, int r_pad) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32 per pass
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < R && c < C) ? a[static_cast<int64_t>(r) * C + c] : 0.0f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < C && r < r_pad) {  // rows R..r_pad-1 of the source are the zero padding of the contraction length
      __half h, l;
      split_f32(tile[tx][k], h, l);
      hi[static_cast<int64_t>(c) * ld + col_off + r] = h;
      if (lo) lo[static_cast<int64_t>(c) * ld + col_off + r] = l;
    }
  }
}

template <typename G>
__global__ void __launch_bounds__(256) sgd_kernel(double* __restrict__ w, const G* __restrict__ g, int64_t n, double lr) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    w[i] -= lr * static_cast<double>(g[i]);
}

// out = (dx + extra) * keep[r % mask_rows]   (gradient through x_l = h_{l-1} * mask_l, plus the label gradient)
__global__ void __launch_bounds__(256)
mask_grad_kernel(const float* __restrict__ dx, const float* __restrict__ extra, const float* __restrict__ keep, int R,
                 int C, int mask_rows, float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(R) * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / C), c = static_cast<int>(i % C);
    float v = dx[i];
    if (extra) v += extra[i];
    if (keep) v *= keep[static_cast<int64_t>(r % mask_rows) * C + c];
    out[i] = v;
  }
}

static int grid_for(int64_t total, int block) {
  return static_cast<int>(std::min<int64_t>((total + block - 1) / block, 148 * 16));
}

}  // namespace dlc

using namespace dlc;

extern "C" int dlc_train_corrupt(const float* x_dev, const float* keep_dev, const float* add_dev, int R, int C,
                                 int mask_rows, float* out_f32_dev, void* out_hi_dev, void* out_lo_dev, int ld,
                                 void* stream) {
  DLC_CHECK_ARG(x_dev && (out_f32_dev || out_hi_dev));
  DLC_CHECK_ARG(R > 0 && C > 0 && mask_rows > 0);
  DLC_CHECK_ARG(ld >= C && ld % 8 == 0);
  const int64_t total = static_cast<int64_t>(R) * (ld / 8);
  corrupt_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
      x_dev, keep_dev, add_dev, R, C, mask_rows, out_f32_dev, static_cast<__half*>(out_hi_dev),
      static_cast<__half*>(out_lo_dev), ld);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_train_xent_grad(const float* y_dev, const float* labels_dev, int R, int C, float* dzy_f32_dev,
                                   void* dzy_hi_dev, void* dzy_lo_dev, int ld, float* dlabel_dev, double* loss_dev,
                                   int exact_gradient, void* stream) {
  DLC_CHECK_ARG(y_dev && labels_dev && loss_dev);
  DLC_CHECK_ARG(R > 0 && C > 0);
  DLC_CHECK_ARG(!dzy_hi_dev || ld >= C);
  xent_grad_kernel<<<R, 256, 0, as_stream(stream)>>>(y_dev, labels_dev, R, C, dzy_f32_dev,
                                                     static_cast<__half*>(dzy_hi_dev),
                                                     static_cast<__half*>(dzy_lo_dev), dzy_hi_dev ? ld : C, dlabel_dev,
                                                     loss_dev, exact_gradient);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_train_hidden_grad(const float* h_dev, const float* dh_rec_dev, const float* dh_up_dev, int B, int P,
                                     int C, float sparse_level, double cs_coef, double cc_coef, double* norms_dev,
                                     float* dzh_f32_dev, void* dzh_hi_dev, void* dzh_lo_dev, int ld, double* loss_dev,
                                     void* stream) {
  DLC_CHECK_ARG(h_dev && (dzh_f32_dev || dzh_hi_dev));
  DLC_CHECK_ARG(B > 0 && P > 0 && C > 0);
  DLC_CHECK_ARG(cc_coef == 0.0 || (B >= 2 && norms_dev));
  DLC_CHECK_ARG(!dzh_hi_dev || ld >= C);
  cudaStream_t s = as_stream(stream);
  if (cc_coef != 0.0)
    frame_diff_norm_kernel<<<B - 1, 256, 0, s>>>(h_dev, P, C, norms_dev, cc_coef, loss_dev);
  const int eff_ld = dzh_hi_dev ? ld : C;
  const int64_t total = static_cast<int64_t>(B) * P * eff_ld;
  // cc_coef of the gradient is penalty / (B - 1) as well: d mean_b(norm_b) / d h = (1 / (B - 1)) * d norm_b / d h
  hidden_grad_kernel<<<grid_for(total, 256), 256, 0, s>>>(h_dev, dh_rec_dev, dh_up_dev, norms_dev, B, P, C, sparse_level,
                                                          cs_coef, cc_coef, dzh_f32_dev,
                                                          static_cast<__half*>(dzh_hi_dev),
                                                          static_cast<__half*>(dzh_lo_dev), eff_ld, loss_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_train_colsum(const float* a_dev, int R, int C, double* out_dev, void* stream) {
  DLC_CHECK_ARG(a_dev && out_dev && R > 0 && C > 0);
  colsum_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(a_dev, R, C, out_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_train_transpose_planes(const float* a_dev, int R, int C, void* hi_dev, void* lo_dev, int ld,
                                          int col_off, int r_pad, void* stream) {
  DLC_CHECK_ARG(a_dev && hi_dev && R > 0 && C > 0);
  DLC_CHECK_ARG(r_pad >= R && col_off >= 0 && col_off + r_pad <= ld);
  transpose_planes_kernel<<<dim3(ceil_div(r_pad, 32), ceil_div(C, 32)), 256, 0, as_stream(stream)>>>(
      a_dev, R, C, static_cast<__half*>(hi_dev), static_cast<__half*>(lo_dev), ld, col_off, r_pad);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_train_sgd(double* w_dev, const void* grad_dev, int grad_dtype, int64_t n, double lr, void* stream) {
  DLC_CHECK_ARG(w_dev && grad_dev && n > 0);
  DLC_CHECK_ARG(grad_dtype == DLC_F32 || grad_dtype == DLC_F64);
  if (grad_dtype == DLC_F32)
    sgd_kernel<float><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w_dev, static_cast<const float*>(grad_dev), n, lr);
  else
    sgd_kernel<double><<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w_dev, static_cast<const double*>(grad_dev), n, lr);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_train_mask_grad(const float* dx_dev, const float* extra_dev, const float* keep_dev, int R, int C,
                                   int mask_rows, float* out_dev, void* stream) {
  DLC_CHECK_ARG(dx_dev && out_dev && R > 0 && C > 0 && mask_rows > 0);
  mask_grad_kernel<<<grid_for(static_cast<int64_t>(R) * C, 256), 256, 0, as_stream(stream)>>>(
      dx_dev, extra_dev, keep_dev, R, C, mask_rows, out_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

// SDA encoder forward: H_0 = sigmoid(X W_0 + b_0), H_l = sigmoid(H_{l-1} W_l + b_l).
// Replaces SDAV._define_model + SDAV.transform (src/sdav/network/SDAV.py:120-163, 293-302; weights :188-217)
// and DA._define_transforming_model + DA.transform (src/sdav/network/DenoisingAutoencoderVariant.py:116-119, 254-259).
// Each layer is ONE fused tcgen05 kernel (dlc_gemm_planes): GEMM + bias + sigmoid + re-split of the activations into
// the next layer's fp16 operand planes; only the last layer writes float32 descriptors.
#include <string.h>

#include <algorithm>
#include <vector>

#include "util.h"

struct dlc_sda {
  int n_layers = 0;
  int precision = DLC_PREC_FP16X2;
  std::vector<int> dims;     // n_layers + 1
  std::vector<int> ld;       // plane ld of each layer's input  (dlc_plane_ld(dims[l]))
  std::vector<int> n_pad;    // padded output width of each layer (= ld of the next layer's input)
  std::vector<int> n_alloc;  // rows of the weight planes / bias entries (>= every width gemm_pad() may choose; zeros)
  std::vector<void*> w_hi;   // [n_alloc[l], ld[l]] fp16 (or bf16)
  std::vector<void*> w_lo;
  std::vector<float*> bias;  // [n_alloc[l]]
  std::vector<bool> is_set;
  bool input_u8 = false;     // x planes hold raw pixel values 0..255 (dlc_sda_set_input_u8)
  int chosen = -1;           // DLC_PREC_AUTO: the probe's choice (-1 = not probed yet); else = precision
  double probe_err[2] = {0.0, 0.0};  // one-product / two-product forward vs three products on the probe sample
};

using namespace dlc;

namespace dlc {
extern thread_local int g_gemm_k_valid;  // planes.cu
extern thread_local float g_gemm_alpha;  // planes.cu
int device_sm_count();                   // planes.cu
bool gemm_pairs_enabled();               // planes.cu
bool encoder_wave_pad_enabled();         // planes.cu (dlc_debug_set key 11)
}

namespace {
// precisions whose handles keep the residual (lo) weight planes
bool needs_lo(int precision) {
  return precision == DLC_PREC_FP16X2 || precision == DLC_PREC_FP16X2_A16 || precision == DLC_PREC_AUTO;
}
int out_pad(int n) {
  // Output width padded so that (a) it is the next layer's K (multiple of 64) and (b) a 32-multiple accumulator
  // width <= 256 divides it with little waste: multiples of 256 when n > 256, else multiples of 64.
  if (n > 256) return (n + 255) / 256 * 256;
  return (n + 63) / 64 * 64;
}

// The accumulator width dlc_gemm_planes derives from a padded N (largest multiple of 32 <= 256 that divides it).
int tile_of_pad(int n_pad) {
  for (int cand = 256; cand >= 32; cand -= 32)
    if (n_pad % cand == 0) return cand;
  return 0;
}
// Relative cost of one layer on CTA pairs: tiles are dealt round-robin to sm/2 pairs, so the layer takes
// ceil(tiles / pairs) tile times, and a tile time grows with its width (+ a fixed part: A tile, pipeline fill).
int64_t pair_cost(int rows, int n_pad, int sms) {
  const int tile = tile_of_pad(n_pad);
  const int64_t tiles = static_cast<int64_t>((ceil_div(rows, 128) + 1) / 2) * (n_pad / tile);
  const int pairs = std::max(1, sms / 2);
  return (tiles + pairs - 1) / pairs * (tile + 16);
}
constexpr int kPadTiles[] = {256, 224, 192};  // accumulator widths tried per call (wide tiles only)
// Padded GEMM width of a layer for this call's row count. With few rows per call (a block of a sequence split over
// GPUs, a streaming batch) the default width leaves the last wave of tiles nearly empty - e.g. 3 990 rows x 2 500
// columns: 160 tiles of 256 on 74 pairs = 3 waves for 2.16 waves of work; 192 tiles of 224 = 3 shorter waves.
// The weight planes hold zeros beyond dims[l+1], so any width <= n_alloc gives the same values; the output planes keep
// their width n_pad (columns beyond it are clipped by the store).
int gemm_pad(const dlc_sda* h, int l, int rows) {
  const int n = h->dims[l + 1], sms = device_sm_count();
  int best = h->n_pad[l];
  const int64_t default_tiles = static_cast<int64_t>((ceil_div(rows, 128) + 1) / 2) * (best / tile_of_pad(best));
  // narrow layers, pairs switched off, or too few tiles for the pair kernel (dlc_gemm_planes then runs single CTAs)
  if (n <= 256 || !gemm_pairs_enabled() || !encoder_wave_pad_enabled() || default_tiles < sms / 2) return best;
  int64_t best_cost = pair_cost(rows, best, sms);
  for (int t : kPadTiles) {
    const int cand = ceil_div(n, t) * t;
    if (cand > h->n_alloc[l]) continue;
    const int64_t c = pair_cost(rows, cand, sms);
    if (c < best_cost) {
      best_cost = c;
      best = cand;
    }
  }
  return best;
}
int alloc_pad(int n) {
  int a = out_pad(n);
  if (n > 256)
    for (int t : kPadTiles) a = std::max(a, ceil_div(n, t) * t);
  return a;
}
}  // namespace

extern "C" int dlc_sda_create(dlc_sda** h, int n_layers, const int* dims, int precision) {
  DLC_CHECK_ARG(h && dims);
  DLC_CHECK_ARG(n_layers >= 1 && n_layers <= 64);
  DLC_CHECK_ARG(precision == DLC_PREC_FP16 || precision == DLC_PREC_FP16X2 || precision == DLC_PREC_BF16 ||
                precision == DLC_PREC_AUTO || precision == DLC_PREC_FP16X2_A16);
  for (int i = 0; i <= n_layers; ++i) DLC_CHECK_ARG(dims[i] > 0);
  if (int rc = dlc_device_check()) return rc;
  dlc_sda* s = new dlc_sda();
  s->n_layers = n_layers;
  s->precision = precision;
  s->chosen = precision == DLC_PREC_AUTO ? -1 : precision;
  s->dims.assign(dims, dims + n_layers + 1);
  s->ld.resize(n_layers);
  s->n_pad.resize(n_layers);
  s->n_alloc.resize(n_layers);
  s->w_hi.assign(n_layers, nullptr);
  s->w_lo.assign(n_layers, nullptr);
  s->bias.assign(n_layers, nullptr);
  s->is_set.assign(n_layers, false);
  for (int l = 0; l < n_layers; ++l) {
    s->n_pad[l] = out_pad(dims[l + 1]);
    s->n_alloc[l] = alloc_pad(dims[l + 1]);
    s->ld[l] = l == 0 ? dlc_plane_ld(dims[0]) : s->n_pad[l - 1];
  }
  for (int l = 0; l < n_layers; ++l) {
    const size_t plane = static_cast<size_t>(s->n_alloc[l]) * s->ld[l] * 2;
    cudaError_t e = cudaMalloc(&s->w_hi[l], plane);
    if (e == cudaSuccess && needs_lo(precision)) e = cudaMalloc(&s->w_lo[l], plane);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->bias[l]), sizeof(float) * s->n_alloc[l]);
    if (e != cudaSuccess) {
      dlc_sda_destroy(s);
      return fail(DLC_ENOMEM, "dlc_sda_create: cudaMalloc failed: %s", cudaGetErrorString(e));
    }
  }
  *h = s;
  return DLC_OK;
}

extern "C" int dlc_sda_destroy(dlc_sda* h) {
  if (!h) return DLC_OK;
  for (int l = 0; l < h->n_layers; ++l) {
    if (h->w_hi[l]) cudaFree(h->w_hi[l]);
    if (h->w_lo[l]) cudaFree(h->w_lo[l]);
    if (h->bias[l]) cudaFree(h->bias[l]);
  }
  delete h;
  return DLC_OK;
}

extern "C" int dlc_sda_set_input_u8(dlc_sda* h, int on) {
  DLC_CHECK_ARG(h);
  const bool want = on != 0;
  if (want != h->input_u8) {
    h->input_u8 = want;
    h->is_set[0] = false;  // layer 0 is packed differently: dlc_sda_set_layer(h, 0, ...) again
  }
  return DLC_OK;
}

extern "C" int dlc_sda_set_layer(dlc_sda* h, int l, const double* w_host, const double* b_host) {
  DLC_CHECK_ARG(h && w_host && b_host);
  DLC_CHECK_ARG(l >= 0 && l < h->n_layers);
  const int k = h->dims[l], n = h->dims[l + 1];
  double* w_dev = nullptr;
  const size_t wbytes = sizeof(double) * static_cast<size_t>(k) * n;
  DLC_CUDA(cudaMalloc(reinterpret_cast<void**>(&w_dev), wbytes));
  // Raw-pixel input: sigmoid((p / 255) W + b) = sigmoid((p (W * 256/255)) / 256 + b). The pixel values are exact in
  // fp16, the weights keep their magnitude (hi/lo split stays in the normal fp16 range) and the 1/256 in the epilogue
  // is exact - layer 0 needs two tensor-core products per K step instead of three.
  std::vector<double> scaled;
  if (l == 0 && h->input_u8) {
    scaled.resize(static_cast<size_t>(k) * n);
    for (size_t i = 0; i < scaled.size(); ++i) scaled[i] = w_host[i] * (256.0 / 255.0);
    w_host = scaled.data();
  }
  cudaError_t e = cudaMemcpy(w_dev, w_host, wbytes, cudaMemcpyHostToDevice);
  int rc = DLC_OK;
  if (e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_sda_set_layer: H2D copy failed: %s", cudaGetErrorString(e));
  if (rc == DLC_OK) {
    if (h->precision == DLC_PREC_BF16)
      rc = fail(DLC_EUNSUPPORTED, "dlc_sda_set_layer: bf16 weight packing is not implemented for the encoder");
    else
      rc = dlc_pack_weight_planes(w_dev, DLC_F64, k, n, h->n_alloc[l], h->w_hi[l], h->w_lo[l], h->ld[l], nullptr);
  }
  if (rc == DLC_OK) {
    std::vector<float> b(h->n_alloc[l], 0.0f);
    for (int i = 0; i < n; ++i) b[i] = static_cast<float>(b_host[i]);
    e = cudaMemcpy(h->bias[l], b.data(), sizeof(float) * b.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_sda_set_layer: bias copy failed: %s", cudaGetErrorString(e));
  }
  e = cudaDeviceSynchronize();
  if (rc == DLC_OK && e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_sda_set_layer: %s", cudaGetErrorString(e));
  cudaFree(w_dev);
  if (rc == DLC_OK) h->is_set[l] = true;
  if (h->precision == DLC_PREC_AUTO) h->chosen = -1;  // new weights: probe again
  return rc;
}

extern "C" size_t dlc_sda_workspace_bytes(const dlc_sda* h, int rows) {
  if (!h || rows <= 0 || h->n_layers < 2) return 0;
  int max_pad = 0;
  for (int l = 0; l + 1 < h->n_layers; ++l) max_pad = std::max(max_pad, h->n_pad[l]);
  const size_t plane = align_up(static_cast<size_t>(rows) * max_pad * 2, 256);
  const int planes_per_buf = (h->precision == DLC_PREC_FP16X2 || h->precision == DLC_PREC_AUTO) ? 2 : 1;
  const int bufs = h->n_layers >= 3 ? 2 : 1;
  return plane * planes_per_buf * bufs + 256;
}

namespace {
// All layers in one arithmetic `mode` (DLC_PREC_FP16: one product; DLC_PREC_FP16X2_A16: two - split weights, the
// activations rounded to fp16 between layers; DLC_PREC_FP16X2: three). The workspace layout is the widest one.
int run_chain(dlc_sda* h, int mode, const void* x_hi_dev, const void* x_lo_dev, int rows, float* out_dev,
              void* ws_dev, void* stream) {
  const bool three = mode == DLC_PREC_FP16X2;
  const int gemm_prec = mode == DLC_PREC_FP16 ? DLC_PREC_FP16 : (mode == DLC_PREC_BF16 ? DLC_PREC_BF16 : DLC_PREC_FP16X2);
  int max_pad = 0;
  for (int l = 0; l + 1 < h->n_layers; ++l) max_pad = std::max(max_pad, h->n_pad[l]);
  const size_t plane = align_up(static_cast<size_t>(rows) * max_pad * 2, 256);
  char* base = static_cast<char*>(ws_dev);
  void* buf_hi[2] = {base, base + plane * (three ? 2 : 1)};
  void* buf_lo[2] = {three ? base + plane : nullptr, three ? base + plane * 3 : nullptr};

  const void* a_hi = x_hi_dev;
  const void* a_lo = (three && !h->input_u8) ? x_lo_dev : nullptr;  // raw pixels are exact in fp16: no residual plane
  for (int l = 0; l < h->n_layers; ++l) {
    const bool last = l + 1 == h->n_layers;
    void* o_hi = last ? nullptr : buf_hi[l & 1];
    void* o_lo = last ? nullptr : buf_lo[l & 1];
    g_gemm_k_valid = h->dims[l];  // columns dims[l]..ld of both operands are zero padding
    if (l == 0 && h->input_u8) g_gemm_alpha = 1.0f / 256.0f;
    int rc = dlc_gemm_planes(a_hi, a_lo, h->w_hi[l], h->w_lo[l], rows, h->dims[l + 1], gemm_pad(h, l, rows), h->ld[l],
                             h->bias[l], DLC_ACT_SIGMOID, gemm_prec, last ? out_dev : nullptr, h->dims[l + 1], o_hi,
                             o_lo, h->n_pad[l], stream);
    if (rc != DLC_OK) return rc;
    a_hi = o_hi;
    a_lo = o_lo;
  }
  return DLC_OK;
}

// max over elements of |a - b| / max(1, |b|) as the bits of a non-negative float (atomicMax on the bit pattern)
__global__ void max_rel_err_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                   unsigned int* __restrict__ out_bits) {
  float worst = 0.0f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float e = fabsf(a[i] - b[i]) / fmaxf(1.0f, fabsf(b[i]));
    worst = fmaxf(worst, e == e ? e : INFINITY);  // NaN counts as a failure
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, off));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(worst));
}

constexpr int kProbeBlocks = 4, kProbeBlockRows = 128;
constexpr double kProbeBudget = 3e-4;  // of the 1e-3 descriptor tolerance; the rest covers the tail beyond the sample
}  // namespace

extern "C" int dlc_sda_chosen_precision(const dlc_sda* h) { return h ? h->chosen : -1; }

extern "C" int dlc_sda_probe_stats(const dlc_sda* h, double* out_host) {
  DLC_CHECK_ARG(h && out_host);
  out_host[0] = h->probe_err[0];
  out_host[1] = h->probe_err[1];
  return DLC_OK;
}

extern "C" int dlc_sda_probe(dlc_sda* h, const void* x_hi_dev, const void* x_lo_dev, int rows, void* ws_dev,
                             size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(h && x_hi_dev);
  DLC_CHECK_ARG(rows > 0);
  if (h->precision != DLC_PREC_AUTO) return DLC_OK;
  DLC_CHECK_ARG(x_lo_dev || h->input_u8);
  for (int l = 0; l < h->n_layers; ++l)
    if (!h->is_set[l]) return fail(DLC_EINVAL, "dlc_sda_probe: layer %d has no weights (dlc_sda_set_layer)", l);
  // sample: up to four blocks of 128 consecutive rows spread evenly over the batch, copied into contiguous planes
  const int n_blocks = std::max(1, std::min(kProbeBlocks, rows / kProbeBlockRows));
  const int brows = std::min(rows, kProbeBlockRows);
  const int srows = n_blocks * brows;
  if (ws_bytes < dlc_sda_workspace_bytes(h, srows) || (h->n_layers > 1 && !ws_dev))
    return fail(DLC_ENOMEM, "dlc_sda_probe: workspace of %zu bytes needed, %zu given",
                dlc_sda_workspace_bytes(h, srows), ws_bytes);
  cudaStream_t s = as_stream(stream);
  const size_t row_bytes = static_cast<size_t>(h->ld[0]) * 2;
  const size_t out_elems = static_cast<size_t>(srows) * h->dims[h->n_layers];
  char* scratch = nullptr;
  const size_t off_lo = align_up(srows * row_bytes, 256);
  const size_t off_out = 2 * off_lo;
  const size_t out_bytes = align_up(out_elems * sizeof(float), 256);
  const size_t off_err = off_out + 3 * out_bytes;
  DLC_CUDA(cudaMalloc(reinterpret_cast<void**>(&scratch), off_err + 256));
  int rc = DLC_OK;
  cudaError_t e = cudaMemsetAsync(scratch + off_err, 0, 8, s);
  for (int b = 0; b < n_blocks && e == cudaSuccess; ++b) {
    const size_t r0 = n_blocks > 1 ? static_cast<size_t>(b) * (rows - brows) / (n_blocks - 1) : 0;
    e = cudaMemcpyAsync(scratch + b * brows * row_bytes, static_cast<const char*>(x_hi_dev) + r0 * row_bytes,
                        brows * row_bytes, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && x_lo_dev && !h->input_u8)
      e = cudaMemcpyAsync(scratch + off_lo + b * brows * row_bytes, static_cast<const char*>(x_lo_dev) + r0 * row_bytes,
                          brows * row_bytes, cudaMemcpyDeviceToDevice, s);
  }
  if (e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_sda_probe: sample copy failed: %s", cudaGetErrorString(e));
  float* outs[3] = {reinterpret_cast<float*>(scratch + off_out), reinterpret_cast<float*>(scratch + off_out + out_bytes),
                    reinterpret_cast<float*>(scratch + off_out + 2 * out_bytes)};
  const int modes[3] = {DLC_PREC_FP16, DLC_PREC_FP16X2_A16, DLC_PREC_FP16X2};
  for (int m = 0; m < 3 && rc == DLC_OK; ++m)
    rc = run_chain(h, modes[m], scratch, h->input_u8 ? nullptr : scratch + off_lo, srows, outs[m], ws_dev, stream);
  unsigned int bits[2] = {0, 0};
  if (rc == DLC_OK) {
    unsigned int* err = reinterpret_cast<unsigned int*>(scratch + off_err);
    const int grid = static_cast<int>(std::min<size_t>((out_elems + 255) / 256, 1024));
    max_rel_err_kernel<<<grid, 256, 0, s>>>(outs[0], outs[2], static_cast<int64_t>(out_elems), err);
    max_rel_err_kernel<<<grid, 256, 0, s>>>(outs[1], outs[2], static_cast<int64_t>(out_elems), err + 1);
    e = cudaMemcpyAsync(bits, err, sizeof(bits), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(DLC_ECUDA, "dlc_sda_probe: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(s);
  }
  cudaFree(scratch);
  if (rc != DLC_OK) return rc;
  float e1, e2;
  memcpy(&e1, &bits[0], 4);
  memcpy(&e2, &bits[1], 4);
  h->probe_err[0] = e1;
  h->probe_err[1] = e2;
  h->chosen = e1 <= kProbeBudget ? DLC_PREC_FP16 : (e2 <= kProbeBudget ? DLC_PREC_FP16X2_A16 : DLC_PREC_FP16X2);
  return DLC_OK;
}

extern "C" int dlc_sda_encode(dlc_sda* h, const void* x_hi_dev, const void* x_lo_dev, int rows, float* out_dev,
                              void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(h && x_hi_dev && out_dev);
  DLC_CHECK_ARG(rows > 0);
  DLC_CHECK_ARG(!(h->precision == DLC_PREC_FP16X2 || h->precision == DLC_PREC_AUTO) || x_lo_dev || h->input_u8);
  for (int l = 0; l < h->n_layers; ++l)
    if (!h->is_set[l]) return fail(DLC_EINVAL, "dlc_sda_encode: layer %d has no weights (dlc_sda_set_layer)", l);
  if (ws_bytes < dlc_sda_workspace_bytes(h, rows) || (h->n_layers > 1 && !ws_dev))
    return fail(DLC_ENOMEM, "dlc_sda_encode: workspace of %zu bytes needed, %zu given",
                dlc_sda_workspace_bytes(h, rows), ws_bytes);
  if (h->precision == DLC_PREC_AUTO && h->chosen < 0) {  // first encode with these weights: choose the arithmetic
    if (int rc = dlc_sda_probe(h, x_hi_dev, x_lo_dev, rows, ws_dev, ws_bytes, stream)) return rc;
  }
  return run_chain(h, h->chosen, x_hi_dev, x_lo_dev, rows, out_dev, ws_dev, stream);
}

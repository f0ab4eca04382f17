// Operand planes (fp16 hi/lo split, K-major, zero padded) and the generic fused contraction
//   out = act(A * B^T + bias)
// on tcgen05 tensor cores. Replaces the reference's float64 `x.matmul(W).add(b).sigmoid()` chain
// (src/utils/TensorflowWrapper.py:57-78, used at src/sdav/network/SDAV.py:129,136,143,150,157 and
// src/sdav/network/DenoisingAutoencoderVariant.py:119) and the matmul core of tf.layers.conv2d
// (src/cnn_vtl/network/cnn_vtl.py:33-93).
#include <cuda_bf16.h>

#include <algorithm>
#include <atomic>

#include "bias_act.cuh"
#include "gemm_pair_sm100.cuh"
#include "util.h"

namespace dlc {

// ------------------------------------------------------------------------------------------------
// split: [rows, cols] f32/f64 -> hi/lo planes, regrouping rows (group_in -> group_out)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void split_planes_kernel(const T* __restrict__ src, int rows, int cols, int src_ld, int group_in,
                                    int group_out, int rows_out, __half* __restrict__ hi, __half* __restrict__ lo,
                                    int ld) {
  const int chunks = ld >> 3;
  const int64_t total = static_cast<int64_t>(rows_out) * chunks;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int pr = static_cast<int>(t / chunks);
    const int c0 = static_cast<int>(t % chunks) << 3;
    const int g = pr / group_out;
    const int w = pr - g * group_out;
    const int r = g * group_in + w;
    const bool row_ok = (w < group_in) && (r < rows);
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      if (row_ok && c < cols) {
        const T x = src[static_cast<int64_t>(r) * src_ld + c];
        if (sizeof(T) == 8) split_f64(static_cast<double>(x), h[j], l[j]);
        else split_f32(static_cast<float>(x), h[j], l[j]);
      } else {
        h[j] = __float2half_rn(0.f);
        l[j] = __float2half_rn(0.f);
      }
    }
    const int64_t o = static_cast<int64_t>(pr) * ld + c0;
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l);
  }
}

// W [k, n] row-major -> Wt planes [n_pad, ld]: Wt[j][i] = W[i][j]
template <typename T>
__global__ void pack_weight_kernel(const T* __restrict__ w, int k, int n, int n_pad, __half* __restrict__ hi,
                                   __half* __restrict__ lo, int ld) {
  const int chunks = ld >> 3;
  const int64_t total = static_cast<int64_t>(n_pad) * chunks;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(t % n_pad);  // output row (fastest across threads -> coalesced reads of W)
    const int i0 = static_cast<int>(t / n_pad) << 3;
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + q;
      if (j < n && i < k) {
        const T x = w[static_cast<int64_t>(i) * n + j];
        if (sizeof(T) == 8) split_f64(static_cast<double>(x), h[q], l[q]);
        else split_f32(static_cast<float>(x), h[q], l[q]);
      } else {
        h[q] = __float2half_rn(0.f);
        l[q] = __float2half_rn(0.f);
      }
    }
    const int64_t o = static_cast<int64_t>(j) * ld + i0;
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l);
  }
}

static std::atomic<int> g_split_bk{32};  // smem ring of the 3-product kernel: BK=32 -> 4 stages, BK=64 -> 2 stages
std::atomic<int> g_promote_k{256};  // K elements accumulated inside the tensor core before promotion to fp32 registers
// Valid K of the next dlc_gemm_planes call on this thread (0 = the whole ld). Callers that know their operands are zero
// beyond K (the SDA encoder: 2500 of 2560, 1681 of 1728) set it so that all-zero K blocks are not loaded or multiplied.
thread_local int g_gemm_k_valid = 0;
// Scale of the accumulator before the bias of the next dlc_gemm_planes call on this thread (z = alpha * acc + bias;
// reset to 1 by the call). The SDA encoder's first layer on raw 8-bit pixels uses 1/256 (see dlc_sda_set_input_u8).
thread_local float g_gemm_alpha = 1.0f;
std::atomic<int> g_tma_store{1};   // plane outputs through staged TMA stores (dlc_debug_set key 5 = 0: direct 16-byte stores)
std::atomic<int> g_cta_pair{1};    // contractions on the CTA-pair kernel (dlc_debug_set key 6: 0 never, 1 auto, 2 always)
std::atomic<int> g_sm_reserve{0};  // dlc_set_sm_reserve
std::atomic<int> g_wave_pad{1};    // dlc_debug_set key 11: the encoder picks its GEMM width per call (sda.cu gemm_pad)
std::atomic<int> g_gram_pair{1};   // the SDAV Gram kernel on CTA pairs (dlc_debug_set key 7)
extern std::atomic<int> g_sim_mgroup;  // sdav_sim.cu
extern std::atomic<int> g_refine_cap;  // sdav_sim.cu
static std::atomic<int> g_dbg_flags{0};  // epilogue A/B flags (dlc_debug_set key 3, see BiasActParams::dbg)

template <class Policy>
static cudaError_t launch_gemm_pair_if(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                                       const CUtensorMap& b1, const BiasActParams& p, int clusters,
                                       cudaStream_t stream) {
  if constexpr (!policy_im2col_a<Policy>::value) return launch_gemm_pair<Policy>(a0, a1, b0, b1, p, clusters, stream);
  else return cudaErrorInvalidValue;
}

template <class Policy>
static int run_bias_act(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int m, int n_pad,
                        int ld, BiasActParams p, cudaStream_t stream) {
  constexpr int BK = Policy::Cfg::BK;
  CUtensorMap ta0, ta1, tb0, tb1;
  const bool split = Policy::Cfg::NPROD == 3;
  if (!make_tmap_k_major(&ta0, a_hi, p.ab_fmt, ld, m, ld, BK, kTileM) ||
      !make_tmap_k_major(&tb0, b_hi, p.ab_fmt, ld, n_pad, ld, BK, p.n_tile))
    return fail(DLC_ECUDA, "dlc_gemm_planes: cuTensorMapEncodeTiled failed");
  ta1 = ta0;
  tb1 = tb0;
  if (split) {
    // a_lo == NULL: the A values are exact in fp16 (their residual is zero) -> two products instead of three
    if ((a_lo && !make_tmap_k_major(&ta1, a_lo, p.ab_fmt, ld, m, ld, BK, kTileM)) ||
        !make_tmap_k_major(&tb1, b_lo, p.ab_fmt, ld, n_pad, ld, BK, p.n_tile))
      return fail(DLC_ECUDA, "dlc_gemm_planes: cuTensorMapEncodeTiled failed (lo planes)");
    p.cv_a_lo_zero = a_lo ? 0 : 1;
  }
  p.alpha = g_gemm_alpha;
  g_gemm_alpha = 1.0f;
  const int k_valid = (g_gemm_k_valid > 0 && g_gemm_k_valid <= ld) ? g_gemm_k_valid : ld;
  g_gemm_k_valid = 0;
  p.k_blocks = ceil_div(k_valid, BK);
  p.kc = std::max(1, g_promote_k.load() / BK);
  p.dbg = g_dbg_flags;
  if (!attach_plane_store_maps(p)) return fail(DLC_ECUDA, "dlc_gemm_planes: cuTensorMapEncodeTiled failed (outputs)");
  const int total = p.m_tiles * p.n_tiles;
  cudaError_t e;
  // CTA pairs (cta_group::2) for large 3-product contractions: each CTA stages half of the B tile, see
  // gemm_pair_sm100.cuh. Needs an even split of n_tile into UMMA-legal halves and enough tiles to fill the GPU.
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  // One-product contractions run on pairs too: a single CTA needs (128 + n_tile) x BK operand bytes per MMA step,
  // more than L2 -> SM sustains at the tensor peak (what the one-product Gram kernel showed).
  if (!policy_im2col_a<Policy>::value && g_cta_pair && p.n_tile % 32 == 0 &&
      (pair_tiles >= sm_count() / 2 || g_cta_pair == 2)) {
    if (!make_tmap_k_major(&tb0, b_hi, p.ab_fmt, ld, n_pad, ld, BK, p.n_tile / 2) ||
        (split && !make_tmap_k_major(&tb1, b_lo, p.ab_fmt, ld, n_pad, ld, BK, p.n_tile / 2)))
      return fail(DLC_ECUDA, "dlc_gemm_planes: cuTensorMapEncodeTiled failed (pair B)");
    if (!split) tb1 = tb0;
    e = launch_gemm_pair_if<Policy>(ta0, ta1, tb0, tb1, p, std::min(pair_tiles, sm_count() / 2), stream);
  } else {
    const int grid = total < sm_count() ? total : sm_count();
    e = launch_gemm<Policy>(ta0, ta1, tb0, tb1, p, grid, stream);
  }
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_gemm_planes: launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

// for translation units that do not include the GEMM headers (sda.cu)
int device_sm_count() { return sm_count(); }
bool gemm_pairs_enabled() { return g_cta_pair.load() != 0; }
bool encoder_wave_pad_enabled() { return g_wave_pad.load() != 0; }
int gemm_debug_flags() { return g_dbg_flags.load(); }  // developer A/B switches of the epilogue (dlc_debug_set key 3)

}  // namespace dlc

using namespace dlc;
extern std::atomic<int> g_probe_side_stream;  // sdav_sim.cu
namespace dlc {
extern std::atomic<int> g_surf_fast;  // surf.cu
}

extern "C" int dlc_plane_ld(int cols) { return cols <= 0 ? 0 : (cols + 63) / 64 * 64; }

extern "C" int dlc_set_sm_reserve(int sms) {
  DLC_CHECK_ARG(sms >= 0 && sms <= 64);
  g_sm_reserve = sms;
  return DLC_OK;
}

extern "C" int dlc_debug_set(int key, int value) {
  if (key == 0 && (value == 32 || value == 64)) {
    g_split_bk = value;
    return DLC_OK;
  }
  if (key == 3) {
    g_dbg_flags = value;
    return DLC_OK;
  }
  if (key == 7) {
    g_gram_pair = value ? 1 : 0;
    return DLC_OK;
  }
  if (key == 6) {  // 0: never, 1: when the problem fills the GPU with pairs, 2: whenever the shape allows (tests)
    g_cta_pair = value;
    return DLC_OK;
  }
  if (key == 5) {
    g_tma_store = value ? 1 : 0;
    return DLC_OK;
  }
  if (key == 4 && value >= 1) {
    g_sim_mgroup = value;
    return DLC_OK;
  }
  if (key == 2 && value >= 32) {
    g_promote_k = value;
    return DLC_OK;
  }
  if (key == 9) {  // 0: the similarity precision probe runs on the caller's stream instead of its side stream
    g_probe_side_stream = value ? 1 : 0;
    return DLC_OK;
  }
  if (key == 11) {  // 0: the encoder keeps its default GEMM width whatever the row count
    g_wave_pad = value ? 1 : 0;
    return DLC_OK;
  }
  if (key == 10) {  // 0: the keypoint detector always uses its generic octave kernel
    g_surf_fast = value ? 1 : 0;
    return DLC_OK;
  }
  if (key == 8) {  // capacity of the deferred-refinement list of the SDAV score kernel (-1: default, 0: refine in place)
    g_refine_cap = value;
    return DLC_OK;
  }
  return fail(DLC_EINVAL, "dlc_debug_set: unknown key/value %d/%d", key, value);
}

extern "C" int dlc_split_planes(const void* src_dev, int src_dtype, int rows, int cols, int src_ld, int group_in,
                                int group_out, void* hi_dev, void* lo_dev, int ld, void* stream) {
  DLC_CHECK_ARG((src_dev && hi_dev) || rows == 0);
  DLC_CHECK_ARG(src_dtype == DLC_F32 || src_dtype == DLC_F64);
  DLC_CHECK_ARG(rows >= 0 && cols > 0 && src_ld >= cols);
  DLC_CHECK_ARG(group_in >= 1 && group_out >= group_in);
  DLC_CHECK_ARG(ld >= cols && ld % 8 == 0);
  if (rows == 0) return DLC_OK;
  const int rows_out = ceil_div(rows, group_in) * group_out;
  const int64_t total = static_cast<int64_t>(rows_out) * (ld / 8);
  const int block = 256;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(total, block), static_cast<int64_t>(sm_count()) * 16));
  if (src_dtype == DLC_F64)
    split_planes_kernel<double><<<grid, block, 0, as_stream(stream)>>>(
        static_cast<const double*>(src_dev), rows, cols, src_ld, group_in, group_out, rows_out,
        static_cast<__half*>(hi_dev), static_cast<__half*>(lo_dev), ld);
  else
    split_planes_kernel<float><<<grid, block, 0, as_stream(stream)>>>(
        static_cast<const float*>(src_dev), rows, cols, src_ld, group_in, group_out, rows_out,
        static_cast<__half*>(hi_dev), static_cast<__half*>(lo_dev), ld);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_pack_weight_planes(const void* w_dev, int src_dtype, int k, int n, int n_pad, void* wt_hi_dev,
                                      void* wt_lo_dev, int ld, void* stream) {
  DLC_CHECK_ARG(w_dev && wt_hi_dev);
  DLC_CHECK_ARG(src_dtype == DLC_F32 || src_dtype == DLC_F64);
  DLC_CHECK_ARG(k > 0 && n > 0 && n_pad >= n);
  DLC_CHECK_ARG(ld >= k && ld % 8 == 0);
  const int64_t total = static_cast<int64_t>(n_pad) * (ld / 8);
  const int block = 256;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(total, block), static_cast<int64_t>(sm_count()) * 16));
  if (src_dtype == DLC_F64)
    pack_weight_kernel<double><<<grid, block, 0, as_stream(stream)>>>(static_cast<const double*>(w_dev), k, n, n_pad,
                                                                      static_cast<__half*>(wt_hi_dev),
                                                                      static_cast<__half*>(wt_lo_dev), ld);
  else
    pack_weight_kernel<float><<<grid, block, 0, as_stream(stream)>>>(static_cast<const float*>(w_dev), k, n, n_pad,
                                                                     static_cast<__half*>(wt_hi_dev),
                                                                     static_cast<__half*>(wt_lo_dev), ld);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_gemm_planes(const void* a_hi_dev, const void* a_lo_dev, const void* b_hi_dev, const void* b_lo_dev,
                               int m, int n, int n_pad, int ld, const float* bias_dev, int act, int precision,
                               float* out_f32_dev, int out_ld, void* out_hi_dev, void* out_lo_dev, int out_plane_ld,
                               void* stream) {
  DLC_CHECK_ARG(a_hi_dev && b_hi_dev);
  DLC_CHECK_ARG(m > 0 && n > 0 && n_pad >= n && n_pad % 32 == 0);
  DLC_CHECK_ARG(ld > 0 && ld % 64 == 0);
  DLC_CHECK_ARG(act == DLC_ACT_NONE || act == DLC_ACT_SIGMOID || act == DLC_ACT_RELU);
  DLC_CHECK_ARG(precision == DLC_PREC_FP16 || precision == DLC_PREC_FP16X2 || precision == DLC_PREC_BF16);
  DLC_CHECK_ARG(precision != DLC_PREC_FP16X2 || b_lo_dev);  // a_lo_dev == NULL: A is exact in fp16
  DLC_CHECK_ARG(out_f32_dev || out_hi_dev);
  DLC_CHECK_ARG(!out_f32_dev || out_ld >= n);
  DLC_CHECK_ARG(!out_hi_dev || (out_plane_ld >= 32 && out_plane_ld % 8 == 0));
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(a_hi_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_hi_dev) & 15) == 0);
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(bias_dev) & 15) == 0);  // read as float4; must hold n_pad entries

  BiasActParams p{};
  // largest accumulator width (multiple of 32, <= 256) that divides the padded N
  int n_tile = 0;
  for (int cand = 256; cand >= 32; cand -= 32)
    if (n_pad % cand == 0) {
      n_tile = cand;
      break;
    }
  DLC_CHECK_ARG(n_tile > 0);
  p.n_tile = n_tile;
  p.ab_fmt = precision == DLC_PREC_BF16 ? 1 : 0;
  p.m_tiles = ceil_div(m, kTileM);
  p.n_tiles = n_pad / n_tile;
  p.M = m;
  p.N = n;
  p.bias = bias_dev;
  p.act = act;
  p.out_f32 = out_f32_dev;
  p.out_ld = out_ld;
  p.out_hi = out_hi_dev;
  p.out_lo = precision == DLC_PREC_FP16X2 ? out_lo_dev : nullptr;
  p.out_plane_ld = out_plane_ld;
  cudaStream_t s = as_stream(stream);
  if (precision == DLC_PREC_FP16X2) {
    if (g_split_bk == 64)
      return run_bias_act<BiasActPolicy<64, 3>>(a_hi_dev, a_lo_dev, b_hi_dev, b_lo_dev, m, n_pad, ld, p, s);
    return run_bias_act<BiasActPolicy<32, 3>>(a_hi_dev, a_lo_dev, b_hi_dev, b_lo_dev, m, n_pad, ld, p, s);
  }
  return run_bias_act<BiasActPolicy<64, 1>>(a_hi_dev, a_lo_dev, b_hi_dev, b_lo_dev, m, n_pad, ld, p, s);
}

// Bias + activation epilogue policy of the tcgen05 contraction (shared by the SDA encoder layers, the generic
// dlc_gemm_planes entry point and - with CONV = true - the cnn_vtl convolutions, which add an im2col-mode A operand
// and the per-image min/max of the descriptor tail).
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "gemm_sm100.cuh"
#include "util.h"

#ifndef DLC_WIDE_STORE
#define DLC_WIDE_STORE 1  // 0: every 32-column chunk leaves as its own 32 x 32 store (developer A/B build)
#endif

namespace dlc {

// ------------------------------------------------------------------------------------------------
// bias + activation epilogue
// ------------------------------------------------------------------------------------------------
struct BiasActParams {
  int n_tile, k_blocks, ab_fmt, kc;
  int dbg;  // developer switches for kernel experiments: 1 skip stores, 2 skip activation math, 4 skip the TMA store
            // instructions (values still staged), 8 skip the conv min/max
  int m_tiles, n_tiles;
  int M, N;
  const float* bias;
  float alpha = 1.0f;  // z = alpha * acc + bias (one fused rounding; alpha = 1 is bit-identical to acc + bias)
  int act;
  float* out_f32;
  int out_ld;
  void* out_hi;
  void* out_lo;
  int out_plane_ld;
  // Plane outputs leave through TMA stores: each epilogue warp stages its 32 rows in shared memory and one lane issues
  // cp.async.bulk.tensor stores. Two adjacent 32-column chunks of a tile are staged side by side (128-byte rows,
  // SWIZZLE_128B) and leave as ONE 32 x 64 store per plane; a chunk without a partner (odd chunk count, edge of the
  // plane) uses the 32 x 32 box (64-byte rows, SWIZZLE_64B). Halving the number of store operations matters: the
  // stores share the SM's TMA unit with the operand loads (ncu A/B on the cnn_vtl conv2 layer: 1.54 ms with the
  // 32-column stores, 1.31 ms with the store instructions removed). Rows beyond M are clipped by the tensor map.
  int use_tma_store;
  alignas(64) CUtensorMap tm_out_hi;    // 32 x 32 box
  alignas(64) CUtensorMap tm_out_lo;
  alignas(64) CUtensorMap tm_out_hi64;  // 32 rows x 64 columns
  alignas(64) CUtensorMap tm_out_lo64;
  // cv_a_lo_zero: the A operand is exact in fp16, no residual plane (any BiasActPolicy)
  // ---- CONV only (cnn_vtl): implicit-GEMM A operand, see gemm_sm100.cuh (policy_im2col_a)
  int cv_implicit, cv_a_lo_zero, cv_ohw, cv_ow, cv_pad_t, cv_pad_l, cv_kw, cv_cblocks;
  // ---- CONV only: descriptor tail. Row m = output pixel of image m / cv_ohw; the per-image min / max of every conv
  // output (cnn_vtl.py:109-111) is reduced here, as order-preserving ints.
  int* mm;  // [images, 2]
};

// Tensor maps of the plane outputs for the epilogue's staged TMA stores: boxes of 32 rows x 32 columns (SWIZZLE_64B)
// and 32 rows x 64 columns (SWIZZLE_128B).
extern std::atomic<int> g_tma_store;  // planes.cu
inline bool attach_plane_store_maps(BiasActParams& p) {
  p.use_tma_store = 0;
  if (!p.out_hi || !g_tma_store) return true;
  if (!make_tmap_k_major(&p.tm_out_hi, p.out_hi, p.ab_fmt, p.out_plane_ld, p.M, p.out_plane_ld, 32, 32)) return false;
  p.tm_out_lo = p.tm_out_hi;
  if (p.out_lo && p.ab_fmt != 1 &&
      !make_tmap_k_major(&p.tm_out_lo, p.out_lo, p.ab_fmt, p.out_plane_ld, p.M, p.out_plane_ld, 32, 32))
    return false;
  if (!make_tmap_k_major(&p.tm_out_hi64, p.out_hi, p.ab_fmt, p.out_plane_ld, p.M, p.out_plane_ld, 64, 32)) return false;
  p.tm_out_lo64 = p.tm_out_hi64;
  if (p.out_lo && p.ab_fmt != 1 &&
      !make_tmap_k_major(&p.tm_out_lo64, p.out_lo, p.ab_fmt, p.out_plane_ld, p.M, p.out_plane_ld, 64, 32))
    return false;
  p.use_tma_store = 1;
  return true;
}

// float <-> int whose signed order equals the float order (for atomicMin / atomicMax on floats)
__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat16 x = __float2bfloat16_rn(a), y = __float2bfloat16_rn(b);
  return static_cast<uint32_t>(__bfloat16_as_ushort(x)) | (static_cast<uint32_t>(__bfloat16_as_ushort(y)) << 16);
}

template <int BK, int NPROD, bool CONV = false>
struct BiasActPolicy {
  using Cfg = GemmCfg<BK, NPROD>;
  using Params = BiasActParams;
  static constexpr bool kIm2colA = CONV;
  static constexpr bool kAltTiles = CONV;
  static constexpr int kStageBytes = DLC_WIDE_STORE ? 8192 : 4096;  // per epilogue warp: a hi and a lo staging tile
  static constexpr int kScratchBytes = 8 * kStageBytes;
  static constexpr bool kPromote = NPROD == 3;  // the high-precision mode also needs accurate accumulation
  static constexpr int kEpiWarps = 4;
  static __device__ __forceinline__ bool enabled(const Params&) { return true; }
  static __device__ __forceinline__ bool a_lo_zero(const Params& p) { return p.cv_a_lo_zero != 0; }
  static constexpr uint64_t kHintA = kEvictNormal;
  static constexpr uint64_t kHintB = kEvictLast;  // weights are re-read by every M tile: keep them in L2

  static __device__ __forceinline__ int num_tiles(const Params& p, int cta, int ncta) {
    const int total = p.m_tiles * p.n_tiles;
    return cta < total ? (total - cta + ncta - 1) / ncta : 0;
  }
  // CTA-pair kernel (gemm_pair_sm100.cuh): a pair covers two adjacent 128-row tiles; tc.mt = the first of them
  static __device__ __forceinline__ int num_tiles_pair(const Params& p, int cluster, int nclusters) {
    const int total = ((p.m_tiles + 1) >> 1) * p.n_tiles;
    return cluster < total ? (total - cluster + nclusters - 1) / nclusters : 0;
  }
  static __device__ __forceinline__ TileCoord tile_pair(const Params& p, int cluster, int nclusters, int i) {
    const int t = cluster + i * nclusters;
    TileCoord tc;
    const int mp = t / p.n_tiles;
    tc.mt = 2 * mp;
    tc.nt = t - mp * p.n_tiles;
    return tc;
  }
  // N fastest: the CTAs running concurrently share one A row-block (read from HBM once, then L2).
  static __device__ __forceinline__ TileCoord tile(const Params& p, int cta, int ncta, int i) {
    const int t = cta + i * ncta;
    TileCoord tc;
    tc.mt = t / p.n_tiles;
    tc.nt = t - tc.mt * p.n_tiles;
    return tc;
  }

  struct Epilogue {
    const Params& p;
    const int quarter, lane;
    uint8_t* stage;  // this warp's staging tiles: hi at +0, lo at +kStageBytes / 2
    __device__ Epilogue(const Params& p_, int quarter_, int half_, int lane_, void* scratch)
        : p(p_), quarter(quarter_), lane(lane_),
          stage(static_cast<uint8_t*>(scratch) + (half_ * 4 + quarter_) * kStageBytes) {}

    int row;
    bool row_ok;
    // CONV: image / pixel of this thread's row and the running min / max of the values it produced in this tile
    int img, pix;
    float t_lo, t_hi;
    __device__ __forceinline__ void begin_tile(TileCoord tc) {
      row = tc.mt * kTileM + quarter * 32 + lane;
      row_ok = row < p.M;
      if (CONV) {
        img = row / p.cv_ohw;
        pix = row - img * p.cv_ohw;
        t_lo = INFINITY;
        t_hi = -INFINITY;
      }
    }
    __device__ __forceinline__ void end_tile(TileCoord) {
      if (CONV) {
        if (!p.mm) return;
        // rows of a warp are consecutive pixels: usually one image -> one atomic pair per warp
        const int img0 = __shfl_sync(0xffffffffu, img, 0);
        const bool same = __all_sync(0xffffffffu, !row_ok || img == img0);
        if (same) {
          float lo = t_lo, hi = t_hi;  // rows beyond M still hold +inf / -inf
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
          }
          if (lane == 0 && lo <= hi) {
            atomicMin(p.mm + 2 * img0, float_to_ordered(lo));
            atomicMax(p.mm + 2 * img0 + 1, float_to_ordered(hi));
          }
        } else if (row_ok && t_lo <= t_hi) {
          atomicMin(p.mm + 2 * img, float_to_ordered(t_lo));
          atomicMax(p.mm + 2 * img + 1, float_to_ordered(t_hi));
        }
      }
    }
    __device__ __forceinline__ void post_tile(TileCoord) {}

    // bias + activation on one 32-column chunk. Branch-free per element: the activation switch and the
    // "chunk fully inside N" test are hoisted out of the element loop; sigmoid uses the SFU approximations
    // (ex2.approx / rcp.approx, ~1e-6 relative - three orders below the 1e-3 descriptor tolerance).
    template <int ACT>
    __device__ __forceinline__ void activate(float (&v)[32], float (&h)[32], int col0) {
      float b[32];
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);  // bias has n_pad entries, 128-B aligned rows
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = __ldg(b4 + q);
          b[4 * q] = t.x;
          b[4 * q + 1] = t.y;
          b[4 * q + 2] = t.z;
          b[4 * q + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) b[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float z = fmaf(v[j], p.alpha, b[j]);
        if (ACT == DLC_ACT_SIGMOID) z = __fdividef(1.0f, 1.0f + __expf(-z));
        else if (ACT == DLC_ACT_RELU) z = fmaxf(z, 0.0f);
        h[j] = z;
      }
      if (col0 + 32 > p.N) {  // chunk straddles / lies beyond the valid width: zero the padding columns
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j >= p.N) h[j] = 0.0f;
      }
    }

    template <int SLOT>
    __device__ __forceinline__ void chunk(TileCoord tc, int c, float (&v)[32]) {
      const int col0 = tc.nt * p.n_tile + c * 32;
      float h[32];
      if (p.dbg & 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = v[j];
      } else if (p.act == DLC_ACT_SIGMOID) activate<DLC_ACT_SIGMOID>(v, h, col0);
      else if (p.act == DLC_ACT_RELU) activate<DLC_ACT_RELU>(v, h, col0);
      else activate<DLC_ACT_NONE>(v, h, col0);
      if (p.dbg & 1) {  // keep the values alive without the global stores
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += h[j];
        if (acc == 123456.789f && p.out_f32) p.out_f32[0] = acc;
        return;
      }
      if (row_ok) {
        if (CONV) {
          if (p.mm && !(p.dbg & 8)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              t_lo = fminf(t_lo, h[j]);
              t_hi = fmaxf(t_hi, h[j]);
            }
          }
        }
        if (p.out_f32) {
          float* o = p.out_f32 + static_cast<int64_t>(row) * p.out_ld + col0;
          const bool vec_ok = ((p.out_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (vec_ok && col0 + j + 3 < p.N) {
              *reinterpret_cast<float4*>(o + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (col0 + j + q < p.N) o[j + q] = h[j + q];
            }
          }
        }
      }
      // plane outputs: every lane of the warp takes part in the staged store (rows beyond M are clipped by TMA)
      if (p.out_hi && col0 < p.out_plane_ld) {
        const bool bf16 = p.ab_fmt == 1;  // bf16 planes have no residual plane
        const bool want_lo = !bf16 && p.out_lo;
        // 8 values -> one 16-byte piece of the hi plane and one of the lo plane (packed just before they are stored,
        // so only 8 words are live at a time next to the 128 running sums of the two-level accumulation)
        auto pack8 = [&](int q, uint4& vh, uint4& vl) {
          uint32_t wh[4], wl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = h[8 * q + 2 * j], b = h[8 * q + 2 * j + 1];
            if (bf16) {
              wh[j] = pack_bf2(a, b);
              wl[j] = 0u;
            } else {
              split_f32x2(a, b, wh[j], wl[j]);
            }
          }
          vh = make_uint4(wh[0], wh[1], wh[2], wh[3]);
          vl = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        };
        if (!p.use_tma_store) {  // direct 16-byte stores (developer A/B switch)
          if (row_ok) {
            const int64_t off = static_cast<int64_t>(row) * p.out_plane_ld + col0;
            uint4* oh = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_hi) + off);
            uint4* ol = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out_lo) + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 vh, vl;
              pack8(q, vh, vl);
              oh[q] = vh;
              if (want_lo) ol[q] = vl;
            }
          }
          return;
        }
        // ---- staged TMA store (all 32 lanes take part; rows beyond M are clipped by the tensor map).
        // Contract with the kernels: the warp that handles an even chunk of a tile also handles the next odd one.
        constexpr int kLo = kStageBytes / 2;
        const bool odd = DLC_WIDE_STORE && (c & 1) != 0;
        const bool partner = DLC_WIDE_STORE && !odd && (c + 1) * 32 < p.n_tile && col0 + 32 < p.out_plane_ld;
        const bool wide = odd || partner;
        if (!odd) {  // a new staging tile: the previous store of this warp has left the buffer
          if (lane == 0 && store_pending && !(p.dbg & 4)) tma_store_wait_read();
          __syncwarp();
        }
        store_pending = true;
        if (wide) {
          const int sw = lane & 7;                // SWIZZLE_128B: 16-byte chunk index ^= row % 8
          const int q0 = odd ? 4 : 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 vh, vl;
            pack8(q, vh, vl);
            const int pos = ((q0 + q) ^ sw) * 16;
            *reinterpret_cast<uint4*>(stage + lane * 128 + pos) = vh;
            if (want_lo) *reinterpret_cast<uint4*>(stage + kLo + lane * 128 + pos) = vl;
          }
          if (!odd) return;                       // the partner chunk completes the rows and issues the store
        } else {
          const int sw = (lane >> 1) & 3;         // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 vh, vl;
            pack8(q, vh, vl);
            const int pos = (q ^ sw) * 16;
            *reinterpret_cast<uint4*>(stage + lane * 64 + pos) = vh;
            if (want_lo) *reinterpret_cast<uint4*>(stage + kLo + lane * 64 + pos) = vl;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.dbg & 4)) {
          const int row0 = row - lane;
          if (wide) {
            tma_store_2d(&p.tm_out_hi64, stage, col0 - 32, row0);
            if (want_lo) tma_store_2d(&p.tm_out_lo64, stage + kLo, col0 - 32, row0);
          } else {
            tma_store_2d(&p.tm_out_hi, stage, col0, row0);
            if (want_lo) tma_store_2d(&p.tm_out_lo, stage + kLo, col0, row0);
          }
          tma_store_commit();
        }
      }
    }
    bool store_pending = false;
    __device__ __forceinline__ void finish() {
      if (lane == 0 && store_pending) tma_store_wait_all();
    }
  };
};


}  // namespace dlc

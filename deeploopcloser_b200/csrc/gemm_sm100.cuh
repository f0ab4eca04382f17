// Warp-specialised, persistent tcgen05 GEMM main loop for sm_100a, shared by every dense contraction on the
// loop-closure path (SDA encoder layers, SDAV Gram/score, query x database similarity, conv-as-GEMM).
//
//   D[128 x n_tile] (TMEM, fp32) = sum over products of  A_p[128 x K] * B_p[n_tile x K]^T     (both K-major)
//
// NPROD = 1 : one fp16 (or bf16) product.
// NPROD = 3 : error-compensated fp16 split, A ~= A_hi + A_lo, B ~= B_hi + B_lo,
//             D = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  (the dropped lo*lo term is ~2^-22 relative).
//
// Roles: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane), warps 2.. = epilogue
// (4 warps, or 8 with two-level accumulation; each reads the 32 TMEM lanes of its quarter = warp_idx % 4).
// Pipelines: smem ring full/empty (TMA <-> MMA; depth chosen at launch from n_tile, up to 12 stages), TMEM accumulator double buffer full/empty (MMA <-> epilogue).
// What happens to an accumulator tile is decided by the Policy's Epilogue (bias+sigmoid+re-split, argmin/score,
// running top-k, ...), which reads TMEM directly - accumulators never visit HBM.
#pragma once
#include <atomic>
#include <type_traits>
#include <utility>

#include "ptx.cuh"

namespace dlc {

constexpr int kTileM = 128;
constexpr int kMaxTileN = 256;
constexpr int kGemmThreads = 192;
constexpr int kTmemCols = 512;  // two 256-column fp32 accumulators

struct TileCoord {
  int mt;
  int nt;
};

template <int BK_, int NPROD_>
struct GemmCfg {
  static constexpr int BK = BK_;
  static constexpr int NPROD = NPROD_;
  static_assert(BK == 64 || BK == 32, "BK is one swizzle span: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B) fp16");
  static_assert(NPROD == 1 || NPROD == 3, "1 product or the 3-product fp16 split");
  static constexpr int kPlanes = NPROD == 3 ? 2 : 1;
  static constexpr int kABytes = kTileM * BK * 2;     // one A plane tile
  static constexpr int kBBytes = kMaxTileN * BK * 2;  // one B plane tile at n_tile = 256
  // The smem ring is carved at run time: a stage holds the A plane tile(s) and B plane tile(s) of the launch's
  // actual n_tile, so narrow accumulators (n_tile = 96 of cnn_vtl's conv1) get a deeper ring - with little MMA
  // work per stage, the bytes in flight are what hides the TMA latency.
  static constexpr int kMaxStages = 12;
  static constexpr int kRingBytes = 208 * 1024;
  static constexpr uint32_t kSwizzleMode = BK == 64 ? 2u : 4u;  // UMMA layout code: 2 = 128B, 4 = 64B
  static constexpr uint32_t kSBO = 8 * BK * 2;                  // bytes between 8-row groups
  static constexpr int kBarrierBytes = 256;
  static_assert((2 * kMaxStages + 4) * 8 + 8 <= kBarrierBytes, "barrier block too small");
  static_assert(kRingBytes / (kPlanes * (kABytes + kBBytes)) >= 2, "need at least a double buffer");
  // bytes of one stage / number of stages for a launch (a_planes = 1 when the A residual plane is skipped)
  __host__ __device__ static constexpr int stage_bytes(int n_tile, int a_planes) {
    return a_planes * kABytes + kPlanes * n_tile * BK * 2;
  }
  __host__ __device__ static constexpr int stages(int n_tile, int a_planes, int ring = kRingBytes) {
    const int s = ring / stage_bytes(n_tile, a_planes);
    return s < kMaxStages ? s : kMaxStages;
  }
};

// Policy contract:
//   using Cfg = GemmCfg<BK, NPROD>;
//   static constexpr bool kPromote;      two-level accumulation (see below)
//   struct Params { int n_tile; int k_blocks; int ab_fmt; int kc; ... };            (POD, passed by value)
//   static __device__ int  num_tiles(const Params&, int cta, int ncta);
//   static __device__ TileCoord tile(const Params&, int cta, int ncta, int i);
//   static constexpr int kEpiWarps;      4 or 8 epilogue warps (forced to 8 by kPromote)
//   static __device__ bool enabled(const Params&);                                  (device-side launch gate)
//   struct Epilogue { __device__ Epilogue(const Params&, int quarter, int half, int lane, void* scratch);
//                     __device__ void begin_tile(TileCoord);
//                     template <int SLOT> __device__ void chunk(TileCoord, int c, float (&v)[32]);
//                                       // columns [32c, 32c+32) of this thread's accumulator row; SLOT = compile-time
//                                       // index of the chunk among this warp's chunks (for per-chunk register state)
//                     __device__ void end_tile(TileCoord);    // last call that may rely on the accumulator
//                     __device__ void post_tile(TileCoord);   // runs after the TMEM buffer went back to the MMA warp
//                     __device__ void finish(); };
//
// Two-level accumulation (kPromote). The tensor core adds into its fp32 accumulator with truncation, so a K = 2500
// dot product accumulated in ~470 sequential steps carries a ~1e-5 relative bias (measured on B200) - too much for
// the 1e-3 descriptor tolerance once five saturating layers amplify it. With kPromote the MMA warp accumulates only
// `kc` K-blocks into a TMEM buffer, hands it to the epilogue warps and continues in the other buffer; the epilogue
// warps (8 instead of 4: two per lane quarter, 128 columns each) add the partial sums in registers with
// round-to-nearest fp32 adds and run the real epilogue on the register tile after the last K chunk.
// Thread layout with two-level accumulation: warp group 0 = {TMA, MMA, 2 idle warps} shrinks to 56 registers per
// thread (setmaxnreg), warp groups 1 and 2 = the 8 epilogue warps grow to 224 (128 of them hold the running sums).
constexpr int kGemmThreadsPromote = 384;
constexpr int kRegsLean = 56;
constexpr int kRegsEpilogue = 224;

// Optional policy member `static constexpr bool kIm2colA = true`: the A operand may be an NHWC activation tensor
// read with im2col-mode TMA (implicit-GEMM convolution) instead of a materialised K-major matrix. Params then
// carries: cv_implicit (runtime switch), cv_a_lo_zero (A has no residual plane), cv_ohw (output pixels per image), cv_ow, cv_pad_t, cv_pad_l, cv_kw,
// cv_cblocks (K blocks per filter tap); K block kb = (tap kh*KW+kw, channel block), B's K order is (kh, kw, c).
template <class P, class = void>
struct policy_im2col_a : std::false_type {};
template <class P>
struct policy_im2col_a<P, std::enable_if_t<P::kIm2colA>> : std::true_type {};

// Optional policy member `static bool a_lo_zero(const Params&)` (3-product policies): the A operand of this launch is
// exactly representable in fp16 (8-bit pixel values), so its residual plane is neither staged nor multiplied -
// two products per K step instead of three.
template <class P, class = void>
struct policy_a_lo_zero {
  static __device__ __forceinline__ bool get(const typename P::Params&) { return false; }
};
template <class P>
struct policy_a_lo_zero<P, std::void_t<decltype(P::a_lo_zero(std::declval<const typename P::Params&>()))>> {
  static __device__ __forceinline__ bool get(const typename P::Params& p) { return P::Cfg::NPROD == 3 && P::a_lo_zero(p); }
};

// Optional policy member `static constexpr bool kAltTiles = true` (two-level-accumulation policies): when the
// accumulator is at most 128 columns wide and the whole K range is one chunk, the two groups of four epilogue warps
// take alternate tiles (group g owns TMEM buffer g) instead of splitting the columns of every tile, so a narrow,
// short-K contraction - whose epilogue costs more than its MMAs - has two tile times to drain each tile.
template <class P, class = void>
struct policy_alt_tiles : std::false_type {};
template <class P>
struct policy_alt_tiles<P, std::enable_if_t<P::kAltTiles>> : std::true_type {};

// Optional policy member `static constexpr bool kFrameMaps = true`: operands are stored at P <= 32 rows per frame and
// read through 3-D tensor maps {K, P, frames} with 32-row boxes (rows P..31 arrive as zeros): the A tile is always
// 4 frames x 32 rows; the B tile is frame-mapped when Params::b_frame_map != 0 (else a plain 2-D tile of n_tile rows).
template <class P, class = void>
struct policy_frame_maps : std::false_type {};
template <class P>
struct policy_frame_maps<P, std::enable_if_t<P::kFrameMaps>> : std::true_type {};

// Optional policy member `static constexpr int kScratchBytes`: 1024-byte-aligned shared memory handed to the
// epilogue (e.g. per-warp staging tiles for TMA stores); it is taken out of the operand ring.
template <class P, class = void>
struct policy_scratch {
  static constexpr int value = 0;
};
template <class P>
struct policy_scratch<P, std::enable_if_t<(P::kScratchBytes > 0)>> {
  static constexpr int value = P::kScratchBytes;
};
constexpr int kSmemLimit = 227 * 1024;
constexpr int kSmemFixed = 1024 + 16 /*alignment slack*/ + 1024 /*barrier block, padded*/;
template <class Policy>
__host__ __device__ constexpr int ring_bytes() {
  const int avail = (kSmemLimit - kSmemFixed - policy_scratch<Policy>::value) / 1024 * 1024;
  return avail < Policy::Cfg::kRingBytes ? avail : Policy::Cfg::kRingBytes;
}
template <class Policy>
__host__ __device__ constexpr int smem_bytes() {
  return kSmemFixed + policy_scratch<Policy>::value + ring_bytes<Policy>();
}

// Optional policy member `static int chunk_stride(const Params&)`: distance in accumulator columns between the
// 32-column chunks handed to the epilogue (default 32). The SDAV Gram kernel packs its N side at 30 rows per frame,
// so chunk c (= frame c of the tile) starts at column 30 c and n_tile is 240: no MMA work on the 2 pad rows.
template <class P, class = void>
struct policy_chunk_stride {
  static __device__ __forceinline__ int get(const typename P::Params&) { return 32; }
};
template <class P>
struct policy_chunk_stride<P, std::void_t<decltype(P::chunk_stride(std::declval<const typename P::Params&>()))>> {
  static __device__ __forceinline__ int get(const typename P::Params& p) { return P::chunk_stride(p); }
};

// Epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, 128 accumulator columns each).
template <class Policy>
__host__ __device__ constexpr int epi_warps() {
  return Policy::kPromote ? 8 : Policy::kEpiWarps;
}
template <class Policy>
__host__ __device__ constexpr int gemm_threads() {
  return Policy::kPromote ? kGemmThreadsPromote : (Policy::kEpiWarps == 8 ? 320 : kGemmThreads);
}

// One 32-column chunk of this thread's accumulator row; SLOT is a compile-time constant so that per-chunk epilogue
// state indexed by it stays in registers.
template <class Epi, int SLOT>
__device__ __forceinline__ void epi_slot_from_tmem(Epi& epi, TileCoord tc, uint32_t taddr, int half, int n_cchunks,
                                                   int cstride) {
  const int c = half * 4 + SLOT;
  if (c < n_cchunks) {
    uint32_t r[32];
    tmem_ld_x32(taddr + c * cstride, r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    epi.template chunk<SLOT>(tc, c, v);
  }
}
template <class Epi, int SLOT>
__device__ __forceinline__ void epi_slot_from_regs(Epi& epi, TileCoord tc, const float (&sums)[128], int half,
                                                   int n_cchunks) {
  const int c = half * 4 + SLOT;
  if (c < n_cchunks) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = sums[SLOT * 32 + j];
    epi.template chunk<SLOT>(tc, c, v);
  }
}
// chunk c = first + SLOT for SLOT < count (the pair kernel's assignment)
template <class Epi, int SLOT>
__device__ __forceinline__ void epi_slot_from_regs_at(Epi& epi, TileCoord tc, const float (&sums)[128], int first,
                                                      int count) {
  if (SLOT < count) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = sums[SLOT * 32 + j];
    epi.template chunk<SLOT>(tc, first + SLOT, v);
  }
}

template <class Policy>
__global__ void __launch_bounds__(gemm_threads<Policy>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ typename Policy::Params p) {
  using Cfg = typename Policy::Cfg;
  constexpr int BK = Cfg::BK;
  constexpr int SMAX = Cfg::kMaxStages;
  constexpr bool kPromote = Policy::kPromote;
  constexpr int kEpiWarps = epi_warps<Policy>();
  static_assert(kEpiWarps == 4 || kEpiWarps == 8, "4 or 8 epilogue warps");

  // smem: [barriers | pad to 1024 | epilogue scratch (policy) | ring of S stages]
  extern __shared__ uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 15) & ~uintptr_t(15));
  uint64_t* full = bars;
  uint64_t* empty = bars + SMAX;
  uint64_t* tfull = bars + 2 * SMAX;
  uint64_t* tempty = bars + 2 * SMAX + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SMAX + 4);
  uint8_t* scratch_b = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(bars) + Cfg::kBarrierBytes + 1023) & ~uintptr_t(1023));
  void* scratch = scratch_b;
  uint8_t* smem = scratch_b + policy_scratch<Policy>::value;  // scratch sizes are multiples of 1024

  // the A operand is exactly representable in fp16 (8-bit pixels): its residual plane is neither staged nor multiplied
  const bool a_lo_zero = policy_a_lo_zero<Policy>::get(p);
  const int a_planes = a_lo_zero ? 1 : Cfg::kPlanes;
  const int stage_bytes = Cfg::stage_bytes(p.n_tile, a_planes);
  const int S = Cfg::stages(p.n_tile, a_planes, ring_bytes<Policy>());
  const int b_plane_bytes = p.n_tile * BK * 2;
  bool alt_tiles = false;
  if constexpr (kPromote && policy_alt_tiles<Policy>::value)
    alt_tiles = p.n_tile <= 128 && (p.kc <= 0 || p.kc >= p.k_blocks);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int ncta = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (Cfg::NPROD == 3) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], alt_tiles ? kEpiWarps / 2 : kEpiWarps);  // one arrival per epilogue warp that reads it
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr int kEpiWarp0 = kPromote ? 4 : 2;  // first epilogue warp

  // a launch can be gated off by device-side state (e.g. the Gram precision probe picks one of two kernels)
  const int my_tiles = Policy::enabled(p) ? Policy::num_tiles(p, cta, ncta) : 0;
  const int n_tile = p.n_tile;
  const int k_blocks = p.k_blocks;
  // K chunks per tile: one (plain accumulation) or ceil(k_blocks / kc) (two-level accumulation)
  const int kc = kPromote ? (p.kc > 0 ? p.kc : k_blocks) : k_blocks;
  const int n_chunks = (k_blocks + kc - 1) / kc;

  if (warp < kEpiWarp0) {
  if (kPromote) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLean));
  if (warp == 0) {
    if (elect_one_sync()) {
      // ===================== TMA producer =====================
      const uint32_t tx_bytes = static_cast<uint32_t>(stage_bytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const TileCoord tc = Policy::tile(p, cta, ncta, i);
        const int m0 = tc.mt * kTileM;
        const int n0 = tc.nt * n_tile;
        // implicit-GEMM A operand: first output pixel of the tile -> base input pixel (w, h, image)
        int cv_w = 0, cv_h = 0, cv_n = 0, cv_kh = 0, cv_kwi = 0, cv_cb = 0;
        bool implicit_a = false;
        if constexpr (policy_im2col_a<Policy>::value) {
          implicit_a = p.cv_implicit != 0;
          if (implicit_a) {
            cv_n = m0 / p.cv_ohw;
            const int rem = m0 - cv_n * p.cv_ohw;
            const int oh = rem / p.cv_ow;
            cv_h = oh - p.cv_pad_t;
            cv_w = rem - oh * p.cv_ow - p.cv_pad_l;
          }
        }
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, 1);
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          uint8_t* st = smem + stage * stage_bytes;
          if constexpr (policy_im2col_a<Policy>::value) {
            if (implicit_a) {
              const uint16_t off_h = static_cast<uint16_t>(cv_kh), off_w = static_cast<uint16_t>(cv_kwi);
              tma_load_im2col_4d(st, &tmA0, &full[stage], cv_cb * BK, cv_w, cv_h, cv_n, off_w, off_h);
              if (Cfg::NPROD == 3 && !a_lo_zero)
                tma_load_im2col_4d(st + Cfg::kABytes, &tmA1, &full[stage], cv_cb * BK, cv_w, cv_h, cv_n, off_w, off_h);
              if (++cv_cb == p.cv_cblocks) {
                cv_cb = 0;
                if (++cv_kwi == p.cv_kw) {
                  cv_kwi = 0;
                  ++cv_kh;
                }
              }
            }
          }
          uint8_t* sb = st + a_planes * Cfg::kABytes;
          if constexpr (policy_frame_maps<Policy>::value) {
            tma_load_3d(st, &tmA0, &full[stage], kb * BK, 0, m0 >> 5, Policy::kHintA);
            if (Cfg::NPROD == 3 && !a_lo_zero)
              tma_load_3d(st + Cfg::kABytes, &tmA1, &full[stage], kb * BK, 0, m0 >> 5, Policy::kHintA);
            if (p.b_frame_map) {
              tma_load_3d(sb, &tmB0, &full[stage], kb * BK, 0, n0 >> 5, Policy::kHintB);
              if (Cfg::NPROD == 3)
                tma_load_3d(sb + b_plane_bytes, &tmB1, &full[stage], kb * BK, 0, n0 >> 5, Policy::kHintB);
            } else {
              tma_load_2d(sb, &tmB0, &full[stage], kb * BK, n0, Policy::kHintB);
              if (Cfg::NPROD == 3) tma_load_2d(sb + b_plane_bytes, &tmB1, &full[stage], kb * BK, n0, Policy::kHintB);
            }
          } else {
            if (!implicit_a) {
              tma_load_2d(st, &tmA0, &full[stage], kb * BK, m0, Policy::kHintA);
              if (Cfg::NPROD == 3 && !a_lo_zero)
                tma_load_2d(st + Cfg::kABytes, &tmA1, &full[stage], kb * BK, m0, Policy::kHintA);
            }
            tma_load_2d(sb, &tmB0, &full[stage], kb * BK, n0, Policy::kHintB);
            if (Cfg::NPROD == 3) tma_load_2d(sb + b_plane_bytes, &tmB1, &full[stage], kb * BK, n0, Policy::kHintB);
          }
          if (++stage == S) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_idesc_f16(kTileM, n_tile, p.ab_fmt);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int kb = 0;
        for (int ch = 0; ch < n_chunks; ++ch) {
          mbar_wait(&tempty[acc], acc_phase ^ 1u, 2);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kMaxTileN);
          const int kb_end = kb + kc < k_blocks ? kb + kc : k_blocks;
          for (int kk = 0; kb < kb_end; ++kb, ++kk) {
            mbar_wait(&full[stage], phase, 3);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * stage_bytes);
            const uint32_t a_lo = a_hi + Cfg::kABytes;
            const uint32_t b_hi = a_hi + a_planes * Cfg::kABytes;
            const uint32_t b_lo = b_hi + b_plane_bytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t koff = k * 32;  // 16 fp16 along K inside the swizzle span
              const uint64_t dah = make_smem_desc(a_hi + koff, Cfg::kSBO, Cfg::kSwizzleMode);
              const uint64_t dbh = make_smem_desc(b_hi + koff, Cfg::kSBO, Cfg::kSwizzleMode);
              if (Cfg::NPROD == 3) {
                // small cross terms first, dominant term last
                const uint64_t dal = make_smem_desc(a_lo + koff, Cfg::kSBO, Cfg::kSwizzleMode);
                const uint64_t dbl = make_smem_desc(b_lo + koff, Cfg::kSBO, Cfg::kSwizzleMode);
                if (!a_lo_zero) {
                  umma_f16(d_tmem, dal, dbh, idesc, (kk | k) != 0 ? 1u : 0u);
                  umma_f16(d_tmem, dah, dbl, idesc, 1u);
                } else {
                  umma_f16(d_tmem, dah, dbl, idesc, (kk | k) != 0 ? 1u : 0u);
                }
                umma_f16(d_tmem, dah, dbh, idesc, 1u);
              } else {
                umma_f16(d_tmem, dah, dbh, idesc, (kk | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&empty[stage]);  // smem slot is free once these MMAs have read it
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&tfull[acc]);  // (partial) accumulator complete -> epilogue
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  }
  } else {
    // ===================== epilogue warps =====================
    if (kPromote) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    const int quarter = warp & 3;               // TMEM lane quarter this warp may read
    const int half = kEpiWarps == 8 ? (warp - kEpiWarp0) >> 2 : 0;  // which 128-column half this warp owns
    const int cstride = policy_chunk_stride<Policy>::get(p);  // columns between chunks (32 unless the policy packs tighter)
    const int n_cchunks = n_tile / cstride;     // chunks in the accumulator
    typename Policy::Epilogue epi(p, quarter, half, lane, scratch);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int col_half = alt_tiles ? 0 : half;  // alternate-tile mode: this group reads all (<= 4) column chunks
    for (int i = 0; i < my_tiles; ++i) {
      if (kPromote && alt_tiles) {
        if ((i & 1) != half) continue;            // the other group's tile
        acc = half;                               // one chunk per tile: tile i accumulates in buffer i & 1
        acc_phase = static_cast<uint32_t>(i >> 1) & 1u;
      }
      const TileCoord tc = Policy::tile(p, cta, ncta, i);
      if (kPromote) {
        float sums[128];
        for (int ch = 0; ch < n_chunks; ++ch) {
          mbar_wait(&tfull[acc], acc_phase, 4);
          tc_fence_after();
          const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kMaxTileN) +
                                 (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(col_half * 4 * cstride);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (col_half * 4 + cc < n_cchunks) {
              uint32_t v[32];
              tmem_ld_x32(taddr + cc * cstride, v);
              tmem_ld_wait();
              if (ch == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sums[cc * 32 + j] = __uint_as_float(v[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) sums[cc * 32 + j] += __uint_as_float(v[j]);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
        epi.begin_tile(tc);
        using Epi = typename Policy::Epilogue;
        bool rolled = false;
        if constexpr (policy_alt_tiles<Policy>::value) rolled = alt_tiles;
        if (rolled) {
          // Narrow accumulator, short K (conv1): the MMA of a tile is brief and the epilogue is a latency chain of
          // straight-line code that never repeats within a tile - ncu showed `no_instruction` among the top stalls.
          // ONE copy of the chunk code, executed per chunk; the running sums rotate down by 32 registers.
#pragma unroll 1
          for (int cc = 0; cc < n_cchunks; ++cc) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = sums[j];
            epi.template chunk<0>(tc, cc, v);
#pragma unroll
            for (int j = 0; j < 96; ++j) sums[j] = sums[j + 32];
          }
        } else {
          epi_slot_from_regs<Epi, 0>(epi, tc, sums, col_half, n_cchunks);
          epi_slot_from_regs<Epi, 1>(epi, tc, sums, col_half, n_cchunks);
          epi_slot_from_regs<Epi, 2>(epi, tc, sums, col_half, n_cchunks);
          epi_slot_from_regs<Epi, 3>(epi, tc, sums, col_half, n_cchunks);
        }
        epi.end_tile(tc);
        epi.post_tile(tc);
      } else {
        mbar_wait(&tfull[acc], acc_phase, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kMaxTileN) + (static_cast<uint32_t>(quarter * 32) << 16);
        epi.begin_tile(tc);
        using Epi = typename Policy::Epilogue;
        epi_slot_from_tmem<Epi, 0>(epi, tc, taddr, half, n_cchunks, cstride);
        epi_slot_from_tmem<Epi, 1>(epi, tc, taddr, half, n_cchunks, cstride);
        epi_slot_from_tmem<Epi, 2>(epi, tc, taddr, half, n_cchunks, cstride);
        epi_slot_from_tmem<Epi, 3>(epi, tc, taddr, half, n_cchunks, cstride);
        if (kEpiWarps == 4) {  // one warp per lane quarter walks all 8 chunks
          epi_slot_from_tmem<Epi, 4>(epi, tc, taddr, half, n_cchunks, cstride);
          epi_slot_from_tmem<Epi, 5>(epi, tc, taddr, half, n_cchunks, cstride);
          epi_slot_from_tmem<Epi, 6>(epi, tc, taddr, half, n_cchunks, cstride);
          epi_slot_from_tmem<Epi, 7>(epi, tc, taddr, half, n_cchunks, cstride);
        }
        epi.end_tile(tc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        epi.post_tile(tc);  // work that no longer needs the accumulator (TMEM buffer already handed back)
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps and launch
// ------------------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

// Row-major [rows, inner] 16-bit matrix with `pitch_elems` elements per row; box = [box_rows, bk] with the swizzle
// span equal to bk*2 bytes. Out-of-range rows/columns are zero-filled by the TMA unit.
inline bool make_tmap_k_major(CUtensorMap* m, const void* base, int ab_fmt, uint64_t inner, uint64_t rows,
                              uint64_t pitch_elems, int bk, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMapDataType dt = ab_fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Frame-grouped K-major planes [frames * rows_per_frame, ld] seen as {ld, rows_per_frame, frames}; box = bk columns x
// 32 rows x box_frames frames. Rows rows_per_frame..31 of every frame (and frames beyond the last) read as zeros.
inline bool make_tmap_frames(CUtensorMap* m, const void* base, uint64_t ld, int rows_per_frame, uint64_t frames, int bk,
                             int box_frames) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[3] = {ld, static_cast<cuuint64_t>(rows_per_frame), frames};
  cuuint64_t gstride[2] = {ld * 2, ld * 2 * static_cast<cuuint64_t>(rows_per_frame)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(bk), 32u, static_cast<cuuint32_t>(box_frames)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// NHWC activation planes [N, H, W, C] (pixel pitch `ld_elems`) read in im2col mode for a stride-1 convolution with
// a KH x KW filter and pad_t / pad_l zero pixels before (pad_b / pad_r after): one load = 128 consecutive output
// pixels x bk channels of one filter tap. Bounding box of the base pixel: [-pad, size + pad_after - (K - 1)).
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeIm2col get_encode_im2col() {
  static PFN_encodeIm2col fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeIm2col>(sym);
  }
  return fn;
}

inline bool make_tmap_im2col_nhwc(CUtensorMap* m, const void* base, int ab_fmt, int N, int H, int W, int C,
                                  uint64_t ld_elems, int KH, int KW, int pad_t, int pad_l, int pad_b, int pad_r,
                                  int bk) {
  PFN_encodeIm2col enc = get_encode_im2col();
  if (!enc) return false;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t gstride[3] = {ld_elems * 2, ld_elems * 2 * W, ld_elems * 2 * W * H};
  int lower[2] = {-pad_l, -pad_t};                          // {W, H}
  int upper[2] = {pad_r - (KW - 1), pad_b - (KH - 1)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMapDataType dt = ab_fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(m, dt, 4, const_cast<void*>(base), gdim, gstride, lower, upper, static_cast<cuuint32_t>(bk),
                   static_cast<cuuint32_t>(kTileM), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <class Policy>
inline cudaError_t launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                               const CUtensorMap& b1, const typename Policy::Params& p, int grid,
                               cudaStream_t stream) {
  using Cfg = typename Policy::Cfg;
  static_assert(smem_bytes<Policy>() <= kSmemLimit, "exceeds the 227 KB per-CTA shared memory limit");
  static_assert(policy_scratch<Policy>::value % 1024 == 0, "epilogue scratch must be a multiple of 1024 bytes");
  // the opt-in shared-memory size is a per-device function attribute: remember it per device, not per process
  static std::atomic<bool> attr_set[kMaxDevices] = {};
  const int dev = current_device();
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes<Policy>());
    if (e != cudaSuccess) return e;
    attr_set[dev].store(true, std::memory_order_release);
  }
  gemm_tc_kernel<Policy><<<grid, gemm_threads<Policy>(), smem_bytes<Policy>(), stream>>>(a0, a1, b0, b1, p);
  return cudaGetLastError();
}

// SMs the persistent kernels of this library size their grids for: the CURRENT device's count (cached per device)
// minus dlc_set_sm_reserve(). A sequence split over GPUs reserves a few SMs so that the NCCL kernels of the exchange
// stage - launched on another stream while a persistent tensor kernel holds every SM - start at once instead of
// waiting for the running layer to end.
extern std::atomic<int> g_sm_reserve;  // planes.cu
inline int sm_count() {
  static std::atomic<int> cache[kMaxDevices] = {};
  const int dev = current_device();
  int n = cache[dev].load(std::memory_order_relaxed);
  if (!n) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  const int keep = n - g_sm_reserve.load(std::memory_order_relaxed);
  return keep < 2 ? 2 : keep & ~1;
}

}  // namespace dlc

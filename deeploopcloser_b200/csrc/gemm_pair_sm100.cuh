// CTA-pair variant of the tcgen05 GEMM main loop (gemm_sm100.cuh): two CTAs of a 2-CTA cluster (one TPC) compute a
// 256 x n_tile accumulator with ONE `tcgen05.mma.cta_group::2`. Each CTA stages its own 128 rows of A and only HALF
// of the B tile, so the shared-memory fill per MMA cycle drops from 48 KB to 32 KB per K block of the 3-product
// kernel (64 -> 43 bytes/clk/SM) and the operand ring gets 6 stages instead of 4 - the single-CTA encoder layer was
// starved for operands (ncu: the MMA thread waited on the `full` barrier 31 % of the time, tensor pipe 78-84 %).
//
// Roles per CTA: warp 0 = TMA producer (its loads signal the LEADER's full barrier), warp 1 = TMEM owner, and in the
// leader (cluster rank 0) also the MMA issuer; warps 4..11 = epilogue (two-level accumulation, 8 warps, as in the
// single-CTA kernel) on the CTA's own 128 accumulator rows in its own TMEM.
// Barriers: full[s]   leader only; 1 arrival (leader producer, expect_tx = both CTAs' bytes) + TMA bytes of both CTAs
//           empty[s]  per CTA; arrived by the leader's MMA commit, multicast to both CTAs
//           tfull[a]  per CTA; same multicast commit
//           tempty[a] leader only; 16 arrivals = the 8 epilogue warps of each CTA (remote arrive from the peer)
// The epilogue always drains the accumulator into registers first (the two-level accumulation path; with one K
// chunk - kc >= k_blocks, what one-product policies pass - that is a plain copy that hands the TMEM buffer back at
// once). NPROD = 1 policies stage and multiply one plane per operand. Tiles come from Policy::tile_pair().
#pragma once
#include "gemm_sm100.cuh"

namespace dlc {

template <class Policy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreadsPromote, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ typename Policy::Params p) {
  using Cfg = typename Policy::Cfg;
  static_assert(Cfg::NPROD == 3 ? Policy::kPromote : !Policy::kPromote,
                "3-product policies use two-level accumulation, 1-product policies plain accumulation");
  constexpr int kPl = Cfg::kPlanes;
  // implicit-GEMM (im2col) A operands: each CTA walks its own 128 output pixels. The column chunks are then split
  // between the two epilogue warp groups at an even chunk near the middle (conv1: 3 chunks -> 2 + 1) instead of at
  // chunk 4, so that narrow accumulators keep both groups busy; even, because the bias/activation epilogue stores
  // chunk pairs.
  constexpr bool kConv = policy_im2col_a<Policy>::value;
  constexpr int BK = Cfg::BK;
  constexpr int SMAX = Cfg::kMaxStages;
  constexpr int kEpiWarps = 8;

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
  uint8_t* smem = scratch_b + policy_scratch<Policy>::value;

  const int n_tile = p.n_tile;
  const int half_n = n_tile >> 1;                         // B rows staged by each CTA
  const int b_plane_bytes = half_n * BK * 2;
  const bool a_lo_zero = policy_a_lo_zero<Policy>::get(p);   // A exact in fp16: no residual plane, two products
  const int a_planes = a_lo_zero ? 1 : kPl;
  const int stage_bytes = a_planes * Cfg::kABytes + kPl * b_plane_bytes;
  int S = ring_bytes<Policy>() / stage_bytes;
  S = S < SMAX ? S : SMAX;

  // Alternate-tile epilogue (policies with kAltTiles; narrow accumulator, K in one chunk - cnn_vtl conv1): warp group g
  // of each CTA drains the tiles that accumulate in TMEM buffer g, all of their (<= 4) column chunks, instead of
  // splitting the columns of every tile - two tile times per tile for an epilogue-bound contraction.
  bool alt_tiles = false;
  if constexpr (Policy::kPromote && policy_alt_tiles<Policy>::value)
    alt_tiles = p.n_tile <= 128 && (p.kc <= 0 || p.kc >= p.k_blocks);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (kPl == 2) {
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
        mbar_init(&tempty[a], alt_tiles ? kEpiWarps : 2 * kEpiWarps);  // the epilogue warps of both CTAs that read it
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = Policy::enabled(p) ? Policy::num_tiles_pair(p, cluster, nclusters) : 0;
  const int k_blocks = p.k_blocks;
  const int kc = Policy::kPromote ? (p.kc > 0 ? p.kc : k_blocks) : k_blocks;
  const int n_chunks = (k_blocks + kc - 1) / kc;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLean));
    if (warp == 0) {
      if (elect_one_sync()) {
        // ===================== TMA producer (both CTAs) =====================
        const uint32_t tx_pair = 2u * static_cast<uint32_t>(stage_bytes);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < my_tiles; ++i) {
          const TileCoord tc = Policy::tile_pair(p, cluster, nclusters, i);  // tc.mt = the pair's FIRST 128-row tile
          const int m0 = (tc.mt + static_cast<int>(rank)) * kTileM;
          const int n0 = tc.nt * n_tile + static_cast<int>(rank) * half_n;
          // implicit-GEMM A operand: first output pixel of this CTA's rows -> base input pixel (w, h, image)
          int cv_w = 0, cv_h = 0, cv_n = 0, cv_kh = 0, cv_kwi = 0, cv_cb = 0;
          if constexpr (kConv) {
            cv_n = m0 / p.cv_ohw;
            const int rem = m0 - cv_n * p.cv_ohw;
            const int oh = rem / p.cv_ow;
            cv_h = oh - p.cv_pad_t;
            cv_w = rem - oh * p.cv_ow - p.cv_pad_l;
          }
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1u, 1);
            const uint32_t lbar = mapa_shared(&full[stage], 0);
            if (leader) mbar_arrive_expect_tx(&full[stage], tx_pair);
            uint8_t* st = smem + stage * stage_bytes;
            uint8_t* sb = st + a_planes * Cfg::kABytes;
            if constexpr (kConv) {
              const uint16_t off_h = static_cast<uint16_t>(cv_kh), off_w = static_cast<uint16_t>(cv_kwi);
              tma_load_im2col_4d_pair(st, &tmA0, lbar, cv_cb * BK, cv_w, cv_h, cv_n, off_w, off_h);
              if (kPl == 2 && !a_lo_zero)
                tma_load_im2col_4d_pair(st + Cfg::kABytes, &tmA1, lbar, cv_cb * BK, cv_w, cv_h, cv_n, off_w, off_h);
              if (++cv_cb == p.cv_cblocks) {
                cv_cb = 0;
                if (++cv_kwi == p.cv_kw) {
                  cv_kwi = 0;
                  ++cv_kh;
                }
              }
              tma_load_2d_pair(sb, &tmB0, lbar, kb * BK, n0, Policy::kHintB);
              if (kPl == 2) tma_load_2d_pair(sb + b_plane_bytes, &tmB1, lbar, kb * BK, n0, Policy::kHintB);
            } else if constexpr (policy_frame_maps<Policy>::value) {  // operands stored P rows per frame, see gemm_sm100.cuh
              tma_load_3d_pair(st, &tmA0, lbar, kb * BK, 0, m0 >> 5, Policy::kHintA);
              if (kPl == 2 && !a_lo_zero)
                tma_load_3d_pair(st + Cfg::kABytes, &tmA1, lbar, kb * BK, 0, m0 >> 5, Policy::kHintA);
              if (p.b_frame_map) {
                tma_load_3d_pair(sb, &tmB0, lbar, kb * BK, 0, n0 >> 5, Policy::kHintB);
                if (kPl == 2) tma_load_3d_pair(sb + b_plane_bytes, &tmB1, lbar, kb * BK, 0, n0 >> 5, Policy::kHintB);
              } else {
                tma_load_2d_pair(sb, &tmB0, lbar, kb * BK, n0, Policy::kHintB);
                if (kPl == 2) tma_load_2d_pair(sb + b_plane_bytes, &tmB1, lbar, kb * BK, n0, Policy::kHintB);
              }
            } else {
              tma_load_2d_pair(st, &tmA0, lbar, kb * BK, m0, Policy::kHintA);
              if (kPl == 2 && !a_lo_zero) tma_load_2d_pair(st + Cfg::kABytes, &tmA1, lbar, kb * BK, m0, Policy::kHintA);
              tma_load_2d_pair(sb, &tmB0, lbar, kb * BK, n0, Policy::kHintB);
              if (kPl == 2) tma_load_2d_pair(sb + b_plane_bytes, &tmB1, lbar, kb * BK, n0, Policy::kHintB);
            }
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    } else if (warp == 1 && leader) {
      if (elect_one_sync()) {
        // ===================== MMA issuer (leader only) =====================
        const uint32_t idesc = make_idesc_f16(2 * kTileM, n_tile, p.ab_fmt);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int i = 0; i < my_tiles; ++i) {
          int kb = 0;
          for (int ch = 0; ch < n_chunks; ++ch) {
            mbar_wait_cluster(&tempty[acc], acc_phase ^ 1u);
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
                const uint32_t koff = k * 32;
                const uint64_t dah = make_smem_desc(a_hi + koff, Cfg::kSBO, Cfg::kSwizzleMode);
                const uint64_t dbh = make_smem_desc(b_hi + koff, Cfg::kSBO, Cfg::kSwizzleMode);
                if (Cfg::NPROD == 3) {
                  const uint64_t dal = make_smem_desc(a_lo + koff, Cfg::kSBO, Cfg::kSwizzleMode);
                  const uint64_t dbl = make_smem_desc(b_lo + koff, Cfg::kSBO, Cfg::kSwizzleMode);
                  if (!a_lo_zero) {
                    umma_f16_pair(d_tmem, dal, dbh, idesc, (kk | k) != 0 ? 1u : 0u);
                    umma_f16_pair(d_tmem, dah, dbl, idesc, 1u);
                  } else {
                    umma_f16_pair(d_tmem, dah, dbl, idesc, (kk | k) != 0 ? 1u : 0u);
                  }
                  umma_f16_pair(d_tmem, dah, dbh, idesc, 1u);
                } else {
                  umma_f16_pair(d_tmem, dah, dbh, idesc, (kk | k) != 0 ? 1u : 0u);
                }
              }
              umma_commit_pair(&empty[stage], 0x3);  // both CTAs' slots are free once these MMAs have read them
              if (++stage == S) {
                stage = 0;
                phase ^= 1u;
              }
            }
            umma_commit_pair(&tfull[acc], 0x3);      // (partial) accumulator complete in both CTAs' TMEM
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int cstride = policy_chunk_stride<Policy>::get(p);
    const int n_cchunks = n_tile / cstride;
    typename Policy::Epilogue epi(p, quarter, half, lane, scratch);
    // this warp group's chunks [c_first, c_first + c_count), at most 4
    const int c_split = kConv ? min(n_cchunks, 2 * ((n_cchunks + 3) / 4)) : min(n_cchunks, 4);
    const int c_first = alt_tiles ? 0 : (half ? c_split : 0);
    const int c_count = alt_tiles ? n_cchunks : (half ? n_cchunks - c_split : c_split);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int i = 0; i < my_tiles; ++i) {
      if (alt_tiles) {
        if ((i & 1) != half) continue;            // the other group's tile
        acc = half;                               // one K chunk per tile: tile i accumulates in buffer i & 1
        acc_phase = static_cast<uint32_t>(i >> 1) & 1u;
      }
      TileCoord tc = Policy::tile_pair(p, cluster, nclusters, i);
      tc.mt += static_cast<int>(rank);
      float sums[128];
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(&tfull[acc], acc_phase, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(acc * kMaxTileN) +
                               (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c_first * cstride);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc < c_count) {
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
        if (lane == 0) mbar_arrive_cluster(mapa_shared(&tempty[acc], 0));  // the leader's MMA thread counts both CTAs
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
      epi.begin_tile(tc);
      using Epi = typename Policy::Epilogue;
      epi_slot_from_regs_at<Epi, 0>(epi, tc, sums, c_first, c_count);
      epi_slot_from_regs_at<Epi, 1>(epi, tc, sums, c_first, c_count);
      epi_slot_from_regs_at<Epi, 2>(epi, tc, sums, c_first, c_count);
      epi_slot_from_regs_at<Epi, 3>(epi, tc, sums, c_first, c_count);
      epi.end_tile(tc);
      epi.post_tile(tc);
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading its TMEM / the leader's MMAs may still target it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

template <class Policy>
inline cudaError_t launch_gemm_pair(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0,
                                    const CUtensorMap& b1, const typename Policy::Params& p, int clusters,
                                    cudaStream_t stream) {
  static_assert(smem_bytes<Policy>() <= kSmemLimit, "exceeds the 227 KB per-CTA shared memory limit");
  static std::atomic<bool> attr_set[kMaxDevices] = {};  // per device, see launch_gemm
  const int dev = current_device();
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel<Policy>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes<Policy>());
    if (e != cudaSuccess) return e;
    attr_set[dev].store(true, std::memory_order_release);
  }
  gemm_pair_kernel<Policy><<<2 * clusters, kGemmThreadsPromote, smem_bytes<Policy>(), stream>>>(a0, a1, b0, b1, p);
  return cudaGetLastError();
}

}  // namespace dlc

// Keyframe database + global matcher: similarity matrix Q x DB^T on tcgen05 tensor cores with the per-row top-k /
// threshold selection fused into the epilogue (scores are consumed straight out of TMEM; the B x N_db similarity
// matrix never exists in HBM). New capability required by the north star; the reference only ever builds dense
// N x N matrices in Python loops (src/sdav/create_similarity_matrix.py:31-38,
// src/cnn_vtl/create_distance_matrix.py:31-36) and has no database, top-k or threshold step.
//
// Mapping: M (accumulator rows / TMEM lanes) = queries, N (accumulator columns) = database rows. An epilogue thread
// owns one query and scans the 256 database columns of each tile, keeping a sorted top-K list in registers for the
// whole kernel. CTAs are grouped per 128-query tile; inside a group the database tiles are strided over the CTAs
// (concurrent CTAs of different groups read the same database tile -> one HBM read, the rest L2 hits).
// Per-CTA partial lists are merged by the deterministic row top-k kernel (score, then lowest index).
#include <cuda_bf16.h>
#include <math.h>

#include <algorithm>

#include "gemm_sm100.cuh"
#include "topk.h"
#include "util.h"

struct dlc_db {
  int dim = 0;
  int ld = 0;
  int64_t capacity = 0;
  int64_t size = 0;
  int metric = DLC_METRIC_COS;
  int dtype = DLC_F16;
  void* rows = nullptr;   // [capacity, ld] fp16 / bf16
  float* sqn = nullptr;   // [capacity] squared norm of the STORED row (L2 metric)
};

namespace dlc {

// ---------------- append: (normalise) + convert + squared norm, one warp per row ----------------
template <typename SrcT>
__device__ __forceinline__ float load_as_float(const SrcT* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}

template <typename SrcT, bool BF16>
__global__ void __launch_bounds__(256)
db_append_kernel(const SrcT* __restrict__ src, int64_t n, int dim, int ld, int normalise, uint16_t* __restrict__ dst,
                 float* __restrict__ sqn) {
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const SrcT* s = src + warp * dim;
  float scale = 1.0f;
  if (normalise) {
    float acc = 0.0f;
    for (int c = lane; c < dim; c += 32) {
      const float x = load_as_float<SrcT>(s, c);
      acc = fmaf(x, x, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    scale = acc > 0.0f ? 1.0f / sqrtf(acc) : 0.0f;
  }
  float n2 = 0.0f;
  uint16_t* d = dst + warp * ld;
  for (int c = lane; c < ld; c += 32) {
    uint16_t bits = 0;
    if (c < dim) {
      const float x = load_as_float<SrcT>(s, c) * scale;
      float stored;
      if (BF16) {
        const __nv_bfloat16 b = __float2bfloat16_rn(x);
        bits = __bfloat16_as_ushort(b);
        stored = __bfloat162float(b);
      } else {
        const __half h = __float2half_rn(x);
        bits = __half_as_ushort(h);
        stored = __half2float(h);
      }
      n2 = fmaf(stored, stored, n2);
    }
    d[c] = bits;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
  if (lane == 0) sqn[warp] = n2;
}

// ---------------- query prep: (normalise) + convert to the operand plane, per-row aux ----------------
// aux[b] = squared norm of the ROUNDED query (L2) so that dist = aux + dn - 2 G is consistent with G.
template <bool BF16>
__global__ void __launch_bounds__(256)
query_prep_kernel(const float* __restrict__ q, int B, int dim, int ld, int metric, uint16_t* __restrict__ plane,
                  float* __restrict__ aux) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* s = q + static_cast<int64_t>(warp) * dim;
  float scale = 1.0f;
  if (metric == DLC_METRIC_COS) {
    float acc = 0.0f;
    for (int c = lane; c < dim; c += 32) acc = fmaf(s[c], s[c], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    scale = acc > 0.0f ? 1.0f / sqrtf(acc) : 0.0f;
  }
  float n2 = 0.0f;
  uint16_t* d = plane + static_cast<int64_t>(warp) * ld;
  for (int c = lane; c < ld; c += 32) {
    uint16_t bits = 0;
    if (c < dim) {
      const float x = s[c] * scale;
      float stored;
      if (BF16) {
        const __nv_bfloat16 b = __float2bfloat16_rn(x);
        bits = __bfloat16_as_ushort(b);
        stored = __bfloat162float(b);
      } else {
        const __half h = __float2half_rn(x);
        bits = __half_as_ushort(h);
        stored = __half2float(h);
      }
      n2 = fmaf(stored, stored, n2);
    }
    d[c] = bits;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
  if (lane == 0) aux[warp] = n2;
}

// ---------------- fused similarity + running top-K ----------------
struct MatchParams {
  int n_tile, k_blocks, ab_fmt, kc;
  int m_tiles;     // query tiles = CTA groups
  int slots;       // CTAs per group
  int n_tiles;     // database tiles
  int B;
  int64_t db_rows;
  int metric;
  const float* dn;        // [db_rows] squared norms (L2)
  const float* qaux;      // [B] squared norms of the rounded queries (L2)
  int use_thr;
  float thr;
  int32_t* counts;        // [B] (threshold mode; zeroed by the host wrapper)
  int64_t idx_offset;
  float* part_scores;     // [B, slots, KMAX] selection keys (larger = better)
  int64_t* part_idx;      // [B, slots, KMAX]
};

template <int BK, int KMAX>
struct MatchPolicy {
  using Cfg = GemmCfg<BK, 1>;
  using Params = MatchParams;
  static constexpr bool kPromote = false;  // operand rounding (fp16/bf16 storage) dominates the error here
  static constexpr int kEpiWarps = 4;
  static __device__ __forceinline__ bool enabled(const Params&) { return true; }
  static constexpr uint64_t kHintA = kEvictLast;   // queries: tiny, re-read for every database tile
  static constexpr uint64_t kHintB = kEvictFirst;  // database: streamed once per pass

  static __device__ __forceinline__ int num_tiles(const Params& p, int cta, int) {
    const int slot = cta / p.m_tiles;
    return slot < p.n_tiles ? (p.n_tiles - slot + p.slots - 1) / p.slots : 0;
  }
  static __device__ __forceinline__ TileCoord tile(const Params& p, int cta, int, int i) {
    TileCoord tc;
    tc.mt = cta % p.m_tiles;
    tc.nt = cta / p.m_tiles + i * p.slots;
    return tc;
  }

  struct Epilogue {
    const Params& p;
    const int quarter, lane;
    float ls[KMAX];
    int li[KMAX];  // database row (local to this database); INT_MAX = empty
    int cnt;
    __device__ Epilogue(const Params& p_, int quarter_, int, int lane_, void*) : p(p_), quarter(quarter_), lane(lane_) {
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        ls[t] = -INFINITY;
        li[t] = 0x7fffffff;
      }
      cnt = 0;
    }

    int row;
    bool l2;
    float gscale, thr_key;
    __device__ __forceinline__ void begin_tile(TileCoord tc) {
      row = tc.mt * kTileM + quarter * 32 + lane;
      l2 = p.metric == DLC_METRIC_L2;
      // selection key (larger = better): COS/DOT: G ; L2: 2 G - |d|^2   (|q|^2 is constant per row)
      gscale = l2 ? 2.0f : 1.0f;
      thr_key = p.thr;
      if (l2) thr_key = (row < p.B ? p.qaux[row] : 0.0f) - p.thr;  // dist <= thr  <=>  key >= |q|^2 - thr
    }
    __device__ __forceinline__ void end_tile(TileCoord) {}
    __device__ __forceinline__ void post_tile(TileCoord) {}

    // One sorted insertion (larger key first; candidates arrive in increasing index, so the strict > keeps the
    // lowest index on ties).
    __device__ __forceinline__ void insert(float cs, int ci) {
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        const bool up = cs > ls[t];
        const float ts = ls[t];
        const int ti = li[t];
        ls[t] = up ? cs : ts;
        li[t] = up ? ci : ti;
        cs = up ? ts : cs;
        ci = up ? ti : ci;
      }
    }

    // 32 scores of this thread's query against 32 consecutive database rows. Fast path (almost every chunk once the
    // list is warm: a new row beats the current K-th best with probability ~K / rows seen): one max over the 32 keys
    // and one compare. Slow path: the keys go through a small local array and ONE rolled insertion loop - the
    // unrolled form (32 x K-step insertions per chunk, 8 chunks) was 29k instructions and ran out of the
    // instruction cache (ncu: `no_inst` was the top stall of the epilogue warps and the tensor pipe sat at 30 %).
    template <int SLOT>
    __device__ __forceinline__ void chunk(TileCoord tc, int c, float (&v)[32]) {
      const int64_t col0 = static_cast<int64_t>(tc.nt) * p.n_tile + c * 32;
      if (col0 >= p.db_rows) return;  // warp-uniform
      const int nvalid = p.db_rows - col0 < 32 ? static_cast<int>(p.db_rows - col0) : 32;
      if (l2) {
        const float my_dn = (col0 + lane < p.db_rows) ? p.dn[col0 + lane] : 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] * 2.0f - __shfl_sync(0xffffffffu, my_dn, j);
      }
      if (nvalid < 32) {  // last database tile only
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j >= nvalid) v[j] = -INFINITY;
      }
      if (row >= p.B) return;
      if (p.use_thr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) cnt += (v[j] >= thr_key && j < nvalid) ? 1 : 0;
      }
      float cmax = v[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, v[j]);
      if (cmax > ls[KMAX - 1]) {
        float kl[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) kl[j] = v[j];
#pragma unroll 1
        for (int j = 0; j < nvalid; ++j) {
          const float key = kl[j];
          if (key > ls[KMAX - 1]) insert(key, static_cast<int>(col0) + j);
        }
      }
    }

    __device__ __forceinline__ void finish() {
      // Every CTA of a group writes its slot (CTAs without tiles write empty lists) so the merge reads defined data.
      const int cta = blockIdx.x;
      const int mt = cta % p.m_tiles, slot = cta / p.m_tiles;
      const int frow = mt * kTileM + quarter * 32 + lane;
      if (frow >= p.B) return;
      const int row = frow;
      const int64_t o = (static_cast<int64_t>(row) * p.slots + slot) * KMAX;
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        p.part_scores[o + t] = ls[t];
        p.part_idx[o + t] = li[t] == 0x7fffffff ? -1 : static_cast<int64_t>(li[t]) + p.idx_offset;
      }
      if (p.use_thr && cnt) atomicAdd(p.counts + row, cnt);
    }
  };
};

__global__ void threshold_finalize_kernel(int B, int k, float thr, int smaller_is_better, float* scores, int64_t* idx) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * k) return;
  const float s = scores[t];
  const bool pass = idx[t] >= 0 && (smaller_is_better ? s <= thr : s >= thr);
  if (!pass) {
    scores[t] = smaller_is_better ? INFINITY : -INFINITY;
    idx[t] = -1;
  }
}

struct MatchLayout {
  size_t off_q, off_aux, off_ps, off_pi, total;
  int ld, m_tiles, slots, kmax;
};
static MatchLayout match_layout(const dlc_db* db, int B, int k) {
  MatchLayout L{};
  L.ld = db->ld;
  L.m_tiles = ceil_div(B, kTileM);
  L.slots = std::max(1, sm_count() / L.m_tiles);
  L.kmax = k <= 16 ? 16 : 32;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  L.off_q = take(static_cast<size_t>(B) * L.ld * 2);
  L.off_aux = take(sizeof(float) * B);
  L.off_ps = take(sizeof(float) * static_cast<size_t>(B) * L.slots * L.kmax);
  L.off_pi = take(sizeof(int64_t) * static_cast<size_t>(B) * L.slots * L.kmax);
  L.total = o;
  return L;
}

template <class Policy>
static int run_match(const dlc_db* db, const MatchLayout& L, char* ws, MatchParams p, cudaStream_t stream) {
  constexpr int BK = Policy::Cfg::BK;
  CUtensorMap ta, tb;
  if (!make_tmap_k_major(&ta, ws + L.off_q, p.ab_fmt, L.ld, p.B, L.ld, BK, kTileM) ||
      !make_tmap_k_major(&tb, db->rows, p.ab_fmt, L.ld, db->size, L.ld, BK, p.n_tile))
    return fail(DLC_ECUDA, "dlc_match: cuTensorMapEncodeTiled failed");
  p.k_blocks = L.ld / BK;
  const int grid = L.m_tiles * L.slots;
  cudaError_t e = launch_gemm<Policy>(ta, ta, tb, tb, p, grid, stream);
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_match: launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

static int match_impl(dlc_db* db, const float* q, int B, int k, int64_t idx_offset, int use_thr, float thr,
                      int32_t* counts, float* scores, int64_t* idx, void* ws_dev, size_t ws_bytes,
                      cudaStream_t s) {
  const MatchLayout L = match_layout(db, B, k);
  if (ws_bytes < L.total || !ws_dev)
    return fail(DLC_ENOMEM, "dlc_match: workspace of %zu bytes needed, %zu given", L.total, ws_bytes);
  if ((reinterpret_cast<uintptr_t>(ws_dev) & 255) != 0) return fail(DLC_EINVAL, "dlc_match: workspace must be 256-byte aligned");
  if (L.m_tiles > sm_count()) return fail(DLC_EUNSUPPORTED, "dlc_match: at most %d queries per call", sm_count() * kTileM);
  char* ws = static_cast<char*>(ws_dev);
  const bool bf16 = db->dtype == DLC_BF16;
  const bool l2 = db->metric == DLC_METRIC_L2;
  float* aux = reinterpret_cast<float*>(ws + L.off_aux);
  if (bf16)
    query_prep_kernel<true><<<ceil_div(B, 8), 256, 0, s>>>(q, B, db->dim, L.ld, db->metric,
                                                           reinterpret_cast<uint16_t*>(ws + L.off_q), aux);
  else
    query_prep_kernel<false><<<ceil_div(B, 8), 256, 0, s>>>(q, B, db->dim, L.ld, db->metric,
                                                            reinterpret_cast<uint16_t*>(ws + L.off_q), aux);
  DLC_CUDA(cudaGetLastError());
  if (use_thr) DLC_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * B, s));

  if (db->size == 0) {  // empty database: every list is padding
    const int64_t n = static_cast<int64_t>(B) * k;
    DLC_CUDA(cudaMemsetAsync(idx, 0xff, sizeof(int64_t) * n, s));  // -1
    threshold_finalize_kernel<<<ceil_div(static_cast<int>(n), 256), 256, 0, s>>>(B, k, 0.0f, l2, scores, idx);
    DLC_CUDA(cudaGetLastError());
    return DLC_OK;
  }

  MatchParams p{};
  p.n_tile = kMaxTileN;
  p.ab_fmt = bf16 ? 1 : 0;
  p.m_tiles = L.m_tiles;
  p.slots = L.slots;
  p.n_tiles = static_cast<int>(ceil_div64(db->size, kMaxTileN));
  p.B = B;
  p.db_rows = db->size;
  p.metric = db->metric;
  p.dn = db->sqn;
  p.qaux = aux;
  p.use_thr = use_thr;
  p.thr = thr;
  p.counts = counts;
  p.idx_offset = idx_offset;
  p.part_scores = reinterpret_cast<float*>(ws + L.off_ps);
  p.part_idx = reinterpret_cast<int64_t*>(ws + L.off_pi);
  int rc;
  if (L.kmax == 16) rc = run_match<MatchPolicy<64, 16>>(db, L, ws, p, s);
  else rc = run_match<MatchPolicy<64, 32>>(db, L, ws, p, s);
  if (rc != DLC_OK) return rc;
  // merge the per-CTA partial lists; L2: reported score = |q|^2 - key = squared distance
  const int cols = L.slots * L.kmax;
  rc = merge_partials_impl(p.part_scores, p.part_idx, B, cols, cols, k, /*largest=*/1, l2 ? aux : nullptr,
                           l2 ? -1.0f : 1.0f, scores, idx, s);
  if (rc != DLC_OK) return rc;
  if (l2 || use_thr) {
    // L2 padding convention (+inf) and threshold filtering of the listed entries
    threshold_finalize_kernel<<<ceil_div(B * k, 256), 256, 0, s>>>(
        B, k, use_thr ? thr : (l2 ? INFINITY : -INFINITY), l2, scores, idx);
    DLC_CUDA(cudaGetLastError());
  }
  return DLC_OK;
}

// ---- row-sharded matcher: merge of the per-rank lists, read straight from the all-gathered blocks
// block r = [idx B x k int64 | scores B x k float32] of rank r. One warp per query: every lane keeps its share of the
// world * k candidates, k rounds of warp arg-best (score, then lowest index; idx -1 = padding loses to everything).
__global__ void __launch_bounds__(256)
merge_rank_lists_kernel(const uint8_t* __restrict__ blocks, size_t block_bytes, int world, int B, int k, int smaller,
                        float* __restrict__ out_s, int64_t* __restrict__ out_i) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  constexpr int kPerLane = 32;                  // world * k <= 1024
  const int n = world * k;
  float sc[kPerLane];
  int64_t ix[kPerLane];
  const float worst = smaller ? INFINITY : -INFINITY;
#pragma unroll
  for (int t = 0; t < kPerLane; ++t) {
    const int c = lane + 32 * t;
    sc[t] = worst;
    ix[t] = -1;
    if (c < n) {
      const int r = c / k, j = c - r * k;
      const uint8_t* blk = blocks + static_cast<size_t>(r) * block_bytes;
      ix[t] = reinterpret_cast<const int64_t*>(blk)[static_cast<size_t>(b) * k + j];
      sc[t] = reinterpret_cast<const float*>(blk + sizeof(int64_t) * static_cast<size_t>(B) * k)[static_cast<size_t>(b) * k + j];
    }
  }
  auto better = [&](float s1, int64_t i1, float s2, int64_t i2) {  // is candidate 1 strictly better than candidate 2
    if (i1 < 0) return false;
    if (i2 < 0) return true;
    if (s1 != s2) return smaller ? s1 < s2 : s1 > s2;
    return i1 < i2;
  };
  for (int round = 0; round < k; ++round) {
    float bs = worst;
    int64_t bi = -1;
    int bt = -1;
#pragma unroll
    for (int t = 0; t < kPerLane; ++t)
      if (better(sc[t], ix[t], bs, bi)) {
        bs = sc[t];
        bi = ix[t];
        bt = t;
      }
    float ws_ = bs;
    int64_t wi = bi;
    int wl = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ws_, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, wi, off);
      const int ol = __shfl_xor_sync(0xffffffffu, wl, off);
      if (better(os, oi, ws_, wi) || (oi == wi && os == ws_ && ol < wl)) {
        ws_ = os;
        wi = oi;
        wl = ol;
      }
    }
    if (lane == 0) {
      out_s[static_cast<size_t>(b) * k + round] = wi < 0 ? worst : ws_;
      out_i[static_cast<size_t>(b) * k + round] = wi;
    }
    if (lane == wl && bt >= 0 && wi >= 0) {     // the winner leaves the pool
#pragma unroll
      for (int t = 0; t < kPerLane; ++t)
        if (t == bt) ix[t] = -1;
    }
  }
}

int comm_all_gather(dlc_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t stream);  // comm.cu
int comm_rank(const dlc_comm* c);
int comm_world(const dlc_comm* c);

}  // namespace dlc

using namespace dlc;

extern "C" size_t dlc_match_sharded_workspace_bytes(const dlc_db* db, int B, int k, int world) {
  if (!db || B <= 0 || k <= 0 || k > 32 || world <= 0) return 0;
  const size_t block = align_up(static_cast<size_t>(B) * k * 12, 256);
  return align_up(match_layout(db, B, k).total, 256) + block * (static_cast<size_t>(world) + 1);
}

// Row-sharded top-k over the `world` ranks of `comm` (one process per GPU; every rank calls this with the SAME
// queries): fused similarity + top-k on the local shard -> ONE ncclAllGather of the packed (score, index) lists on
// `stream` -> merge kernel reading the gathered blocks in place. Every rank ends with the same lists.
extern "C" int dlc_match_topk_sharded(dlc_db* db, dlc_comm* comm, const float* q_dev, int B, int k, int64_t idx_offset,
                                      float* scores_dev, int64_t* idx_dev, void* ws_dev, size_t ws_bytes,
                                      void* stream) {
  DLC_CHECK_ARG(db && comm && q_dev && scores_dev && idx_dev && ws_dev);
  DLC_CHECK_ARG(B >= 1 && k >= 1 && k <= 32);
  const int world = comm_world(comm), rank = comm_rank(comm);
  DLC_CHECK_ARG(world * k <= 1024);
  if (ws_bytes < dlc_match_sharded_workspace_bytes(db, B, k, world))
    return fail(DLC_ENOMEM, "dlc_match_topk_sharded: workspace of %zu bytes needed, %zu given",
                dlc_match_sharded_workspace_bytes(db, B, k, world), ws_bytes);
  cudaStream_t s = as_stream(stream);
  char* ws = static_cast<char*>(ws_dev);
  const size_t local_ws = align_up(match_layout(db, B, k).total, 256);
  const size_t block = align_up(static_cast<size_t>(B) * k * 12, 256);
  uint8_t* gathered = reinterpret_cast<uint8_t*>(ws + local_ws);              // world blocks
  uint8_t* mine = gathered + static_cast<size_t>(rank) * block;               // in-place all-gather: own slot
  int64_t* my_i = reinterpret_cast<int64_t*>(mine);
  float* my_s = reinterpret_cast<float*>(mine + sizeof(int64_t) * static_cast<size_t>(B) * k);
  if (int rc = match_impl(db, q_dev, B, k, idx_offset, 0, 0.0f, nullptr, my_s, my_i, ws, local_ws, s)) return rc;
  if (world > 1) {
    if (int rc = comm_all_gather(comm, mine, gathered, block, s)) return rc;
  }
  merge_rank_lists_kernel<<<ceil_div(B, 8), 256, 0, s>>>(gathered, block, world, B, k,
                                                         db->metric == DLC_METRIC_L2 ? 1 : 0, scores_dev, idx_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_db_create(dlc_db** db, int dim, int64_t capacity_rows, int metric, int dtype) {
  DLC_CHECK_ARG(db);
  DLC_CHECK_ARG(dim > 0 && capacity_rows > 0);
  DLC_CHECK_ARG(metric == DLC_METRIC_COS || metric == DLC_METRIC_DOT || metric == DLC_METRIC_L2);
  DLC_CHECK_ARG(dtype == DLC_F16 || dtype == DLC_BF16);
  if (int rc = dlc_device_check()) return rc;
  dlc_db* d = new dlc_db();
  d->dim = dim;
  d->ld = dlc_plane_ld(dim);
  d->capacity = capacity_rows;
  d->metric = metric;
  d->dtype = dtype;
  cudaError_t e = cudaMalloc(&d->rows, static_cast<size_t>(capacity_rows) * d->ld * 2);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d->sqn), sizeof(float) * capacity_rows);
  if (e != cudaSuccess) {
    dlc_db_destroy(d);
    return fail(DLC_ENOMEM, "dlc_db_create: cudaMalloc of %lld rows x %d failed: %s",
                static_cast<long long>(capacity_rows), dim, cudaGetErrorString(e));
  }
  *db = d;
  return DLC_OK;
}

extern "C" int dlc_db_destroy(dlc_db* db) {
  if (!db) return DLC_OK;
  if (db->rows) cudaFree(db->rows);
  if (db->sqn) cudaFree(db->sqn);
  delete db;
  return DLC_OK;
}

extern "C" int64_t dlc_db_size(const dlc_db* db) { return db ? db->size : 0; }

extern "C" int dlc_db_clear(dlc_db* db) {
  DLC_CHECK_ARG(db);
  db->size = 0;
  return DLC_OK;
}

extern "C" int dlc_db_append(dlc_db* db, const void* rows_dev, int src_dtype, int64_t n, void* stream) {
  DLC_CHECK_ARG(db && (rows_dev || n == 0));
  DLC_CHECK_ARG(n >= 0);
  DLC_CHECK_ARG(src_dtype == DLC_F32 || src_dtype == DLC_F16 || src_dtype == DLC_BF16);
  if (db->size + n > db->capacity)
    return fail(DLC_ENOMEM, "dlc_db_append: %lld + %lld rows exceed the capacity %lld", static_cast<long long>(db->size),
                static_cast<long long>(n), static_cast<long long>(db->capacity));
  if (n == 0) return DLC_OK;
  cudaStream_t s = as_stream(stream);
  uint16_t* dst = static_cast<uint16_t*>(db->rows) + db->size * db->ld;
  float* sqn = db->sqn + db->size;
  const int normalise = db->metric == DLC_METRIC_COS;
  const int grid = static_cast<int>(ceil_div64(n, 8));
  const bool bf = db->dtype == DLC_BF16;
#define DLC_APPEND(SRC)                                                                                               \
  do {                                                                                                                \
    if (bf) db_append_kernel<SRC, true><<<grid, 256, 0, s>>>(static_cast<const SRC*>(rows_dev), n, db->dim, db->ld,    \
                                                              normalise, dst, sqn);                                   \
    else db_append_kernel<SRC, false><<<grid, 256, 0, s>>>(static_cast<const SRC*>(rows_dev), n, db->dim, db->ld,      \
                                                            normalise, dst, sqn);                                     \
  } while (0)
  if (src_dtype == DLC_F32) DLC_APPEND(float);
  else if (src_dtype == DLC_F16) DLC_APPEND(__half);
  else DLC_APPEND(__nv_bfloat16);
#undef DLC_APPEND
  DLC_CUDA(cudaGetLastError());
  db->size += n;
  return DLC_OK;
}

extern "C" size_t dlc_match_workspace_bytes(const dlc_db* db, int B, int k) {
  if (!db || B <= 0 || k <= 0 || k > 32) return 0;
  return match_layout(db, B, k).total;
}

extern "C" int dlc_match_topk(dlc_db* db, const float* q_dev, int B, int k, int64_t idx_offset, float* scores_dev,
                              int64_t* idx_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(db && q_dev && scores_dev && idx_dev);
  DLC_CHECK_ARG(B >= 1);
  DLC_CHECK_ARG(k >= 1 && k <= 32);
  return match_impl(db, q_dev, B, k, idx_offset, 0, 0.0f, nullptr, scores_dev, idx_dev, ws_dev, ws_bytes,
                    as_stream(stream));
}

extern "C" int dlc_match_threshold(dlc_db* db, const float* q_dev, int B, float thr, int max_per_row,
                                   int64_t idx_offset, int32_t* counts_dev, float* scores_dev, int64_t* idx_dev,
                                   void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(db && q_dev && counts_dev && scores_dev && idx_dev);
  DLC_CHECK_ARG(B >= 1);
  DLC_CHECK_ARG(max_per_row >= 1 && max_per_row <= 32);
  return match_impl(db, q_dev, B, max_per_row, idx_offset, 1, thr, counts_dev, scores_dev, idx_dev, ws_dev, ws_bytes,
                    as_stream(stream));
}

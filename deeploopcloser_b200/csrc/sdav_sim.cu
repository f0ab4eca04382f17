// a5/a6: SDAV frame-pair similarity matrix.
// Replaces SimilarityCalculator.similarity_score and helpers (src/sdav/similarity/SimilarityCalculator.py:12-49)
// plus the i<j double loop of src/sdav/create_similarity_matrix.py:31-38.
//
// Reference per pair (h1 = frame i, h2 = frame j, both [P, D]):
//   w      = exp(-(mean_over_all_rows(dataset) - mu)^2 / (2 sigma^2))                  (:19-27)
//   j*(k)  = argmin_j || h2[j] - h1[k] ||_2     (first minimum)                         (:29-37)
//   s_k    = | w . (h1[k] - h2[j*(k)]) |                                                (:39-45)
//   S      = sum_k (a + b ln s_k)                                                       (:47-49)
// Restated as ONE Gram contraction over all patch rows: G = H H^T,  ||h2[j]-h1[k]||^2 = n_k + n_j - 2 G[k,j]
// (n_k is constant inside the argmin), and w.(h1[k]-h2[j]) = p[i,k] - p[j,j] with p = H w (one GEMV).
// Frames are padded to 32 rows, so in a 128 x 256 accumulator tile each epilogue warp owns exactly one frame i and
// every 32-column TMEM chunk is exactly one frame j: the argmin is a per-thread scan of 32 registers, the sum over
// k a warp shuffle reduction. The (30N)^2 Gram matrix never leaves TMEM.
#include <math.h>

#include <algorithm>
#include <vector>

#include "gemm_sm100.cuh"
#include "util.h"

namespace dlc {

constexpr int kFrameRows = 32;                          // padded patch rows per frame
constexpr int kFramesPerMTile = kTileM / kFrameRows;    // 4
constexpr int kFramesPerNTile = kMaxTileN / kFrameRows; // 8
constexpr int kMGroup = 8;                              // M tiles per L2 super-block of the tile order

// ---------------- column mean -> distinctive weights (deterministic two-stage reduction) ----------------
constexpr int kColSumSlabs = 128;
__global__ void colsum_partial_kernel(const float* __restrict__ H, int64_t rows, int D, double* __restrict__ part) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= D) return;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * per;
  const int64_t r1 = r0 + per < rows ? r0 + per : rows;
  double acc = 0.0;
  for (int64_t r = r0; r < r1; ++r) acc += static_cast<double>(H[r * D + col]);
  part[static_cast<int64_t>(blockIdx.y) * D + col] = acc;
}
__global__ void weights_kernel(const double* __restrict__ part, int slabs, int64_t rows, int D, double mu, double sigma,
                               double* __restrict__ w) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= D) return;
  double acc = 0.0;
  for (int s = 0; s < slabs; ++s) acc += part[static_cast<int64_t>(s) * D + col];
  const double mean = acc / static_cast<double>(rows);
  const double d = mean - mu;
  w[col] = exp(-(d * d) / (2.0 * sigma * sigma));
}
// per patch row: squared norm (float) and projection p = h . w (double); one warp per row, padded-row indexing
__global__ void __launch_bounds__(256)
rowstats_kernel(const float* __restrict__ H, int N, int P, int D, const double* __restrict__ w,
                float* __restrict__ sqn, double* __restrict__ pw) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N * kFrameRows) return;
  const int f = warp / kFrameRows, k = warp % kFrameRows;
  double n2 = 0.0, pr = 0.0;
  if (k < P) {
    const float* h = H + (static_cast<int64_t>(f) * P + k) * D;
    for (int c = lane; c < D; c += 32) {
      const double x = static_cast<double>(h[c]);
      n2 += x * x;
      pr += x * w[c];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    pr += __shfl_xor_sync(0xffffffffu, pr, off);
  }
  if (lane == 0) {
    sqn[warp] = k < P ? static_cast<float>(n2) : INFINITY;
    pw[warp] = pr;
  }
}

// ---------------- Gram + argmin + score epilogue ----------------
extern int g_promote_k;  // planes.cu

struct GramParams {
  int n_tile, k_blocks, ab_fmt, kc;
  const int2* tiles;  // (mt, nt) work list in L2-friendly order
  int num_tiles;
  int N, P;
  const float* sqn;   // [N*32]
  const double* pw;   // [N*32]
  float a, b;
  int full;           // 1: every ordered pair i != j; 0: i < j, mirrored
  float* S;           // [N, N]
};

template <int BK, int NPROD>
struct GramPolicy {
  using Cfg = GemmCfg<BK, NPROD>;
  using Params = GramParams;
  static constexpr bool kPromote = NPROD == 3;
  static constexpr uint64_t kHintA = kEvictNormal;
  static constexpr uint64_t kHintB = kEvictNormal;

  static __device__ __forceinline__ int num_tiles(const Params& p, int cta, int ncta) {
    return cta < p.num_tiles ? (p.num_tiles - cta + ncta - 1) / ncta : 0;
  }
  static __device__ __forceinline__ TileCoord tile(const Params& p, int cta, int ncta, int i) {
    const int2 t = __ldg(p.tiles + cta + i * ncta);
    TileCoord tc;
    tc.mt = t.x;
    tc.nt = t.y;
    return tc;
  }

  struct Epilogue {
    const Params& p;
    const int quarter, lane;
    __device__ Epilogue(const Params& p_, int quarter_, int lane_, void*) : p(p_), quarter(quarter_), lane(lane_) {}

    int fa;
    bool fa_ok;
    double pa;
    __device__ __forceinline__ void begin_tile(TileCoord tc) {
      fa = tc.mt * kFramesPerMTile + quarter;  // this warp's frame i (lane = patch k)
      fa_ok = fa < p.N;
      pa = fa_ok ? p.pw[fa * kFrameRows + lane] : 0.0;
    }
    __device__ __forceinline__ void end_tile(TileCoord) {}

    // chunk c = the 32 accumulator columns of frame j = 8 nt + c
    __device__ __forceinline__ void chunk(TileCoord tc, int c, float (&v)[32]) {
      const int fb = tc.nt * kFramesPerNTile + c;
      if (!fa_ok || fb >= p.N) return;  // warp-uniform
      if (fa == fb) {
        if (lane == 0) p.S[static_cast<int64_t>(fa) * p.N + fa] = -1.0f;  // reference fill value (:31)
        return;
      }
      if (!p.full && fa > fb) return;
      const float my_nb = p.sqn[fb * kFrameRows + lane];
      const double my_pb = p.pw[fb * kFrameRows + lane];
      float best = INFINITY;
      int bj = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float nb = __shfl_sync(0xffffffffu, my_nb, j);
        // squared distance up to the per-row constant n_k; pad columns carry nb = +inf and never win
        const float d = fmaf(-2.0f, v[j], nb);
        if (d < best) {  // strict: first minimum wins, like np.argmin
          best = d;
          bj = j;
        }
      }
      const double pb = __shfl_sync(0xffffffffu, my_pb, bj);
      const float s = static_cast<float>(fabs(pa - pb));
      float val = lane < p.P ? p.a + p.b * logf(s) : 0.0f;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off);
      if (lane == 0) {
        p.S[static_cast<int64_t>(fa) * p.N + fb] = val;
        if (!p.full) p.S[static_cast<int64_t>(fb) * p.N + fa] = val;
      }
    }
    __device__ __forceinline__ void finish() {}
  };
};

// Work list: super-blocks of kMGroup M tiles; inside a super-block N tile outermost so the ~148 concurrently
// running tiles touch kMGroup A row-blocks and ~148/kMGroup B row-blocks (fits L2) instead of streaming all of H.
static void build_tile_list(int N, int full, std::vector<int2>& out) {
  const int m_tiles = ceil_div(N, kFramesPerMTile), n_tiles = ceil_div(N, kFramesPerNTile);
  out.clear();
  for (int g0 = 0; g0 < m_tiles; g0 += kMGroup) {
    const int g1 = std::min(g0 + kMGroup, m_tiles);
    for (int nt = 0; nt < n_tiles; ++nt)
      for (int mt = g0; mt < g1; ++mt) {
        const int fa_min = mt * kFramesPerMTile;
        const int fb_max = std::min(nt * kFramesPerNTile + kFramesPerNTile - 1, N - 1);
        // upper-triangle mode needs a pair fa < fb, or the diagonal block (to write the -1 fill)
        if (full || fb_max >= fa_min) out.push_back(make_int2(mt, nt));
      }
  }
}

struct SimWorkspace {
  size_t off_hi, off_lo, off_part, off_w, off_sqn, off_pw, off_tiles, total;
  int ld, rows_pad, max_tiles;
};
static SimWorkspace sim_layout(int N, int P, int D) {
  SimWorkspace w{};
  w.ld = dlc_plane_ld(D);
  w.rows_pad = N * kFrameRows;
  w.max_tiles = ceil_div(N, kFramesPerMTile) * ceil_div(N, kFramesPerNTile);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  const size_t plane = static_cast<size_t>(w.rows_pad) * w.ld * 2;
  w.off_hi = take(plane);
  w.off_lo = take(plane);
  w.off_part = take(sizeof(double) * kColSumSlabs * D);
  w.off_w = take(sizeof(double) * D);
  w.off_sqn = take(sizeof(float) * w.rows_pad);
  w.off_pw = take(sizeof(double) * w.rows_pad);
  w.off_tiles = take(sizeof(int2) * w.max_tiles);
  w.total = o;
  (void)P;
  return w;
}

template <class Policy>
static int run_gram(const SimWorkspace& L, char* ws, GramParams p, cudaStream_t stream) {
  constexpr int BK = Policy::Cfg::BK;
  CUtensorMap ta0, ta1, tb0, tb1;
  const void* hi = ws + L.off_hi;
  const void* lo = ws + L.off_lo;
  if (!make_tmap_k_major(&ta0, hi, 0, L.ld, L.rows_pad, L.ld, BK, kTileM) ||
      !make_tmap_k_major(&tb0, hi, 0, L.ld, L.rows_pad, L.ld, BK, kMaxTileN) ||
      !make_tmap_k_major(&ta1, lo, 0, L.ld, L.rows_pad, L.ld, BK, kTileM) ||
      !make_tmap_k_major(&tb1, lo, 0, L.ld, L.rows_pad, L.ld, BK, kMaxTileN))
    return fail(DLC_ECUDA, "dlc_sdav_similarity: cuTensorMapEncodeTiled failed");
  p.k_blocks = L.ld / BK;
  p.kc = std::max(1, g_promote_k / BK);
  const int grid = std::min(p.num_tiles, sm_count());
  cudaError_t e = launch_gemm<Policy>(ta0, ta1, tb0, tb1, p, grid, stream);
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_sdav_similarity: launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

}  // namespace dlc

using namespace dlc;

// Developer switch (dlc_debug_set key 1): launch only the Gram/score kernel, reusing the operand planes, statistics
// and tile list a previous full call left in the workspace. Lets bench.py time that kernel alone with CUDA events.
static int g_gram_only = 0;
extern "C" int dlc_sdav_debug_gram_only(int on) {
  g_gram_only = on ? 1 : 0;
  return DLC_OK;
}

extern "C" size_t dlc_sdav_similarity_workspace_bytes(int N, int P, int D) {
  if (N <= 0 || P <= 0 || D <= 0) return 0;
  return sim_layout(N, P, D).total;
}

extern "C" int dlc_sdav_weights(const float* desc_dev, int N, int P, int D, double mu, double sigma, double* w_dev,
                                void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(desc_dev && w_dev && ws_dev);
  DLC_CHECK_ARG(N >= 1 && P >= 1 && D >= 1 && sigma != 0.0);
  if (ws_bytes < sizeof(double) * kColSumSlabs * D)
    return fail(DLC_ENOMEM, "dlc_sdav_weights: workspace of %zu bytes needed", sizeof(double) * kColSumSlabs * D);
  cudaStream_t s = as_stream(stream);
  double* part = static_cast<double*>(ws_dev);
  const int64_t rows = static_cast<int64_t>(N) * P;
  colsum_partial_kernel<<<dim3(ceil_div(D, 128), kColSumSlabs), 128, 0, s>>>(desc_dev, rows, D, part);
  weights_kernel<<<ceil_div(D, 128), 128, 0, s>>>(part, kColSumSlabs, rows, D, mu, sigma, w_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_sdav_similarity(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a,
                                   double b, const double* w_dev, int precision, int full_asymmetric,
                                   float* S_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(desc_dev && S_dev && ws_dev);
  DLC_CHECK_ARG(N >= 1 && N <= (1 << 20));
  DLC_CHECK_ARG(P >= 1 && P <= kFrameRows);
  DLC_CHECK_ARG(D >= 1);
  DLC_CHECK_ARG(sigma != 0.0);
  DLC_CHECK_ARG(precision == DLC_PREC_FP16 || precision == DLC_PREC_FP16X2);
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0);
  const SimWorkspace L = sim_layout(N, P, D);
  if (ws_bytes < L.total)
    return fail(DLC_ENOMEM, "dlc_sdav_similarity: workspace of %zu bytes needed, %zu given", L.total, ws_bytes);
  cudaStream_t s = as_stream(stream);
  char* ws = static_cast<char*>(ws_dev);
  const int64_t rows = static_cast<int64_t>(N) * P;

  double* part = reinterpret_cast<double*>(ws + L.off_part);
  double* w = reinterpret_cast<double*>(ws + L.off_w);
  float* sqn = reinterpret_cast<float*>(ws + L.off_sqn);
  double* pw = reinterpret_cast<double*>(ws + L.off_pw);
  static thread_local std::vector<int2> tiles;
  if (!g_gram_only) {
  // 1. operand planes: frames padded P -> 32 rows, K padded to a multiple of 64 (zeros)
  if (int rc = dlc_split_planes(desc_dev, DLC_F32, static_cast<int>(rows), D, D, P, kFrameRows, ws + L.off_hi,
                                ws + L.off_lo, L.ld, stream))
    return rc;
  // 2. dataset mean -> distinctive weights w; 3. per-row squared norms and projections p = h . w
  if (w_dev) {  // weights of another dataset (SimilarityCalculator.similarity_score on frames outside it)
    w = const_cast<double*>(w_dev);
  } else {
    colsum_partial_kernel<<<dim3(ceil_div(D, 128), kColSumSlabs), 128, 0, s>>>(desc_dev, rows, D, part);
    weights_kernel<<<ceil_div(D, 128), 128, 0, s>>>(part, kColSumSlabs, rows, D, mu, sigma, w);
  }
  rowstats_kernel<<<ceil_div(N * kFrameRows, 8), 256, 0, s>>>(desc_dev, N, P, D, w, sqn, pw);
  DLC_CUDA(cudaGetLastError());

  // 4. tile work list (host-built, tiny) -> device
  build_tile_list(N, full_asymmetric, tiles);
  DLC_CUDA(cudaMemcpyAsync(ws + L.off_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, s));
  }  // !g_gram_only

  // 5. Gram + argmin + score
  GramParams p{};
  p.n_tile = kMaxTileN;
  p.ab_fmt = 0;
  p.tiles = reinterpret_cast<const int2*>(ws + L.off_tiles);
  p.num_tiles = static_cast<int>(tiles.size());
  p.N = N;
  p.P = P;
  p.sqn = sqn;
  p.pw = pw;
  p.a = static_cast<float>(a);
  p.b = static_cast<float>(b);
  p.full = full_asymmetric ? 1 : 0;
  p.S = S_dev;
  if (precision == DLC_PREC_FP16X2) return run_gram<GramPolicy<32, 3>>(L, ws, p, s);
  return run_gram<GramPolicy<64, 1>>(L, ws, p, s);
}

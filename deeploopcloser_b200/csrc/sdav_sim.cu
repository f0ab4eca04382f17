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
#include <atomic>
#include <mutex>
#include <vector>

#include "gemm_pair_sm100.cuh"
#include "util.h"

namespace dlc {

constexpr int kFrameRows = 32;                          // padded patch rows per frame
constexpr int kFramesPerMTile = kTileM / kFrameRows;    // 4
constexpr int kFramesPerNTile = kMaxTileN / kFrameRows; // 8
std::atomic<int> g_sim_mgroup{32};                                  // M tiles per L2 super-block of the tile order (dlc_debug_set key 4)

// ---------------- column mean -> distinctive weights (deterministic two-stage reduction) ----------------
constexpr int kColSumSlabs = 128;
// part[slab][0..D) = sum_r x, part[slab][D..2D) = sum_r x^2 over the slab's rows (float64). Block = 32 column lanes
// x 8 row lanes: eight independent row streams per column keep enough loads in flight; the eight partial sums are
// combined in a fixed order (deterministic).
constexpr int kColSumRowLanes = 8;
__global__ void __launch_bounds__(32 * kColSumRowLanes)
colsum_partial_kernel(const float* __restrict__ H, int64_t rows, int D, double* __restrict__ part) {
  __shared__ double s_a[kColSumRowLanes][33], s_b[kColSumRowLanes][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = blockIdx.y * per;
  const int64_t r1 = r0 + per < rows ? r0 + per : rows;
  double acc = 0.0, acc2 = 0.0;
  if (col < D) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += kColSumRowLanes) {
      const double x = static_cast<double>(H[r * D + col]);
      acc += x;
      acc2 = fma(x, x, acc2);
    }
  }
  s_a[ty][tx] = acc;
  s_b[ty][tx] = acc2;
  __syncthreads();
  if (ty == 0 && col < D) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < kColSumRowLanes; ++i) {
      a += s_a[i][tx];
      b += s_b[i][tx];
    }
    part[static_cast<int64_t>(blockIdx.y) * 2 * D + col] = a;
    part[static_cast<int64_t>(blockIdx.y) * 2 * D + D + col] = b;
  }
}
// width = 2 D: the column sums and the column sums of squares of all slabs -> one row
__global__ void colsum_reduce_kernel(const double* __restrict__ part, int slabs, int width, double* __restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= width) return;
  double acc = 0.0;
  for (int s = 0; s < slabs; ++s) acc += part[static_cast<int64_t>(s) * width + col];
  out[col] = acc;
}
// Column sums -> dataset mean -> distinctive weights w (SimilarityCalculator.py:19-27) and the CENTRING vector of the
// operand planes (see prep_rows_kernel). One block: the centring decision needs the total over all columns.
//   centre = mean  when the descriptors nearly coincide: sum ||h - mean||^2 < sum ||h||^2 / 16
//          = 0     otherwise
// Why a decision: centring is what makes n_k + n_j - 2G resolvable in fp32 when the rows are close together (a
// trained-like encoder: row norms^2 ~ 7e2, distances^2 ~ 1e-4), but on spread-out, saturated descriptors (N(0,1)
// weights: 89 % exact 0 / 1 values, centred energy 1/6 of the total) the uncentred planes are BETTER operands: 0 and 1
// are exact in fp16 (no rounding error at all on those elements) and all-zero / all-one mantissas keep the tensor
// pipe's power down, i.e. its clock up (measured: 1.46 -> 1.65 ms for the same Gram kernel with centred planes).
constexpr int kWeightsThreads = 1024;
__global__ void __launch_bounds__(kWeightsThreads)
weights_centre_kernel(const double* __restrict__ part, int slabs, int64_t rows, int D, double mu, double sigma,
                      double* __restrict__ w, float* __restrict__ centre) {
  __shared__ double s_e[kWeightsThreads / 32], s_c[kWeightsThreads / 32];
  __shared__ int s_flag;
  double e_tot = 0.0, e_cen = 0.0;
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    double acc = 0.0, acc2 = 0.0;
    for (int s = 0; s < slabs; ++s) {
      acc += part[static_cast<int64_t>(s) * 2 * D + col];
      acc2 += part[static_cast<int64_t>(s) * 2 * D + D + col];
    }
    const double mean = acc / static_cast<double>(rows);
    const double d = mean - mu;
    if (w) w[col] = exp(-(d * d) / (2.0 * sigma * sigma));
    if (centre) centre[col] = static_cast<float>(mean);
    e_tot += acc2;
    e_cen += fmax(acc2 - static_cast<double>(rows) * mean * mean, 0.0);
  }
  if (!centre) return;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    e_tot += __shfl_xor_sync(0xffffffffu, e_tot, off);
    e_cen += __shfl_xor_sync(0xffffffffu, e_cen, off);
  }
  if ((threadIdx.x & 31) == 0) {
    s_e[threadIdx.x >> 5] = e_tot;
    s_c[threadIdx.x >> 5] = e_cen;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < kWeightsThreads / 32; ++i) {
      a += s_e[i];
      b += s_c[i];
    }
    s_flag = b * 16.0 < a ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag)
    for (int col = threadIdx.x; col < D; col += blockDim.x) centre[col] = 0.0f;
}

// One pass over the descriptors for everything the Gram kernel needs per row: the operand plane(s), the squared norm
// and the projection p = h . w. One warp per PADDED row (frame f, k < 32); rows k >= P only get their statistics.
//
// CENTRING. The planes hold h - c (c = `centre`: the float32-rounded dataset mean, or zero - weights_centre_kernel
// decides), not h: distances are translation invariant, ||h2_j - h1_k||^2 = n'_k + n'_j - 2 G' with
// G' = (H - c)(H - c)^T, and for close-together rows the centred Gram does not cancel. Without it a dataset whose
// descriptors nearly coincide (a trained-like, well-scaled encoder gives row norms^2 ~ 7e2 but distances^2 ~ 1e-4) is
// unresolvable in ANY fp32 accumulator: n_k + n_j - 2G loses 1e-5 absolute, far above the 2e-7 gaps between
// candidates. h - c is ONE float32 subtraction (exact when h and c are within a factor two, Sterbenz - the case that
// matters - and good to 6e-8 relative otherwise). The planes are stored P rows per frame: both sides of the Gram
// kernel read them (the M side through a 3-D tensor map whose 32-row boxes zero-fill rows P..31).
__global__ void __launch_bounds__(256, 4)
prep_rows_kernel(const float* __restrict__ H, int N, int P, int D, const double* __restrict__ w,
                 const float* __restrict__ centre, __half* __restrict__ b_hi, __half* __restrict__ b_lo, int ld,
                 float* __restrict__ sqn, double* __restrict__ pw, unsigned int* __restrict__ nmax_bits,
                 unsigned long long* __restrict__ rowhash) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N * kFrameRows) return;
  const int f = warp / kFrameRows, k = warp % kFrameRows;
  if (k >= P) {  // pad row: never wins an argmin, contributes no score
    if (lane == 0) {
      sqn[warp] = INFINITY;
      pw[warp] = 0.0;
      if (rowhash) rowhash[warp] = 0ull;
    }
    return;
  }
  uint32_t h0 = 0x811c9dc5u, h1 = 0x9747b28cu;   // 64-bit content hash of the row's float32 bits (rep_from_hash_kernel)
  const int64_t r = static_cast<int64_t>(f) * P + k;          // source row = plane row
  const float* h = H + r * D;
  const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0;
  const bool wvec = (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(centre) & 15) == 0;
  double n2 = 0.0, pr = 0.0;
#pragma unroll 4
  for (int c0 = lane * 8; c0 < ld; c0 += 256) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.0f;
    if (vec && c0 + 8 <= D) {
      const float4 u = *reinterpret_cast<const float4*>(h + c0);
      const float4 v = *reinterpret_cast<const float4*>(h + c0 + 4);
      x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w;
      x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < D) x[j] = h[c0 + j];
    }
    __align__(16) __half hh[8];
    __align__(16) __half ll[8];
    double wv[8];
    float mv[8];
    if (wvec && c0 + 8 <= D) {   // 16-byte loads of the weights and the centring vector
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 t = *reinterpret_cast<const double2*>(w + c0 + 2 * j);
        wv[2 * j] = t.x;
        wv[2 * j + 1] = t.y;
      }
      const float4 m0 = *reinterpret_cast<const float4*>(centre + c0);
      const float4 m1 = *reinterpret_cast<const float4*>(centre + c0 + 4);
      mv[0] = m0.x; mv[1] = m0.y; mv[2] = m0.z; mv[3] = m0.w;
      mv[4] = m1.x; mv[5] = m1.y; mv[6] = m1.z; mv[7] = m1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wv[j] = c0 + j < D ? w[c0 + j] : 0.0;
        mv[j] = c0 + j < D ? centre[c0 + j] : 0.0f;
      }
    }
    float s8 = 0.0f;   // squared norm of the centred values: eight float32 terms at a time, summed in float64
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float xc = 0.0f;
      if (c0 + j < D) {
        const uint32_t bits = __float_as_uint(x[j]);
        h0 = (h0 ^ bits) * 0x01000193u;
        h1 = (h1 + bits) * 0x9E3779B1u + (h1 >> 15);
        xc = x[j] - mv[j];
        s8 = fmaf(xc, xc, s8);
        pr = fma(static_cast<double>(x[j]), wv[j], pr);
      }
      split_f32(xc, hh[j], ll[j]);
    }
    n2 += static_cast<double>(s8);
    const int64_t ob = r * ld + c0;
    *reinterpret_cast<uint4*>(b_hi + ob) = *reinterpret_cast<const uint4*>(hh);
    if (b_lo) *reinterpret_cast<uint4*>(b_lo + ob) = *reinterpret_cast<const uint4*>(ll);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    pr += __shfl_xor_sync(0xffffffffu, pr, off);
  }
  if (rowhash) {   // lanes combined position-dependently: identical rows -> identical hashes, always
    h0 *= 2u * lane + 1u;
    h1 ^= h1 >> 13;
    h1 *= 2u * lane + 0x85ebca6bu;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      h0 += __shfl_xor_sync(0xffffffffu, h0, off);
      h1 ^= __shfl_xor_sync(0xffffffffu, h1, off);
    }
    if (lane == 0) rowhash[warp] = (static_cast<unsigned long long>(h0) << 32) | h1;
  }
  if (lane == 0) {
    sqn[warp] = static_cast<float>(n2);
    pw[warp] = pr;
    // largest squared norm (the probe's margin scales an allowance with it); non-negative floats order like their
    // bit patterns, and the plain read first keeps all but a few rows off the atomic
    if (nmax_bits) {
      const unsigned int bits = __float_as_uint(static_cast<float>(n2));
      if (bits > *reinterpret_cast<volatile unsigned int*>(nmax_bits)) atomicMax(nmax_bits, bits);
    }
  }
}

// Residual planes only (lo = fp16(x - fp16(x)) of the centred values), for the three-product kernel when the
// precision probe decided against the one-product kernel: launched after the probe and skipped (device-side gate)
// otherwise, so the common one-product path never writes the residual planes. One warp per plane row.
struct GramControl;
__device__ __forceinline__ bool lo_planes_wanted(const GramControl* ctl);
__global__ void __launch_bounds__(256)
lo_planes_kernel(const float* __restrict__ H, int64_t rows, int D, const float* __restrict__ centre,
                 __half* __restrict__ b_lo, int ld, const GramControl* __restrict__ ctl) {
  if (!lo_planes_wanted(ctl)) return;
  const int64_t r = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* h = H + r * D;
  for (int c0 = lane * 8; c0 < ld; c0 += 256) {
    __align__(16) __half ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = c0 + j < D ? h[c0 + j] - centre[c0 + j] : 0.0f;
      __half hh;
      split_f32(x, hh, ll[j]);
    }
    *reinterpret_cast<uint4*>(b_lo + r * ld + c0) = *reinterpret_cast<const uint4*>(ll);
  }
}

// rep_mask[f]: bit c set iff patch row c of frame f is not bit-identical to an earlier row of the same frame.
// Duplicate patches are common (keypoints near a corner are all shifted onto the same corner patch); their squared
// distances to any row are bit-identical in every arithmetic, np.argmin keeps the first of them, and so does the
// strict '<' scan of the kernels - they never need the exact re-evaluation.
// The mask comes from the row hashes prep_rows_kernel computed while it had the rows in registers (no second pass over
// the descriptors): one warp per frame, lane c = row c. A hash match is only a FILTER - the two rows are then compared
// in full by the whole warp, so the mask is exact (identical rows always hash alike; a 64-bit
// collision between different rows is caught by the comparison).
__global__ void __launch_bounds__(256)
rep_from_hash_kernel(const float* __restrict__ H, int N, int P, int D, const unsigned long long* __restrict__ rowhash,
                     uint32_t* __restrict__ rep_mask) {
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= N) return;
  const unsigned long long mine = rowhash[static_cast<int64_t>(f) * kFrameRows + lane];
  int cand = -1;   // first earlier row with my hash
  for (int e = 0; e < P; ++e) {
    const unsigned long long he = __shfl_sync(0xffffffffu, mine, e);
    if (cand < 0 && e < lane && lane < P && he == mine) cand = e;
  }
  uint32_t pending = __ballot_sync(0xffffffffu, cand >= 0);
  uint32_t dup_mask = 0;
  while (pending) {
    const int c = __ffs(pending) - 1;
    pending &= pending - 1;
    int e = __shfl_sync(0xffffffffu, cand, c);
    bool dup = false;
    while (e >= 0 && !dup) {   // verify against the candidate; on a (never observed) collision try later rows
      const uint32_t* a = reinterpret_cast<const uint32_t*>(H + (static_cast<int64_t>(f) * P + c) * D);
      const uint32_t* b = reinterpret_cast<const uint32_t*>(H + (static_cast<int64_t>(f) * P + e) * D);
      uint32_t d = 0;
      if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0) {   // 16-byte loads, no early exit: they pipeline
        const uint4* a4 = reinterpret_cast<const uint4*>(a);
        const uint4* b4 = reinterpret_cast<const uint4*>(b);
#pragma unroll 10
        for (int i = lane; i < (D >> 2); i += 32) {
          const uint4 x = __ldg(a4 + i), y = __ldg(b4 + i);
          d |= (x.x ^ y.x) | (x.y ^ y.y) | (x.z ^ y.z) | (x.w ^ y.w);
        }
      } else {
#pragma unroll 8
        for (int i = lane; i < D; i += 32) d |= __ldg(a + i) ^ __ldg(b + i);
      }
      dup = !__any_sync(0xffffffffu, d != 0);
      if (!dup) {
        const unsigned long long hc = __shfl_sync(0xffffffffu, mine, c);
        int nxt = -1;
        for (int e2 = e + 1; e2 < c; ++e2)
          if (nxt < 0 && __shfl_sync(0xffffffffu, mine, e2) == hc) nxt = e2;
        e = nxt;
      }
    }
    if (dup) dup_mask |= 1u << c;
  }
  if (lane == 0) rep_mask[f] = ((P >= 32 ? 0xffffffffu : ((1u << P) - 1u)) & ~dup_mask);
}

// ---------------- Gram + argmin + score epilogue ----------------
extern std::atomic<int> g_promote_k;  // planes.cu

// Device-side precision decision written by the probe (see below); both Gram kernels are always launched and the
// one that is not selected returns immediately, so the choice needs no host round trip.
struct GramControl {
  int use_refine;       // 1: one-product kernel + exact refinement, 0: three-product kernel
  float margin;         // candidates within `margin` of the approximate minimum are re-evaluated exactly
  float sigma;          // estimated std of the approximate squared-distance error (diagnostic)
  float flagged_frac;   // estimated fraction of rows needing refinement (diagnostic)
  unsigned long long flagged_rows;   // counted by the refine kernels (diagnostic)
  unsigned long long refined_cands;  // "
  unsigned int n_entries;            // frame pairs the Gram kernel deferred to gram_refine_fix_kernel (may exceed the capacity)
  unsigned int pad_;
};

__device__ __forceinline__ bool lo_planes_wanted(const GramControl* ctl) { return ctl->use_refine == 0; }

// A frame pair with at least one ambiguous row, handed from the Gram epilogue to gram_refine_fix_kernel: the
// approximate match of every row and, for the ambiguous rows, the candidates inside the margin.
struct RefineEntry {
  int fa, fb;
  uint32_t flagged;   // ballot of the lanes (rows of fa) that need the exact re-evaluation
  uint32_t pad_;
  uint8_t bj[32];     // approximate argmin per row
  uint32_t mk[32];    // candidate mask per row (0 = row is not ambiguous)
};

struct GramParams {
  int n_tile, k_blocks, ab_fmt, kc;
  const int2* tiles;  // (mt, nt) work list in L2-friendly order
  int num_tiles;
  const int2* pair_tiles;  // CTA-pair kernel: (first of two adjacent M tiles, nt)
  int num_pair_tiles;
  int N, P, D;
  const float* desc;  // [N, P, D] float32 descriptors (exact refinement reads them)
  const float* sqn;   // [N*32]
  const double* pw;   // [N*32]
  float a, b;
  int full;           // 1: every ordered pair i != j; 0: i < j, mirrored
  float* S;           // [N, N]
  GramControl* ctl;   // NULL: always enabled
  int want_refine;    // this launch runs only when ctl->use_refine == want_refine
  const uint32_t* rep_mask;  // [N] bit c set = row c of the frame is the FIRST of its class of bit-identical rows
  int col_stride;     // accumulator columns per frame on the N side: P when the N tile is read as 8 P plane rows (even
                      // P < 32; no MMA work on pad rows), else 32
  int b_frame_map;    // 1: the N side is read through the 3-D frame map (32-row boxes), 0: plain rows (col_stride = P)
  RefineEntry* work;  // deferred-refinement work list (NULL: refine inside the epilogue)
  int work_cap;       // entries the list holds; pairs beyond it are refined inside the epilogue
};

// score of one frame pair from the per-row matches: sum_k (a + b ln |p_ik - p_j,bj(k)|), warp-wide
__device__ __forceinline__ void pair_score(const GramParams& p, int fa, int fb, int lane, double pa, int bj) {
  const double my_pb = p.pw[fb * kFrameRows + lane];
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

// exact squared distance between two float32 rows, accumulated in float64, whole warp cooperates
__device__ __forceinline__ double exact_d2(const float* __restrict__ a, const float* __restrict__ b, int D, int lane) {
  double acc0 = 0.0, acc1 = 0.0;
  if ((D & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const int n4 = D >> 2;
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
      const double d0 = static_cast<double>(x.x - y.x), d1 = static_cast<double>(x.y - y.y);
      const double d2 = static_cast<double>(x.z - y.z), d3 = static_cast<double>(x.w - y.w);
      acc0 = fma(d0, d0, acc0);
      acc1 = fma(d1, d1, acc1);
      acc0 = fma(d2, d2, acc0);
      acc1 = fma(d3, d3, acc1);
    }
  } else {
    for (int i = lane; i < D; i += 32) {
      const double d = static_cast<double>(a[i] - b[i]);
      acc0 = fma(d, d, acc0);
    }
  }
  acc0 += acc1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
  return acc0;
}

// exact squared distances of row a to TWO rows at once (a is read once): the usual refinement case
__device__ __forceinline__ void exact_d2_pair(const float* __restrict__ a, const float* __restrict__ b0,
                                              const float* __restrict__ b1, int D, int lane, double& r0, double& r1) {
  double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
  if ((D & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* p4 = reinterpret_cast<const float4*>(b0);
    const float4* q4 = reinterpret_cast<const float4*>(b1);
    const int n4 = D >> 2;
#pragma unroll 4
    for (int i = lane; i < n4; i += 32) {
      const float4 x = __ldg(a4 + i), y = __ldg(p4 + i), z = __ldg(q4 + i);
      double d;
      d = static_cast<double>(x.x - y.x); s0 = fma(d, d, s0);
      d = static_cast<double>(x.y - y.y); s1 = fma(d, d, s1);
      d = static_cast<double>(x.z - y.z); s0 = fma(d, d, s0);
      d = static_cast<double>(x.w - y.w); s1 = fma(d, d, s1);
      d = static_cast<double>(x.x - z.x); t0 = fma(d, d, t0);
      d = static_cast<double>(x.y - z.y); t1 = fma(d, d, t1);
      d = static_cast<double>(x.z - z.z); t0 = fma(d, d, t0);
      d = static_cast<double>(x.w - z.w); t1 = fma(d, d, t1);
    }
  } else {
    for (int i = lane; i < D; i += 32) {
      const double d = static_cast<double>(a[i] - b0[i]), e = static_cast<double>(a[i] - b1[i]);
      s0 = fma(d, d, s0);
      t0 = fma(e, e, t0);
    }
  }
  s0 += s1;
  t0 += t1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, off);
    t0 += __shfl_xor_sync(0xffffffffu, t0, off);
  }
  r0 = s0;
  r1 = t0;
}

// ---- (A) three fp16 products + two-level accumulation: every Gram entry good to ~1e-7 relative
template <int BK, int NPROD>
struct GramPolicy {
  using Cfg = GemmCfg<BK, NPROD>;
  using Params = GramParams;
  static constexpr bool kFrameMaps = true;  // planes stored P rows per frame, M side read in 32-row boxes
  static constexpr bool kPromote = NPROD == 3;
  static constexpr int kEpiWarps = 4;
  static constexpr uint64_t kHintA = kEvictNormal;
  static constexpr uint64_t kHintB = kEvictNormal;

  static __device__ __forceinline__ bool enabled(const Params& p) {
    return p.ctl == nullptr || p.ctl->use_refine == p.want_refine;
  }
  static __device__ __forceinline__ int chunk_stride(const Params& p) { return p.col_stride; }
  static __device__ __forceinline__ int num_tiles_pair(const Params& p, int cluster, int nclusters) {
    return cluster < p.num_pair_tiles ? (p.num_pair_tiles - cluster + nclusters - 1) / nclusters : 0;
  }
  static __device__ __forceinline__ TileCoord tile_pair(const Params& p, int cluster, int nclusters, int i) {
    const int2 t = __ldg(p.pair_tiles + cluster + i * nclusters);
    TileCoord tc;
    tc.mt = t.x;
    tc.nt = t.y;
    return tc;
  }
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
    __device__ Epilogue(const Params& p_, int quarter_, int, int lane_, void*) : p(p_), quarter(quarter_), lane(lane_) {}

    int fa;
    bool fa_ok;
    double pa;
    __device__ __forceinline__ void begin_tile(TileCoord tc) {
      fa = tc.mt * kFramesPerMTile + quarter;  // this warp's frame i (lane = patch k)
      fa_ok = fa < p.N;
      pa = fa_ok ? p.pw[fa * kFrameRows + lane] : 0.0;
    }
    __device__ __forceinline__ void end_tile(TileCoord) {}
    __device__ __forceinline__ void post_tile(TileCoord) {}

    // chunk c = the 32 accumulator columns of frame j = 8 nt + c
    template <int SLOT>
    __device__ __forceinline__ void chunk(TileCoord tc, int c, float (&v)[32]) {
      const int fb = tc.nt * kFramesPerNTile + c;
      if (!fa_ok || fb >= p.N) return;  // warp-uniform
      if (fa == fb) {
        if (lane == 0) p.S[static_cast<int64_t>(fa) * p.N + fa] = -1.0f;  // reference fill value (:31)
        return;
      }
      if (!p.full && fa > fb) return;
      const float my_nb = p.sqn[fb * kFrameRows + lane];
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
      pair_score(p, fa, fb, lane, pa, bj);
    }
    __device__ __forceinline__ void finish() {}
  };
};

// Exact re-evaluation of the ambiguous rows of frame pair (fa, fb): `fl` = ballot of the lanes (rows of fa) whose
// candidate mask `mk` holds more than one class of rows of fb inside the margin. Returns this lane's match.
__device__ __forceinline__ int refine_rows(const GramParams& p, int fa, int fb, int lane, uint32_t fl, uint32_t my_mk,
                                           int my_bj) {
  if (lane == 0 && p.ctl) atomicAdd(&p.ctl->flagged_rows, static_cast<unsigned long long>(__popc(fl)));
  const float* abase = p.desc + static_cast<int64_t>(fa) * p.P * p.D;
  const float* bbase = p.desc + static_cast<int64_t>(fb) * p.P * p.D;
  while (fl) {
    const int L = __ffs(fl) - 1;
    fl &= fl - 1;
    uint32_t m = __shfl_sync(0xffffffffu, my_mk, L);
    if (lane == 0 && p.ctl) atomicAdd(&p.ctl->refined_cands, static_cast<unsigned long long>(__popc(m)));
    const float* arow = abase + static_cast<int64_t>(L) * p.D;
    double best = INFINITY;
    int bx = 0;
    while (m) {  // ascending j + strict '<'  ->  first exact minimum, like np.argmin
      const int j0 = __ffs(m) - 1;
      m &= m - 1;
      if (m) {     // two candidates per pass over the row
        const int j1 = __ffs(m) - 1;
        m &= m - 1;
        double e0, e1;
        exact_d2_pair(arow, bbase + static_cast<int64_t>(j0) * p.D, bbase + static_cast<int64_t>(j1) * p.D, p.D, lane,
                      e0, e1);
        if (e0 < best) {
          best = e0;
          bx = j0;
        }
        if (e1 < best) {
          best = e1;
          bx = j1;
        }
      } else {
        const double dd = exact_d2(arow, bbase + static_cast<int64_t>(j0) * p.D, p.D, lane);
        if (dd < best) {
          best = dd;
          bx = j0;
        }
      }
    }
    if (lane == L) my_bj = bx;
  }
  return my_bj;
}

// ---- (B) one fp16 product + exact refinement. The single-product Gram entry carries the fp16 rounding of the
// operands (std `sigma` on the squared distance, estimated by the probe). A row whose runner-up lies within
// `margin` (>= 8 sigma) of the approximate minimum is re-evaluated: the candidates inside the margin get their exact
// float64-accumulated squared distance straight from the float32 descriptors and the first exact minimum wins.
// Rows outside the margin already have the reference's argmin. 3x fewer tensor-core FLOPs than (A).
struct GramRefinePolicy {
  using Cfg = GemmCfg<64, 1>;
  using Params = GramParams;
  static constexpr bool kFrameMaps = true;
  static constexpr bool kPromote = false;
  static constexpr int kEpiWarps = 8;
  static constexpr uint64_t kHintA = kEvictNormal;
  static constexpr uint64_t kHintB = kEvictNormal;

  static __device__ __forceinline__ bool enabled(const Params& p) {
    return p.ctl == nullptr || p.ctl->use_refine == p.want_refine;
  }
  static __device__ __forceinline__ int chunk_stride(const Params& p) { return p.col_stride; }
  static __device__ __forceinline__ int num_tiles_pair(const Params& p, int cluster, int nclusters) {
    return cluster < p.num_pair_tiles ? (p.num_pair_tiles - cluster + nclusters - 1) / nclusters : 0;
  }
  static __device__ __forceinline__ TileCoord tile_pair(const Params& p, int cluster, int nclusters, int i) {
    const int2 t = __ldg(p.pair_tiles + cluster + i * nclusters);
    TileCoord tc;
    tc.mt = t.x;
    tc.nt = t.y;
    return tc;
  }
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
    const int quarter, half, lane;
    float margin;
    __device__ Epilogue(const Params& p_, int quarter_, int half_, int lane_, void*)
        : p(p_), quarter(quarter_), half(half_), lane(lane_) {
      margin = p.ctl ? p.ctl->margin : 0.0f;
    }

    int fa;
    bool fa_ok;
    double pa;
    int bj[4];          // approximate argmin of my row for each of this warp's 4 chunks
    uint32_t mk[4];     // candidate mask when the row needs refinement, else 0
    __device__ __forceinline__ void begin_tile(TileCoord tc) {
      fa = tc.mt * kFramesPerMTile + quarter;
      fa_ok = fa < p.N;
      pa = fa_ok ? p.pw[fa * kFrameRows + lane] : 0.0;
    }

    __device__ __forceinline__ bool chunk_active(int fb) const {
      return fa_ok && fb < p.N && fa != fb && (p.full || fa < fb);
    }

    template <int SLOT>
    __device__ __forceinline__ void chunk(TileCoord tc, int c, float (&v)[32]) {
      const int fb = tc.nt * kFramesPerNTile + c;
      bj[SLOT] = 0;
      mk[SLOT] = 0;
      if (!chunk_active(fb)) return;  // warp-uniform
      const float my_nb = p.sqn[fb * kFrameRows + lane];
      float d[32];
      float best = INFINITY;
      int b = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float nb = __shfl_sync(0xffffffffu, my_nb, j);
        d[j] = fmaf(-2.0f, v[j], nb);
        if (d[j] < best) {
          best = d[j];
          b = j;
        }
      }
      uint32_t m = 0;
      const float lim = best + margin;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (d[j] <= lim) m |= 1u << j;
      bj[SLOT] = b;
      // candidates inside the margin, one per class of bit-identical rows (the first minimum b is always the first
      // of its class): more than one left = the row needs the exact re-evaluation
      m &= __ldg(p.rep_mask + fb);
      mk[SLOT] = (lane < p.P && (m & (m - 1)) != 0) ? m : 0u;
    }
    __device__ __forceinline__ void end_tile(TileCoord) {}

    template <int SLOT>
    __device__ __forceinline__ void finish_chunk(TileCoord tc) {
      const int fb = tc.nt * kFramesPerNTile + half * 4 + SLOT;
      if (fa_ok && fb < p.N && fa == fb) {
        if (lane == 0) p.S[static_cast<int64_t>(fa) * p.N + fa] = -1.0f;
        return;
      }
      if (!chunk_active(fb)) return;
      const uint32_t fl = __ballot_sync(0xffffffffu, mk[SLOT] != 0);
      if (fl) {
        // Ambiguous rows: hand the pair to gram_refine_fix_kernel (one short burst of stores) - re-evaluating here
        // stalls this warp for microseconds per row on 10 KB row reads, the accumulator double buffer runs dry and
        // the tensor pipe idles (measured: 1.4 ms of Gram + 1.4 ms of stalls on the bench workload).
        if (p.work) {
          unsigned int slot = 0;
          if (lane == 0) slot = atomicAdd(&p.ctl->n_entries, 1u);
          slot = __shfl_sync(0xffffffffu, slot, 0);
          if (slot < static_cast<unsigned int>(p.work_cap)) {
            RefineEntry* e = p.work + slot;
            if (lane == 0) {
              e->fa = fa;
              e->fb = fb;
              e->flagged = fl;
            }
            e->bj[lane] = static_cast<uint8_t>(bj[SLOT]);
            e->mk[lane] = mk[SLOT];
            return;
          }
        }
        bj[SLOT] = refine_rows(p, fa, fb, lane, fl, mk[SLOT], bj[SLOT]);  // list full (or absent): refine in place
      }
      pair_score(p, fa, fb, lane, pa, bj[SLOT]);
    }

    // runs after the accumulator went back to the MMA warp: refinement + scores never stall the tensor pipe
    __device__ __forceinline__ void post_tile(TileCoord tc) {
      finish_chunk<0>(tc);
      finish_chunk<1>(tc);
      finish_chunk<2>(tc);
      finish_chunk<3>(tc);
    }
    __device__ __forceinline__ void finish() {}
  };
};

// Second pass of mode (B): one warp per deferred frame pair re-evaluates its ambiguous rows exactly and writes the
// pair's score. Thousands of independent warps keep the row reads in flight (the list is in tile order, so
// neighbouring entries share frames and the reads mostly hit L2); same arithmetic as the in-epilogue path.
constexpr int kFixThreads = 256;
__global__ void __launch_bounds__(kFixThreads, 4) gram_refine_fix_kernel(const GramParams p) {
  if (p.ctl->use_refine != p.want_refine) return;
  const unsigned int n = min(p.ctl->n_entries, static_cast<unsigned int>(p.work_cap));
  const int lane = threadIdx.x & 31;
  const unsigned int warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const RefineEntry* e = p.work + i;
    const int fa = e->fa, fb = e->fb;
    const uint32_t fl = e->flagged;
    const double pa = p.pw[fa * kFrameRows + lane];
    const int bj = refine_rows(p, fa, fb, lane, fl, e->mk[lane], e->bj[lane]);
    pair_score(p, fa, fb, lane, pa, bj);
  }
}

// ---- probe: estimate the single-product error and the share of rows it would leave ambiguous
constexpr int kProbeSamples = 256;   // x (P - 1) ~ 7 k error samples: the margin rests on their rms (8 sigma), not on a tail
struct ProbeAccum {
  double sum_err2;        // sum over sampled (row, candidate) of (approximate - exact squared distance)^2
  unsigned long long n_err;
  unsigned long long max_row_bits;
  unsigned int nmax_bits;   // largest squared row norm (float bits), written by prep_rows_kernel
  unsigned int pad_;
};
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// one CTA per sample: a random row k of frame fa against all P rows of a random other frame fb (warp j = row j)
__global__ void __launch_bounds__(1024)
gram_probe_kernel(const float* __restrict__ desc, int N, int P, int D, const float* __restrict__ centre,
                  const uint32_t* __restrict__ rep_mask, ProbeAccum* acc, float* gaps) {
  __shared__ double s_e[32];
  __shared__ double s_d[32];
  const int sample = blockIdx.x;
  const int j = threadIdx.x >> 5;  // candidate row of frame fb handled by this warp
  const int lane = threadIdx.x & 31;
  const uint32_t h = hash_u32(0x9e3779b9u * (sample + 1));
  const int fa = h % N;
  int fb = hash_u32(h) % N;
  if (fb == fa) fb = (fb + 1) % N;
  const int k = hash_u32(h ^ 0x5bd1e995u) % P;
  const uint32_t rep = rep_mask[fb];
  // a bit-identical copy of an earlier row never needs refinement: only class representatives take part
  const bool active = j < P && ((rep >> j) & 1u);
  if (active) {
    const float* a = desc + (static_cast<int64_t>(fa) * P + k) * D;
    const float* b = desc + (static_cast<int64_t>(fb) * P + j) * D;
    // err  = sum_c x y - fp16(x) fp16(y) = sum_c (x - xh) y + xh (y - yh): both residuals are exact in float32 and the
    //        sum is a statistic (three digits are plenty) -> float32 throughout, no float64 conversions;
    // dist = sum_c y (y - 2 x)  (squared distance up to the row constant): float32 partial sums of four elements,
    //        added in float64 (absolute error ~1e-5 on values ~1e3; it is compared with a margin ~1e-2).
    float err = 0.0f;
    double dist = 0.0;
    // (x, y are the CENTRED values h - c the planes hold, see prep_rows_kernel)
    auto term = [&](float xr, float yr, float m) -> float {
      const float x = xr - m, y = yr - m;
      const float xh = __half2float(__float2half_rn(x)), yh = __half2float(__float2half_rn(y));
      err = fmaf(x - xh, y, fmaf(xh, y - yh, err));
      return y * (y - 2.0f * x);
    };
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(desc) & 15) == 0 && (reinterpret_cast<uintptr_t>(centre) & 15) == 0) {
      const float4* a4 = reinterpret_cast<const float4*>(a);
      const float4* b4 = reinterpret_cast<const float4*>(b);
      const float4* m4 = reinterpret_cast<const float4*>(centre);
#pragma unroll 4
      for (int c = lane; c < (D >> 2); c += 32) {
        const float4 x = __ldg(a4 + c), y = __ldg(b4 + c), m = __ldg(m4 + c);
        dist += static_cast<double>((term(x.x, y.x, m.x) + term(x.y, y.y, m.y)) +
                                    (term(x.z, y.z, m.z) + term(x.w, y.w, m.w)));
      }
    } else {
      for (int c = lane; c < D; c += 32) dist += static_cast<double>(term(a[c], b[c], centre[c]));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      err += __shfl_xor_sync(0xffffffffu, err, off);
      dist += __shfl_xor_sync(0xffffffffu, dist, off);
    }
    if (lane == 0) {
      s_e[j] = 2.0 * static_cast<double>(err);  // error of the approximate squared distance n_j - 2 G
      s_d[j] = dist;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // The single-product error of a squared distance has a large COMMON part (saturated values just below 1 round
    // up to 1.0 in fp16: every Gram entry of the row is biased the same way) that cancels in the comparison of two
    // candidates of the same row. What decides an argmin is the error of a DIFFERENCE d_a - d_b, i.e. the spread of
    // the errors around their row mean: accumulate sum_j (e_j - mean_j e)^2.
    double sum_e = 0.0, sum_e2 = 0.0, d1 = INFINITY, d2 = INFINITY;
    int n_rep = 0;
    for (int c = 0; c < P; ++c) {
      if (!((rep >> c) & 1u)) continue;
      ++n_rep;
      sum_e += s_e[c];
      sum_e2 += s_e[c] * s_e[c];
      const double dist = s_d[c];
      if (dist < d1) {
        d2 = d1;
        d1 = dist;
      } else if (dist < d2) {
        d2 = dist;
      }
    }
    const double centered = n_rep > 1 ? fmax(sum_e2 - sum_e * sum_e / n_rep, 0.0) : 0.0;  // sum (e - mean)^2
    atomicAdd(&acc->sum_err2, centered);
    atomicAdd(&acc->n_err, static_cast<unsigned long long>(n_rep > 1 ? n_rep - 1 : 0));
    const double var_row = n_rep > 1 ? centered / (n_rep - 1) : 0.0;
    atomicMax(&acc->max_row_bits, static_cast<unsigned long long>(__double_as_longlong(var_row)));  // >= 0: bit order = value order
    gaps[sample] = static_cast<float>(d2 - d1);
  }
}
__global__ void gram_probe_finalize_kernel(const ProbeAccum* acc, const float* gaps, float max_flag_frac, int force,
                                           GramControl* ctl) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const double rms = sqrt(acc->sum_err2 / static_cast<double>(acc->n_err > 0 ? acc->n_err : 1));
  const double rms_max = sqrt(__longlong_as_double(static_cast<long long>(acc->max_row_bits)));
  const float s_nmax = __uint_as_float(acc->nmax_bits);
  // The gap between two candidates carries the difference of two such errors (std sqrt(2) sigma): the margin is
  // >= 8 standard deviations of that difference (global estimate; >= 4 for the worst sampled row) plus an allowance
  // for the tensor core's truncating fp32 accumulation (differential part ~1e-6 of the largest Gram entry)
  const float margin = static_cast<float>(1.41421356 * fmax(8.0 * rms, 4.0 * rms_max)) + 4e-6f * s_nmax;
  int cnt = 0;
  for (int i = threadIdx.x; i < kProbeSamples; i += blockDim.x) cnt += gaps[i] < margin ? 1 : 0;
  atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    const float frac = static_cast<float>(s_cnt) / kProbeSamples;
    ctl->margin = margin;
    ctl->sigma = static_cast<float>(rms);
    ctl->flagged_frac = frac;
    ctl->use_refine = force >= 0 ? force : (frac <= max_flag_frac ? 1 : 0);
    ctl->flagged_rows = 0;
    ctl->refined_cands = 0;
    ctl->n_entries = 0;
  }
}

// Work list: super-blocks of g_sim_mgroup M tiles; inside a super-block N tile outermost so the ~148 concurrently
// running tiles touch g_sim_mgroup A row-blocks and ~148/g_sim_mgroup B row-blocks (fits L2) instead of streaming all of H.
static void build_tile_list(int N, int full, int part, int n_parts, std::vector<int2>& out,
                            std::vector<int2>& out_pairs) {
  const int m_tiles = ceil_div(N, kFramesPerMTile), n_tiles = ceil_div(N, kFramesPerNTile);
  const int m_pairs = (m_tiles + 1) / 2;
  out.clear();
  out_pairs.clear();
  // Ownership is dealt in PAIRS of adjacent M tiles (what one CTA pair computes): every n_parts-th pair, interleaved
  // so the triangle's long and short rows are spread evenly; super-blocks of g_sim_mgroup M tiles for L2 locality.
  std::vector<int> owned;
  for (int q = part; q < m_pairs; q += n_parts) owned.push_back(q);
  const size_t group = std::max(1, g_sim_mgroup.load() / 2);
  for (size_t g0 = 0; g0 < owned.size(); g0 += group) {
    const size_t g1 = std::min(g0 + group, owned.size());
    for (int nt = 0; nt < n_tiles; ++nt)
      for (size_t g = g0; g < g1; ++g) {
        const int q = owned[g];
        const int fb_max = std::min(nt * kFramesPerNTile + kFramesPerNTile - 1, N - 1);
        bool any = false;
        for (int mt = 2 * q; mt < std::min(2 * q + 2, m_tiles); ++mt) {
          // upper-triangle mode needs a pair fa < fb, or the diagonal block (to write the -1 fill)
          if (full || fb_max >= mt * kFramesPerMTile) {
            out.push_back(make_int2(mt, nt));
            any = true;
          }
        }
        if (any) out_pairs.push_back(make_int2(2 * q, nt));
      }
  }
}

// The work lists depend only on (N, full, part, n_parts, super-block size): they are built once per key and kept in
// LIBRARY-OWNED device memory (a caller workspace could be overwritten between calls), so a call neither rebuilds the
// list on the host nor copies it. A handful of keys per device is all a process uses (one per sequence length).
struct TileListKey {
  int dev, N, full, part, n_parts, mgroup;
  bool operator==(const TileListKey& o) const {
    return dev == o.dev && N == o.N && full == o.full && part == o.part && n_parts == o.n_parts && mgroup == o.mgroup;
  }
};
struct TileListEntry {
  TileListKey key;
  int2* tiles = nullptr;
  int2* pair_tiles = nullptr;
  int num_tiles = 0, num_pair_tiles = 0;
  uint64_t stamp = 0;
};
constexpr int kTileCacheSlots = 16;
static std::mutex g_tile_mutex;
static TileListEntry g_tile_cache[kTileCacheSlots];
static uint64_t g_tile_stamp = 0;

// Returns the cached device lists for the key (building + uploading them on a miss: synchronous, first call only).
static int get_tile_lists(int N, int full, int part, int n_parts, TileListEntry* out) {
  TileListKey key{current_device(), N, full ? 1 : 0, part, n_parts, g_sim_mgroup.load()};
  std::lock_guard<std::mutex> lock(g_tile_mutex);
  int victim = 0;
  for (int i = 0; i < kTileCacheSlots; ++i) {
    if (g_tile_cache[i].tiles && g_tile_cache[i].key == key) {
      g_tile_cache[i].stamp = ++g_tile_stamp;
      *out = g_tile_cache[i];
      return DLC_OK;
    }
    if (g_tile_cache[i].stamp < g_tile_cache[victim].stamp) victim = i;
  }
  std::vector<int2> tiles, pair_tiles;
  build_tile_list(N, full, part, n_parts, tiles, pair_tiles);
  TileListEntry e;
  e.key = key;
  e.num_tiles = static_cast<int>(tiles.size());
  e.num_pair_tiles = static_cast<int>(pair_tiles.size());
  int2* mem = nullptr;
  const size_t n_all = tiles.size() + pair_tiles.size();
  DLC_CUDA(cudaMalloc(reinterpret_cast<void**>(&mem), sizeof(int2) * std::max<size_t>(n_all, 1)));
  cudaError_t err = cudaSuccess;
  if (!tiles.empty()) err = cudaMemcpy(mem, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice);
  if (err == cudaSuccess && !pair_tiles.empty())
    err = cudaMemcpy(mem + tiles.size(), pair_tiles.data(), sizeof(int2) * pair_tiles.size(), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    cudaFree(mem);
    return fail(DLC_ECUDA, "dlc_sdav_similarity: tile list upload failed: %s", cudaGetErrorString(err));
  }
  e.tiles = mem;
  e.pair_tiles = mem + tiles.size();
  e.stamp = ++g_tile_stamp;
  // an evicted list may still be read by a kernel in flight on some stream: free it only after the device drained
  if (g_tile_cache[victim].tiles) {
    cudaDeviceSynchronize();
    cudaFree(g_tile_cache[victim].tiles);
  }
  g_tile_cache[victim] = e;
  *out = e;
  return DLC_OK;
}

constexpr int64_t kMaxRefineEntries = 1 << 18;
std::atomic<int> g_refine_cap{-1};  // developer override of the list capacity (dlc_debug_set key 8; 0 = refine in the epilogue)
struct SimWorkspace {
  size_t off_bhi, off_blo, off_part, off_w, off_mean, off_sqn, off_pw, off_hash, off_rep, off_ctl, off_probe, off_gaps, off_work, total;
  int ld, rows_pad;
  int col_stride;  // 32, or P when the N tile is read as 8 P plane rows
  int rows_b;      // rows of the planes (N * P)
  int work_cap;    // entries of the deferred-refinement list
};
static SimWorkspace sim_layout(int N, int P, int D) {
  SimWorkspace w{};
  w.ld = dlc_plane_ld(D);
  w.rows_pad = N * kFrameRows;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  // ONE pair of planes, P rows per frame. The M side always reads them through the 3-D frame map (32-row boxes, rows
  // P..31 zero-filled by TMA). The N side reads 8 P plain rows per tile when that keeps UMMA's N = 8 P a multiple of
  // 16 (even P: no MMA work on the pad rows of every frame column block), else it uses the frame map as well.
  w.col_stride = (P < kFrameRows && (P & 1) == 0) ? P : kFrameRows;
  w.rows_b = N * P;
  const size_t plane_b = static_cast<size_t>(w.rows_b) * w.ld * 2;
  w.off_bhi = take(plane_b);
  w.off_blo = take(plane_b);
  w.off_part = take(sizeof(double) * (kColSumSlabs + 1) * 2 * D);   // per-slab sums + their total
  w.off_w = take(sizeof(double) * D);
  w.off_mean = take(sizeof(float) * D);   // centring vector (float32)
  w.off_sqn = take(sizeof(float) * w.rows_pad);
  w.off_pw = take(sizeof(double) * w.rows_pad);
  w.off_hash = take(sizeof(unsigned long long) * w.rows_pad);
  w.off_rep = take(sizeof(uint32_t) * N);
  w.off_ctl = take(sizeof(GramControl));
  w.off_probe = take(sizeof(ProbeAccum));
  w.off_gaps = take(sizeof(float) * kProbeSamples);
  // deferred refinement: auto mode selects the one-product kernel only below ~1 % ambiguous rows, so a quarter of a
  // million ambiguous PAIRS (46 MB) covers sequences of thousands of frames; beyond it the epilogue refines in place
  w.work_cap = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(N) * N, kMaxRefineEntries));
  w.off_work = take(sizeof(RefineEntry) * static_cast<size_t>(w.work_cap));
  w.total = o;
  return w;
}

// Tensor maps of one Gram launch: A = frame map (4 frames x 32 rows per tile); B = frame map of `b_box_frames` frames,
// or plain rows (col_stride = P) of `b_rows` rows per load.
struct GramPlanes {
  const void* hi;   // [N * P, ld] fp16, centred
  const void* lo;   // residual plane (three-product kernel only; may alias hi when unused)
  int ld, col_stride;
};
static bool gram_maps(const GramPlanes& L, int N, int P, int BK, int n_tile, bool pair, CUtensorMap* ta0,
                      CUtensorMap* ta1, CUtensorMap* tb0, CUtensorMap* tb1) {
  const void* bhi = L.hi;
  const void* blo = L.lo ? L.lo : L.hi;
  if (!make_tmap_frames(ta0, bhi, L.ld, P, N, BK, kFramesPerMTile) ||
      !make_tmap_frames(ta1, blo, L.ld, P, N, BK, kFramesPerMTile))
    return false;
  const int b_rows = pair ? n_tile / 2 : n_tile;
  if (L.col_stride == kFrameRows)
    return make_tmap_frames(tb0, bhi, L.ld, P, N, BK, b_rows / kFrameRows) &&
           make_tmap_frames(tb1, blo, L.ld, P, N, BK, b_rows / kFrameRows);
  const uint64_t rows_b = static_cast<uint64_t>(N) * P;
  return make_tmap_k_major(tb0, bhi, 0, L.ld, rows_b, L.ld, BK, b_rows) &&
         make_tmap_k_major(tb1, blo, 0, L.ld, rows_b, L.ld, BK, b_rows);
}

template <class Policy>
static int run_gram(const GramPlanes& L, GramParams p, cudaStream_t stream) {
  constexpr int BK = Policy::Cfg::BK;
  CUtensorMap ta0, ta1, tb0, tb1;
  if (!gram_maps(L, p.N, p.P, BK, p.n_tile, false, &ta0, &ta1, &tb0, &tb1))
    return fail(DLC_ECUDA, "dlc_sdav_similarity: cuTensorMapEncodeTiled failed");
  p.k_blocks = ceil_div(p.D, BK);  // K blocks that hold data: the planes are zero from D to ld
  p.kc = std::max(1, g_promote_k.load() / BK);
  if (p.num_tiles == 0) return DLC_OK;  // a part that owns no tile (more parts than M tiles)
  const int grid = std::min(p.num_tiles, sm_count());
  cudaError_t e = launch_gemm<Policy>(ta0, ta1, tb0, tb1, p, grid, stream);
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_sdav_similarity: launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

// The same launch on CTA pairs (gemm_pair_sm100.cuh): each CTA stages half of the N tile.
extern std::atomic<int> g_cta_pair;   // planes.cu
extern std::atomic<int> g_gram_pair;  // planes.cu
template <class Policy>
static int run_gram_pair(const GramPlanes& L, GramParams p, cudaStream_t stream) {
  constexpr int BK = Policy::Cfg::BK;
  CUtensorMap ta0, ta1, tb0, tb1;
  if (!gram_maps(L, p.N, p.P, BK, p.n_tile, true, &ta0, &ta1, &tb0, &tb1))
    return fail(DLC_ECUDA, "dlc_sdav_similarity: cuTensorMapEncodeTiled failed");
  p.k_blocks = ceil_div(p.D, BK);
  p.kc = std::max(1, g_promote_k.load() / BK);
  if (p.num_pair_tiles == 0) return DLC_OK;  // a part that owns no tile
  const int clusters = std::min(p.num_pair_tiles, sm_count() / 2);
  cudaError_t e = launch_gemm_pair<Policy>(ta0, ta1, tb0, tb1, p, clusters, stream);
  if (e != cudaSuccess) return fail(DLC_ECUDA, "dlc_sdav_similarity: launch failed: %s", cudaGetErrorString(e));
  return DLC_OK;
}

}  // namespace dlc

using namespace dlc;

std::atomic<int> g_probe_side_stream{1};  // dlc_debug_set key 9: kept for compatibility, no effect (the probe is serial)

// Developer switch (dlc_debug_set key 1): launch only the Gram/score kernel, reusing the operand planes, statistics
// and tile list a previous full call left in the workspace. Lets bench.py time that kernel alone with CUDA events.
static std::atomic<int> g_gram_only{0};
// AUTO uses the one-product + refinement kernel only when the probe expects at most this share of rows to need the
// exact re-evaluation. Measured on B200 (1063 frames, 17 M row-vs-frame decisions): the refinement costs ~11 us per
// 1000 flagged rows (each candidate streams 10 KB float32 rows from L2), i.e. ~0.2 ms per 0.1 % of flagged rows,
// against ~3 ms saved by issuing one tensor product instead of three: break-even near 1.5 %.
static float g_max_flag_frac = 0.012f;
extern "C" int dlc_sdav_debug_gram_only(int on) {  // 2: additionally skip the second (refinement) pass - timing only
  g_gram_only = on;
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
  if (ws_bytes < sizeof(double) * (kColSumSlabs + 1) * 2 * D)
    return fail(DLC_ENOMEM, "dlc_sdav_weights: workspace of %zu bytes needed", sizeof(double) * (kColSumSlabs + 1) * 2 * D);
  cudaStream_t s = as_stream(stream);
  double* part = static_cast<double*>(ws_dev);
  const int64_t rows = static_cast<int64_t>(N) * P;
  colsum_partial_kernel<<<dim3(ceil_div(D, 32), kColSumSlabs), 32 * kColSumRowLanes, 0, s>>>(desc_dev, rows, D, part);
  double* total = part + static_cast<size_t>(kColSumSlabs) * 2 * D;
  colsum_reduce_kernel<<<ceil_div(2 * D, 128), 128, 0, s>>>(part, kColSumSlabs, 2 * D, total);
  weights_centre_kernel<<<1, kWeightsThreads, 0, s>>>(total, 1, rows, D, mu, sigma, w_dev, nullptr);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

static int sdav_similarity_impl(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a,
                                double b, const double* w_dev, int precision, int full_asymmetric, int part,
                                int n_parts, float* S_dev, void* ws_dev, size_t ws_bytes, void* stream);
static bool use_pairs(const GramParams& p) {
  return g_cta_pair && g_gram_pair && p.n_tile >= 32 && p.n_tile % 16 == 0 && (p.n_tile / 2) % 8 == 0 &&
         (p.num_pair_tiles >= sm_count() / 2 || g_cta_pair == 2);
}

extern "C" int dlc_sdav_similarity(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a,
                                   double b, const double* w_dev, int precision, int full_asymmetric,
                                   float* S_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  return sdav_similarity_impl(desc_dev, N, P, D, mu, sigma, a, b, w_dev, precision, full_asymmetric, 0, 1, S_dev,
                              ws_dev, ws_bytes, stream);
}

extern "C" int dlc_sdav_similarity_part(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a,
                                        double b, const double* w_dev, int precision, int full_asymmetric, int part,
                                        int n_parts, float* S_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(n_parts >= 1 && part >= 0 && part < n_parts);
  return sdav_similarity_impl(desc_dev, N, P, D, mu, sigma, a, b, w_dev, precision, full_asymmetric, part, n_parts,
                              S_dev, ws_dev, ws_bytes, stream);
}

static int sdav_similarity_impl(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a,
                                double b, const double* w_dev, int precision, int full_asymmetric, int part,
                                int n_parts, float* S_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(desc_dev && S_dev && ws_dev);
  DLC_CHECK_ARG(N >= 1 && N <= (1 << 20));
  DLC_CHECK_ARG(P >= 1 && P <= kFrameRows);
  DLC_CHECK_ARG(D >= 1);
  DLC_CHECK_ARG(sigma != 0.0);
  DLC_CHECK_ARG(precision == DLC_PREC_FP16 || precision == DLC_PREC_FP16X2 || precision == DLC_PREC_AUTO ||
                precision == DLC_PREC_FP16_REFINED);
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0);
  const SimWorkspace L = sim_layout(N, P, D);
  if (ws_bytes < L.total)
    return fail(DLC_ENOMEM, "dlc_sdav_similarity: workspace of %zu bytes needed, %zu given", L.total, ws_bytes);
  cudaStream_t s = as_stream(stream);
  char* ws = static_cast<char*>(ws_dev);
  const int64_t rows = static_cast<int64_t>(N) * P;

  double* colsum_part = reinterpret_cast<double*>(ws + L.off_part);
  double* w = reinterpret_cast<double*>(ws + L.off_w);
  float* mean = reinterpret_cast<float*>(ws + L.off_mean);   // centring vector: the dataset mean, or zero
  float* sqn = reinterpret_cast<float*>(ws + L.off_sqn);
  double* pw = reinterpret_cast<double*>(ws + L.off_pw);
  uint32_t* rep = reinterpret_cast<uint32_t*>(ws + L.off_rep);
  const bool probe = precision == DLC_PREC_AUTO || precision == DLC_PREC_FP16_REFINED;
  ProbeAccum* acc = reinterpret_cast<ProbeAccum*>(ws + L.off_probe);
  float* gaps = reinterpret_cast<float*>(ws + L.off_gaps);
  const int gram_only = g_gram_only.load();
  TileListEntry lists;
  if (int rc = get_tile_lists(N, full_asymmetric, part, n_parts, &lists)) return rc;
  if (!gram_only) {
    // 1. dataset mean -> distinctive weights w and the centring vector of the planes (w_dev given: weights of another
    //    dataset, SimilarityCalculator.similarity_score on frames outside it; the centring is still this dataset's)
    colsum_partial_kernel<<<dim3(ceil_div(D, 32), kColSumSlabs), 32 * kColSumRowLanes, 0, s>>>(desc_dev, rows, D, colsum_part);
    // the slabs are summed by a wide kernel first: the single-block weights kernel reading all of them took 0.14 ms
    double* colsum_total = colsum_part + static_cast<size_t>(kColSumSlabs) * 2 * D;
    colsum_reduce_kernel<<<ceil_div(2 * D, 128), 128, 0, s>>>(colsum_part, kColSumSlabs, 2 * D, colsum_total);
    weights_centre_kernel<<<1, kWeightsThreads, 0, s>>>(colsum_total, 1, rows, D, mu, sigma, w_dev ? nullptr : w, mean);
    if (w_dev) w = const_cast<double*>(w_dev);
    // 2. one pass: centred operand planes (P rows per frame, K padded with zeros), per-row squared norms of the centred
    //    rows, projections p = h . w and a content hash of every row. With a precision probe (auto / fp16r) the
    //    residual planes are written later and only if the probe picks the three-product kernel.
    const bool lo_now = precision == DLC_PREC_FP16X2;
    unsigned long long* rowhash = reinterpret_cast<unsigned long long*>(ws + L.off_hash);
    if (probe) DLC_CUDA(cudaMemsetAsync(acc, 0, sizeof(ProbeAccum), s));
    prep_rows_kernel<<<ceil_div(N * kFrameRows, 8), 256, 0, s>>>(
        desc_dev, N, P, D, w, mean, reinterpret_cast<__half*>(ws + L.off_bhi),
        lo_now ? reinterpret_cast<__half*>(ws + L.off_blo) : nullptr, L.ld, sqn, pw, probe ? &acc->nmax_bits : nullptr,
        probe ? rowhash : nullptr);
    // 3. precision probe: classes of bit-identical rows (from the hashes, verified), the single-product error of the
    //    centred values on sampled rows, then the margin and the device-side choice between the two Gram kernels
    if (probe) {
      GramControl* ctl = reinterpret_cast<GramControl*>(ws + L.off_ctl);
      rep_from_hash_kernel<<<ceil_div(N, 8), 256, 0, s>>>(desc_dev, N, P, D, rowhash, rep);
      if (N >= 2) {
        gram_probe_kernel<<<kProbeSamples, 1024, 0, s>>>(desc_dev, N, P, D, mean, rep, acc, gaps);
      } else {
        DLC_CUDA(cudaMemsetAsync(gaps, 0x7f, sizeof(float) * kProbeSamples, s));  // large gaps: nothing to refine
      }
      gram_probe_finalize_kernel<<<1, 256, 0, s>>>(acc, gaps, g_max_flag_frac,
                                                   precision == DLC_PREC_FP16_REFINED ? 1 : -1, ctl);
      if (precision == DLC_PREC_AUTO)  // fp16r never runs the three-product kernel
        lo_planes_kernel<<<static_cast<int>(ceil_div64(rows, 8)), 256, 0, s>>>(
            desc_dev, rows, D, mean, reinterpret_cast<__half*>(ws + L.off_blo), L.ld, ctl);
    }
    DLC_CUDA(cudaGetLastError());
    // a part of the matrix: entries other parts own stay zero, so the parts combine with a sum (all-reduce)
    if (n_parts > 1) DLC_CUDA(cudaMemsetAsync(S_dev, 0, sizeof(float) * static_cast<size_t>(N) * N, s));
  }  // !gram_only

  // 5. Gram + argmin + score
  GramParams p{};
  p.rep_mask = rep;
  p.col_stride = L.col_stride;
  p.b_frame_map = L.col_stride == kFrameRows ? 1 : 0;
  p.n_tile = kFramesPerNTile * L.col_stride;  // 256, or 240 for 30 patches per frame
  p.ab_fmt = 0;
  p.tiles = lists.tiles;
  p.num_tiles = lists.num_tiles;
  p.pair_tiles = lists.pair_tiles;
  p.num_pair_tiles = lists.num_pair_tiles;
  p.N = N;
  p.P = P;
  p.D = D;
  p.desc = desc_dev;
  p.sqn = sqn;
  p.pw = pw;
  p.a = static_cast<float>(a);
  p.b = static_cast<float>(b);
  p.full = full_asymmetric ? 1 : 0;
  p.S = S_dev;
  p.ctl = nullptr;
  p.want_refine = 0;
  // enough pair tiles to fill the GPU and an N tile that splits into two UMMA-legal halves: CTA pairs
  const bool pairs = use_pairs(p);
  const GramPlanes planes{ws + L.off_bhi, ws + L.off_blo, L.ld, L.col_stride};
  if (precision == DLC_PREC_FP16X2)
    return pairs ? run_gram_pair<GramPolicy<32, 3>>(planes, p, s) : run_gram<GramPolicy<32, 3>>(planes, p, s);
  if (precision == DLC_PREC_FP16)
    return pairs ? run_gram_pair<GramPolicy<64, 1>>(planes, p, s) : run_gram<GramPolicy<64, 1>>(planes, p, s);

  // DLC_PREC_AUTO / DLC_PREC_FP16_REFINED: probe the single-product error on the data, then launch both kernels; the
  // device-side control block lets exactly one of them run (no host round trip).
  GramControl* ctl = reinterpret_cast<GramControl*>(ws + L.off_ctl);
  p.ctl = ctl;
  p.want_refine = 1;
  const int cap = g_refine_cap.load();
  p.work_cap = cap >= 0 ? std::min(cap, L.work_cap) : L.work_cap;
  p.work = p.work_cap > 0 ? reinterpret_cast<RefineEntry*>(ws + L.off_work) : nullptr;
  if (gram_only) DLC_CUDA(cudaMemsetAsync(&ctl->n_entries, 0, sizeof(unsigned int), s));  // else reset by the probe
  if (int rc = pairs ? run_gram_pair<GramRefinePolicy>(planes, p, s) : run_gram<GramRefinePolicy>(planes, p, s)) return rc;
  if (p.work && gram_only != 2) {
    gram_refine_fix_kernel<<<4 * sm_count(), kFixThreads, 0, s>>>(p);
    DLC_CUDA(cudaGetLastError());
  }
  p.work = nullptr;
  p.want_refine = 0;
  return pairs ? run_gram_pair<GramPolicy<32, 3>>(planes, p, s) : run_gram<GramPolicy<32, 3>>(planes, p, s);
}

// ------------------------------------------------------------------------------------------------------------
// Staged form of the same computation, for ONE sequence whose frames are split over the GPUs of a box:
// every stage works on rank-local rows and writes into this rank's slice of arrays that the host exchanges with
// NCCL all-gathers between the stages (frames are dealt in contiguous blocks of `frames_per_part`, so the gathered
// arrays are the flat global ones). Nothing is computed twice: the dataset mean, the row statistics, the operand
// planes and the precision probe are all produced once, by the rank that encoded the frame.
//   stage_colsum   local descriptors -> column sums [D]                       | all-gather [n_parts, D] doubles
//   stage_weights  all column sums   -> mean [D], weights w [D]               | (identical on every rank)
//   stage_prepare  local descriptors -> centred plane slice + stats block     | all-gather planes, all-gather stats
//   stage_gram     planes + stats    -> this part's tile rows of S + the list of ambiguous pairs
//   stage_fix      float32 descriptors of ALL frames -> exact scores of the listed pairs   (after their all-gather,
//                  which the host overlaps with stage_gram)
// ------------------------------------------------------------------------------------------------------------
namespace dlc {
struct StageStats {  // layout of one part's stats block
  size_t off_sqn, off_pw, off_rep, off_probe, off_gaps, off_hash, total;
};
static StageStats stage_stats_layout(int frames_per_part) {
  StageStats t{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up(o + bytes, 16);
    return at;
  };
  t.off_sqn = take(sizeof(float) * frames_per_part * kFrameRows);
  t.off_pw = take(sizeof(double) * frames_per_part * kFrameRows);
  t.off_rep = take(sizeof(uint32_t) * frames_per_part);
  t.off_probe = take(sizeof(ProbeAccum));
  t.off_gaps = take(sizeof(float) * kProbeSamples);
  t.off_hash = take(sizeof(unsigned long long) * frames_per_part * kFrameRows);  // scratch of stage_prepare (exchanged, unused)
  t.total = align_up(o, 256);
  return t;
}
struct StageWorkspace {
  size_t off_part, off_sqn, off_pw, off_rep, off_ctl, off_work, total;
  int work_cap;
};
static StageWorkspace stage_layout(int N, int D, int n_parts) {
  StageWorkspace w{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o = align_up(o + bytes, 256);
    return at;
  };
  w.off_part = take(sizeof(double) * kColSumSlabs * 2 * D);
  w.off_sqn = take(sizeof(float) * N * kFrameRows);
  w.off_pw = take(sizeof(double) * N * kFrameRows);
  w.off_rep = take(sizeof(uint32_t) * N);
  w.off_ctl = take(sizeof(GramControl));
  // one entry per frame pair this part can own (+ slack for the interleaving): the list never overflows, so the Gram
  // kernel itself never reads the float32 descriptors - their exchange overlaps it
  const int64_t pairs = (static_cast<int64_t>(N) * N / 2) / n_parts + 64LL * N + 1024;
  w.work_cap = static_cast<int>(std::min<int64_t>(pairs, int64_t{1} << 22));
  w.off_work = take(sizeof(RefineEntry) * static_cast<size_t>(w.work_cap));
  w.total = o;
  return w;
}

// stats blocks of all parts -> flat sqn [N*32], pw [N*32], rep [N] (frame g lives in block g / per at g % per)
__global__ void stage_unpack_kernel(const uint8_t* __restrict__ blocks, size_t block_bytes, StageStats t, int per,
                                    int N, float* __restrict__ sqn, double* __restrict__ pw,
                                    uint32_t* __restrict__ rep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // padded row index
  if (i >= N * kFrameRows) return;
  const int g = i / kFrameRows, k = i % kFrameRows;
  const uint8_t* b = blocks + static_cast<size_t>(g / per) * block_bytes;
  const int l = g % per;
  sqn[i] = reinterpret_cast<const float*>(b + t.off_sqn)[l * kFrameRows + k];
  pw[i] = reinterpret_cast<const double*>(b + t.off_pw)[l * kFrameRows + k];
  if (k == 0) rep[g] = reinterpret_cast<const uint32_t*>(b + t.off_rep)[l];
}

// probe blocks of all parts -> one control block (same rule as gram_probe_finalize_kernel, over all samples)
__global__ void stage_probe_finalize_kernel(const uint8_t* __restrict__ blocks, size_t block_bytes, StageStats t,
                                            int n_parts, GramControl* ctl) {
  __shared__ int s_cnt, s_tot;
  __shared__ float s_margin, s_rms;
  if (threadIdx.x == 0) {
    double sum_err2 = 0.0;
    unsigned long long n_err = 0, max_row = 0;
    unsigned int nmax = 0;
    for (int r = 0; r < n_parts; ++r) {
      const ProbeAccum* a = reinterpret_cast<const ProbeAccum*>(blocks + r * block_bytes + t.off_probe);
      sum_err2 += a->sum_err2;
      n_err += a->n_err;
      max_row = a->max_row_bits > max_row ? a->max_row_bits : max_row;
      nmax = a->nmax_bits > nmax ? a->nmax_bits : nmax;
    }
    const double rms = sqrt(sum_err2 / static_cast<double>(n_err > 0 ? n_err : 1));
    const double rms_max = sqrt(__longlong_as_double(static_cast<long long>(max_row)));
    s_margin = static_cast<float>(1.41421356 * fmax(8.0 * rms, 4.0 * rms_max)) + 4e-6f * __uint_as_float(nmax);
    s_rms = static_cast<float>(rms);
    s_cnt = 0;
    s_tot = 0;
  }
  __syncthreads();
  int cnt = 0, tot = 0;
  for (int i = threadIdx.x; i < n_parts * kProbeSamples; i += blockDim.x) {
    const float g = reinterpret_cast<const float*>(blocks + (i / kProbeSamples) * block_bytes + t.off_gaps)[i % kProbeSamples];
    if (g < 1e30f) {   // parts with fewer than two frames sample nothing (their gaps are a large sentinel)
      ++tot;
      cnt += g < s_margin ? 1 : 0;
    }
  }
  atomicAdd(&s_cnt, cnt);
  atomicAdd(&s_tot, tot);
  __syncthreads();
  if (threadIdx.x == 0) {
    ctl->margin = s_margin;
    ctl->sigma = s_rms;
    ctl->flagged_frac = s_tot ? static_cast<float>(s_cnt) / s_tot : 0.0f;
    ctl->use_refine = 1;
    ctl->flagged_rows = 0;
    ctl->refined_cands = 0;
    ctl->n_entries = 0;
  }
}
}  // namespace dlc

extern "C" size_t dlc_sdav_stage_stats_bytes(int frames_per_part) {
  return frames_per_part > 0 ? stage_stats_layout(frames_per_part).total : 0;
}
extern "C" size_t dlc_sdav_stage_workspace_bytes(int N, int P, int D, int n_parts) {
  if (N <= 0 || P <= 0 || D <= 0 || n_parts <= 0) return 0;
  return stage_layout(N, D, n_parts).total;
}

extern "C" int dlc_sdav_stage_colsum(const float* desc_local_dev, int64_t rows_local, int D, double* colsum_dev,
                                     void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(colsum_dev && ws_dev && D >= 1 && rows_local >= 0);
  DLC_CHECK_ARG(desc_local_dev || rows_local == 0);
  if (ws_bytes < sizeof(double) * kColSumSlabs * 2 * D)
    return fail(DLC_ENOMEM, "dlc_sdav_stage_colsum: workspace of %zu bytes needed", sizeof(double) * kColSumSlabs * 2 * D);
  cudaStream_t s = as_stream(stream);
  double* part = static_cast<double*>(ws_dev);
  colsum_partial_kernel<<<dim3(ceil_div(D, 32), kColSumSlabs), 32 * kColSumRowLanes, 0, s>>>(desc_local_dev, rows_local, D, part);
  colsum_reduce_kernel<<<ceil_div(2 * D, 128), 128, 0, s>>>(part, kColSumSlabs, 2 * D, colsum_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_sdav_stage_weights(const double* colsums_dev, int n_parts, int64_t rows_total, int D, double mu,
                                      double sigma, double* w_dev, float* centre_dev, void* stream) {
  DLC_CHECK_ARG(colsums_dev && w_dev && centre_dev);
  DLC_CHECK_ARG(n_parts >= 1 && rows_total >= 1 && D >= 1 && sigma != 0.0);
  weights_centre_kernel<<<1, kWeightsThreads, 0, as_stream(stream)>>>(colsums_dev, n_parts, rows_total, D, mu, sigma,
                                                                      w_dev, centre_dev);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

extern "C" int dlc_sdav_stage_prepare(const float* desc_local_dev, int n_local, int frames_per_part, int P, int D,
                                      const double* w_dev, const float* centre_dev, int precision,
                                      void* plane_hi_local_dev, void* plane_lo_local_dev, void* stats_local_dev,
                                      void* stream) {
  DLC_CHECK_ARG(stats_local_dev && w_dev && centre_dev);
  DLC_CHECK_ARG(n_local >= 0 && n_local <= frames_per_part && P >= 1 && P <= kFrameRows && D >= 1);
  DLC_CHECK_ARG(n_local == 0 || (desc_local_dev && plane_hi_local_dev));
  DLC_CHECK_ARG(precision != DLC_PREC_FP16X2 || plane_lo_local_dev || n_local == 0);
  cudaStream_t s = as_stream(stream);
  const StageStats t = stage_stats_layout(frames_per_part);
  uint8_t* st = static_cast<uint8_t*>(stats_local_dev);
  DLC_CUDA(cudaMemsetAsync(st, 0, t.total, s));
  float* gaps = reinterpret_cast<float*>(st + t.off_gaps);
  DLC_CUDA(cudaMemsetAsync(gaps, 0x7f, sizeof(float) * kProbeSamples, s));  // "nothing sampled" sentinel (3.4e38)
  if (n_local == 0) return DLC_OK;
  const int ld = dlc_plane_ld(D);
  ProbeAccum* acc = reinterpret_cast<ProbeAccum*>(st + t.off_probe);
  uint32_t* rep = reinterpret_cast<uint32_t*>(st + t.off_rep);
  const bool probe = precision == DLC_PREC_AUTO || precision == DLC_PREC_FP16_REFINED;
  prep_rows_kernel<<<ceil_div(n_local * kFrameRows, 8), 256, 0, s>>>(
      desc_local_dev, n_local, P, D, w_dev, centre_dev, static_cast<__half*>(plane_hi_local_dev),
      precision == DLC_PREC_FP16X2 ? static_cast<__half*>(plane_lo_local_dev) : nullptr, ld,
      reinterpret_cast<float*>(st + t.off_sqn), reinterpret_cast<double*>(st + t.off_pw),
      probe ? &acc->nmax_bits : nullptr, probe ? reinterpret_cast<unsigned long long*>(st + t.off_hash) : nullptr);
  if (probe) {
    rep_from_hash_kernel<<<ceil_div(n_local, 8), 256, 0, s>>>(
        desc_local_dev, n_local, P, D, reinterpret_cast<const unsigned long long*>(st + t.off_hash), rep);
    if (n_local >= 2)
      gram_probe_kernel<<<kProbeSamples, 1024, 0, s>>>(desc_local_dev, n_local, P, D, centre_dev, rep, acc, gaps);
  }
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

static int stage_params(const void* plane_hi_all_dev, const void* plane_lo_all_dev, const float* desc_all_dev, int N,
                        int P, int D, double a, double b, int precision, int full_asymmetric, int part, int n_parts,
                        float* S_dev, char* ws, const StageWorkspace& L, GramParams* out, GramPlanes* planes) {
  TileListEntry lists;
  if (int rc = get_tile_lists(N, full_asymmetric, part, n_parts, &lists)) return rc;
  const int col_stride = (P < kFrameRows && (P & 1) == 0) ? P : kFrameRows;
  GramParams p{};
  p.rep_mask = reinterpret_cast<const uint32_t*>(ws + L.off_rep);
  p.col_stride = col_stride;
  p.b_frame_map = col_stride == kFrameRows ? 1 : 0;
  p.n_tile = kFramesPerNTile * col_stride;
  p.ab_fmt = 0;
  p.tiles = lists.tiles;
  p.num_tiles = lists.num_tiles;
  p.pair_tiles = lists.pair_tiles;
  p.num_pair_tiles = lists.num_pair_tiles;
  p.N = N;
  p.P = P;
  p.D = D;
  p.desc = desc_all_dev;
  p.sqn = reinterpret_cast<const float*>(ws + L.off_sqn);
  p.pw = reinterpret_cast<const double*>(ws + L.off_pw);
  p.a = static_cast<float>(a);
  p.b = static_cast<float>(b);
  p.full = full_asymmetric ? 1 : 0;
  p.S = S_dev;
  const bool refine = precision == DLC_PREC_AUTO || precision == DLC_PREC_FP16_REFINED;
  p.ctl = refine ? reinterpret_cast<GramControl*>(ws + L.off_ctl) : nullptr;
  p.want_refine = refine ? 1 : 0;
  p.work = refine ? reinterpret_cast<RefineEntry*>(ws + L.off_work) : nullptr;
  p.work_cap = refine ? L.work_cap : 0;
  *out = p;
  *planes = GramPlanes{plane_hi_all_dev, plane_lo_all_dev, dlc_plane_ld(D), col_stride};
  return DLC_OK;
}

extern "C" int dlc_sdav_stage_gram(const void* plane_hi_all_dev, const void* plane_lo_all_dev,
                                   const void* stats_all_dev, int n_parts, int frames_per_part, int N, int P, int D,
                                   double a, double b, int precision, int full_asymmetric, int part, float* S_dev,
                                   void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(plane_hi_all_dev && stats_all_dev && S_dev && ws_dev);
  DLC_CHECK_ARG(N >= 1 && P >= 1 && P <= kFrameRows && D >= 1);
  DLC_CHECK_ARG(n_parts >= 1 && part >= 0 && part < n_parts && frames_per_part >= 1 &&
                static_cast<int64_t>(frames_per_part) * n_parts >= N);
  DLC_CHECK_ARG(precision == DLC_PREC_FP16 || precision == DLC_PREC_FP16X2 || precision == DLC_PREC_AUTO ||
                precision == DLC_PREC_FP16_REFINED);
  DLC_CHECK_ARG(precision != DLC_PREC_FP16X2 || plane_lo_all_dev);
  DLC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws_dev) & 255) == 0);
  const StageWorkspace L = stage_layout(N, D, n_parts);
  if (ws_bytes < L.total)
    return fail(DLC_ENOMEM, "dlc_sdav_stage_gram: workspace of %zu bytes needed, %zu given", L.total, ws_bytes);
  cudaStream_t s = as_stream(stream);
  char* ws = static_cast<char*>(ws_dev);
  const StageStats t = stage_stats_layout(frames_per_part);
  const uint8_t* blocks = static_cast<const uint8_t*>(stats_all_dev);
  stage_unpack_kernel<<<ceil_div(N * kFrameRows, 256), 256, 0, s>>>(
      blocks, t.total, t, frames_per_part, N, reinterpret_cast<float*>(ws + L.off_sqn),
      reinterpret_cast<double*>(ws + L.off_pw), reinterpret_cast<uint32_t*>(ws + L.off_rep));
  GramParams p;
  GramPlanes planes;
  if (int rc = stage_params(plane_hi_all_dev, plane_lo_all_dev, nullptr, N, P, D, a, b, precision, full_asymmetric,
                            part, n_parts, S_dev, ws, L, &p, &planes))
    return rc;
  if (p.ctl) stage_probe_finalize_kernel<<<1, 256, 0, s>>>(blocks, t.total, t, n_parts, p.ctl);
  DLC_CUDA(cudaGetLastError());
  if (n_parts > 1) DLC_CUDA(cudaMemsetAsync(S_dev, 0, sizeof(float) * static_cast<size_t>(N) * N, s));
  const bool pairs = use_pairs(p);
  if (precision == DLC_PREC_FP16X2)
    return pairs ? run_gram_pair<GramPolicy<32, 3>>(planes, p, s) : run_gram<GramPolicy<32, 3>>(planes, p, s);
  if (precision == DLC_PREC_FP16)
    return pairs ? run_gram_pair<GramPolicy<64, 1>>(planes, p, s) : run_gram<GramPolicy<64, 1>>(planes, p, s);
  // one product + deferred exact refinement (AUTO is treated as FP16_REFINED here: choosing the three-product kernel
  // on the device would need the residual planes of every rank, i.e. a second exchange decided without the host)
  return pairs ? run_gram_pair<GramRefinePolicy>(planes, p, s) : run_gram<GramRefinePolicy>(planes, p, s);
}

extern "C" int dlc_sdav_stage_fix(const void* plane_hi_all_dev, const float* desc_all_dev, int N, int P, int D, double a,
                                  double b, int precision, int full_asymmetric, int part, int n_parts, float* S_dev,
                                  void* ws_dev, size_t ws_bytes, void* stream) {
  DLC_CHECK_ARG(desc_all_dev && S_dev && ws_dev && plane_hi_all_dev);
  DLC_CHECK_ARG(N >= 1 && P >= 1 && P <= kFrameRows && D >= 1 && n_parts >= 1 && part >= 0 && part < n_parts);
  if (!(precision == DLC_PREC_AUTO || precision == DLC_PREC_FP16_REFINED)) return DLC_OK;  // nothing was deferred
  const StageWorkspace L = stage_layout(N, D, n_parts);
  if (ws_bytes < L.total)
    return fail(DLC_ENOMEM, "dlc_sdav_stage_fix: workspace of %zu bytes needed, %zu given", L.total, ws_bytes);
  GramParams p;
  GramPlanes planes;
  if (int rc = stage_params(plane_hi_all_dev, nullptr, desc_all_dev, N, P, D, a, b, precision, full_asymmetric, part,
                            n_parts, S_dev, static_cast<char*>(ws_dev), L, &p, &planes))
    return rc;
  gram_refine_fix_kernel<<<4 * sm_count(), kFixThreads, 0, as_stream(stream)>>>(p);
  DLC_CUDA(cudaGetLastError());
  return DLC_OK;
}

// Diagnostics of the last AUTO / FP16_REFINED call that used this workspace: out_host[0..5] = use_refine, margin,
// sigma, estimated flagged fraction, flagged rows, refined candidates. Synchronises the stream.
extern "C" int dlc_sdav_similarity_stats(int N, int P, int D, const void* ws_dev, double* out_host, void* stream) {
  DLC_CHECK_ARG(ws_dev && out_host && N >= 1 && P >= 1 && D >= 1);
  const SimWorkspace L = sim_layout(N, P, D);
  GramControl c;
  DLC_CUDA(cudaMemcpyAsync(&c, static_cast<const char*>(ws_dev) + L.off_ctl, sizeof(c), cudaMemcpyDeviceToHost,
                           as_stream(stream)));
  DLC_CUDA(cudaStreamSynchronize(as_stream(stream)));
  out_host[0] = c.use_refine;
  out_host[1] = c.margin;
  out_host[2] = c.sigma;
  out_host[3] = c.flagged_frac;
  out_host[4] = static_cast<double>(c.flagged_rows);
  out_host[5] = static_cast<double>(c.refined_cands);
  return DLC_OK;
}

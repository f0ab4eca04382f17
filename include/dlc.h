/* libdlc - C ABI of the B200-native loop-closure hot path (descriptor extraction + keyframe matching).
 *
 * This header is the drop-in boundary. The reference (nschejtman/deepLoopCloser) has no FFI of its own: its hot
 * path is a handful of Python methods that hand NumPy arrays to TensorFlow-1 / NumPy / CPython loops. Each entry
 * point below names the reference call it replaces (file:line relative to the reference tree); INTEGRATION.md
 * shows the ctypes stub a maintainer of the reference would add for each.
 *
 * Conventions
 *  - every function returns 0 (DLC_OK) or a negative DLC_E* code; dlc_last_error() gives a thread-local message;
 *  - pointers named *_dev are DEVICE pointers owned by the caller, pointers named *_host are host pointers;
 *  - every launch goes on the cudaStream_t passed in (as void*; 0 = legacy default stream); no hidden syncs,
 *    no allocations on the hot path: scratch memory is caller-provided (`ws_dev`, size from *_workspace_bytes);
 *  - a handle is not thread-safe; distinct handles may be used from distinct threads; one process per GPU;
 *  - ONE STREAM PER HANDLE / WORKSPACE: calls that share a handle (dlc_db_append then dlc_match_topk; dlc_sda_* on one
 *    encoder) or a workspace buffer must be issued on the same stream, or the caller orders them with events - the
 *    library records none. (dlc_db_append advances the row count on the host when the append kernel is ENQUEUED; a
 *    match on another stream would read rows that kernel has not written yet.)
 *  - dlc_debug_set / dlc_sdav_debug_gram_only are process-wide developer switches for A/B measurements (atomic
 *    integers); they are not part of the drop-in surface and must not be flipped while other threads are in a call;
 *  - sm_100a only. There is no CPU fallback anywhere behind this ABI.
 */
#ifndef DLC_H_
#define DLC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLC_OK 0
#define DLC_EINVAL (-1)       /* bad argument */
#define DLC_ECUDA (-2)        /* CUDA runtime / driver error (message has the CUDA string) */
#define DLC_ENOMEM (-3)       /* workspace too small or allocation failed */
#define DLC_EUNSUPPORTED (-4) /* device is not sm_100 or shape outside the supported envelope */

/* Arithmetic modes of the tensor-core contractions (fp32 accumulation in TMEM in all of them). */
#define DLC_PREC_FP16 0   /* one fp16 product                           (~1e-3 on well-scaled weights)          */
#define DLC_PREC_FP16X2 1 /* fp16 hi/lo split, 3 products, ~22-bit operands (meets 1e-3 on N(0,1) weights too) */
#define DLC_PREC_BF16 2   /* one bf16 product (matcher databases stored as bf16)                                */
/* dlc_sdav_similarity only: */
#define DLC_PREC_AUTO 3         /* probe the data on the device, then FP16_REFINED if few rows need refinement,  */
                                /* else FP16X2                                                                    */
#define DLC_PREC_FP16_REFINED 4 /* one fp16 product; rows whose nearest-neighbour choice is ambiguous under the   */
                                /* fp16 rounding error are re-evaluated exactly (float64-accumulated distances)   */

/* dlc_sda_* only: */
#define DLC_PREC_FP16X2_A16 5   /* two products: weights split hi/lo, activations rounded to fp16 between layers    */
/* DLC_PREC_AUTO on dlc_sda_create: the handle keeps split weights and picks the cheapest of FP16 (1 product),
 * FP16X2_A16 (2) and FP16X2 (3) whose sampled descriptors stay within 3e-4 of the three-product forward on the
 * handle's OWN weights and the caller's own input (dlc_sda_probe; run implicitly by the first dlc_sda_encode). The
 * reference restores trained checkpoints when it has them (SDAV.py:232-240) - well-scaled weights pass with one
 * product - and otherwise draws N(0,1) weights (:189-217), whose saturating pre-activations need all three. */

/* dtypes for untyped buffers */
#define DLC_F32 0
#define DLC_F64 1
#define DLC_F16 2
#define DLC_BF16 3
#define DLC_U8 4
#define DLC_I32 5

/* activations of dlc_gemm_planes */
#define DLC_ACT_NONE 0
#define DLC_ACT_SIGMOID 1
#define DLC_ACT_RELU 2

/* matcher metrics */
#define DLC_METRIC_COS 0 /* cosine similarity, larger = closer  */
#define DLC_METRIC_DOT 1 /* inner product, larger = closer      */
#define DLC_METRIC_L2 2  /* squared L2 distance, smaller = closer; reported score is the distance */

const char* dlc_last_error(void);
int dlc_version(void);
/* Developer switches for A/B measurements and tests; not part of the drop-in surface, defaults in brackets.
 *   0: K block of the 3-product kernel, 32 or 64 [32]         1: (see dlc_sdav_debug_gram_only)
 *   2: K elements accumulated in TMEM before promotion [256]   3: kernel experiment flags [0]
 *   4: M tiles per L2 super-block of the SDAV tile order [32]  5: plane outputs through staged TMA stores [1]
 *   6: CTA-pair kernels: 0 never, 1 when the problem fills the GPU, 2 whenever the shape allows [1]
 *   7: SDAV Gram kernels on CTA pairs [1]
 *   8: capacity of the deferred-refinement list of the SDAV score kernel; 0 = refine inside the epilogue [-1: default]
 *   9: SDAV precision probe on its side stream [1] */
/* SMs left free by the library's persistent kernels (tensor-core contractions, second pass): their grids are sized for
 * (SM count - sms), 0 <= sms <= 64; 0 is the default. Meant for a sequence split over GPUs, so that NCCL's kernels,
 * launched on another stream during a persistent kernel, find SMs at once (measured on 8 GPUs: no gain beyond the
 * run-to-run noise, DESIGN.md 6 - hence off by default; ShardedSequencePipeline.sm_reserve / DLC_SM_RESERVE set it).
 * Process-wide. */
int dlc_set_sm_reserve(int sms);
/* Developer A/B switches (key: value): 0 K block of split-precision GEMMs (32 | 64); 2 K elements accumulated in TMEM
 * before promotion; 3 epilogue flags (1 skip stores, 2 skip activation, 4 skip TMA store issue, 8 skip min/max);
 * 4 Gram M-group; 5 TMA plane stores on/off; 6 CTA pairs (0 never, 1 auto, 2 always); 7 Gram pairs; 8 capacity of
 * the deferred-refinement list; 9 (no-op); 10 keypoint detector: shared-memory octave kernel on/off; 11 encoder:
 * per-call GEMM width on/off. */
int dlc_debug_set(int key, int value);
/* Timing aid: 1 = dlc_sdav_similarity launches only the Gram / score kernels (+ refinement pass) on the operand
 * planes, statistics and tile list a previous full call left in the workspace; 2 = without the refinement pass
 * (results incomplete); 0 = normal. */
int dlc_sdav_debug_gram_only(int on);
/* 0 when the current CUDA device can run this library (compute capability 10.x), DLC_EUNSUPPORTED otherwise. */
int dlc_device_check(void);
int dlc_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Operand planes. Every tensor-core contraction consumes K-major fp16 "planes": a row-major [rows, ld] fp16
 * matrix with ld a multiple of 64 (zero padded), optionally accompanied by a second plane holding the fp16
 * rounding residual (DLC_PREC_FP16X2).
 * ------------------------------------------------------------------------------------------------------------ */
/* ld (in elements) of a plane holding `cols` valid columns: cols rounded up to a multiple of 64. */
int dlc_plane_ld(int cols);

/* src_dev [rows, cols] (DLC_F32 or DLC_F64, row pitch src_ld elements) -> hi_dev / lo_dev [rows_out, ld] fp16.
 * Rows are regrouped on the way: source row r goes to plane row (r / group_in) * group_out + r % group_in
 * (group_in = group_out = 1 for a plain copy). Pad rows/columns are written as zeros. lo_dev may be NULL.
 * Replaces: the float64 feed of tf.Session.run (src/sdav/network/SDAV.py:297-302). */
int dlc_split_planes(const void* src_dev, int src_dtype, int rows, int cols, int src_ld, int group_in, int group_out,
                     void* hi_dev, void* lo_dev, int ld, void* stream);

/* W_dev [k, n] row-major (DLC_F32/DLC_F64; the reference's weight layout, SDAV.py:189-217) -> transposed K-major
 * planes wt_hi_dev / wt_lo_dev [n_pad, ld] with ld = dlc_plane_ld(k); rows n..n_pad-1 and columns k..ld-1 zero. */
int dlc_pack_weight_planes(const void* w_dev, int src_dtype, int k, int n, int n_pad, void* wt_hi_dev,
                           void* wt_lo_dev, int ld, void* stream);

/* out = act(A * B^T + bias):  A planes [m, ld] , B planes [n_pad, ld] (n_pad multiple of n_tile), k = ld.
 * Outputs (each optional): out_f32_dev [m, n] with pitch out_ld; out_hi/out_lo planes [m, out_plane_ld] whose
 * columns >= n are written as zero (so they can feed the next contraction directly).
 * precision: DLC_PREC_FP16 (lo planes ignored), DLC_PREC_FP16X2, DLC_PREC_BF16 (planes hold bf16). With
 * DLC_PREC_FP16X2, a_lo_dev == NULL declares the A values exact in fp16 (two products per K step instead of three).
 * Replaces: TensorWrapper.matmul/add/sigmoid (src/utils/TensorflowWrapper.py:57-78) and tf.layers.conv2d's
 * matmul core (src/cnn_vtl/network/cnn_vtl.py:33-93). */
int dlc_gemm_planes(const void* a_hi_dev, const void* a_lo_dev, const void* b_hi_dev, const void* b_lo_dev, int m,
                    int n, int n_pad, int ld, const float* bias_dev, int act, int precision, float* out_f32_dev,
                    int out_ld, void* out_hi_dev, void* out_lo_dev, int out_plane_ld, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * a1  patch gather + normalise.   Replaces get_vectorized_patches_from_key_points + get_1d/2d_boundaries + /255.0
 *     (src/sdav/input/CvInputParser.py:100-123, 49-97, 27).
 * img_dev uint8 [B,H,W]; xy_dev float32 [B,P,2] keypoint (x = column, y = row) centres, rounded half-to-even
 * like Python's round(); patch is odd. swap_xy_quirk = 1 reproduces the reference (rows are indexed by x and
 * clamped against H, columns by y against W); 0 is the geometrically intended behaviour.
 * ------------------------------------------------------------------------------------------------------------ */
/* fp16 planes [B*P, ld] (ld = dlc_plane_ld(patch*patch)), value = pixel/255 split into hi/lo. lo may be NULL. */
int dlc_patch_gather(const uint8_t* img_dev, int B, int H, int W, const float* xy_dev, int P, int patch,
                     int swap_xy_quirk, void* out_hi_dev, void* out_lo_dev, int ld, void* stream);
/* One fp16 plane [B*P, ld] holding the pixel VALUE (0..255, exact in fp16) instead of pixel/255: input of an encoder
 * in raw-pixel mode (dlc_sda_set_input_u8), which folds the /255.0 of CvInputParser.py:27 into its first layer. */
int dlc_patch_gather_u8(const uint8_t* img_dev, int B, int H, int W, const float* xy_dev, int P, int patch,
                        int swap_xy_quirk, void* out_dev, int ld, void* stream);
/* float64 [B*P, patch*patch], bit-identical to the reference's ndarray. */
int dlc_patch_gather_f64(const uint8_t* img_dev, int B, int H, int W, const float* xy_dev, int P, int patch,
                         int swap_xy_quirk, double* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * f3  keypoint detector in front of the patch gather. Replaces get_top_n_key_points
 *     (src/sdav/input/CvInputParser.py:36-46: cv2.xfeatures2d.SURF_create().detect(img), sort by -response, first n).
 *     The fast-Hessian ("SURF") detector of Bay et al. with the documented constants of opencv-contrib 3.4.2 (9x9 base
 *     box filters + 6 per layer, doubling per octave; det = Dxx Dyy - 0.81 Dxy^2; strict 3x3x3 maxima of the middle
 *     layers; quadratic sub-sample interpolation; orientation-window test). That module is absent from the reference
 *     tree and from this image: PARITY UNPINNED against OpenCV; bit-identical to oracle/surf.py.
 * img_dev uint8 [B,H,W] -> xy_dev float32 [B, top_n, 2] (x = column, y = row; what dlc_patch_gather* consume), best
 * response first, ties in detection order (octave, layer, row, column); info_dev (optional) float32 [B, top_n, 2] =
 * (size, response); found_dev int32 [B] = keypoints found in the frame before the cut (entries beyond it are the
 * image centre with response 0; more than 16384 means the candidate list overflowed and the selection is incomplete).
 * SURF_create() defaults: hessian_threshold 100, n_octaves 4, n_layers 3.
 * ------------------------------------------------------------------------------------------------------------ */
size_t dlc_surf_workspace_bytes(int B, int H, int W, int n_octaves, int n_layers);
int dlc_surf_detect(const uint8_t* img_dev, int B, int H, int W, float hessian_threshold, int n_octaves, int n_layers,
                    int top_n, float* xy_dev, float* info_dev, int32_t* found_dev, void* ws_dev, size_t ws_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * a3/a4  SDA encoder forward: H_l = sigmoid(H_{l-1} W_l + b_l).
 *        Replaces SDAV.transform (src/sdav/network/SDAV.py:120-163, 293-302) and DA.transform
 *        (src/sdav/network/DenoisingAutoencoderVariant.py:116-119, 254-259).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct dlc_sda dlc_sda;
/* dims has n_layers+1 entries (1681,2500,2500,2500,2500,2500 for SDAV; in,hidden for one DA). */
/* precision: DLC_PREC_FP16, DLC_PREC_FP16X2_A16, DLC_PREC_FP16X2 or DLC_PREC_AUTO. */
int dlc_sda_create(dlc_sda** h, int n_layers, const int* dims, int precision);
int dlc_sda_destroy(dlc_sda* h);
/* Raw-pixel input mode (call before dlc_sda_set_layer(h, 0, ...); changing it un-sets layer 0): the x planes given to
 * dlc_sda_encode hold pixel values 0..255 (dlc_patch_gather_u8; exact in fp16, x_lo_dev is ignored) and layer 0
 * evaluates sigmoid((p / 255) W + b) as sigmoid((p (W * 256/255)) / 256 + b): two tensor-core products per K step
 * instead of three, and no rounding of the input at all. */
int dlc_sda_set_input_u8(dlc_sda* h, int on);
/* W_host [dims[l], dims[l+1]] row-major float64, b_host [dims[l+1]] float64 (the reference's variable layout). */
int dlc_sda_set_layer(dlc_sda* h, int l, const double* w_host, const double* b_host);
/* bytes of scratch needed to encode `rows` patch rows */
size_t dlc_sda_workspace_bytes(const dlc_sda* h, int rows);
/* DLC_PREC_AUTO handles: choose the arithmetic now, from up to 512 sampled rows of this input (four blocks of 128
 * rows spread over the batch) run through all layers in the three modes. Allocates and frees ~25 MB of device scratch
 * and SYNCHRONISES the stream (it reads two error figures back): an initialisation step, not part of the hot path.
 * dlc_sda_set_layer invalidates the choice. No-op (DLC_OK) on handles created with a fixed precision. */
int dlc_sda_probe(dlc_sda* h, const void* x_hi_dev, const void* x_lo_dev, int rows, void* ws_dev, size_t ws_bytes,
                  void* stream);
/* The arithmetic dlc_sda_encode uses: DLC_PREC_FP16, DLC_PREC_FP16X2_A16 or DLC_PREC_FP16X2 (the creation-time
 * precision, or the probe's choice); -1 for an AUTO handle that has not been probed yet. */
int dlc_sda_chosen_precision(const dlc_sda* h);
/* out_host[2] = max over the sample of |d - d3| / max(1, |d3|) for the one-product and the two-product forward
 * against the three-product one (the figures the choice was made on; zeros before a probe). */
int dlc_sda_probe_stats(const dlc_sda* h, double* out_host);
/* x planes [rows, dlc_plane_ld(dims[0])] (from dlc_patch_gather or dlc_split_planes); out_dev float32
 * [rows, dims[n_layers]] dense (the reference's flat [B*30, 2500] result, SDAV.py:163). */
int dlc_sda_encode(dlc_sda* h, const void* x_hi_dev, const void* x_lo_dev, int rows, float* out_dev, void* ws_dev,
                   size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * a5/a6  SDAV frame-pair similarity matrix.  Replaces SimilarityCalculator.similarity_score and the i<j double
 *        loop around it (src/sdav/similarity/SimilarityCalculator.py:12-49, src/sdav/create_similarity_matrix.py:31-38).
 * desc_dev float32 [N, P, D] descriptors (P <= 32). S_dev float32 [N, N]: S[i][j] = S[j][i] = score(i, j) for i < j
 * (the reference evaluates only i < j and mirrors), diagonal = -1 (the reference's fill value).
 * full_asymmetric = 1 instead evaluates score(i, j) for every ordered pair i != j.
 * The nearest-neighbour distances come from ONE Gram contraction over all patch rows,
 * ||h2_j - h1_k||^2 = n_k + n_j - 2 G_kj, evaluated on h - c: c is the dataset's column mean when the descriptors
 * nearly coincide (sum ||h - mean||^2 < sum ||h||^2 / 16 - a trained-like encoder; without the shift the subtraction
 * cancels below any fp32 accumulator's resolution) and zero otherwise (decided on the device, no host round trip).
 * ------------------------------------------------------------------------------------------------------------ */
size_t dlc_sdav_similarity_workspace_bytes(int N, int P, int D);
/* w_dev: optional float64 [D] distinctive weights (from dlc_sdav_weights on another dataset); NULL = derive them
 * from desc_dev itself, which is what the reference does for the dataset it was constructed with. */
int dlc_sdav_similarity(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a, double b,
                        const double* w_dev, int precision, int full_asymmetric, float* S_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);
/* One part of the same matrix, for a sequence whose score matrix is split over `n_parts` GPUs: part
 * `part` evaluates the 128-row tile rows part, part + n_parts, ... (interleaved, which balances the triangle) against
 * all columns and leaves every entry it does not own at zero, so the parts combine with a sum (NCCL all-reduce);
 * each pair (i, j) and the diagonal fill are produced by exactly one part. */
int dlc_sdav_similarity_part(const float* desc_dev, int N, int P, int D, double mu, double sigma, double a, double b,
                             const double* w_dev, int precision, int full_asymmetric, int part, int n_parts,
                             float* S_dev, void* ws_dev, size_t ws_bytes, void* stream);
/* ---- The same computation in stages, for ONE sequence whose frames are split over the GPUs of a box (SURVEY 8e row 3;
 * workload: the one N x N matrix of src/sdav/create_similarity_matrix.py:31-38). Frames are dealt to the parts (ranks) in
 * contiguous blocks of frames_per_part = ceil(N / n_parts); every stage works on the part's own rows and writes into the
 * part's slice of arrays that the HOST exchanges between stages (NCCL all-gather; gathered slices form the flat global
 * arrays). No work is replicated: mean, row statistics, operand planes and the precision probe are produced once, by
 * the part that holds the frame.
 *   1. dlc_sdav_stage_colsum   local desc -> colsum [2 D] float64 (column sums, then column sums of squares)
 *                                                                          -> all-gather: colsums [n_parts, 2 D]
 *   2. dlc_sdav_stage_weights  colsums -> w [D] float64 and the centring vector of the planes, centre [D] float32
 *                              (the dataset mean when the descriptors nearly coincide, else zero: see
 *                              dlc_sdav_similarity)                        (same result on every part)
 *   3. dlc_sdav_stage_prepare  local desc -> centred fp16 plane slice [frames_per_part * P, dlc_plane_ld(D)] (+ the
 *                              residual plane for DLC_PREC_FP16X2) and a stats block (dlc_sdav_stage_stats_bytes)
 *                                                                          -> all-gather planes, all-gather stats
 *   4. dlc_sdav_stage_gram     planes + stats of all parts -> this part's entries of S (others zero: combine with a
 *                              sum / all-reduce) + the list of frame pairs with ambiguous rows (in ws_dev)
 *   5. dlc_sdav_stage_fix      float32 descriptors of ALL frames [N, P, D] -> exact scores of the listed pairs. Only
 *                              this stage reads them, so their all-gather can overlap stage 4.
 * precision: DLC_PREC_FP16, DLC_PREC_FP16X2, DLC_PREC_FP16_REFINED; DLC_PREC_AUTO means DLC_PREC_FP16_REFINED here.
 * Stages 4 and 5 share ws_dev (dlc_sdav_stage_workspace_bytes, 256-byte aligned). */
size_t dlc_sdav_stage_stats_bytes(int frames_per_part);
size_t dlc_sdav_stage_workspace_bytes(int N, int P, int D, int n_parts);
int dlc_sdav_stage_colsum(const float* desc_local_dev, int64_t rows_local, int D, double* colsum_dev, void* ws_dev,
                          size_t ws_bytes, void* stream);
int dlc_sdav_stage_weights(const double* colsums_dev, int n_parts, int64_t rows_total, int D, double mu, double sigma,
                           double* w_dev, float* centre_dev, void* stream);
int dlc_sdav_stage_prepare(const float* desc_local_dev, int n_local, int frames_per_part, int P, int D,
                           const double* w_dev, const float* centre_dev, int precision, void* plane_hi_local_dev,
                           void* plane_lo_local_dev, void* stats_local_dev, void* stream);
int dlc_sdav_stage_gram(const void* plane_hi_all_dev, const void* plane_lo_all_dev, const void* stats_all_dev,
                        int n_parts, int frames_per_part, int N, int P, int D, double a, double b, int precision,
                        int full_asymmetric, int part, float* S_dev, void* ws_dev, size_t ws_bytes, void* stream);
int dlc_sdav_stage_fix(const void* plane_hi_all_dev, const float* desc_all_dev, int N, int P, int D, double a, double b,
                       int precision, int full_asymmetric, int part, int n_parts, float* S_dev, void* ws_dev,
                       size_t ws_bytes, void* stream);

/* Diagnostics of the last AUTO / FP16_REFINED call on this workspace: out_host[6] = {use_refine, margin, sigma,
 * estimated flagged fraction, flagged rows, refined candidates}. Synchronises the stream. */
int dlc_sdav_similarity_stats(int N, int P, int D, const void* ws_dev, double* out_host, void* stream);
/* w = exp(-(mean_rows(desc) - mu)^2 / (2 sigma^2)), float64 [D] (SimilarityCalculator.py:19-27).
 * ws_dev needs dlc_sdav_similarity_workspace_bytes(N, P, D) bytes (or at least 129*2*D*8). */
int dlc_sdav_weights(const float* desc_dev, int N, int P, int D, double mu, double sigma, double* w_dev,
                     void* ws_dev, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Row-wise candidate selection on a dense score matrix (loop candidates from S, or a k-way merge of partial
 * top-k lists). New capability (north star); nearest reference analogue: np.argmin in
 * src/sdav/similarity/SimilarityCalculator.py:33-35. Order: best score first, ties -> lowest index.
 * cand_idx_dev (optional, int64 [rows, cols]) maps a column to the index reported; NULL reports the column.
 * Columns with |col - row| <= exclude_band are skipped when exclude_band >= 0 (temporal neighbours / diagonal).
 * Rows with fewer than k candidates are padded with idx = -1 and score = -inf (+inf when smallest-first).
 * ------------------------------------------------------------------------------------------------------------ */
int dlc_topk_rows(const float* scores_dev, const int64_t* cand_idx_dev, int rows, int cols, int ld, int k,
                  int largest, int exclude_band, float* out_scores_dev, int64_t* out_idx_dev, void* stream);

/* Frame-level descriptor for the global matcher: mean over the `group_rows` patch descriptors of each frame,
 * x_dev float32 [groups*group_rows, cols] -> out_dev float32 [groups, cols]. New definition (the reference never
 * forms one descriptor per frame); the database L2-normalises it on append when the metric is cosine. */
int dlc_mean_pool_rows(const float* x_dev, int groups, int group_rows, int cols, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Keyframe database + global matcher (cosine / dot / L2 similarity matrix with fused per-row top-k). New capability.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct dlc_db dlc_db;
/* dtype: DLC_F16 or DLC_BF16 storage. Device memory for `capacity_rows` rows is allocated once here. */
int dlc_db_create(dlc_db** db, int dim, int64_t capacity_rows, int metric, int dtype);
int dlc_db_destroy(dlc_db* db);
int64_t dlc_db_size(const dlc_db* db);
int dlc_db_clear(dlc_db* db);
/* rows_dev [n, dim] (DLC_F32, DLC_F16 or DLC_BF16). COS: rows are L2-normalised before storage. */
int dlc_db_append(dlc_db* db, const void* rows_dev, int src_dtype, int64_t n, void* stream);
size_t dlc_match_workspace_bytes(const dlc_db* db, int B, int k);
/* q_dev float32 [B, dim]; out scores [B,k] / idx [B,k] (database row + idx_offset). k <= 32. */
int dlc_match_topk(dlc_db* db, const float* q_dev, int B, int k, int64_t idx_offset, float* scores_dev,
                   int64_t* idx_dev, void* ws_dev, size_t ws_bytes, void* stream);
/* Threshold selection: counts_dev[b] = number of database rows whose score passes `thr` (>= for COS/DOT, <= for L2);
 * the best min(count, max_per_row) of them are listed (max_per_row <= 32), the rest of the row is padded. */
int dlc_match_threshold(dlc_db* db, const float* q_dev, int B, float thr, int max_per_row, int64_t idx_offset,
                        int32_t* counts_dev, float* scores_dev, int64_t* idx_dev, void* ws_dev, size_t ws_bytes,
                        void* stream);

/* ---- Row-sharded database over the GPUs of one box (one process per GPU). A dlc_comm wraps an NCCL
 * communicator: rank 0 obtains a 128-byte unique id (dlc_comm_unique_id), the host distributes it to the other ranks
 * by any channel, every rank calls dlc_comm_create. libnccl.so.2 is resolved with dlopen at the first call (the copy
 * the host process already loaded is reused); libdlc.so has no link-time NCCL dependency. */
typedef struct dlc_comm dlc_comm;
int dlc_comm_unique_id(void* id_out_host /* 128 bytes */);
int dlc_comm_create(dlc_comm** c, const void* id_host /* 128 bytes */, int rank, int world);
int dlc_comm_destroy(dlc_comm* c);
size_t dlc_match_sharded_workspace_bytes(const dlc_db* db, int B, int k, int world);
/* Every rank holds its shard in `db` and calls this with the SAME q_dev [B, dim]: fused similarity + top-k on the
 * shard (indices = local row + idx_offset) -> ONE ncclAllGather of the packed (index, score) lists, B*k*12 bytes per
 * rank, stream-ordered on `stream` -> merge kernel reading the gathered blocks in place (best score first, ties ->
 * lowest global index). Every rank ends with identical scores_dev / idx_dev [B, k]. world * k <= 1024. */
int dlc_match_topk_sharded(dlc_db* db, dlc_comm* comm, const float* q_dev, int B, int k, int64_t idx_offset,
                           float* scores_dev, int64_t* idx_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * a9/a10  cnn_vtl Hamming matrix.  Replaces DistanceCalculator.calculate_distance and the N x N loop around it
 *         (src/cnn_vtl/similarity/DistanceCalculator.py:4-12, src/cnn_vtl/create_distance_matrix.py:31-36).
 * desc_dev int8 [N, M]; D_dev int32 [N, N]. signed_bin_quirk = 1 reproduces the reference exactly
 * (bin() of the signed XOR: popcount of |int8(a ^ b)|); 0 = plain two's-complement popcount.
 * ------------------------------------------------------------------------------------------------------------ */
size_t dlc_hamming_workspace_bytes(int N, int M);
int dlc_hamming_matrix(const int8_t* desc_dev, int N, int M, int signed_bin_quirk, int32_t* D_dev, void* ws_dev,
                       size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * a7/a8  cnn_vtl conv head, fused.  Replaces CnnVtl._define_model + CnnVtl.transform
 *        (src/cnn_vtl/network/cnn_vtl.py:28-128, 130-133): conv1 11x11/4 VALID 96 ReLU -> maxpool 3/2 -> conv2 5x5
 *        SAME 256 ReLU -> maxpool 3/2 -> conv3/conv4 3x3 SAME 384 ReLU -> conv5 3x3 SAME 256 linear; the five conv
 *        outputs flattened (NHWC) and concatenated; per-image (d - min) * 255 / (max - min); int8 cast; column
 *        sub-sampling. One tcgen05 kernel per convolution (conv2..5 are implicit GEMMs over im2col-mode TMA; bias,
 *        ReLU, the min/max and the gather of the kept columns live in the epilogue), so neither an im2col matrix
 *        nor the 546,944-wide float descriptor is ever written to memory.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct dlc_cnnvtl dlc_cnnvtl;
/* H x W = spatial size of the input images (cnn_vtl.py:29 input_shape[1:3]); precision DLC_PREC_FP16X2 or _FP16. */
int dlc_cnnvtl_create(dlc_cnnvtl** h, int H, int W, int precision);
int dlc_cnnvtl_destroy(dlc_cnnvtl* h);
/* layer 0..4 = conv1..conv5. w_host float64 HWIO [kh, kw, cin, cout] (tf.layers.conv2d kernel layout), b_host [cout]. */
int dlc_cnnvtl_set_conv(dlc_cnnvtl* h, int layer, const double* w_host, const double* b_host);
/* Length of the concatenated descriptor before sub-sampling (546,944 for 192x240 inputs). */
int64_t dlc_cnnvtl_descriptor_len(const dlc_cnnvtl* h);
/* The kept columns of cnn_vtl.py:119-128 as strictly increasing indices into the concatenated descriptor. */
int dlc_cnnvtl_set_keep_cols(dlc_cnnvtl* h, const int64_t* keep_cols_host, int M);
size_t dlc_cnnvtl_workspace_bytes(const dlc_cnnvtl* h, int n);
/* x_dev [n, H, W, 3] NHWC (DLC_U8, DLC_F32 or DLC_F64; BGR 0..255 like cv2.imread, no mean subtraction) ->
 * out_dev int8 [n, M]. seg_f32_dev_host: optional host array of 5 device pointers (entries may be NULL); entry l
 * receives conv(l+1)'s float32 NHWC output [n, OH, OW, cout] (diagnostics / parity tests). out_dev may be NULL when
 * only the layer outputs are wanted. ws_dev must be 256-byte aligned. */
int dlc_cnnvtl_forward(dlc_cnnvtl* h, const void* x_dev, int x_dtype, int n, int8_t* out_dev,
                       float* const* seg_f32_dev_host, void* ws_dev, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * a7  cnn_vtl conv head building blocks: the explicit-im2col formulation of the same head (im2col planes +
 *     dlc_gemm_planes, NHWC throughout), kept as an independent cross-check of the fused path.
 * ------------------------------------------------------------------------------------------------------------ */
/* x planes NHWC [N,H,W,C] (row = pixel, ld_in >= C) -> im2col planes [N*OH*OW, ld] with column (kh*KW + kw)*C + c,
 * zero padding of pad_t/pad_l pixels (TF 'SAME' puts the extra pixel at the bottom/right). */
int dlc_im2col_planes(const void* x_hi_dev, const void* x_lo_dev, int N, int H, int W, int C, int ld_in, int KH,
                      int KW, int stride, int pad_t, int pad_l, int OH, int OW, void* out_hi_dev, void* out_lo_dev,
                      int ld, void* stream);
/* 3x3/2 VALID max-pool on NHWC float32 -> planes [N*OH*OW, ld] (cnn_vtl.py:42-45, 58-61). */
int dlc_maxpool_planes(const float* x_dev, int N, int H, int W, int C, int window, int stride, int OH, int OW,
                       void* out_hi_dev, void* out_lo_dev, int ld, void* stream);
/* Per-image min/max over `n_seg` NHWC float32 segments (the concatenated descriptor of cnn_vtl.py:96-111), then
 * q = int8(trunc((d - min) * (255 / (max - min)))) with two's-complement wrap, evaluated only at the kept columns
 * keep_cols_dev[M] (indices into the concatenated descriptor; cnn_vtl.py:113-128). minmax_dev float32 [N,2] scratch. */
int dlc_cnnvtl_quantise(const float* const* seg_ptrs_host, const int64_t* seg_sizes_host, int n_seg, int N,
                        const int64_t* keep_cols_dev, int M, float* minmax_dev, int8_t* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training step of the denoising autoencoders (callers: train.py -> SDAV.fit_dataset,
 * train-sdav.py -> SDA.fit -> DA.fit_dataset). Replaces the TensorFlow graph pieces of SDAV._define_model /
 * _define_loss_for_layer / _define_optimizer (src/sdav/network/SDAV.py:120-186, 223-226) and DA._define_fitting_model /
 * _define_loss / _define_optimizer / _corrupt_tensor (src/sdav/network/DenoisingAutoencoderVariant.py:103-158,
 * 182-202). These are the elementwise / reduction kernels; the contractions between them are dlc_gemm_planes calls
 * (forward, tied-weight decoder, gradient into the hidden layer, gradient into the input, and ONE GEMM for the
 * weight gradient x~^T dzh + dzy^T h with the two contractions concatenated along K). All activations float32
 * [R, C] row-major (R = batch * patches), master weights float64.
 * ------------------------------------------------------------------------------------------------------------ */
/* out = x * keep[r % mask_rows] + add[r % mask_rows] (keep / add optional, [mask_rows, C] float32 0/1 masks:
 * SDAV's shared [P, C] masking noise, TensorflowWrapper.py:34-38, or DA's [R, C] zeros/ones masks) -> float32
 * and/or operand planes [R, ld]. */
int dlc_train_corrupt(const float* x_dev, const float* keep_dev, const float* add_dev, int R, int C, int mask_rows,
                      float* out_f32_dev, void* out_hi_dev, void* out_lo_dev, int ld, void* stream);
/* softmax_cross_entropy_with_logits_v2(labels, logits = y) averaged over the R rows (SDAV.py:172): adds the loss to
 * loss_dev[0]; dzy = dy * y * (1 - y) (gradient at the decoder pre-activation) as float32 and/or planes; dlabel
 * (optional) = d loss / d labels. exact_gradient = 0 (what the reference's optimizer.minimize follows): dy is
 * TensorFlow's registered gradient (softmax(y) - labels) / R, which is not the derivative of the loss when a row's
 * labels do not sum to one (they are patch rows here) [TF1-doc: xent_op.h, nn_grad.py]; != 0: the mathematical
 * derivative (softmax(y) * sum(labels) - labels) / R. */
int dlc_train_xent_grad(const float* y_dev, const float* labels_dev, int R, int C, float* dzy_f32_dev,
                        void* dzy_hi_dev, void* dzy_lo_dev, int ld, float* dlabel_dev, double* loss_dev,
                        int exact_gradient, void* stream);
/* dzh = (dh_rec + dh_up + cs_coef * sign(h - sparse_level) + cc_coef * d(sum_b ||h[b] - h[b+1]||_F)/dh) * h * (1 - h)
 * for h [B*P, C]; dh_rec / dh_up optional. cs_coef = sparse_penalty / count, cc_coef = consecutive_penalty / (B - 1)
 * (SDAV.py:174-183); the two loss terms are added to loss_dev[0]. norms_dev: B - 1 doubles of scratch. */
int dlc_train_hidden_grad(const float* h_dev, const float* dh_rec_dev, const float* dh_up_dev, int B, int P, int C,
                          float sparse_level, double cs_coef, double cc_coef, double* norms_dev, float* dzh_f32_dev,
                          void* dzh_hi_dev, void* dzh_lo_dev, int ld, double* loss_dev, void* stream);
/* out[c] = sum_r a[r, c] (bias gradients), float64. */
int dlc_train_colsum(const float* a_dev, int R, int C, double* out_dev, void* stream);
/* a [R, C] -> transposed planes [C, ld]: plane[c][col_off + r] = a[r][c], zero for R <= r < r_pad. */
int dlc_train_transpose_planes(const float* a_dev, int R, int C, void* hi_dev, void* lo_dev, int ld, int col_off,
                               int r_pad, void* stream);
/* w -= lr * grad (GradientDescentOptimizer, SDAV.py:225); grad float32 or float64. */
int dlc_train_sgd(double* w_dev, const void* grad_dev, int grad_dtype, int64_t n, double lr, void* stream);
/* out = (dx + extra) * keep[r % mask_rows]: gradient through x_l = h_{l-1} * mask_l (+ the label gradient). */
int dlc_train_mask_grad(const float* dx_dev, const float* extra_dev, const float* keep_dev, int R, int C,
                        int mask_rows, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Matrix -> 8-bit image: the tails of src/sdav/create_similarity_matrix.py:41-48 (DLC_IMG_SIMILARITY:
 * 255 * ((M - min) / (max - min)), written as in the reference: move = 0 - min, divide = max + move) and of
 * src/cnn_vtl/create_distance_matrix.py:40-41 (DLC_IMG_DISTANCE: 255 - M / max * 255), followed by cv2.imwrite's
 * float64 -> uint8 conversion (round half to even, clamp to 0..255). m_dev [rows, cols] DLC_F32 (scores) or DLC_I32
 * (Hamming distances); truncate_int != 0 truncates float scores toward zero first, as the reference's int64 matrix
 * (np.full([n, n], -1), :31) does on store. Non-finite scores stay out of the range and saturate (+inf -> 255).
 * Float64 arithmetic in the reference's operation order. out_dev uint8 [rows, cols].
 * ------------------------------------------------------------------------------------------------------------ */
#define DLC_IMG_SIMILARITY 0
#define DLC_IMG_DISTANCE 1
size_t dlc_matrix_image_workspace_bytes(void);
int dlc_matrix_image(const void* m_dev, int dtype, int rows, int cols, int mode, int truncate_int, uint8_t* out_dev,
                     void* ws_dev, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DLC_H_ */

// Internal: row top-k with an optional affine post-transform of the reported score
// (out = row_add[row] + scale * score), shared by dlc_topk_rows and the matcher's partial-list merge.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dlc {
int topk_rows_impl(const float* scores, const int64_t* cand_idx, int rows, int cols, int ld, int k, int largest,
                   int exclude_band, const float* row_add, float scale, float* out_scores, int64_t* out_idx,
                   cudaStream_t stream);
// merge of per-CTA partial top-k lists (cand_idx required, unique non-negative indices, -1 = padding): one CTA per row,
// candidates read once
int merge_partials_impl(const float* scores, const int64_t* cand_idx, int rows, int cols, int ld, int k, int largest,
                        const float* row_add, float scale, float* out_scores, int64_t* out_idx, cudaStream_t stream);
}

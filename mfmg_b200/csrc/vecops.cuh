// vecops.cuh -- BLAS-1 helpers (internal C++ API; all stream-ordered on ctx->stream).
#pragma once
#include "common.cuh"

namespace mfmgb
{
int vec_fill(mfmgb_ctx *ctx, double *v, double value, int64_t n);
int vec_axpy(mfmgb_ctx *ctx, double *y, double a, const double *x, int64_t n);
// result_dev[0] = sum_i a[i]*b[i], summed in a fixed order that depends only on n
int vec_dot_async(mfmgb_ctx *ctx, const double *a, const double *b, int64_t n, double *result_dev);
// number of blocks the deterministic reductions use for a length-n vector
int reduce_blocks(const mfmgb_ctx *ctx, int64_t n);
// final pass: result_dev[j] = sum_b partials[j*stride + b] for j < n_results (one launch)
int reduce_finalize(mfmgb_ctx *ctx, const double *partials, int nblocks, int stride, int n_results,
                    double *result_dev);
} // namespace mfmgb

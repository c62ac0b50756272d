// dense.cuh -- dense coarse solver state (internal).
#pragma once
#include "common.cuh"

struct mfmgb_dense
{
  int64_t n = 0, lda = 0;
  double *inv = nullptr; // packed [n][lda]: strictly lower = L^-1, upper incl. diagonal = U^-1
  int *perm = nullptr;   // composed row permutation: (P b)[i] = b[perm[i]]
  double *work0 = nullptr, *work1 = nullptr;
  int64_t num_swaps = 0;
};

namespace mfmgb
{
int dense_solve_async(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x);
}

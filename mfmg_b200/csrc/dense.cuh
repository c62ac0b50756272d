// dense.cuh -- dense coarse solver state (internal).
#pragma once
#include "common.cuh"

struct mfmgb_dense
{
  int64_t n = 0, lda = 0;
  double *inv = nullptr; // M = U^-1 L^-1 P (= A^-1), row-major [n][lda], padding columns zero
  int *perm = nullptr;   // composed row permutation: (P b)[i] = b[perm[i]]
  double *work0 = nullptr, *work1 = nullptr;
  int64_t num_swaps = 0;
  // multi-GPU: the GEMV is split by rows across the ranks (every rank holds M and the full right-hand side);
  // chunk = this rank's rows_per_rank results, gathered = all ranks' chunks.
  bool distributed = false;
  int nranks = 1, rank = 0;
  int64_t rows_per_rank = 0;
  double *chunk = nullptr, *gathered = nullptr;
};

namespace mfmgb
{
int dense_solve_async(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x);
// switch the solve of D to the row-split multi-GPU form (needs an initialised communicator)
int dense_enable_distributed(mfmgb_ctx *ctx, mfmgb_dense *D);
}

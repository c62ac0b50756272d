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
  // multi-GPU: the two triangular GEMVs are split by rows across the ranks (every rank holds the factors and the
  // full right-hand side); rank r takes h rows from the top and h mirrored rows from the bottom so that both the
  // lower and the upper sweep are balanced.  chunk = this rank's 2h results, gathered = all ranks' chunks.
  bool distributed = false;
  int nranks = 1, rank = 0;
  int64_t half = 0;
  double *chunk = nullptr, *gathered = nullptr;
};

namespace mfmgb
{
int dense_solve_async(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x);
// switch the solve of D to the row-split multi-GPU form (needs an initialised communicator)
int dense_enable_distributed(mfmgb_ctx *ctx, mfmgb_dense *D);
}

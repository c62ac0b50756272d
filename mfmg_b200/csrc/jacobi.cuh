// jacobi.cuh -- Jacobi smoother state (internal).
#pragma once
#include "csr.cuh"

struct mfmgb_jacobi
{
  int64_t n = 0;
  int64_t n_vec = 0; // length of the vectors the operator gathers from (n + ghosts for a partitioned block)
  double omega = 1.;
  double *dinv = nullptr; // 1 / a_ii
  double *tmp = nullptr;  // previous iterate for the in-place entry point
};

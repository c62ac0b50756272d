// jacobi.cuh -- Jacobi smoother state (internal).
#pragma once
#include "csr.cuh"

struct mfmgb_jacobi
{
  int64_t n = 0;
  double omega = 1.;
  double *dinv = nullptr; // 1 / a_ii
  double *tmp = nullptr;  // previous iterate for the in-place entry point
};

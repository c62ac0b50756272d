// mf.cuh -- matrix-free Laplace/diffusion operator state (internal).
#pragma once
#include "csr.cuh"

struct mfmgb_mf
{
  int dim = 3, degree = 1;
  int64_t cells[3] = {1, 1, 1};
  int64_t nodes[3] = {1, 1, 1};
  double h[3] = {1., 1., 1.};
  int64_t n = 0, n_cells = 0;
  int nq = 0;                 // (degree+1)^dim quadrature points per cell
  double *coef = nullptr;     // device [n_cells][nq]
  uint8_t *constr = nullptr;  // device [n]
  // host copies of the 1D tables of this degree: S[q][a], D[q][a], Gauss weights (unit interval)
  double S[9] = {0}, D[9] = {0}, W[3] = {0};
};

namespace mfmgb
{
// y = epilogue(A_mf x), same epilogues as the CSR kernels
int mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &args);
} // namespace mfmgb

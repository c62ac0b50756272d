// mf.cuh -- matrix-free Laplace/diffusion operator state (internal).
#pragma once
#include "csr.cuh"

struct mfmgb_mf
{
  int dim = 3, degree = 1;
  int64_t cells[3] = {1, 1, 1};
  int64_t nodes[3] = {1, 1, 1};
  double h[3] = {1., 1., 1.};
  int64_t n = 0, n_cells = 0; // rows (owned nodes), cells of the local box
  int64_t n_local = 0;        // nodes of the local box = length of the vectors it reads (owned + ghost planes)
  int nq = 0;                 // (degree+1)^dim quadrature points per cell
  double *coef = nullptr;     // device [n_cells][nq]
  uint8_t *constr = nullptr;  // device [n]
  // 3D Q1 fast path (mf_q1.cuh)
  bool q1_cell_constant = false;  // every cell's (cell, q) entries are equal: coef_cell + Kref are used
  double *coef_cell = nullptr;    // device [n_cells]
  double Kref[64] = {0};          // sum_q G[q][a][b]: reference cell matrix with the Jacobian folded in
  // one coefficient for the whole grid and every owned unconstrained node interior to the local box: the factorised
  // 27-point stencil z-sweep (mf_q1_sweep.cuh) serves the operator
  bool q1_stencil = false;
  double q1_const_coef = 0.;
  // the constrained nodes of the local box are exactly its x / y faces plus (bottom_bc / top_bc) its first / last
  // plane: the stencil kernel then computes the flags instead of loading them
  bool q1_arith_flags = false, q1_bottom_bc = false, q1_top_bc = false;
  bool force_generic = false;     // tests: run the generic colour-phase kernel instead
  int q1_tz = 6;                  // owned node planes per CTA of mf_q1_kernel
  uint8_t *brick_flags = nullptr; // device, one byte per CTA brick: does it contain a constrained node?
  // slab layout of the row-partitioned hierarchy: vectors are [owned planes | ghost planes below | ghost planes
  // above]; owned node planes are [own0, own1) of the local box (single GPU: all of them)
  int64_t own0 = 0, own1 = 0;
  // host copies of the 1D tables of this degree: S[q][a], D[q][a], Gauss weights (unit interval)
  double S[9] = {0}, D[9] = {0}, W[3] = {0};
};

namespace mfmgb
{
// y = epilogue(A_mf x), same epilogues as the CSR kernels
int mf_apply(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &args);
// z chunks [zc0, zc1) of a 3D Q1 slab operator; chunk 0 is the only one that reads the ghost plane below, the last one
// the only one that reads the ghost plane above
int mf_apply_chunks(mfmgb_ctx *ctx, const mfmgb_mf *M, const double *x, Epi epi, const EpiArgs &args, int zc0, int zc1);
int mf_num_chunks(const mfmgb_mf *M);
} // namespace mfmgb

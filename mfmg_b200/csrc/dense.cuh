// dense.cuh -- dense coarse solver state (internal).
#pragma once
#include "csr.cuh"
#include "common.cuh"

struct mfmgb_dense
{
  int64_t n = 0, lda = 0;
  double *inv = nullptr; // M = U^-1 L^-1 P (= A^-1), row-major [n][lda], padding columns zero
  int *perm = nullptr;   // composed row permutation: (P b)[i] = b[perm[i]]
  // Factor-and-solve form (getrs: source/cuda/dealii_operator_device_helpers.cu:214) for factorisations whose pivots
  // span more than 12 decades: the explicit inverse of a numerically singular coarse operator (the reference's own
  // gold configuration, 4^3 cells with 2 eigenvectors per agglomerate, has cond(A_c) ~ 1e17) has entries ~1/sigma_min and
  // its product with a consistent right-hand side cancels catastrophically, substitution does not.
  bool substitution = false;
  double *lu = nullptr;  // the packed factors L\U (row-major [n][lda]) when substitution is on
  double pivot_ratio = 1.; // min |u_kk| / max |u_kk|
  bool direct_gemv = false; // always the direct-load GEMV (no shared-memory ring): for solves that run NEXT TO another kernel
  double *work0 = nullptr, *work1 = nullptr;
  int64_t num_swaps = 0;
  // multi-GPU: the GEMV is split by rows across the ranks (every rank holds M and the full right-hand side);
  // chunk = this rank's rows_per_rank results, gathered = all ranks' chunks.
  bool distributed = false;
  int nranks = 1, rank = 0;
  int64_t rows_per_rank = 0;
  double *chunk = nullptr, *gathered = nullptr;
};

struct mfmgb_coarse_dd;

namespace mfmgb
{
int dense_solve_async(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *x);
// switch the solve of D to the row-split multi-GPU form (needs an initialised communicator)
int dense_enable_distributed(mfmgb_ctx *ctx, mfmgb_dense *D);
// domain-decomposed coarse solve of a row-partitioned hierarchy (coarse_dd.cu): on return x_c is valid on this
// rank's coarse rows and on all separator rows
// g_below (may be NULL): this rank's share of the restricted residual on the separator rows of the rank below
// (coarse_dd_n_sep_below values), added to the all-reduced separator right-hand side
int coarse_dd_solve_async(mfmgb_ctx *ctx, const mfmgb_coarse_dd *d, const double *b_c, double *x_c,
                          const double *g_below = nullptr);
int64_t coarse_dd_n_sep_below(const mfmgb_coarse_dd *d);
// building blocks shared with the domain-decomposed coarse solver (coarse_dd.cu)
int csr_to_dense_device(mfmgb_ctx *ctx, const mfmgb_csr *A, int64_t lda, double **out); // n_rows x lda, zero padded
int dense_factor_device(mfmgb_ctx *ctx, double *lu, int64_t n, mfmgb_dense **out);     // takes ownership of lu
int dense_apply_rows(mfmgb_ctx *ctx, const mfmgb_dense *D, const double *b, double *out, int64_t row0, int64_t n_out);
int dense_gemm(mfmgb_ctx *ctx, int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
               int64_t ldb, double *C, int64_t ldc, double alpha, double beta);
}

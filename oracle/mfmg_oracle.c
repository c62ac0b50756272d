/*
 * mfmg_oracle.c -- CPU ORACLE for the mfmg V-cycle-apply hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (mfmg_b200/, include/) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's HOST algorithm for the path; every
 * function cites the reference file:line it follows (paths relative to the mfmg tree).
 * The reference itself cannot be built here (needs deal.II, Trilinos, MPI, Boost, LAPACKE,
 * ARPACK, p4est; its CUDA path needs cuSPARSE entry points removed in CUDA 12), so there is
 * no oracle/_ref.  Parity pins: the known-answer tests of the reference's own test-suite
 * (tests/golden/, see tests/test_oracle_kat.py).  PCG iteration counts / residual
 * histories are "parity unpinned": no reference test asserts a value for them
 * (tests/hierarchy_driver.cc:211-212), so the oracle's own history is the contract.
 *
 * Threading: rows of an SpMV are independent, so "#pragma omp parallel for" over rows
 * does not change any result bit (each row is still summed left to right).  Reductions
 * (dot, norm) are summed serially in index order so that results do not depend on the
 * thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef int32_t i32;

/* ------------------------------------------------------------------------------------------
 * CSR y = A x.  Follows the host operator apply, source/dealii/dealii_trilinos_matrix_operator.cc:28-35
 * (Epetra CrsMatrix::Multiply, row sums left to right in stored column order) and is the
 * semantics the device SparseMatrixDevice::vmult must reproduce
 * (include/mfmg/cuda/sparse_matrix_device.templates.cuh:351-371, tests/test_sparse_matrix_device.cu:98-107).
 * ---------------------------------------------------------------------------------------- */
void orc_spmv(i64 n_rows, const i64 *rowptr, const i32 *col, const double *val,
              const double *x, double *y)
{
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n_rows; ++i)
  {
    double s = 0.;
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
      s += val[k] * x[col[k]];
    y[i] = s;
  }
}

/* y = A^T x with the implicit transpose (Epetra Multiply(TransA=true), used by the host
 * operator for OperatorMode::TRANS, source/dealii/dealii_trilinos_matrix_operator.cc:33-34;
 * the V-cycle prolongation, include/mfmg/common/hierarchy.hpp:297-298).  Serial scatter:
 * contributions to y[c] arrive in ascending row order, which is exactly the order of an
 * explicitly stored, column-sorted transpose (source/cuda/cuda_matrix_operator.cu:93-123). */
void orc_spmv_transpose(i64 n_rows, i64 n_cols, const i64 *rowptr, const i32 *col,
                        const double *val, const double *x, double *y)
{
  for (i64 c = 0; c < n_cols; ++c)
    y[c] = 0.;
  for (i64 i = 0; i < n_rows; ++i)
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
      y[col[k]] += val[k] * x[i];
}

/* Explicit transpose (counting sort, stable in row order => columns of A^T ascending).
 * Restates what CudaMatrixOperator::transpose builds through EpetraExt on the host
 * (source/cuda/cuda_matrix_operator.cu:93-123).  Output arrays are caller-allocated:
 * t_rowptr[n_cols+1], t_col[nnz], t_val[nnz]. */
void orc_csr_transpose(i64 n_rows, i64 n_cols, const i64 *rowptr, const i32 *col,
                       const double *val, i64 *t_rowptr, i32 *t_col, double *t_val)
{
  for (i64 c = 0; c <= n_cols; ++c)
    t_rowptr[c] = 0;
  for (i64 k = 0; k < rowptr[n_rows]; ++k)
    t_rowptr[col[k] + 1]++;
  for (i64 c = 0; c < n_cols; ++c)
    t_rowptr[c + 1] += t_rowptr[c];
  i64 *next = (i64 *)malloc(sizeof(i64) * (size_t)(n_cols > 0 ? n_cols : 1));
  memcpy(next, t_rowptr, sizeof(i64) * (size_t)n_cols);
  for (i64 i = 0; i < n_rows; ++i)
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
    {
      i64 p = next[col[k]]++;
      t_col[p] = (i32)i;
      t_val[p] = val[k];
    }
  free(next);
}

/* Inverse diagonal: D^-1_ii = 1 / a_ii (kernel extract_inv_diag, source/cuda/cuda_smoother.cu:86-96).
 * Returns the number of rows without a stored diagonal entry (dinv left at 0 for those). */
i64 orc_inv_diag(i64 n, const i64 *rowptr, const i32 *col, const double *val, double *dinv)
{
  i64 missing = 0;
  for (i64 i = 0; i < n; ++i)
  {
    int found = 0;
    dinv[i] = 0.;
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i)
      {
        dinv[i] = 1. / val[k];
        found = 1;
      }
    if (!found)
      ++missing;
  }
  return missing;
}

/* r = A x - b : the NEGATIVE residual of include/mfmg/common/hierarchy.hpp:282-286. */
void orc_residual_neg(i64 n, const i64 *rowptr, const i32 *col, const double *val,
                      const double *x, const double *b, double *r)
{
  orc_spmv(n, rowptr, col, val, x, r);
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i)
    r[i] = r[i] - b[i];
}

/* One Jacobi sweep exactly as SmootherOperator::apply, source/cuda/cuda_smoother.cu:49-59
 * (same algebra on the host: source/dealii/dealii_smoother.cc:72-80):
 *     r = A x;  r -= b;  t = D^-1 r;  x -= t.
 * omega is NOT in the reference (it is pinned to 1 by tests/test_smoother_device.cu:71-78);
 * for omega != 1 the update is x -= omega * t.  work has n entries. */
void orc_jacobi_apply(i64 n, const i64 *rowptr, const i32 *col, const double *val,
                      const double *dinv, double omega, const double *b, double *x, double *work)
{
  orc_spmv(n, rowptr, col, val, x, work);
  if (omega == 1.)
  {
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i)
    {
      double r = work[i] - b[i];
      double t = dinv[i] * r;
      x[i] = x[i] - t;
    }
  }
  else
  {
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < n; ++i)
    {
      double r = work[i] - b[i];
      double t = dinv[i] * r;
      x[i] = x[i] - omega * t;
    }
  }
}

/* CSR -> dense row-major (cusparseDcsr2dense at source/cuda/dealii_operator_device_helpers.cu:182,
 * duplicates summed). */
void orc_csr_to_dense(i64 n, const i64 *rowptr, const i32 *col, const double *val, double *dense)
{
  memset(dense, 0, sizeof(double) * (size_t)n * (size_t)n);
  for (i64 i = 0; i < n; ++i)
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
      dense[i * n + col[k]] += val[k];
}

/* Dense LU with partial (row) pivoting, in place, row-major: P A = L U, unit-diagonal L.
 * Restates getrf as used by lu_factorization, source/cuda/dealii_operator_device_helpers.cu:169-228
 * (pivot = first entry of maximal magnitude in the column, as LAPACK idamax).
 * piv[k] = row swapped with k at step k.  Returns 0, or k+1 if U(k,k) == 0. */
int orc_lu_factor(i64 n, double *a, i32 *piv)
{
  int info = 0;
  for (i64 k = 0; k < n; ++k)
  {
    i64 p = k;
    double amax = fabs(a[k * n + k]);
    for (i64 i = k + 1; i < n; ++i)
    {
      double v = fabs(a[i * n + k]);
      if (v > amax)
      {
        amax = v;
        p = i;
      }
    }
    piv[k] = (i32)p;
    if (amax == 0.)
    {
      if (!info)
        info = (int)(k + 1);
      continue;
    }
    if (p != k)
      for (i64 j = 0; j < n; ++j)
      {
        double t = a[k * n + j];
        a[k * n + j] = a[p * n + j];
        a[p * n + j] = t;
      }
    double inv = 1. / a[k * n + k];
#pragma omp parallel for schedule(static) if (n - k > 256)
    for (i64 i = k + 1; i < n; ++i)
    {
      double l = a[i * n + k] * inv;
      a[i * n + k] = l;
      if (l != 0.)
        for (i64 j = k + 1; j < n; ++j)
          a[i * n + j] -= l * a[k * n + j];
    }
  }
  return info;
}

/* x = A^-1 b from the factors (getrs, source/cuda/dealii_operator_device_helpers.cu:214):
 * apply the row interchanges, forward substitution with unit L, back substitution with U. */
void orc_lu_solve(i64 n, const double *lu, const i32 *piv, const double *b, double *x)
{
  if (x != b)
    memcpy(x, b, sizeof(double) * (size_t)n);
  for (i64 k = 0; k < n; ++k)
  {
    i64 p = piv[k];
    if (p != k)
    {
      double t = x[k];
      x[k] = x[p];
      x[p] = t;
    }
  }
  for (i64 i = 0; i < n; ++i)
  {
    double s = x[i];
    const double *row = lu + i * n;
    for (i64 j = 0; j < i; ++j)
      s -= row[j] * x[j];
    x[i] = s;
  }
  for (i64 i = n - 1; i >= 0; --i)
  {
    double s = x[i];
    const double *row = lu + i * n;
    for (i64 j = i + 1; j < n; ++j)
      s -= row[j] * x[j];
    x[i] = s / row[i];
  }
}

/* ---- the same LU with partial pivoting on BAND storage -------------------------------------------------
 * The coarse operators of structured agglomerate grids are banded (lexicographic agglomerates: bandwidth
 * ~ agglomerates per plane), and at n_c = 32 768 (BASELINE configs[3]) the dense array above would be 8.6 GB and its
 * factorisation 2.3e13 flops.  This is the SAME algorithm -- getrf's pivot rule (first entry of maximal magnitude in
 * the column), the same multipliers, the same update order -- restricted to the entries that can be non-zero:
 * with lower / upper bandwidths kl / ku, step k only touches rows (k, k + kl] and columns (k, k + ku + kl].
 * Every skipped operation of the dense routine is x - l * 0 (exact), so the factors and the solutions are
 * bit-identical to orc_lu_factor / orc_lu_solve (checked in tests/test_oracle_kat.py).  As in LAPACK's band
 * routines the multipliers are not interchanged; the solve applies interchange k right before eliminating with
 * column k, which performs the same subtractions in the same order as permuting first.
 * Storage: row i holds columns [i - kl, i + ku + kl] in ab[i * w + (j - i + kl)], w = 2 kl + ku + 1. */
#define BAND(i, j) ab[(size_t)(i) * (size_t)w + (size_t)((j) - (i) + kl)]

void orc_csr_bandwidths(i64 n, const i64 *rowptr, const i32 *col, i64 *kl_out, i64 *ku_out)
{
  i64 kl = 0, ku = 0;
  for (i64 i = 0; i < n; ++i)
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
    {
      const i64 d = (i64)col[k] - i;
      if (-d > kl)
        kl = -d;
      if (d > ku)
        ku = d;
    }
  *kl_out = kl;
  *ku_out = ku;
}

void orc_csr_to_band(i64 n, const i64 *rowptr, const i32 *col, const double *val, i64 kl, i64 ku, double *ab)
{
  const i64 w = 2 * kl + ku + 1;
  memset(ab, 0, sizeof(double) * (size_t)n * (size_t)w);
  for (i64 i = 0; i < n; ++i)
    for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k)
      BAND(i, (i64)col[k]) += val[k];
}

int orc_band_lu_factor(i64 n, i64 kl, i64 ku, double *ab, i32 *piv)
{
  const i64 w = 2 * kl + ku + 1;
  int info = 0;
  for (i64 k = 0; k < n; ++k)
  {
    const i64 i_end = k + kl < n - 1 ? k + kl : n - 1;      /* last row with an entry in column k */
    const i64 j_end = k + ku + kl < n - 1 ? k + ku + kl : n - 1; /* last column of row k after fill-in */
    i64 p = k;
    double amax = fabs(BAND(k, k));
    for (i64 i = k + 1; i <= i_end; ++i)
    {
      double v = fabs(BAND(i, k));
      if (v > amax)
      {
        amax = v;
        p = i;
      }
    }
    piv[k] = (i32)p;
    if (amax == 0.)
    {
      if (!info)
        info = (int)(k + 1);
      continue;
    }
    if (p != k)
      for (i64 j = k; j <= j_end; ++j)
      {
        double t = BAND(k, j);
        BAND(k, j) = BAND(p, j);
        BAND(p, j) = t;
      }
    const double inv = 1. / BAND(k, k);
#pragma omp parallel for schedule(static) if ((i_end - k) * (j_end - k) > 200000)
    for (i64 i = k + 1; i <= i_end; ++i)
    {
      double l = BAND(i, k) * inv;
      BAND(i, k) = l;
      if (l != 0.)
        for (i64 j = k + 1; j <= j_end; ++j)
          BAND(i, j) -= l * BAND(k, j);
    }
  }
  return info;
}

void orc_band_lu_solve(i64 n, i64 kl, i64 ku, const double *ab, const i32 *piv, const double *b, double *x)
{
  const i64 w = 2 * kl + ku + 1;
  if (x != b)
    memcpy(x, b, sizeof(double) * (size_t)n);
  for (i64 k = 0; k < n; ++k)
  {
    const i64 p = piv[k];
    if (p != k)
    {
      double t = x[k];
      x[k] = x[p];
      x[p] = t;
    }
    const i64 i_end = k + kl < n - 1 ? k + kl : n - 1;
    const double xk = x[k];
    for (i64 i = k + 1; i <= i_end; ++i)
      x[i] -= BAND(i, k) * xk;
  }
  for (i64 i = n - 1; i >= 0; --i)
  {
    const i64 j_end = i + ku + kl < n - 1 ? i + ku + kl : n - 1;
    double s = x[i];
    for (i64 j = i + 1; j <= j_end; ++j)
      s -= BAND(i, j) * x[j];
    x[i] = s / BAND(i, i);
  }
}
#undef BAND

/* ------------------------------------------------------------------------------------------
 * Hierarchy: levels[0] finest ... levels[L-1] coarsest.  Mirrors mfmg::Hierarchy / Level
 * (include/mfmg/common/hierarchy.hpp:159-309, level.hpp:22-76).  Operators are given, not
 * built: setup (agglomeration, eigensolves, R assembly, R A R^T) is outside the hot path.
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  i64 n;
  /* A of this level (CSR); on the coarsest level only used to build the dense factors */
  const i64 *a_rowptr;
  const i32 *a_col;
  const double *a_val;
  /* matrix-free fine operator (level 0 only): if mf != NULL, A x is evaluated by it */
  void *mf;
  /* restrictor from the next-finer level to this one: n x n_finer (levels >= 1) */
  i64 r_ncols;
  const i64 *r_rowptr;
  const i32 *r_col;
  const double *r_val;
  /* explicit transpose of the restrictor (built at finalize) */
  i64 *p_rowptr;
  i32 *p_col;
  double *p_val;
  double *dinv;
  /* Chebyshev smoother (dealii::PreconditionChebyshev as used by DealIIMatrixFreeSmoother): theta, delta from the
   * eigenvalue estimate; work vectors */
  double cheb_theta, cheb_delta, cheb_lambda_min, cheb_lambda_max;
  double *cw1, *cw2, *cw3, *cw4;
  /* coarsest level: dense LU, or the same factorisation on band storage (band_kl >= 0) */
  double *lu;
  i32 *piv;
  i64 band_kl, band_ku;
  /* work vectors */
  double *res, *work, *bc, *xc;
} orc_level;

typedef struct
{
  int n_levels;
  int n_smoothing_steps; /* smoother.n_smoothing_steps, hierarchy.hpp:169 */
  int is_preconditioner; /* "is preconditioner", hierarchy.hpp:168 */
  double omega;
  int explicit_transpose; /* 1: prolong with stored R^T; 0: implicit Tvmult */
  int coarse_storage;     /* 0: automatic, 1: dense array, 2: band storage (same arithmetic, see orc_band_lu_factor) */
  /* smoother.type: 0 = Jacobi (source/cuda/cuda_smoother.cu), 1 = Chebyshev (source/dealii/dealii_matrix_free_smoother.cc) */
  int smoother_type;
  int cheb_degree;        /* smoother.degree            (PreconditionChebyshev::AdditionalData::degree, default 0)          */
  double cheb_range;      /* smoother.smoothing_range   (default 0: alpha = min(0.9 lambda_max, lambda_min))                 */
  double cheb_max_ev;     /* smoother.max_eigenvalue    (default 1; only used when cheb_cg_its == 0)                         */
  int cheb_cg_its;        /* eig_cg_n_iterations        (default 8; the reference never changes it)                          */
  int cheb_guess;         /* initial vector of the eigenvalue CG: 0 = (i mod 11) minus its mean, 1 = 1/sqrt(n) with v_0 = 0  */
  orc_level *lev;
} orc_hierarchy;

void orc_mf_apply(const void *mf, const double *x, double *y);
void orc_mf_diag(const void *mf, double *diag);
i64 orc_mf_n(const void *mf);

orc_hierarchy *orc_hierarchy_new(int n_levels, int n_smoothing_steps, int is_preconditioner,
                                 double omega)
{
  orc_hierarchy *h = (orc_hierarchy *)calloc(1, sizeof(orc_hierarchy));
  h->n_levels = n_levels;
  h->n_smoothing_steps = n_smoothing_steps;
  h->is_preconditioner = is_preconditioner;
  h->omega = omega;
  h->explicit_transpose = 1;
  h->lev = (orc_level *)calloc((size_t)n_levels, sizeof(orc_level));
  return h;
}

void orc_hierarchy_set_explicit_transpose(orc_hierarchy *h, int flag) { h->explicit_transpose = flag; }
void orc_hierarchy_set_coarse_storage(orc_hierarchy *h, int mode) { h->coarse_storage = mode; }
void orc_hierarchy_set_chebyshev(orc_hierarchy *h, int degree, double smoothing_range, double max_eigenvalue,
                                 int eig_cg_n_iterations, int initial_guess)
{
  h->smoother_type = 1;
  h->cheb_degree = degree;
  h->cheb_range = smoothing_range;
  h->cheb_max_ev = max_eigenvalue;
  h->cheb_cg_its = eig_cg_n_iterations;
  h->cheb_guess = initial_guess;
}

/* The arrays are borrowed: the caller keeps them alive for the life of the hierarchy. */
void orc_hierarchy_set_operator(orc_hierarchy *h, int level, i64 n, const i64 *rowptr,
                                const i32 *col, const double *val)
{
  orc_level *l = &h->lev[level];
  l->n = n;
  l->a_rowptr = rowptr;
  l->a_col = col;
  l->a_val = val;
}

void orc_hierarchy_set_mf_operator(orc_hierarchy *h, void *mf)
{
  h->lev[0].mf = mf;
  h->lev[0].n = orc_mf_n(mf);
}

void orc_hierarchy_set_restrictor(orc_hierarchy *h, int level, i64 n_rows, i64 n_cols,
                                  const i64 *rowptr, const i32 *col, const double *val)
{
  orc_level *l = &h->lev[level];
  (void)n_rows;
  l->r_ncols = n_cols;
  l->r_rowptr = rowptr;
  l->r_col = col;
  l->r_val = val;
}

static void level_apply_A(const orc_level *l, const double *x, double *y);
double orc_dot(i64 n, const double *a, const double *b);

/* eigenvalues of a symmetric tridiagonal matrix (diagonal d[0..k), off-diagonal e[0..k-1)) by cyclic Jacobi rotations
 * on the dense k x k matrix (k <= 64); ascending on return in w */
static void tridiag_eigenvalues(int k, const double *d, const double *e, double *w)
{
  double a[64 * 64];
  memset(a, 0, sizeof(double) * (size_t)k * (size_t)k);
  for (int i = 0; i < k; ++i)
  {
    a[i * k + i] = d[i];
    if (i + 1 < k)
      a[i * k + i + 1] = a[(i + 1) * k + i] = e[i];
  }
  for (int sweep = 0; sweep < 100; ++sweep)
  {
    double off = 0.;
    for (int i = 0; i < k; ++i)
      for (int j = i + 1; j < k; ++j)
        off += a[i * k + j] * a[i * k + j];
    if (off < 1e-300)
      break;
    for (int p = 0; p < k; ++p)
      for (int q = p + 1; q < k; ++q)
      {
        const double apq = a[p * k + q];
        if (apq == 0.)
          continue;
        const double tau = (a[q * k + q] - a[p * k + p]) / (2. * apq);
        const double t = (tau >= 0. ? 1. : -1.) / (fabs(tau) + sqrt(1. + tau * tau));
        const double c = 1. / sqrt(1. + t * t), sn = t * c;
        for (int r = 0; r < k; ++r)
        {
          const double arp = a[r * k + p], arq = a[r * k + q];
          a[r * k + p] = c * arp - sn * arq;
          a[r * k + q] = sn * arp + c * arq;
        }
        for (int r = 0; r < k; ++r)
        {
          const double apr = a[p * k + r], aqr = a[q * k + r];
          a[p * k + r] = c * apr - sn * aqr;
          a[q * k + r] = sn * apr + c * aqr;
        }
      }
  }
  for (int i = 0; i < k; ++i)
    w[i] = a[i * k + i];
  for (int i = 1; i < k; ++i) /* insertion sort */
  {
    double v = w[i];
    int j = i - 1;
    for (; j >= 0 && w[j] > v; --j)
      w[j + 1] = w[j];
    w[j + 1] = v;
  }
}

/* dealii::PreconditionChebyshev::estimate_eigenvalues (deal.II @89057dff, the version mfmg's ci/Dockerfile pins;
 * third-party, restated from its published algorithm): eig_cg_n_iterations steps of CG preconditioned with D^-1 on
 * A x = v from x = 0, v = the "high-frequency" start vector; the CG coefficients give the Lanczos tridiagonal matrix
 *   T_jj = 1/alpha_j + beta_{j-1}/alpha_{j-1},  T_{j,j+1} = sqrt(beta_j)/alpha_j
 * whose extreme eigenvalues estimate those of D^-1 A; lambda_max gets the safety factor 1.2.  Then
 *   alpha = smoothing_range > 1 ? lambda_max / smoothing_range : min(0.9 lambda_max, lambda_min)
 *   delta = (lambda_max - alpha) / 2,  theta = (lambda_max + alpha) / 2.
 * CG recurrence and stopping rule as SolverCG with ReductionControl(n_its, sqrt(eps), 1e-10). */
static void cheb_estimate(orc_hierarchy *h, orc_level *l)
{
  const i64 n = l->n;
  l->cw1 = (double *)calloc((size_t)n, sizeof(double));
  l->cw2 = (double *)calloc((size_t)n, sizeof(double));
  l->cw3 = (double *)calloc((size_t)n, sizeof(double));
  l->cw4 = (double *)calloc((size_t)n, sizeof(double));
  double lmin = 1., lmax = 1.;
  if (h->cheb_cg_its > 0 && n > 0)
  {
    double *x = l->cw1, *g = l->cw2, *hh = l->cw3, *d = l->cw4;
    double *Ad = (double *)calloc((size_t)n, sizeof(double));
    /* right-hand side = start vector (set_initial_guess) */
    if (h->cheb_guess == 0)
    {
      double mean = 0.;
      for (i64 i = 0; i < n; ++i)
      {
        g[i] = (double)(i % 11);
        mean += g[i];
      }
      mean /= (double)n;
      for (i64 i = 0; i < n; ++i)
        g[i] -= mean;
    }
    else
    {
      for (i64 i = 0; i < n; ++i)
        g[i] = 1. / sqrt((double)n);
      g[0] = 0.;
    }
    /* x = 0: g = A x - b = -b */
    for (i64 i = 0; i < n; ++i)
    {
      x[i] = 0.;
      g[i] = -g[i];
    }
    double res = sqrt(orc_dot(n, g, g));
    const double res0 = res, tol = sqrt(2.220446049250313e-16), reduce = 1e-10;
    double diag[64], offd[64];
    int k = 0;
    if (res > tol)
    {
      for (i64 i = 0; i < n; ++i)
      {
        hh[i] = l->dinv[i] * g[i];
        d[i] = -hh[i];
      }
      double gh = orc_dot(n, g, hh), beta_alpha = 0.;
      const int max_it = h->cheb_cg_its < 64 ? h->cheb_cg_its : 64;
      for (int it = 1; it <= max_it; ++it)
      {
        level_apply_A(l, d, Ad);
        double alpha = orc_dot(n, d, Ad);
        alpha = gh / alpha;
        for (i64 i = 0; i < n; ++i)
        {
          g[i] += alpha * Ad[i];
          x[i] += alpha * d[i];
        }
        res = sqrt(orc_dot(n, g, g));
        for (i64 i = 0; i < n; ++i)
          hh[i] = l->dinv[i] * g[i];
        double beta = gh;
        gh = orc_dot(n, g, hh);
        beta = gh / beta;
        diag[k] = 1. / alpha + beta_alpha;
        beta_alpha = beta / alpha;
        offd[k] = sqrt(beta) / alpha;
        ++k;
        if (res <= tol || res <= reduce * res0)
          break;
        for (i64 i = 0; i < n; ++i)
          d[i] = beta * d[i] - hh[i];
      }
    }
    free(Ad);
    if (k > 0)
    {
      double w[64];
      tridiag_eigenvalues(k, diag, offd, w);
      lmin = w[0];
      lmax = 1.2 * w[k - 1];
    }
  }
  else
  {
    lmax = h->cheb_max_ev;
    lmin = h->cheb_range != 0. ? h->cheb_max_ev / h->cheb_range : h->cheb_max_ev;
  }
  const double alpha = h->cheb_range > 1. ? lmax / h->cheb_range : (0.9 * lmax < lmin ? 0.9 * lmax : lmin);
  l->cheb_lambda_min = lmin;
  l->cheb_lambda_max = lmax;
  l->cheb_delta = (lmax - alpha) * 0.5;
  l->cheb_theta = (lmax + alpha) * 0.5;
}

void orc_hierarchy_chebyshev_info(const orc_hierarchy *h, int level, double *out4)
{
  const orc_level *l = &h->lev[level];
  out4[0] = l->cheb_lambda_min;
  out4[1] = l->cheb_lambda_max;
  out4[2] = l->cheb_theta;
  out4[3] = l->cheb_delta;
}

/* Build smoothers (all but last level) and the coarse dense LU (last level):
 * Hierarchy ctor, include/mfmg/common/hierarchy.hpp:183-234. Returns LU info. */
int orc_hierarchy_finalize(orc_hierarchy *h)
{
  int info = 0;
  for (int li = 0; li < h->n_levels; ++li)
  {
    orc_level *l = &h->lev[li];
    i64 n = l->n;
    l->res = (double *)calloc((size_t)n, sizeof(double));
    l->work = (double *)calloc((size_t)n, sizeof(double));
    l->bc = (double *)calloc((size_t)n, sizeof(double));
    l->xc = (double *)calloc((size_t)n, sizeof(double));
    if (li > 0)
    {
      i64 nnz = l->r_rowptr[n];
      l->p_rowptr = (i64 *)malloc(sizeof(i64) * (size_t)(l->r_ncols + 1));
      l->p_col = (i32 *)malloc(sizeof(i32) * (size_t)(nnz > 0 ? nnz : 1));
      l->p_val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
      orc_csr_transpose(n, l->r_ncols, l->r_rowptr, l->r_col, l->r_val, l->p_rowptr, l->p_col,
                        l->p_val);
    }
    if (li < h->n_levels - 1)
    {
      l->dinv = (double *)malloc(sizeof(double) * (size_t)n);
      if (l->mf)
      {
        orc_mf_diag(l->mf, l->dinv);
        for (i64 i = 0; i < n; ++i)
          l->dinv[i] = 1. / l->dinv[i];
      }
      else
        orc_inv_diag(n, l->a_rowptr, l->a_col, l->a_val, l->dinv);
      if (h->smoother_type == 1)
        cheb_estimate(h, l);
    }
    else
    {
      i64 kl = 0, ku = 0;
      orc_csr_bandwidths(n, l->a_rowptr, l->a_col, &kl, &ku);
      const int band = h->coarse_storage == 2 || (h->coarse_storage == 0 && n >= 1024 && 4 * (2 * kl + ku + 1) <= n);
      l->piv = (i32 *)malloc(sizeof(i32) * (size_t)(n > 0 ? n : 1));
      l->band_kl = -1;
      if (band)
      {
        l->band_kl = kl;
        l->band_ku = ku;
        l->lu = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (size_t)(2 * kl + ku + 1));
        orc_csr_to_band(n, l->a_rowptr, l->a_col, l->a_val, kl, ku, l->lu);
        info = orc_band_lu_factor(n, kl, ku, l->lu, l->piv);
      }
      else
      {
        l->lu = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
        orc_csr_to_dense(n, l->a_rowptr, l->a_col, l->a_val, l->lu);
        info = orc_lu_factor(n, l->lu, l->piv);
      }
    }
  }
  return info;
}

void orc_hierarchy_free(orc_hierarchy *h)
{
  if (!h)
    return;
  for (int li = 0; li < h->n_levels; ++li)
  {
    orc_level *l = &h->lev[li];
    free(l->res);
    free(l->work);
    free(l->bc);
    free(l->xc);
    free(l->p_rowptr);
    free(l->p_col);
    free(l->p_val);
    free(l->dinv);
    free(l->cw1);
    free(l->cw2);
    free(l->cw3);
    free(l->cw4);
    free(l->lu);
    free(l->piv);
  }
  free(h->lev);
  free(h);
}

static void level_apply_A(const orc_level *l, const double *x, double *y)
{
  if (l->mf)
    orc_mf_apply(l->mf, x, y);
  else
    orc_spmv(l->n, l->a_rowptr, l->a_col, l->a_val, x, y);
}

/* dealii::PreconditionChebyshev::vmult (deal.II @89057dff): dst = p_degree(D^-1 A) D^-1 src from a zero start;
 * `degree` further matrix-vector products after the first damped-Jacobi step (degree 0 == Jacobi with omega = 1/theta) */
static void cheb_vmult(const orc_hierarchy *h, orc_level *l, double *dst, const double *src)
{
  const i64 n = l->n;
  double *u1 = l->cw1, *u2 = l->cw2;
  const double theta = l->cheb_theta, delta = l->cheb_delta;
  const double f2 = 1. / theta;
  for (i64 i = 0; i < n; ++i) /* vector_updates(start_zero = true) */
  {
    u1[i] = f2 * src[i];
    dst[i] = l->dinv[i] * u1[i];
    u1[i] = -dst[i];
  }
  if (fabs(delta) < 1e-40)
    return;
  double rhok = delta / theta;
  const double sigma = theta / delta;
  for (int k = 0; k < h->cheb_degree; ++k)
  {
    level_apply_A(l, dst, u2);
    const double rhokp = 1. / (2. * sigma - rhok);
    const double factor1 = rhokp * rhok, factor2 = 2. * rhokp / delta;
    rhok = rhokp;
    for (i64 i = 0; i < n; ++i)
    {
      double t = u2[i] - src[i];
      t = l->dinv[i] * t;
      u1[i] = factor1 * u1[i] + factor2 * t;
      dst[i] -= u1[i];
    }
  }
}

static void level_smooth(const orc_hierarchy *h, orc_level *l, const double *b, double *x)
{
  i64 n = l->n;
  if (h->smoother_type == 1)
  {
    /* DealIIMatrixFreeSmoother::apply, source/dealii/dealii_matrix_free_smoother.cc:67-79:
     * r = A x - b;  tmp = Chebyshev(r);  x -= tmp */
    level_apply_A(l, x, l->cw3);
    for (i64 i = 0; i < n; ++i)
      l->cw3[i] -= b[i];
    cheb_vmult(h, l, l->cw4, l->cw3);
    for (i64 i = 0; i < n; ++i)
      x[i] -= l->cw4[i];
    return;
  }
  /* source/cuda/cuda_smoother.cu:49-59 */
  level_apply_A(l, x, l->work);
  const double om = h->omega;
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i)
  {
    double r = l->work[i] - b[i];
    double t = l->dinv[i] * r;
    x[i] = (om == 1.) ? x[i] - t : x[i] - om * t;
  }
}

/* Hierarchy::apply, include/mfmg/common/hierarchy.hpp:246-309, line by line. */
void orc_hierarchy_apply(orc_hierarchy *h, const double *b, double *x, int level_index)
{
  orc_level *fine = &h->lev[level_index];
  i64 n = fine->n;
  if (level_index > 0 || h->is_preconditioner) /* :253-259 */
    for (i64 i = 0; i < n; ++i)
      x[i] = 0.;

  if (level_index == h->n_levels - 1)
  {
    if (fine->band_kl >= 0)
      orc_band_lu_solve(n, fine->band_kl, fine->band_ku, fine->lu, fine->piv, b, x);
    else
      orc_lu_solve(n, fine->lu, fine->piv, b, x); /* :261-268 -> cuda_solver.cu:496-515 */
    return;
  }
  orc_level *coarse = &h->lev[level_index + 1];
  i64 nc = coarse->n;

  for (int s = 0; s < h->n_smoothing_steps; ++s) /* :277-279 */
    level_smooth(h, fine, b, x);

  level_apply_A(fine, x, fine->res); /* :284-286 */
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i)
    fine->res[i] = fine->res[i] - b[i];

  /* :289-290 b_c = R res */
  orc_spmv(nc, coarse->r_rowptr, coarse->r_col, coarse->r_val, fine->res, coarse->bc);

  /* :293-294 recurse */
  orc_hierarchy_apply(h, coarse->bc, coarse->xc, level_index + 1);

  /* :297-298 x_corr = R^T x_c  (reuse fine->work) */
  if (h->explicit_transpose)
    orc_spmv(n, coarse->p_rowptr, coarse->p_col, coarse->p_val, coarse->xc, fine->work);
  else
    orc_spmv_transpose(nc, n, coarse->r_rowptr, coarse->r_col, coarse->r_val, coarse->xc,
                       fine->work);
    /* :302 x -= x_corr */
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; ++i)
    x[i] = x[i] - fine->work[i];

  for (int s = 0; s < h->n_smoothing_steps; ++s) /* :305-306 */
    level_smooth(h, fine, b, x);
}

/* Hierarchy::vmult(x, b), hierarchy.hpp:238-244 */
void orc_hierarchy_vmult(orc_hierarchy *h, double *x, const double *b)
{
  orc_hierarchy_apply(h, b, x, 0);
}

/* serial, index-ordered reductions (thread-count independent) */
double orc_dot(i64 n, const double *a, const double *b)
{
  double s = 0.;
  for (i64 i = 0; i < n; ++i)
    s += a[i] * b[i];
  return s;
}

/* Preconditioned CG, the recurrence of dealii::SolverCG::solve (deal.II 9.x solver_cg.h, third
 * party, un-vendored) as called at tests/hierarchy_driver.cc:204-210 with the hierarchy as
 * preconditioner; stops when the l2 norm of g = A x - b is <= tol (absolute) or after max_it
 * steps (SolverControl(max_it, tol), hierarchy_driver.cc:202-203).
 *   g = A x - b; res0 = |g|; h = M^-1 g; d = -h; gh = g.h
 *   loop: h = A d; alpha = gh / (d.h); g += alpha h; x += alpha d; res = |g|; check;
 *         h = M^-1 g; beta = gh; gh = g.h; beta = gh / beta; d = beta d - h
 * res_hist must hold max_it + 1 doubles.  Returns the number of iterations (last_step()),
 * negative if not converged.  If hier == NULL the preconditioner is the identity. */
int orc_pcg(orc_hierarchy *hier, i64 n, const i64 *rowptr, const i32 *col, const double *val,
            const double *b, double *x, double tol, int max_it, double *res_hist)
{
  double *g = (double *)malloc(sizeof(double) * (size_t)n);
  double *hh = (double *)malloc(sizeof(double) * (size_t)n);
  double *d = (double *)malloc(sizeof(double) * (size_t)n);
  int it = 0, converged = 0;
  const orc_level *l0 = hier ? &hier->lev[0] : NULL;

  if (l0 && l0->mf)
    orc_mf_apply(l0->mf, x, g);
  else
    orc_spmv(n, rowptr, col, val, x, g);
  for (i64 i = 0; i < n; ++i)
    g[i] = g[i] - b[i];
  double res = sqrt(orc_dot(n, g, g));
  res_hist[0] = res;
  if (res <= tol)
    converged = 1;
  else
  {
    if (hier)
      orc_hierarchy_vmult(hier, hh, g);
    else
      memcpy(hh, g, sizeof(double) * (size_t)n);
    for (i64 i = 0; i < n; ++i)
      d[i] = -hh[i];
    double gh = orc_dot(n, g, hh);
    while (!converged && it < max_it)
    {
      ++it;
      if (l0 && l0->mf)
        orc_mf_apply(l0->mf, d, hh);
      else
        orc_spmv(n, rowptr, col, val, d, hh);
      double alpha = orc_dot(n, d, hh);
      alpha = gh / alpha;
      for (i64 i = 0; i < n; ++i)
        g[i] = g[i] + alpha * hh[i];
      for (i64 i = 0; i < n; ++i)
        x[i] = x[i] + alpha * d[i];
      res = sqrt(orc_dot(n, g, g));
      res_hist[it] = res;
      if (res <= tol)
      {
        converged = 1;
        break;
      }
      if (hier)
        orc_hierarchy_vmult(hier, hh, g);
      else
        memcpy(hh, g, sizeof(double) * (size_t)n);
      double beta = gh;
      gh = orc_dot(n, g, hh);
      beta = gh / beta;
      for (i64 i = 0; i < n; ++i)
        d[i] = beta * d[i] - hh[i];
    }
  }
  free(g);
  free(hh);
  free(d);
  return converged ? it : -it - 1;
}

/* ------------------------------------------------------------------------------------------
 * Matrix-free Laplace/diffusion operator on a uniform Cartesian grid of Q_p hexahedra /
 * quadrilaterals, lexicographic DoF numbering.  Restates LaplaceOperator::local_apply,
 * tests/laplace_matrix_free.hpp:129-156, inside deal.II's MatrixFreeOperators::Base::vmult
 * semantics (constrained DoFs: input treated as 0 in the cell loop, output y_i = x_i), and
 * compute_diagonal, tests/laplace_matrix_free.hpp:75-98,158-199 (constrained entries := 1).
 *   per cell: u_loc = gather(x); grad at the (p+1)^d Gauss points; multiply by coef(cell,q);
 *   integrate against grad(phi_i); scatter-add.
 * ---------------------------------------------------------------------------------------- */
typedef struct
{
  int dim, degree;
  i64 cells[3];
  double h[3];
  i64 nodes[3];
  i64 n;
  int nq1;          /* p+1 */
  double *shape;    /* [nq1][nq1]  phi_a(x_q) on the unit interval */
  double *dshape;   /* [nq1][nq1]  phi_a'(x_q) on the unit interval */
  double *qw;       /* [nq1] Gauss weights on the unit interval */
  const double *coef;          /* [n_cells][nq1^dim] (borrowed) */
  const unsigned char *constr; /* [n] (borrowed) */
} orc_mf;

static void gauss_unit(int nq, double *pts, double *wts)
{
  /* Gauss-Legendre on [0,1] (dealii::QGauss<1>(p+1), tests/laplace_matrix_free.hpp:304) */
  for (int i = 0; i < nq; ++i)
  {
    double z = cos(M_PI * (i + 0.75) / (nq + 0.5)), pp = 0.;
    for (int it = 0; it < 100; ++it)
    {
      double p1 = 1., p2 = 0.;
      for (int j = 0; j < nq; ++j)
      {
        double p3 = p2;
        p2 = p1;
        p1 = ((2. * j + 1.) * z * p2 - j * p3) / (j + 1.);
      }
      pp = nq * (z * p1 - p2) / (z * z - 1.);
      double z1 = z;
      z = z1 - p1 / pp;
      if (fabs(z - z1) < 1e-16)
        break;
    }
    pts[nq - 1 - i] = 0.5 * (z + 1.);
    wts[nq - 1 - i] = 1. / ((1. - z * z) * pp * pp);
  }
}

orc_mf *orc_mf_new(int dim, int degree, const i64 *cells, const double *h, const double *coef,
                   const unsigned char *constrained)
{
  orc_mf *m = (orc_mf *)calloc(1, sizeof(orc_mf));
  m->dim = dim;
  m->degree = degree;
  m->n = 1;
  for (int d = 0; d < 3; ++d)
  {
    m->cells[d] = d < dim ? cells[d] : 1;
    m->h[d] = d < dim ? h[d] : 1.;
    m->nodes[d] = d < dim ? cells[d] * degree + 1 : 1;
    m->n *= m->nodes[d];
  }
  int nq = degree + 1;
  m->nq1 = nq;
  m->shape = (double *)malloc(sizeof(double) * nq * nq);
  m->dshape = (double *)malloc(sizeof(double) * nq * nq);
  m->qw = (double *)malloc(sizeof(double) * nq);
  double *qp = (double *)malloc(sizeof(double) * nq);
  gauss_unit(nq, qp, m->qw);
  /* Lagrange basis on equidistant support points (FE_Q(p) for p <= 2 uses equidistant
   * Gauss-Lobatto points), a = node index, q = quadrature index. */
  for (int q = 0; q < nq; ++q)
    for (int a = 0; a < nq; ++a)
    {
      double xa = (double)a / degree, v = 1., dv = 0.;
      for (int c = 0; c < nq; ++c)
        if (c != a)
          v *= (qp[q] - (double)c / degree) / (xa - (double)c / degree);
      for (int e = 0; e < nq; ++e)
        if (e != a)
        {
          double t = 1. / (xa - (double)e / degree);
          for (int c = 0; c < nq; ++c)
            if (c != a && c != e)
              t *= (qp[q] - (double)c / degree) / (xa - (double)c / degree);
          dv += t;
        }
      m->shape[q * nq + a] = v;
      m->dshape[q * nq + a] = dv;
    }
  free(qp);
  m->coef = coef;
  m->constr = constrained;
  return m;
}

void orc_mf_free(orc_mf *m)
{
  if (!m)
    return;
  free(m->shape);
  free(m->dshape);
  free(m->qw);
  free(m);
}

i64 orc_mf_n(const void *mf) { return ((const orc_mf *)mf)->n; }

/* cell kernel: out_loc = K_cell(coef) u_loc  (dense evaluation, no sum factorisation) */
static void mf_cell(const orc_mf *m, const double *cq, const double *u, double *out)
{
  const int nq = m->nq1, dim = m->dim;
  const int n3 = dim == 3 ? nq : 1;
  const int ndof = nq * nq * n3;
  for (int i = 0; i < ndof; ++i)
    out[i] = 0.;
  for (int qz = 0; qz < n3; ++qz)
    for (int qy = 0; qy < nq; ++qy)
      for (int qx = 0; qx < nq; ++qx)
      {
        const int q = qx + nq * (qy + nq * qz);
        double g[3] = {0., 0., 0.};
        for (int az = 0; az < n3; ++az)
          for (int ay = 0; ay < nq; ++ay)
            for (int ax = 0; ax < nq; ++ax)
            {
              const double uu = u[ax + nq * (ay + nq * az)];
              const double sx = m->shape[qx * nq + ax], sy = m->shape[qy * nq + ay];
              const double sz = dim == 3 ? m->shape[qz * nq + az] : 1.;
              g[0] += uu * m->dshape[qx * nq + ax] / m->h[0] * sy * sz;
              g[1] += uu * sx * m->dshape[qy * nq + ay] / m->h[1] * sz;
              if (dim == 3)
                g[2] += uu * sx * sy * m->dshape[qz * nq + az] / m->h[2];
            }
        double jxw = m->qw[qx] * m->qw[qy] * m->h[0] * m->h[1];
        if (dim == 3)
          jxw *= m->qw[qz] * m->h[2];
        const double c = cq[q] * jxw;
        for (int az = 0; az < n3; ++az)
          for (int ay = 0; ay < nq; ++ay)
            for (int ax = 0; ax < nq; ++ax)
            {
              const double sx = m->shape[qx * nq + ax], sy = m->shape[qy * nq + ay];
              const double sz = dim == 3 ? m->shape[qz * nq + az] : 1.;
              double v = g[0] * m->dshape[qx * nq + ax] / m->h[0] * sy * sz +
                         g[1] * sx * m->dshape[qy * nq + ay] / m->h[1] * sz;
              if (dim == 3)
                v += g[2] * sx * sy * m->dshape[qz * nq + az] / m->h[2];
              out[ax + nq * (ay + nq * az)] += c * v;
            }
      }
}

static void mf_loop(const orc_mf *m, const double *x, double *y, int diag_mode)
{
  const int nq = m->nq1, p = m->degree, dim = m->dim;
  const int n3 = dim == 3 ? nq : 1;
  const int ndof = nq * nq * n3;
  const int nqp = ndof;
  double u[125], out[125], dg[125];
  for (i64 i = 0; i < m->n; ++i)
    y[i] = 0.;
  for (i64 cz = 0; cz < m->cells[2]; ++cz)
    for (i64 cy = 0; cy < m->cells[1]; ++cy)
      for (i64 cx = 0; cx < m->cells[0]; ++cx)
      {
        const i64 cell = cx + m->cells[0] * (cy + m->cells[1] * cz);
        const double *cq = m->coef + cell * nqp;
        i64 idx[125];
        for (int az = 0; az < n3; ++az)
          for (int ay = 0; ay < nq; ++ay)
            for (int ax = 0; ax < nq; ++ax)
              idx[ax + nq * (ay + nq * az)] =
                  (cx * p + ax) + m->nodes[0] * ((cy * p + ay) + m->nodes[1] * (cz * p + az));
        if (!diag_mode)
        {
          for (int a = 0; a < ndof; ++a)
            u[a] = m->constr[idx[a]] ? 0. : x[idx[a]];
          mf_cell(m, cq, u, out);
          for (int a = 0; a < ndof; ++a)
            if (!m->constr[idx[a]])
              y[idx[a]] += out[a];
        }
        else
        {
          for (int i = 0; i < ndof; ++i)
          {
            for (int a = 0; a < ndof; ++a)
              u[a] = 0.;
            u[i] = 1.;
            mf_cell(m, cq, u, out);
            dg[i] = out[i];
          }
          for (int a = 0; a < ndof; ++a)
            if (!m->constr[idx[a]])
              y[idx[a]] += dg[a];
        }
      }
  for (i64 i = 0; i < m->n; ++i)
    if (m->constr[i])
      y[i] = diag_mode ? 1. : x[i];
}

void orc_mf_apply(const void *mf, const double *x, double *y)
{
  mf_loop((const orc_mf *)mf, x, y, 0);
}
void orc_mf_diag(const void *mf, double *diag) { mf_loop((const orc_mf *)mf, NULL, diag, 1); }

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

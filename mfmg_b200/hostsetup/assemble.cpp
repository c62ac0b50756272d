// assemble.cpp -- HOST SETUP (not the hot path): structured-grid Q_p Laplace/diffusion assembly.
//
// Produces the CSR operator that mfmg's user code hands to the hierarchy through
// MeshEvaluator::evaluate_global / evaluate_agglomerate.  It follows the test problem of the
// reference, tests/laplace.hpp:154-204 (cell matrix K_ij = sum_q kappa(x_q) grad phi_i . grad phi_j JxW,
// constraints applied while assembling), on the mesh the benchmark configs use: a uniform
// Cartesian grid of the unit cube (GridGenerator::hyper_cube + refine_global, tests/laplace.hpp:95-97),
// with DoFs numbered lexicographically instead of in deal.II's cell-traversal order (a pure
// permutation).
//
// The matrix is generated row by row (no COO intermediate), in ascending column order, with
// the full tensor-product stencil kept in the pattern -- deal.II keeps constrained rows and
// columns in the sparsity pattern with explicit zeros (tests/laplace.hpp:146-147), which is
// what makes nnz = (3N-2)^d for Q1.
//
// Built with plain g++ -fopenmp into libmfmg_b200_host.so; no CUDA here.
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace
{
struct Grid
{
  int dim, p;
  int64_t cells[3];
  int64_t nodes[3];
};

inline Grid make_grid(int dim, int degree, const int64_t *cells)
{
  Grid g;
  g.dim = dim;
  g.p = degree;
  for (int d = 0; d < 3; ++d)
  {
    g.cells[d] = d < dim ? cells[d] : 1;
    g.nodes[d] = d < dim ? cells[d] * degree + 1 : 1;
  }
  return g;
}

// 1D column range [lo, hi] (inclusive) coupled to node i, and the adjacent 1D cells [c0, c1]
inline void range_1d(const Grid &g, int d, int64_t i, int64_t &lo, int64_t &hi, int64_t &c0,
                     int64_t &c1)
{
  if (d >= g.dim)
  {
    lo = hi = 0;
    c0 = c1 = 0;
    return;
  }
  const int p = g.p;
  if (i % p == 0)
  {
    const int64_t v = i / p;
    c0 = v > 0 ? v - 1 : 0;
    c1 = v < g.cells[d] ? v : g.cells[d] - 1;
  }
  else
    c0 = c1 = i / p;
  lo = c0 * p;
  hi = c1 * p + p;
}
} // namespace

extern "C"
{
  // launchers such as torchrun export OMP_NUM_THREADS=1; the setup path sets its own thread count
  void hs_set_num_threads(int n)
  {
#ifdef _OPENMP
    if (n > 0)
      omp_set_num_threads(n);
#else
    (void)n;
#endif
  }
  int hs_get_max_threads(void)
  {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
  }

  // Row offsets of the rows [row_begin, row_end) of the global matrix (lexicographic numbering).
  // rowptr has row_end - row_begin + 1 entries and starts at 0.  Returns nnz of those rows.
  int64_t hs_assemble_count(int dim, int degree, const int64_t *cells, int64_t row_begin,
                            int64_t row_end, int64_t *rowptr)
  {
    const Grid g = make_grid(dim, degree, cells);
    const int64_t nloc = row_end - row_begin;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nloc; ++r)
    {
      const int64_t row = row_begin + r;
      const int64_t ix = row % g.nodes[0], iy = (row / g.nodes[0]) % g.nodes[1],
                    iz = row / (g.nodes[0] * g.nodes[1]);
      const int64_t idx[3] = {ix, iy, iz};
      int64_t len = 1;
      for (int d = 0; d < 3; ++d)
      {
        int64_t lo, hi, c0, c1;
        range_1d(g, d, idx[d], lo, hi, c0, c1);
        len *= hi - lo + 1;
      }
      rowptr[r + 1] = len;
    }
    rowptr[0] = 0;
    for (int64_t r = 0; r < nloc; ++r)
      rowptr[r + 1] += rowptr[r];
    return rowptr[nloc];
  }

  // Fill col/val (and optionally diag, may be NULL) for rows [row_begin, row_end).
  //   G            [nq][ndof][ndof]  per-quadrature-point reference matrices
  //                G_q[a][b] = (grad phi_a . grad phi_b)(x_q) * JxW_q  (lexicographic local DoFs)
  //   coef         [n_cells][coef_stride]; coef_stride == nq (value per quadrature point) or 1
  //                (cell-wise constant; then Gsum = sum_q G_q is used)
  //   constrained  [n_nodes] 1 = homogeneous Dirichlet (deal.II AffineConstraints semantics in
  //                distribute_local_to_global: no off-diagonal contributions in constrained rows
  //                and columns, the diagonal of a constrained DoF accumulates the local K_ii)
  //   col          global column indices (int32)
  int hs_assemble_fill(int dim, int degree, const int64_t *cells, int nq, const double *G,
                       const double *coef, int coef_stride, const unsigned char *constrained,
                       int64_t row_begin, int64_t row_end, const int64_t *rowptr, int32_t *col,
                       double *val, double *diag)
  {
    const Grid g = make_grid(dim, degree, cells);
    const int p = degree;
    const int n1 = p + 1;
    const int ndof = dim == 3 ? n1 * n1 * n1 : (dim == 2 ? n1 * n1 : n1);
    if (ndof > 125)
      return 1;
    std::vector<double> gsum((size_t)ndof * ndof, 0.);
    for (int q = 0; q < nq; ++q)
      for (int e = 0; e < ndof * ndof; ++e)
        gsum[e] += G[(size_t)q * ndof * ndof + e];

    const int64_t nloc = row_end - row_begin;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < nloc; ++r)
    {
      const int64_t row = row_begin + r;
      const int64_t idx[3] = {row % g.nodes[0], (row / g.nodes[0]) % g.nodes[1],
                              row / (g.nodes[0] * g.nodes[1])};
      int64_t lo[3], hi[3], c0[3], c1[3], len[3];
      for (int d = 0; d < 3; ++d)
      {
        range_1d(g, d, idx[d], lo[d], hi[d], c0[d], c1[d]);
        len[d] = hi[d] - lo[d] + 1;
      }
      double acc[125];
      const int nacc = (int)(len[0] * len[1] * len[2]);
      for (int e = 0; e < nacc; ++e)
        acc[e] = 0.;
      const bool row_constrained = constrained[row] != 0;
      double krow[125];
      for (int64_t cz = c0[2]; cz <= c1[2]; ++cz)
        for (int64_t cy = c0[1]; cy <= c1[1]; ++cy)
          for (int64_t cx = c0[0]; cx <= c1[0]; ++cx)
          {
            const int64_t cell = cx + g.cells[0] * (cy + g.cells[1] * cz);
            const int ax = (int)(idx[0] - cx * p), ay = dim > 1 ? (int)(idx[1] - cy * p) : 0,
                      az = dim > 2 ? (int)(idx[2] - cz * p) : 0;
            const int a = ax + n1 * (ay + n1 * az);
            if (coef_stride == 1)
            {
              const double c = coef[cell];
              for (int b = 0; b < ndof; ++b)
                krow[b] = c * gsum[(size_t)a * ndof + b];
            }
            else
            {
              for (int b = 0; b < ndof; ++b)
                krow[b] = 0.;
              for (int q = 0; q < nq; ++q)
              {
                const double c = coef[cell * (int64_t)coef_stride + q];
                const double *gq = G + ((size_t)q * ndof + a) * ndof;
                for (int b = 0; b < ndof; ++b)
                  krow[b] += c * gq[b];
              }
            }
            const int nz = dim > 2 ? n1 : 1, ny = dim > 1 ? n1 : 1;
            for (int bz = 0; bz < nz; ++bz)
              for (int by = 0; by < ny; ++by)
                for (int bx = 0; bx < n1; ++bx)
                {
                  const int b = bx + n1 * (by + n1 * bz);
                  const int64_t jx = cx * p + bx, jy = dim > 1 ? cy * p + by : 0,
                                jz = dim > 2 ? cz * p + bz : 0;
                  const int e =
                      (int)((jx - lo[0]) + len[0] * ((jy - lo[1]) + len[1] * (jz - lo[2])));
                  if (row_constrained)
                  {
                    if (b == a)
                      acc[e] += std::fabs(krow[b]);
                  }
                  else
                  {
                    const int64_t j = jx + g.nodes[0] * (jy + g.nodes[1] * jz);
                    if (!constrained[j])
                      acc[e] += krow[b];
                  }
                }
          }
      int64_t k = rowptr[r];
      for (int64_t jz = lo[2]; jz <= hi[2]; ++jz)
        for (int64_t jy = lo[1]; jy <= hi[1]; ++jy)
          for (int64_t jx = lo[0]; jx <= hi[0]; ++jx)
          {
            const int64_t j = jx + g.nodes[0] * (jy + g.nodes[1] * jz);
            const int e = (int)((jx - lo[0]) + len[0] * ((jy - lo[1]) + len[1] * (jz - lo[2])));
            col[k] = (int32_t)j;
            val[k] = acc[e];
            if (j == row && diag)
              diag[r] = acc[e];
            ++k;
          }
    }
    return 0;
  }

  // Galerkin triple product pieces used in setup are done with scipy; this helper only
  // densifies a CSR (rows sorted or not, duplicates summed) into row-major storage.
  void hs_csr_to_dense(int64_t n_rows, int64_t n_cols, const int64_t *rowptr, const int32_t *col,
                       const double *val, double *dense)
  {
    std::memset(dense, 0, sizeof(double) * (size_t)n_rows * (size_t)n_cols);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_rows; ++i)
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
        dense[i * n_cols + col[k]] += val[k];
  }
}

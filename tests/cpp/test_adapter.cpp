// test_adapter.cpp -- the reference's device unit tests, restated against the C++ adapter (include/mfmg_b200/mfmg.hpp):
//   tests/test_smoother_device.cu:28-119          tridiag(-1,4,-1), b = 1, x0 = 0 -> x = 0.25
//   tests/test_direct_solver_device.cu:23-110     x_ref ~ N(10,2), b = A x_ref, all solver names within 1e-12 %
//   tests/test_sparse_matrix_device_operator.cu   30x39 banded matrix, apply / transpose, exact
//   Hierarchy::apply (fused) == the same algorithm composed from the abstract objects == dense host algebra
// Build: g++ -std=c++17 -Iinclude tests/cpp/test_adapter.cpp -Lmfmg_b200/csrc -lmfmg_b200 (+rpath).  Needs a GPU to run.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>

#include "mfmg_b200/mfmg.hpp"

using V = mfmg::DeviceVector;

// the reference's own constructor argument list (sparse_matrix_device.cuh:37-47) is accepted as written
#include <type_traits>
static_assert(std::is_constructible<mfmg::SparseMatrixDevice<double>, mfmg::MPI_Comm, double *, int *, int *, unsigned int,
                                    mfmg::IndexSet const &, mfmg::IndexSet const &,
                                    std::shared_ptr<mfmg::CudaHandle const>>::value,
              "SparseMatrixDevice(MPI_Comm, val, col, rowptr, local_nnz, range, domain, handle)");

static int failures = 0;
#define CHECK(cond)                                                                                 \
  do                                                                                                \
  {                                                                                                 \
    if (!(cond))                                                                                    \
    {                                                                                               \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                 \
      ++failures;                                                                                   \
    }                                                                                               \
  } while (0)

struct HostCsr
{
  unsigned n_rows, n_cols;
  std::vector<int64_t> rp;
  std::vector<int> col;
  std::vector<double> val;
};

static HostCsr tridiag(unsigned size)
{
  HostCsr a{size, size, {0}, {}, {}};
  for (unsigned i = 0; i < size; ++i)
  {
    for (unsigned j = (i == 0 ? 0 : i - 1); j < std::min(size, i + 2); ++j)
    {
      a.col.push_back((int)j);
      a.val.push_back(i == j ? 4. : -1.);
    }
    a.rp.push_back((int64_t)a.col.size());
  }
  return a;
}

static std::vector<double> host_spmv(HostCsr const &a, std::vector<double> const &x)
{
  std::vector<double> y(a.n_rows, 0.);
  for (unsigned i = 0; i < a.n_rows; ++i)
    for (int64_t k = a.rp[i]; k < a.rp[i + 1]; ++k)
      y[i] += a.val[k] * x[a.col[k]];
  return y;
}

int main()
{
  auto handle = std::make_shared<mfmg::CudaHandle>(0);
  auto params = std::make_shared<mfmg::ParameterTree>();

  // ---- smoother ----
  {
    auto a = tridiag(30);
    auto m = std::make_shared<mfmg::SparseMatrixDevice<double>>(handle, a.n_rows, a.n_cols, a.rp, a.col, a.val);
    std::shared_ptr<mfmg::Operator<V>> op = std::make_shared<mfmg::CudaMatrixOperator<V>>(m);
    mfmg::CudaSmoother<V> smoother(op, params);
    auto b = op->build_domain_vector();
    auto x = op->build_range_vector();
    *b = 1.;
    *x = 0.;
    smoother.apply(*b, *x);
    for (double v : x->export_to_host())
      CHECK(v == 0.25);
    auto bad = std::make_shared<mfmg::ParameterTree>();
    bad->put("smoother.type", "Gauss-Seidel");
    bool threw = false;
    try
    {
      mfmg::CudaSmoother<V> s2(op, bad);
    }
    catch (std::runtime_error const &)
    {
      threw = true;
    }
    CHECK(threw);
  }

  // ---- the reference's own constructor: cuda_malloc'ed arrays WITHOUT slack, ownership taken
  //      (tests/test_sparse_matrix_device.cu:59-75, sparse_matrix_device.templates.cuh:244-272) ----
  for (unsigned size : {30u, 5000u, 70000u})
  {
    auto a = tridiag(size);
    std::vector<int> rp32(a.rp.begin(), a.rp.end());
    double *val_dev = nullptr;
    int *col_dev = nullptr, *rp_dev = nullptr;
    mfmg::cuda_malloc(val_dev, (unsigned)a.val.size());
    mfmg::cuda_malloc(col_dev, (unsigned)a.col.size());
    mfmg::cuda_malloc(rp_dev, (unsigned)rp32.size());
    mfmg::cuda_mem_copy_to_dev(a.val, val_dev);
    mfmg::cuda_mem_copy_to_dev(a.col, col_dev);
    mfmg::cuda_mem_copy_to_dev(rp32, rp_dev);
    auto m = std::make_shared<mfmg::SparseMatrixDevice<double>>(handle, val_dev, col_dev, rp_dev,
                                                                (unsigned)a.val.size(), size, size);
    CHECK(m->m() == size && m->n() == size && m->local_nnz() == a.val.size());
    CHECK(m->val_dev == val_dev && m->column_index_dev == col_dev && m->row_ptr_dev == rp_dev);
    std::vector<double> xh(size);
    std::default_random_engine gen(3);
    std::uniform_real_distribution<double> dist(-1., 1.);
    for (auto &v : xh)
      v = dist(gen);
    V x(handle, size), y(handle, size);
    x.import_from_host(xh);
    m->vmult(y, x);
    auto ref = host_spmv(a, xh);
    auto got = y.export_to_host();
    for (unsigned i = 0; i < size; ++i)
      CHECK(std::abs(got[i] - ref[i]) <= 1e-14 * (std::abs(ref[i]) + 1.));
    // smoother on the adopted matrix: one sweep from x = 0 with b = 1 gives D^-1 b = 0.25
    std::shared_ptr<mfmg::Operator<V>> op = std::make_shared<mfmg::CudaMatrixOperator<V>>(m);
    mfmg::CudaSmoother<V> smoother(op, params);
    V b(handle, size);
    b = 1.;
    y = 0.;
    smoother.apply(b, y);
    for (double v : y.export_to_host())
      CHECK(v == 0.25);
  } // (the destructor frees the three adopted arrays)

  // ---- SparseMatrixDevice::mmult (tests/test_sparse_matrix_device.cu:242-361): 30 x 30, five random columns per row
  //      from std::default_random_engine(i) plus the diagonal, A(i,j) = i + j, B(i,j) = i - j, C = A B against the
  //      dense product ----
  {
    unsigned const size = 30;
    HostCsr a{size, size, {0}, {}, {}}, b{size, size, {0}, {}, {}};
    std::vector<std::vector<double>> Ad(size, std::vector<double>(size, 0.)), Bd = Ad;
    for (unsigned i = 0; i < size; ++i)
    {
      std::vector<unsigned> indices;
      std::default_random_engine generator(i);
      std::uniform_int_distribution<int> distribution(0, size - 1);
      for (unsigned j = 0; j < 5; ++j)
        indices.push_back((unsigned)distribution(generator));
      indices.push_back(i);
      std::sort(indices.begin(), indices.end());
      indices.erase(std::unique(indices.begin(), indices.end()), indices.end());
      for (unsigned j : indices)
      {
        a.col.push_back((int)j);
        a.val.push_back((double)(i + j));
        b.col.push_back((int)j);
        b.val.push_back((double)i - (double)j);
        Ad[i][j] = (double)(i + j);
        Bd[i][j] = (double)i - (double)j;
      }
      a.rp.push_back((int64_t)a.col.size());
      b.rp.push_back((int64_t)b.col.size());
    }
    mfmg::SparseMatrixDevice<double> A_dev(handle, size, size, a.rp, a.col, a.val);
    mfmg::SparseMatrixDevice<double> B_dev(handle, size, size, b.rp, b.col, b.val);
    mfmg::SparseMatrixDevice<double> C_dev(handle, size, size, b.rp, b.col, b.val); // overwritten, like the reference's
    A_dev.mmult(C_dev, B_dev);
    CHECK(C_dev.m() == size && C_dev.n() == size);
    std::vector<int64_t> crp;
    std::vector<int> cc;
    std::vector<double> cv;
    C_dev.copy_to_host(crp, cc, cv);
    std::vector<std::vector<double>> Cd(size, std::vector<double>(size, 0.)), Cgot = Cd;
    for (unsigned i = 0; i < size; ++i)
      for (unsigned k = 0; k < size; ++k)
        for (unsigned j = 0; j < size; ++j)
          Cd[i][j] += Ad[i][k] * Bd[k][j];
    for (unsigned i = 0; i < size; ++i)
      for (int64_t k = crp[i]; k < crp[i + 1]; ++k)
      {
        CHECK(k == crp[i] || cc[(std::size_t)k - 1] < cc[(std::size_t)k]); // ascending columns, no duplicates
        Cgot[i][(unsigned)cc[(std::size_t)k]] = cv[(std::size_t)k];
      }
    for (unsigned i = 0; i < size; ++i)
      for (unsigned j = 0; j < size; ++j)
        CHECK(std::abs(Cgot[i][j] - Cd[i][j]) <= 1e-12 * (std::abs(Cd[i][j]) + 1.));
    // and the product acts like A (B x)
    std::vector<double> xh(size);
    for (unsigned i = 0; i < size; ++i)
      xh[i] = 1. + 0.1 * i;
    V x(handle, size), t(handle, size), y1(handle, size), y2(handle, size);
    x.import_from_host(xh);
    B_dev.vmult(t, x);
    A_dev.vmult(y1, t);
    C_dev.vmult(y2, x);
    auto h1 = y1.export_to_host(), h2 = y2.export_to_host();
    for (unsigned i = 0; i < size; ++i)
      CHECK(std::abs(h1[i] - h2[i]) <= 1e-10 * (std::abs(h1[i]) + 1.));
  }

  // ---- direct solver ----
  {
    auto a = tridiag(30);
    auto m = std::make_shared<mfmg::SparseMatrixDevice<double>>(handle, a.n_rows, a.n_cols, a.rp, a.col, a.val);
    std::shared_ptr<mfmg::Operator<V>> op = std::make_shared<mfmg::CudaMatrixOperator<V>>(m);
    std::vector<double> sol_ref(30);
    std::default_random_engine generator;
    std::normal_distribution<> distribution(10., 2.);
    for (auto &v : sol_ref)
      v = distribution(generator);
    auto rhs = host_spmv(a, sol_ref);
    for (auto solver : {"cholesky", "lu_dense", "lu_sparse_host"})
    {
      auto p = std::make_shared<mfmg::ParameterTree>();
      p->put("solver.type", solver);
      mfmg::CudaSolver<V> direct(*handle, op, p);
      V b(handle, 30), x(handle, 30);
      b.import_from_host(rhs);
      direct.apply(b, x);
      auto xh = x.export_to_host();
      for (unsigned i = 0; i < 30; ++i)
        CHECK(std::abs(xh[i] - sol_ref[i]) <= 1e-14 * std::abs(sol_ref[i]));
    }
    bool threw = false;
    try
    {
      auto p = std::make_shared<mfmg::ParameterTree>();
      p->put("solver.type", "amgx");
      mfmg::CudaSolver<V> s(*handle, op, p);
    }
    catch (mfmg::NotImplementedExc const &)
    {
      threw = true;
    }
    CHECK(threw);
  }

  // ---- operator apply / transpose (exact) ----
  {
    unsigned const n_rows = 30, nnz_per_row = 10, n_cols = n_rows + nnz_per_row - 1;
    HostCsr a{n_rows, n_cols, {0}, {}, {}};
    for (unsigned i = 0; i < n_rows; ++i)
    {
      for (unsigned j = 0; j < nnz_per_row; ++j)
      {
        a.col.push_back((int)(i + j));
        a.val.push_back((double)(i + i + j));
      }
      a.rp.push_back((int64_t)a.col.size());
    }
    auto m = std::make_shared<mfmg::SparseMatrixDevice<double>>(handle, a.n_rows, a.n_cols, a.rp, a.col, a.val);
    mfmg::CudaMatrixOperator<V> op(m);
    auto dom = op.build_domain_vector();
    auto rng = op.build_range_vector();
    CHECK(dom->size() == n_cols && rng->size() == n_rows);
    *dom = 1.;
    op.apply(*dom, *rng);
    auto ref = host_spmv(a, std::vector<double>(n_cols, 1.));
    auto got = rng->export_to_host();
    for (unsigned i = 0; i < n_rows; ++i)
      CHECK(got[i] == ref[i]);
    auto top = op.transpose();
    auto tdom = top->build_domain_vector();
    auto trng = top->build_range_vector();
    CHECK(tdom->size() == n_rows && trng->size() == n_cols);
    *tdom = 1.;
    top->apply(*tdom, *trng);
    std::vector<double> tref(n_cols, 0.);
    for (unsigned i = 0; i < n_rows; ++i)
      for (int64_t k = a.rp[i]; k < a.rp[i + 1]; ++k)
        tref[a.col[k]] += a.val[k];
    auto tgot = trng->export_to_host();
    for (unsigned i = 0; i < n_cols; ++i)
      CHECK(tgot[i] == tref[i]);
    V t2(handle, n_cols);
    op.apply(*tdom, t2, mfmg::OperatorMode::TRANS);
    auto t2h = t2.export_to_host();
    for (unsigned i = 0; i < n_cols; ++i)
      CHECK(t2h[i] == tref[i]);
  }

  // ---- two-level hierarchy: fused == generic composition ----
  {
    unsigned const n = 255, nc = 127;
    auto a = tridiag(n);
    // linear-interpolation restrictor (full weighting): R[i, 2i..2i+2] = (0.5, 1, 0.5)
    HostCsr r{nc, n, {0}, {}, {}};
    for (unsigned i = 0; i < nc; ++i)
    {
      for (unsigned j = 0; j < 3; ++j)
      {
        r.col.push_back((int)(2 * i + j));
        r.val.push_back(j == 1 ? 1. : 0.5);
      }
      r.rp.push_back((int64_t)r.col.size());
    }
    // A_c = R A R^T (dense on the host, then CSR with the tridiagonal pattern it has)
    std::vector<double> ac(nc * nc, 0.);
    for (unsigned i = 0; i < nc; ++i)
      for (unsigned j = 0; j < nc; ++j)
      {
        double s = 0.;
        for (int64_t ki = r.rp[i]; ki < r.rp[i + 1]; ++ki)
          for (int64_t kj = r.rp[j]; kj < r.rp[j + 1]; ++kj)
          {
            int const p = r.col[ki], q = r.col[kj];
            double apq = p == q ? 4. : (std::abs(p - q) == 1 ? -1. : 0.);
            s += r.val[ki] * apq * r.val[kj];
          }
        ac[i * nc + j] = s;
      }
    HostCsr c{nc, nc, {0}, {}, {}};
    for (unsigned i = 0; i < nc; ++i)
    {
      for (unsigned j = 0; j < nc; ++j)
        if (ac[i * nc + j] != 0.)
        {
          c.col.push_back((int)j);
          c.val.push_back(ac[i * nc + j]);
        }
      c.rp.push_back((int64_t)c.col.size());
    }
    auto mk = [&](HostCsr const &h) {
      return std::make_shared<mfmg::CudaMatrixOperator<V>>(
          std::make_shared<mfmg::SparseMatrixDevice<double>>(handle, h.n_rows, h.n_cols, h.rp, h.col, h.val));
    };
    for (bool precond : {true, false})
      for (unsigned nu : {1u, 2u})
      {
        auto p = std::make_shared<mfmg::ParameterTree>();
        p->put("is preconditioner", precond);
        p->put("smoother.n_smoothing_steps", nu);
        mfmg::Hierarchy<V> hierarchy(handle, {mk(a), mk(c)}, {mk(r)}, p);
        std::vector<double> bh(n), xh(n);
        std::default_random_engine gen(7);
        std::uniform_real_distribution<double> dist(0., 1.);
        for (auto &v : bh)
          v = dist(gen);
        for (auto &v : xh)
          v = dist(gen);
        V b(handle, n), x1(handle, n), x2(handle, n);
        b.import_from_host(bh);
        x1.import_from_host(xh);
        x2.import_from_host(xh);
        hierarchy.vmult(x1, b);
        hierarchy.apply_generic(b, x2);
        auto h1 = x1.export_to_host(), h2 = x2.export_to_host();
        double num = 0., den = 0.;
        for (unsigned i = 0; i < n; ++i)
        {
          num += (h1[i] - h2[i]) * (h1[i] - h2[i]);
          den += h2[i] * h2[i];
        }
        CHECK(std::sqrt(num / den) < 1e-13);
        CHECK(hierarchy.grid_complexity() > 1.0 && hierarchy.operator_complexity() > 1.0);
        // the cycle contracts the error of A x = b: |b - A x_new| < |b - A x_old| in solver mode
        if (!precond)
        {
          auto r0 = host_spmv(a, xh), r1 = host_spmv(a, h1);
          double n0 = 0., n1 = 0.;
          for (unsigned i = 0; i < n; ++i)
          {
            n0 += (r0[i] - bh[i]) * (r0[i] - bh[i]);
            n1 += (r1[i] - bh[i]) * (r1[i] - bh[i]);
          }
          CHECK(n1 < 0.05 * n0);
        }
      }
  }

  // ---- Hierarchy(comm, MeshEvaluator, params, timer): the reference constructor (hierarchy.hpp:159-236) through
  //      create_hierarchy_helpers / CudaHierarchyHelpers with a user CudaMeshEvaluator (matrix-based) ----
  {
    struct TridiagEvaluator : mfmg::CudaMeshEvaluator<2>
    {
      using mfmg::CudaMeshEvaluator<2>::CudaMeshEvaluator;
      void evaluate_global(mfmg::SparseMatrixDevice<double> &m) const override
      {
        auto a = tridiag(255);
        m.reinit(_cuda_handle, a.n_rows, a.n_cols, a.rp, a.col, a.val);
      }
      void build_restrictor_matrix(std::shared_ptr<mfmg::ParameterTree const>,
                                   mfmg::SparseMatrixDevice<double> &m) const override
      {
        HostCsr r{127, 255, {0}, {}, {}};
        for (unsigned i = 0; i < 127; ++i)
        {
          for (unsigned j = 0; j < 3; ++j)
          {
            r.col.push_back((int)(2 * i + j));
            r.val.push_back(j == 1 ? 1. : 0.5);
          }
          r.rp.push_back((int64_t)r.col.size());
        }
        m.reinit(_cuda_handle, r.n_rows, r.n_cols, r.rp, r.col, r.val);
      }
    };
    auto evaluator = std::make_shared<TridiagEvaluator>(handle);
    auto p = std::make_shared<mfmg::ParameterTree>();
    p->put("is preconditioner", false);
    p->put("smoother.n_smoothing_steps", 2u);
    auto timer = std::make_shared<mfmg::TimerOutput>();
    mfmg::Hierarchy<V> hierarchy(mfmg::MPI_COMM_SELF, evaluator, p, timer);
    CHECK(hierarchy.is_fused());
    auto b = hierarchy.build_range_vector();
    CHECK(b->size() == 255);
    std::vector<double> bh(255), xh(255);
    std::default_random_engine gen(11);
    std::uniform_real_distribution<double> dist(0., 1.);
    for (auto &v : bh)
      v = dist(gen);
    for (auto &v : xh)
      v = dist(gen);
    V x1(handle, 255), x2(handle, 255);
    b->import_from_host(bh);
    x1.import_from_host(xh);
    x2.import_from_host(xh);
    hierarchy.vmult(x1, *b);
    hierarchy.apply_generic(*b, x2);
    auto h1 = x1.export_to_host(), h2 = x2.export_to_host();
    double num = 0., den = 0.;
    for (unsigned i = 0; i < 255; ++i)
    {
      num += (h1[i] - h2[i]) * (h1[i] - h2[i]);
      den += h2[i] * h2[i];
    }
    CHECK(std::sqrt(num / den) < 1e-13);
    CHECK(timer->get_summary_data().count("Setup: build coarse operator") == 1);
    CHECK(timer->get_summary_data().count("Apply") == 1);
    // unknown evaluator types are rejected like hierarchy.hpp:102-105
    struct HostEvaluator : mfmg::MeshEvaluator
    {
      int get_dim() const override { return 3; }
      std::string get_mesh_evaluator_type() const override { return "DealIIMeshEvaluator"; }
    };
    bool threw = false;
    try
    {
      mfmg::Hierarchy<V> h2(mfmg::MPI_COMM_SELF, std::make_shared<HostEvaluator>(), p);
    }
    catch (mfmg::NotImplementedExc const &)
    {
      threw = true;
    }
    CHECK(threw);
  }

  // ---- CudaMatrixFreeOperator / CudaMatrixFreeMeshEvaluator (cuda_matrix_free_operator.cuh, *_mesh_evaluator.cuh):
  //      the library's evaluator of the reference's matrix-free Laplace problem, and a user evaluator ----
  {
    int const c = 12; // 12^3 cells, Q1, constant coefficient, Dirichlet on the whole boundary
    int const nn = c + 1;
    unsigned const n = (unsigned)(nn * nn * nn);
    std::vector<double> coef((std::size_t)c * c * c * 8, 1.);
    std::vector<uint8_t> constrained(n, 0);
    for (int k = 0; k < nn; ++k)
      for (int j = 0; j < nn; ++j)
        for (int i = 0; i < nn; ++i)
          if (i == 0 || j == 0 || k == 0 || i == c || j == c || k == c)
            constrained[(std::size_t)(k * nn + j) * nn + i] = 1;
    // piecewise-constant restrictor over 3x3x3-cell agglomerates (rows: agglomerates; unconstrained nodes only)
    struct Evaluator : mfmg::LaplaceMatrixFreeMeshEvaluator<3>
    {
      using mfmg::LaplaceMatrixFreeMeshEvaluator<3>::LaplaceMatrixFreeMeshEvaluator;
      HostCsr r;
      void build_restrictor_matrix(std::shared_ptr<mfmg::ParameterTree const>,
                                   mfmg::SparseMatrixDevice<double> &m) const override
      {
        m.reinit(_cuda_handle, r.n_rows, r.n_cols, r.rp, r.col, r.val);
      }
    };
    double const hh = 1. / c;
    auto evaluator = std::make_shared<Evaluator>(handle, 1, std::vector<int64_t>{c, c, c}, std::vector<double>{hh, hh, hh},
                                                 coef, constrained);
    int const na = c / 3;
    evaluator->r = HostCsr{(unsigned)(na * na * na), n, {0}, {}, {}};
    for (int ak = 0; ak < na; ++ak)
      for (int aj = 0; aj < na; ++aj)
        for (int ai = 0; ai < na; ++ai)
        {
          for (int k = 3 * ak; k < 3 * ak + 3; ++k)
            for (int j = 3 * aj; j < 3 * aj + 3; ++j)
              for (int i = 3 * ai; i < 3 * ai + 3; ++i)
                if (!constrained[(std::size_t)(k * nn + j) * nn + i])
                {
                  evaluator->r.col.push_back((k * nn + j) * nn + i);
                  evaluator->r.val.push_back(1.);
                }
          evaluator->r.rp.push_back((int64_t)evaluator->r.col.size());
        }
    mfmg::CudaMatrixFreeOperator<3, V> op(evaluator);
    auto x = op.build_domain_vector(), y = op.build_range_vector();
    CHECK(x->size() == n && y->size() == n && op.grid_complexity() == n);
    *x = 1.;
    op.apply(*x, *y);
    // A 1 restricted to interior rows: rows whose 27 neighbours are all unconstrained sum to zero (Laplace of a constant)
    auto yh = y->export_to_host();
    CHECK(std::abs(yh[(std::size_t)((6 * nn + 6) * nn + 6)]) < 1e-13);
    CHECK(yh[0] == 1.); // constrained rows act as identity
    bool threw = false;
    try
    {
      op.apply(*x, *y, mfmg::OperatorMode::TRANS);
    }
    catch (mfmg::NotImplementedExc const &)
    {
      threw = true;
    }
    CHECK(threw);
    for (auto smoother : {"Jacobi", "Chebyshev"})
    {
      auto p = std::make_shared<mfmg::ParameterTree>();
      p->put("is preconditioner", false);
      p->put("smoother.type", smoother);
      mfmg::Hierarchy<V> hierarchy(mfmg::MPI_COMM_SELF, evaluator, p);
      CHECK(hierarchy.is_fused()); // the library's own matrix-free operator feeds the fused V-cycle
      std::vector<double> xh(n);
      std::default_random_engine gen(5);
      std::uniform_real_distribution<double> dist(0., 1.);
      for (unsigned i = 0; i < n; ++i)
        xh[i] = constrained[i] ? 0. : dist(gen);
      V b(handle, n), x1(handle, n), r(handle, n);
      b = 0.;
      x1.import_from_host(xh);
      double prev = 0., rate = 0.;
      for (int cycle = 0; cycle < 12; ++cycle)
      {
        hierarchy.vmult(x1, b);
        op.apply(x1, r);
        double const nrm = r.l2_norm();
        if (cycle > 0)
          rate = nrm / prev;
        prev = nrm;
      }
      CHECK(rate > 0. && rate < 0.9); // the two-grid iteration contracts
      if (std::string(smoother) == "Jacobi")
      {
        // fused == the abstract composition (CudaMatrixFreeSmoother + CudaMatrixFreeOperator + CudaSolver)
        V x2(handle, n), x3(handle, n);
        x2.import_from_host(xh);
        x3.import_from_host(xh);
        std::vector<double> bh(n);
        for (unsigned i = 0; i < n; ++i)
          bh[i] = constrained[i] ? 0. : dist(gen);
        b.import_from_host(bh);
        hierarchy.vmult(x2, b);
        hierarchy.apply_generic(b, x3);
        auto h2 = x2.export_to_host(), h3 = x3.export_to_host();
        double num = 0., den = 0.;
        for (unsigned i = 0; i < n; ++i)
        {
          num += (h2[i] - h3[i]) * (h2[i] - h3[i]);
          den += h3[i] * h3[i];
        }
        CHECK(std::sqrt(num / den) < 1e-12);
      }
    }
  }

  if (failures == 0)
    std::printf("test_adapter: ALL OK\n");
  return failures == 0 ? 0 : 1;
}

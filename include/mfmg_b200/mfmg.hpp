// mfmg.hpp -- header-only C++ adapter: mfmg's operator API on top of the mfmg_b200 C ABI.
//
// Same class names, member functions, argument meaning and error behaviour as the reference, so
// code written against mfmg's device path keeps compiling with s/#include <mfmg/cuda/...>/this/:
//
//   mfmg::OperatorMode, Operator<V>      include/mfmg/common/operator.hpp:19-52
//   mfmg::Smoother<V>, Solver<V>         include/mfmg/common/smoother.hpp:23-42, solver.hpp:23-42
//   mfmg::Level<V>                       include/mfmg/common/level.hpp:22-76
//   mfmg::CudaHandle                     include/mfmg/cuda/cuda_handle.cuh
//   mfmg::SparseMatrixDevice<double>     include/mfmg/cuda/sparse_matrix_device.cuh:28-104
//   mfmg::CudaMatrixOperator<V>          include/mfmg/cuda/cuda_matrix_operator.cuh
//   mfmg::CudaSmoother<V>                include/mfmg/cuda/cuda_smoother.cuh
//   mfmg::CudaSolver<V>                  include/mfmg/cuda/cuda_solver.cuh
//   mfmg::Hierarchy<V>::vmult/apply      include/mfmg/common/hierarchy.hpp:238-309
//
// Third-party types of the reference that are not available without deal.II/Boost are replaced
// by minimal stand-ins with the members the path uses:
//   DeviceVector  ~ dealii::LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>
//                   (get_values(), size(), local_size(), operator=(double), add(a, v), l2_norm())
//   HostVector    ~ dealii::LinearAlgebra::distributed::Vector<double, MemorySpace::Host>
//   ParameterTree ~ boost::property_tree::ptree (get<T>(path, default), put(path, value))
// With deal.II present a maintainer binds the real types instead.
#ifndef MFMG_B200_MFMG_HPP
#define MFMG_B200_MFMG_HPP

#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mfmg_b200.h"

namespace mfmg
{
// include/mfmg/common/exceptions.hpp:65-84
class NotImplementedExc : public std::runtime_error
{
public:
  explicit NotImplementedExc(std::string const &what = "not implemented") : std::runtime_error(what) {}
};

inline void ASSERT_THROW(bool cond, std::string const &msg)
{
  if (!cond)
    throw std::runtime_error(msg); // exceptions.hpp:59-63
}

// every C-ABI status is checked in all build types (the reference's ASSERT_CUDA is Debug-only)
inline void check_status(mfmgb_ctx *ctx, int rc)
{
  if (rc == MFMGB_OK)
    return;
  std::string msg = mfmgb_last_error(ctx);
  if (rc == MFMGB_ERR_NOT_IMPLEMENTED)
    throw NotImplementedExc(msg);
  throw std::runtime_error(msg);
}

// include/mfmg/cuda/utils.cuh:66-99 -- the raw device-memory helpers callers use to fill the arrays they hand to
// SparseMatrixDevice's take-ownership constructor (same names and argument order; no CUDA headers needed here)
template <typename T>
inline void cuda_malloc(T *&pointer, unsigned int n_elements)
{
  void *p = nullptr;
  check_status(nullptr, mfmgb_dev_malloc(nullptr, (int64_t)(sizeof(T) * (std::size_t)n_elements), &p));
  pointer = static_cast<T *>(p);
}
template <typename T>
inline void cuda_free(T *&pointer)
{
  check_status(nullptr, mfmgb_dev_free(nullptr, pointer));
  pointer = nullptr;
}
template <typename T>
inline void cuda_mem_copy_to_dev(std::vector<T> const &vector_host, T *pointer_dev)
{
  check_status(nullptr, mfmgb_dev_upload(nullptr, pointer_dev, vector_host.data(), (int64_t)(sizeof(T) * vector_host.size())));
}
template <typename T>
inline void cuda_mem_copy_to_host(T const *pointer_dev, std::vector<T> &vector_host)
{
  check_status(nullptr, mfmgb_dev_download(nullptr, pointer_dev, vector_host.data(), (int64_t)(sizeof(T) * vector_host.size())));
}

enum class OperatorMode
{
  NO_TRANS,
  TRANS
};

// boost::property_tree::ptree stand-in: flat map keyed by the dotted path
class ParameterTree
{
public:
  template <typename T>
  T get(std::string const &path, T const &default_value) const
  {
    auto it = _data.find(path);
    if (it == _data.end())
      return default_value;
    std::istringstream is(it->second);
    T v;
    if (!(is >> std::boolalpha >> v))
      return default_value;
    return v;
  }
  std::string get(std::string const &path, char const *default_value) const
  {
    auto it = _data.find(path);
    return it == _data.end() ? std::string(default_value) : it->second;
  }
  template <typename T>
  void put(std::string const &path, T const &value)
  {
    std::ostringstream os;
    os << std::boolalpha << value;
    _data[path] = os.str();
  }

private:
  std::map<std::string, std::string> _data;
};

// Owns the device context (replaces the cusparse/cusolver handles of mfmg::CudaHandle)
class CudaHandle
{
public:
  explicit CudaHandle(int device = 0, void *stream = nullptr)
  {
    check_status(nullptr, mfmgb_ctx_create(device, stream, &ctx));
  }
  ~CudaHandle() { mfmgb_ctx_destroy(ctx); }
  CudaHandle(CudaHandle const &) = delete;
  CudaHandle &operator=(CudaHandle const &) = delete;
  mfmgb_ctx *ctx = nullptr;
};

class DeviceVector
{
public:
  using value_type = double;
  DeviceVector(std::shared_ptr<CudaHandle const> handle, std::size_t n) : _handle(std::move(handle)), _n(n)
  {
    check_status(_handle->ctx, mfmgb_vec_alloc(_handle->ctx, (int64_t)n, &_values));
  }
  DeviceVector(DeviceVector const &other) : DeviceVector(other._handle, other._n)
  {
    check_status(_handle->ctx, mfmgb_vec_copy(_handle->ctx, _values, other._values, (int64_t)_n));
  }
  DeviceVector &operator=(DeviceVector const &other)
  {
    ASSERT_THROW(_n == other._n, "DeviceVector: size mismatch");
    check_status(_handle->ctx, mfmgb_vec_copy(_handle->ctx, _values, other._values, (int64_t)_n));
    return *this;
  }
  ~DeviceVector() { mfmgb_vec_free(_handle->ctx, _values); }
  DeviceVector &operator=(double s)
  {
    check_status(_handle->ctx, mfmgb_vec_fill(_handle->ctx, _values, s, (int64_t)_n));
    return *this;
  }
  void add(double a, DeviceVector const &v)
  {
    check_status(_handle->ctx, mfmgb_vec_axpy(_handle->ctx, _values, a, v._values, (int64_t)_n));
  }
  double operator*(DeviceVector const &v) const
  {
    double r = 0.;
    check_status(_handle->ctx, mfmgb_vec_dot(_handle->ctx, _values, v._values, (int64_t)_n, &r));
    return r;
  }
  double l2_norm() const { return std::sqrt((*this) * (*this)); }
  double *get_values() { return _values; }
  double const *get_values() const { return _values; }
  std::size_t size() const { return _n; }
  std::size_t local_size() const { return _n; }
  void import_from_host(std::vector<double> const &h)
  {
    ASSERT_THROW(h.size() == _n, "DeviceVector: size mismatch");
    check_status(_handle->ctx, mfmgb_vec_upload(_handle->ctx, _values, h.data(), (int64_t)_n));
  }
  std::vector<double> export_to_host() const
  {
    std::vector<double> h(_n);
    check_status(_handle->ctx, mfmgb_vec_download(_handle->ctx, _values, h.data(), (int64_t)_n));
    return h;
  }
  std::shared_ptr<CudaHandle const> const &handle() const { return _handle; }

private:
  std::shared_ptr<CudaHandle const> _handle;
  std::size_t _n;
  double *_values = nullptr;
};

// include/mfmg/common/operator.hpp:25-52
template <typename VectorType>
class Operator
{
public:
  using vector_type = VectorType;
  using size_type = std::size_t;
  virtual ~Operator() = default;
  virtual void apply(vector_type const &x, vector_type &y, OperatorMode mode = OperatorMode::NO_TRANS) const = 0;
  virtual std::shared_ptr<Operator<vector_type>> transpose() const = 0;
  virtual std::shared_ptr<Operator<vector_type>> multiply(std::shared_ptr<Operator<vector_type> const> b) const = 0;
  virtual std::shared_ptr<Operator<vector_type>>
  multiply_transpose(std::shared_ptr<Operator<vector_type> const> b) const = 0;
  virtual std::shared_ptr<vector_type> build_domain_vector() const = 0;
  virtual std::shared_ptr<vector_type> build_range_vector() const = 0;
  virtual size_type grid_complexity() const = 0;
  virtual size_type operator_complexity() const = 0;
};

// include/mfmg/cuda/sparse_matrix_device.cuh:28-104.  The reference's ctor takes an MPI_Comm and two
// dealii::IndexSet; without deal.II the (local) sizes are passed directly.
template <typename ScalarType>
class SparseMatrixDevice
{
  static_assert(sizeof(ScalarType) == sizeof(double), "FP64 only on this path");

public:
  // TAKES OWNERSHIP of the three cudaMalloc'ed arrays and frees them in the destructor
  // (sparse_matrix_device.templates.cuh:244-272)
  SparseMatrixDevice(std::shared_ptr<CudaHandle const> handle, ScalarType *val_dev_, int *column_index_dev_,
                     int *row_ptr_dev_, unsigned int local_nnz, unsigned int n_rows, unsigned int n_cols)
      : val_dev(val_dev_), column_index_dev(column_index_dev_), row_ptr_dev(row_ptr_dev_), _handle(std::move(handle))
  {
    check_status(_handle->ctx, mfmgb_csr_adopt_device(_handle->ctx, n_rows, n_cols, local_nnz, val_dev,
                                                      column_index_dev, row_ptr_dev, &_csr));
  }
  // convert_matrix semantics (source/cuda/utils.cu:39-168): copy a host CSR to the device
  SparseMatrixDevice(std::shared_ptr<CudaHandle const> handle, unsigned int n_rows, unsigned int n_cols,
                     std::vector<int64_t> const &row_ptr, std::vector<int> const &column_index,
                     std::vector<double> const &val)
      : _handle(std::move(handle))
  {
    check_status(_handle->ctx, mfmgb_csr_upload(_handle->ctx, n_rows, n_cols, row_ptr.data(), column_index.data(),
                                                val.data(), &_csr));
    void *rp = nullptr;
    int is64 = 0;
    mfmgb_csr_device_arrays(_csr, &val_dev, &column_index_dev, &rp, &is64);
    row_ptr_dev = is64 ? nullptr : static_cast<int *>(rp);
  }
  // adopt an existing C handle (used by transpose())
  SparseMatrixDevice(std::shared_ptr<CudaHandle const> handle, mfmgb_csr *csr) : _handle(std::move(handle)), _csr(csr)
  {
    void *rp = nullptr;
    int is64 = 0;
    mfmgb_csr_device_arrays(_csr, &val_dev, &column_index_dev, &rp, &is64);
    row_ptr_dev = is64 ? nullptr : static_cast<int *>(rp);
  }
  ~SparseMatrixDevice() { mfmgb_csr_destroy(_handle->ctx, _csr); }
  SparseMatrixDevice(SparseMatrixDevice const &) = delete;
  SparseMatrixDevice &operator=(SparseMatrixDevice const &) = delete;

  unsigned int m() const { return (unsigned int)info(0); }
  unsigned int n() const { return (unsigned int)info(1); }
  unsigned int n_local_rows() const { return (unsigned int)info(0); }
  unsigned int local_nnz() const { return (unsigned int)info(2); }
  unsigned int n_nonzero_elements() const { return (unsigned int)info(2); }

  void vmult(DeviceVector &dst, DeviceVector const &src) const
  {
    ASSERT_THROW(src.size() == n() && dst.size() == m(), "SparseMatrixDevice::vmult: size mismatch");
    check_status(_handle->ctx, mfmgb_spmv(_handle->ctx, _csr, src.get_values(), dst.get_values()));
  }

  mfmgb_csr *c_handle() const { return _csr; }
  std::shared_ptr<CudaHandle const> const &handle() const { return _handle; }

  ScalarType *val_dev = nullptr;
  int *column_index_dev = nullptr;
  int *row_ptr_dev = nullptr;

private:
  int64_t info(int which) const
  {
    int64_t v[3];
    mfmgb_csr_info(_csr, &v[0], &v[1], &v[2]);
    return v[which];
  }
  std::shared_ptr<CudaHandle const> _handle;
  mfmgb_csr *_csr = nullptr;
};

// source/cuda/cuda_matrix_operator.cu
template <typename VectorType>
class CudaMatrixOperator : public Operator<VectorType>
{
public:
  using vector_type = VectorType;
  using size_type = std::size_t;
  explicit CudaMatrixOperator(std::shared_ptr<SparseMatrixDevice<double>> matrix) : _matrix(std::move(matrix)) {}

  void apply(vector_type const &x, vector_type &y, OperatorMode mode = OperatorMode::NO_TRANS) const override
  {
    if (mode == OperatorMode::NO_TRANS)
      _matrix->vmult(y, x);
    else
    {
      // explicit transpose built lazily on first use (cuda_matrix_operator.cu:84-88); not re-entrant,
      // like the reference's mutable _transposed_matrix
      if (!_transposed_matrix)
        build_transpose();
      _transposed_matrix->vmult(y, x);
    }
  }
  std::shared_ptr<Operator<vector_type>> transpose() const override
  {
    if (!_transposed_matrix)
      build_transpose();
    return std::make_shared<CudaMatrixOperator<vector_type>>(_transposed_matrix);
  }
  // Setup operations: the reference computes products with cuSPARSE csrgemm (serial) or on the host through
  // Trilinos (parallel).  Setup stays on the host path here, so these are not provided by the device library.
  std::shared_ptr<Operator<vector_type>> multiply(std::shared_ptr<Operator<vector_type> const>) const override
  {
    throw NotImplementedExc("CudaMatrixOperator::multiply: setup stays on the host path (mfmg_b200.hostsetup)");
  }
  std::shared_ptr<Operator<vector_type>>
  multiply_transpose(std::shared_ptr<Operator<vector_type> const>) const override
  {
    throw NotImplementedExc("CudaMatrixOperator::multiply_transpose: setup stays on the host path");
  }
  std::shared_ptr<vector_type> build_domain_vector() const override
  {
    return std::make_shared<vector_type>(_matrix->handle(), _matrix->n());
  }
  std::shared_ptr<vector_type> build_range_vector() const override
  {
    return std::make_shared<vector_type>(_matrix->handle(), _matrix->m());
  }
  size_type grid_complexity() const override { return _matrix->m(); }
  size_type operator_complexity() const override { return _matrix->n_nonzero_elements(); }
  std::shared_ptr<SparseMatrixDevice<double>> get_matrix() const { return _matrix; }

private:
  void build_transpose() const
  {
    mfmgb_csr *t = nullptr;
    check_status(_matrix->handle()->ctx, mfmgb_csr_transpose(_matrix->handle()->ctx, _matrix->c_handle(), &t));
    _transposed_matrix = std::make_shared<SparseMatrixDevice<double>>(_matrix->handle(), t);
  }
  std::shared_ptr<SparseMatrixDevice<double>> _matrix;
  mutable std::shared_ptr<SparseMatrixDevice<double>> _transposed_matrix;
};

// include/mfmg/common/smoother.hpp:23-42
template <typename VectorType>
class Smoother
{
public:
  using vector_type = VectorType;
  Smoother(std::shared_ptr<Operator<vector_type> const> op, std::shared_ptr<ParameterTree const> params)
      : _operator(std::move(op)), _params(std::move(params))
  {
  }
  virtual ~Smoother() = default;
  virtual void apply(vector_type const &b, vector_type &x) const = 0;

protected:
  std::shared_ptr<Operator<vector_type> const> _operator;
  std::shared_ptr<ParameterTree const> _params;
};

// source/cuda/cuda_smoother.cu:99-172
template <typename VectorType>
class CudaSmoother : public Smoother<VectorType>
{
public:
  using vector_type = VectorType;
  CudaSmoother(std::shared_ptr<Operator<vector_type> const> op, std::shared_ptr<ParameterTree const> params)
      : Smoother<vector_type>(op, params)
  {
    std::string prec_type = this->_params ? this->_params->get("smoother.type", "Jacobi") : std::string("Jacobi");
    std::transform(prec_type.begin(), prec_type.end(), prec_type.begin(), ::tolower);
    ASSERT_THROW(prec_type == "jacobi", "Only Jacobi smoother is implemented."); // cuda_smoother.cu:110
    _cuda_operator = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(this->_operator);
    ASSERT_THROW(_cuda_operator != nullptr, "CudaSmoother needs a CudaMatrixOperator");
    auto m = _cuda_operator->get_matrix();
    ASSERT_THROW(m->m() == m->n(), "The matrix is not square. The matrix is a " + std::to_string(m->m()) + " by " +
                                       std::to_string(m->n()) + " .");
    check_status(m->handle()->ctx, mfmgb_jacobi_setup(m->handle()->ctx, m->c_handle(), 1.0, &_jacobi));
  }
  ~CudaSmoother() override { mfmgb_jacobi_destroy(_cuda_operator->get_matrix()->handle()->ctx, _jacobi); }
  void apply(vector_type const &b, vector_type &x) const override
  {
    auto m = _cuda_operator->get_matrix();
    check_status(m->handle()->ctx,
                 mfmgb_jacobi_apply(m->handle()->ctx, _jacobi, m->c_handle(), b.get_values(), x.get_values()));
  }

private:
  std::shared_ptr<CudaMatrixOperator<vector_type> const> _cuda_operator;
  mfmgb_jacobi *_jacobi = nullptr;
};

// include/mfmg/common/solver.hpp:23-42
template <typename VectorType>
class Solver
{
public:
  using vector_type = VectorType;
  Solver(std::shared_ptr<Operator<vector_type> const> op, std::shared_ptr<ParameterTree const> params)
      : _operator(std::move(op)), _params(std::move(params))
  {
  }
  virtual ~Solver() = default;
  virtual void apply(vector_type const &b, vector_type &x) const = 0;

protected:
  std::shared_ptr<Operator<vector_type> const> _operator;
  std::shared_ptr<ParameterTree const> _params;
};

// source/cuda/cuda_solver.cu:196-515 ("lu_dense" is the default, :215)
template <typename VectorType>
class CudaSolver : public Solver<VectorType>
{
public:
  using vector_type = VectorType;
  CudaSolver(CudaHandle const &cuda_handle, std::shared_ptr<Operator<vector_type> const> op,
             std::shared_ptr<ParameterTree const> params)
      : Solver<vector_type>(op, params), _ctx(cuda_handle.ctx)
  {
    std::string solver = this->_params ? this->_params->get("solver.type", "lu_dense") : std::string("lu_dense");
    if (solver == "amgx")
      throw NotImplementedExc("solver.type amgx is not available in mfmg_b200");
    ASSERT_THROW(solver == "lu_dense" || solver == "cholesky" || solver == "lu_sparse_host",
                 "The provided solver name " + solver + " is invalid."); // cuda_solver.cu:70
    auto cuda_operator = std::dynamic_pointer_cast<CudaMatrixOperator<vector_type> const>(this->_operator);
    ASSERT_THROW(cuda_operator != nullptr, "CudaSolver needs a CudaMatrixOperator");
    check_status(_ctx, mfmgb_dense_factor(_ctx, cuda_operator->get_matrix()->c_handle(), &_dense));
  }
  ~CudaSolver() override { mfmgb_dense_destroy(_ctx, _dense); }
  void apply(vector_type const &b, vector_type &x) const override
  {
    check_status(_ctx, mfmgb_dense_solve(_ctx, _dense, b.get_values(), x.get_values()));
  }

private:
  mfmgb_ctx *_ctx;
  mfmgb_dense *_dense = nullptr;
};

// include/mfmg/common/level.hpp:22-76
template <typename VectorType>
class Level
{
public:
  using vector_type = VectorType;
  std::shared_ptr<Operator<vector_type> const> get_operator() const { return _operator; }
  std::shared_ptr<Operator<vector_type> const> get_restrictor() const { return _restrictor; }
  std::shared_ptr<Smoother<vector_type> const> get_smoother() const { return _smoother; }
  std::shared_ptr<Solver<vector_type> const> get_solver() const { return _solver; }
  void set_operator(std::shared_ptr<Operator<vector_type> const> op) { _operator = op; }
  void set_restrictor(std::shared_ptr<Operator<vector_type> const> r) { _restrictor = r; }
  void set_smoother(std::shared_ptr<Smoother<vector_type> const> s) { _smoother = s; }
  void set_solver(std::shared_ptr<Solver<vector_type> const> s) { _solver = s; }
  std::shared_ptr<vector_type> build_vector() const { return _operator->build_range_vector(); }

private:
  std::shared_ptr<Operator<vector_type> const> _operator, _restrictor;
  std::shared_ptr<Smoother<vector_type> const> _smoother;
  std::shared_ptr<Solver<vector_type> const> _solver;
};

// include/mfmg/common/hierarchy.hpp:159-309.  The reference ctor runs the SETUP from a MeshEvaluator;
// setup stays on the host path, so this ctor receives the level operators it produced (A_l, R_l) and
// builds smoothers / coarse solver exactly as hierarchy.hpp:183-234 does.
template <typename VectorType>
class Hierarchy
{
public:
  using vector_type = VectorType;
  Hierarchy(std::shared_ptr<CudaHandle const> handle,
            std::vector<std::shared_ptr<CudaMatrixOperator<vector_type>>> const &operators,
            std::vector<std::shared_ptr<CudaMatrixOperator<vector_type>>> const &restrictors,
            std::shared_ptr<ParameterTree> params = nullptr)
      : _handle(std::move(handle))
  {
    if (!params)
      params = std::make_shared<ParameterTree>(); // (the reference dereferences a null default, hierarchy.hpp:160,168)
    _is_preconditioner = params->get("is preconditioner", true);             // :168
    _n_smoothing_steps = params->get("smoother.n_smoothing_steps", 1u);      // :169
    unsigned int const num_levels = (unsigned int)operators.size();
    ASSERT_THROW(num_levels >= 1 && restrictors.size() + 1 == num_levels, "Hierarchy: inconsistent level description");
    _levels.resize(num_levels);
    for (unsigned int l = 0; l < num_levels; ++l)
    {
      _levels[l].set_operator(operators[l]);
      if (l > 0)
        _levels[l].set_restrictor(restrictors[l - 1]);
      if (l + 1 < num_levels)
        _levels[l].set_smoother(std::make_shared<CudaSmoother<vector_type>>(operators[l], params)); // :204
      else
        _levels[l].set_solver(std::make_shared<CudaSolver<vector_type>>(*_handle, operators[l], params)); // :194
    }
    // fused device path
    check_status(_handle->ctx, mfmgb_hierarchy_create(_handle->ctx, (int)num_levels, (int)_n_smoothing_steps,
                                                      _is_preconditioner ? 1 : 0, 1.0, &_fused));
    for (unsigned int l = 0; l < num_levels; ++l)
    {
      check_status(_handle->ctx, mfmgb_hierarchy_set_operator(_fused, (int)l, operators[l]->get_matrix()->c_handle()));
      if (l > 0)
        check_status(_handle->ctx, mfmgb_hierarchy_set_restrictor(_fused, (int)l,
                                                                  restrictors[l - 1]->get_matrix()->c_handle(), nullptr));
    }
    check_status(_handle->ctx, mfmgb_hierarchy_finalize(_handle->ctx, _fused));
  }
  ~Hierarchy() { mfmgb_hierarchy_destroy(_handle->ctx, _fused); }
  Hierarchy(Hierarchy const &) = delete;

  // hierarchy.hpp:238-244
  void vmult(vector_type &x, vector_type const &b) const { apply(b, x, 0); }

  // hierarchy.hpp:246-309 -- fused kernels, preallocated workspaces, one launch sequence
  void apply(vector_type const &b, vector_type &x, int level_index = 0) const
  {
    check_status(_handle->ctx, mfmgb_hierarchy_apply(_handle->ctx, _fused, b.get_values(), x.get_values(), level_index));
  }

  // The same algorithm written against the abstract Operator / Smoother / Solver interfaces, line for line as
  // the reference composes it (unfused C-ABI calls).  Kept to show the drop-in objects are interchangeable.
  void apply_generic(vector_type const &b, vector_type &x, int level_index = 0) const
  {
    auto const num_levels = _levels.size();
    auto &level_fine = _levels[level_index];
    auto a = level_fine.get_operator();
    if (level_index > 0 || _is_preconditioner)
      x = 0.;
    if (level_index == (int)num_levels - 1)
    {
      level_fine.get_solver()->apply(b, x);
      return;
    }
    auto &level_coarse = _levels[level_index + 1];
    auto restrictor = level_coarse.get_restrictor();
    auto smoother = level_fine.get_smoother();
    for (unsigned int i = 0; i < _n_smoothing_steps; ++i)
      smoother->apply(b, x);
    auto res = level_fine.build_vector();
    a->apply(x, *res);
    res->add(-1., b);
    auto b_coarse = level_coarse.build_vector();
    restrictor->apply(*res, *b_coarse);
    auto x_coarse = level_coarse.build_vector();
    apply_generic(*b_coarse, *x_coarse, level_index + 1);
    auto x_correction = level_fine.build_vector();
    restrictor->apply(*x_coarse, *x_correction, OperatorMode::TRANS);
    x.add(-1., *x_correction);
    for (unsigned int i = 0; i < _n_smoothing_steps; ++i)
      smoother->apply(b, x);
  }

  double grid_complexity() const
  {
    double s = 0;
    for (auto const &l : _levels)
      s += (double)l.get_operator()->grid_complexity();
    return s / (double)_levels[0].get_operator()->grid_complexity();
  }
  double operator_complexity() const
  {
    double s = 0;
    for (auto const &l : _levels)
      s += (double)l.get_operator()->operator_complexity();
    return s / (double)_levels[0].get_operator()->operator_complexity();
  }
  mfmgb_hierarchy *c_handle() const { return _fused; }

private:
  std::shared_ptr<CudaHandle const> _handle;
  bool _is_preconditioner = true;
  unsigned int _n_smoothing_steps = 1;
  std::vector<Level<vector_type>> _levels;
  mfmgb_hierarchy *_fused = nullptr;
};

// names used by older mfmg snapshots / BASELINE.json
template <typename V>
using SparseMatrixDeviceOperator = CudaMatrixOperator<V>;
template <typename V>
using SmootherDevice = CudaSmoother<V>;
template <typename V>
using DirectSolverDevice = CudaSolver<V>;
using HierarchyDevice = Hierarchy<DeviceVector>;
} // namespace mfmg

#endif
